"""per-call device times of the named C-ABI entry points inside a config #2 forward"""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from sam6d_b200 import _lib, ops, synth
from sam6d_b200.pem import Net
torch.manual_seed(0)
net = Net(precision="bf16").cuda().eval()
net.load_state_dict(synth.make_pem_state_dict(seed=1), strict=True)
inp = {k: v.cuda() for k, v in synth.make_pem_inputs(B=32, n=2048, n_model=1024, seed=3).items() if k in ("pts", "dense_fm", "dense_po", "dense_fo", "model")}
rand = torch.rand(32, synth.N_PROPOSAL1 * 3, device='cuda')
for _ in range(3): net(inp, rand=rand)
for nm in ("sam6d_pe_mlp_max_tc", "sam6d_ball_query_pair", "sam6d_coarse_select", "sam6d_transformer_tail_bf16", "sam6d_geo_embed_tc"):
    _lib.time_kernel(nm, True)
for _ in range(5): net(inp, rand=rand)
torch.cuda.synchronize()
for nm in ("sam6d_pe_mlp_max_tc", "sam6d_ball_query_pair", "sam6d_coarse_select", "sam6d_transformer_tail_bf16", "sam6d_geo_embed_tc"):
    ev = [a.elapsed_time(b) for a, b in _lib.timed_events(nm)]
    per = len(ev) // 5
    print(nm, 'calls/step', per, 'ms/step', sum(ev) / 5, 'per call', [round(sum(ev[i::per]) / 5, 4) for i in range(per)] if per <= 4 else '', flush=True)
