"""scratch diagnostics run on the GPU box (not a test)"""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pem_oracle as po
from sam6d_b200 import ops
from sam6d_b200.pem import Net

def G(s): return torch.Generator().manual_seed(s)
d = torch.randn(2, 196, 3, generator=G(3))
pts = d / d.norm(dim=2, keepdim=True) * (0.5 + 0.5 * torch.rand(2, 196, 1, generator=G(4)))
pts = torch.cat([torch.ones(2, 1, 3) * 100, pts], dim=1).contiguous()
d_ref, a_ref = po.geo_embedding_indices(pts)
true = torch.cdist(pts.double(), pts.double()) / 0.2
T = ops.geo_indices(pts.cuda(), 0.2, 180.0 / (15 * math.pi)).cpu()
print("cpu-oracle vs true: max", (d_ref.double() - true).abs().max().item())
print("gpu vs true: max", (T[..., 3].double() - true).abs().max().item())
e = (T[..., 3].double() - true).abs()
i = e.argmax().item(); print("worst gpu idx", i // (197 * 197), (i // 197) % 197, i % 197, "true", true.flatten()[i].item(), "gpu", T[..., 3].flatten()[i].item(), "cpu", d_ref.flatten()[i].item())
e2 = (d_ref.double() - true).abs()
i = e2.argmax().item(); print("worst cpu idx", i // (197 * 197), (i // 197) % 197, i % 197, "true", true.flatten()[i].item(), "cpu", d_ref.flatten()[i].item())
print("torch threads", torch.get_num_threads(), "mkldnn", torch.backends.mkldnn.is_available(), "fp32 matmul precision", torch.get_float32_matmul_precision())
x = pts[0]
print("cpu matmul check:", (x @ x.t() - (x.double() @ x.double().t()).float()).abs().max().item())

# end-to-end golden full: per-proposal diffs
gold = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/pem_full.pt"), weights_only=False)
m = gold["meta"]
sd = po.make_state_dict(seed=m["seed"])
net = Net().cuda().eval(); net.load_state_dict(sd, strict=True)
inp = po.make_inputs(B=m["B"], n=m["n"], seed=m["seed"])
torch.manual_seed(1); rand = torch.rand(m["B"], 18000)
out = net({k: inp[k].cuda() for k in ("pts", "dense_fm", "dense_po", "dense_fo", "model")}, rand=rand.cuda())
for k in ("init_R", "init_t", "pred_R", "pred_t", "pred_pose_score"):
    dd = (out[k].cpu() - gold[k]).abs().reshape(m["B"], -1).amax(dim=1)
    print(k, dd.tolist())
