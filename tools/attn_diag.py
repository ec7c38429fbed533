import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam6d_b200 import ops
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
B, S, H, D = 64, 197, 4, 64
qkv = torch.randn(B * S, 3 * H * D, device="cuda").bfloat16()
vt = ops.transpose_tokens(qkv, 2 * H * D, H * D, B, S)
bias = torch.randn(B, H, S, S, device="cuda")
ref = ops.attn_tc(qkv, 0, qkv, H * D, vt, B, H, S, S, D, 0.125, bias=bias, bias_variant=1)
alt = ops.attn_tc(qkv, 0, qkv, H * D, vt, B, H, S, S, D, 0.125, bias=bias, bias_variant=3)
print("variants agree:", (ref - alt).abs().max().item())
print("no bias        %8.1f us" % timeit(lambda: ops.attn_tc(qkv, 0, qkv, H * D, vt, B, H, S, S, D, 0.125)))
print("bias, tmem wb  %8.1f us" % timeit(lambda: ops.attn_tc(qkv, 0, qkv, H * D, vt, B, H, S, S, D, 0.125, bias=bias, bias_variant=1)))
print("bias, restage  %8.1f us" % timeit(lambda: ops.attn_tc(qkv, 0, qkv, H * D, vt, B, H, S, S, D, 0.125, bias=bias, bias_variant=3)))
# SAM window shape
nW, L, H2, D2 = 800, 196, 16, 80
qkv2 = torch.randn(nW * L, 3 * H2 * D2, device="cuda").bfloat16()
vt2 = ops.transpose_tokens(qkv2, 2 * H2 * D2, H2 * D2, nW, L)
rh = torch.randn(27, 80, device="cuda") * 0.1; rw = torch.randn(27, 80, device="cuda") * 0.1
blob = ops.pack_rel_pos(rh, rw)
print("sam win no bias %8.1f us" % timeit(lambda: ops.attn_tc(qkv2, 0, qkv2, H2 * D2, vt2, nW, H2, L, L, D2, 0.11, out_dtype=torch.bfloat16), n=3))
print("sam win rel-pos %8.1f us" % timeit(lambda: ops.attn_tc(qkv2, 0, qkv2, H2 * D2, vt2, nW, H2, L, L, D2, 0.11, rel=(blob, 14, 14), out_dtype=torch.bfloat16), n=3))
qf = qkv2.float()
print("sam win simt    %8.1f us" % timeit(lambda: ops.attn_relpos(qf, nW, 14, 14, H2, rh, rw, 0.11, out_dtype=torch.bfloat16), n=3))
# SAM global attention shape: 16 images x 16 heads x 4096 tokens
Bg, Hg, Sg, Dg = 16, 16, 64, 80
qkvg = torch.randn(Bg * Sg * Sg, 3 * Hg * Dg, device="cuda").bfloat16()
vtg = ops.transpose_tokens(qkvg, 2 * Hg * Dg, Hg * Dg, Bg, Sg * Sg)
rhg = torch.randn(127, 80, device="cuda") * 0.1; rwg = torch.randn(127, 80, device="cuda") * 0.1
blobg = ops.pack_rel_pos(rhg, rwg, slab_rows=128)
t = timeit(lambda: ops.attn_global_tc(qkvg, vtg, blobg, Bg, Hg, Sg, 0.11), n=3)
print("sam global tc   %8.1f us  (%.1f TFLOP/s)" % (t, 4.0 * Bg * Hg * 4096 * 4096 * 80 / t / 1e6))
