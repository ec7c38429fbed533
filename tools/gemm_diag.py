import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sam6d_b200 import ops
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(65536, 3840, 1280), (65536, 256, 256)]:
    Ab = torch.randn(M, K, device="cuda").bfloat16(); Wb = torch.randn(N, K, device="cuda").bfloat16(); b = torch.randn(N, device="cuda")
    for act, name in ((0, "full epilogue"), (-2, "math, no stores"), (-1, "tmem read only")):
        for odt in (torch.bfloat16, torch.float32):
            t = timeit(lambda: ops.gemm_tma(Ab, Wb, b, act=act, out_dtype=odt))
            print(f"M={M} N={N} K={K} {name:18s} out={str(odt)[6:]:9s} {t*1e3:8.1f} us ({2.0*M*N*K/t/1e9:7.1f} TF)")
    t = timeit(lambda: ops.gemm_tma(Ab, Wb, None, act=0, out_dtype=torch.bfloat16))
    print(f"   no bias: {t*1e3:8.1f} us")
