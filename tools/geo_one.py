"""one geometric-embedding call at config #2 size (for ncu captures)"""
import sys, torch
sys.path.insert(0, '/root/repo')
from sam6d_b200 import ops
torch.manual_seed(0)
B, S = 64, 197
T = torch.rand(B, S, S, 4, device='cuda') * 3
div = torch.exp(torch.arange(0, 256, 2, device='cuda').float() * (-9.21 / 256))
Wa = (torch.randn(256, 256, device='cuda') * 0.05).bfloat16(); Wd = (torch.randn(256, 256, device='cuda') * 0.05).bfloat16()
bias = torch.zeros(256, device='cuda')
for _ in range(2):
    ops.geo_embed_tc(T, div, Wa, Wd, bias)
torch.cuda.synchronize()
