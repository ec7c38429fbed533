"""Host side of the table-interpolated geometric embedding (csrc/geo_lut.cu): the tables GeometricStructureEmbedding._tables builds
(float64 evaluation of W emb(x) on a uniform grid, stored bf16) interpolate the projected sinusoidal embedding to within the bf16
rounding of the result -- checked against the float64 embedding with the kernel's arithmetic emulated in torch, next to the
bf16-operand tensor-core product they replace (tools/geo_lut_error.py prints the same study)."""
import math

import torch

from oracle import pem_oracle as po
from sam6d_b200 import pem


def _bf(t):
    return t.float().to(torch.bfloat16).double()


def test_tables_interpolate_the_projected_embedding():
    sd = po.make_state_dict(seed=3)
    geo = pem.GeometricStructureEmbedding(pem.DEFAULT_MODEL_CFG["geo_embedding"])
    geo.load_state_dict({k[len("geo_embedding."):]: v for k, v in sd.items() if k.startswith("geo_embedding.")})
    w = geo._weights()
    inv_h = pem.GEO_LUT_INV_H
    assert w["tab_a"].dtype == torch.bfloat16 and w["tab_d"].dtype == torch.bfloat16
    assert w["tab_a"].shape == (int(math.ceil(180.0 / geo.sigma_a * inv_h)) + 1, 256)         # angle indices: [0, 180 / sigma_a]
    assert w["tab_d"].shape == (int(pem.GEO_LUT_D_MAX * inv_h) + 1, 256)
    assert (w["tab_a"].shape[0] + w["tab_d"].shape[0]) * 512 <= 200 * 1024                    # both tables fit one CTA's shared memory
    assert torch.equal(w["wdT_bf"], geo.proj_d.weight.detach().t().contiguous().to(torch.bfloat16))
    Wa, Wd = geo.proj_a.weight.detach().double(), geo.proj_d.weight.detach().double()
    bias = (geo.proj_a.bias + geo.proj_d.bias).detach().double()
    div = geo.embedding.div_term.double()

    def emb(x):
        om = x[..., None].double() * div
        return torch.stack([torch.sin(om), torch.cos(om)], dim=-1).reshape(*x.shape, 256)

    # table rows are the exact function values (to bf16 rounding), the distance table carries both biases
    i = torch.tensor([0, 1, 17, w["tab_d"].shape[0] - 1])
    torch.testing.assert_close(w["tab_d"][i].double(), _bf(emb(i.double() / inv_h) @ Wd.T + bias), atol=0, rtol=0)
    torch.testing.assert_close(w["tab_a"][i[:3]].double(), _bf(emb(i[:3].double() / inv_h) @ Wa.T), atol=0, rtol=0)

    g = torch.Generator().manual_seed(0)
    n = 4000
    xa, xd = torch.rand(n, 3, generator=g) * 12.0, torch.rand(n, generator=g) * (pem.GEO_LUT_D_MAX - 1e-3)
    exact = emb(xd) @ Wd.T + bias + (emb(xa) @ Wa.T).max(dim=1).values

    def lerp(tab, x):                                   # the kernel's fp32 interpolation of the bf16 table
        tab = tab.double()
        u = x.float() * inv_h
        k = u.floor().clamp(0, tab.shape[0] - 2).long()
        t = (u - k.float()).double()
        return (tab[k] + t[..., None] * (tab[k + 1] - tab[k])).float().double()

    e_lut = _bf(lerp(w["tab_d"], xd) + lerp(w["tab_a"], xa.reshape(-1)).reshape(n, 3, 256).max(dim=1).values)
    e_tc = _bf(_bf((_bf(emb(xd).float()) @ _bf(Wd).T).float().double() + bias.float().double()) +
               _bf((_bf(emb(xa.reshape(-1)).float()) @ _bf(Wa).T).float().double().reshape(n, 3, 256)).max(dim=1).values)
    rms = lambda e: (e - exact).pow(2).mean().sqrt().item()          # noqa: E731
    floor = rms(_bf(exact))                                           # bf16 rounding of the exact result alone
    assert rms(e_lut) < rms(e_tc), (rms(e_lut), rms(e_tc))            # closer to the float64 embedding than the tensor-core product
    assert rms(e_lut) < 1.5 * floor, (rms(e_lut), floor)
    assert (e_lut - exact).abs().max().item() < 2e-2 * exact.abs().max().item()
