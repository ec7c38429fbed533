"""GPU parity of the PEM RGB branch (sam6d_b200/vit.py) against oracle/vit_oracle.py: the bilinear pixel gather,
a small ViT (width 192, depth 4) in both arithmetic modes, ViT-B/16 at full size, state_dict compatibility, and Net end to end
from rgb + rgb_choose.  The timm trunk of the reference is unpinned (oracle header); tolerances are written below."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import vit_oracle as vo      # noqa: E402
from oracle import pem_oracle as po      # noqa: E402


def G(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_bilinear_gather_kernel(dt):
    from sam6d_b200 import ops
    B, C, Gr, sub, H = 3, 256, 14, 4, 224
    up = torch.randn(B, Gr * Gr, sub * sub * C, generator=G(1)).to(dt)
    choose = torch.randint(0, H * H, (B, 2048), generator=G(2))
    choose[0, :4] = torch.tensor([0, H - 1, H * (H - 1), H * H - 1])
    fmap = up.float().reshape(B, Gr, Gr, sub, sub, C).permute(0, 5, 1, 3, 2, 4).contiguous().reshape(B, C, Gr * sub, Gr * sub)
    want = vo.chosen_pixel_feats(F.interpolate(fmap, (H, H), mode="bilinear", align_corners=False), choose)
    got = ops.bilinear_gather(up.cuda(), choose.cuda(), Gr, sub, C, H, H).cpu()
    torch.testing.assert_close(got, want, atol=2e-6, rtol=1e-5)


def _encoder(embed_dim, depth, heads, precision):
    from sam6d_b200.vit import ViTEncoder
    cfg = SimpleNamespace(vit_type="vit_base", up_type="linear", embed_dim=embed_dim, out_dim=256, use_pyramid_feat=True,
                          pretrained=False, depth=depth, num_heads=heads)
    return ViTEncoder(cfg, npoint=2048, precision=precision).cuda().eval()


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 4e-2)])
def test_small_vit_encoder_matches_oracle(precision, tol):
    sd = vo.make_state_dict(embed_dim=192, depth=4, seed=3)
    enc = _encoder(192, 4, 3, precision)
    missing = enc.load_state_dict(sd, strict=True)              # same key set as the (timm-named) reference checkpoint
    assert not missing.missing_keys and not missing.unexpected_keys
    img = torch.randn(3, 3, 224, 224, generator=G(4))
    choose = torch.randint(0, 224 * 224, (3, 2048), generator=G(5))
    ref = vo.get_img_feats(sd, img, choose, depth=4, num_heads=3)
    got = enc.get_img_feats(img.cuda(), choose.cuda()).cpu()
    assert got.shape == (3, 2048, 256)
    scale = ref.abs().mean().item()
    err = (got - ref).abs()
    print(f"vit small {precision}: mean |ref| {scale:.3f}, max err {err.max().item():.2e}, mean err {err.mean().item():.2e}")
    torch.testing.assert_close(got, ref, atol=tol * max(scale, 1.0), rtol=tol)


def test_vit_base_full_size_bf16_and_full_map_contract():
    sd = vo.make_state_dict(embed_dim=768, depth=12, seed=6)
    enc = _encoder(768, 12, 12, "bf16")
    enc.load_state_dict(sd, strict=True)
    img = torch.randn(2, 3, 224, 224, generator=G(7))
    choose = torch.randint(0, 224 * 224, (2, 2048), generator=G(8))
    outs = vo.vit_forward(sd, img, 12, 12)
    fmap, cls_ref = vo.upscale_map(sd, outs, 224, 224, 256)
    ref = vo.chosen_pixel_feats(fmap, choose)
    got = enc.get_img_feats(img.cuda(), choose.cuda()).cpu()
    scale = ref.abs().mean().item()
    err = (got - ref).abs()
    print(f"vit-b bf16: mean |ref| {scale:.3f}, max err {err.max().item():.2e}, mean err {err.mean().item():.2e}")
    assert err.mean().item() < 1.5e-2 * max(scale, 1.0)         # bf16 operands through 12 blocks and a K = 3072 projection
    torch.testing.assert_close(got, ref, atol=0.12 * max(scale, 1.0), rtol=0.1)
    full, cls = enc.rgb_net(img.cuda())                          # the reference's ViT_AE.forward contract: (B,256,224,224), cls
    assert full.shape == (2, 256, 224, 224) and cls.shape == (2, 768)
    torch.testing.assert_close(vo.chosen_pixel_feats(full.cpu(), choose), got, atol=1e-6, rtol=0)
    torch.testing.assert_close(cls.cpu(), cls_ref, atol=0.1, rtol=0.1)


def test_net_from_rgb_equals_net_from_features():
    """Net with the RGB branch plugged in (pose_estimation_model.py:23-33): same poses as feeding its features directly"""
    from sam6d_b200.pem import Net
    enc = _encoder(192, 4, 3, "bf16")
    enc.load_state_dict(vo.make_state_dict(embed_dim=192, depth=4, seed=3), strict=True)
    net = Net(feature_extraction=enc, precision="bf16").cuda().eval()
    net.load_state_dict({**po.make_state_dict(seed=1), **{"feature_extraction." + k: v for k, v in enc.state_dict().items()}}, strict=True)
    B = 2
    inp = po.make_inputs(B=B, n=2048, seed=12)
    rgb = torch.randn(B, 3, 224, 224, generator=G(9)).cuda()
    choose = torch.randint(0, 224 * 224, (B, 2048), generator=G(10)).cuda()
    torch.manual_seed(1)
    rand = torch.rand(B, po.N_PROPOSAL1 * 3).cuda()
    base = {k: inp[k].cuda() for k in ("pts", "dense_po", "dense_fo", "model")}
    a = net(dict(base, rgb=rgb, rgb_choose=choose), rand=rand)
    b = net(dict(base, dense_fm=enc.get_img_feats(rgb, choose)), rand=rand)
    assert torch.equal(a["pred_R"], b["pred_R"]) and torch.equal(a["pred_t"], b["pred_t"])
    R = a["pred_R"].cpu()
    torch.testing.assert_close(R @ R.transpose(1, 2), torch.eye(3).expand_as(R), atol=1e-5, rtol=0)
