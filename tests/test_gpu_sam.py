"""GPU parity of the SAM ViT image-encoder path: the attention kernel against the oracle restatement of
Attention.forward / add_decomposed_rel_pos, and the drop-in ImageEncoderViT against the golden fixture produced by the
vendored reference module (tests/golden/sam_small.pt: ViT-H width, one windowed + one global block, 1024^2 input)."""
import os
from functools import partial

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sam_oracle as so      # noqa: E402


def _ref_attention(qkv, nW, Hs, Ws, nH, rel_h, rel_w, scale):
    T, C3 = qkv.shape
    C = C3 // 3
    hd = C // nH
    x = qkv.view(nW, Hs * Ws, 3, nH, hd).permute(2, 0, 3, 1, 4).reshape(3, nW * nH, Hs * Ws, hd)
    q, k, v = x.unbind(0)
    attn = (q * scale) @ k.transpose(-2, -1)
    Rh, Rw = so.rel_pos_table(Hs, rel_h), so.rel_pos_table(Ws, rel_w)
    rq = q.reshape(nW * nH, Hs, Ws, hd)
    attn = (attn.view(-1, Hs, Ws, Hs, Ws) + torch.einsum("bhwc,hkc->bhwk", rq, Rh)[:, :, :, :, None] +
            torch.einsum("bhwc,wkc->bhwk", rq, Rw)[:, :, :, None, :]).view(-1, Hs * Ws, Hs * Ws).softmax(dim=-1)
    return (attn @ v).view(nW, nH, Hs * Ws, hd).permute(0, 2, 1, 3).reshape(T, C)


@pytest.mark.parametrize("nW,S,nH", [(3, 14, 4), (1, 64, 2), (2, 9, 1)])
def test_attn_relpos_kernel(nW, S, nH):
    from sam6d_b200 import ops
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(nW * S * S, 3 * nH * 80, generator=g)
    rel_h = torch.randn(2 * S - 1, 80, generator=g) * 0.1
    rel_w = torch.randn(2 * S - 1, 80, generator=g) * 0.1
    ref = _ref_attention(qkv, nW, S, S, nH, rel_h, rel_w, 80 ** -0.5)
    got = ops.attn_relpos(qkv.cuda(), nW, S, S, nH, rel_h.cuda(), rel_w.cuda(), 80 ** -0.5).cpu()
    torch.testing.assert_close(got, ref, atol=2e-5, rtol=1e-4)


def _encoder(cfg, precision):
    from sam6d_b200.sam import ImageEncoderViT
    return ImageEncoderViT(depth=cfg["depth"], embed_dim=cfg["embed_dim"], img_size=1024, mlp_ratio=4,
                           norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=cfg["num_heads"], patch_size=16, qkv_bias=True,
                           use_rel_pos=True, global_attn_indexes=cfg["global_attn_indexes"], window_size=14, out_chans=256,
                           precision=precision).cuda().eval()


@pytest.mark.parametrize("precision,atol", [("fp32", 2e-3), ("bf16", 6e-2)])
def test_image_encoder_matches_reference_golden(golden_dir, precision, atol):
    gold = torch.load(os.path.join(golden_dir, "sam_small.pt"), weights_only=False)
    cfg = gold["meta"]["cfg"]
    enc = _encoder(cfg, precision)
    enc.load_state_dict(so.make_state_dict(seed=gold["meta"]["seed"], **cfg), strict=True)
    img = so.make_images(B=1, seed=gold["meta"]["img_seed"])
    out = enc(img.cuda()).cpu()
    assert out.shape == (1, 256, 64, 64)
    err = (out[:, :, ::4, ::4] - gold["out_sub"]).abs()
    print(f"SAM encoder {precision}: max err {err.max().item():.3e}, mean err {err.mean().item():.3e}, |ref| mean {gold['out_abs_mean']:.3f}")
    assert err.max().item() < atol
    assert abs(out.double().sum().item() - gold["out_sum"]) < atol * out.numel() * 0.05


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_vit_h_32_blocks_matches_reference_golden(golden_dir, precision):
    """the full SAM ViT-H of build_sam.py:14-21 (32 blocks, 4 global) on one 1024 x 1024 frame against the vendored reference
    module's CPU output (tests/golden/sam_vith.pt, tools/make_golden.py sam_vith).  Seeded random weights are a worst case for
    error growth (no trained-in contraction): the bounds below are the measured drift plus margin, stated relative to the
    output's standard deviation."""
    gold = torch.load(os.path.join(golden_dir, "sam_vith.pt"), weights_only=False)
    cfg = gold["meta"]["cfg"]
    assert cfg["depth"] == 32
    enc = _encoder(cfg, precision)
    enc.load_state_dict(so.make_state_dict(seed=gold["meta"]["seed"], **cfg), strict=True)
    img = so.make_images(B=1, seed=gold["meta"]["img_seed"])
    out = enc(img.cuda()).cpu()
    ref = gold["out_sub"]
    err = (out[:, :, ::4, ::4] - ref).abs()
    rel_rms = (err.pow(2).mean().sqrt() / gold["out_std"]).item()
    cos = torch.nn.functional.cosine_similarity(out[:, :, ::4, ::4].flatten(), ref.flatten(), dim=0).item()
    print(f"SAM ViT-H 32 blocks {precision}: max err {err.max().item():.3e}, rms err / output std {rel_rms:.3e}, cosine {cos:.6f}, "
          f"|ref| max {gold['out_abs_max']:.2f} std {gold['out_std']:.3f}")
    assert torch.isfinite(out).all()
    if precision == "fp32":
        assert rel_rms < 2e-3 and cos > 0.99999
    else:
        assert rel_rms < 0.12 and cos > 0.993


def test_image_encoder_batch_and_independence():
    """two frames in one batch give the same embeddings as one at a time (no cross-image leakage through the window maps)"""
    cfg = dict(embed_dim=1280, depth=2, num_heads=16, global_attn_indexes=(1,))
    enc = _encoder(cfg, "bf16")
    enc.load_state_dict(so.make_state_dict(seed=2, **cfg), strict=True)
    img = so.make_images(B=2, seed=5).cuda()
    both = enc(img)
    one = enc(img[1:2].contiguous())
    torch.testing.assert_close(both[1:2], one, atol=1e-5, rtol=1e-5)
    assert torch.isfinite(both).all()


def test_sharded_semantic_score_single_rank_matches_oracle():
    """dist.sharded_semantic_score with the CUDA scorer (world 1 = one shard holding every object)"""
    from oracle import ism_oracle as io
    from sam6d_b200 import dist as sdist
    desc, refs = io.make_descriptors(P=40, O=6, T=42, C=256, seed=9)[:2]
    got = sdist.sharded_semantic_score(desc.cuda(), refs.cuda(), 0, confidence_thresh=0.2)
    ref = io.compute_semantic_score(desc, refs, confidence_thresh=0.2)[:4]
    assert torch.equal(got[0].cpu(), ref[0]) and torch.equal(got[1].cpu(), ref[1]) and torch.equal(got[3].cpu(), ref[3])
    torch.testing.assert_close(got[2].cpu(), ref[2], atol=1e-5, rtol=1e-5)
