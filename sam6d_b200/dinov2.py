"""ISM descriptor branch on B200 kernels: drop-ins for
    DinoVisionTransformer (ViT-L/14 = dinov2_vitl14)   ISM/model/vision_transformer.py:43-375
    CustomDINOv2                                        ISM/model/dinov2.py:92-258
    MaskedPatch_MatrixSimilarity (compute_straight / compute_visible_ratio)   ISM/model/loss.py:46-77
    compute_appearance_score / compute_geometric_score  ISM/model/detector.py:298-323

The trunk is a pre-norm ViT with LayerScale: parameter names are the reference's (`cls_token`, `pos_embed` (1, 1370, C),
`mask_token`, `patch_embed.proj`, `blocks.N.{norm1, attn.{qkv,proj}, ls1.gamma, norm2, mlp.{fc1,fc2}, ls2.gamma}`, `norm`), so
`dinov2_vitl14_pretrain.pth` loads unchanged.  Every forward runs through the C ABI:
    patch embedding, qkv / proj / fc1 / fc2                -> sam6d_gemm_tma / sam6d_gemm_tc (tcgen05; bias, GELU, residual epilogues;
                                                              LayerScale folded into proj / fc2 when the weights are packed)
    LayerNorm                                              -> sam6d_layernorm_bf16
    attention over 257 tokens (16 heads x 64)              -> sam6d_attn_tc_ex on keys 0..255 (tensor cores, log-sum-exp out)
                                                              + sam6d_attn_merge_key for the 257th token
    crop / mask / nearest resize / pad of all proposals    -> sam6d_crop_resize_pad
    masked, normalised patch tokens                        -> sam6d_masked_patch_normalize
    appearance score + visible ratio                       -> sam6d_gemm_tma_batched (256 x 256 x 1024 per proposal) + sam6d_appearance_reduce
There is no CPU path.  The positional-embedding interpolation (bicubic, once per input size) is weight preprocessing in torch."""
import ctypes
import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .pem import _W, _f32, _Packed, _param_key

_ACT_GELU = 2


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _PatchEmbed(nn.Module):
    def __init__(self, patch, in_chans, dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim, bias=True)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _LayerScale(nn.Module):
    def __init__(self, dim, init_values):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class _Block(nn.Module):
    def __init__(self, dim, mlp_ratio, init_values):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.ls1 = _LayerScale(dim, init_values)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = _LayerScale(dim, init_values)


class DinoVisionTransformer(nn.Module):
    """forward(x (B,3,H,W), is_training=False) -> x_norm_clstoken (B,C), or with is_training=True the reference's dict with
    'x_norm_clstoken' and 'x_norm_patchtokens' (vision_transformer.py:232-267, 325-330).  H, W multiples of the patch size with at
    most 256 patches (the 224 x 224 proposal crops of SAM-6D give 16 x 16)."""

    def __init__(self, img_size=518, patch_size=14, in_chans=3, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0, init_values=1.0,
                 interpolate_offset=0.1, interpolate_antialias=False, num_register_tokens=0):
        super().__init__()
        if embed_dim // num_heads != 64 or num_register_tokens or interpolate_antialias:
            raise ValueError("sam6d_b200 DinoVisionTransformer: head dim 64, no register tokens (dinov2_vit{s,b,l}14)")
        self.embed_dim, self.num_heads, self.patch_size, self.depth = embed_dim, num_heads, patch_size, depth
        self.interpolate_offset = interpolate_offset
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim)
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, mlp_ratio, init_values) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self._packed = _Packed()
        self._pos_cache = {}

    # ---- weights in kernel form ------------------------------------------------------------------------------------
    def _weights(self):
        key = _param_key(self)
        if self._packed.key != key:
            C, P = self.embed_dim, self.patch_size
            K = 3 * P * P
            Kp = (K + 7) // 8 * 8                                 # 588 -> 592: the GEMM wants K % 8 == 0 (zero columns)
            pw = torch.zeros(C, Kp, dtype=torch.float32, device=self.cls_token.device)
            pw[:, :K] = _f32(self.patch_embed.proj.weight).reshape(C, K)
            w = dict(pe_w=_W(pw), pe_b=_f32(self.patch_embed.proj.bias), K=K, Kp=Kp, nw=_f32(self.norm.weight), nb=_f32(self.norm.bias),
                     blocks=[])
            for blk in self.blocks:
                g1, g2 = _f32(blk.ls1.gamma).double(), _f32(blk.ls2.gamma).double()
                # x + gamma * (W y + b) = x + (diag(gamma) W) y + gamma * b : LayerScale folded into the projection
                w["blocks"].append(dict(
                    n1w=_f32(blk.norm1.weight), n1b=_f32(blk.norm1.bias), qkv=_W(blk.attn.qkv.weight), qkv_b=_f32(blk.attn.qkv.bias),
                    proj=_W((_f32(blk.attn.proj.weight).double() * g1[:, None]).float()), proj_b=(_f32(blk.attn.proj.bias).double() * g1).float().contiguous(),
                    n2w=_f32(blk.norm2.weight), n2b=_f32(blk.norm2.bias), f1=_W(blk.mlp.fc1.weight), f1b=_f32(blk.mlp.fc1.bias),
                    f2=_W((_f32(blk.mlp.fc2.weight).double() * g2[:, None]).float()), f2b=(_f32(blk.mlp.fc2.bias).double() * g2).float().contiguous()))
            self._packed.w, self._packed.key = w, key
            self._pos_cache = {}
        return self._packed.w

    def _pos(self, npatch, w, h):
        """interpolate_pos_encoding (vision_transformer.py:179-207) -> (cls row (C,), patch rows (npatch, C)) with cls_token added"""
        ck = (npatch, w, h, self._packed.key)
        if ck not in self._pos_cache:
            pe = self.pos_embed.detach().float()
            N = pe.shape[1] - 1
            if not (npatch == N and w == h):
                dim = pe.shape[-1]
                w0, h0 = w // self.patch_size + self.interpolate_offset, h // self.patch_size + self.interpolate_offset
                sq = math.sqrt(N)
                patch_pe = F.interpolate(pe[:, 1:].reshape(1, int(sq), int(sq), dim).permute(0, 3, 1, 2), scale_factor=(float(w0) / sq, float(h0) / sq),
                                         mode="bicubic", antialias=False)
                pe = torch.cat((pe[:, :1], patch_pe.permute(0, 2, 3, 1).reshape(1, -1, dim)), dim=1)
            cls = (self.cls_token.detach().float().reshape(-1) + pe[0, 0]).contiguous()
            self._pos_cache = {ck: (cls, pe[0, 1:].contiguous())}
        return self._pos_cache[ck]

    def _attention(self, qk, vt, B, S, C):
        H, d = self.num_heads, C // self.num_heads
        scale = d ** -0.5
        if S <= 256:
            return ops.attn_tc(qk, 0, qk, C, vt, B, H, S, S, d, scale, out_dtype=torch.bfloat16)
        # 257 tokens: keys 0..255 (class token + 255 patches) on the tensor cores, then the last patch token merged by log-sum-exp.
        # (The window starts at key 0 because a TMA box must start on a 16-byte boundary of the V^T rows.)
        out, lse = ops.attn_tc_ex(qk, 0, qk, C, vt, B, H, S, S - 1, d, scale, k_brows=S, k_row0=0, v_col0=0, want_lse=True)
        ops.attn_merge_key(qk, 0, qk, C, S, S - 1, vt, S - 1, lse, B, H, S, scale, out)
        return out

    @torch.no_grad()
    def forward_features(self, x, masks=None):
        if masks is not None:
            raise NotImplementedError("mask tokens are a training feature of DINOv2")
        if not x.is_cuda:
            raise RuntimeError("sam6d_b200 DinoVisionTransformer needs CUDA tensors: there is no CPU path")
        w = self._weights()
        B, Cin, Himg, Wimg = x.shape
        P, C = self.patch_size, self.embed_dim
        Gh, Gw = Himg // P, Wimg // P
        L, S = Gh * Gw, Gh * Gw + 1
        if Himg % P or Wimg % P or L > 256:
            raise RuntimeError("input must be a multiple of the patch size with at most 256 patches")
        cls, pos = self._pos(L, Himg, Wimg)
        K, Kp = w["K"], w["Kp"]
        patches = torch.zeros(B * L, Kp, dtype=torch.float32, device=x.device)
        patches[:, :K] = x.float().reshape(B, Cin, Gh, P, Gw, P).permute(0, 2, 4, 1, 3, 5).reshape(B * L, K)
        tok = torch.empty(B, S, C, dtype=torch.float32, device=x.device)
        tok[:, 0, :] = cls
        ops.gemm_tc_raw(patches.data_ptr(), 0, w["pe_w"].bf16.data_ptr(), 1, w["pe_b"], pos.data_ptr(), tok.data_ptr() + C * 4, 0,
                        L, C, Kp, Kp, Kp, C, C, batch=B, sA=L * Kp, sW=0, sC=S * C, sR=0)
        tok = tok.view(B * S, C)
        for bw in w["blocks"]:
            xn = ops.layernorm_bf16(tok, bw["n1w"], bw["n1b"], eps=1e-6)
            qk, vt = ops.gemm_tma_vt(xn, bw["qkv"].bf16, bw["qkv_b"], 2 * C, S, slot=4)
            att = self._attention(qk, vt, B, S, C)
            tok = ops.gemm_tma(att, bw["proj"].bf16, bw["proj_b"], residual=tok)
            xn = ops.layernorm_bf16(tok, bw["n2w"], bw["n2b"], eps=1e-6)
            hid = ops.gemm_tma(xn, bw["f1"].bf16, bw["f1b"], act=_ACT_GELU, out_dtype=torch.bfloat16)
            tok = ops.gemm_tma(hid, bw["f2"].bf16, bw["f2b"], residual=tok)
        xn = ops.layernorm(tok, w["nw"], w["nb"], eps=1e-6).view(B, S, C)
        return {"x_norm_clstoken": xn[:, 0], "x_norm_regtokens": xn[:, 1:1], "x_norm_patchtokens": xn[:, 1:], "x_prenorm": tok.view(B, S, C),
                "masks": masks}

    @torch.no_grad()
    def forward(self, *args, is_training=False, **kwargs):
        ret = self.forward_features(*args, **kwargs)
        return ret if is_training else ret["x_norm_clstoken"]


def vit_large(patch_size=14, **kwargs):
    """dinov2_vitl14 (vision_transformer.py:364-375 with the arguments of _make_dinov2_model, dinov2.py:46-90)"""
    kw = dict(img_size=518, init_values=1.0, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4)
    kw.update(kwargs)
    return DinoVisionTransformer(patch_size=patch_size, **kw)


# =====================================================================================================================
def crop_resize_pad(image_u8: Optional[torch.Tensor], masks: torch.Tensor, boxes: torch.Tensor, target: int = 224, want_rgb=True, want_mask=True):
    """all proposals of a frame in one launch pair: image (H,W,3) uint8, masks (P,H,W) float, boxes (P,4) xyxy
    -> (rgb (P,3,T,T) or None, mask (P,T,T) or None)   (CustomDINOv2.process_rgb_proposals / process_masks_proposals)"""
    P, H, W = masks.shape
    dev = masks.device
    m = masks.float().contiguous()
    b = boxes.to(torch.int32).contiguous()
    rgb = torch.empty(P, 3, target, target, dtype=torch.float32, device=dev) if want_rgb else None
    pm = torch.empty(P, target, target, dtype=torch.float32, device=dev) if want_mask else None
    img = image_u8.contiguous() if want_rgb else None
    _lib.call("sam6d_crop_resize_pad", _p(img), _p(m), _p(b), P, H, W, target, _p(rgb), _p(pm), _s())
    return rgb, pm


class CustomDINOv2(nn.Module):
    """ISM/model/dinov2.py:92-258 without the Lightning base: forward(image_np (H,W,3) uint8, proposals with .masks (P,H,W) and
    .boxes (P,4)) -> (cls_features (P,C), patch_features (P,256,C) masked + L2-normalised)."""

    def __init__(self, model_name="dinov2_vitl14", token_name="x_norm_clstoken", image_size=224, chunk_size=16, descriptor_width_size=640,
                 checkpoint_dir=None, patch_size=14, validpatch_thresh=0.5, model: Optional[nn.Module] = None):
        super().__init__()
        if model_name != "dinov2_vitl14" and model is None:
            raise NotImplementedError("SAM-6D configures dinov2_vitl14")
        self.model_name = model_name
        self.model = model if model is not None else vit_large(patch_size)
        if checkpoint_dir is not None:
            import os.path as osp
            self.model.load_state_dict(torch.load(osp.join(checkpoint_dir, f"{model_name}_pretrain.pth"), map_location="cpu"))
        self.validpatch_thresh, self.token_name, self.chunk_size = validpatch_thresh, token_name, chunk_size
        self.patch_size, self.proposal_size, self.descriptor_width_size = patch_size, image_size, descriptor_width_size

    def _image(self, image_np, device):
        img = image_np if torch.is_tensor(image_np) else torch.from_numpy(image_np)
        return img.to(device=device, dtype=torch.uint8).contiguous()

    @torch.no_grad()
    def process_rgb_proposals(self, image_np, masks, boxes):
        return crop_resize_pad(self._image(image_np, masks.device), masks, boxes, self.proposal_size, True, False)[0]

    @torch.no_grad()
    def process_masks_proposals(self, masks, boxes):
        return crop_resize_pad(None, masks, boxes, self.proposal_size, False, True)[1]

    @torch.no_grad()
    def compute_cls_and_patch_features(self, images, masks, want_bf16=False):
        """dinov2.py:248-258 (+ the bf16 copy / validity flags the scoring kernels use)"""
        P = images.shape[0]
        C, G = self.model.embed_dim, self.proposal_size // self.patch_size
        cls = torch.empty(P, C, dtype=torch.float32, device=images.device)
        pf = torch.empty(P, G * G, C, dtype=torch.float32, device=images.device)
        pb = torch.empty(P, G * G, C, dtype=torch.bfloat16, device=images.device) if want_bf16 else None
        valid = torch.empty(P, G * G, dtype=torch.uint8, device=images.device)
        chunk = max(self.chunk_size, 64)          # the reference's 16 bounds eager-attention memory; the kernels batch more
        for i in range(0, P, chunk):
            f = self.model(images[i:i + chunk].contiguous(), is_training=True)
            n = f["x_norm_clstoken"].shape[0]
            cls[i:i + n] = f["x_norm_clstoken"]
            pt = f["x_norm_patchtokens"]                                     # view of (n, S, C): row stride C, batch stride S*C
            mk = masks[i:i + n].contiguous()                                 # named: must outlive the launch
            _lib.call("sam6d_masked_patch_normalize", _p(pt), ctypes.c_longlong(pt.stride(1)), ctypes.c_longlong(pt.stride(0)),
                      _p(mk), n, G, self.patch_size, C, ctypes.c_float(self.validpatch_thresh), _p(pf[i:i + n]),
                      _p(pb[i:i + n]) if want_bf16 else None, _p(valid[i:i + n]), _s())
        self.last_patch_bf16, self.last_valid = pb, valid
        return cls, pf

    @torch.no_grad()
    def forward(self, image_np, proposals, want_bf16=False):
        masks, boxes = proposals.masks, proposals.boxes
        rgbs, pmasks = crop_resize_pad(self._image(image_np, masks.device), masks, boxes, self.proposal_size, True, True)
        return self.compute_cls_and_patch_features(rgbs, pmasks, want_bf16)

    @torch.no_grad()
    def forward_cls_token(self, image_np, proposals):
        return self.forward(image_np, proposals)[0]

    @torch.no_grad()
    def forward_patch_tokens(self, image_np, proposals):
        return self.forward(image_np, proposals)[1]


class MaskedPatch_MatrixSimilarity(nn.Module):
    """ISM/model/loss.py:46-77: compute_straight (appearance score) and compute_visible_ratio on (P, N, C) patch descriptors
    (masked rows are zero), N <= 256.  One batched tensor-core GEMM + one reduction kernel produce both."""

    def __init__(self, metric="cosine", chunk_size=64):
        super().__init__()
        self.metric, self.chunk_size = metric, chunk_size

    @torch.no_grad()
    def scores(self, query, reference, thred=0.5):
        P, N, C = query.shape
        q = query.to(torch.bfloat16).contiguous()
        r = reference.to(torch.bfloat16).contiguous()
        ld = (N + 3) // 4 * 4
        sim = torch.empty(P, N, ld, dtype=torch.float32, device=query.device)
        ops.gemm_tma_batched(q, r, sim, N, N, ld, N * ld)
        qvalid = (query.abs().amax(dim=-1) > 0).to(torch.uint8).contiguous()
        appe = torch.empty(P, dtype=torch.float32, device=query.device)
        vis = torch.empty(P, dtype=torch.float32, device=query.device)
        _lib.call("sam6d_appearance_reduce", _p(sim), ctypes.c_longlong(ld), ctypes.c_longlong(N * ld), P, N, _p(qvalid), ctypes.c_float(thred),
                  _p(appe), _p(vis), _s())
        return appe, vis

    def compute_straight(self, query, reference):
        return self.scores(query, reference)[0]

    def compute_visible_ratio(self, query, reference, thred=0.5):
        return self.scores(query, reference, thred)[1]


def compute_appearance_score(ref_appe_descriptors_all, best_pose, pred_objects_idx, query_appe_descriptors):
    """detector.py:298-309: gather the best template's patch descriptors per proposal, then compute_straight"""
    ref = ref_appe_descriptors_all[pred_objects_idx, best_pose, ...]
    return MaskedPatch_MatrixSimilarity().compute_straight(query_appe_descriptors, ref), ref
