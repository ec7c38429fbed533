"""oracle/sam_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32) of SAM's ViT image encoder as the Instance Segmentation Model uses it:
    ImageEncoderViT.forward   ISM/segment_anything/modeling/image_encoder.py:106-116
    Block.forward             :166-182   (window partition AFTER norm1, padding 64 -> 70)
    Attention.forward         :224-240   (q*scale for QK^T, UNSCALED q for the rel-pos bias)
    window_partition / window_unpartition :243-290,   get_rel_pos / add_decomposed_rel_pos :293-361
    PatchEmbed :364-395,  MLPBlock / LayerNorm2d  ISM/segment_anything/modeling/common.py:13-43
over a flat state_dict with the reference's key names (`sam_vit_h_4b8939.pth: image_encoder.*`).

Parity status: PINNED -- tools/make_golden.py imports the vendored reference module (it imports cleanly here), loads the same
seeded state dict and checks this restatement against it; fixture tests/golden/sam_small.pt.
"""
import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def vit_cfg(name: str = "vit_h"):
    """build_sam.py:14-46"""
    if name == "vit_h":
        return dict(embed_dim=1280, depth=32, num_heads=16, global_attn_indexes=(7, 15, 23, 31))
    if name == "vit_l":
        return dict(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23))
    return dict(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11))


def window_partition(x: torch.Tensor, ws: int):
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)
    return x, (Hp, Wp)


def window_unpartition(w: torch.Tensor, ws: int, pad_hw: Tuple[int, int], hw: Tuple[int, int]):
    Hp, Wp = pad_hw
    H, W = hw
    B = w.shape[0] // (Hp * Wp // ws // ws)
    x = w.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
    return x[:, :H, :W, :].contiguous()


def rel_pos_table(size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """get_rel_pos for q_size == k_size == size (no interpolation: the table already has 2*size-1 rows)"""
    assert rel_pos.shape[0] == 2 * size - 1
    coords = torch.arange(size)[:, None] - torch.arange(size)[None, :] + (size - 1)
    return rel_pos[coords.long()]                                        # (size, size, head_dim)


def attention(sd: SD, p: str, x: torch.Tensor, num_heads: int) -> torch.Tensor:
    B, H, W, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"]).reshape(B, H * W, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * num_heads, H * W, -1).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh, Rw = rel_pos_table(H, sd[p + ".rel_pos_h"]), rel_pos_table(W, sd[p + ".rel_pos_w"])
    rq = q.reshape(B * num_heads, H, W, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(-1, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(-1, H * W, H * W)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).view(B, num_heads, H, W, -1).permute(0, 2, 3, 1, 4).reshape(B, H, W, -1)
    return F.linear(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def block(sd: SD, p: str, x: torch.Tensor, num_heads: int, window: int, eps: float) -> torch.Tensor:
    C = x.shape[-1]
    shortcut = x
    x = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    if window > 0:
        H, W = x.shape[1], x.shape[2]
        x, pad_hw = window_partition(x, window)
    x = attention(sd, p + ".attn", x, num_heads)
    if window > 0:
        x = window_unpartition(x, window, pad_hw, (H, W))
    x = shortcut + x
    y = F.layer_norm(x, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    y = F.linear(F.gelu(F.linear(y, sd[p + ".mlp.lin1.weight"], sd[p + ".mlp.lin1.bias"])), sd[p + ".mlp.lin2.weight"],
                 sd[p + ".mlp.lin2.bias"])
    return x + y


def layernorm2d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def image_encoder(sd: SD, img: torch.Tensor, num_heads: int, global_attn_indexes, window_size: int = 14, patch: int = 16,
                  eps: float = 1e-6, prefix: str = "") -> torch.Tensor:
    """img (B,3,S,S) -> (B,256,S/16,S/16)"""
    x = F.conv2d(img, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=patch).permute(0, 2, 3, 1)
    x = x + sd[prefix + "pos_embed"]
    depth = 1 + max(int(k[len(prefix):].split(".")[1]) for k in sd if k.startswith(prefix + "blocks."))
    for i in range(depth):
        x = block(sd, f"{prefix}blocks.{i}", x, num_heads, 0 if i in global_attn_indexes else window_size, eps)
    x = x.permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[prefix + "neck.0.weight"])
    x = layernorm2d(x, sd[prefix + "neck.1.weight"], sd[prefix + "neck.1.bias"])
    x = F.conv2d(x, sd[prefix + "neck.2.weight"], padding=1)
    return layernorm2d(x, sd[prefix + "neck.3.weight"], sd[prefix + "neck.3.bias"])


# seeded weights / images: shared with bench.py (data, not algorithm)
from sam6d_b200.synth import make_sam_state_dict as make_state_dict, make_images  # noqa: E402,F401
