"""oracle/dinov2_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32) of the ISM descriptor branch:
    vit_forward                DinoVisionTransformer.forward_features   ISM/model/vision_transformer.py:212-267 (ViT-L/14: :364-375)
                               with interpolate_pos_encoding :179-207, Block.forward ISM/model/layers/block.py:79-104 (LayerScale
                               ISM/model/layers/layer_scale.py:16-28), Attention.forward ISM/model/layers/attention.py:47-62
    crop_resize_pad            CropResizePad.__call__                   ISM/utils/bbox_utils.py:89-126
    process_rgb_proposals / process_masks_proposals / cls_and_patch_features
                               CustomDINOv2                             ISM/model/dinov2.py:131-147, 175-186, 228-258
over a flat state_dict with the reference's key names (`dinov2_vitl14_pretrain.pth`).
Parity status: PINNED -- tools/make_golden_dinov2.py instantiates the reference's own vit_large(patch_size=14, img_size=518,
init_values=1.0, block_chunks=0) and CropResizePad / CustomDINOv2 methods (imported from /root/reference, pytorch_lightning
stubbed), loads the same seeded state dict and finds this restatement bit-identical; fixture tests/golden/dinov2.pt."""
import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def interpolate_pos_encoding(pos_embed: torch.Tensor, npatch: int, w: int, h: int, patch: int = 14, offset: float = 0.1) -> torch.Tensor:
    """vision_transformer.py:179-207"""
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    pe = pos_embed.float()
    cls_pe, patch_pe = pe[:, 0], pe[:, 1:]
    dim = pe.shape[-1]
    w0, h0 = w // patch + offset, h // patch + offset
    sqrt_N = math.sqrt(N)
    sx, sy = float(w0) / sqrt_N, float(h0) / sqrt_N
    patch_pe = F.interpolate(patch_pe.reshape(1, int(sqrt_N), int(sqrt_N), dim).permute(0, 3, 1, 2), scale_factor=(sx, sy), mode="bicubic",
                             antialias=False)
    assert int(w0) == patch_pe.shape[-2] and int(h0) == patch_pe.shape[-1]
    return torch.cat((cls_pe.unsqueeze(0), patch_pe.permute(0, 2, 3, 1).view(1, -1, dim)), dim=1)


def vit_forward(sd: SD, x: torch.Tensor, num_heads: int = 16, patch: int = 14):
    """-> dict(x_norm_clstoken (B,C), x_norm_patchtokens (B,L,C))"""
    B, _, w, h = x.shape
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)
    t = t + interpolate_pos_encoding(sd["pos_embed"], t.shape[1] - 1, w, h, patch)
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    C = t.shape[-1]
    hd = C // num_heads
    for i in range(depth):
        p = f"blocks.{i}."
        y = F.layer_norm(t, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, -1, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        y = (a @ v).transpose(1, 2).reshape(B, -1, C)
        y = F.linear(y, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        t = t + y * sd[p + "ls1.gamma"]
        y = F.layer_norm(t, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        t = t + y * sd[p + "ls2.gamma"]
    n = F.layer_norm(t, (C,), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return dict(x_norm_clstoken=n[:, 0], x_norm_patchtokens=n[:, 1:])


def crop_resize_pad(images: torch.Tensor, boxes: torch.Tensor, target: int = 224) -> torch.Tensor:
    """bbox_utils.py:98-126: images (N,C,H,W), boxes (N,4) int64 xyxy -> (N,C,target,target)"""
    box_sizes = boxes[:, 2:] - boxes[:, :2]
    scale_factor = target / torch.max(box_sizes, dim=-1)[0]
    out = []
    for image, box, scale in zip(images, boxes, scale_factor):
        image = image[:, box[1]:box[3], box[0]:box[2]]
        image = F.interpolate(image.unsqueeze(0), scale_factor=scale.item())[0]
        oh, ow = image.shape[1:]
        if 1.0 != ow / oh:
            pt = max((target - oh) // 2, 0)
            pl = max((target - ow) // 2, 0)
            image = F.pad(image, (pl, target - ow - pl, pt, target - oh - pt))
        assert image.shape[1] == image.shape[2]
        image = F.interpolate(image.unsqueeze(0), scale_factor=target / image.shape[1])[0]
        out.append(image)
    return torch.stack(out)


def rgb_normalize(image_u8: torch.Tensor) -> torch.Tensor:
    """T.ToTensor + T.Normalize on an (H,W,3) uint8 image -> (3,H,W) float32"""
    x = image_u8.permute(2, 0, 1).float().div(255)
    return (x - torch.tensor(MEAN).view(3, 1, 1)) / torch.tensor(STD).view(3, 1, 1)


def process_rgb_proposals(image_u8: torch.Tensor, masks: torch.Tensor, boxes: torch.Tensor, target: int = 224) -> torch.Tensor:
    """dinov2.py:131-147"""
    rgb = rgb_normalize(image_u8)
    masked = rgb.unsqueeze(0).repeat(len(masks), 1, 1, 1) * masks.unsqueeze(1)
    return crop_resize_pad(masked, boxes, target)


def process_masks_proposals(masks: torch.Tensor, boxes: torch.Tensor, target: int = 224) -> torch.Tensor:
    """dinov2.py:175-186"""
    return crop_resize_pad(masks.unsqueeze(1), boxes, target).squeeze(1)


def cls_and_patch_features(sd: SD, images: torch.Tensor, masks: torch.Tensor, num_heads: int = 16, patch: int = 14, thresh: float = 0.5):
    """dinov2.py:248-258: cls tokens (P,C); patch tokens masked by AvgPool2d(14)(mask) > 0.5 and L2-normalised (P,L,C)"""
    f = vit_forward(sd, images, num_heads, patch)
    keep = F.avg_pool2d(masks.unsqueeze(1), patch, patch).flatten(-2).squeeze(1) > thresh
    pf = F.normalize(f["x_norm_patchtokens"] * keep.unsqueeze(-1), dim=-1)
    return f["x_norm_clstoken"], pf, keep


from sam6d_b200.synth import make_dinov2_state_dict as make_state_dict, make_proposals  # noqa: E402,F401  (seeded data, shared with bench.py)
