"""oracle/sam_dec_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32) of the SAM prompt encoder, mask decoder and automatic-mask-generator filtering as the Instance
Segmentation Model runs them:
    dense_pe / embed_points        PromptEncoder.get_dense_pe, _embed_points, PositionEmbeddingRandom
                                   ISM/segment_anything/modeling/prompt_encoder.py:62-214
    attention / two_way_transformer TwoWayTransformer, TwoWayAttentionBlock, Attention
                                   ISM/segment_anything/modeling/transformer.py:16-240
    mask_decoder                   MaskDecoder.predict_masks + the multimask slice   modeling/mask_decoder.py:71-176
    postprocess_masks              Sam.postprocess_masks                             modeling/sam.py:133-162
    process_batch / nms / generate SamAutomaticMaskGenerator._process_batch, _process_crop (one crop: crop_n_layers = 0),
                                   calculate_stability_score, batched_mask_to_box    automatic_mask_generator.py:225-321,
                                   utils/amg.py:156-176,303-345; CustomSamAutomaticMaskGenerator ISM/model/sam.py:52-155
over a flat state_dict with the reference's key names (`sam_vit_h_4b8939.pth`: `prompt_encoder.*`, `mask_decoder.*`).
Parity status: PINNED -- tools/make_golden_sam_dec.py builds the vendored reference modules (they import here), loads the same
seeded state dict and finds this restatement bit-identical; fixture tests/golden/sam_dec.pt."""
import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
C, HEADS, T_OUT = 256, 8, 5            # transformer dim, heads, output tokens (iou + 4 mask tokens)


def _lin(sd: SD, n: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[n + ".weight"], sd[n + ".bias"])


def _ln(sd: SD, n: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[n + ".weight"], sd[n + ".bias"], 1e-5)


# ---- prompt encoder ----------------------------------------------------------------------------------------------------
def _pe_encoding(sd: SD, coords: torch.Tensor) -> torch.Tensor:
    coords = 2 * coords - 1
    coords = coords @ sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    coords = 2 * np.pi * coords
    return torch.cat([torch.sin(coords), torch.cos(coords)], dim=-1)


def dense_pe(sd: SD, size: int = 64) -> torch.Tensor:
    """get_dense_pe -> (1, 256, size, size)"""
    grid = torch.ones((size, size), dtype=torch.float32)
    y = (grid.cumsum(dim=0) - 0.5) / size
    x = (grid.cumsum(dim=1) - 0.5) / size
    return _pe_encoding(sd, torch.stack([x, y], dim=-1)).permute(2, 0, 1).unsqueeze(0)


def embed_points(sd: SD, points: torch.Tensor, labels: torch.Tensor, image_size: int = 1024) -> torch.Tensor:
    """_embed_points(pad=True): points (B,N,2) in the 1024-frame, labels (B,N) -> (B, N+1, 256)"""
    points = points + 0.5
    points = torch.cat([points, torch.zeros((points.shape[0], 1, 2))], dim=1)
    labels = torch.cat([labels, -torch.ones((labels.shape[0], 1))], dim=1)
    coords = points.clone()
    coords[:, :, 0] = coords[:, :, 0] / image_size
    coords[:, :, 1] = coords[:, :, 1] / image_size
    e = _pe_encoding(sd, coords.to(torch.float))
    e[labels == -1] = 0.0
    e[labels == -1] += sd["prompt_encoder.not_a_point_embed.weight"]
    e[labels == 0] += sd["prompt_encoder.point_embeddings.0.weight"]
    e[labels == 1] += sd["prompt_encoder.point_embeddings.1.weight"]
    return e


# ---- two-way transformer ---------------------------------------------------------------------------------------------------
def attention(sd: SD, p: str, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    q, k, v = _lin(sd, p + ".q_proj", q), _lin(sd, p + ".k_proj", k), _lin(sd, p + ".v_proj", v)
    sep = lambda x: x.reshape(x.shape[0], x.shape[1], HEADS, x.shape[2] // HEADS).transpose(1, 2)   # noqa: E731
    q, k, v = sep(q), sep(k), sep(v)
    a = torch.softmax(q @ k.permute(0, 1, 3, 2) / math.sqrt(q.shape[-1]), dim=-1)
    o = (a @ v).transpose(1, 2)
    return _lin(sd, p + ".out_proj", o.reshape(o.shape[0], o.shape[1], -1))


def two_way_transformer(sd: SD, src: torch.Tensor, pos: torch.Tensor, tokens: torch.Tensor, p: str = "mask_decoder.transformer"):
    """src, pos (B,256,h,w); tokens (B,T,256) -> (queries (B,T,256), keys (B,h*w,256))"""
    keys = src.flatten(2).permute(0, 2, 1)
    key_pe = pos.flatten(2).permute(0, 2, 1)
    queries, query_pe = tokens, tokens
    for i in range(2):
        l = f"{p}.layers.{i}"
        if i == 0:
            queries = attention(sd, l + ".self_attn", queries, queries, queries)
        else:
            q = queries + query_pe
            queries = queries + attention(sd, l + ".self_attn", q, q, queries)
        queries = _ln(sd, l + ".norm1", queries)
        q, k = queries + query_pe, keys + key_pe
        queries = _ln(sd, l + ".norm2", queries + attention(sd, l + ".cross_attn_token_to_image", q, k, keys))
        m = _lin(sd, l + ".mlp.lin2", F.relu(_lin(sd, l + ".mlp.lin1", queries)))
        queries = _ln(sd, l + ".norm3", queries + m)
        q, k = queries + query_pe, keys + key_pe
        keys = _ln(sd, l + ".norm4", keys + attention(sd, l + ".cross_attn_image_to_token", k, q, queries))
    q, k = queries + query_pe, keys + key_pe
    queries = _ln(sd, p + ".norm_final_attn", queries + attention(sd, p + ".final_attn_token_to_image", q, k, keys))
    return queries, keys


def mask_decoder(sd: SD, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse: torch.Tensor, multimask_output: bool = True):
    """predict_masks with the no-mask dense embedding -> (low-res masks (B,3,256,256), iou predictions (B,3))"""
    B = sparse.shape[0]
    out_tokens = torch.cat([sd["mask_decoder.iou_token.weight"], sd["mask_decoder.mask_tokens.weight"]], dim=0)
    tokens = torch.cat((out_tokens.unsqueeze(0).expand(B, -1, -1), sparse), dim=1)
    h, w = image_embeddings.shape[-2:]
    dense = sd["prompt_encoder.no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(B, -1, h, w)
    src = torch.repeat_interleave(image_embeddings, B, dim=0) + dense
    pos = torch.repeat_interleave(image_pe, B, dim=0)
    hs, src = two_way_transformer(sd, src, pos, tokens)
    iou_tok, mask_tok = hs[:, 0, :], hs[:, 1:T_OUT, :]
    src = src.transpose(1, 2).view(B, C, h, w)
    u = "mask_decoder.output_upscaling"
    x = F.conv_transpose2d(src, sd[u + ".0.weight"], sd[u + ".0.bias"], stride=2)
    m, s_ = x.mean(1, keepdim=True), (x - x.mean(1, keepdim=True)).pow(2).mean(1, keepdim=True)          # LayerNorm2d, eps 1e-6
    x = sd[u + ".1.weight"][:, None, None] * ((x - m) / torch.sqrt(s_ + 1e-6)) + sd[u + ".1.bias"][:, None, None]
    x = F.gelu(x)
    x = F.gelu(F.conv_transpose2d(x, sd[u + ".3.weight"], sd[u + ".3.bias"], stride=2))
    hyper = []
    for i in range(4):
        y = mask_tok[:, i, :]
        for j in range(3):
            y = _lin(sd, f"mask_decoder.output_hypernetworks_mlps.{i}.layers.{j}", y)
            if j < 2:
                y = F.relu(y)
        hyper.append(y)
    hyper = torch.stack(hyper, dim=1)
    b, c, hh, ww = x.shape
    masks = (hyper @ x.view(b, c, hh * ww)).view(b, -1, hh, ww)
    y = iou_tok
    for j in range(3):
        y = _lin(sd, f"mask_decoder.iou_prediction_head.layers.{j}", y)
        if j < 2:
            y = F.relu(y)
    sl = slice(1, None) if multimask_output else slice(0, 1)
    return masks[:, sl], y[:, sl]


def postprocess_masks(masks: torch.Tensor, input_size: Tuple[int, int], original_size: Tuple[int, int], img_size: int = 1024) -> torch.Tensor:
    masks = F.interpolate(masks, (img_size, img_size), mode="bilinear", align_corners=False)
    masks = masks[..., : input_size[0], : input_size[1]]
    return F.interpolate(masks, original_size, mode="bilinear", align_corners=False)


# ---- automatic mask generator (single crop = the whole image, crop_n_layers = 0) -----------------------------------------------
def build_point_grid(n_per_side: int = 32) -> np.ndarray:
    """utils/amg.py:179-187"""
    offset = 1 / (2 * n_per_side)
    pts = np.linspace(offset, 1 - offset, n_per_side)
    return np.stack([np.tile(pts[None, :], (n_per_side, 1)), np.tile(pts[:, None], (1, n_per_side))], axis=-1).reshape(-1, 2)


def preprocess_shape(oldh: int, oldw: int, long_side: int = 1024) -> Tuple[int, int]:
    """ResizeLongestSide.get_preprocess_shape (utils/transforms.py:93-101)"""
    scale = long_side * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


def apply_coords(points: np.ndarray, original_size: Tuple[int, int]) -> np.ndarray:
    new_h, new_w = preprocess_shape(*original_size)
    c = points.astype(float).copy()
    c[..., 0] = c[..., 0] * (new_w / original_size[1])
    c[..., 1] = c[..., 1] * (new_h / original_size[0])
    return c


def stability_score(masks: torch.Tensor, thr: float = 0.0, off: float = 1.0) -> torch.Tensor:
    inter = (masks > (thr + off)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    union = (masks > (thr - off)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    return inter / union


def batched_mask_to_box(masks: torch.Tensor) -> torch.Tensor:
    """utils/amg.py:303-345 for (N,H,W) bool masks -> (N,4) xyxy, [0,0,0,0] for an empty mask"""
    if masks.numel() == 0:
        return torch.zeros(masks.shape[0], 4)
    h, w = masks.shape[-2:]
    in_h, _ = torch.max(masks, dim=-1)
    hc = in_h * torch.arange(h)[None, :]
    bottom, _ = torch.max(hc, dim=-1)
    top, _ = torch.min(hc + h * (~in_h), dim=-1)
    in_w, _ = torch.max(masks, dim=-2)
    wc = in_w * torch.arange(w)[None, :]
    right, _ = torch.max(wc, dim=-1)
    left, _ = torch.min(wc + w * (~in_w), dim=-1)
    empty = (right < left) | (bottom < top)
    return torch.stack([left, top, right, bottom], dim=-1) * (~empty).unsqueeze(-1)


def process_batch(sd: SD, features: torch.Tensor, image_pe: torch.Tensor, points: np.ndarray, im_size: Tuple[int, int],
                  pred_iou_thresh: float = 0.88, stability_score_thresh: float = 0.95, return_all: bool = False):
    """_process_batch for the whole-image crop: points (n,2) in image pixels -> dict(masks bool (k,H,W), boxes (k,4), iou_preds (k))"""
    tp = torch.as_tensor(apply_coords(points, im_size))
    sparse = embed_points(sd, tp[:, None, :].float(), torch.ones(tp.shape[0], 1))
    low, iou = mask_decoder(sd, features, image_pe, sparse)
    masks = postprocess_masks(low, preprocess_shape(*im_size), im_size).flatten(0, 1)
    iou = iou.flatten(0, 1)
    keep = iou > pred_iou_thresh
    masks_k, iou_k = masks[keep], iou[keep]
    st = stability_score(masks_k)
    keep2 = st >= stability_score_thresh
    masks_k, iou_k, st = masks_k[keep2], iou_k[keep2], st[keep2]
    mb = masks_k > 0.0
    boxes = batched_mask_to_box(mb)
    # is_box_near_crop_edge with crop box == image box never filters (near_crop_edge & ~near_image_edge is empty)
    out = dict(masks=mb, boxes=boxes, iou_preds=iou_k, stability=st)
    if return_all:
        out.update(low_res=low, iou_all=iou, logits=masks)
    return out


def nms(boxes: torch.Tensor, scores: torch.Tensor, thr: float) -> torch.Tensor:
    """torchvision.ops.nms semantics (all boxes in one category): indices kept, by decreasing score"""
    order = torch.argsort(scores, descending=True, stable=True)
    b = boxes.float()
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    keep: List[int] = []
    dead = torch.zeros(len(b), dtype=torch.bool)
    for i in order.tolist():
        if dead[i]:
            continue
        keep.append(i)
        xx1, yy1 = torch.maximum(b[i, 0], b[:, 0]), torch.maximum(b[i, 1], b[:, 1])
        xx2, yy2 = torch.minimum(b[i, 2], b[:, 2]), torch.minimum(b[i, 3], b[:, 3])
        inter = (xx2 - xx1).clamp(min=0) * (yy2 - yy1).clamp(min=0)
        dead |= inter / (area[i] + area - inter) > thr
    return torch.tensor(keep, dtype=torch.long)


def generate_masks(sd: SD, features: torch.Tensor, im_size: Tuple[int, int], points_per_side: int = 32, points_per_batch: int = 64,
                   pred_iou_thresh: float = 0.88, stability_score_thresh: float = 0.95, box_nms_thresh: float = 0.7):
    """CustomSamAutomaticMaskGenerator.generate_masks after set_image (segmentor_width_size = the image width: no resize):
    features (1,256,64,64) of the frame -> {"masks": (N,H,W) bool, "boxes": (N,4)}"""
    pe = dense_pe(sd)
    pts = build_point_grid(points_per_side) * np.array(im_size)[None, ::-1]
    data = dict(masks=[], boxes=[], iou_preds=[])
    for i in range(0, len(pts), points_per_batch):
        r = process_batch(sd, features, pe, pts[i:i + points_per_batch], im_size, pred_iou_thresh, stability_score_thresh)
        for k in data:
            data[k].append(r[k])
    masks, boxes, iou = torch.cat(data["masks"]), torch.cat(data["boxes"]), torch.cat(data["iou_preds"])
    keep = nms(boxes, iou, box_nms_thresh) if len(boxes) else torch.zeros(0, dtype=torch.long)
    return dict(masks=masks[keep], boxes=boxes[keep], iou_preds=iou[keep])


from sam6d_b200.synth import make_sam_decoder_state_dict as make_state_dict  # noqa: E402,F401  (seeded weights, shared with bench.py)
