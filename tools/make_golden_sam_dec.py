"""tools/make_golden_sam_dec.py -- DEV CONTAINER ONLY (needs /root/reference).

Pins oracle/sam_dec_oracle.py against the vendored reference modules (ISM/segment_anything: PromptEncoder, MaskDecoder,
TwoWayTransformer, Sam.postprocess_masks, SamAutomaticMaskGenerator._process_batch helpers, torchvision batched_nms) on seeded
weights and a synthetic image embedding, and writes tests/golden/sam_dec.pt.

Usage: python tools/make_golden_sam_dec.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/SAM-6D/Instance_Segmentation_Model")

from oracle import sam_dec_oracle as so  # noqa: E402
from sam6d_b200 import synth  # noqa: E402


def main():
    torch.set_num_threads(8)
    from segment_anything.modeling import MaskDecoder, PromptEncoder, TwoWayTransformer, Sam
    from segment_anything.modeling.image_encoder import ImageEncoderViT
    from segment_anything.utils import amg
    from segment_anything.utils.transforms import ResizeLongestSide
    from torchvision.ops.boxes import batched_nms
    sd = so.make_state_dict(seed=1)
    pe = PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16).eval()
    md = MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                     transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256).eval()
    print("PromptEncoder strict load:", pe.load_state_dict({k[len("prompt_encoder."):]: v for k, v in sd.items() if k.startswith("prompt_encoder.")}, strict=True))
    print("MaskDecoder strict load:", md.load_state_dict({k[len("mask_decoder."):]: v for k, v in sd.items() if k.startswith("mask_decoder.")}, strict=True))
    feat = synth.make_image_embedding(seed=1)
    im_size = (480, 640)
    pts = so.build_point_grid(8) * np.array(im_size)[None, ::-1]
    assert np.array_equal(so.build_point_grid(32), amg.build_point_grid(32))
    tr = ResizeLongestSide(1024)
    tp = tr.apply_coords(pts, im_size)
    assert np.array_equal(tp, so.apply_coords(pts, im_size)) and tr.get_preprocess_shape(480, 640, 1024) == so.preprocess_shape(480, 640)
    with torch.no_grad():
        in_points = torch.as_tensor(tp)
        sparse, dense = pe(points=(in_points[:, None, :], torch.ones(len(tp), 1, dtype=torch.int)), boxes=None, masks=None)
        low, iou = md(image_embeddings=feat, image_pe=pe.get_dense_pe(), sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                      multimask_output=True)
        # Sam.postprocess_masks needs only image_encoder.img_size
        sam_like = type("S", (), {"image_encoder": type("E", (), {"img_size": 1024})()})()
        logits = Sam.postprocess_masks(sam_like, low, tr.get_preprocess_shape(480, 640, 1024), im_size)
    r = so.process_batch(sd, feat, so.dense_pe(sd), pts, im_size, 0.88, 0.95, return_all=True)
    for name, a, b in (("dense pe", pe.get_dense_pe(), so.dense_pe(sd)), ("sparse embeddings", sparse, so.embed_points(sd, in_points[:, None, :].float(), torch.ones(len(tp), 1))),
                       ("low-res masks", low, r["low_res"]), ("iou predictions", iou.flatten(0, 1), r["iou_all"]),
                       ("postprocessed logits", logits.flatten(0, 1), r["logits"])):
        d = (a - b).abs().max().item()
        print(f"  {name:22s} max|ref - oracle| = {d:.3e}")
        assert d == 0.0, name
    # the filters of _process_batch with the reference helpers
    m_all, i_all = logits.flatten(0, 1), iou.flatten(0, 1)
    k1 = i_all > 0.88
    st = amg.calculate_stability_score(m_all[k1], 0.0, 1.0)
    k2 = st >= 0.95
    mb = m_all[k1][k2] > 0.0
    boxes = amg.batched_mask_to_box(mb)
    assert torch.equal(mb, r["masks"]) and torch.equal(boxes, r["boxes"]) and torch.equal(i_all[k1][k2], r["iou_preds"])
    assert not amg.is_box_near_crop_edge(boxes, [0, 0, 640, 480], [0, 0, 640, 480]).any()
    # NMS against torchvision on boxes with overlaps (the synthetic masks are frame-filling, so use constructed boxes too)
    g = torch.Generator().manual_seed(0)
    xy = torch.randint(0, 400, (300, 2), generator=g).float()
    wh = torch.randint(20, 200, (300, 2), generator=g).float()
    bx = torch.cat([xy, xy + wh], dim=1)
    sc = torch.rand(300, generator=g)
    assert torch.equal(so.nms(bx, sc, 0.7), batched_nms(bx, sc, torch.zeros(300), 0.7))
    assert torch.equal(so.nms(boxes.float(), r["iou_preds"], 0.7), batched_nms(boxes.float(), r["iou_preds"], torch.zeros(len(boxes)), 0.7))
    print(f"  filters: {int(k1.sum())} of {len(k1)} pass the IoU threshold, {len(boxes)} pass stability; boxes / masks / NMS == reference helpers")
    gold = dict(meta=dict(source="segment_anything PromptEncoder / MaskDecoder / TwoWayTransformer / Sam.postprocess_masks / utils.amg imported from /root/reference",
                          torch=torch.__version__, seed=1, im_size=im_size, points_per_side=8),
                low_sub=low[:, :, ::8, ::8].clone(), low_abs_mean=low.abs().mean().item(), iou_all=i_all.clone(),
                logits_sub=logits[:, :, ::16, ::16].clone(), stability_all=amg.calculate_stability_score(m_all, 0.0, 1.0),
                kept_index=torch.nonzero(k1).flatten()[k2].clone(), boxes=boxes.clone(), mask_area=mb.flatten(1).sum(1).clone(),
                dense_pe_sub=pe.get_dense_pe()[:, :, ::8, ::8].clone(), sparse=sparse.clone(),
                feat_checksum=feat.double().sum().item())
    path = os.path.join(ROOT, "tests", "golden", "sam_dec.pt")
    torch.save(gold, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
