"""tools/make_golden_dinov2.py -- DEV CONTAINER ONLY (needs /root/reference).

Pins oracle/dinov2_oracle.py against the reference's OWN modules, imported unmodified from /root/reference:
    ISM/model/vision_transformer.py  vit_large(patch_size=14, img_size=518, init_values=1.0, block_chunks=0)  (= dinov2_vitl14)
    ISM/utils/bbox_utils.py          CropResizePad
    ISM/model/dinov2.py              CustomDINOv2.process_rgb_proposals / process_masks_proposals / compute_cls_and_patch_features
    ISM/model/loss.py                MaskedPatch_MatrixSimilarity.compute_straight / compute_visible_ratio
and writes tests/golden/dinov2.pt (seeded ViT-L/14 weights and a synthetic 6-proposal frame are regenerated from their seeds).

Usage: python tools/make_golden_dinov2.py"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import dinov2_oracle as do, ism_oracle as io  # noqa: E402
from ref_ism_import import import_reference_ism, STUBBED  # noqa: E402


def main():
    torch.set_num_threads(8)
    loss, detector = import_reference_ism()
    from model import vision_transformer as vits, dinov2 as rdino           # the reference's modules
    from utils.bbox_utils import CropResizePad
    print("stubbed third-party imports:", STUBBED)
    sd = do.make_state_dict(seed=1)
    ref = vits.vit_large(patch_size=14, img_size=518, init_values=1.0, ffn_layer="mlp", block_chunks=0, num_register_tokens=0,
                         interpolate_antialias=False, interpolate_offset=0.1).eval()
    print("reference vit_large accepted the oracle state_dict (strict):", ref.load_state_dict(sd, strict=True))
    image, masks, boxes = do.make_proposals(P=6, seed=1)

    # CustomDINOv2 without its constructor (which needs a checkpoint file): the methods read these attributes only
    host = types.SimpleNamespace(model=ref, rgb_normalize=None, rgb_proposal_processor=CropResizePad(224),
                                 patch_kernel=torch.nn.AvgPool2d(kernel_size=14, stride=14), validpatch_thresh=0.5, chunk_size=16)
    import torchvision.transforms as T
    host.rgb_normalize = T.Compose([T.ToTensor(), T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))])
    D = rdino.CustomDINOv2
    with torch.no_grad():
        r_rgbs = D.process_rgb_proposals(host, image.numpy(), masks.clone(), boxes)
        r_masks = D.process_masks_proposals(host, masks.clone(), boxes)
        r_cls, r_patch = D.compute_cls_and_patch_features(host, r_rgbs, r_masks)
        o_rgbs = do.process_rgb_proposals(image, masks.clone(), boxes)
        o_masks = do.process_masks_proposals(masks.clone(), boxes)
        o_cls, o_patch, o_keep = do.cls_and_patch_features(sd, o_rgbs, o_masks)
    for name, a, b in (("processed rgbs", r_rgbs, o_rgbs), ("processed masks", r_masks, o_masks), ("cls tokens", r_cls, o_cls),
                       ("masked patch tokens", r_patch, o_patch)):
        d = (a - b).abs().max().item()
        print(f"  {name:22s} max|ref - oracle| = {d:.3e}")
        assert d == 0.0, name
    # appearance score / visible ratio: the query patches against the patches of another proposal standing in for the best template
    ref_patch = r_patch.roll(1, dims=0)
    m = loss.MaskedPatch_MatrixSimilarity(metric="cosine", chunk_size=64)
    r_appe = m.compute_straight(r_patch, ref_patch)
    r_vis = m.compute_visible_ratio(r_patch, ref_patch, 0.5)
    assert torch.equal(io.appearance_score(o_patch, o_patch.roll(1, dims=0)), r_appe)
    assert torch.equal(io.visible_ratio(o_patch, o_patch.roll(1, dims=0), 0.5), r_vis)
    print("  appearance score / visible ratio restatements == reference")
    gold = dict(meta=dict(source="ISM/model/vision_transformer.py vit_large + ISM/model/dinov2.py + ISM/utils/bbox_utils.py + ISM/model/loss.py "
                                 "imported from /root/reference (CPU, fp32)", torch=torch.__version__, seed=1, P=6, stubbed_imports=list(STUBBED)),
                boxes=boxes, rgbs_sum=r_rgbs.double().sum(dim=(1, 2, 3)), rgbs_sub=r_rgbs[:, :, ::7, ::7].clone(),
                pmasks_packed=(r_masks > 0.5).to(torch.uint8), cls=r_cls.clone(), patch_sub=r_patch[:, ::5, :].clone(), keep=o_keep,
                appe=r_appe, vis=r_vis, cls_abs_mean=r_cls.abs().mean().item(),
                input_checksum=dict(image=image.double().sum().item(), masks=masks.double().sum().item()))
    path = os.path.join(ROOT, "tests", "golden", "dinov2.pt")
    torch.save(gold, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
