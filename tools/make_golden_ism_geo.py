"""tools/make_golden_ism_geo.py -- DEV CONTAINER ONLY (needs /root/reference).

Pins the geometric-score restatement of oracle/ism_oracle.py against the reference's OWN code:
    Instance_Segmentation_Model.Calculate_the_query_translation   ISM/model/detector.py:237-250
      -> depth_image_to_pointcloud_translate_torch                ISM/utils/trimesh_utils.py:77-105
    Instance_Segmentation_Model.project_template_to_image         ISM/model/detector.py:209-235
    Instance_Segmentation_Model.compute_geometric_score (IoU part) ISM/model/detector.py:311-323 -> compute_iou, bbox_utils.py:197-221
imported unmodified from /root/reference (tools/ref_ism_import.py stubs the absent third-party imports) and called on a bare
object carrying `ref_data`, with a batch of the dtypes run_inference_custom.py:83-93 builds (int32 depth, float64 cam_K, float64
depth_scale of shape (1,)).  Writes tests/golden/ism_geo.pt.

Usage: python tools/make_golden_ism_geo.py"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import ism_oracle as io  # noqa: E402
from ref_ism_import import import_reference_ism, STUBBED  # noqa: E402


def main():
    loss, detector = import_reference_ism()
    ISMModel = detector.Instance_Segmentation_Model
    cases = {}
    for tag, kw in dict(frame12=dict(N=12, seed=0), frame64=dict(N=64, seed=1)).items():
        inp = io.make_geometric_inputs(**kw)
        H, W = inp["depth"].shape
        host = types.SimpleNamespace(ref_data={"poses": inp["poses"], "pointcloud": inp["pointcloud"]})
        host.Calculate_the_query_translation = types.MethodType(ISMModel.Calculate_the_query_translation, host)
        batch = {"depth": inp["depth"].unsqueeze(0), "cam_intrinsic": inp["K"].unsqueeze(0), "depth_scale": inp["depth_scale"]}
        with torch.no_grad():
            tr_ref = ISMModel.Calculate_the_query_translation(host, inp["masks"].clone(), batch["depth"][0], batch["cam_intrinsic"][0],
                                                              batch["depth_scale"])
            vu_ref = ISMModel.project_template_to_image(host, inp["best_pose"], inp["pred_obj"], batch, inp["masks"].clone())
            proposals = types.SimpleNamespace(boxes=inp["boxes"])
            # compute_geometric_score also returns the visible ratio (pinned separately: ism_scoring / dinov2 goldens); feed it
            # one-patch descriptors so that only the IoU part matters here
            one = torch.ones(inp["masks"].shape[0], 1, 4)
            iou_ref, _ = ISMModel.compute_geometric_score(host, vu_ref, proposals, one, one, 0.5)
        tr = io.query_translation(inp["masks"], inp["depth"], inp["K"], inp["depth_scale"])
        vu = io.project_template_to_image(inp["poses"], inp["pointcloud"], inp["best_pose"], inp["pred_obj"], tr, inp["K"], H, W)
        xyxy, iou = io.geometric_iou(vu, inp["boxes"])
        assert torch.equal(tr, tr_ref), "translation restatement differs"
        assert torch.equal(vu, vu_ref), "projection restatement differs"
        assert torch.is_tensor(iou_ref) and torch.equal(iou, iou_ref), "IoU restatement differs"
        print(f"  {tag}: N={kw['N']}: oracle == reference bit for bit (translation, pixel coordinates, IoU); IoU range "
              f"{iou.min().item():.3f} .. {iou.max().item():.3f}")
        cases[tag] = dict(kw=kw, translate=tr_ref, xyxy=xyxy, iou=iou_ref, vu_checksum=int(vu_ref.long().sum()),
                          input_checksum=float(inp["masks"].double().sum() + inp["depth"].double().sum() + inp["pointcloud"].double().sum()))
    # the batch-wide rule: one proposal whose projected box misses its proposal box zeroes the score of the whole batch
    inp = io.make_geometric_inputs(N=12, seed=0)
    inp["boxes"][3] = torch.tensor([0, 0, 2, 2])
    tr = io.query_translation(inp["masks"], inp["depth"], inp["K"], inp["depth_scale"])
    vu = io.project_template_to_image(inp["poses"], inp["pointcloud"], inp["best_pose"], inp["pred_obj"], tr, inp["K"], 480, 640)
    from utils.bbox_utils import compute_iou
    xyxy, iou = io.geometric_iou(vu, inp["boxes"])
    assert compute_iou(xyxy, inp["boxes"]) == 0.0 and iou == 0.0
    out = os.path.join(ROOT, "tests", "golden", "ism_geo.pt")
    torch.save(dict(meta=dict(source="ISM/model/detector.py project_template_to_image / Calculate_the_query_translation / "
                              "compute_geometric_score + utils/trimesh_utils.py + utils/bbox_utils.py imported from /root/reference "
                              "(CPU)", torch=torch.__version__, stubbed_imports=list(STUBBED)), cases=cases), out)
    print(f"wrote {out} ({os.path.getsize(out) / 1e3:.1f} KB)")


if __name__ == "__main__":
    main()
