"""CPU: oracle/ism_oracle.py against tests/golden/ism_scoring.pt -- outputs of the reference's own PairwiseSimilarity /
compute_semantic_score / best_template_pose (tools/make_golden_ism.py)."""
import os

import pytest
import torch

from oracle import ism_oracle as io


@pytest.mark.parametrize("case", ["config5_ycbv", "config3_ism"])
def test_oracle_matches_reference_scoring(golden_dir, case):
    gold = torch.load(os.path.join(golden_dir, "ism_scoring.pt"), weights_only=False)
    c = gold["cases"][case]
    q, r = io.make_descriptors(P=c["P"], O=c["O"], T=c["T"], C=c["C"], seed=c["seed"])
    assert q.double().sum().item() == c["input_checksum"]["q"] and r.double().sum().item() == c["input_checksum"]["ref"]
    idx_sel, pred_obj, sem, best_t, scores, _ = io.compute_semantic_score(q, r, 0.2)
    assert torch.equal(idx_sel, c["idx_selected"])
    assert torch.equal(pred_obj, c["pred_idx_objects"])
    assert torch.equal(best_t, c["best_template"])                 # bit-exact argmax template indices
    torch.testing.assert_close(sem, c["semantic_score"], atol=1e-6, rtol=0)
    torch.testing.assert_close(scores, c["sim"], atol=1e-6, rtol=0)
    assert 0 < len(idx_sel) < c["P"]                               # the threshold splits the synthetic proposals


@pytest.mark.parametrize("case", ["frame12", "frame64"])
def test_oracle_matches_reference_geometric_score(golden_dir, case):
    """tests/golden/ism_geo.pt = outputs of the reference's own Calculate_the_query_translation / project_template_to_image /
    compute_geometric_score (tools/make_golden_ism_geo.py), on int32 depth and float64 intrinsics like its run_inference_custom.py"""
    gold = torch.load(os.path.join(golden_dir, "ism_geo.pt"), weights_only=False)
    c = gold["cases"][case]
    inp = io.make_geometric_inputs(**c["kw"])
    assert float(inp["masks"].double().sum() + inp["depth"].double().sum() + inp["pointcloud"].double().sum()) == c["input_checksum"]
    H, W = inp["depth"].shape
    tr = io.query_translation(inp["masks"], inp["depth"], inp["K"], inp["depth_scale"])
    vu = io.project_template_to_image(inp["poses"], inp["pointcloud"], inp["best_pose"], inp["pred_obj"], tr, inp["K"], H, W)
    xyxy, iou = io.geometric_iou(vu, inp["boxes"])
    assert torch.equal(tr, c["translate"]) and int(vu.long().sum()) == c["vu_checksum"]
    assert torch.equal(xyxy, c["xyxy"]) and torch.equal(iou, c["iou"])
    assert (iou > 0.2).all() and (iou < 1).all()
    # batch-wide rule of compute_iou: one empty intersection zeroes the whole batch
    boxes = inp["boxes"].clone()
    boxes[1] = torch.tensor([0, 0, 2, 2])
    assert io.geometric_iou(vu, boxes)[1] == 0.0
