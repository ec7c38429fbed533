"""GPU: the PEM input builder kernels (csrc/inputs.cu, sam6d_b200/inputs.py) against tests/golden/pem_input.pt -- the
reference's get_test_data loop (PEM/run_inference_custom.py:165-253) evaluated by the pinned oracle on the repository's example
frame (BASELINE config #1 data).  Integer outputs (kept detections, bounding boxes, valid counts, rgb_choose, the uint8 crops)
are bit-exact; points agree to float32 rounding."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "pem_input.pt"), weights_only=False)


def test_input_builder_matches_reference_loop(gold):
    from sam6d_b200 import inputs
    rgb, depth = gold["rgb"].numpy(), gold["depth"].numpy().astype(np.uint16)
    dets = [d for d in gold["dets"] if d["score"] > 0.2]
    frame = inputs.FrameInputs(dets, rgb, depth, gold["cam_K"], gold["depth_scale"], gold["radius"])
    keep = frame.kept()
    assert keep.tolist() == gold["det_index"]
    assert frame.bbox()[keep].tolist() == gold["bbox"]
    assert frame.n_valid()[keep].tolist() == gold["n_valid"]
    choose_idx = np.stack([c.numpy() for c in gold["choose_idx"]])
    pts, rgb_choose, rgb_out, u8 = frame.sample(keep, choose_idx, 224, True, want_u8=True)
    assert torch.equal(rgb_choose.cpu(), gold["rgb_choose"])
    assert torch.equal(u8.cpu(), gold["rgb_u8"])                       # cv2 INTER_LINEAR fixed point, bit for bit
    torch.testing.assert_close(pts.cpu(), gold["pts"], atol=0, rtol=2e-7)
    from oracle import input_oracle as io
    want = np.stack([io.rgb_transform(x) for x in gold["rgb_u8"].numpy()])
    torch.testing.assert_close(rgb_out.cpu(), torch.from_numpy(want), atol=1e-6, rtol=0)


def test_get_test_data_signature_and_edge_cases(gold):
    """the drop-in call: same keys / shapes as the reference's ret_dict; an empty detection list and a frame without any kept
    detection return empty batches; sampling with numpy's RNG is reproducible from a RandomState"""
    from sam6d_b200 import inputs
    rgb, depth = gold["rgb"].numpy(), gold["depth"].numpy().astype(np.uint16)
    args = (rgb, depth, gold["cam_K"], gold["depth_scale"], gold["model_points"].numpy())
    ret, img, whole_pts, mp, kept = inputs.get_test_data(gold["dets"], *args, rng=np.random.RandomState(3))
    assert ret["pts"].shape == (5, 2048, 3) and ret["rgb"].shape == (5, 3, 224, 224) and ret["rgb_choose"].shape == (5, 2048)
    assert ret["rgb_choose"].dtype == torch.int64 and ret["model"].shape == (5, 1024, 3) and ret["K"].shape == (5, 3, 3)
    assert whole_pts.shape == (480 * 640, 3) and len(kept) == 5
    assert int(ret["rgb_choose"].min()) >= 0 and int(ret["rgb_choose"].max()) < 224 * 224
    ret2 = inputs.get_test_data(gold["dets"], *args, rng=np.random.RandomState(3))[0]
    assert torch.equal(ret["pts"], ret2["pts"]) and torch.equal(ret["rgb_choose"], ret2["rgb_choose"])
    # every sampled point belongs to its detection's mask and lies within 1.2 x radius of the cloud centre
    ret0, _, _, _, kept0 = inputs.get_test_data([], *args)
    assert ret0["pts"].shape == (0, 2048, 3) and kept0 == []
    only_small = [d for d in gold["dets"] if d is gold["dets"][5]]
    only_small[0] = dict(only_small[0], segmentation=dict(only_small[0]["segmentation"]))
    from oracle import input_oracle as io
    tiny = np.zeros((480, 640), bool)
    tiny[100:103, 100:103] = True                                         # 9 pixels: skipped (np.sum(mask) > 32 fails)
    only_small[0]["segmentation"] = io.mask_to_rle(tiny)
    ret1 = inputs.get_test_data(only_small, *args)[0]
    assert ret1["pts"].shape[0] == 0


def test_input_builder_many_detections(gold):
    """200 detections of one frame (BASELINE config #5 proposal count) in one pair of launches agree with the per-detection oracle"""
    from sam6d_b200 import inputs
    from oracle import input_oracle as io
    rgb, depth = gold["rgb"].numpy(), gold["depth"].numpy().astype(np.uint16)
    g = np.random.RandomState(5)
    yy, xx = np.mgrid[0:480, 0:640]
    dets = []
    for i in range(200):
        cy, cx, ry, rx = g.randint(40, 440), g.randint(40, 600), g.randint(8, 70), g.randint(8, 70)
        m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
        dets.append(dict(score=0.5, segmentation=io.mask_to_rle(m)))
    frame = inputs.FrameInputs(dets, rgb, depth, gold["cam_K"], gold["depth_scale"], gold["radius"])
    whole_depth = depth.astype(np.float32) * gold["depth_scale"] / 1000.0
    whole_pts = io.get_point_cloud_from_depth(whole_depth, np.array(gold["cam_K"]).reshape(3, 3))
    keep = frame.kept()
    ref_keep, ref_n, ref_bbox = [], [], []
    for i, d in enumerate(dets):
        r = io.build_instance(d["segmentation"], whole_depth, whole_pts, rgb, np.float32(gold["radius"]), choose_idx=np.zeros(4, np.int64))
        if r is not None:
            ref_keep.append(i); ref_n.append(r["n_valid"]); ref_bbox.append(r["bbox"])
    assert keep.tolist() == ref_keep
    assert frame.n_valid()[keep].tolist() == ref_n
    assert frame.bbox()[keep].tolist() == ref_bbox
