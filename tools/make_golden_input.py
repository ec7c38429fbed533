"""tools/make_golden_input.py -- DEV CONTAINER ONLY (needs /root/reference).

Pins oracle/input_oracle.py against the reference's own PEM/utils/data_utils.py (imported unmodified; `imageio`, used only
by load_im, is stubbed) on the repository's example frame (SAM-6D/Data/Example: rgb.png, depth.png, camera.json,
obj_000005.ply -- BASELINE config #1), and writes tests/golden/pem_input.pt: the frame, six synthetic detections in the
reference's own RLE format, the sample indices, and the reference-side outputs of the input builder
(PEM/run_inference_custom.py:165-253) for each kept detection.

Usage: python tools/make_golden_input.py"""
import json
import os
import sys
import types

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/SAM-6D"

from oracle import input_oracle as io  # noqa: E402


def load_ply_vertices(path):
    with open(path, "rb") as fh:
        n = 0
        while True:
            line = fh.readline().decode("ascii", "replace").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        v = np.loadtxt(fh, max_rows=n, usecols=(0, 1, 2), dtype=np.float32)
    return v


def main():
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    sys.path.insert(0, os.path.join(REF, "Pose_Estimation_Model", "utils"))
    import data_utils as du                                                    # the reference's own module

    ex = os.path.join(REF, "Data", "Example")
    rgb = cv2.imread(os.path.join(ex, "rgb.png"), cv2.IMREAD_UNCHANGED)[:, :, ::-1].copy()      # imageio order: RGB
    depth = cv2.imread(os.path.join(ex, "depth.png"), cv2.IMREAD_UNCHANGED)
    cam = json.load(open(os.path.join(ex, "camera.json")))
    K = np.array(cam["cam_K"]).reshape(3, 3)
    verts = load_ply_vertices(os.path.join(ex, "obj_000005.ply"))
    g = np.random.RandomState(7)
    model_points = (verts[g.choice(len(verts), 1024, replace=False)] / 1000.0).astype(np.float32)
    H, W = depth.shape

    # synthetic detections (no ISM output ships with the reference): ellipse, box with a hole, border-clipped box, a large one,
    # one below the score threshold, one with fewer than 32 valid pixels (skipped by the reference)
    yy, xx = np.mgrid[0:H, 0:W]
    masks = [
        ((yy - 250) / 40.0) ** 2 + ((xx - 380) / 55.0) ** 2 < 1.0,
        (abs(yy - 200) < 45) & (abs(xx - 200) < 30) & ~((abs(yy - 200) < 10) & (abs(xx - 200) < 10)),
        (yy > 400) & (xx > 560),
        (abs(yy - 260) < 120) & (abs(xx - 320) < 150) & (((yy // 7) + (xx // 5)) % 3 != 0),
        (abs(yy - 100) < 30) & (abs(xx - 500) < 30),
        (abs(yy - 300) < 3) & (abs(xx - 100) < 4),
    ]
    scores = [0.9, 0.55, 0.7, 0.35, 0.1, 0.8]
    dets = []
    for i, (m, s) in enumerate(zip(masks, scores)):
        rle = io.mask_to_rle(m)
        assert np.array_equal(du.rle_to_binary_mask(rle), m), "RLE round trip through the reference decoder"
        assert np.array_equal(io.rle_to_binary_mask(rle), m)
        x0, x1, y0, y1 = xx[m].min(), xx[m].max(), yy[m].min(), yy[m].max()
        dets.append(dict(scene_id=0, image_id=0, category_id=5, bbox=[int(x0), int(y0), int(x1 - x0), int(y1 - y0)], score=s, time=0.0,
                         segmentation=rle))

    # pin the restated helpers against the reference's own
    whole_depth = depth.astype(np.float32) * cam["depth_scale"] / 1000.0
    assert np.array_equal(du.get_point_cloud_from_depth(whole_depth, K), io.get_point_cloud_from_depth(whole_depth, K))
    for m in masks:
        mm = np.logical_and(m, whole_depth > 0)
        bb = du.get_bbox(mm)
        assert [int(v) for v in bb] == io.get_bbox(mm), (bb, io.get_bbox(mm))
        ch = mm[bb[0]:bb[1], bb[2]:bb[3]].flatten().nonzero()[0]
        assert np.array_equal(du.get_resize_rgb_choose(ch, bb, 224), io.get_resize_rgb_choose(ch, bb, 224))
    print("oracle helpers == reference data_utils (get_point_cloud_from_depth, get_bbox, get_resize_rgb_choose, rle_to_binary_mask)")

    out, whole_pts, radius = io.get_test_data(dets, rgb, depth, cam["cam_K"], cam["depth_scale"], model_points, 0.2, seed=11)
    print(f"kept {len(out['pts'])} of {len(dets)} detections: det indices {out['det_index']}, valid points {out['n_valid']}, bboxes {out['bbox']}")
    whole_depth_m = depth.astype(np.float32) * cam["depth_scale"] / 1000.0
    rgb_u8 = []
    for k, di in enumerate(out["det_index"]):
        seg = [d for d in dets if d["score"] > 0.2][di]["segmentation"]
        y1, y2, x1, x2 = out["bbox"][k]
        m = np.logical_and(io.rle_to_binary_mask(seg) > 0, whole_depth_m > 0)[y1:y2, x1:x2]
        u8 = io.crop_resize_rgb(rgb, m, out["bbox"][k], 224, True, return_u8=True)
        assert np.array_equal(io.rgb_transform(u8), out["rgb"][k])
        rgb_u8.append(u8)
    gold = dict(
        meta=dict(source="oracle/input_oracle.py pinned against PEM/utils/data_utils.py on SAM-6D/Data/Example", numpy=np.__version__,
                  cv2=cv2.__version__, img_size=224, n_sample=2048, det_score_thresh=0.2),
        rgb=torch.from_numpy(rgb), depth=torch.from_numpy(depth.astype(np.int16)), cam_K=cam["cam_K"], depth_scale=cam["depth_scale"],
        model_points=torch.from_numpy(model_points), dets=dets, radius=float(radius),
        det_index=out["det_index"], bbox=out["bbox"], n_valid=out["n_valid"],
        choose_idx=[torch.from_numpy(np.asarray(c)) for c in out["choose_idx"]],
        pts=torch.from_numpy(np.stack(out["pts"])),
        # the resized, masked crop before ToTensor / Normalize (uint8: cv2's fixed-point INTER_LINEAR, the exactness target);
        # the normalised tensor is rgb_transform of it
        rgb_u8=torch.from_numpy(np.stack(rgb_u8)),
        rgb_choose=torch.from_numpy(np.stack(out["rgb_choose"])),
    )
    path = os.path.join(ROOT, "tests", "golden", "pem_input.pt")
    torch.save(gold, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
