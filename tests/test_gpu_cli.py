"""GPU: the run_inference_custom.py drop-in (SURVEY.md 8b CLI row, BASELINE config #1) on the repository's example frame:
rgb / depth / camera.json from tests/golden/pem_input.pt, a CAD PLY, templates in the reference's on-disk format (written by the
point-splat stand-in for the BlenderProc renderer), an ISM detection JSON -> detection_pem.json in the reference's record format
(PEM/run_inference_custom.py:301-307).  No checkpoint ships (no network): seeded random weights, so poses are checked for
validity and determinism, not accuracy -- accuracy of every stage is covered by the parity tests."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_ply(path, verts_mm, faces, colors):
    with open(path, "w") as fh:
        fh.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
                 "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face %d\nproperty list uchar int vertex_indices\nend_header\n"
                 % (len(verts_mm), len(faces)))
        for v, c in zip(verts_mm, colors):
            fh.write("%f %f %f %d %d %d\n" % (v[0], v[1], v[2], c[0], c[1], c[2]))
        for f in faces:
            fh.write("3 %d %d %d\n" % tuple(f))


def test_pem_cli_end_to_end(tmp_path, golden_dir):
    import cv2
    from scipy.spatial import ConvexHull
    from sam6d_b200 import meshio
    from sam6d_b200.cli import pem_run_inference_custom as cli, render_point_templates as rpt
    gold = torch.load(os.path.join(golden_dir, "pem_input.pt"), weights_only=False)
    out = str(tmp_path)
    cv2.imwrite(os.path.join(out, "rgb.png"), gold["rgb"].numpy()[:, :, ::-1])
    cv2.imwrite(os.path.join(out, "depth.png"), gold["depth"].numpy().astype(np.uint16))
    json.dump(dict(cam_K=gold["cam_K"], depth_scale=gold["depth_scale"]), open(os.path.join(out, "camera.json"), "w"))
    json.dump(gold["dets"], open(os.path.join(out, "detection_ism.json"), "w"))
    pts_mm = gold["model_points"].numpy().astype(np.float64) * 1000.0               # CAD = convex hull of the example object's samples
    hull = ConvexHull(pts_mm)
    verts = pts_mm[hull.vertices]
    remap = {v: i for i, v in enumerate(hull.vertices)}
    faces = np.array([[remap[a] for a in s] for s in hull.simplices])
    colors = np.random.RandomState(0).randint(40, 255, (len(verts), 3))
    cad = os.path.join(out, "obj.ply")
    _write_ply(cad, verts, faces, colors)
    v2, f2, c2 = meshio.load_ply(cad)
    assert v2.shape == verts.shape and f2.shape == faces.shape and c2.shape == colors.shape
    rpt.main(["--cad_path", cad, "--output_dir", out, "--size", "192"])
    assert os.path.exists(os.path.join(out, "templates", "xyz_41.npy"))
    argv = ["--output_dir", out, "--cad_path", cad, "--rgb_path", os.path.join(out, "rgb.png"), "--depth_path", os.path.join(out, "depth.png"),
            "--cam_path", os.path.join(out, "camera.json"), "--seg_path", os.path.join(out, "detection_ism.json"), "--random_weights"]
    np.random.seed(0)
    assert cli.main(argv) == 0
    res = json.load(open(os.path.join(out, "sam6d_results", "detection_pem.json")))
    assert len(res) == 5                                                             # 6 detections, one under the score threshold
    for r in res:
        assert set(["scene_id", "image_id", "category_id", "bbox", "score", "time", "segmentation", "R", "t"]) <= set(r)
        R, t = np.array(r["R"]), np.array(r["t"])
        assert R.shape == (3, 3) and t.shape == (3,) and np.isfinite(R).all() and np.isfinite(t).all()
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-4) and abs(np.linalg.det(R) - 1) < 1e-4
        assert 0.0 <= r["score"] <= 1.0
    assert os.path.exists(os.path.join(out, "sam6d_results", "vis_pem.png"))
    # same numpy seed -> same sampled points -> same record (the whole path is deterministic)
    np.random.seed(0)
    assert cli.main(argv) == 0
    res2 = json.load(open(os.path.join(out, "sam6d_results", "detection_pem.json")))
    assert [r["R"] for r in res] == [r["R"] for r in res2]


def test_ism_cli_then_pem_cli(tmp_path, golden_dir):
    """BASELINE config #1 as the reference's demo.sh chains it: templates -> ISM CLI (SAM automatic mask generation, DINOv2
    descriptors, semantic / appearance / geometric scores -> detection_ism.json) -> PEM CLI consuming that file.  Seeded random
    weights (no checkpoints ship): record format, validity and determinism are checked, not detection quality."""
    import cv2
    from scipy.spatial import ConvexHull
    from sam6d_b200.cli import ism_run_inference_custom as ism_cli, pem_run_inference_custom as pem_cli, render_point_templates as rpt
    gold = torch.load(os.path.join(golden_dir, "pem_input.pt"), weights_only=False)
    out = str(tmp_path)
    cv2.imwrite(os.path.join(out, "rgb.png"), gold["rgb"].numpy()[:, :, ::-1])
    cv2.imwrite(os.path.join(out, "depth.png"), gold["depth"].numpy().astype(np.uint16))
    json.dump(dict(cam_K=gold["cam_K"], depth_scale=gold["depth_scale"]), open(os.path.join(out, "camera.json"), "w"))
    pts_mm = gold["model_points"].numpy().astype(np.float64) * 1000.0
    hull = ConvexHull(pts_mm)
    remap = {v: i for i, v in enumerate(hull.vertices)}
    cad = os.path.join(out, "obj.ply")
    _write_ply(cad, pts_mm[hull.vertices], np.array([[remap[a] for a in s] for s in hull.simplices]),
               np.random.RandomState(0).randint(40, 255, (len(hull.vertices), 3)))
    rpt.main(["--cad_path", cad, "--output_dir", out, "--size", "192"])
    assert os.path.exists(os.path.join(out, "templates", "template_poses.npy"))
    common = ["--output_dir", out, "--cad_path", cad, "--rgb_path", os.path.join(out, "rgb.png"), "--depth_path", os.path.join(out, "depth.png"),
              "--cam_path", os.path.join(out, "camera.json")]
    assert ism_cli.main(common + ["--random_weights", "--stability_score_thresh", "0.0", "--pred_iou_thresh", "-10", "--confidence_thresh", "-1",
                                  "--points_per_side", "8"]) == 0
    dets = json.load(open(os.path.join(out, "sam6d_results", "detection_ism.json")))
    print(f"ISM CLI: {len(dets)} detections")
    for d in dets:
        assert set(["scene_id", "image_id", "category_id", "bbox", "score", "time", "segmentation"]) <= set(d)
        assert d["segmentation"]["size"] == [480, 640] and sum(d["segmentation"]["counts"]) == 480 * 640
        assert len(d["bbox"]) == 4 and np.isfinite(d["score"])
    assert len(dets) >= 1
    np.random.seed(0)
    assert pem_cli.main(common + ["--seg_path", os.path.join(out, "sam6d_results", "detection_ism.json"), "--random_weights",
                                  "--det_score_thresh", "-1"]) == 0
    res = json.load(open(os.path.join(out, "sam6d_results", "detection_pem.json")))
    assert len(res) <= len(dets)
    for r in res:
        R = np.array(r["R"])
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-4) and np.isfinite(np.array(r["t"])).all()
