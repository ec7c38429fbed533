"""CPU: oracle/input_oracle.py (the PEM input builder, PEM/run_inference_custom.py:165-253) against tests/golden/pem_input.pt
-- produced by tools/make_golden_input.py where the restatement is pinned against the reference's own data_utils.py on the
repository's example frame -- and the fixed-point emulation of cv2.resize(INTER_LINEAR) the CUDA kernel implements."""
import os

import numpy as np
import pytest
import torch

from oracle import input_oracle as io


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "pem_input.pt"), weights_only=False)


def test_input_oracle_reproduces_fixture(gold):
    rgb, depth = gold["rgb"].numpy(), gold["depth"].numpy().astype(np.uint16)
    out, _, radius = io.get_test_data(gold["dets"], rgb, depth, gold["cam_K"], gold["depth_scale"], gold["model_points"].numpy(), 0.2,
                                      choose_idx=[c.numpy() for c in gold["choose_idx"]])
    assert out["det_index"] == gold["det_index"] and out["bbox"] == gold["bbox"] and out["n_valid"] == gold["n_valid"]
    assert abs(float(radius) - gold["radius"]) < 1e-7
    assert np.array_equal(np.stack(out["pts"]), gold["pts"].numpy())
    assert np.array_equal(np.stack(out["rgb_choose"]), gold["rgb_choose"].numpy())
    for k, u8 in enumerate(gold["rgb_u8"].numpy()):
        assert np.array_equal(io.rgb_transform(u8), out["rgb"][k])
    # edge cases the fixture covers: a detection under the score threshold, one with 35 pixels (> 32: kept, 30 survive the radius
    # filter -> sampled with replacement), a crop clipped at the image border, a mask with holes
    assert len(gold["dets"]) == 6 and len(out["pts"]) == 5 and min(out["n_valid"]) < 2048 < max(out["n_valid"])


def _emulate_resize(src, S):
    """OpenCV's uint8 INTER_LINEAR as csrc/inputs.cu implements it"""
    h, w, _ = src.shape

    def coefs(n_src, n_dst, horizontal):
        scale = 1.0 / (n_dst / n_src)
        i0 = np.zeros(n_dst, np.int64); i1 = np.zeros(n_dst, np.int64); a0 = np.zeros(n_dst, np.int64); a1 = np.zeros(n_dst, np.int64)
        for d in range(n_dst):
            f = np.float32((d + 0.5) * scale - 0.5)
            s = int(np.floor(f))
            f = np.float32(f - np.float32(s))
            if horizontal:
                if s < 0:
                    f, s = np.float32(0), 0
                if s >= n_src - 1:
                    f, s = np.float32(0), n_src - 1
            i0[d], i1[d] = min(max(s, 0), n_src - 1), min(max(s + 1, 0), n_src - 1)
            a0[d] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(2048)))
            a1[d] = int(np.rint(f * np.float32(2048)))
        return i0, i1, a0, a1
    s = src.astype(np.int64)
    if h == 2 * S and w == 2 * S:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, xa0, xa1 = coefs(w, S, True)
    y0, y1, ya0, ya1 = coefs(h, S, False)
    hor = s[:, x0, :] * xa0[None, :, None] + s[:, x1, :] * xa1[None, :, None]
    out = (((ya0[:, None, None] * (hor[y0] >> 4)) >> 16) + ((ya1[:, None, None] * (hor[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("n", [6, 37, 78, 112, 223, 224, 225, 298, 447, 448, 449, 480])
def test_cv2_linear_resize_emulation_is_bit_exact(n):
    cv2 = pytest.importorskip("cv2")
    src = np.random.RandomState(n).randint(0, 256, (n, n, 3)).astype(np.uint8)
    assert np.array_equal(cv2.resize(src, (224, 224), interpolation=cv2.INTER_LINEAR), _emulate_resize(src, 224))


def test_rle_round_trip_and_bbox_edges():
    g = np.random.RandomState(0)
    for shape in [(480, 640), (7, 5), (1, 9)]:
        m = g.rand(*shape) > 0.6
        assert np.array_equal(io.rle_to_binary_mask(io.mask_to_rle(m)), m)
    m = np.zeros((480, 640), bool)
    m[0:5, 630:640] = True                        # corner: the square box is shifted back into the image
    y1, y2, x1, x2 = io.get_bbox(m)
    assert 0 <= y1 < y2 <= 480 and 0 <= x1 < x2 <= 640 and (y2 - y1) == (x2 - x1)
    m[:] = True                                    # full frame: side capped at min(H, W)
    y1, y2, x1, x2 = io.get_bbox(m)
    assert (y2 - y1) == 480 and (x2 - x1) == 480
