"""oracle/input_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + cv2, the libraries the reference itself calls) of the PEM input builder:
    get_test_data            PEM/run_inference_custom.py:165-253   (per-detection loop: mask, bbox, cloud, radius filter,
                                                                     2048-point sampling, crop / mask / resize / normalise, rgb_choose)
    _get_template            PEM/run_inference_custom.py:117-146   (template crop, 5000-point sampling)
    rle_to_binary_mask       PEM/utils/data_utils.py:73-89         (uncompressed COCO RLE, column-major)
    get_point_cloud_from_depth, get_resize_rgb_choose, get_bbox     PEM/utils/data_utils.py:92-160
    rgb_transform            PEM/run_inference_custom.py:97-99     (ToTensor + ImageNet Normalize)
Parity status: PINNED -- tools/make_golden_input.py imports the reference's data_utils.py (imageio stubbed: only load_im uses
it) and checks get_bbox / get_point_cloud_from_depth / get_resize_rgb_choose / rle_to_binary_mask of this file against it on
the example frame (SAM-6D/Data/Example) and writes tests/golden/pem_input.pt.  pycocotools (mask decode) and trimesh (CAD
sampling) are absent: detections carry the reference's own uncompressed RLE (ISM/model/utils.py:25-43 mask_to_rle), which
cocomask.frPyObjects + decode round-trips exactly; CAD samples are an input.

The reference draws its sample indices from numpy's unseeded global RNG (np.random.choice, :214-217), so index-level parity is
defined given the indices: every function here takes them as `choose_idx`."""
from typing import Dict, List, Optional

import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def mask_to_rle(binary_mask: np.ndarray) -> Dict:
    """ISM/model/utils.py:25-43 (vectorised; same counts)"""
    flat = np.asarray(binary_mask).ravel(order="F").astype(np.uint8)
    change = np.flatnonzero(np.diff(flat)) + 1
    bounds = np.concatenate([[0], change, [flat.size]])
    counts = np.diff(bounds).tolist()
    if flat.size and flat[0] == 1:
        counts = [0] + counts
    return {"counts": counts, "size": list(binary_mask.shape)}


def rle_to_binary_mask(rle: Dict) -> np.ndarray:
    """data_utils.py:73-89"""
    size = rle["size"]
    out = np.zeros(int(np.prod(size)), dtype=bool)
    counts = rle["counts"]
    start = 0
    for i in range(len(counts) - 1):
        start += counts[i]
        out[start:start + counts[i + 1]] = (i + 1) % 2
    return out.reshape(*size, order="F")


def get_point_cloud_from_depth(depth: np.ndarray, K: np.ndarray) -> np.ndarray:
    """data_utils.py:92-110 (no bbox)"""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    H, W = depth.shape
    xmap = np.tile(np.arange(W), (H, 1))
    ymap = np.tile(np.arange(H)[:, None], (1, W))
    pt2 = depth.astype(np.float32)
    pt0 = (xmap.astype(np.float32) - cx) * pt2 / fx
    pt1 = (ymap.astype(np.float32) - cy) * pt2 / fy
    return np.stack([pt0, pt1, pt2]).transpose((1, 2, 0))


def get_resize_rgb_choose(choose: np.ndarray, bbox, img_size: int) -> np.ndarray:
    """data_utils.py:113-124"""
    rmin, rmax, cmin, cmax = bbox
    crop_h, crop_w = rmax - rmin, cmax - cmin
    ratio_h, ratio_w = img_size / crop_h, img_size / crop_w
    row_idx, col_idx = choose // crop_w, choose % crop_w
    return (np.floor(row_idx * ratio_h) * img_size + np.floor(col_idx * ratio_w)).astype(np.int64)


def get_bbox(label: np.ndarray):
    """data_utils.py:127-160"""
    img_width, img_length = label.shape
    rows, cols = np.any(label, axis=1), np.any(label, axis=0)
    rmin, rmax = np.where(rows)[0][[0, -1]]
    cmin, cmax = np.where(cols)[0][[0, -1]]
    rmax += 1
    cmax += 1
    b = min(max(rmax - rmin, cmax - cmin), min(img_width, img_length))
    center = [int((rmin + rmax) / 2), int((cmin + cmax) / 2)]
    rmin, rmax = center[0] - int(b / 2), center[0] + int(b / 2)
    cmin, cmax = center[1] - int(b / 2), center[1] + int(b / 2)
    if rmin < 0:
        rmax += -rmin
        rmin = 0
    if cmin < 0:
        cmax += -cmin
        cmin = 0
    if rmax > img_width:
        rmin -= rmax - img_width
        rmax = img_width
    if cmax > img_length:
        cmin -= cmax - img_length
        cmax = img_length
    return [int(rmin), int(rmax), int(cmin), int(cmax)]


def rgb_transform(rgb_u8: np.ndarray) -> np.ndarray:
    """ToTensor + Normalize (run_inference_custom.py:97-99): (H,W,3) uint8 -> (3,H,W) float32"""
    x = rgb_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    return (x - MEAN[:, None, None]) / STD[:, None, None]


def crop_resize_rgb(image_u8: np.ndarray, mask_crop: Optional[np.ndarray], bbox, img_size: int, mask_flag: bool = True,
                    return_u8: bool = False) -> np.ndarray:
    """run_inference_custom.py:221-225: crop, channel flip, mask, cv2 INTER_LINEAR resize, normalise"""
    import cv2
    y1, y2, x1, x2 = bbox
    rgb = image_u8.copy()[y1:y2, x1:x2, :][:, :, ::-1]
    if mask_flag:
        rgb = rgb * (mask_crop[:, :, None] > 0).astype(np.uint8)
    rgb = cv2.resize(rgb, (img_size, img_size), interpolation=cv2.INTER_LINEAR)
    if return_u8:
        return np.array(rgb)
    return rgb_transform(np.array(rgb))


def build_instance(seg: Dict, whole_depth: np.ndarray, whole_pts: np.ndarray, whole_image: np.ndarray, radius: float,
                   n_sample: int = 2048, img_size: int = 224, rgb_mask_flag: bool = True, choose_idx: Optional[np.ndarray] = None,
                   rng: Optional[np.random.RandomState] = None):
    """one iteration of the detection loop (run_inference_custom.py:189-234).  Returns None where the reference `continue`s,
    else dict(bbox, n_valid, choose_idx, pts, rgb, rgb_choose)."""
    mask = rle_to_binary_mask(seg)
    mask = np.logical_and(mask > 0, whole_depth > 0)
    if np.sum(mask) > 32:
        bbox = get_bbox(mask)
        y1, y2, x1, x2 = bbox
    else:
        return None
    mask = mask[y1:y2, x1:x2]
    choose = mask.astype(np.float32).flatten().nonzero()[0]
    cloud = whole_pts.copy()[y1:y2, x1:x2, :].reshape(-1, 3)[choose, :]
    center = np.mean(cloud, axis=0)
    tmp_cloud = cloud - center[None, :]
    flag = np.linalg.norm(tmp_cloud, axis=1) < radius * 1.2
    if np.sum(flag) < 4:
        return None
    choose = choose[flag]
    cloud = cloud[flag]
    if choose_idx is None:
        rng = rng or np.random
        if len(choose) <= n_sample:
            choose_idx = rng.choice(np.arange(len(choose)), n_sample)
        else:
            choose_idx = rng.choice(np.arange(len(choose)), n_sample, replace=False)
    n_valid = len(choose)
    choose = choose[choose_idx]
    cloud = cloud[choose_idx]
    rgb = crop_resize_rgb(whole_image, mask, bbox, img_size, rgb_mask_flag)
    rgb_choose = get_resize_rgb_choose(choose, [y1, y2, x1, x2], img_size)
    return dict(bbox=bbox, n_valid=n_valid, center=center, choose_idx=np.asarray(choose_idx), pts=cloud.astype(np.float32), rgb=rgb,
                rgb_choose=rgb_choose)


def get_test_data(dets: List[Dict], whole_image: np.ndarray, depth_raw: np.ndarray, cam_K: np.ndarray, depth_scale: float,
                  model_points: np.ndarray, det_score_thresh: float = 0.2, n_sample: int = 2048, img_size: int = 224,
                  choose_idx: Optional[List[np.ndarray]] = None, seed: int = 0):
    """run_inference_custom.py:165-253 after the file reads; model_points (m) replace trimesh's mesh.sample.
    choose_idx: per kept detection sample indices (in detection order) or None (drawn from RandomState(seed))."""
    dets = [d for d in dets if d["score"] > det_score_thresh]
    K = np.array(cam_K).reshape(3, 3)
    if whole_image.ndim == 2:
        whole_image = np.concatenate([whole_image[:, :, None]] * 3, axis=2)
    whole_depth = depth_raw.astype(np.float32) * depth_scale / 1000.0
    whole_pts = get_point_cloud_from_depth(whole_depth, K)
    radius = np.max(np.linalg.norm(model_points, axis=1))
    rng = np.random.RandomState(seed)
    out = dict(pts=[], rgb=[], rgb_choose=[], score=[], dets=[], bbox=[], n_valid=[], choose_idx=[], det_index=[])
    for di, inst in enumerate(dets):
        ci = choose_idx[di] if choose_idx is not None else None
        r = build_instance(inst["segmentation"], whole_depth, whole_pts, whole_image, radius, n_sample, img_size, True, ci, rng)
        if r is None:
            continue
        out["pts"].append(r["pts"]); out["rgb"].append(r["rgb"]); out["rgb_choose"].append(r["rgb_choose"])
        out["score"].append(inst["score"]); out["dets"].append(inst); out["bbox"].append(r["bbox"]); out["n_valid"].append(r["n_valid"])
        out["choose_idx"].append(r["choose_idx"]); out["det_index"].append(di)
    return out, whole_pts, radius


def get_template(rgb_u8: np.ndarray, mask_u8: np.ndarray, xyz_mm: np.ndarray, n_sample: int = 5000, img_size: int = 224,
                 choose_idx: Optional[np.ndarray] = None, rng: Optional[np.random.RandomState] = None):
    """_get_template (run_inference_custom.py:117-146) after the file reads: rgb (H,W,3) uint8 as loaded, mask (H,W) uint8,
    xyz (H,W,3) object coordinates in mm -> (rgb (3,S,S), rgb_choose (n), xyz (n,3) in m)"""
    xyz = xyz_mm.astype(np.float32) / 1000.0
    mask = mask_u8 == 255
    bbox = get_bbox(mask)
    y1, y2, x1, x2 = bbox
    mask = mask[y1:y2, x1:x2]
    rgb = crop_resize_rgb(rgb_u8, mask, bbox, img_size, True)
    choose = (mask > 0).astype(np.float32).flatten().nonzero()[0]
    if choose_idx is None:
        rng = rng or np.random
        if len(choose) <= n_sample:
            choose_idx = rng.choice(np.arange(len(choose)), n_sample)
        else:
            choose_idx = rng.choice(np.arange(len(choose)), n_sample, replace=False)
    choose = choose[choose_idx]
    xyz = xyz[y1:y2, x1:x2, :].reshape((-1, 3))[choose, :]
    return rgb, get_resize_rgb_choose(choose, bbox, img_size), xyz, bbox, np.asarray(choose_idx)
