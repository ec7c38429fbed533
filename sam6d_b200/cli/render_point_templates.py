"""Stand-in for SAM-6D/Render/render_custom_templates.py (BlenderProc, OUT OF SCOPE -- SURVEY.md section 2.1): writes the 42
template views a CAD model needs, in the reference's on-disk format
    $OUT/templates/rgb_i.png, mask_i.png (255 = object), xyz_i.npy (object coordinates in mm, float16),  i = 0..41
by z-buffered point splatting of coloured surface samples from the 42 viewpoints of a once-subdivided icosahedron (the CNOS
level-0 viewpoint set has the same construction).  No shading, no textures beyond vertex colours: good enough to exercise the
CLIs on the repository's example data, not a renderer.

    python -m sam6d_b200.cli.render_point_templates --cad_path obj.ply --output_dir OUT [--size 256]"""
import argparse
import os

import numpy as np

from .. import meshio


def icosphere_42():
    """12 icosahedron vertices + 30 edge midpoints, normalised"""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    d = np.linalg.norm(v[:, None] - v[None], axis=2)
    edge = np.argwhere((d > 1e-6) & (d < d[d > 1e-6].min() * 1.01))
    mids = {tuple(sorted(e)): None for e in edge.tolist()}
    m = np.array([v[a] + v[b] for a, b in mids], dtype=np.float64)
    m /= np.linalg.norm(m, axis=1, keepdims=True)
    out = np.concatenate([v, m])
    assert out.shape == (42, 3)
    return out


def look_at(cam_pos):
    """world (object) -> camera rotation with the camera at cam_pos looking at the origin, OpenCV convention (z forward, y down)"""
    z = -cam_pos / np.linalg.norm(cam_pos)
    up = np.array([0.0, 0.0, 1.0]) if abs(z[2]) < 0.95 else np.array([0.0, 1.0, 0.0])
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z])


def render_templates(verts_mm, faces, colors, size=256, n_points=400000, seed=0):
    """-> list of 42 (rgb (S,S,3) uint8, mask (S,S) uint8, xyz (S,S,3) float16)"""
    rng = np.random.RandomState(seed)
    if len(faces):
        a, b, c = verts_mm[faces[:, 0]].astype(np.float64), verts_mm[faces[:, 1]].astype(np.float64), verts_mm[faces[:, 2]].astype(np.float64)
        area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
        f = np.minimum(np.searchsorted(np.cumsum(area), rng.random_sample(n_points) * area.sum()), len(faces) - 1)
        u = rng.random_sample((n_points, 2))
        flip = u.sum(1) > 1
        u[flip] = 1 - u[flip]
        w0, w1, w2 = 1 - u[:, :1] - u[:, 1:], u[:, :1], u[:, 1:]
        pts = w0 * a[f] + w1 * b[f] + w2 * c[f]
        if colors is not None:
            col = w0 * colors[faces[f, 0]] + w1 * colors[faces[f, 1]] + w2 * colors[faces[f, 2]]
        else:
            col = np.full((n_points, 3), 160.0)
    else:
        pts, col = verts_mm.astype(np.float64), (colors if colors is not None else np.full((len(verts_mm), 3), 160)).astype(np.float64)
    radius = np.linalg.norm(pts, axis=1).max()
    focal = 2.2 * size                                        # object fills ~70 % of the frame at distance 3.2 radius
    dist = 3.2 * radius
    views = []
    poses = []
    for cam_dir in icosphere_42():
        R = look_at(cam_dir * dist)
        Tm = np.eye(4)
        Tm[:3, :3], Tm[:3, 3] = R, [0, 0, dist / 1000.0]       # object -> camera (metres), the layout of the reference's template poses
        poses.append(Tm)
        pc = pts @ R.T + np.array([0, 0, dist])
        uu = np.round(focal * pc[:, 0] / pc[:, 2] + size / 2).astype(np.int64)
        vv = np.round(focal * pc[:, 1] / pc[:, 2] + size / 2).astype(np.int64)
        ok = (uu >= 0) & (uu < size) & (vv >= 0) & (vv < size)
        lin = vv[ok] * size + uu[ok]
        z = pc[ok, 2]
        order = np.lexsort((z, lin))                          # per pixel: nearest point first
        lin_s, first = np.unique(lin[order], return_index=True)
        sel = np.flatnonzero(ok)[order][first]
        # points within 2 % of the radius behind the front surface would be visible through holes: z-buffer handles it per pixel
        rgb = np.zeros((size * size, 3), np.uint8)
        mask = np.zeros(size * size, np.uint8)
        xyz = np.zeros((size * size, 3), np.float16)
        rgb[lin_s] = np.clip(col[sel], 0, 255).astype(np.uint8)
        mask[lin_s] = 255
        xyz[lin_s] = pts[sel].astype(np.float16)
        views.append((rgb.reshape(size, size, 3), mask.reshape(size, size), xyz.reshape(size, size, 3)))
    render_templates.last_poses = np.stack(poses)
    return views


def write_templates(views, out_dir):
    import cv2
    tdir = os.path.join(out_dir, "templates")
    os.makedirs(tdir, exist_ok=True)
    for i, (rgb, mask, xyz) in enumerate(views):
        cv2.imwrite(os.path.join(tdir, f"rgb_{i}.png"), rgb[:, :, ::-1])       # files hold RGB as load_im reads it
        cv2.imwrite(os.path.join(tdir, f"mask_{i}.png"), mask)
        np.save(os.path.join(tdir, f"xyz_{i}.npy"), xyz)
    if getattr(render_templates, "last_poses", None) is not None:
        np.save(os.path.join(tdir, "template_poses.npy"), render_templates.last_poses)
    return tdir


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--cad_path", required=True)
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--size", type=int, default=256)
    a = ap.parse_args(argv)
    v, f, c = meshio.load_ply(a.cad_path)
    print("templates written to", write_templates(render_templates(v, f, c, a.size), a.output_dir))


if __name__ == "__main__":
    main()
