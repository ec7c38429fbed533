"""CAD input of the CLIs: PLY reading and area-weighted surface sampling -- what the reference gets from
`trimesh.load_mesh(cad_path).sample(n)` (PEM/run_inference_custom.py:182-183; trimesh is not a dependency here).
Host-side numpy; ASCII and binary_little_endian PLY with `vertex` (x, y, z first) and `face` (vertex_indices list) elements."""
import numpy as np

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4", "double": "f8",
              "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def load_ply(path):
    """-> (vertices (V,3) float32, faces (F,3) int64 or empty, vertex colours (V,3) uint8 or None)"""
    with open(path, "rb") as fh:
        fmt, elements = None, []
        while True:
            line = fh.readline()
            if not line:
                raise ValueError("PLY header not terminated")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append(dict(name=tok[1], count=int(tok[2]), props=[]))
            elif tok[0] == "property":
                elements[-1]["props"].append(tok[1:])
            elif tok[0] == "end_header":
                break
        verts, faces, colors = None, np.zeros((0, 3), np.int64), None
        for el in elements:
            names = [p[-1] for p in el["props"]]
            if el["name"] == "vertex":
                if fmt == "ascii":
                    data = np.loadtxt(fh, max_rows=el["count"], dtype=np.float64, ndmin=2)
                    cols = {n: data[:, i] for i, n in enumerate(names)}
                else:
                    dt = np.dtype([(p[-1], ("<" if fmt == "binary_little_endian" else ">") + _PLY_TYPES[p[0]]) for p in el["props"]])
                    rec = np.frombuffer(fh.read(dt.itemsize * el["count"]), dtype=dt)
                    cols = {n: rec[n] for n in names}
                verts = np.stack([cols["x"], cols["y"], cols["z"]], axis=1).astype(np.float32)
                if all(c in cols for c in ("red", "green", "blue")):
                    colors = np.stack([cols["red"], cols["green"], cols["blue"]], axis=1).astype(np.uint8)
            elif el["name"] == "face":
                if fmt == "ascii":
                    rows = [fh.readline().split() for _ in range(el["count"])]
                    tris = []
                    for r in rows:
                        n = int(r[0])
                        idx = [int(v) for v in r[1:1 + n]]
                        tris += [[idx[0], idx[k], idx[k + 1]] for k in range(1, n - 1)]       # fan triangulation
                    faces = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
                else:
                    end = "<" if fmt == "binary_little_endian" else ">"
                    ct, it = _PLY_TYPES[el["props"][0][1]], _PLY_TYPES[el["props"][0][2]]
                    tris = []
                    for _ in range(el["count"]):
                        n = int(np.frombuffer(fh.read(np.dtype(ct).itemsize), dtype=end + ct)[0])
                        idx = np.frombuffer(fh.read(np.dtype(it).itemsize * n), dtype=end + it).astype(np.int64)
                        tris += [[idx[0], idx[k], idx[k + 1]] for k in range(1, n - 1)]
                    faces = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
            else:
                if fmt == "ascii":
                    for _ in range(el["count"]):
                        fh.readline()
                else:
                    raise ValueError(f"binary PLY element {el['name']} is not supported")
    if verts is None:
        raise ValueError("PLY without a vertex element")
    return verts, faces, colors


def sample_surface(verts, faces, n, rng=None):
    """area-weighted random points on the triangles (trimesh.sample.sample_surface semantics); vertices when there are no faces"""
    rng = rng if rng is not None else np.random
    if len(faces) == 0:
        return verts[rng.choice(len(verts), n, replace=len(verts) < n)].astype(np.float32)
    a, b, c = verts[faces[:, 0]].astype(np.float64), verts[faces[:, 1]].astype(np.float64), verts[faces[:, 2]].astype(np.float64)
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    f = np.searchsorted(np.cumsum(area), rng.random_sample(n) * area.sum())
    f = np.minimum(f, len(faces) - 1)
    u = rng.random_sample((n, 2))
    flip = u.sum(axis=1) > 1.0
    u[flip] = 1.0 - u[flip]
    return (a[f] + u[:, :1] * (b[f] - a[f]) + u[:, 1:] * (c[f] - a[f])).astype(np.float32)
