"""Mesh dump helpers with the reference's on-disk format (monoport/lib/mesh_util.py:223-242): text OBJ, `%.4f`
coordinates, 1-based faces, optional per-vertex colour appended to the `v` line.  Used to dump `reconstruction()`
output for parity inspection; vectorised (one formatted write) instead of a Python loop per vertex."""
import numpy as np


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def save_obj_mesh(mesh_path, verts, faces):
    verts, faces = _np(verts), _np(faces)
    with open(mesh_path, "w") as f:
        if len(verts):
            np.savetxt(f, verts[:, :3], fmt="v %.4f %.4f %.4f")
        if len(faces):
            np.savetxt(f, faces[:, :3].astype(np.int64) + 1, fmt="f %d %d %d")


def save_obj_mesh_with_color(mesh_path, verts, faces, colors):
    verts, faces, colors = _np(verts), _np(faces), _np(colors)
    with open(mesh_path, "w") as f:
        if len(verts):
            np.savetxt(f, np.concatenate([verts[:, :3], colors[:, :3]], 1), fmt="v %.4f %.4f %.4f %.4f %.4f %.4f")
        if len(faces):
            np.savetxt(f, faces[:, :3].astype(np.int64) + 1, fmt="f %d %d %d")


def save_ply_mesh(mesh_path, verts, faces, colors=None):
    """Binary little-endian PLY (float32 x y z [+ uchar red green blue], int32 triangle lists): the compact dump format for
    benchmark / parity runs -- a 257^3 mesh of 80 k vertices is 1.9 MB instead of 6.3 MB of `%.4f` text, and it
    round-trips the float32 coordinates exactly (the OBJ writers above keep the reference's 4 decimals)."""
    verts, faces = _np(verts).astype("<f4")[:, :3], _np(faces).astype("<i4")[:, :3]
    header = ["ply", "format binary_little_endian 1.0", "element vertex %d" % len(verts),
              "property float x", "property float y", "property float z"]
    vdtype = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    if colors is not None:
        header += ["property uchar red", "property uchar green", "property uchar blue"]
        vdtype += [("r", "u1"), ("g", "u1"), ("b", "u1")]
    header += ["element face %d" % len(faces), "property list uchar int vertex_indices", "end_header"]
    v = np.empty(len(verts), dtype=vdtype)
    v["x"], v["y"], v["z"] = verts[:, 0], verts[:, 1], verts[:, 2]
    if colors is not None:
        c = np.clip(np.rint(_np(colors)[:, :3] * 255.0), 0, 255).astype("u1")      # colours in [0,1] like the OBJ writer's
        v["r"], v["g"], v["b"] = c[:, 0], c[:, 1], c[:, 2]
    f = np.empty(len(faces), dtype=[("n", "u1"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4")])
    f["n"] = 3
    f["a"], f["b"], f["c"] = faces[:, 0], faces[:, 1], faces[:, 2]
    with open(mesh_path, "wb") as fh:
        fh.write(("\n".join(header) + "\n").encode("ascii"))
        fh.write(v.tobytes())
        fh.write(f.tobytes())


def load_ply_mesh(mesh_path):
    """Reader of save_ply_mesh's files -> (verts [V,3] float32, faces [F,3] int32, colors [V,3] uint8 or None)."""
    with open(mesh_path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    head = raw[:end].decode("ascii").split("\n")
    nv = int(next(l for l in head if l.startswith("element vertex")).split()[-1])
    nf = int(next(l for l in head if l.startswith("element face")).split()[-1])
    has_c = any("red" in l for l in head)
    vdtype = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")] + ([("r", "u1"), ("g", "u1"), ("b", "u1")] if has_c else [])
    v = np.frombuffer(raw, dtype=vdtype, count=nv, offset=end)
    f = np.frombuffer(raw, dtype=[("n", "u1"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4")], count=nf, offset=end + v.nbytes)
    verts = np.stack([v["x"], v["y"], v["z"]], 1)
    faces = np.stack([f["a"], f["b"], f["c"]], 1)
    return verts, faces, (np.stack([v["r"], v["g"], v["b"]], 1) if has_c else None)
