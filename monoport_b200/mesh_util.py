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
