"""An independent check of the 256-case marching-cubes table the CUDA kernels read (monoport_b200/csrc/mc_table.inc).
The oracle (oracle/spec.py:marching_cubes_ref) and the kernels share the generator tools/gen_mc_table.py, so a wrong table
would pass the bit-exact comparisons.  Nothing here imports that generator: the table is parsed from the .inc file and
checked against properties that follow from the geometry alone --
  * every triangle corner is a cut edge (its two cell corners differ in the inside flag), every cut edge is used;
  * inside a cell the patch is an oriented manifold: a directed triangle side that does not lie in a cell face is matched
    by exactly one reversed side; a side lying in a cell face occurs once (the patch boundary lies on the faces);
  * watertightness across cells: for every pair of cases that agree on a shared face (all 3 axes, all 256 x 16 compatible
    pairs) the directed sides the two cells leave on that face cancel exactly;
  * orientation: stepping from a triangle's centroid along its normal lowers the trilinear interpolant of the corner flags
    (normals point from inside to outside)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# geometry conventions stated in mcubes_kernels.cuh (case_of / mesh_emit_kernel), restated here by hand
CORNER = [(c & 1, (c >> 1) & 1, (c >> 2) & 1) for c in range(8)]           # (x, y, z) of corner c


def edge_ends(e):
    """edge id -> (corner a, corner b): edges 0-3 along x at (y,z) = (q&1, q>>1), 4-7 along y at (x,z), 8-11 along z at (x,y)."""
    axis, q = e >> 2, e & 3
    lo = [0, 0, 0]
    others = [a for a in range(3) if a != axis]
    lo[others[0]], lo[others[1]] = q & 1, q >> 1
    hi = list(lo)
    hi[axis] = 1
    idx = lambda p: p[0] | (p[1] << 1) | (p[2] << 2)
    return idx(lo), idx(hi)


def load_table():
    txt = open(os.path.join(ROOT, "monoport_b200", "csrc", "mc_table.inc")).read()
    body = txt[txt.index("g_mc_tri[256][16]"):]
    rows = re.findall(r"\{([-0-9, ]+)\}", body)
    assert len(rows) == 256
    tab = np.array([[int(v) for v in r.split(",")] for r in rows], dtype=np.int64)
    assert tab.shape == (256, 16)
    return tab


def tris_of(tab, k):
    n = int(tab[k, 15])
    t = tab[k, :3 * n].reshape(n, 3)
    assert (t >= 0).all() and (tab[k, 3 * n:15] == -1).all(), k
    return t


def mid(e):
    a, b = edge_ends(e)
    return (np.array(CORNER[a], float) + np.array(CORNER[b], float)) / 2


def on_face(e1, e2):
    """(axis, side) of the cell face containing both edge midpoints, or None."""
    m1, m2 = mid(e1), mid(e2)
    for ax in range(3):
        for side in (0.0, 1.0):
            if m1[ax] == side and m2[ax] == side:
                return ax, int(side)
    return None


def test_cases_are_oriented_manifold_patches():
    tab = load_table()
    assert tab[0, 15] == 0 and tab[255, 15] == 0
    for k in range(256):
        inside = [(k >> c) & 1 for c in range(8)]
        cut = {e for e in range(12) if inside[edge_ends(e)[0]] != inside[edge_ends(e)[1]]}
        t = tris_of(tab, k)
        assert set(t.reshape(-1).tolist()) == cut, k                       # vertices exactly on the sign-changing edges
        sides = {}
        for a, b, c in t.tolist():
            assert len({a, b, c}) == 3, k
            for s in ((a, b), (b, c), (c, a)):
                sides[s] = sides.get(s, 0) + 1
        for (a, b), cnt in sides.items():
            assert cnt == 1, (k, a, b)                                      # no directed side twice
            if on_face(a, b) is None:
                assert sides.get((b, a), 0) == 1, (k, a, b)                 # interior side: matched by its reverse
            else:
                assert (b, a) not in sides or on_face(a, b) is not None


def face_sides(tab, k, axis, side):
    """Directed patch-boundary sides of case k lying in face (axis, side), as pairs of face-local edge keys."""
    out = []
    tris = tris_of(tab, k).tolist()
    all_sides = {s for a, b, c in tris for s in ((a, b), (b, c), (c, a))}
    for a, b, c in tris:
        for s in ((a, b), (b, c), (c, a)):
            f = on_face(*s)
            # a side lying in the face is part of the patch BOUNDARY unless its reverse is present too (a fan triangle may lie
            # flat in an ambiguous face: the chord it shares with the next fan triangle is interior although it is in the plane)
            if f == (axis, side) and (s[1], s[0]) not in all_sides:
                key = []
                for e in s:
                    m = mid(e)
                    key.append(tuple(np.delete(m, axis)))
                out.append(tuple(key))
    return out


def test_shared_faces_cancel_for_every_compatible_pair():
    tab = load_table()
    sides_cache = {(k, ax, sd): face_sides(tab, k, ax, sd) for k in range(256) for ax in range(3) for sd in (0, 1)}
    for ax in range(3):
        hi = [c for c in range(8) if CORNER[c][ax] == 1]
        lo = [c for c in range(8) if CORNER[c][ax] == 0]
        # corner on the + face of cell A <-> the corner of cell B (its neighbour along +axis) with the same other coordinates
        pair = {a: next(b for b in lo if all(CORNER[a][o] == CORNER[b][o] for o in range(3) if o != ax)) for a in hi}
        for ka in range(256):
            flags = {pair[a]: (ka >> a) & 1 for a in hi}
            base = sum(v << c for c, v in flags.items())
            for rest in range(16):                                          # the four far corners of B are free
                kb = base | sum(((rest >> j) & 1) << c for j, c in enumerate(hi))
                sa = sides_cache[(ka, ax, 1)]
                sb = sides_cache[(kb, ax, 0)]
                assert sorted(sa) == sorted((q, p) for p, q in sb), (ax, ka, kb)


def trilinear(flags, p):
    x, y, z = p
    v = 0.0
    for c in range(8):
        cx, cy, cz = CORNER[c]
        v += flags[c] * (x if cx else 1 - x) * (y if cy else 1 - y) * (z if cz else 1 - z)
    return v


def test_normals_point_from_inside_to_outside():
    tab = load_table()
    for k in range(1, 255):
        flags = [float((k >> c) & 1) for c in range(8)]
        for a, b, c in tris_of(tab, k).tolist():
            pa, pb, pc = mid(a), mid(b), mid(c)
            n = np.cross(pb - pa, pc - pa)
            assert np.linalg.norm(n) > 0, k
            n = n / np.linalg.norm(n)
            ctr = (pa + pb + pc) / 3
            eps = 0.04
            assert trilinear(flags, ctr + eps * n) < trilinear(flags, ctr - eps * n), (k, a, b, c)
