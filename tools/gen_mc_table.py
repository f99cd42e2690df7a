#!/usr/bin/env python
"""Generate the 256-case marching-cubes triangle table used by the CUDA kernel and the CPU oracle.

No marching-cubes implementation exists in the reference (SURVEY.md finding 3) and no copy of the
classic Lorensen-Cline table is available offline, so the table is *derived*, not transcribed:

  * corner i of a cell sits at offset (i&1, (i>>1)&1, (i>>2)&1)  (x fastest);
  * case index = sum_i [v_i > iso] << i  ("inside" = occupied, v > iso, same strict test as
    RTL/recon.py:56,60);
  * edges 0-3 run along x, 4-7 along y, 8-11 along z (see EDGE_CORNERS);
  * on every cell face the iso-contour segments are fixed by that face's four corner flags alone;
    the ambiguous face (two diagonally opposite inside corners) always *isolates the inside
    corners*.  Because the rule only looks at the face, the two cells sharing a face agree, hence
    the extracted surface is watertight (the classic table + complement cases is not);
  * segments are chained into closed directed loops, each loop is fan-triangulated from its
    smallest edge id.  Orientation: normals point from inside (occupied) to outside.

Run as a script to (re)write monoport_b200/csrc/mc_table.inc.
"""
import os
import numpy as np

CORNER_OFF = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], dtype=np.int64)

# edge -> (corner a, corner b); a is always the lower corner along the edge axis
EDGE_CORNERS = [
    (0, 1), (2, 3), (4, 5), (6, 7),      # x edges at (y,z) = (0,0) (1,0) (0,1) (1,1)
    (0, 2), (1, 3), (4, 6), (5, 7),      # y edges at (x,z) = (0,0) (1,0) (0,1) (1,1)
    (0, 4), (1, 5), (2, 6), (3, 7),      # z edges at (x,y) = (0,0) (1,0) (0,1) (1,1)
]
EDGE_AXIS = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2]
_EDGE_OF = {}
for _e, (_a, _b) in enumerate(EDGE_CORNERS):
    _EDGE_OF[(_a, _b)] = _e
    _EDGE_OF[(_b, _a)] = _e


def _faces():
    """Six faces, each as 4 corner ids in counter-clockwise order seen from OUTSIDE the cube."""
    faces = []
    for axis in range(3):
        for side in (0, 1):
            u, v = [a for a in range(3) if a != axis]
            # corners of the face in (u,v) order (0,0) (1,0) (1,1) (0,1)
            ring = []
            for (du, dv) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                off = [0, 0, 0]
                off[axis] = side
                off[u] = du
                off[v] = dv
                ring.append(off[0] | (off[1] << 1) | (off[2] << 2))
            # orientation: (e_u x e_v) . outward_normal > 0  <=> ccw seen from outside
            eu = np.zeros(3); eu[u] = 1
            ev = np.zeros(3); ev[v] = 1
            n = np.zeros(3); n[axis] = 1 if side else -1
            if np.dot(np.cross(eu, ev), n) < 0:
                ring = ring[::-1]
            faces.append(ring)
    return faces


FACES = _faces()


def _case_segments(case):
    """Directed segments (edge_from -> edge_to) on the cube surface for one case."""
    inside = [(case >> i) & 1 for i in range(8)]
    segs = []
    for ring in FACES:
        b = [inside[c] for c in ring]
        # crossings between ring[k] and ring[k+1]
        cross = [k for k in range(4) if b[k] != b[(k + 1) % 4]]
        if not cross:
            continue
        # Walk ccw (seen from outside).  A segment is emitted for every maximal run of inside
        # corners: it enters the run at crossing k_in (outside->inside) and leaves at k_out
        # (inside->outside).  With two separate runs (ambiguous face) each run gets its own
        # segment => inside corners are isolated.
        for k_in in cross:
            if b[k_in] == 0 and b[(k_in + 1) % 4] == 1:
                k = (k_in + 1) % 4
                while b[(k + 1) % 4] == 1:
                    k = (k + 1) % 4
                k_out = k
                e_in = _EDGE_OF[(ring[k_in], ring[(k_in + 1) % 4])]
                e_out = _EDGE_OF[(ring[k_out], ring[(k_out + 1) % 4])]
                # inside run is ccw-after e_in; direct the segment so that the inside region lies
                # to its RIGHT seen from outside => loop normals point inside->outside
                segs.append((e_in, e_out))
    return segs


def _case_triangles(case):
    segs = _case_segments(case)
    nxt = {}
    for a, b in segs:
        assert a not in nxt, "edge leaves twice"
        nxt[a] = b
    tris = []
    seen = set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        cur = nxt[start]
        while cur != start:
            loop.append(cur)
            seen.add(cur)
            cur = nxt[cur]
        assert len(loop) >= 3
        for k in range(1, len(loop) - 1):
            tris.append((loop[0], loop[k], loop[k + 1]))
    return tris


def build_table():
    """Returns (ntri[256] uint8, tri[256, MAXT*3] int8 padded with -1, edge_mask[256] uint16)."""
    all_tris = [_case_triangles(c) for c in range(256)]
    maxt = max(len(t) for t in all_tris)
    ntri = np.array([len(t) for t in all_tris], dtype=np.uint8)
    tri = -np.ones((256, maxt * 3), dtype=np.int8)
    emask = np.zeros(256, dtype=np.uint16)
    for c, ts in enumerate(all_tris):
        flat = [e for t in ts for e in t]
        tri[c, :len(flat)] = flat
        for e, (a, b) in enumerate(EDGE_CORNERS):
            if ((c >> a) & 1) != ((c >> b) & 1):
                emask[c] |= 1 << e
        assert set(flat) == {e for e in range(12) if emask[c] >> e & 1} or not flat
    return ntri, tri, emask


def _self_check():
    """Orientation check on the single-corner case: normal must point away from corner 0."""
    ntri, tri, _ = build_table()
    mid = np.array([(CORNER_OFF[a] + CORNER_OFF[b]) / 2.0 for a, b in EDGE_CORNERS])
    t = tri[1, :3]
    n = np.cross(mid[t[1]] - mid[t[0]], mid[t[2]] - mid[t[0]])
    assert np.dot(n, np.array([1.0, 1.0, 1.0])) > 0, "orientation must be inside->outside"
    assert ntri[0] == 0 and ntri[255] == 0
    return ntri, tri


def write_inc(path):
    ntri, tri, emask = build_table()
    _self_check()
    maxt = tri.shape[1] // 3
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_mc_table.py -- do not edit.  See that file for the derivation.\n")
        f.write("#define MC_MAX_TRI %d\n" % maxt)
        # one 16-byte row per case in global memory: bytes 0..14 the edge ids of up to five triangles (-1 padded), byte 15
        # the triangle count.  Per-lane lookups (one case per active cell) go through L1 as ONE 16-byte load instead of
        # serialising on the constant cache.
        assert tri.shape[1] == 15
        f.write("static __device__ const signed char g_mc_tri[256][16] = {\n")
        for c in range(256):
            f.write("  {%s},\n" % ",".join([str(int(v)) for v in tri[c]] + [str(int(ntri[c]))]))
        f.write("};\n")
    return maxt


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "..", "monoport_b200", "csrc", "mc_table.inc")
    m = write_inc(out)
    nt, _, _ = build_table()
    print("wrote", os.path.normpath(out), "max triangles/cell =", m, "total tris over cases =", int(nt.sum()))
