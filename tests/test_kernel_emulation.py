"""CPU-side checks of the integer / byte CUDA kernels: the kernel sources (monoport_b200/csrc/*_kernels.cuh, mp_scan.cuh)
are compiled UNMODIFIED against tests/emu/cuda_emu.h (one OS thread per CUDA thread, barriers for __syncthreads and
warp shuffles) and compared with the oracle.  The build container has no GPU; this catches indexing / scan / table
mistakes before a kernel is sent to a B200.  The `-m gpu` tests remain the parity tests of the real library."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import spec

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


def _build(tmp, name):
    exe = os.path.join(tmp, name)
    cmd = ["g++", "-std=c++17", "-O1", "-pthread", "-ffp-contract=off", "-Wno-unknown-pragmas",
           "-o", exe, os.path.join(EMU, name + ".cpp")]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


@pytest.fixture(scope="module")
def emu_mcubes(tmp_path_factory):
    return _build(str(tmp_path_factory.mktemp("emu")), "emu_mcubes")


def _run_mcubes(exe, vol, tmp_path, iso=0.5):
    D, H, W = vol.shape
    vin, vout, fout = (str(tmp_path / n) for n in ("vol.f32", "verts.f32", "faces.i32"))
    np.ascontiguousarray(vol, dtype=np.float32).tofile(vin)
    r = subprocess.run([exe, str(D), str(H), str(W), repr(float(iso)), vin, vout, fout], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    nv, nf = (int(x) for x in r.stdout.split())
    v = np.fromfile(vout, dtype=np.float32).reshape(-1, 3)
    f = np.fromfile(fout, dtype=np.int32).reshape(-1, 3)
    assert v.shape[0] == nv and f.shape[0] == nf
    return v, f


@pytest.mark.parametrize("kind,R", [("sphere", 33), ("ellipsoid", 40), ("two_blobs", 49)])
def test_mcubes_kernels_match_oracle_on_analytic_volumes(emu_mcubes, tmp_path, kind, R):
    vol = spec.analytic_volume(R, kind)
    V, F = spec.marching_cubes_ref(vol)
    v, f = _run_mcubes(emu_mcubes, vol, tmp_path)
    assert np.array_equal(f, F), "topology must be bit-exact"
    assert v.shape == V.shape and np.array_equal(v, V)


@pytest.mark.parametrize("shape", [(20, 17, 23), (9, 40, 35), (2, 2, 2), (3, 70, 2), (3, 5, 131), (2, 3, 257), (2, 9, 128),
                                   (2, 2, 33)])
def test_mcubes_kernels_match_oracle_on_noise(emu_mcubes, tmp_path, shape):
    """Dense noise: nearly every node is active (full shared-memory queues, every table case), odd and tiny shapes
    (rows shorter than a warp, rows of 2^k+1 nodes whose last chunk holds one node, rows longer than one chunk group,
    row tiles hanging over H, scan chunks hanging over n)."""
    vol = np.random.default_rng(sum(shape)).random(shape, dtype=np.float32)
    V, F = spec.marching_cubes_ref(vol)
    v, f = _run_mcubes(emu_mcubes, vol, tmp_path)
    assert np.array_equal(f, F.reshape(-1, 3))
    assert v.shape == V.shape and np.array_equal(v, V)


@pytest.mark.parametrize("fill", [0.0, 1.0])
def test_mcubes_kernels_empty_volume(emu_mcubes, tmp_path, fill):
    v, f = _run_mcubes(emu_mcubes, np.full((9, 9, 9), fill, dtype=np.float32), tmp_path)
    assert v.shape[0] == 0 and f.shape[0] == 0


# ---- the ordered scan itself (mp_scan.cuh): exclusive prefixes of two packed counters -----------------------------------
@pytest.fixture(scope="module")
def emu_scan(tmp_path_factory):
    return _build(str(tmp_path_factory.mktemp("emu_scan")), "emu_scan")


@pytest.mark.parametrize("vec", [0, 1])
@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 2047, 2048, 2049, 70001])
def test_ordered_scan_matches_cumsum(emu_scan, tmp_path, n, vec):
    _check_scan(emu_scan, tmp_path, n, vec)


def test_ordered_scan_many_ctas(emu_scan, tmp_path):
    """More CTA totals than threads in the last CTA: every thread of it owns a run of several totals."""
    _check_scan(emu_scan, tmp_path, 2048 * 256 + 4097, 1)


def _check_scan(exe, tmp_path, n, vec):
    rng = np.random.default_rng(n + vec)
    b = (rng.integers(0, 64, size=n, dtype=np.uint8) * (rng.random(n) < 0.3)).astype(np.uint8)
    fin, fout = str(tmp_path / "in.u8"), str(tmp_path / "out.u64")
    b.tofile(fin)
    r = subprocess.run([exe, str(n), str(vec), fin, fout], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lo, hi = (b & 7).astype(np.uint64), (b >> 3).astype(np.uint64)
    assert [int(x) for x in r.stdout.split()] == [int(lo.sum()), int(hi.sum())]
    if n:
        want = (np.cumsum(lo) - lo) + ((np.cumsum(hi) - hi) << np.uint64(32))
        assert np.array_equal(np.fromfile(fout, dtype=np.uint64), want)
