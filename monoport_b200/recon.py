"""Surface extraction + mesh reconstruction on the device.

  pifu_calib, forward_vertices : same signatures/returns as RTL/recon.py:4-25 / :27-89 (the reference file itself can
                                 also be used unchanged on the volumes the engines return);
  marching_cubes, reconstruction: additions asked for by the north star (the reference has no marching cubes;
                                 `reconstruction` follows upstream PIFu's signature).
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib

_DIRS = {"front": 0, "back": 1, "left": 2, "right": 3}


@torch.no_grad()
def pifu_calib(extrinsic, intrinsic, device="cuda:0"):
    """inv(K' E' diag(1,-1,1,1)) in float64 -> float32 [1,4,4]  (RTL/recon.py:4-25; host-side, negligible)."""
    K = np.array(intrinsic, dtype=np.float64, copy=True)
    E = np.array(extrinsic, dtype=np.float64, copy=True)
    K[2, 2] = K[0, 0]
    K[2, 3] = 0
    E[2, 3] = 0
    m = np.linalg.inv(K @ E @ np.diag([1.0, -1.0, 1.0, 1.0]))
    return torch.from_numpy(m).unsqueeze(0).float().to(device)


@torch.no_grad()
def forward_vertices(sdf, direction="front"):
    """Visible-surface vertices + normals of a [1,1,R,R,R] occupancy volume (RTL/recon.py:27-89), one kernel pass."""
    if sdf is None:
        return None, None, None, None
    if sdf.device.type != "cuda":
        raise RuntimeError("monoport_b200.forward_vertices: CUDA tensors only (no CPU path)")
    vol = sdf[0, 0]
    if vol.dtype != torch.float32 or not vol.is_contiguous():
        vol = vol.float().contiguous()
    R = vol.shape[2]
    assert vol.shape[0] == R and vol.shape[1] == R, "cubic volumes only"
    dev = vol.device
    cap = R * R
    X = torch.empty(cap, dtype=torch.int64, device=dev)
    Y = torch.empty(cap, dtype=torch.int64, device=dev)
    Z = torch.empty(cap, dtype=torch.float32, device=dev)
    N = torch.empty((cap, 3), dtype=torch.float32, device=dev)
    n = ctypes.c_int64(0)
    with _lib.device_guard(dev):
        _lib.check(_lib.load().mp_forward_vertices(
            ctypes.c_void_p(vol.data_ptr()), R, _DIRS[direction], ctypes.c_void_p(X.data_ptr()),
            ctypes.c_void_p(Y.data_ptr()), ctypes.c_void_p(Z.data_ptr()), ctypes.c_void_p(N.data_ptr()),
            ctypes.byref(n), _lib.stream_ptr(dev)), "mp_forward_vertices")
    k = n.value
    return X[:k], Y[:k], Z[:k], N[:k]


class _McubesWorkspace:
    _cache = {}

    @classmethod
    def get(cls, shape, device):
        key = (tuple(shape), device.index, threading.get_ident())
        h = cls._cache.get(key)
        if h is None:
            h = ctypes.c_void_p()
            with _lib.device_guard(device):
                _lib.check(_lib.load().mp_mcubes_create(shape[0], shape[1], shape[2], ctypes.byref(h)), "mp_mcubes_create")
            cls._cache[key] = h
        return h


@torch.no_grad()
def marching_cubes(vol, iso=0.5):
    """[D,H,W] (z,y,x) float32 CUDA volume -> (verts [V,3] float32 in index space (x,y,z), faces [F,3] int32).
    Indexed, watertight, deterministic ordering (see csrc/mcubes.cu)."""
    if vol.device.type != "cuda":
        raise RuntimeError("monoport_b200.marching_cubes: CUDA tensors only (no CPU path)")
    if vol.dim() == 5:
        vol = vol[0, 0]
    if vol.dtype != torch.float32 or not vol.is_contiguous():
        vol = vol.float().contiguous()
    dev = vol.device
    lib = _lib.load()
    h = _McubesWorkspace.get(vol.shape, dev)
    nv, nf = ctypes.c_int64(0), ctypes.c_int64(0)
    with _lib.device_guard(dev):
        st = _lib.stream_ptr(dev)
        _lib.check(lib.mp_mcubes_count(h, ctypes.c_void_p(vol.data_ptr()), ctypes.c_float(iso), ctypes.byref(nv),
                                       ctypes.byref(nf), st), "mp_mcubes_count")
        verts = torch.empty((nv.value, 3), dtype=torch.float32, device=dev)
        faces = torch.empty((nf.value, 3), dtype=torch.int32, device=dev)
        _lib.check(lib.mp_mcubes_emit(h, ctypes.c_void_p(vol.data_ptr()), ctypes.c_float(iso),
                                      ctypes.c_void_p(verts.data_ptr()), ctypes.c_void_p(faces.data_ptr()), st),
                   "mp_mcubes_emit")
    return verts, faces


@torch.no_grad()
def reconstruction(net, cuda, calib_tensor, resolution, b_min, b_max, use_octree=False, num_samples=10000,
                   transform=None, feats=None, engine=None):
    """PIFu-shaped mesh reconstruction (upstream signature): evaluate the occupancy field on a
    `resolution`^3 grid over [b_min,b_max] (dense through the fused kernel, or coarse-to-fine when `use_octree`),
    run marching cubes at 0.5, map vertices to world space.  `feats` = net.filter(image) output.
    Returns (verts [V,3] world, faces [F,3], normals None, values None) or -1 when the volume is empty
    (upstream returns -1 on marching-cubes failure).  `num_samples` is accepted for signature compatibility
    (the fused kernel needs no batching)."""
    if feats is None:
        raise ValueError("pass feats=net.filter(image)")
    R = int(resolution)
    device = torch.device(cuda)
    b_min_t = torch.as_tensor(np.asarray(b_min, dtype=np.float32)).view(3)
    b_max_t = torch.as_tensor(np.asarray(b_max, dtype=np.float32)).view(3)
    if use_octree:
        if engine is None and (R < 3 or (R - 1) & (R - 2)):
            raise ValueError("use_octree=True needs a resolution of 2^k+1 nodes per axis (17, 33, ..., 257, 513: RTL/main.py:187); "
                             "got %d -- use %d, or the dense path (use_octree=False), which takes any resolution"
                             % (R, (1 << max(1, (R - 1).bit_length())) + 1))
        if engine is None:
            from .engine import Seg3dLossless, make_query_func
            res = [R]
            while res[0] > 17 and (res[0] - 1) % 2 == 0:
                res.insert(0, (res[0] - 1) // 2 + 1)
            engine = Seg3dLossless(make_query_func(net), b_min_t.numpy()[None], b_max_t.numpy()[None], res,
                                   balance_value=0.5, faster=False).to(device)
        sdf = engine(im_feat_list=feats, calib_tensor=calib_tensor)
        if sdf is None:
            return -1
        vol = sdf[0, 0]
    else:
        vol = net.query_grid(feats[-1][0], calib_tensor, R, b_min_t, b_max_t)
    verts, faces = marching_cubes(vol, 0.5)
    if verts.shape[0] == 0:
        return -1
    # index space -> world:  (v + 0.5)/R * (b_max-b_min) + b_min  (node-centre convention of the engines)
    scale = ((b_max_t - b_min_t) / R).to(verts.device)
    verts = (verts + 0.5) * scale + b_min_t.to(verts.device)
    if transform is not None:
        T = torch.as_tensor(transform, dtype=torch.float32, device=verts.device)
        verts = verts @ T[:3, :3].T + T[:3, 3]
    return verts, faces, None, None


@torch.no_grad()
def make_mat_color(resolution, b_min, b_max, device):
    """Voxel index -> world matrix of the demo (RTL/main.py:201-210): diag((b_max-b_min)/R) with translation b_min."""
    b_min = torch.as_tensor(np.asarray(b_min, dtype=np.float32)).view(3)
    b_max = torch.as_tensor(np.asarray(b_max, dtype=np.float32)).view(3)
    mat = torch.eye(4, dtype=torch.float32)
    length = b_max - b_min
    for a in range(3):
        mat[a, a] = length[a] / resolution
    mat[0:3, 3] = b_min
    return mat.to(device)


@torch.no_grad()
def colorization(netC, feat_tensor_C, X, Y, Z, calib_tensor, norm=None, resolution=257, b_min=(-1, -1, -1),
                 b_max=(1, 1, 1), canvas=None):
    """Direct rendering of the visible surface (RTL/main.py:212-249): either normals as colours, or netC queried at the
    visible vertices (X, Y, R - Z) mapped to world space by mat_color, `pred * 0.5 + 0.5`, scattered into an [R,R,3]
    canvas.  Returns None when there is no surface (X is None), like the reference."""
    if X is None:
        return None
    device = calib_tensor.device
    if canvas is None:
        canvas = torch.ones((resolution, resolution, 3), dtype=torch.float32, device=device)
    image = canvas.clone()
    if norm is not None:
        image[X, Y, :] = ((norm + 1) / 2).clamp(0, 1)
        return image
    from .modeling.geometry import orthogonal, perspective
    head = netC.surface_classifier
    feat = feat_tensor_C[-1][0]
    if (netC.precision in ("auto", "tc") and not netC.training and X.is_cuda and feat.is_cuda and feat.device == device
            and head.filter_channels[-1] == 3 and head.tc_supported()):
        # one launch: vertices -> world (mat_color) -> netC -> pred*0.5+0.5 -> canvas  (tensor-core program of the colour head)
        n = int(X.numel())
        rc = _lib.MP_OK
        if n:
            Xc, Yc, Zc = X.contiguous(), Y.contiguous(), Z.float().contiguous()
            with _lib.device_guard(device):
                fh = netC.feature_handle(feat)
                proj = _lib.PROJ_PERSPECTIVE if netC.projection is perspective else _lib.PROJ_ORTHOGONAL
                rc = _lib.load().mp_colorize_surface(
                    head.handle(), fh.ptr, ctypes.c_void_p(Xc.data_ptr()), ctypes.c_void_p(Yc.data_ptr()),
                    ctypes.c_void_p(Zc.data_ptr()), n, int(resolution), _lib.f3(b_min), _lib.f3(b_max), _lib.calib12(calib_tensor),
                    proj, ctypes.c_float(netC.normalizer.scale), ctypes.c_void_p(image.data_ptr()), _lib.stream_ptr(device))
                if rc != _lib.MP_E_RANGE:
                    _lib.check(rc, "mp_colorize_surface")
        if rc != _lib.MP_E_RANGE:
            return image
        # features outside the validated range of the tensor-core colour program: the generic route below (its query()
        # takes the exact fp32 kernel for such a frame)
    verts = torch.stack([X.float(), Y.float(), resolution - Z.float()], dim=1)          # RTL/main.py:231-233
    samples = verts.unsqueeze(0).permute(0, 2, 1).contiguous()                          # [1,3,N]
    samples = orthogonal(samples, make_mat_color(resolution, b_min, b_max, device).unsqueeze(0))
    feats = [[f.to(device) for f in fs] for fs in feat_tensor_C]
    preds = netC.query(feats, points=samples, calibs=calib_tensor)[0]
    image[X, Y, :] = (preds[0] * 0.5 + 0.5).t()
    return image
