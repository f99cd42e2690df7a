"""Stand-alone projection / sampling helpers with the reference's signatures (monoport/lib/modeling/geometry.py).
The fused kernel does all of this internally; these exist because callers import them directly
(RTL/main.py:22,237 uses `orthogonal` to map voxel indices to world space)."""
import torch
import torch.nn.functional as F


def index(feat, uv):
    """[B,C,H,W], [B,2,N] in [-1,1] -> [B,C,N]; bilinear, zeros padding, align_corners=True (geometry.py:4-16)."""
    grid = uv.transpose(1, 2).unsqueeze(2)
    return F.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[..., 0]


def _affine(points, calibrations):
    return torch.baddbmm(calibrations[:, :3, 3:4], calibrations[:, :3, :3], points)


def _reject_transforms(transforms):
    if transforms is not None:
        # the reference slices `transforms[:2, :2]` on the BATCH axis (geometry.py:31-32), which only type-checks
        # for B>=2 square cases and is never exercised on the recon path (RTL/main.py:179-182)
        raise NotImplementedError("image-space `transforms` are not supported (unused on the reconstruction path)")


def orthogonal(points, calibrations, transforms=None):
    """[B,3,N] world -> [B,3,N] image space:  R p + t  (geometry.py:19-34)."""
    _reject_transforms(transforms)
    return _affine(points, calibrations)


def perspective(points, calibrations, transforms=None):
    """geometry.py:37-55:  (x/z, y/z, z) of R p + t."""
    _reject_transforms(transforms)
    h = _affine(points, calibrations)
    return torch.cat([h[:, :2] / h[:, 2:3], h[:, 2:3]], 1)
