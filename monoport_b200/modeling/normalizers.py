"""Depth feature (monoport/lib/modeling/normalizers/DepthNormalizer.py).  Only the scale matters to the fused
kernel (`z * scale`, :32); the soft one-hot branch (:17-30) is dead for every shipped config."""
import torch.nn as nn

from ..config import CN


class DepthNormalizer(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        if getattr(opt, "soft_onehot", False):
            raise NotImplementedError("soft_onehot depth features are not supported by the fused kernel")

    @property
    def scale(self):
        return float(self.opt.scale)

    def forward(self, z, calibs=None, index_feat=None):
        return z * self.opt.scale


def PIFuNomalizer(*args, **kwargs):
    opt = CN()
    opt.soft_onehot = False
    opt.scale = 512 // 2 / 200.0
    return DepthNormalizer(opt)
