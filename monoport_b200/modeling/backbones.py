"""Image encoders that FEED the hot path -- they stay plain PyTorch (north star).
Written against the reference's state-dict layout so its checkpoints load unchanged:
  HGFilter      monoport/lib/modeling/backbones/HGFilters.py:117-204   (4-stack hourglass, GroupNorm, avg-pool down)
  ResnetFilter  monoport/lib/modeling/backbones/ResBlkFilters.py:87-139 (reflect-pad ResNet, GroupNorm)
Only the configurations the shipped factories select (PIFuHGFilters / PIFuResBlkFilters) are implemented.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..config import CN


def _gn(c):
    return nn.GroupNorm(32, c)


class ConvBlock(nn.Module):
    """Pre-activation block: three 3x3 convs whose outputs (C/2, C/4, C/4) are concatenated, plus a residual
    (1x1-projected when the width changes).  HGFilters.py:12-62."""

    def __init__(self, cin, cout, norm="group"):
        super().__init__()
        if norm != "group":
            raise NotImplementedError("only GroupNorm encoders are shipped (HGFilters.py:210)")
        h, q = cout // 2, cout // 4
        self.conv1 = nn.Conv2d(cin, h, 3, 1, 1, bias=False)
        self.conv2 = nn.Conv2d(h, q, 3, 1, 1, bias=False)
        self.conv3 = nn.Conv2d(q, q, 3, 1, 1, bias=False)
        self.bn1, self.bn2, self.bn3, self.bn4 = _gn(cin), _gn(h), _gn(q), _gn(cin)
        # `downsample.0` aliases bn4 in the reference's state dict (HGFilters.py:30-35)
        self.downsample = nn.Sequential(self.bn4, nn.ReLU(True), nn.Conv2d(cin, cout, 1, bias=False)) if cin != cout else None

    def forward(self, x):
        a = self.conv1(F.relu(self.bn1(x)))
        b = self.conv2(F.relu(self.bn2(a)))
        c = self.conv3(F.relu(self.bn3(b)))
        res = x if self.downsample is None else self.downsample(x)
        return torch.cat((a, b, c), 1) + res


class HourGlass(nn.Module):
    """Recursive hourglass (HGFilters.py:65-114): avg-pool down, bicubic(align_corners=True) up (:108)."""

    def __init__(self, num_modules, depth, num_features, norm="group"):
        super().__init__()
        self.depth = depth
        for level in range(depth, 0, -1):
            self.add_module("b1_%d" % level, ConvBlock(num_features, num_features, norm))
            self.add_module("b2_%d" % level, ConvBlock(num_features, num_features, norm))
            if level == 1:
                self.add_module("b2_plus_1", ConvBlock(num_features, num_features, norm))
        for level in range(1, depth + 1):
            self.add_module("b3_%d" % level, ConvBlock(num_features, num_features, norm))

    def _run(self, level, x):
        up = self._modules["b1_%d" % level](x)
        low = self._modules["b2_%d" % level](F.avg_pool2d(x, 2, stride=2))
        low = self._run(level - 1, low) if level > 1 else self._modules["b2_plus_1"](low)
        low = self._modules["b3_%d" % level](low)
        return up + F.interpolate(low, scale_factor=2, mode="bicubic", align_corners=True)

    def forward(self, x):
        return self._run(self.depth, x)


class HGFilter(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.num_modules = opt.num_stack
        if opt.hg_down != "ave_pool" or opt.norm != "group":
            raise NotImplementedError("only the shipped HG configuration (group norm, ave_pool) is implemented")
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3)
        self.bn1 = _gn(64)
        self.conv2 = ConvBlock(64, 128)
        self.conv3 = ConvBlock(128, 128)
        self.conv4 = ConvBlock(128, 256)
        for i in range(self.num_modules):
            self.add_module("m%d" % i, HourGlass(1, opt.num_hourglass, 256))
            self.add_module("top_m_%d" % i, ConvBlock(256, 256))
            self.add_module("conv_last%d" % i, nn.Conv2d(256, 256, 1))
            self.add_module("bn_end%d" % i, _gn(256))
            self.add_module("l%d" % i, nn.Conv2d(256, opt.hourglass_dim, 1))
            if i < self.num_modules - 1:
                self.add_module("bl%d" % i, nn.Conv2d(256, 256, 1))
                self.add_module("al%d" % i, nn.Conv2d(opt.hourglass_dim, 256, 1))

    def forward(self, x):
        m = self._modules
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.avg_pool2d(self.conv2(x), 2, stride=2)
        prev = self.conv4(self.conv3(x))
        outputs = []
        for i in range(self.num_modules):
            ll = m["top_m_%d" % i](m["m%d" % i](prev))
            ll = F.relu(m["bn_end%d" % i](m["conv_last%d" % i](ll)))
            out = m["l%d" % i](ll)
            outputs.append((out,))
            if i < self.num_modules - 1:
                prev = prev + m["bl%d" % i](ll) + m["al%d" % i](out)
        return outputs


def PIFuHGFilters(*args, **kwargs):
    opt = CN()
    opt.norm = "group"
    opt.num_stack = 4
    opt.num_hourglass = 2
    opt.skip_hourglass = False
    opt.hg_down = "ave_pool"
    opt.hourglass_dim = 256
    return HGFilter(opt)


class ResnetBlock(nn.Module):
    def __init__(self, dim, last=False):
        super().__init__()
        layers = [nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3, bias=False), _gn(dim), nn.ReLU(True),
                  nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3, bias=False)]
        if not last:
            layers.append(_gn(dim))
        self.conv_block = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.conv_block(x)


class ResnetFilter(nn.Module):
    """ResBlkFilters.py:87-139 with norm='group', 6 blocks, reflect padding, no tanh."""

    def __init__(self, opt, input_nc=3, ngf=64, n_blocks=6):
        super().__init__()
        layers = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, 7, bias=False), _gn(ngf), nn.ReLU(True)]
        c = ngf
        for _ in range(2):
            layers += [nn.Conv2d(c, 2 * c, 3, 2, 1, bias=False), _gn(2 * c), nn.ReLU(True)]
            c *= 2
        layers += [ResnetBlock(c, last=(i == n_blocks - 1)) for i in range(n_blocks)]
        if getattr(opt, "use_tanh", False):
            layers.append(nn.Tanh())
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return [(self.model(x),)]


def PIFuResBlkFilters(*args, **kwargs):
    opt = CN()
    opt.use_tanh = False
    return ResnetFilter(opt)
