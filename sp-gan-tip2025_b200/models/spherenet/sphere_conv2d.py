"""Mirror of models/spherenet/sphere_conv2d.py."""
import math

import torch
from torch import nn

from ... import functional as SF
from ...functional import ConvGeom
from ... import grids
from ...grids import GRID_CACHE
from ..ops import leaky_relu
from .grid_generator import GridSampler, GridSamplerNewTextureNoGrad

_SPHERE_GEOM = ConvGeom(3, 3, stride=3, pad=0)


class SphereConvBatchDiffFixBorderGNoGrad(nn.Conv2d):
    """sphere_conv2d.py:124-205: gather the RGB skip at the spherical taps, 3x3 stride-3 conv with W/sqrt(Cin*9) and
    bias, LeakyReLU(0.01).  Weight initialises to the centre-delta kernel (:137-145)."""

    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1, padding=0, dilation=1, groups=1,
                 bias=True, padding_mode='zeros'):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        self.grid_shape = None
        self.grid = None
        self.scale = 1 / math.sqrt(in_channels * self.kernel_size[0] ** 2)
        self.sampler = GridSamplerNewTextureNoGrad()
        self.activation = nn.LeakyReLU()
        delta = torch.tensor([[0., 0., 0.], [0., 1., 0.], [0., 0., 0.]])
        self.weight = nn.Parameter(delta.repeat(out_channels, in_channels, 1, 1))

    def genSamplingPattern(self, h, w, stride, coords_partial):
        grid = GRID_CACHE.get(h, w, coords_partial, self.weight.device)
        if coords_partial.get("test_flag", False):
            self.grid = grid
            return None
        return grid

    def forward(self, x, coords_partial):
        B, C, H, W = x.shape
        if tuple(self.kernel_size) != (3, 3) or tuple(self.padding) != (0, 0):
            raise NotImplementedError("only the 3x3 / padding-0 configuration of spgan.yaml is implemented")
        self.grid_shape = (H, W)
        grid = GRID_CACHE.batch(H, W, coords_partial, B, x.device)
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            g = self.sampler(x, grid)
            y = SF.conv2d(g, self.weight, _SPHERE_GEOM, out_scale=self.scale if self.scale else 1.0)
            if self.bias is not None:
                y = y + self.bias.view(1, -1, 1, 1)
            return leaky_relu(y, 0.01)
        g = SF.sphere_gather_raw(x, grid)
        return SF.conv_apply(g, self.weight, _SPHERE_GEOM, out_scale=self.scale if self.scale else 1.0, bias=self.bias,
                             act=(0.01, 1.0), precision=0)


class SphereConv2d(nn.Conv2d):
    """sphere_conv2d.py:16-67: SphereNet's full-sphere conv — sample the whole equirectangular feature at the tangent-plane
    taps of every pixel ('nearest' sampler), then a Kh x Kw conv with stride = kernel size."""

    _pattern = staticmethod(grids.full_sphere_pattern)

    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1, padding=0, dilation=1, scale=None,
                 groups=1, bias=True, padding_mode='zeros'):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        self.grid_shape = None
        self.grid = None
        self.scale = scale
        self.sampler = GridSampler()

    def genSamplingPattern(self, h, w):
        pattern = self._pattern(h, w, tuple(self.kernel_size), tuple(self.stride))
        self.grid = torch.from_numpy(grids.full_sphere_grid(pattern, h, w))

    def forward(self, x):
        B, C, H, W = x.shape
        if self.grid_shape is None or self.grid_shape != (H, W):
            self.grid_shape = (H, W)
            self.genSamplingPattern(H, W)
        if self.grid.device != x.device:
            self.grid = self.grid.to(x.device)
        if self.groups != 1 or tuple(self.dilation) != (1, 1) or self.kernel_size[0] != self.kernel_size[1] or \
                self.padding[0] != self.padding[1]:
            raise NotImplementedError("SphereConv2d: groups / dilation / rectangular kernels are not used by the reference")
        g = self.sampler(x, self.grid)  # one grid shared by the batch (the reference repeats it B times)
        k = self.kernel_size[0]
        geom = ConvGeom(k, k, stride=k, pad=self.padding[0])
        y = SF.conv2d(g, self.weight, geom, out_scale=self.scale if self.scale else 1.0)
        if self.bias is not None:
            y = y + self.bias.view(1, -1, 1, 1)
        return y


class IncreIntervalSphereConv2d(SphereConv2d):
    """sphere_conv2d.py:70-121: same op on IncreIntervalGridGenerator's pattern."""

    _pattern = staticmethod(grids.incre_interval_pattern)
