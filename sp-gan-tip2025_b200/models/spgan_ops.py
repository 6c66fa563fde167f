"""Mirror of the reference's models/spgan_ops.py.  The shipped generator instantiates only `ToRGB` (and through it
`ModulatedConv2d` 1x1 and `Upsample`) from this module (models/spgan/spgan.py:13-15, 724); `SphereModulatedConv2d` /
`StyledConv` are the signature set north_star pins: same constructor arguments, parameter names and arithmetic as the
reference classes, including their two differences from the live models/spgan_ops_gs.py variants."""
import torch

from .. import functional as SF
from ..grids import GRID_CACHE
from .custom_ops import FusedLeakyReLU, fused_leaky_relu, upfirdn2d  # noqa: F401
from .ops import (Blur, ConstantInput, Downsample, EqualConv2d, EqualLinear, ModulatedConv2d, NoiseInjection,  # noqa: F401
                  PixelNorm, ScaledLeakyReLU, ToRGB, Upsample, create_gaussian_kernel, make_kernel)
from . import spgan_ops_gs as _gs
from .spherenet import GridSamplerNewTexture


class SphereModulatedConv2d(_gs.ModulatedConv2d):
    """models/spgan_ops.py:736-1379.  Differs from the live spgan_ops_gs.ModulatedConv2d in two ways, both reproduced:
    the sampler is `GridSamplerNewTexture` (grid_sample_github: bilinear weights from the UNCLIPPED coordinate, clamped
    indices, true gradient — models/spherenet/grid_sample_ops.py:5-55) instead of the border-clipped gather with the surrogate
    gradient, and with `deal_coords` the gathered features are viewed as (1, batch * 256, ...) (:1202), so any feature width
    other than 256 is an error.  The constructor keyword is `size_cut` (:751), not `cut_size`."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False, downsample=False,
                 blur_kernel=[1, 2, 1], no_zero_pad=False, config=None, side=None, deal_coords=False, size_cut=False):
        super().__init__(in_channel, out_channel, kernel_size, style_dim, demodulate=demodulate, upsample=upsample,
                         downsample=downsample, blur_kernel=blur_kernel, no_zero_pad=no_zero_pad, config=config, side=side,
                         deal_coords=deal_coords, cut_size=size_cut)
        self.sampler = GridSamplerNewTexture()

    def forward(self, input, style, coords=None, coords_partial=None, calc_flops=False):
        batch, C, H, W = input.shape
        flops = self.get_flops(input, style) if calc_flops else 0
        if style is not None and style.ndim == 4:
            mean_style = style.mean([2, 3], keepdim=True)
            if ((style - mean_style) < 1e-8).all():
                style = mean_style.squeeze()
        if style.ndim != 2 or self.upsample:
            out, _ = ModulatedConv2d.forward(self, input, style)  # spatial styles / upsampling: the plain op (:1263-1379, :1143-1155)
            return out, flops
        if self.deal_coords and C != 256:
            raise RuntimeError("shape '[1, %d, %d, %d]' is invalid for input of size %d (models/spgan_ops.py:1202 views the "
                               "gathered features as batch * 256 channels)" % (batch * 256, 3 * H, 3 * W, batch * C * 9 * H * W))
        s, w, d = self._mod_demod(style, batch)
        self.grid_shape = (H, W)
        grid = GRID_CACHE.batch(H, W, coords_partial, batch, input.device)
        out = SF.sphere_modconv(input, coords if self.deal_coords else None, grid, w, s, d, self.scale, flat_concat=True,
                                sampler="texture")
        if self.cut_size:  # (:1216-1217, :1259-1260): both branches crop when size_cut is set
            out = out[:, :, 1:-1, 1:-1]
        return out, flops

    def forward_fused(self, *a, **k):
        raise NotImplementedError("SphereModulatedConv2d has no fused inference path (it is not instantiated by spgan.yaml)")


class StyledConv(_gs.StyledConv):
    """models/spgan_ops.py:1448-1520: StyledConv over `SphereModulatedConv2d`."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 2, 1],
                 demodulate=True, no_zero_pad=False, disable_noise=False, activation="LeakyReLU", config=None, side=None,
                 deal_coords=False):
        super().__init__(in_channel, out_channel, kernel_size, style_dim, upsample=upsample, blur_kernel=blur_kernel,
                         demodulate=demodulate, no_zero_pad=no_zero_pad, disable_noise=disable_noise, activation=activation,
                         config=config, side=side, deal_coords=deal_coords)
        old = self.conv
        self.conv = SphereModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                          blur_kernel=blur_kernel, demodulate=demodulate, no_zero_pad=no_zero_pad, config=config,
                                          side=side, deal_coords=deal_coords)
        del old

    def forward(self, input, style, noise=None, coords=None, coords_partial=None, test_ids=None, calc_flops=False):
        flops = 0
        out, cur = self.conv(input, style, coords=coords, coords_partial=coords_partial, calc_flops=calc_flops)
        flops += cur
        if self.noise is not None:
            out, cur = self.noise(out, noise=noise, test_ids=test_ids, calc_flops=calc_flops)
            flops += cur
        out = self._activate(out)
        return out, flops
