"""Mirror of the reference's models/spgan_ops_gs.py: the module set of models/ops.py plus the LIVE spherical modulated
conv (`ModulatedConv2d` with `deal_coords`, used through `StyledConv` by SphereConditionalBlock,
models/spgan/spgan.py:143).  Everything that is identical to models/ops.py is re-exported from there."""
import math

import numpy as np
import torch
from torch import nn

from .. import functional as SF
from ..grids import GRID_CACHE
from . import ops as norm_ops
from .custom_ops import FusedLeakyReLU, fused_leaky_relu, upfirdn2d  # noqa: F401
from .ops import (ConstantInput, Downsample, EqualLinear, NoiseInjection, PixelNorm, ScaledLeakyReLU, ToRGB,  # noqa: F401
                  Upsample, _grad_needed, create_gaussian_kernel, leaky_relu, make_kernel)
from .spherenet import GridSamplerNewTextureNoGrad


class Blur(nn.Module):
    """models/spgan_ops_gs.py:89-193: the spherical Blur variant is dead code in the reference (its forward always
    raises, :133-134, 161); the name is kept for import compatibility."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, *args, **kwargs):
        raise NotImplementedError("spgan_ops_gs.Blur is unreachable in the reference as well (its forward raises)")


class EqualConv2d(norm_ops.EqualConv2d):
    """models/spgan_ops_gs.py:196-263 (spherical variant never instantiated by spgan.yaml): plain behaviour kept."""


class ModulatedConv2d(norm_ops.ModulatedConv2d):
    """models/spgan_ops_gs.py:311-972: style-modulated, demodulated SPHERICAL 3x3 conv.

    forward(input, style, coords=None, coords_partial=None, calc_flops=False) -> (out, flops).  With `deal_coords`
    the raw coordinate planes are gathered at the same taps, encoded (tanh / cos pi / sin pi) and concatenated; the
    reference's flat (1, B*C) ++ (1, B*3) concatenation under groups=B (:792-814) is reproduced exactly."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 2, 1], no_zero_pad=False, config=None, side=None, deal_coords=False,
                 cut_size=False, _init_zero=True):
        super().__init__(in_channel, out_channel, kernel_size, style_dim, demodulate=demodulate, upsample=upsample,
                         downsample=downsample, blur_kernel=blur_kernel, no_zero_pad=no_zero_pad, config=config,
                         side=side)
        assert kernel_size == 3, f"{kernel_size} != 3 is not supported"
        self.deal_coords = deal_coords
        self.cut_size = cut_size
        delta = torch.tensor([[0., 0., 0.], [0., 1., 0.], [0., 0., 0.]]).repeat(1, out_channel, in_channel, 1, 1)
        if not _init_zero:
            delta = delta + torch.randn(1, out_channel, in_channel, 3, 3) * 1e-4
        self.weight = nn.Parameter(delta)  # centre-delta init (:374-395), not randn
        self.grid_shape = None
        self.grid = None
        self.sampler = GridSamplerNewTextureNoGrad()
        self.sp_k_size = (kernel_size, kernel_size)

    def genSamplingPattern(self, h, w, stride, coords_partial):
        """:410-428 — returns the (1, 3h, 3w, 2) grid in training mode, stores it in `self.grid` in test mode."""
        grid = GRID_CACHE.get(h, w, coords_partial, self.weight.device)
        if coords_partial.get("test_flag", False):
            self.grid = grid
            return None
        return grid

    def _coord_dim(self):
        n = self.config.train_params.coord_num_dir
        assert isinstance(n, int)
        return n

    def forward(self, input, style, coords=None, coords_partial=None, calc_flops=False):
        batch, C, H, W = input.shape
        flops = self.get_flops(input, style) if calc_flops else 0
        if style is not None and style.ndim == 4:
            mean_style = style.mean([2, 3], keepdim=True)
            if ((style - mean_style) < 1e-8).all():
                style = mean_style.squeeze()
        if style.ndim != 2:
            raise NotImplementedError("spatially-shaped styles are outside the B200 hot path")
        if self.upsample:
            out, _ = norm_ops.ModulatedConv2d.forward(self, input, style)
            return out, flops
        self._check_config()
        s, w, d = self._mod_demod(style, batch)  # in_channel already counts the coordinate planes
        self.grid_shape = (H, W)
        grid = GRID_CACHE.batch(H, W, coords_partial, batch, input.device)
        out = SF.sphere_modconv(input, coords if self.deal_coords else None, grid, w, s, d, self.scale, flat_concat=True)
        if (not self.deal_coords) and self.cut_size:
            out = out[:, :, 1:-1, 1:-1]
        return out, flops

    def _check_config(self):
        """The configurations the spherical conv supports; shared by the autograd path and the fused no_grad path so that eval
        mode cannot silently differ from train mode (models/spgan_ops_gs.py:700-760)."""
        if self.padding != 0:
            raise NotImplementedError("the spherical conv is only used with no_zero_pad=True (padding 0) in spgan.yaml")
        if self.deal_coords:
            nc = self._coord_dim()
            if nc != 3:
                raise NotImplementedError("coord_num_dir == %d: only the 3-channel encoding of spgan.yaml is implemented" % nc)
            tp = self.config.train_params
            if (hasattr(tp, "no_coord_encode_all") and tp.no_coord_encode) or (hasattr(tp, "no_coord_encode") and tp.no_coord_encode):
                raise NotImplementedError("no_coord_encode variants are not used by spgan.yaml")

    def forward_fused(self, input, style, coords, coords_partial, act, residual=None):
        """no_grad fast path: gather + encode + modulate + conv + LeakyReLU (+ residual) in two kernels.  Same guards and
        the same `cut_size` crop as forward(); a residual cannot be folded in when the output is cropped afterwards."""
        batch, C, H, W = input.shape
        self._check_config()
        crop = (not self.deal_coords) and self.cut_size
        s, w, d = self._mod_demod(style, batch)
        grid = GRID_CACHE.batch(H, W, coords_partial, batch, input.device)
        out = SF.sphere_modconv_fused(input, coords if self.deal_coords else None, grid, w, s, d, self.scale, act=act,
                                      flat_concat=True, residual=None if crop else residual)
        if crop:
            out = out[:, :, 1:-1, 1:-1]
            if residual is not None:
                out = out + residual
        return out


class StyledConv(nn.Module):
    """models/spgan_ops_gs.py:1041-1112."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 2, 1],
                 demodulate=True, no_zero_pad=False, disable_noise=False, activation="LeakyReLU", config=None,
                 side=None, deal_coords=False, _init_zero=True):
        super().__init__()
        self.no_zero_pad = no_zero_pad
        self.upsample = upsample
        self.deal_coords = deal_coords
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate, no_zero_pad=no_zero_pad,
                                    config=config, side=side, deal_coords=deal_coords, _init_zero=_init_zero)
        self.noise = None if disable_noise else NoiseInjection()
        if activation.lower() == "leakyrelu":
            self.activate = FusedLeakyReLU(out_channel)
        elif activation.lower() == "leakyrelu_n":
            self.activate = nn.LeakyReLU()
        else:
            raise NotImplementedError("Unknown activation {}".format(activation))

    def calc_in_spatial_size(self, out_spatial_size):
        return self.conv.calc_in_spatial_size(out_spatial_size)

    def calc_out_spatial_size(self, in_spatial_size):
        return self.conv.calc_out_spatial_size(in_spatial_size)

    def calibrate_spatial_shape(self, spatial_latent, direction, padding_mode="replicate", verbose=False, pin_loc=None):
        return self.conv.calibrate_spatial_shape(spatial_latent, direction, padding_mode=padding_mode, verbose=verbose,
                                                 pin_loc=pin_loc)

    def get_noise_nch(self):
        return self.conv.out_channel

    def _activate(self, out):
        if isinstance(self.activate, nn.LeakyReLU):
            return leaky_relu(out, self.activate.negative_slope)
        return self.activate(out)

    def forward(self, input, style, noise=None, coords=None, coords_partial=None, test_ids=None, calc_flops=False,
                residual=None):
        """`residual` is an extension over the reference signature (default None keeps it drop-in): the caller's
        `out + shortcut` (models/spgan/spgan.py:169) folded into the conv epilogue on the no_grad path."""
        plain_lrelu = isinstance(self.activate, nn.LeakyReLU)
        fusable = (not calc_flops and self.noise is None and plain_lrelu and not self.upsample and style is not None
                   and style.ndim == 2 and not _grad_needed(input, style, *self.parameters()))
        if fusable:
            out = self.conv.forward_fused(input, style, coords, coords_partial,
                                          act=(self.activate.negative_slope, 1.0), residual=residual)
            return out, 0
        flops = 0
        out, cur = self.conv(input, style, coords=coords, coords_partial=coords_partial, calc_flops=calc_flops)
        flops += cur
        if self.noise is not None:
            out, cur = self.noise(out, noise=noise, test_ids=test_ids, calc_flops=calc_flops)
            flops += cur
        out = self._activate(out)
        if residual is not None:
            out = out + residual
        if calc_flops:
            flops += np.prod(out.shape[1:])
        return out, flops
