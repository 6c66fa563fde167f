"""Mirror of the reference's models/ops.py (plain StyleGAN2 / InfinityGAN op modules): same class names,
constructor and forward signatures, parameter / buffer names, returned tuples.  The arithmetic runs in
libspgan_b200.so: the modulated conv never materialises per-sample weights (the style scales the activations in the
operand packer, the demodulation scales the accumulator in the epilogue) and, under no_grad, noise injection, bias and
leaky-ReLU are fused into the conv epilogue.
"""
import math

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from .. import functional as SF
from ..functional import ConvGeom
from .custom_ops import FusedLeakyReLU, fused_leaky_relu, upfirdn2d


def _grad_needed(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class PixelNorm(nn.Module):
    """models/ops.py:13-20 (tiny (B, 512) elementwise glue in front of the mapping network)."""

    def get_flops(self, input):
        return np.prod(input.shape[1:])

    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    """models/ops.py:24-29."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class Upsample(nn.Module):
    """models/ops.py:32-61.  no_zero_pad: the reference runs a depthwise conv_transpose2d(stride 2) and drops one
    pixel per side, which is upfirdn2d(up=2) with pad (k-2, k-3)."""

    def __init__(self, kernel, factor=2, no_zero_pad=False):
        super().__init__()
        self.no_zero_pad = no_zero_pad
        self.factor = factor
        kernel = make_kernel(kernel) * (factor ** 2)
        self.register_buffer('kernel', kernel)
        if no_zero_pad:
            self.pad = (0, 0)
        else:
            p = kernel.shape[0] - factor
            self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        if self.no_zero_pad:
            k = self.kernel.shape[0]
            return upfirdn2d(input, self.kernel, up=2, down=1, pad=(k - 2, k - 3))
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """models/ops.py:64-79."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel)
        self.register_buffer('kernel', kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


def create_gaussian_kernel(kernel_size, std=1):
    """models/ops.py:82-85 (scipy.signal.gaussian no longer exists; same window written out)."""
    n = np.arange(kernel_size) - (kernel_size - 1) / 2.0
    g = np.exp(-0.5 * (n / std) ** 2).reshape(kernel_size, 1)
    k = np.outer(g, g)
    return k / k.sum()


class Blur(nn.Module):
    """models/ops.py:88-140."""

    def __init__(self, kernel, pad, upsample_factor=1, padding_mode="zero", prior="gaussian"):
        super().__init__()
        if isinstance(kernel, int):
            if prior.lower() == "gaussian":
                kernel = create_gaussian_kernel(kernel_size=kernel)
            elif prior.lower() == "mean":
                kernel = torch.ones(kernel, kernel, dtype=torch.float32)
            else:
                raise NotImplementedError("Unknown prior {}".format(prior))
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer('kernel', kernel)
        if padding_mode == "replicate":
            self.zero_pad = (0, 0)
            self.replicate_pad = pad if isinstance(pad, tuple) else (pad, pad, pad, pad)
            self.use_replicate_pad = True
        elif padding_mode == "zero":
            self.zero_pad = pad if isinstance(pad, tuple) else (pad, pad)
            self.replicate_pad = 0
            self.use_replicate_pad = False
        else:
            raise NotImplementedError("Unknown padding_mode {}".format(padding_mode))
        self.upsample_factor = upsample_factor

    def get_output_shape(self, shape):
        B, C, H, W = shape
        ks = self.kernel.shape[0]
        if self.use_replicate_pad:
            H += self.replicate_pad[2] + self.replicate_pad[3]
            W += self.replicate_pad[0] + self.replicate_pad[1]
        else:
            H += self.zero_pad[0]
            W += self.zero_pad[1]
        return B, C, H - ks // 2 * 2, W - ks // 2 * 2

    def get_flops(self, shape):
        _, C, _, _ = shape
        _, _, oh, ow = self.get_output_shape(shape)
        return oh * ow * C * (self.kernel.shape[0] ** 2)

    def forward(self, input):
        if self.use_replicate_pad:
            input = F.pad(input, self.replicate_pad, mode="replicate")
        return upfirdn2d(input, self.kernel, pad=self.zero_pad)


class EqualConv2d(nn.Module):
    """models/ops.py:143-187 (discriminator convs)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding_type = padding
        self.extra_padding_layer = None
        if type(padding) is str:
            if padding == "reflect":
                self.extra_padding_layer = nn.ReflectionPad2d(kernel_size // 2)
                self.zero_pad_size = 0
            elif padding == "zero":
                self.zero_pad_size = kernel_size // 2
            else:
                raise NotImplementedError("Unknown padding type {}".format(padding))
        else:
            self.zero_pad_size = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        if self.extra_padding_layer is not None:
            input = self.extra_padding_layer(input)
        k = self.weight.shape[2]
        geom = ConvGeom(k, k, stride=self.stride, pad=self.zero_pad_size)
        if _grad_needed(input, self.weight, self.bias):
            out = SF.conv2d(input, self.weight, geom, out_scale=self.scale)
            if self.bias is not None:
                out = out + self.bias.view(1, -1, 1, 1)
            return out
        return SF.conv_apply(input, self.weight, geom, out_scale=self.scale, bias=self.bias)

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]},'
                f' {self.weight.shape[2]}, stride={self.stride}, padding_type={self.padding_type})')


class EqualLinear(nn.Module):
    """models/ops.py:190-222."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def get_flops(self, input):
        flops = 0
        if self.activation:
            flops += self.bias.shape[0] * 2
        flops += np.prod(self.weight.shape) * 2
        flops += self.bias.shape[0] * 2
        return flops

    def forward(self, input):
        return SF.equal_linear(input, self.weight, self.bias, self.scale, self.lr_mul, bool(self.activation))

    def __repr__(self):
        return f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})'


class ScaledLeakyReLU(nn.Module):
    """models/ops.py:225-232."""

    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return _scaled_lrelu(input, self.negative_slope)


class _ScaledLReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, slope, scale):
        out = SF.bias_act(x, None, None, 3, 0, slope, scale)
        ctx.save_for_backward(out)
        ctx.slope, ctx.scale = slope, scale
        return out

    @staticmethod
    def backward(ctx, g):
        out, = ctx.saved_tensors
        return _ScaledLReLUGradFn.apply(g, out, ctx.slope, ctx.scale), None, None


class _ScaledLReLUGradFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, out, slope, scale):
        ctx.save_for_backward(out)
        ctx.slope, ctx.scale = slope, scale
        return SF.bias_act(g, None, out, 3, 1, slope, scale)

    @staticmethod
    def backward(ctx, gg):
        out, = ctx.saved_tensors
        return SF.bias_act(gg, None, out, 3, 1, ctx.slope, ctx.scale), None, None, None


def _scaled_lrelu(x, slope, scale=math.sqrt(2)):
    return _ScaledLReLUFn.apply(x, slope, scale)


def leaky_relu(x, slope=0.01):
    """nn.LeakyReLU() of the spherical blocks (models/spgan_ops_gs.py:1085-1086) through K1 (no bias, gain 1)."""
    return _ScaledLReLUFn.apply(x, slope, 1.0)


class ModulatedConv2d(nn.Module):
    """models/ops.py:235-729."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 2, 1], no_zero_pad=False, config=None, side=None):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.style_dim = style_dim
        self.upsample = upsample
        self.downsample = downsample
        self.no_zero_pad = no_zero_pad
        self.config = config
        self.side = side
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        if upsample:
            factor = 2
            assert kernel_size == 3, "only 3x3 kernels are supported for the upsampling conv (as in the reference)"
            if len(blur_kernel) % 2 == 1:
                pad0 = pad1 = len(blur_kernel) // 2
            else:
                p = (len(blur_kernel) - factor) - (kernel_size - 1)
                pad0 = (p + 1) // 2 + factor - 1
                pad1 = p // 2 + 1
            if no_zero_pad:
                self.dirty_rm_size = (pad0, pad1)
                self.blur = Blur(blur_kernel, pad=(0, 0), upsample_factor=factor)
            else:
                self.dirty_rm_size = (0, 0)
                self.blur = Blur(blur_kernel, pad=(pad0, pad1), upsample_factor=factor)
        elif downsample:
            raise NotImplementedError("Never used.")
        else:
            if no_zero_pad:
                self.padding = 0
                self.dirty_rm_size = (kernel_size // 2, kernel_size // 2)
            else:
                self.padding = kernel_size // 2
                self.dirty_rm_size = (0, 0)
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.demodulate = demodulate
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1) if style_dim > 0 else None

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, '
                f'upsample={self.upsample}, downsample={self.downsample})')

    # ---- spatial bookkeeping used by the test managers (base_test_manager.py:133-145) ----
    def calc_in_spatial_size(self, out_spatial_size, verbose=False):
        """models/ops.py:313-336."""
        if self.upsample:
            v = out_spatial_size + 1 + self.dirty_rm_size[0] + self.dirty_rm_size[1]
            return (v if v % 2 == 0 else v + 1) // 2
        return out_spatial_size + self.dirty_rm_size[0] + self.dirty_rm_size[1]

    def calc_out_spatial_size(self, in_spatial_size):
        """models/ops.py:338-349."""
        if self.upsample:
            return in_spatial_size * 2 - 1 - self.dirty_rm_size[0] - self.dirty_rm_size[1]
        return in_spatial_size - self.dirty_rm_size[0] - self.dirty_rm_size[1]

    def calibrate_spatial_shape(self, feature, direction, padding_mode="replicate", verbose=False, pin_loc=None):
        """models/ops.py:352-489: geometric re-alignment of a spatial latent across this layer (style fusion /
        interactive tools; not on the generation path).  Host-side glue on small tensors."""
        _, _, h, w = feature.shape
        d0, d1 = self.dirty_rm_size
        if direction == "forward":
            if self.upsample:
                nh, nw = h * 2 - 1, w * 2 - 1
                feature = F.interpolate(feature, size=[nh, nw], mode="bilinear", align_corners=True)[:, :, 1:-1, 1:-1]
                if pin_loc is not None:
                    pin_loc = [(pin_loc[0] - h // 2) * 2 + nh // 2, (pin_loc[1] - w // 2) * 2 + nw // 2]
            elif self.downsample:
                raise NotImplementedError("Never used.")
            else:
                if self.padding == 0:
                    assert d0 != 0 and d1 != 0
                    feature = feature[:, :, d0:-d0, d1:-d1]
                if pin_loc is not None:
                    pin_loc = [pin_loc[0] - d0, pin_loc[1] - d1]
        elif direction == "backward":
            rec = (self.calc_in_spatial_size(h), self.calc_in_spatial_size(w))
            if self.upsample:
                if self.dirty_rm_size != (0, 0):
                    feature = F.pad(feature, (d1, d1, d0, d0), mode=padding_mode)
                feature = F.interpolate(feature, size=rec, mode="bilinear", align_corners=True)
                if pin_loc is not None:
                    pin = [pin_loc[0] + d0, pin_loc[1] + d1]
                    old_c = [h + d0, w + d1]
                    new_c = [old_c[0] // 2, old_c[1] // 2]
                    pin_loc = [(pin[0] - old_c[0]) // 2 + new_c[0], (pin[1] - old_c[1]) // 2 + new_c[1]]
            elif self.downsample:
                raise NotImplementedError("Never used.")
            else:
                if self.padding == 0:
                    feature = F.pad(feature, (d1, d1, d0, d0), mode=padding_mode)
                if pin_loc is not None:
                    pin_loc = [pin_loc[0] + d0, pin_loc[1] + d1]
        else:
            raise NotImplementedError("Unknown direction {} (valid: 'forward' or 'backward')".format(direction))
        return feature, pin_loc

    def get_flops(self, input, style):
        """Analytic MAC-style counter in the spirit of models/ops.py:502-577 (used only by --calc-flops)."""
        B, C, H, W = input.shape
        k = self.kernel_size
        if self.upsample:
            oh, ow = H * 2 + 1, W * 2 + 1
        else:
            oh, ow = H + 2 * self.padding - k + 1, W + 2 * self.padding - k + 1
        flops = self.out_channel * C * k * k * oh * ow * 2
        if self.modulation is not None:
            flops += self.modulation.get_flops(style)
        return flops

    # ---- the op ----
    def _geom(self):
        k = self.kernel_size
        if self.upsample:
            return ConvGeom(k, k, stride=2, transposed=True, crop=1 if self.no_zero_pad else 0)
        return ConvGeom(k, k, stride=1, pad=self.padding)

    def _mod_demod(self, style, batch):
        """s = modulation(style) (B, Cin); d = rsqrt(scale^2 sum_c s^2 sum_t w^2 + eps) (B, Cout) or None
        (models/ops.py:598-604), with autograd when needed.

        Under no_grad the pair is memoised per module for as long as the SAME style storage (kept alive by the cache,
        so its address cannot be recycled) and the same parameter versions are presented: a panorama runs 60 patch
        calls with one global latent, and the reference recomputes these 20 small GEMMs in every call."""
        w = self.weight[0]
        if _grad_needed(style, self.weight, self.modulation.weight, self.modulation.bias):
            s = self.modulation(style).view(batch, self.in_channel)
            d = None
            if self.demodulate:
                wsq = (w * w).sum(dim=(2, 3))
                d = torch.rsqrt(SF._LinearFn.apply(s * s, wsq, None, self.scale * self.scale, 1.0) + 1e-8)
            return s, w, d
        key = self._md_key(style)
        cache = self.__dict__.setdefault("_md_cache", {})
        cached = cache.get(key)
        if cached is not None:
            return cached[1], w, cached[2]
        s = self.modulation(style).view(batch, self.in_channel)
        d = SF.demod_coefficients(w, s, self.scale, 1e-8) if self.demodulate else None
        self._md_store(key, style, s, d)
        return s, w, d

    def _md_key(self, style):
        return (SF.epoch(), style.data_ptr(), style._version, tuple(style.shape), tuple(style.stride()), str(style.device),
                self.weight._version, self.modulation.weight._version,
                self.modulation.bias._version if self.modulation.bias is not None else -1,
                self.weight.data_ptr(), self.modulation.weight.data_ptr())

    def _md_store(self, key, style, s, d):
        """Memoise a (modulation, demodulation) pair for `style` (also used by Generator.prepare_modulation, which computes the
        pairs of all layers in one launch)."""
        cache = self.__dict__.setdefault("_md_cache", {})
        # several live styles per module (one per position-group size of a panorama engine); entries of an older epoch are
        # dead.  Never evict a live entry: a concurrent branch of a captured graph may only READ what the launching stream
        # computed before the fork (panorama.PanoramaEngine._body)
        for k in [k for k in cache if k[0] != key[0]]:
            del cache[k]
        while len(cache) >= 16:
            del cache[next(iter(cache))]  # oldest first
        cache[key] = (style, s, d)

    def weight_sq(self):
        """(Cout, Cin) sum over the taps of W^2 (the demodulation's weight factor, models/ops.py:603), cached per weight version."""
        w = self.weight[0]
        key = (SF.epoch()[0], self.weight._version, self.weight.data_ptr())
        hit = self.__dict__.get("_wsq_cache")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, (w * w).sum(dim=(2, 3)).contiguous())
            self.__dict__["_wsq_cache"] = hit
        return hit[1]

    def forward(self, input, style, coords=None, calc_flops=False):
        batch = input.shape[0]
        flops = self.get_flops(input, style) if calc_flops else 0
        if style is not None and style.ndim == 4:
            mean_style = style.mean([2, 3], keepdim=True)
            if ((style - mean_style) < 1e-8).all():
                style = mean_style.squeeze()
        if style.ndim != 2:
            return self._forward_spatial_style(input, style), flops
        s, w, d = self._mod_demod(style, batch)
        out = SF.conv2d(input, w, self._geom(), in_mul=s, out_mul=d, out_scale=self.scale)
        if self.upsample:
            out = self.blur(out)
        return out, flops

    def _forward_spatial_style(self, input, style):
        """Spatially-shaped styles, test-time style fusion (models/ops.py:637-729): the style modulates the ACTIVATIONS per
        pixel, the conv runs with the un-modulated weight, and the demodulation is the per-pixel estimate
        rsqrt(sum_c (sum_t (scale W)^2)[o, c] * s[b, c, y, x]^2 + 1e-8), bilinearly resized for the upsampling conv and — a
        quirk of the reference kept here — applied to the plain conv only when it is unpadded (:721-723)."""
        assert not self.training, "Only accepts spatially-shaped global-latent for testing-time manipulation!"
        assert style.ndim == 4, "Only considered BxCxHxW case, but got shape {}".format(style.shape)
        assert style.shape[2] >= input.shape[2] and style.shape[3] >= input.shape[3]
        assert (style.shape[2] - input.shape[2]) % 2 == 0 and (style.shape[3] - input.shape[3]) % 2 == 0
        ph, pw = (style.shape[2] - input.shape[2]) // 2, (style.shape[3] - input.shape[3]) // 2
        style = style[:, :, ph:ph + input.shape[2], pw:pw + input.shape[3]]
        sb, sc, sh, sw = style.shape
        flat = style.permute(0, 2, 3, 1).reshape(-1, sc)
        smod_flat = self.modulation(flat)                                    # (B*H*W, Cin) through the linear kernel
        style_mod = smod_flat.view(sb, sh, sw, self.in_channel).permute(0, 3, 1, 2)
        input_st = style_mod * input
        w = self.weight[0]
        demod = None
        if self.demodulate:
            wsq = (w * w).sum(dim=(2, 3))                                       # (Cout, Cin)
            dflat = torch.rsqrt(SF._LinearFn.apply(smod_flat * smod_flat, wsq, None, self.scale * self.scale, 1.0) + 1e-8)
            demod = dflat.view(sb, sh, sw, self.out_channel).permute(0, 3, 1, 2)
        # the reference's spatial branch crops the transposed conv by one pixel whatever no_zero_pad says (:708)
        geom = ConvGeom(self.kernel_size, self.kernel_size, stride=2, transposed=True, crop=1) if self.upsample else self._geom()
        out = SF.conv2d(input_st, w, geom, out_scale=self.scale)
        if self.upsample:
            if demod is not None:
                out = out * F.interpolate(demod, size=(out.shape[2], out.shape[3]), mode="bilinear", align_corners=True)
            out = self.blur(out)
        elif self.downsample:
            raise NotImplementedError("Never used.")
        elif demod is not None and self.padding == 0:
            d0, d1 = self.dirty_rm_size
            out = out * demod[:, :, d0:-d0, d1:-d1]
        return out.contiguous()

    def forward_fused(self, input, style, noise, noise_weight, act_bias, act=(0.2, 2 ** 0.5)):
        """no_grad fast path for StyledConv: conv + noise + bias + leaky-ReLU in the GEMM epilogue (plain conv), or
        conv -> FIR -> fused noise/bias/act (upsampling conv)."""
        batch = input.shape[0]
        s, w, d = self._mod_demod(style, batch)
        if self.upsample:
            geom = self._geom()
            if tuple(self.blur.kernel.shape) == (3, 3) and self.blur.zero_pad == (0, 0) and not self.blur.use_replicate_pad:
                # polyphase transposed conv -> one fused interleave + FIR + noise + bias + leaky-ReLU kernel
                pp = SF.conv_apply(input, w, geom, in_mul=s, out_mul=d, out_scale=self.scale, polyphase=True)
                return SF.upblur_act(pp, self.blur.kernel, geom.out_size(input.shape[2], input.shape[3]), noise,
                                     noise_weight, act_bias, act[0], act[1])
            out = SF.conv_apply(input, w, geom, in_mul=s, out_mul=d, out_scale=self.scale)
            out = self.blur(out)
            return SF.noise_bias_act(out, noise, noise_weight, act_bias, act[0], act[1])
        return SF.conv_apply(input, w, self._geom(), in_mul=s, out_mul=d, out_scale=self.scale, noise=noise,
                             noise_w=noise_weight, bias=act_bias, act=act)


class NoiseInjection(nn.Module):
    """models/ops.py:732-785."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))
        self.testing_noise = {}

    def resolve_noise(self, image, noise=None, test_ids=None):
        """The noise tensor the reference would add (fixed per-test-id noise, given noise, or fresh normal noise)."""
        if (not self.training) and (test_ids is not None):
            assert noise is None, "`test_ids` and `noise` are mutually exclusive!"
            batch, _, height, width = image.shape
            assert len(test_ids) == batch
            cur = []
            for test_id in test_ids:
                test_id = test_id.item() if isinstance(test_id, torch.Tensor) else test_id
                if test_id not in self.testing_noise:
                    t = image.new_empty(1, height, width).normal_()
                    self.testing_noise[test_id] = t.cpu().detach()
                    ch, cw = height, width
                else:
                    t = self.testing_noise[test_id].detach().to(image.device)
                    _, ch, cw = t.shape
                    if ch < height or cw < width:
                        nt = image.new_empty(1, height, width).normal_()
                        ph, pw = (height - ch) // 2, (width - cw) // 2
                        nt[:, ph:ph + ch, pw:pw + cw] = t
                        self.testing_noise[test_id] = nt
                        t = nt
                        ch, cw = height, width
                ph, pw = (ch - height) // 2, (cw - width) // 2
                cur.append(t[:, ph:ph + height, pw:pw + width])
            noise = torch.stack(cur)
        elif noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return noise

    def forward(self, image, noise=None, test_ids=None, calc_flops=False):
        noise = self.resolve_noise(image, noise, test_ids)
        flops = np.prod(image.shape[1:]) * 2 if calc_flops else 0
        return image + self.weight * noise, flops


class ConstantInput(nn.Module):
    """models/ops.py:788-795."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, batch_size):
        return self.input.repeat(batch_size, 1, 1, 1)


class StyledConv(nn.Module):
    """models/ops.py:798-863."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 2, 1],
                 demodulate=True, no_zero_pad=False, disable_noise=False, activation="LeakyReLU", config=None,
                 side=None):
        super().__init__()
        self.no_zero_pad = no_zero_pad
        self.upsample = upsample
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate, no_zero_pad=no_zero_pad,
                                    config=config, side=side)
        self.noise = None if disable_noise else NoiseInjection()
        if activation.lower() == "leakyrelu":
            self.activate = FusedLeakyReLU(out_channel)
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

    def forward(self, input, style, noise=None, coords=None, test_ids=None, calc_flops=False):
        fusable = (not calc_flops and style is not None and style.ndim == 2
                   and not _grad_needed(input, style, noise, *self.parameters()))
        if fusable:
            nz = nw = None
            if self.noise is not None:
                B = input.shape[0]
                oh = self.conv.calc_out_spatial_size(input.shape[2])
                ow = self.conv.calc_out_spatial_size(input.shape[3])
                nz = self.noise.resolve_noise(input.new_empty(B, 1, oh, ow), noise, test_ids)
                nw = self.noise.weight
            out = self.conv.forward_fused(input, style, nz, nw, self.activate.bias,
                                          (self.activate.negative_slope, self.activate.scale))
            return out, 0
        flops = 0
        out, cur = self.conv(input, style, coords=coords, calc_flops=calc_flops)
        flops += cur
        if self.noise is not None:
            out, cur = self.noise(out, noise=noise, test_ids=test_ids, calc_flops=calc_flops)
            flops += cur
        out = self.activate(out)
        if calc_flops:
            flops += np.prod(out.shape[1:])
        return out, flops


class ToRGB(nn.Module):
    """models/ops.py:866-930 / models/spgan_ops.py:1523-1586."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 2, 1], no_zero_pad=False, config=None,
                 side=None):
        super().__init__()
        self.no_zero_pad = no_zero_pad
        if upsample:
            self.upsample = Upsample(blur_kernel, no_zero_pad=no_zero_pad)
        self.conv = ModulatedConv2d(in_channel=in_channel, out_channel=3, kernel_size=1, style_dim=style_dim,
                                    demodulate=False, no_zero_pad=no_zero_pad, config=config, side=side)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def align_spatial_size(self, source, target):
        if source is None:
            return source
        _, _, cH, cW = target.shape
        _, _, sH, sW = source.shape
        if (cH == sH) and (cW == sW):
            return source
        assert ((sH - cH) % 2 == 0) and ((sW - cW) % 2 == 0), \
            "Should always have equal padding on two sides, got target ({}x{}) and source ({}x{})".format(cH, cW, sH, sW)
        h_st, w_st = (sH - cH) // 2, (sW - cW) // 2
        return source[:, :, h_st:h_st + cH, w_st:w_st + cW]

    def forward(self, input, style, skip=None, coords=None, calc_flops=False):
        flops = 0
        if style.ndim == 4:
            style = self.align_spatial_size(style, target=input)
        if coords is not None:
            coords = self.align_spatial_size(coords, target=input)
        fusable = (not calc_flops and style.ndim == 2 and not _grad_needed(input, style, skip, *self.parameters()))
        if fusable:
            conv = self.conv
            s, w, d = conv._mod_demod(style, input.shape[0])
            res = None
            if skip is not None:
                res = self.upsample(skip)
                if self.no_zero_pad:
                    oh, ow = conv.calc_out_spatial_size(input.shape[2]), conv.calc_out_spatial_size(input.shape[3])
                    res = self.align_spatial_size(res, target=torch.empty(1, 1, oh, ow, device="meta"))
            out = SF.conv_apply(input, w, conv._geom(), in_mul=s, out_mul=d, out_scale=conv.scale,
                                bias=self.bias.view(-1), residual=res)
            return out, 0
        out, cur = self.conv(input, style, coords=coords, calc_flops=calc_flops)
        flops += cur
        out = out + self.bias
        if calc_flops:
            flops += np.prod(out.shape[1:])
        if skip is not None:
            skip = self.upsample(skip)
            if self.no_zero_pad:
                skip = self.align_spatial_size(skip, target=out)
            out = out + skip
            if calc_flops:
                flops += np.prod(out.shape[1:])
        return out, flops
