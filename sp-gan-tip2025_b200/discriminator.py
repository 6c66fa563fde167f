"""Discriminator composition over the mirrored op modules, with the reference's parameter names
(models/stylegan2discriminator.py).  Like generator.py this is a caller of the hot path: 1x1 / 3x3 / stride-2 convs
(EqualConv2d), Blur [1,3,3,1] (upfirdn2d) and fused bias + leaky-ReLU; minibatch-stddev and the two linear heads are
small glue."""
import math

import torch
from torch import nn

from . import functional as SF
from .generator import default_config
from .models import ops


class ConvLayer(nn.Sequential):
    """stylegan2discriminator.py:9-54: [Blur] -> EqualConv2d -> [FusedLeakyReLU | ScaledLeakyReLU]."""

    def __init__(self, in_channel, out_channel, kernel_size, downsample=False, blur_kernel=[1, 3, 3, 1], bias=True,
                 activate=True):
        layers = []
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            layers.append(ops.Blur(blur_kernel, pad=((p + 1) // 2, p // 2)))
            stride = 2
            self.padding = 0
        else:
            stride = 1
            self.padding = kernel_size // 2
        layers.append(ops.EqualConv2d(in_channel, out_channel, kernel_size, padding=self.padding, stride=stride,
                                      bias=bias and not activate))
        if activate:
            layers.append(ops.FusedLeakyReLU(out_channel) if bias else ops.ScaledLeakyReLU(0.2))
        super().__init__(*layers)

    def forward(self, input):
        mods = list(self)
        if not ops._grad_needed(input, *self.parameters()) and isinstance(mods[-1], ops.FusedLeakyReLU):
            # inference: bias + leaky-ReLU ride in the conv epilogue
            x = input
            if isinstance(mods[0], ops.Blur):
                x = mods[0](x)
            conv, act = mods[-2], mods[-1]
            k = conv.weight.shape[2]
            geom = SF.ConvGeom(k, k, stride=conv.stride, pad=conv.zero_pad_size)
            return SF.conv_apply(x, conv.weight, geom, out_scale=conv.scale, bias=act.bias,
                                 act=(act.negative_slope, act.scale))
        return super().forward(input)


class ResBlock(nn.Module):
    """stylegan2discriminator.py:57-77."""

    def __init__(self, in_channel, out_channel, kernel_size=3, blur_kernel=[1, 3, 3, 1], downsample=True):
        super().__init__()
        self.conv1 = ConvLayer(in_channel, in_channel, kernel_size)
        self.conv2 = ConvLayer(in_channel, out_channel, kernel_size, downsample=downsample)
        self.skip = ConvLayer(in_channel, out_channel, kernel_size=1, downsample=downsample, activate=False, bias=False)

    def forward(self, input):
        out = self.conv2(self.conv1(input))
        skip = self.skip(input)
        return (out + skip) / math.sqrt(2)


class Discriminator(nn.Module):
    """StyleGan2Discriminator (stylegan2discriminator.py:80-229) for spgan.yaml: patch 101, channel_multiplier 2,
    auxiliary coordinate regressor (coord_use_ac), no projection head."""

    def __init__(self, config=None):
        super().__init__()
        self.config = config if config is not None else default_config()
        tp = self.config.train_params
        size = tp.patch_size
        cm = tp.channel_multiplier
        channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm, 512: 32 * cm,
                    1024: 16 * cm, 2048: 8 * cm}
        linear_ch = 512
        log_size = int(round(math.log(size, 2)))
        convs = [ConvLayer(3, channels[2 ** log_size], kernel_size=1)]
        in_channel = channels[2 ** log_size]
        cur = size
        for i in range(log_size, 2, -1):
            out_channel = channels[2 ** (i - 1)]
            convs.append(ResBlock(in_channel, out_channel, 3, [1, 3, 3, 1]))
            in_channel = out_channel
            cur //= 2
        self.last_feat_ch = in_channel
        self.use_coord_ac = bool(getattr(tp, "coord_use_ac", False))
        self.convs = nn.Sequential(*convs)
        self.stddev_group = self._smallest_divisor_larger_than(tp.batch_size, start=4)
        self.stddev_feat = 1
        self.final_conv = ConvLayer(in_channel + 1, linear_ch, kernel_size=3)
        self.final_linear = nn.Sequential(ops.EqualLinear(linear_ch * cur * cur, linear_ch, activation='fused_lrelu'),
                                          ops.EqualLinear(linear_ch, 1))
        if self.use_coord_ac:
            self.coord_linear = nn.Sequential(ops.EqualLinear(linear_ch * cur * cur, linear_ch, activation='fused_lrelu'),
                                              ops.EqualLinear(linear_ch, tp.coord_num_dir))

    @staticmethod
    def _smallest_divisor_larger_than(number, start):
        for i in range(start, int(math.sqrt(number))):
            if number % i == 0:
                return i
        return number

    def forward(self, input_data, **kwargs):
        img = input_data if isinstance(input_data, torch.Tensor) else input_data["gen" if "gen" in input_data else "patch"]
        h = self.convs(img)
        batch, channel, height, width = h.shape
        group = min(batch, self.stddev_group)
        if h.is_cuda and self.stddev_feat == 1 and batch % group == 0:
            h = SF.minibatch_stddev(h, group)  # fused stddev + concat (csrc/multi_tensor.cu)
            out = self.final_conv(h).view(batch, -1)
            ret = {"d_patch": self.final_linear(out)}
            if self.use_coord_ac:
                ret["ac_coords_pred"] = self.coord_linear(out)
            return ret
        stddev = h.view(group, -1, self.stddev_feat, channel // self.stddev_feat, height, width)
        stddev = torch.sqrt(stddev.var(0, unbiased=False) + 1e-8)
        stddev = stddev.mean([2, 3, 4], keepdims=True).squeeze(2)
        stddev = stddev.repeat(group, 1, height, width)
        h = torch.cat([h, stddev], 1)
        out = self.final_conv(h).view(batch, -1)
        ret = {"d_patch": self.final_linear(out)}
        if self.use_coord_ac:
            ret["ac_coords_pred"] = self.coord_linear(out)
        return ret
