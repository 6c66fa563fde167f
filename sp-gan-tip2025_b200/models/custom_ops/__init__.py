"""Mirror of models/custom_ops/__init__.py:1-2 of the reference: `upfirdn2d`, `fused_leaky_relu`, `FusedLeakyReLU`,
backed by libspgan_b200.so (K1-K3 replacements).  Unlike the reference there is no native-PyTorch CPU branch
(models/custom_ops/upfirdn2d.py:151-154, fused_act.py:91-98): a CPU tensor raises."""
import torch
from torch import nn

from ...functional import (FusedLeakyReLUFunction, FusedLeakyReLUFunctionBackward, UpFirDn2d, UpFirDn2dBackward,  # noqa: F401
                           bias_act as fused_bias_act, fused_leaky_relu, upfirdn2d)


class FusedLeakyReLU(nn.Module):
    """models/custom_ops/fused_act.py:78-88."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
