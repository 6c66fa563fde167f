"""Mirror of the reference's models/spgan_ops.py.  The shipped generator instantiates only `ToRGB` (and through it
`ModulatedConv2d` 1x1 and `Upsample`) from this module (models/spgan/spgan.py:13-15, 724); `SphereModulatedConv2d` /
`StyledConv` keep their signatures and share the implementation of models/spgan_ops_gs.py."""
from .custom_ops import FusedLeakyReLU, fused_leaky_relu, upfirdn2d  # noqa: F401
from .ops import (Blur, ConstantInput, Downsample, EqualConv2d, EqualLinear, ModulatedConv2d, NoiseInjection,  # noqa: F401
                  PixelNorm, ScaledLeakyReLU, ToRGB, Upsample, create_gaussian_kernel, make_kernel)
from . import spgan_ops_gs as _gs


class SphereModulatedConv2d(_gs.ModulatedConv2d):
    """models/spgan_ops.py:736-1379.  The reference variant samples with the pure-torch gather (true autograd,
    `GridSamplerNewTexture`) and hard-codes `batch * 256` (:1202); it is never instantiated by spgan.yaml.  Here it
    shares the live spherical conv (surrogate gather gradient)."""


class StyledConv(_gs.StyledConv):
    """models/spgan_ops.py:1448-1520."""
