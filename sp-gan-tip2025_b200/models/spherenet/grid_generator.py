"""Mirror of models/spherenet/grid_generator.py: sampling-pattern generators (host numpy, float64) and the sampler
modules (device gather kernels of libspgan_b200.so)."""
import torch
from torch import nn

from ... import functional as SF
from ... import grids


def _as_pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class GridGenerator:
    """Full-sphere SphereNet pattern (grid_generator.py:12-108).  Not reached by configs/model/spgan.yaml."""

    def __init__(self, height, width, kernel_size, stride=1):
        if isinstance(height, torch.Tensor):
            height, width = int(height), int(width)
        self.height, self.width = height, width
        self.kernel_size = _as_pair(kernel_size)
        self.stride = _as_pair(stride)

    def createSamplingPattern(self):
        return grids.full_sphere_pattern(self.height, self.width, self.kernel_size, self.stride)


class IncreIntervalGridGenerator:
    """grid_generator.py:385-582: full-sphere pattern on re-spread row / column centres (optionally the upsampling variant)."""

    def __init__(self, height, width, kernel_size, stride=1, upsample=False):
        if isinstance(height, torch.Tensor):
            height, width = int(height), int(width)
        self.height, self.width, self.upsample = height, width, upsample
        self.kernel_size = _as_pair(kernel_size)
        self.stride = _as_pair(stride)

    def createSamplingPattern(self):
        return grids.incre_interval_pattern(self.height, self.width, self.kernel_size, self.stride, self.upsample)


class GridGeneratorPatchCoordsFixBorder:
    """Patch pattern parameterised by `coords_partial` (grid_generator.py:111-352)."""

    def __init__(self, height, width, kernel_size, stride=1, coords_partial=None):
        if isinstance(height, torch.Tensor):
            height, width = int(height), int(width)
        self.height, self.width = height, width
        self.kernel_size = _as_pair(kernel_size)
        self.stride = _as_pair(stride)
        self.coords_partial = coords_partial
        assert self.coords_partial is not None
        if self.kernel_size != (3, 3) or self.stride != (1, 1):
            raise NotImplementedError("only the 3x3 / stride-1 pattern of spgan.yaml is implemented")

    def createSamplingPattern(self):
        """(1, 3H, 3W, 2) float64, last dim (lat, lon) in grid units."""
        return grids.sampling_pattern_dense(self.height, self.width, self.coords_partial)


GridSamplerFuncNoGrad = SF.GridSamplerFuncNoGrad


class GridSamplerNewTextureNoGrad(nn.Module):
    """grid_generator.py:602-607: bilinear / border / align_corners gather with the surrogate backward."""

    def forward(self, z, grid):
        return SF.sphere_gather(z, grid)


class GridSamplerNew(nn.Module):
    """grid_generator.py:588-592: F.grid_sample(bilinear, border, align_corners=True) with its true input gradient."""

    def forward(self, z, grid):
        return SF.grid_sample(z, grid, "bilinear_border")


class GridSamplerNewTexture(nn.Module):
    """grid_generator.py:595-599 -> grid_sample_github (grid_sample_ops.py:5-55): bilinear weights from the unclipped
    coordinate, corner indices clamped, true gradient (autograd through torch.gather in the reference)."""

    def forward(self, z, grid):
        return SF.grid_sample(z, grid, "texture")


class GridSampler(nn.Module):
    """grid_generator.py:580-585 -> grid_sample_grad_fix.grid_sample: F.grid_sample(nearest, zeros, align_corners=True) with
    aten::grid_sampler_2d_backward's input gradient and a double backward (grid_sample_grad_fix.py:29-88)."""

    def forward(self, z, grid):
        return SF.grid_sample(z, grid, "nearest_zeros")
