"""Mirror of models/spherenet/__init__.py:1-6 of the reference."""
from .grid_generator import (GridGenerator, GridGeneratorPatchCoordsFixBorder, GridSampler, GridSamplerFuncNoGrad,  # noqa: F401
                             GridSamplerNew, GridSamplerNewTexture, GridSamplerNewTextureNoGrad, IncreIntervalGridGenerator)
from .sphere_conv2d import IncreIntervalSphereConv2d, SphereConv2d, SphereConvBatchDiffFixBorderGNoGrad  # noqa: F401
