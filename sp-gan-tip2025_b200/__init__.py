"""spgan-b200: B200-native SP-GAN convolution hot path (hand-written sm_100a CUDA behind a C ABI).

Layout:
  csrc/          CUDA kernels + the C-ABI shared library `libspgan_b200.so` (include/spgan_b200.h)
  lib.py         ctypes binding of that ABI (fails loudly when the library or a B200 is missing)
  functional.py  torch.autograd Functions over the ABI (device memory and streams are torch's; the math is ours)
  grids.py       host-side float64 spherical sampling tables (numpy, cached)
  models/        drop-in mirrors of the reference's `models.custom_ops`, `models.spherenet`, `models.ops`,
                 `models.spgan_ops`, `models.spgan_ops_gs` module signatures
  generator.py   the generator / panorama composition used by bench.py and the parity tests

There is no CPU or PyTorch fallback for any op in this package.
"""
__version__ = "0.1.0"
