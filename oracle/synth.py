"""TEST INFRASTRUCTURE (oracle/): deterministic synthetic tensors shared by make_golden.py, tests and bench.

Weights and inputs are derived from (seed, name) with numpy's PCG64 `default_rng`, whose stream is
stable across numpy versions, so the golden fixtures only need to store OUTPUTS: every consumer can
regenerate bit-identical inputs and weights on any machine.
"""
import zlib

import numpy as np
import torch


def rng_for(seed, name):
    return np.random.default_rng([int(seed), zlib.crc32(name.encode())])


def randn(seed, name, shape, scale=1.0, shift=0.0):
    a = rng_for(seed, name).standard_normal(size=tuple(int(s) for s in shape)).astype(np.float32)
    return (a * np.float32(scale) + np.float32(shift)).astype(np.float32)


def randn_t(seed, name, shape, scale=1.0, shift=0.0):
    return torch.from_numpy(randn(seed, name, shape, scale, shift))


def synthetic_state_dict(manifest, seed=9000):
    """manifest: {key: shape}.  FIR kernels keep their real values; everything else is seeded noise scaled so
    that every term of the forward (noise injection, biases, ToRGB, sphere RGB convs) contributes."""
    sd = {}
    fir = np.outer([1, 2, 1], [1, 2, 1]).astype(np.float32)
    fir = fir / fir.sum() * 4
    for key, shape in manifest.items():
        shape = tuple(shape)
        if key.endswith("blur.kernel") or key.endswith("upsample.kernel"):
            v = fir.copy()
        elif key.endswith("modulation.bias"):
            v = randn(seed, key, shape, 0.1, 1.0)
        elif key.endswith("noise.weight"):
            v = randn(seed, key, shape, 0.1)
        elif key.endswith(".bias") or key.endswith("to_rgbs.0.bias"):
            v = randn(seed, key, shape, 0.1)
        elif "mapping" in key and key.endswith("weight"):
            v = randn(seed, key, shape, 100.0)  # EqualLinear lr_mul=0.01 stores weight / lr_mul
        elif "sp_convs" in key and key.endswith("weight"):
            v = randn(seed, key, shape, 1.0)
        else:
            v = randn(seed, key, shape)
        sd[key] = torch.from_numpy(v)
    return sd
