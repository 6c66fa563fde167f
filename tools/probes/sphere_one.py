"""One spherical modulated conv at the structure-synthesiser size (B = 32, 256 + 3 channels, 35 x 35) for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle"))
import torch

import spgan_b200.functional as SF
import spgan_b200.lib as lib
from spgan_b200 import grids, panorama

lib.require_device()
B, C, O, h = 32, 256, 256, int(os.environ.get("SPH_H", "35"))
pl = panorama.plan(384, 768)
cp, _ = panorama.patch_inputs(pl, 2, 7, 27, pl["lat_h"], pl["lat_w"])
grid = grids.GRID_CACHE.get(h, h, cp, torch.device("cuda"))
x = torch.randn(B, C, h, h, device="cuda")
c = torch.randn(B, 3, h, h, device="cuda")
w = torch.randn(O, C + 3, 3, 3, device="cuda")
s = torch.randn(B, C + 3, device="cuda") * 0.3 + 1
d = torch.randn(B, O, device="cuda") * 0.3 + 1
for fused in (False, True):
    SF.FUSED_SPHERE_GATHER = fused
    for _ in range(3):
        y = SF.sphere_modconv_fused(x, c, grid, w, s, d, 0.05, act=(0.01, 1.0), precision=int(os.environ.get("SPH_PREC", "1")))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = SF.sphere_modconv_fused(x, c, grid, w, s, d, 0.05, act=(0.01, 1.0), precision=int(os.environ.get("SPH_PREC", "1")))
    e1.record()
    torch.cuda.synchronize()
    print("fused=%s: %.3f ms per call" % (fused, e0.elapsed_time(e1) / 10))
