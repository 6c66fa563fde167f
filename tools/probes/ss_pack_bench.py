"""Times spgan_sphere_pack_seg (and the other structure-chain launches) in isolation at the bench shapes: B = 64 (two lattice
positions of 32), 256 + 3 channels, 35 / 29 / 23 / 17 pixels.  SPGAN_SPHERE_PACK_V1=1 selects the scalar producer."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spgan_b200.functional as SF  # noqa: E402
from spgan_b200 import grids  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
G, Bg, C = 2, 32, 256
B = G * Bg
for H in (35, 29, 23, 17):
    cps = [{"p_x_st": 6 / 65, "p_x_ed": (6 + H + 1) / 65, "p_y_st": 12 / 48, "p_y_ed": (12 + H + 1) / 48, "circular_flag": False,
            "x_total": 65, "y_total": 48, "test_flag": True, "partial": 0.6667},
           {"p_x_st": 12 / 65, "p_x_ed": (12 + H + 1) / 65, "p_y_st": 30 / 48, "p_y_ed": (30 + H + 1) / 48, "circular_flag": True,
            "x_total": 65, "y_total": 48, "test_flag": True, "partial": 0.6667}]
    x = torch.randn(B, C, H, H, device=dev)
    coords = torch.randn(B, 3, H, H, device=dev)
    w = torch.randn(256, C + 3, 3, 3, device=dev)
    s = torch.randn(B, C + 3, device=dev)
    d = SF.demod_coefficients(w, s, 0.02)
    xh, xp = SF.ss_input(x, 1)
    grid = grids.GRID_CACHE.group_grid(H, H, cps, dev)
    y_sc = torch.zeros(B, H, H, 256, device=dev)
    nm = torch.ones(B, 256, device=dev)

    def run():
        return SF.ss_sphere(xh, coords, grid, Bg, w, s, d, 0.02, (0.01, 1.0), y_sc, nm, 1)
    for _ in range(3):
        run()
    SF.profile_calls(True)
    for _ in range(10):
        run()
    t = SF.profile_calls(False)
    rows = B * H * H
    wbytes = 2 * 2 * rows * (9 * 256 + 64)
    print("H=%d  %s" % (H, {k: round(v[0] / v[1], 4) for k, v in t.items()}),
          " pack: %.0f GB/s written" % (wbytes / (t["spgan_sphere_pack_seg"][0] / t["spgan_sphere_pack_seg"][1] * 1e-3) / 1e9))
