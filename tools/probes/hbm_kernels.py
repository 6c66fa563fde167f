"""One launch of each HBM-bound kernel at its microbenchmark shape (for `ncu --set full`): FIR 3x3 / 4x4, spherical gather,
up / down 2, pack_act."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spgan_b200.functional as SF  # noqa: E402
from spgan_b200 import grids, panorama  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
k3 = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], device=dev) / 4
k4 = torch.tensor([1., 3., 3., 1.], device=dev)
k4 = (k4[None, :] * k4[:, None]) / 64
pl = panorama.plan(384, 768)
cp, _ = panorama.patch_inputs(pl, 2, 7, 27, pl["lat_h"], pl["lat_w"])
grid = torch.from_numpy(grids.sampling_grid(35, 35, cp)).to(dev)
xg = torch.randn(32, 512, 105, 105, device=dev)
xd = torch.randn(32, 256, 101, 101, device=dev)
z = torch.randn(32, 256, 35, 35, device=dev)
xu = torch.randn(32, 64, 53, 53, device=dev)
for rep in range(2):
    with torch.no_grad():
        SF.upfirdn2d(xg, k3, pad=(0, 0))
        SF.upfirdn2d(xd, k4, pad=(2, 2))
        SF.sphere_gather_raw(z, grid)
        SF.upfirdn2d(xu, k4 * 4, up=2, down=1, pad=(2, 1))
        SF.upfirdn2d(xd, k4, up=1, down=2, pad=(1, 1))
torch.cuda.synchronize()
print("ok")
