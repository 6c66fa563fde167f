"""One position group (2 lattice positions x 32 patches) of the 384x768 panorama workload, eager launches on one stream,
inside an NVTX range "prof" after two warm-up runs: the target of `ncu --nvtx --nvtx-include "prof/"` (tools/gpu_ncu.sh)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spgan_b200.lib as lib  # noqa: E402
from spgan_b200 import panorama  # noqa: E402
from spgan_b200.generator import Generator  # noqa: E402

B = int(os.environ.get("B", "32"))
group = int(os.environ.get("GROUP", "2"))
torch.cuda.set_device(0)
lib.require_device()
torch.manual_seed(9000)
gen = Generator().cuda().eval()
pl = panorama.plan(384, 768)
pos = panorama.positions(pl)
only = set(pos[24:24 + group])
eng = panorama.PanoramaEngine(gen, pl, B, "cuda:0", streams=1, only=only, use_graph=False, group=group)
g = torch.Generator(device="cpu").manual_seed(9000)
gl = torch.randn(B, 2, 512, generator=g)
gl[:, 1] = gl[:, 0]
eng.load(gl.cuda(), torch.randn(B, 256, pl["lat_h"], pl["lat_w"], generator=g).cuda(),
         [torch.randn(B, 1, pl["noise_h"][l], pl["noise_w"][l], generator=g).cuda() for l in range(8)])
for _ in range(2):
    eng.run()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("prof")
eng.run()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
