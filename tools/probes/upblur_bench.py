"""Times spgan_upblur_pack in isolation at the four texture-chain shapes (B = 64: two lattice positions of 32)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spgan_b200.functional as SF  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", "64"))
C = 512
k = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], device=dev) / 4
only = os.environ.get("ONLY")
for Hq in (11, 17, 29, 53):
    if only and int(only) != Hq:
        continue
    zh = 2 * Hq - 1
    oh = zh - 2
    pp = torch.randn(B, 4, Hq, Hq, C, device=dev)
    nz = torch.randn(B, 1, oh, oh, device=dev)
    nw = torch.tensor([0.1], device=dev)
    bias = torch.randn(C, device=dev)
    mul = torch.randn(B, C, device=dev)

    def run():
        return SF.chain_upblur_pack(pp, (zh, zh), k, nz, nw, bias, mul, 1)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    rd = B * zh * zh * C * 4
    wr = B * oh * oh * C * 4
    print("Hq=%d -> %dx%d: %.4f ms, %.0f GB/s (read %.0f MB + write %.0f MB)" % (Hq, oh, oh, ms, (rd + wr) / ms / 1e6, rd / 1e6, wr / 1e6))
