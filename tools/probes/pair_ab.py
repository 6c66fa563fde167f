"""CTA-pair GEMM kernel on / off per conv shape (CUDA events, conv_apply incl. its pack): panorama shapes (B = 64) and training
shapes (B = 8).  Prints ms for pair mode 0 and 2 and the M-tile counts, to calibrate the automatic choice."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spgan_b200.functional as SF  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
shapes = []
for B in (64, 8):
    shapes += [(B, 512, 512, 103, 3, "conv"), (B, 512, 512, 55, 3, "conv"), (B, 512, 512, 31, 3, "conv"), (B, 512, 512, 19, 3, "conv"),
               (B, 512, 512, 53, 3, "up"), (B, 512, 512, 29, 3, "up"), (B, 512, 512, 17, 3, "up"), (B, 256, 512, 11, 3, "up"),
               (B, 259, 256, 35, 7, "conv"), (B, 259, 256, 23, 7, "conv"), (B, 259, 256, 17, 7, "conv"),
               (B, 256, 256, 35, 1, "conv"), (B, 256, 512, 101, 3, "d_s2"), (B, 512, 512, 50, 3, "d_s2"), (B, 512, 512, 25, 3, "conv_p1")]
for B, C, O, H, k, kind in shapes:
    x = torch.randn(B, C, H, H, device=dev)
    w = torch.randn(O, C, k, k, device=dev)
    if kind == "up":
        geom = SF.ConvGeom(k, k, stride=2, transposed=True, crop=1)
    elif kind == "d_s2":
        geom = SF.ConvGeom(k, k, stride=2, pad=0)
    elif kind == "conv_p1":
        geom = SF.ConvGeom(k, k, pad=1)
    else:
        geom = SF.ConvGeom(k, k)
    res = []
    for mode in (0, 2):
        SF.set_gemm_pair_mode(mode)
        for _ in range(2):
            SF.conv_apply(x, w, geom, precision=1)
        SF.profile_gemm(True)
        for _ in range(5):
            SF.conv_apply(x, w, geom, precision=1)
        st = SF.profile_gemm(False)
        res.append(st["ms"] / 5)
    oh = geom.out_size(H, H)[0]
    print("B=%2d C=%3d O=%3d H=%3d k=%d %-7s out %3d rows %7d m_tiles %5d: single %.4f ms  pair %.4f ms  ratio %.3f" % (
        B, C, O, H, k, kind, oh, B * oh * oh, -(-B * oh * oh // 128), res[0], res[1], res[1] / res[0]))
SF.set_gemm_pair_mode(1)
