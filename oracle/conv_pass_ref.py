"""TEST INFRASTRUCTURE (oracle/): numpy restatement of the `SpganConvPass` contract of include/spgan_b200.h.

Used on the CPU to check that the host-side pass planner (`spgan_b200.functional.plan_passes`) decomposes every conv
of the path — F.conv2d(stride, padding), F.conv_transpose2d(stride 2) + crop (reference models/ops.py:617-619, 634,
175) and their data gradients — into passes that reproduce PyTorch's own results.  Never imported by product code.
"""
import numpy as np


def run_pass(p, y, x, w_flat, ws_o, ws_c, Cout, in_mul=None, out_mul=None, out_scale=1.0):
    """Accumulate one pass into y (B, Cout, oh, ow) in float64.  p: dict from plan_passes."""
    B, Cin, H, W = x.shape
    oh, ow = y.shape[2], y.shape[3]
    xs = x.astype(np.float64)
    if in_mul is not None:
        xs = xs * in_mul[:, :, None, None]
    for i in range(p["My"]):
        Y = i * p["out_stride"] + p["off_y"]
        if Y < 0 or Y >= oh:
            continue
        for j in range(p["Mx"]):
            X = j * p["out_stride"] + p["off_x"]
            if X < 0 or X >= ow:
                continue
            acc = np.zeros((B, Cout))
            for dy, dx, wi in p["taps"]:
                yy, xx = i * p["in_stride"] + dy, j * p["in_stride"] + dx
                if yy < 0 or yy >= H or xx < 0 or xx >= W:
                    continue
                wt = np.array([[w_flat[o * ws_o + c * ws_c + wi] for c in range(Cin)] for o in range(Cout)])
                acc += xs[:, :, yy, xx] @ wt.T
            v = acc * out_scale
            if out_mul is not None:
                v = v * out_mul
            y[:, :, Y, X] = v
    return y
