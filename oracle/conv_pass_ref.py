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


# --------------------------------------------------------------------------------------------
# numpy restatement of the PACKED formulation the tcgen05 kernels consume (include/spgan_b200.h:
# spgan_pack_act / spgan_conv_gemm / spgan_conv_wgrad_gemm), used by tests/test_planner.py to check the host-side
# geometry (lattice sizes, polyphase tap mapping, flat row offsets, im2col bounding boxes) without a GPU.
# --------------------------------------------------------------------------------------------
def pack_lattice(x, mul, step, pad_y, pad_x, Hl, Wl):
    """spgan_pack_act: (B, C, H, W) -> (step*step, B, Hl, Wl, C); lattice point (i, j) of phase py*step + px holds pixel
    (i*step + py - pad_y, j*step + px - pad_x), zero outside the image."""
    B, C, H, W = x.shape
    out = np.zeros((step * step, B, Hl, Wl, C))
    for py in range(step):
        for px in range(step):
            for i in range(Hl):
                sy = i * step + py - pad_y
                if sy < 0 or sy >= H:
                    continue
                for j in range(Wl):
                    sx = j * step + px - pad_x
                    if 0 <= sx < W:
                        v = x[:, :, sy, sx]
                        out[py * step + px, :, i, j, :] = v * mul if mul is not None else v
    return out


def gemm_flat(pack, taps, wsel, My, Mx):
    """Flat mode of spgan_conv_gemm: the pack is a matrix of rows (phase, b, i, j); tap (ph, oy, ox) reads row
    r + (ph*B*Hl + oy)*Wl + ox (zero beyond the matrix, wrap-around inside it); lattice points outside My x Mx are dropped.
    wsel[t] is the (O, C) weight slice of tap t.  Returns (B, My, Mx, O)."""
    P, B, Hl, Wl, C = pack.shape
    rows = pack.reshape(P * B * Hl * Wl, C)
    n = B * Hl * Wl
    acc = np.zeros((n, wsel[0].shape[0]))
    for (ph, oy, ox), w in zip(taps, wsel):
        off = (ph * B * Hl + oy) * Wl + ox
        src = np.zeros((n, C))
        hi = min(n, rows.shape[0] - off)
        if hi > 0:
            src[:hi] = rows[off:off + hi]
        acc += src @ w.T
    return acc.reshape(B, Hl, Wl, -1)[:, :My, :Mx]


def gemm_im2col(pack, taps, wsel, My, Mx):
    """im2col mode: base pixels are the My x Mx outputs of each image; tap (ph, oy, ox) reads pixel (i + oy, j + ox) of
    phase plane ph, zero outside the (Hl, Wl) image."""
    P, B, Hl, Wl, C = pack.shape
    out = np.zeros((B, My, Mx, wsel[0].shape[0]))
    for (ph, oy, ox), w in zip(taps, wsel):
        src = np.zeros((B, My, Mx, C))
        h = max(0, min(My, Hl - oy))
        ww = max(0, min(Mx, Wl - ox))
        src[:, :h, :ww] = pack[ph, :, oy:oy + h, ox:ox + ww]
        out += src @ w.T
    return out


def wgrad_lattice(gpack, g_phase, xpack, taps):
    """spgan_conv_wgrad_gemm: dW[t][o][c] = sum_q G'[g_phase][q][o] * X'[ph_t][q + off_t][c] over the flattened lattice
    (rows beyond the X' matrix are zero)."""
    _, B, Hl, Wl, O = gpack.shape
    P, _, _, _, C = xpack.shape
    n = B * Hl * Wl
    G = gpack[g_phase].reshape(n, O)
    rows = xpack.reshape(P * n, C)
    out = []
    for ph, oy, ox in taps:
        off = ph * n + oy * Wl + ox
        src = np.zeros((n, C))
        hi = min(n, rows.shape[0] - off)
        if hi > 0:
            src[:hi] = rows[off:off + hi]
        out.append(G.T @ src)
    return out
