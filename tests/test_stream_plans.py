"""CPU: the host-side launch plans of the streamed FIR / gather kernels (csrc/upfirdn2d.cu plan_fir_stream,
csrc/sphere_gather.cu plan_gather_stream), swept through the host-only C-ABI entry points for their invariants: every output
row and column is covered exactly once, nothing staged exceeds the shared-memory stage, thread blocks stay within the
register-limited size, interior columns really have all taps inside the image."""
import ctypes
import itertools

import numpy as np
import pytest


def _lib():
    import __graft_entry__ as ge
    ge.build()
    import spgan_b200.lib as lib
    return lib.load()


def fir_plan(planes, h, w, k, up, down, pad):
    out = (ctypes.c_int32 * 20)()
    rc = _lib().spgan_upfirdn2d_plan(planes, h, w, k, k, up, down, pad[0], pad[1], pad[0], pad[1], ctypes.cast(out, ctypes.c_void_p))
    assert rc == 0
    keys = ["variant", "P", "bands", "R", "strips", "G", "threads", "stage_floats", "n_int", "int_lo", "n_bord", "Gb",
            "border_base", "items", "grid", "out_h", "out_w", "max_stage", "max_threads", "strip"]
    return dict(zip(keys, list(out)))


def _fir_cases():
    geoms = [(3, 1, 1, (0, 0)), (3, 1, 1, (2, 2)), (4, 1, 1, (2, 2)), (4, 1, 1, (1, 1)), (4, 1, 1, (2, 1)), (2, 1, 1, (1, 0)),
             (4, 2, 1, (2, 1)), (3, 2, 1, (1, 0)), (4, 1, 2, (1, 1)), (4, 1, 2, (1, 2))]
    sizes = [(5, 7), (11, 11), (19, 19), (50, 50), (53, 53), (101, 101), (105, 105), (106, 106), (128, 128), (150, 131),
             (30, 700), (40, 400), (256, 256), (384, 384), (9, 2900)]
    planes = [1, 3, 600, 16384]
    return list(itertools.product(geoms, sizes, planes))


@pytest.mark.parametrize("chunk", range(4))
def test_fir_stream_plan_invariants(chunk):
    cases = _fir_cases()[chunk::4]
    streamed = 0
    for (k, up, down, pad), (h, w), planes in cases:
        p = fir_plan(planes, h, w, k, up, down, pad)
        oh = (h * up + pad[0] + pad[1] - k) // down + 1
        ow = (w * up + pad[0] + pad[1] - k) // down + 1
        assert (p["out_h"], p["out_w"]) == (max(oh, 0), max(ow, 0))
        if p["variant"] == 0:
            continue
        streamed += 1
        K = p["variant"] // 100
        assert p["variant"] % 100 == up * 10 + down and K >= k
        S = p["strip"]
        # stage capacity
        assert p["stage_floats"] <= p["max_stage"]
        if p["bands"] == 1 and p["P"] >= 1 and h * w + 8 <= p["max_stage"]:
            assert p["P"] * h * w + 8 <= p["stage_floats"]
            assert p["R"] == oh
            assert p["items"] == -(-planes // p["P"])
        else:
            assert p["P"] == 1
            # every band's input rows fit the stage (restating the kernel's decode)
            for band in range(p["bands"]):
                oy0 = band * p["R"]
                rows = min(p["R"], oh - oy0)
                assert rows >= 1
                y_first, y_last = oy0 * down - pad[0], (oy0 + rows - 1) * down - pad[0] + K - 1
                lo = max((y_first + 1) >> 1 if up == 2 else y_first, 0)
                hi = min(y_last >> 1 if up == 2 else y_last, h - 1)
                assert (hi - lo + 1) * w + 8 <= p["stage_floats"], (h, w, k, up, down, pad, band)
            assert p["items"] == min(planes * p["bands"], 2 ** 31 - 1)
        # rows covered exactly: bands * R >= out_h, no empty band, whole strips
        assert p["bands"] * p["R"] >= oh and (p["bands"] - 1) * p["R"] < oh
        assert p["strips"] == -(-p["R"] // S)
        if p["bands"] > 1:
            assert p["R"] % S == 0  # every band starts on the same tap parity (up 2)
        # threads
        assert p["threads"] % 32 == 0 and 32 <= p["threads"] <= p["max_threads"]
        assert 1 <= p["grid"] <= max(1, p["items"])
        if p["n_int"] > 0:
            assert up == 1 and down == 1
            n_int, lo = p["n_int"], p["int_lo"]
            assert n_int + p["n_bord"] == ow and lo >= 0
            # interior columns have all K taps inside the image; the columns next to them do not
            for ox in (lo, lo + n_int - 1):
                assert 0 <= ox - pad[0] and ox - pad[0] + K - 1 <= w - 1
            for ox in (lo - 1, lo + n_int):
                if 0 <= ox < ow:
                    assert ox - pad[0] < 0 or ox - pad[0] + K - 1 > w - 1
            assert p["G"] * n_int <= p["border_base"] <= p["threads"]
            assert p["border_base"] % 32 == 0
            assert p["border_base"] + p["Gb"] * p["n_bord"] <= p["threads"]
            assert p["Gb"] >= 1
        elif ow <= p["max_threads"]:
            assert p["G"] * ow <= p["threads"]
        else:
            assert p["G"] == 1
            passes = -(-ow // p["max_threads"])
            assert p["threads"] * passes >= ow
    assert streamed > len(cases) // 2


def test_fir_stream_plan_hot_path_shapes():
    """The generator's and discriminator's blurs take the streamed kernel with whole planes and the split column mapping."""
    g = fir_plan(32 * 512, 105, 105, 3, 1, 1, (0, 0))
    assert g["variant"] == 311 and g["bands"] == 1 and g["P"] == 1 and g["n_bord"] == 0 and g["n_int"] == 103
    d = fir_plan(32 * 256, 101, 101, 4, 1, 1, (2, 2))
    assert d["variant"] == 411 and d["bands"] == 1 and d["n_int"] == 98 and d["n_bord"] == 4 and d["int_lo"] == 2
    u = fir_plan(32 * 256, 53, 53, 4, 2, 1, (2, 1))
    assert u["variant"] == 421 and u["P"] == 4  # four 53 x 53 planes per 46 KB stage
    assert fir_plan(32 * 3, 53, 53, 4, 2, 1, (2, 1))["P"] == 1  # few planes: spread over the CTAs first
    assert fir_plan(8, 64, 64, 5, 1, 1, (2, 2))["variant"] == 0      # 5 x 5: generic kernel
    assert fir_plan(8, 64, 64, 4, 2, 2, (2, 2))["variant"] == 0      # up and down together: polyphase kernel


def gather_plan(B, C, H, W, encode=0):
    out = (ctypes.c_int32 * 12)()
    rc = _lib().spgan_sphere_gather_plan(B, C, H, W, encode, ctypes.cast(out, ctypes.c_void_p))
    assert rc == 0
    keys = ["streamed", "cc", "chunks", "psplit", "pslice", "raw_floats", "smem", "resident", "grid", "items", "smem_max",
            "threads"]
    return dict(zip(keys, list(out)))


def test_gather_stream_plan_invariants():
    n = 0
    for B, C, (H, W), enc in itertools.product([1, 2, 8, 32, 64], [1, 3, 5, 37, 256, 259], [(7, 9), (17, 17), (35, 35), (53, 53),
                                                                                           (83, 83), (130, 70), (384, 384)], [0, 1]):
        if enc and C != 3:
            continue
        p = gather_plan(B, C, H, W, enc)
        if not p["streamed"]:
            assert H * W * 32 > p["smem_max"] - 64  # only planes too large for one 4-channel stage fall back
            continue
        n += 1
        cc = p["cc"]
        assert cc in (4, 8) and (cc == 4 or (C >= 8 and not enc))
        assert p["chunks"] == -(-C // cc)
        assert p["raw_floats"] >= cc * H * W + 8 and p["raw_floats"] % 32 == 0
        assert p["smem"] == 4 * (p["raw_floats"] + H * W * (12 if cc == 8 else 4)) <= p["smem_max"]
        assert p["resident"] in (1, 2) and p["resident"] * (p["smem"] + 1024) <= 228 * 1024
        assert 1 <= p["psplit"] <= 4 and p["psplit"] * p["pslice"] >= 9 * H * W > (p["psplit"] - 1) * p["pslice"]
        assert p["items"] == B * p["chunks"] * p["psplit"]
        assert 1 <= p["grid"] <= p["items"]
    assert n > 100
    hot = gather_plan(32, 256, 35, 35)
    assert hot["cc"] == 8 and hot["resident"] == 2 and hot["psplit"] == 2


# ------------------------------------------------------------------------------------------------ index arithmetic
def _emulate_fir_stream(p, x, kern, up, down, pad):
    """numpy restatement of fir_stream_kernel's index arithmetic (csrc/upfirdn2d.cu: item decode, staged row window, strips,
    zero-word taps, up-2 tap parity) driven by the exported plan.  Every output must be written exactly once."""
    planes, h, w = x.shape
    kh, kw = kern.shape
    K, S = p["variant"] // 100, p["strip"]
    px0 = py0 = pad[0]
    oh, ow = p["out_h"], p["out_w"]
    kf = np.zeros((K, K), x.dtype)
    kf[:kh, :kw] = kern[::-1, ::-1]
    out = np.full((planes, oh, ow), np.nan, x.dtype)
    written = np.zeros((planes, oh, ow), np.int32)
    ox = np.arange(ow)
    for item in range(p["items"]):
        if p["bands"] == 1:
            plane0 = item * p["P"]
            npl = min(p["P"], planes - plane0)
            oy0, rows_out, iy_lo, iy_hi = 0, oh, 0, h - 1
        else:
            plane0, band = divmod(item, p["bands"])
            npl = 1
            oy0 = band * p["R"]
            rows_out = min(p["R"], oh - oy0)
            y_first, y_last = oy0 * down - py0, (oy0 + rows_out - 1) * down - py0 + K - 1
            iy_lo = max((y_first + 1) >> 1 if up == 2 else y_first, 0)
            iy_hi = min(y_last >> 1 if up == 2 else y_last, h - 1)
        nrows = iy_hi - iy_lo + 1
        assert npl * h * w + 8 <= p["stage_floats"] if p["bands"] == 1 else nrows * w + 8 <= p["stage_floats"]
        for pl in range(npl):
            staged = x[plane0 + pl, iy_lo:iy_hi + 1]  # what the bulk copy brings in
            for strip in range(p["strips"]):
                ly0 = strip * S
                rows_left = rows_out - ly0
                if rows_left <= 0:
                    continue
                for i in range(min(S, rows_left)):
                    acc = np.zeros(ow, x.dtype)
                    if up == 1:
                        row0 = (oy0 + ly0) * down - py0 - iy_lo
                        for ky in range(K):
                            ry = row0 + i * down + ky
                            if not 0 <= ry < nrows:
                                continue  # CHECK variant: clamped row times zero
                            for kx in range(K):
                                ix = ox * down - px0 + kx
                                ok = (ix >= 0) & (ix < w)  # else: the zero word
                                acc += kf[ky, kx] * np.where(ok, staged[ry, np.clip(ix, 0, w - 1)], 0)
                    else:
                        c = ox - px0
                        kx0 = c & 1
                        ix0 = (c + kx0) >> 1
                        par0 = (oy0 - py0) & 1
                        t0 = oy0 + ly0 - py0
                        j0 = ((t0 + par0) >> 1) - iy_lo
                        A = (i >> 1) if par0 else ((i + 1) >> 1)
                        par = (par0 + i) & 1
                        for a, ky in ((0, par), (1, par + 2)):
                            ry = j0 + A + a
                            if not 0 <= ry < nrows:
                                continue
                            for b in range(2):
                                ix = ix0 + b
                                ok = (ix >= 0) & (ix < w)
                                wcol = np.where(kx0 == 1, kf[ky, 1 + 2 * b] if 1 + 2 * b < K else 0, kf[ky, 2 * b])
                                acc += wcol * np.where(ok, staged[ry, np.clip(ix, 0, w - 1)], 0)
                    out[plane0 + pl, oy0 + ly0 + i] = acc
                    written[plane0 + pl, oy0 + ly0 + i] += 1
    assert (written == 1).all()
    return out


def test_fir_stream_index_arithmetic_vs_oracle():
    """The kernel's decode / window / parity formulas, restated in numpy and driven by the real plan, reproduce the oracle's
    upfirdn2d over geometries far beyond the GPU test list (all pads 0..K-1, non-square small kernels for up / down 2, bands)."""
    import spgan_oracle as O
    rng = np.random.default_rng(7)
    cases = []
    for k in (2, 3, 4):
        for pad in itertools.product(range(k), range(k)):
            cases.append((3, 13, 11, (k, k), 1, 1, pad))
    for kh, kw in ((4, 4), (3, 3), (2, 4), (4, 1), (1, 3)):
        for pad in ((0, 0), (1, 0), (2, 1), (3, 3), (1, 2), (0, 3)):
            cases.append((2, 9, 12, (kh, kw), 2, 1, pad))
            cases.append((2, 14, 11, (kh, kw), 1, 2, pad))
    # bands: planes larger than a 46 KB stage
    cases += [(1, 120, 131, (4, 4), 1, 1, (2, 1)), (1, 150, 100, (3, 3), 1, 1, (0, 0)), (1, 140, 110, (4, 4), 2, 1, (2, 1)),
              (1, 141, 97, (3, 3), 2, 1, (1, 0)), (1, 260, 90, (4, 4), 1, 2, (1, 1)), (1, 301, 60, (4, 4), 1, 2, (3, 0)),
              (2, 40, 400, (4, 4), 1, 1, (2, 2)), (5, 50, 50, (4, 4), 1, 1, (2, 2))]
    streamed = 0
    for planes, h, w, (kh, kw), up, down, pad in cases:
        out = (ctypes.c_int32 * 20)()
        assert _lib().spgan_upfirdn2d_plan(planes, h, w, kh, kw, up, down, pad[0], pad[1], pad[0], pad[1],
                                           ctypes.cast(out, ctypes.c_void_p)) == 0
        keys = ["variant", "P", "bands", "R", "strips", "G", "threads", "stage_floats", "n_int", "int_lo", "n_bord", "Gb",
                "border_base", "items", "grid", "out_h", "out_w", "max_stage", "max_threads", "strip"]
        p = dict(zip(keys, list(out)))
        if p["variant"] == 0 or p["out_h"] <= 0 or p["out_w"] <= 0:
            continue
        streamed += 1
        x = rng.standard_normal((planes, h, w))
        kern = rng.standard_normal((kh, kw))
        want = O.upfirdn2d(x[None], kern, up=(up, up), down=(down, down), pad=(pad[0], pad[1], pad[0], pad[1]))[0]
        got = _emulate_fir_stream(p, x, kern, up, down, pad)
        assert got.shape == want.shape, (h, w, kh, kw, up, down, pad)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12), (h, w, kh, kw, up, down, pad)
    assert streamed >= 60


def _emulate_gather_stream(p, z, grid):
    """numpy restatement of sphere_gather_stream_kernel's item decode, [pixel][channel] tile and position slices
    (csrc/sphere_gather.cu) driven by the exported plan; corners / weights come from the oracle's index arithmetic."""
    import spgan_oracle as O
    B, C, H, W = z.shape
    cc, TS = p["cc"], (12 if p["cc"] == 8 else 4)
    opix = 9 * H * W
    x0, y0, wx1, wy1 = O.gather_indices(grid, H, W)
    x0, y0, wx1, wy1 = (a.reshape(a.shape[0], -1) for a in (x0, y0, wx1, wy1))
    out = np.full((B, C, opix), np.nan, np.float32)
    written = np.zeros((B, C, opix), np.int32)
    per_sample = p["chunks"] * p["psplit"]
    for item in range(p["items"]):
        b, rem = divmod(item, per_sample)
        c0 = (rem // p["psplit"]) * cc
        sl = rem % p["psplit"]
        nc = min(cc, C - c0)
        tile = np.zeros((H * W, TS), np.float32)  # channels beyond the tensor read as zero
        tile[:, :nc] = z[b, c0:c0 + nc].reshape(nc, H * W).T
        lo, hi = sl * p["pslice"], min(opix, (sl + 1) * p["pslice"])
        pos = np.arange(lo, hi)
        g = 0 if grid.shape[0] == 1 else b
        xa, ya = x0[g, pos], y0[g, pos]
        xb, yb = np.minimum(xa + 1, W - 1), np.minimum(ya + 1, H - 1)
        fx, fy = wx1[g, pos], wy1[g, pos]
        ex, ey = np.float32(1) - fx, np.float32(1) - fy
        v = (tile[ya * W + xa, :nc] * (ex * ey)[:, None] + tile[ya * W + xb, :nc] * (fx * ey)[:, None] +
             tile[yb * W + xa, :nc] * (ex * fy)[:, None] + tile[yb * W + xb, :nc] * (fx * fy)[:, None])
        out[b, c0:c0 + nc, lo:hi] = v.T
        written[b, c0:c0 + nc, lo:hi] += 1
    assert (written == 1).all()
    return out.reshape(B, C, 3 * H, 3 * W)


def test_gather_stream_item_decode_vs_oracle():
    import cases as K
    import spgan_oracle as O
    rng = np.random.default_rng(11)
    cp = K.test_cp(2, 7, 27)
    for B, C, H, W, shared in [(2, 5, 7, 9, False), (3, 37, 11, 11, True), (1, 3, 17, 17, True), (40, 8, 5, 6, True),
                               (2, 9, 35, 35, False), (300, 4, 5, 5, True)]:
        p = gather_plan(B, C, H, W)
        assert p["streamed"]
        g1 = O.gen_sampling_grid(H, W, cp)
        grid = g1 if shared else np.concatenate([g1 + np.float32(0.004 * i) for i in range(B)], 0)
        z = rng.standard_normal((B, C, H, W)).astype(np.float32)
        want = O.grid_sample_border(z, np.repeat(g1, B, 0) if shared else grid)
        got = _emulate_gather_stream(p, z, grid)
        assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max(), (B, C, H, W)
