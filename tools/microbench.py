"""Isolated-kernel microbenchmarks (BASELINE.json configs[4]): bias-act, upfirdn2d, the spherical gather, the fused FIR tail,
the operand packers and the fused spherical modulated conv, each timed with CUDA events on inputs larger than L2 and
reported against its roofline (HBM: MEASURED_PEAKS.json hbm_gbs; tensor: bf16_tflops_sustained).

    python tools/microbench.py [--budget-s 45]

Prints one JSON object per case.  Every case is wrapped so that a failure is reported and the sweep continues.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from bench import ClockSampler
import spgan_b200.functional as SF
import spgan_b200.lib as lib
from spgan_b200 import grids, panorama


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget-s", type=float, default=45.0)
    ap.add_argument("--sweep", action="store_true", help="run the configs[4] sweep (C x H x B, forward and backward)")
    ap.add_argument("--max-gb", type=float, default=6.0, help="skip sweep cases whose tensors exceed this many GiB")
    ap.add_argument("--only", default="", help="comma-separated substrings: run only the fixed cases whose name contains one")
    ap.add_argument("--legacy-ab", action="store_true",
                    help="also time the pre-streaming FIR / gather kernels (SPGAN_LEGACY_HBM_KERNELS=1) beside the defaults")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    lib.require_device()
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6549.8))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1388.5))
    t_start = time.time()

    last_clocks = {}

    def timed(fn, iters=5, warm=2, min_s=0.45):
        """ms per call (CUDA events).  The call is captured in a CUDA graph of `reps` repetitions and the graph is replayed, so
        that small kernels are timed on the device and not on the Python / autograd launch overhead (~0.1 ms per call, which
        made every sub-0.1 ms case of the first sweep read as a fraction of its real bandwidth); eager loop if the capture
        fails.  The loop runs for at least `min_s` so that the nvidia-smi clock sampler (200 ms period) sees the kernel under
        load — the record lands in the case's JSON line."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        one = max(e0.elapsed_time(e1), 1e-3)
        reps = max(1, min(20, int(2.0 / one)))
        graph = None
        if reps > 1:
            try:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.graph(g, stream=side):
                    for _ in range(reps):
                        fn()
                torch.cuda.current_stream().wait_stream(side)
                g.replay()
                torch.cuda.synchronize()
                graph = g
            except Exception:
                torch.cuda.synchronize()
                graph = None
        if graph is None:
            reps = 1
        run = graph.replay if graph is not None else fn
        iters = max(iters, min(4000, int(min_s * 1e3 / (one * reps)) + 1))
        sampler = ClockSampler(0)
        sampler.start()
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        last_clocks.clear()
        last_clocks.update(sampler.stop())
        last_clocks["graph_replay"] = graph is not None
        del graph
        return e0.elapsed_time(e1) / (iters * reps)

    def report(name, ms, bytes_=None, flops=None, **kw):
        out = {"case": name, "ms": round(ms, 4)}
        if bytes_ is not None:
            gbs = bytes_ / (ms / 1e3) / 1e9
            out.update(bound="hbm", algorithmic_MB=round(bytes_ / 1e6, 1), achieved_GBs=round(gbs, 1), peak_GBs=hbm,
                       frac=round(gbs / hbm, 3))
        if flops is not None:
            tfs = flops / (ms / 1e3) / 1e12
            out.update(bound="tensor", algorithmic_GFLOP=round(flops / 1e9, 2), achieved_TFs=round(tfs, 1), peak_TFs=tf_peak,
                       frac=round(tfs / tf_peak, 3))
        out.update(kw)
        out["clocks"] = dict(last_clocks)
        print(json.dumps(out), flush=True)

    only = [t for t in args.only.split(",") if t]

    def legacy_ab(name, fn, **kw):
        """Time `fn` again with the pre-streaming kernels selected (the library reads the variable per call)."""
        if not args.legacy_ab:
            return
        os.environ["SPGAN_LEGACY_HBM_KERNELS"] = "1"
        try:
            report(name + " [legacy kernel]", timed(fn), **kw)
        finally:
            os.environ.pop("SPGAN_LEGACY_HBM_KERNELS", None)

    def case(name, build):
        if only and not any(t in name for t in only):
            return
        if time.time() - t_start > args.budget_s:
            print(json.dumps({"case": name, "skipped": "time budget"}), flush=True)
            return
        try:
            with torch.no_grad():
                build()
        except Exception as e:  # keep sweeping
            if str(e).startswith("skipped:"):
                print(json.dumps({"case": name, "skipped": str(e)[9:]}), flush=True)
            else:
                print(json.dumps({"case": name, "error": "%s: %s" % (type(e).__name__, str(e)[:200])}), flush=True)
        torch.cuda.empty_cache()

    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    SQ2 = 2 ** 0.5
    k3 = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], device=dev) / 16 * 4
    k4 = torch.tensor([1., 3., 3., 1.], device=dev)
    k4 = (k4[None, :] * k4[:, None]) / 64

    pl = panorama.plan(384, 768)
    cp, _ = panorama.patch_inputs(pl, 2, 7, 27, pl["lat_h"], pl["lat_w"])

    # ---- K1 bias + activation (models/custom_ops/fused_bias_act_kernel.cu) ----
    def bias_act_fwd():
        x, b = rn(32, 512, 101, 101), rn(512)
        ms = timed(lambda: SF.bias_act(x, b, None, 3, 0, 0.2, SQ2))
        report("bias_act fwd (32,512,101,101)", ms, bytes_=2 * 4 * x.numel())
    case("bias_act fwd", bias_act_fwd)

    def bias_act_bwd():
        go, out = rn(32, 512, 101, 101), rn(32, 512, 101, 101)
        ms = timed(lambda: SF.FusedLeakyReLUFunctionBackward.apply(go, out, 0.2, SQ2))
        report("bias_act bwd + bias reduction (32,512,101,101)", ms, bytes_=3 * 4 * go.numel())
    case("bias_act bwd", bias_act_bwd)

    def noise_bias_act():
        x, nz, nw, b = rn(32, 512, 101, 101), rn(32, 1, 101, 101), torch.tensor([0.3], device=dev), rn(512)
        ms = timed(lambda: SF.noise_bias_act(x, nz, nw, b))
        report("noise + bias + act (32,512,101,101)", ms, bytes_=2 * 4 * x.numel())
    case("noise_bias_act", noise_bias_act)

    # ---- K2 upfirdn2d (models/custom_ops/upfirdn2d_kernel.cu) ----
    def fir_g():
        x = rn(32, 512, 105, 105)
        ms = timed(lambda: SF.upfirdn2d(x, k3, pad=(0, 0)))
        report("upfirdn2d 3x3 pad 0 (G blur) (32,512,105,105)", ms, bytes_=4 * 32 * 512 * (105 * 105 + 103 * 103))
        legacy_ab("upfirdn2d 3x3 pad 0 (G blur) (32,512,105,105)", lambda: SF.upfirdn2d(x, k3, pad=(0, 0)),
                  bytes_=4 * 32 * 512 * (105 * 105 + 103 * 103))
        x = rn(8, 512, 105, 105)
        ms = timed(lambda: SF.upfirdn2d(x, k3, pad=(0, 0)))
        report("upfirdn2d 3x3 pad 0 (G blur, training batch) (8,512,105,105)", ms, bytes_=4 * 8 * 512 * (105 * 105 + 103 * 103))
    case("upfirdn2d G blur", fir_g)

    def fir_d():
        x = rn(32, 256, 101, 101)
        ms = timed(lambda: SF.upfirdn2d(x, k4, pad=(2, 2)))
        report("upfirdn2d 4x4 pad 2 (D blur) (32,256,101,101)", ms, bytes_=4 * 32 * 256 * (101 * 101 + 102 * 102))
        legacy_ab("upfirdn2d 4x4 pad 2 (D blur) (32,256,101,101)", lambda: SF.upfirdn2d(x, k4, pad=(2, 2)),
                  bytes_=4 * 32 * 256 * (101 * 101 + 102 * 102))
        x = rn(32, 512, 50, 50)
        ms = timed(lambda: SF.upfirdn2d(x, k4, pad=(2, 2)))
        report("upfirdn2d 4x4 pad 2 (D blur) (32,512,50,50)", ms, bytes_=4 * 32 * 512 * (50 * 50 + 51 * 51))
    case("upfirdn2d D blur", fir_d)

    def fir_up():
        x = rn(32, 64, 53, 53)
        ms = timed(lambda: SF.upfirdn2d(x, k3, up=2, down=1, pad=(1, 0)))
        oh = 53 * 2 + 1 - 3 + 1
        report("upfirdn2d up=2 3x3 (32,64,53,53)", ms, bytes_=4 * 32 * 64 * (53 * 53 + oh * oh))
        legacy_ab("upfirdn2d up=2 3x3 (32,64,53,53)", lambda: SF.upfirdn2d(x, k3, up=2, down=1, pad=(1, 0)),
                  bytes_=4 * 32 * 64 * (53 * 53 + oh * oh))
        x = rn(32, 256, 53, 53)
        ms = timed(lambda: SF.upfirdn2d(x, k4 * 4, up=2, down=1, pad=(2, 1)))
        report("upfirdn2d up=2 4x4 pad (2,1) (Upsample) (32,256,53,53)", ms, bytes_=4 * 32 * 256 * (53 * 53 + 106 * 106))
        x = rn(32, 256, 101, 101)
        ms = timed(lambda: SF.upfirdn2d(x, k4, up=1, down=2, pad=(1, 1)))
        report("upfirdn2d down=2 4x4 pad (1,1) (Downsample) (32,256,101,101)", ms, bytes_=4 * 32 * 256 * (101 * 101 + 50 * 50))
        legacy_ab("upfirdn2d down=2 4x4 pad (1,1) (Downsample) (32,256,101,101)",
                  lambda: SF.upfirdn2d(x, k4, up=1, down=2, pad=(1, 1)), bytes_=4 * 32 * 256 * (101 * 101 + 50 * 50))
    case("upfirdn2d up2", fir_up)

    def upblur():
        pp, nz, nw, b = rn(32, 512, 4, 53, 53), rn(32, 1, 103, 103), torch.tensor([0.3], device=dev), rn(512)
        ms = timed(lambda: SF.upblur_act(pp, k3, (105, 105), nz, nw, b))
        report("upblur_act (interleave + FIR + noise + bias + act) -> (32,512,103,103)", ms,
               bytes_=4 * 32 * 512 * (4 * 53 * 53 + 103 * 103))
    case("upblur_act", upblur)

    def upblur_pack():
        pp, nz, nw, b = rn(64, 4, 53, 53, 512), rn(64, 1, 103, 103), torch.tensor([0.3], device=dev), rn(512)
        mul = rn(64, 512)
        ms = timed(lambda: SF.chain_upblur_pack(pp, (105, 105), k3, nz, nw, b, mul, 1))
        report("upblur_pack (channels-last planes -> FIR + noise + bias + act -> packed bf16 hi/lo) -> (64,103,103,512)", ms,
               bytes_=4 * 64 * 512 * (105 * 105 + 103 * 103))
    case("upblur_pack", upblur_pack)

    def sphere_pack_seg():
        B, Bg, C, H = 64, 32, 256, 35
        x, c, s = rn(B, C, H, H), rn(B, 3, H, H), rn(B, C + 3)
        w = rn(256, C + 3, 3, 3)
        d = rn(B, 256).abs() + 0.5
        cp2, _ = panorama.patch_inputs(pl, 2, 8, 28, pl["lat_h"], pl["lat_w"])
        grid = grids.GRID_CACHE.group_grid(H, H, [cp, cp2], dev)
        xh, _ = SF.ss_input(x, 1)
        y_sc, nm = torch.zeros(B, H, H, 256, device=dev), torch.ones(B, 256, device=dev)
        SF.profile_calls(True)
        for _ in range(12):
            SF.ss_sphere(xh, c, grid, Bg, w, s, d, 0.02, (0.01, 1.0), y_sc, nm, 1)
        t = SF.profile_calls(False)
        ms = t["spgan_sphere_pack_seg"][0] / t["spgan_sphere_pack_seg"][1]
        report("sphere_pack_seg (concat repack + vectorised gather producer) (64,256+3,35,35) -> [9*256 | 64] bf16 hi/lo", ms,
               bytes_=4.0 * B * H * H * (9 * 256 + 64) + 4.0 * B * H * H * 259)
        ms = t["spgan_conv_gemm_ex"][0] / t["spgan_conv_gemm_ex"][1]
        report("spherical GEMM K = 9*256 + 64 (64,35,35) -> 256", ms, flops=2.0 * B * H * H * 256 * 259 * 9)
    case("sphere_pack_seg", sphere_pack_seg)

    # ---- L1 spherical gather (F.grid_sample in the reference) ----
    def gather():
        z = rn(32, 256, 35, 35)
        grid = torch.from_numpy(grids.sampling_grid(35, 35, cp)).to(dev)
        ms = timed(lambda: SF.sphere_gather_raw(z, grid))
        report("sphere_gather (32,256,35,35) -> 9x", ms, bytes_=4 * z.numel() * 10 + grid.numel() * 4)
        legacy_ab("sphere_gather (32,256,35,35) -> 9x", lambda: SF.sphere_gather_raw(z, grid),
                  bytes_=4 * z.numel() * 10 + grid.numel() * 4)
        z = rn(8, 256, 35, 35)
        gb = grid.expand(8, -1, -1, -1).contiguous()
        ms = timed(lambda: SF.sphere_gather_raw(z, gb))
        report("sphere_gather, per-sample grids (training) (8,256,35,35) -> 9x", ms, bytes_=4 * z.numel() * 10 + gb.numel() * 4)
    case("sphere_gather", gather)

    # ---- operand packer ----
    def pack():
        x, s = rn(32, 512, 103, 103), rn(32, 512)
        out = torch.empty((2, 32 * 103 * 103, 512), device=dev, dtype=torch.bfloat16)
        import ctypes
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # the capture stream while a graph is recorded
        ms = timed(lambda: lib.call("spgan_pack_act", p(out), p(x), p(s), 32, 512, 103, 103, 512, 0, 0, 103, 103, 1, 0, st()))
        report("pack_act (32,512,103,103) -> bf16 hi/lo channels-last", ms, bytes_=4 * x.numel() + out.numel() * 2)
    case("pack_act", pack)

    # ---- BASELINE configs[4] / SURVEY §8(d)5: C in {64..512} x H in {32..384} x B in {1, 8, 32}, forward and backward, for the
    # spherical modulated conv, upfirdn2d (G blur 3x3 / D blur 4x4 / up 2 / down 2) and bias-act.  Cases whose tensors exceed
    # `--max-gb` are reported as skipped (the autograd path of the spherical conv holds the 9x gathered tensor, as the
    # reference does); the time budget cuts the tail of the sweep, largest shapes first within each (C, B).
    max_bytes = args.max_gb * 2 ** 30

    def sph_case(B, C, H):
        def run():
            if 4.0 * B * C * H * H * 9 * 3 > max_bytes:
                raise RuntimeError("skipped: tensors exceed --max-gb")
            x, c = rn(B, C, H, H), rn(B, 3, H, H)
            w = rn(C, C + 3, 3, 3)
            s, d = rn(B, C + 3), rn(B, C).abs() + 0.5
            grid = torch.from_numpy(grids.sampling_grid(H, H, dict(cp))).to(dev)
            fl = 2.0 * B * H * H * C * (C + 3) * 9
            ms = timed(lambda: SF.sphere_modconv_fused(x, c, grid, w, s, d, 0.05, act=(0.01, 1.0), precision=1), iters=3, warm=1)
            report("sphere_modconv fwd B=%d C=%d H=%d" % (B, C, H), ms, flops=fl)
            with torch.enable_grad():
                xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)

                def fb():
                    y = SF.sphere_modconv(xg, c, grid, wg, s, d, 0.05)
                    gx, gw = torch.autograd.grad(y.sum(), [xg, wg])
                    return gx
                ms = timed(fb, iters=2, warm=1, min_s=0.3)
            report("sphere_modconv fwd+bwd (dX, dW) B=%d C=%d H=%d" % (B, C, H), ms, flops=3 * fl)
        case("sphere_modconv B=%d C=%d H=%d" % (B, C, H), run)

    def fir_case(B, C, H, name, kern, up, down, pad):
        def run():
            if 4.0 * B * C * H * H * (up * up + 1) * 2 > max_bytes:
                raise RuntimeError("skipped: tensors exceed --max-gb")
            x = rn(B, C, H, H)
            y = SF.upfirdn2d(x, kern, up=up, down=down, pad=pad)
            by = 4.0 * (x.numel() + y.numel())
            ms = timed(lambda: SF.upfirdn2d(x, kern, up=up, down=down, pad=pad))
            report("upfirdn2d %s fwd B=%d C=%d H=%d" % (name, B, C, H), ms, bytes_=by)
            with torch.enable_grad():
                xg = x.clone().requires_grad_(True)
                yy = SF.upfirdn2d(xg, kern, up=up, down=down, pad=pad)
                go = torch.ones_like(yy)
                ms = timed(lambda: torch.autograd.grad(yy, xg, go, retain_graph=True))
            report("upfirdn2d %s bwd B=%d C=%d H=%d" % (name, B, C, H), ms, bytes_=by)
        case("upfirdn2d %s B=%d C=%d H=%d" % (name, B, C, H), run)

    def act_case(B, C, H):
        def run():
            if 4.0 * B * C * H * H * 3 > max_bytes:
                raise RuntimeError("skipped: tensors exceed --max-gb")
            x, b = rn(B, C, H, H), rn(C)
            ms = timed(lambda: SF.bias_act(x, b, None, 3, 0, 0.2, SQ2))
            report("bias_act fwd B=%d C=%d H=%d" % (B, C, H), ms, bytes_=2 * 4.0 * x.numel())
            go = rn(B, C, H, H)
            ms = timed(lambda: SF.FusedLeakyReLUFunctionBackward.apply(go, x, 0.2, SQ2))
            report("bias_act bwd + bias reduction B=%d C=%d H=%d" % (B, C, H), ms, bytes_=3 * 4.0 * x.numel())
        case("bias_act B=%d C=%d H=%d" % (B, C, H), run)

    for B in (32, 8, 1):
        for C in (512, 256, 128, 64):
            for H in (384, 256, 128, 64, 32):
                if not args.sweep:
                    continue
                act_case(B, C, H)
                fir_case(B, C, H, "3x3 pad0", k3, 1, 1, (0, 0))
                fir_case(B, C, H, "4x4 pad2", k4, 1, 1, (2, 2))
                fir_case(B, C, H, "up2 4x4", k4 * 4, 2, 1, (2, 1))
                fir_case(B, C, H, "down2 4x4", k4, 1, 2, (1, 1))
                sph_case(B, C, H)

    # ---- plain modulated 3x3 conv fwd / data gradient / weight gradient at the largest layer ----
    def conv_big():
        geom = SF.ConvGeom(3, 3)
        x, w = rn(32, 512, 103, 103), rn(512, 512, 3, 3)
        s, d = rn(32, 512), rn(32, 512).abs() + 0.5
        fl = 2.0 * 32 * 101 * 101 * 512 * 512 * 9
        ms = timed(lambda: SF.conv_apply(x, w, geom, in_mul=s, out_mul=d, out_scale=0.02, precision=1), iters=3, warm=1)
        report("modconv 3x3 512->512 fwd (pack + GEMM) (32,512,103,103)", ms, flops=fl)
        gy = rn(32, 512, 101, 101)
        ms = timed(lambda: SF.conv_apply(gy, w, geom, True, (103, 103), d, s, 0.02, precision=1), iters=3, warm=1)
        report("modconv 3x3 512->512 data gradient (pack + GEMM)", ms, flops=fl)
        ms = timed(lambda: SF.conv_wgrad(gy[:8], x[:8], (512, 512, 3, 3), geom, s[:8], d[:8], 0.02, precision=1), iters=3, warm=1)
        report("modconv 3x3 512->512 weight gradient B=8 (2 packs + GEMM + reduce)", ms, flops=fl / 4)
    case("modconv big", conv_big)


if __name__ == "__main__":
    main()
