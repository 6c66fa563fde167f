"""bench.py — SP-GAN generator forward: 384x768 close-loop panoramas per second (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32] [--precision 1]

A step = one batch of B = 32 panoramas = 60 lattice positions x 32 patches = 1920 generator patch forwards through
the public API (`spgan_b200.panorama.generate` over `spgan_b200.generator.Generator`), random-init weights of
configs/model/spgan.yaml, synthetic latents/noise.  `value` times the steps with all inputs resident in HBM; `e2e`
repeats them with the inputs in pinned host memory (H2D inside the timed region) and the finished panoramas copied
back (D2H).  Under torchrun every rank generates its own 32 panoramas (weak scaling, no data-path collective).
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "panoramas_per_sec_384x768_generator_forward"
UNIT = "panoramas/s"
PATCH_GFLOP = 101.0  # algorithmic GFLOP per 101x101 patch forward (50.51 GMAC)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", type=int, default=1, choices=[0, 1, 2, 3])
    ap.add_argument("--streams", type=int, default=2, help="concurrent lattice-position branches of the panorama graph")
    ap.add_argument("--group", type=int, default=0, help="lattice positions per generator call (panorama.PanoramaEngine group); "
                                                         "0 = automatic: ~64 patches per call, at most 8 positions")
    ap.add_argument("--pair-mode", type=int, default=1, choices=[0, 1, 2], help="CTA-pair GEMM kernel: 0 never, 1 auto, 2 wherever legal")
    ap.add_argument("--ts-precision", default="", help="8 comma-separated per-layer modes for the texture chain")
    ap.add_argument("--fast-tail", action="store_true", help="also measure the fp16x2 tail policy (reported beside the headline)")
    ap.add_argument("--no-strict", action="store_true", help="do not also measure bf16x3-everywhere (reported beside the headline)")
    ap.add_argument("--no-pano768", action="store_true", help="do not append the 768x1536 lattice-sharded object")
    ap.add_argument("--pano768-batch", type=int, default=8)
    ap.add_argument("--skip-profile", action="store_true", help="skip the eager per-launch profiling step")
    ap.add_argument("--cpu-sample-patches", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs under ncu only: skip the end-to-end leg")
    ap.add_argument("--no-hbm", action="store_true", help="skip the hbm_kernels side table (FIR / gather / bias-act timed alone)")
    ap.add_argument("--profile-calls", action="store_true", help="add per-entry-point CUDA-event times of one extra step")
    ap.add_argument("--workload", default="panorama", choices=["panorama", "train", "pano768"],
                    help="panorama = BASELINE configs[1] (default, the headline); train = configs[2], full G+D step; "
                         "pano768 = configs[3], one batch of 768x1536 panoramas with the patch lattice sharded over the ranks")
    ap.add_argument("--train-batch", type=int, default=8)
    ap.add_argument("--no-train", action="store_true", help="panorama workload: do not append the train workload's object")
    ap.add_argument("--no-graphs", action="store_true", help="train workload: launch every kernel eagerly (no CUDA graphs)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_reference_rate(n_patches, threads=None):
    """panoramas/s of the CPU restatement of the reference (oracle/spgan_oracle.py), B = 1, on the first
    `n_patches` lattice positions of a 384x768 panorama (every position costs the same)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import spgan_oracle as O
    import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    plan = O.close_loop_plan(384, 768)
    with open(os.path.join(ROOT, "tests", "golden", "generator_manifest.json")) as f:
        sd = synth.synthetic_state_dict(json.load(f), 9000)
    gl = synth.randn_t(9000, "bench_gl", (1, 512))
    gl = torch.stack([gl, gl], 1)
    canvas = synth.randn_t(9000, "bench_canvas", (1, 256, plan["lat_h"], plan["lat_w"]))
    noises = [synth.randn_t(9000, "bench_noise%d" % l, (1, 1, plan["noise_h"][l], plan["noise_w"][l])) for l in range(8)]
    pos = [(a, b) for a in range(plan["steps_h"]) for b in range(plan["steps_w"])]
    with torch.no_grad():
        O.generate_panorama(sd, plan, gl, canvas, noises, positions=set(pos[:1]))  # warm-up patch
        t0 = time.perf_counter()
        O.generate_panorama(sd, plan, gl, canvas, noises, positions=set(pos[:n_patches]))
        dt = time.perf_counter() - t0
    total = len(pos)
    return (n_patches / total) / dt, threads, dt, total


def reference_rate(steps, warmup):
    """(panoramas/s, seconds per step list, threads, kind, sample) of the reference's CPU path.  The REAL reference (staged
    under oracle/_ref by oracle/build_ref.py, or /root/reference in the build container) driven by its own close-loop
    manager over all 60 lattice positions of one B = 1 panorama; the oracle port on a 12-position sample only when the
    reference cannot be imported."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import refrun
        if not refrun.available():
            raise RuntimeError("reference tree not staged")
        times, threads, total = refrun.time_reference_panorama(steps=steps, warmup=warmup)
        rate = len(times) / sum(times)
        return rate, times, threads, "reference", ("B=1, one full 384x768 panorama per step (all %d lattice positions), the reference's own "
                                                  "InfinityGanGenerator + close-loop manager on CPU (native PyTorch ops)" % total)
    except Exception as e:  # noqa: BLE001 - any import problem of the staged reference falls back to the port
        sys.stderr.write("reference arm: real reference unavailable (%s: %s); timing the oracle port\n" % (type(e).__name__, str(e)[:200]))
        per_step, times, threads, total = 12, [], None, 60
        for i in range(warmup + steps):
            _, threads, dt, total = cpu_reference_rate(per_step)
            if i >= warmup:
                times.append(dt)
        rate = (per_step / total) * len(times) / sum(times)
        return rate, times, threads, "port", ("B=1, %d of %d patch positions of one 384x768 panorama per step, oracle port of the "
                                              "reference on CPU" % (per_step, total))


def cpu_baseline_object(args):
    rate, times, threads, kind, sample = reference_rate(1, 0)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample + " (%.1f s)" % sum(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, times, threads, kind, sample = reference_rate(args.steps, args.warmup)
    ms = 1000.0 * sum(times) / len(times)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "SP-GAN generator forward at 384x768 (close-loop, 60 patch positions), random-init configs/model/spgan.yaml, "
                               "synthetic latents; reference CPU path at batch 1 (BASELINE configs[0])", "batch_per_gpu": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
DTYPES = {0: "fp32 (SIMT)", 1: "fp32-equivalent (bf16x3 split on tcgen05, fp32 accumulate)",
          2: "bf16 (tcgen05, fp32 accumulate)", 3: "fp16x2 split on tcgen05 (fp16 hi+lo activations x fp16 weights, fp32 accumulate)"}
ISSUED = {0: 0, 1: 3, 2: 1, 3: 2}  # tensor-core MMAs issued per algorithmic MMA


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def measure_panorama(args, dev, world, rank, local, th, tw, B, sharded, steps, warmup, layer_precision=None,
                     profile=True, e2e=True):
    """Times `steps` panorama steps (one step = every lattice position of a batch of B panoramas) through
    spgan_b200.panorama.PanoramaEngine.  Returns a dict of raw measurements (all ranks)."""
    import torch
    import torch.distributed as dist
    import spgan_b200.functional as SF
    import spgan_b200.lib as lib
    from spgan_b200 import panorama
    from spgan_b200.generator import Generator

    torch.manual_seed(9000)
    gen = Generator().to(dev).eval()
    gen.texture_synthesizer.layer_precision = layer_precision
    pl = panorama.plan(th, tw)
    n_pos = len(panorama.positions(pl))
    # sharded: every rank holds the SAME canvases (same seed) and runs its share of the lattice (strong scaling)
    g = torch.Generator(device="cpu").manual_seed(9000 + (0 if sharded else rank))
    host = {
        "gl": torch.randn(B, 2, 512, generator=g).pin_memory(),
        "canvas": torch.randn(B, 256, pl["lat_h"], pl["lat_w"], generator=g).pin_memory(),
        "noises": [torch.randn(B, 1, pl["noise_h"][l], pl["noise_w"][l], generator=g).pin_memory() for l in range(8)],
    }
    host["gl"][:, 1] = host["gl"][:, 0]  # the managers use mixing=False: both columns equal
    host_out = torch.empty(B, 3, pl["meta_h"], pl["meta_w"]).pin_memory()
    h2d = host["gl"].numel() * 4 + host["canvas"].numel() * 4 + sum(n.numel() * 4 for n in host["noises"])
    d2h = host_out.numel() * 4
    graphs = not args.no_graphs
    if sharded:
        eng = panorama.ShardedPanoramaEngine(gen, pl, B, dev, rank, world, streams=args.streams, use_graph=graphs, group=args.group or None)
    else:
        eng = panorama.PanoramaEngine(gen, pl, B, dev, streams=args.streams, use_graph=graphs, group=args.group or None)
    group_used = (eng.engine if sharded else eng).group
    eng.load(host["gl"], host["canvas"], host["noises"])

    def step_resident():
        return eng.run()

    def step_e2e():
        eng.load(host["gl"], host["canvas"], host["noises"])
        host_out.copy_(eng.run(), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # a start/end range (process-wide, unlike push/pop): `ncu --nvtx --nvtx-include "spgan_timed"` lists exactly the
        # launches of the timed region, including those issued by autograd's backward thread
        rng = torch.cuda.nvtx.range_start("spgan_timed")
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.nvtx.range_end(rng)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # warm-up: two eager passes (they also count the launches of a step), the capture, then graph replays
    l0, g0 = lib.launches(), lib.load().spgan_gemm_launch_count()
    step_resident()
    l1, g1 = lib.launches(), lib.load().spgan_gemm_launch_count()
    for _ in range(max(warmup, 3) + 2):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(step_resident, steps)
    clocks = sampler.stop() if sampler else None
    out = {"ms_step": ms_total / steps, "ms_total": ms_total, "launches": l1 - l0, "gemm_launches": int(g1 - g0),
           "clocks": clocks, "h2d": h2d, "d2h": d2h, "n_pos": n_pos, "plan": pl, "graphs": graphs, "group": group_used,
           "canvas_mb": host["canvas"].numel() * 4 // 2 ** 20}
    if e2e:
        step_e2e()
        out["ms_e2e"] = timed(step_e2e, steps) / steps
    if profile:
        # per-launch CUDA-event times need eager launches on one stream (events cannot be recorded inside a graph replay,
        # and concurrent branches would overlap the brackets): one extra eager step of the same work after the timed region
        prof = (panorama.ShardedPanoramaEngine(gen, pl, B, dev, rank, world, streams=1, use_graph=False, group=args.group or None) if sharded
                else panorama.PanoramaEngine(gen, pl, B, dev, streams=1, use_graph=False, group=args.group or None))
        prof.load(host["gl"], host["canvas"], host["noises"])
        prof.run()
        SF.profile_gemm(True)
        ms_prof = timed(prof.run, 1)
        out["gemm_stats"] = SF.profile_gemm(False)
        out["ms_eager_1stream"] = ms_prof
        if args.profile_calls:
            SF.profile_calls(True)
            ms_calls = timed(prof.run, 1)
            cm = {k: [round(v[0], 3), v[1]] for k, v in sorted(SF.profile_calls(False).items(), key=lambda kv: -kv[1][0])}
            cm["_step_total_ms"] = ms_calls
            out["call_ms"] = cm
        del prof
    del eng
    torch.cuda.empty_cache()
    return out


def parse_modes(text):
    if not text:
        return None
    m = [int(v) for v in text.split(",")]
    if len(m) != 8:
        raise SystemExit("--ts-precision needs 8 comma-separated modes")
    return m


def run_ours(args):
    import torch
    import torch.distributed as dist
    import spgan_b200.functional as SF
    import spgan_b200.lib as lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib.require_device()  # no fallback: fail loudly if the extension or a B200 is missing
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    SF.set_precision(args.precision)
    SF.set_gemm_pair_mode(args.pair_mode)

    B = args.batch
    sharded = args.workload == "pano768"
    th, tw = (768, 1536) if sharded else (384, 768)
    modes = parse_modes(args.ts_precision)
    warm = args.warmup if args.skip_e2e else max(args.warmup, 3)
    m = measure_panorama(args, dev, world, rank, local, th, tw, B, sharded, args.steps, warm, layer_precision=modes,
                         profile=not args.skip_profile, e2e=not args.skip_e2e)
    pl, n_pos = m["plan"], m["n_pos"]
    ms_step = m["ms_step"]
    jobs = B if sharded else world * B  # sharded: the ranks share ONE batch of panoramas
    value = jobs / (ms_step / 1000.0)
    ms_e2e = m.get("ms_e2e", float("nan"))
    e2e_value = jobs / (ms_e2e / 1000.0)

    # two more points beside the headline policy: bf16x3 in EVERY conv (strict), and the 2-MMA fp16 split on the last three
    # texture layers (stated looser bound, tests/test_gpu_generator.py)
    def side_point(fm, note):
        f = measure_panorama(args, dev, world, rank, local, th, tw, B, False, args.steps, 3, layer_precision=fm, profile=True,
                             e2e=False)
        gs = f.get("gemm_stats")
        return {"layer_precision": fm, "value": jobs / (f["ms_step"] / 1000.0), "unit": UNIT, "ms_per_step": f["ms_step"],
                "gemm_algorithmic_tflops": gs["flops"] / (gs["ms"] / 1000.0) / 1e12 if gs and gs["ms"] > 0 else None, "note": note}

    fast = strict = None
    side_ok = not sharded and modes is None and args.precision == 1 and not args.skip_e2e
    if side_ok and not args.no_strict:
        strict = side_point([1] * 8, "bf16x3 (3 MMAs per product) in every conv, the last texture conv included; generator output "
                                     "2.3e-4..2.6e-4 from the reference against 3.2e-4..3.8e-4 for the headline policy (both held to 5e-4)")
    if side_ok and args.fast_tail:
        fast = side_point([1, 1, 1, 1, 1, 3, 3, 3], "texture layers 5-7 (73 % of the FLOPs) issue 2 MMAs per product instead of 3; "
                          "generator output within 7e-4 of the reference (max-abs over peak) instead of 5e-4 — reported beside the "
                          "headline, not as it")

    # strong-scaling point of BASELINE configs[3] in the same line: one batch of 768x1536 panoramas, lattice sharded
    pano768 = None
    if not sharded and not args.no_pano768 and not args.skip_e2e:
        p7 = measure_panorama(args, dev, world, rank, local, 768, 1536, args.pano768_batch, True, max(1, args.steps // 2 + 1), 1,
                              layer_precision=modes, profile=False, e2e=False)
        pano768 = {"metric": "panoramas_per_sec_768x1536_generator_forward_lattice_sharded",
                   "value": args.pano768_batch / (p7["ms_step"] / 1000.0), "unit": UNIT, "n_gpus": world,
                   "ms_per_step": p7["ms_step"], "scaling": "strong", "batch": args.pano768_batch,
                   "positions": p7["n_pos"], "collective": "one all-gather of the finished patches per step" if world > 1 else "none (1 rank)",
                   "clocks": p7["clocks"]}

    # second half of BASELINE.json's metric ("... & train img/s"): the train workload, same process, same ranks
    train = None
    if not sharded and not args.no_train and not args.skip_e2e:
        train = measure_train(args, dev, world, rank, local)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    gemm_stats = m.get("gemm_stats")
    roofline = None
    if gemm_stats:
        ach = gemm_stats["flops"] / (gemm_stats["ms"] / 1000.0) / 1e12 if gemm_stats["ms"] > 0 else 0.0
        traffic, traffic_note = ncu_traffic()
        roofline = {"bound": "tensor", "kernel": "conv_gemm2_kernel / conv_gemm_kernel (tcgen05 implicit-GEMM modulated conv: cta_group::2 CTA-pair and single-CTA variants)", "achieved": ach,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None, "traffic": traffic,
                    "traffic_note": traffic_note, "peak_source": peak_src, "launches_timed": gemm_stats["launches"],
                    "avg_launch_ms": gemm_stats["ms"] / max(gemm_stats["launches"], 1),
                    "share_of_step": gemm_stats["ms"] / (m["ms_eager_1stream"] if m.get("ms_eager_1stream") else 1.0),
                    "note": "achieved = algorithmic conv FLOPs (2*B*Ho*Wo*Cout*Cin*k^2, valid outputs, real channels) / CUDA-event "
                            "time of the launches, taken from ONE eager single-stream step run right after the timed region (the timed "
                            "steps are CUDA-graph replays with concurrent branches, which cannot be bracketed per launch); mode %d "
                            "issues %dx that many tensor-core FLOPs (2x in the last texture conv under the default policy)"
                            % (args.precision, ISSUED[args.precision])}
        shapes = sorted(gemm_stats.get("shapes", {}).items(), key=lambda kv: -kv[1][0])
        roofline["by_shape"] = [{"shape": k, "ms_per_step": round(v[0], 3), "launches_per_step": v[2],
                                 "algorithmic_tflops": round(v[1] / (v[0] / 1000.0) / 1e12, 1) if v[0] > 0 else None}
                                for k, v in shapes[:24]]
    per_gpu_patches = B * (-(-n_pos // world) if sharded else n_pos)
    out = {
        "metric": METRIC.replace("384x768", "%dx%d" % (th, tw)) + ("_lattice_sharded" if sharded else ""), "value": value,
        "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": ("fp32-equivalent: bf16x3 split on tcgen05 (3 MMAs per product, fp32 accumulate) in every conv except the last "
                  "texture conv, which runs the fp16 hi/lo x fp16 2-MMA split; generator output within 5e-4 of the fp32 reference"
                  if modes is None and args.precision == 1 else DTYPES[args.precision] if modes is None else "per-layer modes %s" % modes),
        "data": "synthetic",
        "config": {"workload": "SP-GAN generator forward batch %d at %dx%d (close-loop, %d patch positions x %d patches of 101x101 per step%s), "
                               "random-init configs/model/spgan.yaml, synthetic latents" % (
                                   B, th, tw, n_pos, B, ", lattice positions sharded over the ranks + one all-gather of the patches" if sharded else ""),
                   "batch_per_gpu": B, "patches_per_step_per_gpu": per_gpu_patches,
                   "l2": "inputs and activations larger than L2 (latent canvas %d MB, > 1 GB of operands per patch batch)" % m["canvas_mb"],
                   "precision_mode": args.precision, "ts_layer_precision": modes,
                   "execution": ("one CUDA graph per step, lattice positions on %d concurrent branches" % args.streams if m["graphs"]
                                 else "eager launches, %d streams" % args.streams) + ", %d lattice positions per generator call" % m["group"],
                   "algorithmic_tflop_per_step_per_gpu": per_gpu_patches * PATCH_GFLOP / 1000.0},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"], "ms_per_step": ms_e2e},
        "gpu_launches": m["launches"] * args.steps, "tcgen05_gemm_launches": m["gemm_launches"] * args.steps,
        "achieved_model_tflops": jobs * n_pos * PATCH_GFLOP / 1000.0 / (ms_step / 1000.0),
        "clocks": m["clocks"], "roofline": roofline,
    }
    if "call_ms" in m:
        out["call_ms"] = m["call_ms"]
    if strict is not None:
        out["strict_bf16x3"] = strict
    if fast is not None:
        out["fast_tail"] = fast
    if pano768 is not None:
        out["pano768"] = pano768
    if train is not None:
        out["train"] = train
    if not sharded and not args.no_hbm and not args.skip_e2e and rank == 0 and world == 1:
        try:
            out["hbm_kernels"] = measure_hbm_kernels(dev, local)
        except Exception as e:  # the headline line must not depend on the side table
            out["hbm_kernels"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if not args.no_cpu_baseline and not sharded:
        out["cpu_baseline"] = cpu_baseline_object(args)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def measure_hbm_kernels(dev, local):
    """The HBM-bound kernels of the path (north_star: gather / upfirdn2d / bias-act against the B200 HBM peak) timed alone in
    this run: CUDA events around ~0.3 s of back-to-back launches after 5 warm-up launches, inputs larger than L2, clocks sampled
    during the loops.  `frac` = algorithmic bytes / time / MEASURED_PEAKS.json hbm_gbs (the copy peak; burst figure: each kernel runs
    alone).  tools/microbench.py is the full version (graph replays, legacy A/B, configs[4] sweep)."""
    import torch
    import spgan_b200.functional as SF
    from spgan_b200 import grids, panorama
    hbm = float(load_peaks().get("hbm_gbs", 6549.8))
    g = torch.Generator(device=dev).manual_seed(3)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    k3 = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]], device=dev) / 4
    k4 = torch.tensor([1., 3., 3., 1.], device=dev)
    k4 = (k4[None, :] * k4[:, None]) / 64
    pl = panorama.plan(384, 768)
    cp, _ = panorama.patch_inputs(pl, 2, 7, 27, pl["lat_h"], pl["lat_w"])
    cases = []
    x = rn(32, 512, 101, 101)
    b = rn(512)
    cases.append(("bias_act fwd (32,512,101,101)", lambda: SF.bias_act(x, b, None, 3, 0, 0.2, 2 ** 0.5), 8.0 * x.numel()))
    go = rn(32, 512, 101, 101)
    cases.append(("bias_act bwd + bias reduction (32,512,101,101)",
                  lambda: SF.FusedLeakyReLUFunctionBackward.apply(go, x, 0.2, 2 ** 0.5), 12.0 * x.numel()))
    xg = rn(32, 512, 105, 105)
    cases.append(("upfirdn2d 3x3 pad 0 (32,512,105,105)", lambda: SF.upfirdn2d(xg, k3, pad=(0, 0)),
                  4.0 * 32 * 512 * (105 * 105 + 103 * 103)))
    xd = rn(32, 256, 101, 101)
    cases.append(("upfirdn2d 4x4 pad 2 (32,256,101,101)", lambda: SF.upfirdn2d(xd, k4, pad=(2, 2)),
                  4.0 * 32 * 256 * (101 * 101 + 102 * 102)))
    cases.append(("upfirdn2d down 2 4x4 (32,256,101,101)", lambda: SF.upfirdn2d(xd, k4, up=1, down=2, pad=(1, 1)),
                  4.0 * 32 * 256 * (101 * 101 + 50 * 50)))
    xu = rn(32, 256, 53, 53)
    cases.append(("upfirdn2d up 2 4x4 (32,256,53,53)", lambda: SF.upfirdn2d(xu, k4 * 4, up=2, down=1, pad=(2, 1)),
                  4.0 * 32 * 256 * (53 * 53 + 106 * 106)))
    z = rn(32, 256, 35, 35)
    grid = torch.from_numpy(grids.sampling_grid(35, 35, cp)).to(dev)
    cases.append(("sphere_gather (32,256,35,35) -> 9x", lambda: SF.sphere_gather_raw(z, grid), 4.0 * z.numel() * 10 + grid.numel() * 4))
    out = []
    sampler = ClockSampler(local)
    sampler.start()
    with torch.no_grad():
        for name, fn, nbytes in cases:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            # ~0.3 s per case, so that the 200 ms clock sampler sees every kernel under load
            n = max(40, min(4000, int(300.0 / max(e0.elapsed_time(e1) / 5, 0.02))))
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            gbs = nbytes / (ms / 1e3) / 1e9
            out.append({"case": name, "ms": round(ms, 4), "achieved_GBs": round(gbs, 1), "frac": round(gbs / hbm, 3)})
    clocks = sampler.stop()
    del cases, x, go, xg, xd, xu, z
    torch.cuda.empty_cache()
    return {"peak_GBs": hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)", "clocks": clocks, "cases": out,
            "note": "each kernel timed alone (~0.3 s of back-to-back launches, CUDA events, eager: the Python launch overhead is inside the figure), "
                    "inputs larger than L2"}


def ncu_traffic():
    """dram bytes per launch of the dominant GEMM launch, read from the committed ncu --set full summary of the final
    kernels (profiles/r2_ncu_full_pano.json, written by tools/ncu_summary.py from the capture)."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return t["dram_bytes_per_launch"], "dram__bytes_read.sum + dram__bytes_write.sum of %s from %s" % (t["launch"], t["source"])
    except Exception:
        return None, "no ncu --set full capture of the final kernels is committed yet"


# ------------------------------------------------------------------------------------------------ train workload
TRAIN_METRIC = "train_images_per_sec_101x101_full_G+D_step"
TRAIN_UNIT = "img/s"


def cpu_train_rate(batch=2, threads=None):
    """img/s of one D step + one G step (forward + backward, no regularisers) of the oracle port on the CPU."""
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases as K
    import spgan_oracle as O
    import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    gsd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "kernel" not in k and "noises." not in k)
           for k, v in K.generator_state_dict().items()}
    dsd = {k: v.clone().requires_grad_("kernel" not in k) for k, v in K.discriminator_state_dict().items()}
    gl, lat, coords, cps, noises, _ = K.generator_train_case()
    t0 = time.perf_counter()
    # D step
    with torch.no_grad():
        fake = O.generator_forward({k: v.detach() for k, v in gsd.items()}, gl, lat, coords, cps, noises, inject_index=5)
    real = synth.randn_t(9000, "bench_real", fake.shape).clamp(-1, 1)
    fp, rp = O.discriminator_forward(dsd, fake), O.discriminator_forward(dsd, real)
    (F.softplus(-rp[0]).mean() + F.softplus(fp[0]).mean()).backward()
    # G step
    fake = O.generator_forward(gsd, gl, lat, coords, cps, noises, inject_index=5)
    fp = O.discriminator_forward({k: v.detach() for k, v in dsd.items()}, fake)
    F.softplus(-fp[0]).mean().backward()
    dt = time.perf_counter() - t0
    return batch / dt, threads, dt


def run_train(args):
    import torch
    import torch.distributed as dist
    import spgan_b200.functional as SF
    import spgan_b200.lib as lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib.require_device()
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    SF.set_precision(args.precision)
    SF.set_gemm_pair_mode(args.pair_mode)
    out = measure_train(args, dev, world, rank, local)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def measure_train(args, dev, world, rank, local):
    """The train workload (BASELINE configs[2]); returns the JSON object on rank 0, None elsewhere.  The process group
    (world > 1) is the caller's."""
    import torch
    import torch.distributed as dist
    import spgan_b200.functional as SF
    import spgan_b200.lib as lib
    from spgan_b200.training import TrainStep

    B = args.train_batch
    ts = TrainStep(batch=B, device=dev, world=world, seed=9000, rank=rank, use_graphs=not args.no_graphs)
    graphs_on = ts.use_graphs
    tp = ts.config.train_params
    # "real" patches live in pinned host memory for the e2e leg (the dataloader side of train.py:205-215)
    g = torch.Generator(device="cpu").manual_seed(9000 + rank)
    host_real = torch.randn(B, 3, 101, 101, generator=g).clamp_(-1, 1).pin_memory()
    host_ac = (torch.rand(B, 3, generator=g) * 2 - 1).pin_memory()
    host_loss = torch.empty(4).pin_memory()

    parts = ("d", "r1", "g", "path", "ema")

    def one_iteration(acc, e2e):
        """Every part of the iteration runs, each bracketed by CUDA events (the schedule weights them afterwards)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(parts) + 1)]
        ev[0].record()
        real = (host_real, host_ac) if e2e else None  # pinned host memory: the step copies it into its static buffers
        ld = ts.d_step(real=real)
        ev[1].record()
        lr = ts.d_r1_step(real=real)
        ev[2].record()
        lg = ts.g_step()
        ev[3].record()
        lp = ts.g_path_step()
        ev[4].record()
        ts.ema_step()
        if e2e:
            host_loss.copy_(torch.stack([ld, lr, lg, lp]), non_blocking=True)
        ev[5].record()
        acc.append(ev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(steps, e2e, mark=False):
        acc = []
        barrier()
        # the timed region only: a process-wide NVTX range (backward kernels come from autograd's own thread)
        rng = torch.cuda.nvtx.range_start("spgan_timed") if mark else None
        for _ in range(steps):
            one_iteration(acc, e2e)
        if mark:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_end(rng)
        barrier()
        ms = {k: sum(ev[i].elapsed_time(ev[i + 1]) for ev in acc) / steps for i, k in enumerate(parts)}
        t = torch.tensor([ms[k] for k in parts], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(zip(parts, [float(v) for v in t.tolist()]))

    def amortised(ms):
        return ms["d"] + ms["g"] + ms["ema"] + ms["r1"] / tp.d_reg_every + ms["path"] / tp.g_reg_every

    # eager pass first: warm-up, launch count and CUDA-event times of the tcgen05 launches (a graph replay runs no Python,
    # so per-launch events cannot be recorded inside it); then the bodies are captured and the timed steps are replays
    ts.use_graphs = False
    measure(1, False)
    SF.profile_gemm(True)
    l0, g0 = lib.launches(), lib.load().spgan_gemm_launch_count()
    ms_eager = measure(1, False)
    l1, g1 = lib.launches(), lib.load().spgan_gemm_launch_count()
    gemm_stats = SF.profile_gemm(False)
    ts.use_graphs = graphs_on
    try:
        measure(args.warmup if args.skip_e2e else max(args.warmup, 3), False)  # two more eager runs, then capture + replay
    except Exception as e:  # capture failed: fall back to eager launches on every rank, and say so
        if not graphs_on:
            raise
        sys.stderr.write("CUDA graph capture failed (%s: %s); falling back to eager launches\n" % (type(e).__name__, str(e)[:300]))
        torch.cuda.synchronize()
        graphs_on = False
        ts = TrainStep(batch=B, device=dev, world=world, seed=9000, rank=rank, use_graphs=False)
        measure(max(args.warmup, 3), False)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = measure(args.steps, False, mark=True)
    clocks = sampler.stop() if sampler else None
    ms_e2e = measure(args.steps, True) if not args.skip_e2e else {k: float("nan") for k in parts}
    call_ms = None
    if args.profile_calls:
        ts.use_graphs = False  # per-call events need the eager path
        SF.profile_calls(True)
        ms_prof = measure(1, False)
        call_ms = {k: [round(v[0], 3), v[1]] for k, v in sorted(SF.profile_calls(False).items(), key=lambda kv: -kv[1][0])}
        call_ms["_all_parts_ms"] = sum(ms_prof.values())
        ts.use_graphs = graphs_on
    from spgan_b200.training import replicas_in_sync
    in_sync = replicas_in_sync([ts.G, ts.D], world)  # collective: every rank calls it
    if rank != 0:
        return None
    ms_step = amortised(ms)
    value = world * B / (ms_step / 1000.0)
    e2e_value = world * B / (amortised(ms_e2e) / 1000.0)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = gemm_stats["flops"] / (gemm_stats["ms"] / 1000.0) / 1e12 if gemm_stats["ms"] > 0 else 0.0
    all_ms = sum(ms_eager.values())
    out = {
        "metric": TRAIN_METRIC, "value": value, "unit": TRAIN_UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {0: "fp32 (SIMT)", 1: "fp32-equivalent (bf16x3 split on tcgen05, fp32 accumulate)", 2: "bf16 (tcgen05, fp32 accumulate)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "full G+D training iteration (StyleGAN2 discriminator, R1 every %d, path-length every %d at batch %d), "
                               "batch %d/GPU data-parallel, 101x101 patches, random-init configs/model/spgan.yaml" % (
                                   tp.d_reg_every, tp.g_reg_every, max(1, B // tp.path_batch_shrink), B),
                   "batch_per_gpu": B, "parallelism": "dp%d" % world,
                   "schedule": "every timed step runs D, R1, G, path-length and EMA; ms_per_step = D + G + EMA + R1/%d + path/%d "
                               "(the reference's lazy-regularisation cadence, train.py:288,379)" % (tp.d_reg_every, tp.g_reg_every),
                   "part_ms": ms, "part_ms_eager_launches": ms_eager, "cuda_graphs": graphs_on,
                   "replicas_in_sync_after_run": in_sync,
                   "l2": "activations of one iteration exceed L2 (> 1 GB)", "precision_mode": args.precision},
        "e2e": {"value": e2e_value, "unit": TRAIN_UNIT, "h2d_bytes_per_step": 2 * (host_real.numel() + host_ac.numel()) * 4,
                "d2h_bytes_per_step": 16, "ms_per_step": amortised(ms_e2e)},
        "gpu_launches": (l1 - l0) * args.steps, "tcgen05_gemm_launches": int(g1 - g0) * args.steps, "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel + conv_wgrad_gemm_kernel (tcgen05 forward, data-gradient and weight-gradient GEMMs)",
                     "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
                     "traffic": None, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s",
                     "launches_timed": gemm_stats["launches"],
                     "avg_launch_ms": gemm_stats["ms"] / max(gemm_stats["launches"], 1),
                     "share_of_step": gemm_stats["ms"] / (all_ms if all_ms else 1.0),
                     "note": "launch times from one eager all-parts iteration (CUDA events cannot be recorded inside a graph replay)"},
    }
    if call_ms is not None:
        out["call_ms"] = call_ms
    if not args.no_cpu_baseline:
        rate, threads, dt = cpu_train_rate(2)
        out["cpu_baseline"] = {"value": rate, "unit": TRAIN_UNIT, "cores": threads, "kind": "port",
                               "sample": "one D step + one G step (forward + backward, no regularisers) at batch 2 (%.1f s), oracle port on CPU" % dt}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "train":
        run_train(a)
    else:
        run_ours(a)
