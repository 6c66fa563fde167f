"""TEST INFRASTRUCTURE (oracle/): drive the REAL reference (imported through refimport) — its InfinityGanGenerator and
its close-loop test manager (test_managers/close_loop_infinite_generation.py:170-305) — on CPU, or on the GPU over the
drop-in mirrors (`spgan_b200.dropin.install()`).  Used by oracle/make_golden*.py (fixtures), by the `-m gpu` drop-in
tests, and by bench.py's reference arm / cpu_baseline leg (timing the reference's own CPU path)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import refimport  # noqa: E402
import synth  # noqa: E402

SEED = 9000


def available():
    return refimport.available()


def load(dropin=False, real_cuda=False):
    """Import the reference (optionally with the op modules replaced by the mirrors) -> (config, InfinityGanGenerator)."""
    config = refimport.load_config(real_cuda=real_cuda)
    if dropin:
        root = os.path.dirname(HERE)
        if root not in sys.path:
            sys.path.insert(0, root)
        import spgan_b200.dropin as dropin_mod
        dropin_mod.install()
    from models.spgan.spgan import InfinityGanGenerator
    return config, InfinityGanGenerator


def synthetic_generator(config, cls, manifest=None, seed=SEED):
    """The reference generator with the synthetic state dict every fixture uses (oracle/synth.py)."""
    import torch
    torch.manual_seed(seed)
    gen = cls(config)
    if manifest is None:
        manifest = {k: list(v.shape) for k, v in gen.state_dict().items()}
    gen.load_state_dict(synth.synthetic_state_dict(manifest, seed))
    return gen.eval()


def manager(gen, config, device, H, W, batch=1):
    from test_managers.close_loop_infinite_generation import InfiniteGenerationManagerPatchCoordsCloseLoop as Mgr
    EasyDict = refimport._AttrDict
    config.task = EasyDict({"height": H, "width": W, "batch_size": batch})
    config.train_params.batch_size = batch
    mgr = Mgr(gen, device, "/tmp", config)
    mgr.task_specific_init()
    return mgr


def testing_vars(mgr, plan, tag, batch=1, seed=SEED, device="cpu"):
    """Inputs named `<tag>_gl`, `<tag>_canvas`, `<tag>_noise<l>` (the names tests/ regenerate), held on the CPU as the
    reference manager expects (close_loop_infinite_generation.py:84-168)."""
    import torch
    from test_managers.testing_vars_wrapper import TestingVars
    gl = synth.randn_t(seed, tag + "_gl", (batch, 512))
    gl = torch.stack([gl, gl], 1)
    canvas = synth.randn_t(seed, tag + "_canvas", (batch, 256, plan["lat_h"], plan["lat_w"]))
    noises = [synth.randn_t(seed, "%s_noise%d" % (tag, l), (batch, 1, plan["noise_h"][l], plan["noise_w"][l])) for l in range(8)]
    meta_coords = mgr.coord_handler.sample_coord_grid(canvas, is_training=False)
    # the canvases stay on the CPU (the manager moves every slice, close_loop_infinite_generation.py:204-230); the global
    # latent is created on the model's device, as create_vars does (:84-102)
    tv = TestingVars(meta_img=torch.zeros(batch, 3, plan["meta_h"], plan["meta_w"]), global_latent=gl.to(device),
                     local_latent=canvas, meta_coords=meta_coords, noises=noises, device=device)
    return tv, gl, canvas, noises


def time_reference_panorama(steps=1, warmup=0, threads=None, H=384, W=768):
    """Wall-clock seconds per full B = 1 panorama of the reference's own CPU path (all `threads` host threads):
    InfinityGanGenerator (random init, seed 9000) driven by its close-loop manager over every lattice position."""
    import time
    import torch
    import spgan_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    config, cls = load()
    torch.manual_seed(SEED)
    gen = cls(config).eval()
    mgr = manager(gen, config, "cpu", H, W)
    plan = O.close_loop_plan(H, W)
    tv, _, _, _ = testing_vars(mgr, plan, "bench")
    import contextlib
    times = []
    with torch.no_grad(), contextlib.redirect_stdout(sys.stderr):  # the manager prints banners on stdout
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            mgr.generate(tv, disable_pbar=True)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, threads, plan["steps_h"] * plan["steps_w"]
