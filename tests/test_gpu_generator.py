"""GPU parity: the whole generator patch forward and the close-loop panorama against the reference's outputs
(golden fixtures written by oracle/make_golden.py from the real reference)."""
import numpy as np
import pytest
import torch

import cases as K
import spgan_oracle as O
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gen():
    import spgan_b200.lib as lib
    from spgan_b200.generator import Generator
    torch.cuda.set_device(0)
    lib.require_device()
    g = Generator()
    g.load_state_dict(K.generator_state_dict())
    return g.cuda().eval()


# (global precision, texture-chain policy, bound): exact fp32 SIMT; the default policy (bf16x3, last conv fp16x2); strict
# bf16x3 everywhere.  The default policy is held to the SAME bound as strict bf16x3.
POLICIES = [(0, None, 2e-4), (1, None, 5e-4), (1, "strict", 5e-4)]


def _run_policy(gen, precision, policy, fn):
    import spgan_b200.functional as SF
    ts = gen.texture_synthesizer
    SF.set_precision(precision)
    ts.layer_precision = list(ts.STRICT) if policy == "strict" else policy
    try:
        with torch.no_grad():
            return fn()
    finally:
        SF.set_precision(1)
        ts.layer_precision = None


@pytest.mark.parametrize("precision,policy,tol", POLICIES)
@pytest.mark.parametrize("name,B,pos", [("b1_p27", 1, (2, 7)), ("b2_p59", 2, (5, 9))])
def test_generator_patch_golden(gen, name, B, pos, precision, policy, tol):
    g = K.load("generator.npz")
    gl, lat, coords, cp, noises = K.generator_case(name, B, *pos)
    img = _run_policy(gen, precision, policy, lambda: gen(gl.cuda(), lat.cuda(), coords.cuda(), cp, noises=[n.cuda() for n in noises]))
    assert img.shape == (B, 3, 101, 101)
    err = K.rel_err(K.t2n(img), g["img_" + name])
    print("precision %d policy %s %s: generator output error %.2e (bound %.1e)" % (precision, policy, name, err, tol))
    assert err < tol


def test_generator_autograd_path_matches_fused_path(gen):
    """The differentiable composition (training path) and the fused no_grad path agree."""
    gl, lat, coords, cp, noises = K.generator_case("b1_p27", 1, 2, 7)
    with torch.no_grad():
        a = gen(gl.cuda(), lat.cuda(), coords.cuda(), cp, noises=[n.cuda() for n in noises])
    lat_g = lat.cuda().requires_grad_(True)
    b = gen(gl.cuda(), lat_g, coords.cuda(), cp, noises=[n.cuda() for n in noises])
    assert K.rel_err(K.t2n(b), K.t2n(a)) < 5e-4
    b.square().mean().backward()
    assert lat_g.grad is not None and torch.isfinite(lat_g.grad).all() and lat_g.grad.abs().max() > 0


def test_panorama_384_golden_strip(gen):
    from spgan_b200 import panorama
    ref = K.load("panorama_384.npz")
    pl = panorama.plan(384, 768)
    gl = synth.randn_t(K.SEED, "pano_gl", (1, 512))
    gl = torch.stack([gl, gl], 1).cuda()
    canvas = synth.randn_t(K.SEED, "pano_canvas", (1, 256, pl["lat_h"], pl["lat_w"])).cuda()
    noises = [synth.randn_t(K.SEED, "pano_noise%d" % l, (1, 1, pl["noise_h"][l], pl["noise_w"][l])).cuda() for l in range(8)]
    # the golden panorama was produced by the reference manager with the same synthetic state dict (make_golden.py
    # loads it into the reference generator in golden_generator() before golden_panorama() runs)
    meta = panorama.generate(gen, pl, gl, canvas, noises)
    img = K.t2n(meta)
    assert img.shape == (1, 3, pl["meta_h"], pl["meta_w"])
    e_strip, e_seam = K.rel_err(img[:, :, 250:290, :], ref["strip"]), K.rel_err(img[:, :, :, 740:768], ref["col_seam"])
    print("384x768 panorama, default policy: strip error %.2e, seam error %.2e (bound 5e-4)" % (e_strip, e_seam))
    assert e_strip < 5e-4
    assert e_seam < 5e-4  # longitude seam columns
    assert abs(img.mean() - float(ref["mean"])) < 1e-3 * float(ref["std"])
    assert abs(img.std() - float(ref["std"])) < 1e-3 * float(ref["std"])


def test_panorama_sharded_positions_assemble_identically(gen):
    """Lattice positions are independent: two disjoint shards written into one meta image in row-major order equal
    the sequential result."""
    from spgan_b200 import panorama
    pl = panorama.plan(384, 768)
    B = 1
    gl = torch.randn(B, 512, device="cuda")
    canvas = torch.randn(B, 256, pl["lat_h"], pl["lat_w"], device="cuda")
    noises = [torch.randn(B, 1, pl["noise_h"][l], pl["noise_w"][l], device="cuda") for l in range(8)]
    pos = panorama.positions(pl)[:12]
    full = panorama.generate(gen, pl, gl, canvas, noises, only=set(pos))
    a = panorama.generate(gen, pl, gl, canvas, noises, only=set(pos[0::2]))
    b = panorama.generate(gen, pl, gl, canvas, noises, only=set(pos[1::2]))
    # merge: replay the row-major overwrite order using each shard's own patches
    merged = torch.zeros_like(full)
    for (ix, iy) in pos:
        src = a if (ix, iy) in set(pos[0::2]) else b
        px, py = ix * pl["pix_step"], iy * pl["pix_step"]
        patch = panorama.circular_slice(src, pl["meta_w"], px, px + 101, py, py + 101)
        panorama.circular_assign(merged, pl["meta_w"], px, px + 101, py, py + 101, patch)
    assert torch.equal(merged, full)


def test_style_memoisation_tracks_latent_updates(gen):
    """The per-layer (modulation, demodulation) memo must not survive an in-place change of the global latent."""
    gl, lat, coords, cp, noises = K.generator_case("b1_p27", 1, 2, 7)
    gl, lat, coords = gl.cuda(), lat.cuda(), coords.cuda()
    noises = [n.cuda() for n in noises]
    with torch.no_grad():
        styles = gen.texture_synthesizer.styles_for(gl)
        a1 = gen(gl, lat, coords, cp, noises=noises, styles=styles)
        a2 = gen(gl, lat, coords, cp, noises=noises, styles=styles)  # memo hit
        assert torch.equal(a1, a2)
        gl.mul_(0.5)
        styles.copy_(gen.texture_synthesizer.styles_for(gl))  # same storage, new contents
        b1 = gen(gl, lat, coords, cp, noises=noises, styles=styles)
        b2 = gen(gl.clone(), lat, coords, cp, noises=noises)  # fresh tensors, no memo
    assert K.rel_err(K.t2n(b1), K.t2n(b2)) < 1e-6
    assert K.rel_err(K.t2n(b1), K.t2n(a1)) > 1e-3


def test_chain_path_matches_module_path(gen):
    """The channels-last texture chain (no fp32 activations between convs, ToRGB in the GEMM epilogue) computes what the
    module-by-module path computes: identical operands bit for bit, ToRGB summed in a different order."""
    gl, lat, coords, cp, noises = K.generator_case("b2_p59", 2, 5, 9)
    args = (gl.cuda(), lat.cuda(), coords.cuda(), cp)
    nz = [n.cuda() for n in noises]
    ts = gen.texture_synthesizer
    ts.layer_precision = list(ts.STRICT)  # the module path runs bf16x3 in every conv
    try:
        with torch.no_grad():
            a = gen(*args, noises=nz)
            type(ts).use_chain = False
            try:
                b = gen(*args, noises=nz)
            finally:
                type(ts).use_chain = True
    finally:
        ts.layer_precision = None
    assert K.rel_err(K.t2n(a), K.t2n(b)) < 2e-5


@pytest.mark.parametrize("modes,tol", [([1, 1, 1, 1, 1, 1, 1, 3], 5e-4), ([1, 1, 1, 1, 1, 3, 3, 3], 7e-4), ([3] * 8, 2e-3)])
def test_generator_mixed_precision_tail_golden(gen, modes, tol):
    """Per-layer precision of the chain: the 2-MMA fp16 split on the layers nearest the output, whose rounding error is
    not amplified by later layers.  With the synthetic weights the activations reach 5e5, far beyond fp16's 65504: the
    operands are pre-scaled by calibrated powers of two (TextureSynthesizer.calibrate_act_scales), which is exact."""
    g = K.load("generator.npz")
    ts = gen.texture_synthesizer
    for name, B, pos in [("b1_p27", 1, (2, 7)), ("b2_p59", 2, (5, 9))]:
        gl, lat, coords, cp, noises = K.generator_case(name, B, *pos)
        ts.layer_precision = modes
        try:
            with torch.no_grad():
                nz = [n.cuda() for n in noises]
                _, styles, structure = gen(gl.cuda(), lat.cuda(), coords.cuda(), cp, noises=nz, return_latents=True)
                scales = ts.calibrate_act_scales(styles, structure, cp, nz)
                assert all(s > 0 and abs(np.log2(s) - round(np.log2(s))) < 1e-9 for s in scales)
                img = gen(gl.cuda(), lat.cuda(), coords.cuda(), cp, noises=nz)
        finally:
            ts.layer_precision = None
            ts.act_scale = None
        err = K.rel_err(K.t2n(img), g["img_" + name])
        print("modes %s %s: generator output error %.2e (bound %.1e)" % (modes, name, err, tol))
        assert err < tol, (name, modes)


@pytest.mark.parametrize("streams", [1, 3])
def test_panorama_engine_graph_equals_eager_loop(gen, streams):
    """PanoramaEngine (static buffers, one CUDA graph, lattice positions on concurrent branches) == panorama.generate,
    bit for bit, also after new inputs are loaded into the static buffers."""
    from spgan_b200 import panorama
    pl = panorama.plan(384, 768)
    B = 2
    only = set(panorama.positions(pl)[:7])
    eng = panorama.PanoramaEngine(gen, pl, B, "cuda:0", streams=streams, only=only, group=1)  # one position per call: bit-exact
    for seed in (1, 2, 3, 4):
        g = torch.Generator(device="cpu").manual_seed(seed)
        gl = torch.randn(B, 512, generator=g).cuda()
        canvas = torch.randn(B, 256, pl["lat_h"], pl["lat_w"], generator=g).cuda()
        noises = [torch.randn(B, 1, pl["noise_h"][l], pl["noise_w"][l], generator=g).cuda() for l in range(8)]
        want = panorama.generate(gen, pl, gl, canvas, noises, only=only)
        eng.load(gl, canvas, noises)
        got = eng.run()
        assert torch.equal(got, want), seed
    assert eng.graph is not None


@pytest.mark.parametrize("precision,policy,tol", POLICIES)
def test_generator_patch_golden_at_bench_batch_32(gen, precision, policy, tol):
    """The BENCHMARKED batch size: at B = 32 the reference's flat (1, B*C) ++ (1, B*3) concatenation under groups = B
    (models/spgan_ops_gs.py:792-814) maps channels across samples differently than at B = 1, 2 (generator.npz); the
    fixture is the real reference's output for this batch (oracle/make_golden_r2.py): strided sample + norm of the whole
    batch and three samples in full."""
    g = K.load("generator_b32.npz")
    gl, lat, coords, cp, noises = K.generator_case("b32_p34", 32, 3, 4)
    img = _run_policy(gen, precision, policy, lambda: gen(gl.cuda(), lat.cuda(), coords.cuda(), cp, noises=[n.cuda() for n in noises]))
    img = K.t2n(img)
    assert img.shape == (32, 3, 101, 101)
    peak = float(g["peak"])
    errs = [float(np.abs(img[b] - g["img_s%d" % b]).max()) / peak for b in (0, 17, 31)]
    print("B=32 precision %d policy %s: per-sample errors %s (bound %.1e)" % (precision, policy, ["%.2e" % e for e in errs], tol))
    assert K.compact_check(g, "img", img, tol)
    assert max(errs) < tol


def test_panorama_768_golden_strip_and_seam(gen):
    """BASELINE configs[3]: one 768x1536 close-loop panorama (180 lattice positions) against the reference manager's own
    output (tests/golden/panorama_768.npz), through the graphed multi-branch engine."""
    from spgan_b200 import panorama
    ref = K.load("panorama_768.npz")
    pl = panorama.plan(768, 1536)
    gl = synth.randn_t(K.SEED, "pano768_gl", (1, 512))
    canvas = synth.randn_t(K.SEED, "pano768_canvas", (1, 256, pl["lat_h"], pl["lat_w"]))
    noises = [synth.randn_t(K.SEED, "pano768_noise%d" % l, (1, 1, pl["noise_h"][l], pl["noise_w"][l])) for l in range(8)]
    eng = panorama.PanoramaEngine(gen, pl, 1, "cuda:0", streams=2, use_graph=False)
    eng.load(gl.cuda(), canvas.cuda(), [n.cuda() for n in noises])
    img = K.t2n(eng.run())
    assert img.shape == (1, 3, pl["meta_h"], pl["meta_w"])
    peak = float(ref["peak"])
    W = pl["target_w"]
    errs = [float(np.abs(img[:, :, 470:486, :] - ref["strip"]).max()) / peak,
            float(np.abs(img[:, :, :, W - 12:W] - ref["col_seam"]).max()) / peak,
            float(np.abs(img[:, :, 0:6, 0:512] - ref["top"]).max()) / peak]
    print("768x1536 panorama, default policy: strip / seam / top errors %s (bound 5e-4)" % ["%.2e" % e for e in errs])
    assert max(errs) < 5e-4
    assert abs(img.mean() - float(ref["mean"])) < 1e-3 * float(ref["std"])
    assert abs(img.std() - float(ref["std"])) < 1e-3 * float(ref["std"])


def test_position_group_at_bench_batch_32(gen):
    """The benchmarked configuration: two lattice positions of B = 32 as ONE generator call of 64 (block-diagonal flat-concat
    table of 2 x 32 groups, two sampling grids) against the two single-position calls."""
    from spgan_b200 import panorama
    pl = panorama.plan(384, 768)
    B = 32
    only = set(panorama.positions(pl)[9:11])
    g = torch.Generator(device="cpu").manual_seed(32)
    gl = torch.randn(B, 512, generator=g).cuda()
    canvas = torch.randn(B, 256, pl["lat_h"], pl["lat_w"], generator=g).cuda()
    noises = [torch.randn(B, 1, pl["noise_h"][l], pl["noise_w"][l], generator=g).cuda() for l in range(8)]
    want = panorama.generate(gen, pl, gl, canvas, noises, only=only)
    eng = panorama.PanoramaEngine(gen, pl, B, "cuda:0", streams=1, only=only, use_graph=False, group=2)
    eng.load(gl, canvas, noises)
    got = eng.run()
    err = K.rel_err(K.t2n(got), K.t2n(want))
    print("two positions x 32 patches as one call: %.2e" % err)
    assert err < 2e-6


def test_structure_chain_matches_module_path(gen):
    """The channels-last structure chain (256-channel main K segment + dense coordinate tail segment, shortcut as the
    residual of the spherical GEMM, packed / NHWC sinks between the convs) computes what the module-by-module path computes
    (259 channels padded to 320 per tap, NCHW fp32 tensors between the convs): same products, different summation order."""
    from spgan_b200.generator import ImplicitFunction
    for name, B, pos in [("b1_p27", 1, (2, 7)), ("b2_p59", 2, (5, 9))]:
        gl, lat, coords, cp, _ = K.generator_case(name, B, *pos)
        ss = gen.structure_synthesizer
        with torch.no_grad():
            a, _ = ss(gl.cuda()[:, 0], lat.cuda(), coords.cuda(), cp)
            ImplicitFunction.use_chain = False
            try:
                b, _ = ss(gl.cuda()[:, 0], lat.cuda(), coords.cuda(), cp)
            finally:
                ImplicitFunction.use_chain = True
        assert a.shape == b.shape == (B, 256, 11, 11)
        err = K.rel_err(K.t2n(a), K.t2n(b))
        print("structure chain vs module path, %s: %.2e" % (name, err))
        assert err < 5e-5


@pytest.mark.parametrize("group,streams,graph", [(3, 1, False), (2, 2, True)])
def test_panorama_engine_position_groups(gen, group, streams, graph):
    """Several lattice positions per generator call (grids.PositionGroup: stacked patch batches, one sampling grid per
    position, block-diagonal flat-concat table) give the patches of the one-position-per-call loop; 7 positions in groups of
    3 / 2 leave a ragged last group."""
    from spgan_b200 import panorama
    pl = panorama.plan(384, 768)
    B = 2
    only = set(panorama.positions(pl)[8:15])  # crosses a lattice row: two different latitude windows in one group
    eng = panorama.PanoramaEngine(gen, pl, B, "cuda:0", streams=streams, only=only, use_graph=graph, group=group)
    for seed in (1, 2, 3, 4):
        g = torch.Generator(device="cpu").manual_seed(seed)
        gl = torch.randn(B, 512, generator=g).cuda()
        canvas = torch.randn(B, 256, pl["lat_h"], pl["lat_w"], generator=g).cuda()
        noises = [torch.randn(B, 1, pl["noise_h"][l], pl["noise_w"][l], generator=g).cuda() for l in range(8)]
        want = panorama.generate(gen, pl, gl, canvas, noises, only=only)
        eng.load(gl, canvas, noises)
        got = eng.run()
        err = K.rel_err(K.t2n(got), K.t2n(want))
        print("position groups of %d, seed %d: %.2e" % (group, seed, err))
        # the GEMM rows are bit-identical; only the ToRGB partial sums are added in a different order (N-tile count)
        assert err < 2e-6, seed
    assert (eng.graph is not None) == graph


def test_fused_mapping_and_modulation_match_per_layer_path(gen):
    """the mapping network as one cluster kernel and the (modulation, demodulation) pairs of all 20 modulated
    convs as one launch against the module-by-module path (8 + 40 launches): styles, every pair, and the generator output."""
    from spgan_b200.generator import Generator, TextureSynthesizer
    gl, lat, coords, cp, noises = K.generator_case("b2_p59", 2, 5, 9)
    gl, lat, coords = gl.cuda(), lat.cuda(), coords.cuda()
    nz = [n.cuda() for n in noises]
    ts = gen.texture_synthesizer
    with torch.no_grad():
        TextureSynthesizer.use_fused_mapping = False
        Generator.use_fused_modulation = False
        try:
            styles_ref = ts.styles_for(gl)
            pairs_ref = [m._mod_demod(gl[:, 0] if sel else styles_ref[:, idx], 2) for m, sel, idx in gen._modulated_layers()]
            img_ref = gen(gl, lat, coords, cp, noises=nz, styles=styles_ref)
        finally:
            TextureSynthesizer.use_fused_mapping = True
            Generator.use_fused_modulation = True
        styles = ts.styles_for(gl)
        assert K.rel_err(K.t2n(styles), K.t2n(styles_ref)) < 2e-6
        gl2 = gl.clone()  # fresh storage: no memo entry yet
        launches = SF_launches()
        assert gen.prepare_modulation(gl2, styles)
        assert SF_launches() == launches + 1
        for (m, sel, idx), (s_ref, _, d_ref) in zip(gen._modulated_layers(), pairs_ref):
            s, _, d = m._mod_demod(gl2[:, 0] if sel else styles[:, idx], 2)
            assert K.rel_err(K.t2n(s), K.t2n(s_ref)) < 5e-6
            assert (d is None) == (d_ref is None)
            if d is not None:
                assert K.rel_err(K.t2n(d), K.t2n(d_ref)) < 5e-6
        assert SF_launches() == launches + 1  # every pair came from the memo
        img = gen(gl2, lat, coords, cp, noises=nz, styles=styles)
    assert K.rel_err(K.t2n(img), K.t2n(img_ref)) < 2e-5


def SF_launches():
    import spgan_b200.lib as lib
    return lib.launches()
