"""TEST INFRASTRUCTURE (oracle/): generate tests/golden/* by RUNNING THE REAL REFERENCE in this container.

    python oracle/make_golden.py            # needs /root/reference; CPU only; ~2 min

The reference ships no tests or golden vectors, so the pins are outputs of its own
Python implementation, imported read-only through `oracle/refimport.py` (five in-process stubs, no edits
to the reference tree), on inputs/weights that `oracle/synth.py` can regenerate anywhere.  Every fixture is
also compared here against `oracle/spgan_oracle.py`; a mismatch aborts the run, so a committed fixture set
means the oracle restatement agreed with the reference at generation time.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refimport  # noqa: E402

config = refimport.load_config()
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import spgan_oracle as O  # noqa: E402
import synth  # noqa: E402

from models.custom_ops import fused_leaky_relu, upfirdn2d  # noqa: E402  (reference, CPU branches)
import models.ops as ref_ops  # noqa: E402
import models.spgan_ops as ref_spops  # noqa: E402
import models.spgan_ops_gs as ref_gs  # noqa: E402
from models.spherenet import SphereConvBatchDiffFixBorderGNoGrad, GridSamplerNewTextureNoGrad  # noqa: E402
from models.spgan.spgan import InfinityGanGenerator  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)
SEED = 9000


def close(a, b, tol, what):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    print("  %-46s max rel-to-peak err %.3e" % (what, err))
    assert err <= tol, (what, err)


def t2n(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------ K1
def golden_bias_act():
    res = {}
    for name, shape in (("4d", (2, 5, 7, 3)), ("2d", (3, 6))):
        x = synth.randn_t(SEED, "ba_x_" + name, shape).requires_grad_(True)
        b = synth.randn_t(SEED, "ba_b_" + name, (shape[1],)).requires_grad_(True)
        go = synth.randn_t(SEED, "ba_go_" + name, shape)
        y = fused_leaky_relu(x, b)
        gx, gb = torch.autograd.grad(y, [x, b], go)
        res["y_" + name], res["gx_" + name], res["gb_" + name] = t2n(y), t2n(gx), t2n(gb)
        yo = O.fused_leaky_relu(t2n(x), t2n(b))
        gxo, gbo = O.fused_leaky_relu_backward(t2n(go), yo)
        close(yo, res["y_" + name], 1e-6, "bias_act fwd " + name)
        close(gxo, res["gx_" + name], 1e-6, "bias_act bwd " + name)
        close(gbo, res["gb_" + name], 1e-5, "bias_act bias-grad " + name)
    np.savez_compressed(os.path.join(OUT, "bias_act.npz"), **res)


# ------------------------------------------------------------------------------------------ K2
UPFIRDN_CASES = [
    # name, shape, taps, gain, up, down, pad
    ("g_blur3", (2, 3, 13, 13), [1, 2, 1], 4.0, 1, 1, (0, 0)),       # G: Blur after convT (ops.py:622)
    ("d_blur4_main", (2, 3, 13, 13), [1, 3, 3, 1], 1.0, 1, 1, (2, 2)),  # D: ConvLayer downsample 3x3 (disc:24-29)
    ("d_blur4_skip", (2, 3, 12, 12), [1, 3, 3, 1], 1.0, 1, 1, (1, 1)),  # D: skip 1x1
    ("up2_4", (1, 2, 7, 9), [1, 3, 3, 1], 4.0, 2, 1, (2, 1)),          # Upsample (ops.py:40-47)
    ("down2_4", (1, 2, 10, 8), [1, 3, 3, 1], 1.0, 1, 2, (1, 1)),       # Downsample (ops.py:72-75)
    ("up2_3", (1, 2, 6, 5), [1, 2, 1], 4.0, 2, 1, (1, 1)),
    ("negpad", (1, 2, 9, 9), [1, 2, 1], 1.0, 1, 1, (-1, 0)),
]


def golden_upfirdn2d():
    res = {}
    for name, shape, taps, gain, up, down, pad in UPFIRDN_CASES:
        k = torch.from_numpy(O.make_kernel(taps) * np.float32(gain))
        x = synth.randn_t(SEED, "ufd_x_" + name, shape).requires_grad_(True)
        y = upfirdn2d(x, k, up=up, down=down, pad=pad)
        go = synth.randn_t(SEED, "ufd_go_" + name, y.shape)
        gx, = torch.autograd.grad(y, x, go)
        res["y_" + name], res["gx_" + name] = t2n(y), t2n(gx)
        p4 = (pad[0], pad[1], pad[0], pad[1])
        yo = O.upfirdn2d(t2n(x), t2n(k), (up, up), (down, down), p4)
        gxo = O.upfirdn2d_backward(t2n(go), t2n(k), (up, up), (down, down), p4, shape)
        close(yo, res["y_" + name], 1e-6, "upfirdn2d fwd " + name)
        close(gxo, res["gx_" + name], 1e-6, "upfirdn2d bwd " + name)
    np.savez_compressed(os.path.join(OUT, "upfirdn2d.npz"), **res)


# ------------------------------------------------------------------------------------------ grids
def train_cp(x_st, y_st, size=35, gx=45, gy=140, circular=None):
    """coords_partial as coord_handler.py:1027-1038 builds it for training."""
    if circular is None:
        circular = y_st + size > gy
    return {"p_x_st": x_st / gx, "p_x_ed": (x_st + size - 1) / gx, "p_y_st": y_st / gy,
            "p_y_ed": (y_st + size - 1) / gy, "circular_flag": bool(circular), "x_total": gx, "y_total": gy,
            "y_st": y_st, "y_ed": y_st + size, "partial": 0.6667}


GRID_CASES = [
    ("train_0_0_35", 35, train_cp(0, 0)),
    ("train_9_120_35", 35, train_cp(9, 120)),       # wraps in longitude
    ("train_4_77_29", 29, train_cp(4, 77)),         # cropped layer reuses the uncropped range
    ("train_7_139_17", 17, train_cp(7, 139)),
    ("train_2_30_53", 53, train_cp(2, 30)),
]


def ref_grid(module_cls_grid_fn, h, cp):
    return module_cls_grid_fn(h, h, 1, cp)


def golden_grids(gen):
    """Reference: ModulatedConv2d.genSamplingPattern (models/spgan_ops_gs.py:410-428)."""
    mod = gen.structure_synthesizer.implicit_model.conv_stack[0].conv.conv
    plan = O.close_loop_plan(384, 768)
    test_cases = []
    for (ix, iy) in ((0, 0), (2, 7), (5, 9), (3, 8)):
        cp, _ = O.patch_coords_partial(plan, ix, iy, plan["lat_h"], plan["lat_w"], ix * plan["steps_w"] + iy)
        test_cases.append(("test_%d_%d_35" % (ix, iy), 35, cp))
        test_cases.append(("test_%d_%d_23" % (ix, iy), 23, cp))
    res = {}
    for name, h, cp in GRID_CASES + test_cases:
        if cp.get("test_flag", False):
            mod.genSamplingPattern(h, h, stride=1, coords_partial=cp)
            g = t2n(mod.grid)
        else:
            g = t2n(mod.genSamplingPattern(h, h, stride=1, coords_partial=cp))
        go = O.gen_sampling_grid(h, h, cp)
        assert g.dtype == np.float32 and go.dtype == np.float32
        assert np.array_equal(g.view(np.uint32), go.view(np.uint32)), "grid not bit-exact: " + name
        res[name] = g
        x0, y0, _, _ = O.gather_indices(g, h, h)
        res[name + "_x0"], res[name + "_y0"] = x0.astype(np.int16), y0.astype(np.int16)
    print("  grids bit-exact vs reference: %d cases" % (len(GRID_CASES) + len(test_cases)))
    np.savez_compressed(os.path.join(OUT, "grids.npz"), **res)
    with open(os.path.join(OUT, "grid_cases.json"), "w") as f:
        json.dump({name: {"h": h, "cp": {k: (bool(v) if isinstance(v, (bool, np.bool_)) else v) for k, v in cp.items()}}
                   for name, h, cp in GRID_CASES + test_cases}, f, indent=1, default=float)

    # checksum-of-checksums over the whole finite training grid set (10 x 140 starts x the 4 SS sizes)
    outer_ref = hashlib.sha256()
    outer_orc = hashlib.sha256()
    n = 0
    for x_st in range(10):
        for y_st in range(0, 140, 7):
            cp = train_cp(x_st, y_st)
            for h in (35, 29, 23, 17):
                g = t2n(mod.genSamplingPattern(h, h, stride=1, coords_partial=cp))
                go = O.gen_sampling_grid(h, h, cp)
                outer_ref.update(hashlib.sha256(g.tobytes()).digest())
                outer_orc.update(hashlib.sha256(go.tobytes()).digest())
                n += 1
    assert outer_ref.hexdigest() == outer_orc.hexdigest()
    print("  grid checksum-of-checksums over %d grids matches" % n)
    with open(os.path.join(OUT, "grid_checksum.json"), "w") as f:
        json.dump({"sha256_of_sha256": outer_ref.hexdigest(), "x_st": list(range(10)), "y_st": list(range(0, 140, 7)),
                   "sizes": [35, 29, 23, 17], "count": n}, f, indent=1)


# ------------------------------------------------------------------------------------------ gather
def golden_gather():
    """Reference: GridSamplerFuncNoGrad (models/spherenet/grid_generator.py:602-623)."""
    res = {}
    sampler = GridSamplerNewTextureNoGrad()
    for name, (B, C, h), cps in (("train", (2, 5, 17), [train_cp(7, 139, 17), train_cp(1, 20, 17)]),
                                 ("border", (1, 3, 11), [train_cp(9, 3, 11)])):
        grid = np.concatenate([O.gen_sampling_grid(h, h, cp) for cp in cps], 0)
        z = synth.randn_t(SEED, "gather_z_" + name, (B, C, h, h)).requires_grad_(True)
        y = sampler(z, torch.from_numpy(grid))
        go = synth.randn_t(SEED, "gather_go_" + name, y.shape)
        gz, = torch.autograd.grad(y, z, go)
        res["grid_" + name], res["y_" + name], res["gz_" + name] = grid, t2n(y), t2n(gz)
        close(O.grid_sample_border(t2n(z), grid), res["y_" + name], 2e-6, "gather fwd " + name)
        close(O.gather_surrogate_backward(t2n(go)), res["gz_" + name], 1e-6, "gather surrogate bwd " + name)
    np.savez_compressed(os.path.join(OUT, "gather.npz"), **res)


# ------------------------------------------------------------------------------------------ modulated convs
MODCONV_CASES = [
    # name, cin, cout, k, demod, upsample, B, H
    ("k3", 6, 5, 3, True, False, 2, 9),
    ("k7", 7, 4, 7, True, False, 2, 11),
    ("k1_nodemod", 8, 3, 1, False, False, 2, 6),
    ("k3_up", 6, 5, 3, True, True, 2, 5),
]
STYLE_DIM = 8


def fill_module(mod, tag):
    with torch.no_grad():
        for n, p in mod.named_parameters():
            p.copy_(synth.randn_t(SEED, tag + n, p.shape, 1.0, 1.0 if n.endswith("modulation.bias") else 0.0))


def golden_modconv():
    res = {}
    for name, cin, cout, k, demod, up, B, H in MODCONV_CASES:
        m = ref_ops.ModulatedConv2d(cin, cout, k, STYLE_DIM, demodulate=demod, upsample=up, no_zero_pad=True,
                                    blur_kernel=[1, 2, 1], config=config, side="ts")
        fill_module(m, "mc_" + name + "_")
        x = synth.randn_t(SEED, "mc_x_" + name, (B, cin, H, H)).requires_grad_(True)
        s = synth.randn_t(SEED, "mc_s_" + name, (B, STYLE_DIM)).requires_grad_(True)
        y, _ = m(x, s)
        go = synth.randn_t(SEED, "mc_go_" + name, y.shape)
        grads = torch.autograd.grad(y, [x, s, m.weight, m.modulation.weight, m.modulation.bias], go)
        res["y_" + name] = t2n(y)
        for gname, g in zip(("gx", "gs", "gw", "gmw", "gmb"), grads):
            res[gname + "_" + name] = t2n(g)
        blur = m.blur.kernel if up else None
        yo = O.modulated_conv2d(x, s, m.weight, m.modulation.weight, m.modulation.bias, demodulate=demod,
                                upsample=up, blur_kernel=blur)
        close(t2n(yo), res["y_" + name], 1e-5, "modconv fwd " + name)
        go_ = torch.autograd.grad(yo, [x, s, m.weight], go)
        close(t2n(go_[0]), res["gx_" + name], 1e-5, "modconv dX " + name)
        close(t2n(go_[2]), res["gw_" + name], 1e-5, "modconv dW " + name)
    np.savez_compressed(os.path.join(OUT, "modconv.npz"), **res)


def golden_sphere_modconv():
    """Reference: spgan_ops_gs.ModulatedConv2d deal_coords=True (models/spgan_ops_gs.py:700-816) incl. the
    batch>1 channel interleave of the (1, B*C) + (1, B*3) concatenation, in train (list) and test (dict) mode."""
    res = {}
    plan = O.close_loop_plan(384, 768)
    cp_test, _ = O.patch_coords_partial(plan, 2, 7, plan["lat_h"], plan["lat_w"], 27)
    for name, B, C, cout, h, cps in (("train_b2", 2, 4, 5, 17, [train_cp(7, 139, 17), train_cp(1, 20, 17)]),
                                     ("test_b3", 3, 4, 5, 11, cp_test),
                                     ("train_b1", 1, 5, 4, 23, [train_cp(3, 60, 23)])):
        m = ref_gs.ModulatedConv2d(C + 3, cout, 3, STYLE_DIM, no_zero_pad=True, config=config, side="ss", deal_coords=True)
        fill_module(m, "smc_" + name + "_")
        x = synth.randn_t(SEED, "smc_x_" + name, (B, C, h, h)).requires_grad_(True)
        c = synth.randn_t(SEED, "smc_c_" + name, (B, 3, h, h))
        s = synth.randn_t(SEED, "smc_s_" + name, (B, STYLE_DIM)).requires_grad_(True)
        y, _ = m(x, s, coords=c.clone(), coords_partial=cps)
        go = synth.randn_t(SEED, "smc_go_" + name, y.shape)
        gx, gs, gw = torch.autograd.grad(y, [x, s, m.weight], go)
        res["y_" + name], res["gx_" + name], res["gs_" + name], res["gw_" + name] = t2n(y), t2n(gx), t2n(gs), t2n(gw)
        grid = torch.from_numpy(O.batch_sampling_grid(h, h, cps, B))
        yo = O.sphere_modulated_conv2d(x, c, s, m.weight, m.modulation.weight, m.modulation.bias, grid)
        close(t2n(yo), res["y_" + name], 1e-5, "sphere modconv fwd " + name)
        gxo, gso, gwo = torch.autograd.grad(yo, [x, s, m.weight], go)
        close(t2n(gxo), res["gx_" + name], 1e-5, "sphere modconv dX " + name)
        close(t2n(gso), res["gs_" + name], 1e-5, "sphere modconv dS " + name)
        close(t2n(gwo), res["gw_" + name], 1e-5, "sphere modconv dW " + name)
    m = SphereConvBatchDiffFixBorderGNoGrad(3, 3)
    fill_module(m, "srgb_")
    x = synth.randn_t(SEED, "srgb_x", (2, 3, 17, 17)).requires_grad_(True)
    cps = [train_cp(7, 139, 17), train_cp(1, 20, 17)]
    y = m(x, cps)
    go = synth.randn_t(SEED, "srgb_go", y.shape)
    gx, gw, gb = torch.autograd.grad(y, [x, m.weight, m.bias], go)
    res["y_srgb"], res["gx_srgb"], res["gw_srgb"], res["gb_srgb"] = t2n(y), t2n(gx), t2n(gw), t2n(gb)
    yo = O.sphere_rgb_conv(x, m.weight, m.bias, torch.from_numpy(O.batch_sampling_grid(17, 17, cps, 2)))
    close(t2n(yo), res["y_srgb"], 1e-5, "sphere rgb conv fwd")
    np.savez_compressed(os.path.join(OUT, "sphere_modconv.npz"), **res)


# ------------------------------------------------------------------------------------------ generator + lattice
def golden_generator(gen):
    manifest = {k: list(v.shape) for k, v in gen.state_dict().items()}
    with open(os.path.join(OUT, "generator_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0)
    sd = synth.synthetic_state_dict(manifest, SEED)
    gen.load_state_dict(sd)
    gen.eval()
    plan = O.close_loop_plan(384, 768)
    res = {}
    with torch.no_grad():
        for name, B, (ix, iy) in (("b1_p27", 1, (2, 7)), ("b2_p59", 2, (5, 9))):
            cp, (zx_st, zx_ed, zy_st, zy_ed) = O.patch_coords_partial(plan, ix, iy, plan["lat_h"], plan["lat_w"], 5)
            gl = synth.randn_t(SEED, "gen_gl_" + name, (B, 512))
            gl = torch.stack([gl, gl], 1)
            canvas = synth.randn_t(SEED, "gen_canvas_" + name, (B, 256, plan["lat_h"], plan["lat_w"]))
            coords_full = O.meta_coord_grid(plan["lat_h"], plan["lat_w"]).unsqueeze(0).repeat(B, 1, 1, 1)
            lat = O.circular_slice(canvas, plan["lat_w"], zx_st, zx_ed, zy_st, zy_ed).contiguous()
            coords = O.circular_slice(coords_full, plan["lat_w"], zx_st, zx_ed, zy_st, zy_ed).contiguous()
            noises = [synth.randn_t(SEED, "gen_noise%d_%s" % (l, name), (B, 1, s, s)) for l, s in enumerate(plan["out_sizes"])]
            out = gen(global_latent=gl, local_latent=lat, override_coords=coords.clone(), coords_partial_override=cp,
                      noises=noises, disable_dual_latents=True)["gen"]
            res["img_" + name] = t2n(out)
            yo = O.generator_forward(sd, gl, lat, coords, cp, noises)
            close(t2n(yo), res["img_" + name], 2e-4, "generator fwd " + name)
    np.savez_compressed(os.path.join(OUT, "generator.npz"), **res)


def golden_lattice(gen):
    """Reference: BaseTestManager.__init__ + task_specific_init + the cursors of generate()."""
    from test_managers.close_loop_infinite_generation import InfiniteGenerationManagerPatchCoordsCloseLoop as Mgr
    EasyDict = refimport._AttrDict
    out = {}
    for (H, W) in ((384, 768), (768, 1536)):
        config.task = EasyDict({"height": H, "width": W, "batch_size": 1})
        config.train_params.batch_size = 1
        mgr = Mgr(gen, "cpu", "/tmp", config)
        mgr.task_specific_init()
        plan = O.close_loop_plan(H, W)
        ref = dict(pix_step=int(mgr.pixelspace_step_size), lat_step=int(mgr.latentspace_step_size),
                   outfeat_step=[int(v) for v in mgr.outfeat_step_sizes], out_sizes=[int(v) for v in mgr.outfeat_sizes_list],
                   steps_h=int(mgr.num_steps_h), steps_w=int(mgr.num_steps_w), steps_w_min=int(mgr.num_steps_w_min),
                   meta_h=int(mgr.meta_height), meta_w=int(mgr.meta_width),
                   noise_h=[int(v) for v in mgr.noise_heights], noise_w=[int(v) for v in mgr.noise_widths])
        tv = mgr.create_vars()
        ref["lat_h"], ref["lat_w"] = int(tv.local_latent.shape[2]), int(tv.local_latent.shape[3])
        for k, v in ref.items():
            assert plan[k] == v, (k, plan[k], v)
        mc = O.meta_coord_grid(ref["lat_h"], ref["lat_w"])
        assert torch.equal(mc, tv.meta_coords[0]), "meta coords differ"
        out["%dx%d" % (H, W)] = ref
        print("  lattice %dx%d matches: %s" % (H, W, ref))
        if H == 384:
            out["meta_coords_384_sha256"] = hashlib.sha256(t2n(tv.meta_coords[0]).tobytes()).hexdigest()
    with open(os.path.join(OUT, "lattice.json"), "w") as f:
        json.dump(out, f, indent=1)


def golden_panorama(gen):
    """One B=1 384x768 panorama through the reference manager vs the oracle's generate_panorama; keep a seam strip."""
    from test_managers.close_loop_infinite_generation import InfiniteGenerationManagerPatchCoordsCloseLoop as Mgr
    from test_managers.testing_vars_wrapper import TestingVars
    EasyDict = refimport._AttrDict
    config.task = EasyDict({"height": 384, "width": 768, "batch_size": 1})
    config.train_params.batch_size = 1
    mgr = Mgr(gen, "cpu", "/tmp", config)
    mgr.task_specific_init()
    plan = O.close_loop_plan(384, 768)
    gl = synth.randn_t(SEED, "pano_gl", (1, 512))
    gl = torch.stack([gl, gl], 1)
    canvas = synth.randn_t(SEED, "pano_canvas", (1, 256, plan["lat_h"], plan["lat_w"]))
    noises = [synth.randn_t(SEED, "pano_noise%d" % l, (1, 1, plan["noise_h"][l], plan["noise_w"][l])) for l in range(8)]
    meta_coords = mgr.coord_handler.sample_coord_grid(canvas, is_training=False)
    tv = TestingVars(meta_img=torch.zeros(1, 3, plan["meta_h"], plan["meta_w"]), global_latent=gl, local_latent=canvas,
                     meta_coords=meta_coords, noises=noises, device="cpu")
    with torch.no_grad():
        mgr.generate(tv, disable_pbar=True)
        mine = O.generate_panorama(gen.state_dict(), plan, gl, canvas, noises)
    close(t2n(mine), t2n(tv.meta_img), 2e-4, "384x768 panorama, oracle vs reference manager")
    img = t2n(tv.meta_img)
    np.savez_compressed(os.path.join(OUT, "panorama_384.npz"),
                        strip=img[:, :, 250:290, :].astype(np.float32),
                        col_seam=img[:, :, :, 740:768].astype(np.float32),
                        mean=np.float64(img.mean()), std=np.float64(img.std()))


# ------------------------------------------------------------------------------------------ training-path pins
def compact(res, limit=8192):
    """Keep the fixtures small: tensors above `limit` elements are stored as a strided sample of the flattened array
    plus its l2 norm (tests/cases.py:compact_view takes the same sample)."""
    out = {}
    for k, v in res.items():
        v = np.asarray(v)
        if v.size <= limit:
            out[k] = v
        else:
            stride = -(-v.size // 4096)
            out[k + "__sample"] = v.reshape(-1)[::stride].copy()
            out[k + "__norm"] = np.float64(np.sqrt((v.astype(np.float64) ** 2).sum()))
    return out


def synthetic_d_state(disc):
    sd = {}
    for k, v in disc.state_dict().items():
        if k.endswith("kernel"):
            sd[k] = v.clone()
        elif k.endswith(".bias"):
            sd[k] = synth.randn_t(SEED, "d_" + k, v.shape, 0.1)
        else:
            sd[k] = synth.randn_t(SEED, "d_" + k, v.shape)
    return sd


def golden_discriminator():
    """Reference: StyleGan2Discriminator (models/stylegan2discriminator.py) forward, first-order grads and the R1
    second-order gradient (models/losses.py:36-41) on a B = 2 batch."""
    from models.stylegan2discriminator import StyleGan2Discriminator
    torch.manual_seed(SEED)
    disc = StyleGan2Discriminator(config)
    manifest = {k: list(v.shape) for k, v in disc.state_dict().items()}
    with open(os.path.join(OUT, "discriminator_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0)
    sd = synthetic_d_state(disc)
    disc.load_state_dict(sd)
    disc.train()
    res = {}
    img = synth.randn_t(SEED, "d_img", (2, 3, 101, 101)).clamp(-1, 1).requires_grad_(True)
    out = disc(img)
    d, ac = out["d_patch"], out["ac_coords_pred"]
    res["d"], res["ac"] = t2n(d), t2n(ac)
    do, ao = O.discriminator_forward(sd, img)
    close(t2n(do), res["d"], 1e-5, "D forward d_patch")
    close(t2n(ao), res["ac"], 1e-5, "D forward ac_coords")
    names = ["convs.0.0.weight", "convs.1.conv1.0.weight", "convs.1.conv2.1.weight", "convs.1.skip.1.weight",
             "convs.3.conv2.2.bias", "final_conv.0.weight", "final_linear.0.weight", "coord_linear.1.weight"]
    params = dict(disc.named_parameters())
    loss = F.softplus(-d).mean() + (ac * synth.randn_t(SEED, "d_acw", ac.shape)).sum()
    grads = torch.autograd.grad(loss, [img] + [params[n] for n in names], retain_graph=True)
    res["g_img"] = t2n(grads[0])
    for n, g in zip(names, grads[1:]):
        res["g_" + n] = t2n(g)
    # R1
    r1 = O.d_r1_penalty(d, img)
    res["r1"] = t2n(r1)
    g2 = torch.autograd.grad(r1, [params[n] for n in names[:6]], allow_unused=True)
    for n, g in zip(names[:6], g2):
        res["r1g_" + n] = t2n(g)
    np.savez_compressed(os.path.join(OUT, "discriminator.npz"), **compact(res))
    print("  discriminator golden: d", res["d"].ravel(), "r1", float(res["r1"]))


def golden_second_order_ops():
    """Path-length style second-order gradients through ONE styled conv of each kind (plain 3x3, upsampling 3x3,
    spherical 3x3): g = d<y, n>/d style (create_graph) ; L = |g|^2 ; dL/d{weight, modulation.weight, x}."""
    res = {}
    for name, up in (("plain", False), ("up", True)):
        m = ref_ops.StyledConv(6, 5, 3, STYLE_DIM, upsample=up, blur_kernel=[1, 2, 1], no_zero_pad=True, config=config, side="ts")
        fill_module(m, "so_" + name + "_")
        with torch.no_grad():
            m.noise.weight.fill_(0.3)
        H = 7
        x = synth.randn_t(SEED, "so_x_" + name, (2, 6, H, H)).requires_grad_(True)
        s = synth.randn_t(SEED, "so_s_" + name, (2, STYLE_DIM)).requires_grad_(True)
        oh = m.calc_out_spatial_size(H)
        nz = synth.randn_t(SEED, "so_nz_" + name, (2, 1, oh, oh))
        y, _ = m(x, s, noise=nz)
        n = synth.randn_t(SEED, "so_n_" + name, y.shape)
        g, = torch.autograd.grad((y * n).sum(), s, create_graph=True)
        L = g.pow(2).sum()
        gw, gmw, gx = torch.autograd.grad(L, [m.conv.weight, m.conv.modulation.weight, x])
        res["y_" + name], res["g_" + name] = t2n(y), t2n(g)
        res["gw_" + name], res["gmw_" + name], res["gx_" + name] = t2n(gw), t2n(gmw), t2n(gx)
    # spherical
    cps = [train_cp(7, 139, 11), train_cp(1, 20, 11)]
    m = ref_gs.StyledConv(4 + 3, 5, 3, STYLE_DIM, no_zero_pad=True, disable_noise=True, config=config, activation="LeakyReLU_n",
                          side="ss", deal_coords=True)
    fill_module(m, "so_sph_")
    x = synth.randn_t(SEED, "so_x_sph", (2, 4, 11, 11)).requires_grad_(True)
    c = synth.randn_t(SEED, "so_c_sph", (2, 3, 11, 11))
    s = synth.randn_t(SEED, "so_s_sph", (2, STYLE_DIM)).requires_grad_(True)
    y, _ = m(x, s, coords=c.clone(), coords_partial=cps)
    n = synth.randn_t(SEED, "so_n_sph", y.shape)
    g, = torch.autograd.grad((y * n).sum(), s, create_graph=True)
    L = g.pow(2).sum()
    gw, gmw, gx = torch.autograd.grad(L, [m.conv.weight, m.conv.modulation.weight, x])
    res["y_sph"], res["g_sph"], res["gw_sph"], res["gmw_sph"], res["gx_sph"] = t2n(y), t2n(g), t2n(gw), t2n(gmw), t2n(gx)
    np.savez_compressed(os.path.join(OUT, "second_order.npz"), **res)
    print("  second-order op goldens written")


TRAIN_GRAD_KEYS = [
    "structure_synthesizer.implicit_model.conv_stack.0.conv.conv.weight",
    "structure_synthesizer.implicit_model.conv_stack.0.sc.weight",
    "structure_synthesizer.implicit_model.conv_stack.1.conv.conv.weight",
    "structure_synthesizer.implicit_model.conv_stack.2.conv.conv.modulation.weight",
    "texture_synthesizer.convs.0.conv.weight",
    "texture_synthesizer.convs.3.conv.weight",
    "texture_synthesizer.convs.6.activate.bias",
    "texture_synthesizer.convs.6.noise.weight",
    "texture_synthesizer.to_rgbs.1.conv.weight",
    "texture_synthesizer.sp_convs.0.weight",
    "texture_synthesizer.mapping.1.weight",
]


def golden_generator_train(gen):
    """Reference generator in train() mode: per-sample coords_partial list, style mixing at inject_index 5, first-order
    gradients of <img, go> w.r.t. the local latent and a spread of parameters."""
    sd = synth.synthetic_state_dict({k: list(v.shape) for k, v in gen.state_dict().items()}, SEED)
    gen.load_state_dict(sd)
    gen.train()
    B = 2
    cps = [train_cp(3, 17), train_cp(8, 120)]
    gl = synth.randn_t(SEED, "tr_gl", (B, 2, 512))
    lat = synth.randn_t(SEED, "tr_lat", (B, 256, 35, 35)).requires_grad_(True)
    coords_full = O.meta_coord_grid(80, 180)
    coords = torch.stack([coords_full[:, 3:38, 17:52], coords_full[:, 8:43, 120:155]]).contiguous()
    noises = [synth.randn_t(SEED, "tr_noise%d" % l, (B, 1, s, s)) for l, s in enumerate(O.TS_FEATURE_SIZES)]
    out = gen(global_latent=gl, local_latent=lat, override_coords=coords.clone(), coords_partial_override=cps,
              noises=noises, inject_index=5, disable_dual_latents=True)["gen"]
    go = synth.randn_t(SEED, "tr_go", out.shape)
    params = dict(gen.named_parameters())
    grads = torch.autograd.grad((out * go).sum(), [lat] + [params[k] for k in TRAIN_GRAD_KEYS])
    res = {"img": t2n(out), "g_lat": t2n(grads[0])}
    for k, g in zip(TRAIN_GRAD_KEYS, grads[1:]):
        res["g_" + k] = t2n(g)
    yo = O.generator_forward(sd, gl, lat, coords, cps, noises, inject_index=5)
    close(t2n(yo), res["img"], 2e-4, "generator train-mode fwd")
    np.savez_compressed(os.path.join(OUT, "generator_train.npz"), **compact(res))
    gen.eval()


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "gen", "pano", "train"]
    if "train" in which:
        golden_discriminator()
        golden_second_order_ops()
        torch.manual_seed(SEED)
        golden_generator_train(InfinityGanGenerator(config))
    if "ops" in which:
        golden_bias_act()
        golden_upfirdn2d()
        golden_gather()
        golden_modconv()
        golden_sphere_modconv()
    if "gen" in which or "pano" in which:
        torch.manual_seed(SEED)
        gen = InfinityGanGenerator(config)
        golden_grids(gen)
        golden_generator(gen)
        golden_lattice(gen)
        if "pano" in which:
            golden_panorama(gen)
    print("golden fixtures written to", OUT)
