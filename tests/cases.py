"""Shared case tables for the parity tests (must stay in step with oracle/make_golden.py, which wrote the fixtures)."""
import json
import os

import numpy as np
import torch

import spgan_oracle as O
import synth

SEED = 9000
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STYLE_DIM = 8

UPFIRDN_CASES = [
    ("g_blur3", (2, 3, 13, 13), [1, 2, 1], 4.0, 1, 1, (0, 0)),
    ("d_blur4_main", (2, 3, 13, 13), [1, 3, 3, 1], 1.0, 1, 1, (2, 2)),
    ("d_blur4_skip", (2, 3, 12, 12), [1, 3, 3, 1], 1.0, 1, 1, (1, 1)),
    ("up2_4", (1, 2, 7, 9), [1, 3, 3, 1], 4.0, 2, 1, (2, 1)),
    ("down2_4", (1, 2, 10, 8), [1, 3, 3, 1], 1.0, 1, 2, (1, 1)),
    ("up2_3", (1, 2, 6, 5), [1, 2, 1], 4.0, 2, 1, (1, 1)),
    ("negpad", (1, 2, 9, 9), [1, 2, 1], 1.0, 1, 1, (-1, 0)),
]

MODCONV_CASES = [
    # name, cin, cout, k, demod, upsample, B, H
    ("k3", 6, 5, 3, True, False, 2, 9),
    ("k7", 7, 4, 7, True, False, 2, 11),
    ("k1_nodemod", 8, 3, 1, False, False, 2, 6),
    ("k3_up", 6, 5, 3, True, True, 2, 5),
]


def train_cp(x_st, y_st, size=35, gx=45, gy=140, circular=None):
    if circular is None:
        circular = y_st + size > gy
    return {"p_x_st": x_st / gx, "p_x_ed": (x_st + size - 1) / gx, "p_y_st": y_st / gy,
            "p_y_ed": (y_st + size - 1) / gy, "circular_flag": bool(circular), "x_total": gx, "y_total": gy,
            "y_st": y_st, "y_ed": y_st + size, "partial": 0.6667}


def test_cp(ix, iy, it):
    plan = O.close_loop_plan(384, 768)
    cp, _ = O.patch_coords_partial(plan, ix, iy, plan["lat_h"], plan["lat_w"], it)
    return cp


SPHERE_CASES = [
    # name, B, C, cout, h, coords_partial
    ("train_b2", 2, 4, 5, 17, [train_cp(7, 139, 17), train_cp(1, 20, 17)]),
    ("test_b3", 3, 4, 5, 11, test_cp(2, 7, 27)),
    ("train_b1", 1, 5, 4, 23, [train_cp(3, 60, 23)]),
]


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def module_params(tag, shapes):
    """Parameters as make_golden.fill_module wrote them: randn(seed, tag + name), modulation.bias shifted by 1."""
    return {n: synth.randn_t(SEED, tag + n, s, 1.0, 1.0 if n.endswith("modulation.bias") else 0.0) for n, s in shapes.items()}


def modconv_params(name, cin, cout, k):
    return module_params("mc_" + name + "_", {"weight": (1, cout, cin, k, k), "modulation.weight": (cin, STYLE_DIM),
                                               "modulation.bias": (cin,)})


def sphere_params(name, cin_total, cout):
    return module_params("smc_" + name + "_", {"weight": (1, cout, cin_total, 3, 3), "modulation.weight": (cin_total, STYLE_DIM),
                                                "modulation.bias": (cin_total,)})


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def t2n(t):
    return t.detach().cpu().numpy()


def generator_case(name, B, ix, iy):
    """Inputs of the golden generator cases (make_golden.golden_generator)."""
    plan = O.close_loop_plan(384, 768)
    cp, (zx_st, zx_ed, zy_st, zy_ed) = O.patch_coords_partial(plan, ix, iy, plan["lat_h"], plan["lat_w"], 5)
    gl = synth.randn_t(SEED, "gen_gl_" + name, (B, 512))
    gl = torch.stack([gl, gl], 1)
    canvas = synth.randn_t(SEED, "gen_canvas_" + name, (B, 256, plan["lat_h"], plan["lat_w"]))
    coords_full = O.meta_coord_grid(plan["lat_h"], plan["lat_w"]).unsqueeze(0).repeat(B, 1, 1, 1)
    lat = O.circular_slice(canvas, plan["lat_w"], zx_st, zx_ed, zy_st, zy_ed).contiguous()
    coords = O.circular_slice(coords_full, plan["lat_w"], zx_st, zx_ed, zy_st, zy_ed).contiguous()
    noises = [synth.randn_t(SEED, "gen_noise%d_%s" % (l, name), (B, 1, s, s)) for l, s in enumerate(plan["out_sizes"])]
    return gl, lat, coords, cp, noises


def generator_state_dict():
    return synth.synthetic_state_dict(load_json("generator_manifest.json"), SEED)


def compact_check(golden, key, got, tol):
    """Compare `got` with a fixture entry that may be stored compacted (oracle/make_golden.py:compact)."""
    got = np.asarray(got)
    if key in golden:
        return rel_err(got, golden[key]) < tol
    ref = golden[key + "__sample"]
    stride = -(-got.size // 4096)
    sample = got.reshape(-1)[::stride]
    norm = float(np.sqrt((got.astype(np.float64) ** 2).sum()))
    ref_norm = float(golden[key + "__norm"])
    scale = max(np.abs(ref).max(), ref_norm / np.sqrt(got.size), 1e-30)
    return float(np.abs(sample - ref).max() / scale) < tol and abs(norm - ref_norm) <= tol * ref_norm


def compact_l2(golden, key, got):
    """(relative L2 error, relative norm difference) of `got` against a (possibly compacted) fixture entry.

    Used for NETWORK-level gradients: the leaky-ReLU gate makes the gradient a discontinuous function of the forward
    activations, so a forward perturbation of relative size e flips a fraction ~0.4 e of the gates and moves the
    gradient by ~sqrt(0.4 e) in relative L2 (1e-3 for fp32 rounding noise of 5e-6) whatever the implementation; the
    per-op gradient tests hold the tight element-wise bounds."""
    got = np.asarray(got)
    if key in golden:
        ref = np.asarray(golden[key]).reshape(-1)  # small tensor stored in full: the L2 error covers the norm
        sample = got.reshape(-1)
        l2 = float(np.sqrt(((sample.astype(np.float64) - ref) ** 2).sum()) / max(np.sqrt((ref.astype(np.float64) ** 2).sum()), 1e-30))
        return l2, 0.0
    else:
        ref = golden[key + "__sample"]
        stride = -(-got.size // 4096)
        sample = got.reshape(-1)[::stride]
        ref_norm = float(golden[key + "__norm"])
    norm = float(np.sqrt((got.astype(np.float64) ** 2).sum()))
    l2 = float(np.sqrt(((sample.astype(np.float64) - ref) ** 2).sum()) / max(np.sqrt((ref.astype(np.float64) ** 2).sum()), 1e-30))
    return l2, abs(norm - ref_norm) / max(ref_norm, 1e-30)


def discriminator_state_dict():
    """As oracle/make_golden.py:synthetic_d_state: FIR buffers keep their values, biases ~ 0.1 N(0,1), weights N(0,1)."""
    sd = {}
    for k, shape in load_json("discriminator_manifest.json").items():
        if k.endswith("kernel"):
            sd[k] = torch.from_numpy(O.make_kernel([1, 3, 3, 1]))
        elif k.endswith(".bias"):
            sd[k] = synth.randn_t(SEED, "d_" + k, shape, 0.1)
        else:
            sd[k] = synth.randn_t(SEED, "d_" + k, shape)
    return sd


D_GRAD_KEYS = ["convs.0.0.weight", "convs.1.conv1.0.weight", "convs.1.conv2.1.weight", "convs.1.skip.1.weight",
               "convs.3.conv2.2.bias", "final_conv.0.weight", "final_linear.0.weight", "coord_linear.1.weight"]

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


def generator_train_case():
    """Inputs of make_golden.golden_generator_train."""
    B = 2
    cps = [train_cp(3, 17), train_cp(8, 120)]
    gl = synth.randn_t(SEED, "tr_gl", (B, 2, 512))
    lat = synth.randn_t(SEED, "tr_lat", (B, 256, 35, 35))
    coords_full = O.meta_coord_grid(80, 180)
    coords = torch.stack([coords_full[:, 3:38, 17:52], coords_full[:, 8:43, 120:155]]).contiguous()
    noises = [synth.randn_t(SEED, "tr_noise%d" % l, (B, 1, s, s)) for l, s in enumerate(O.TS_FEATURE_SIZES)]
    go = synth.randn_t(SEED, "tr_go", (B, 3, 101, 101))
    return gl, lat, coords, cps, noises, go


def reference_gradient_sensitivity(net, key, fwd_err):
    """How far the REFERENCE's own gradient `key` of `net` ("generator" / "discriminator") moves (relative L2) when its
    forward output moves by `fwd_err` (relative L2), read off tests/golden/sensitivity.json: the reference was re-run
    with its input perturbed by eps = 1e-6 .. 1e-3 and both changes were recorded (oracle/make_golden_r2.py).  Log-log
    interpolation on the (forward change, gradient change) curve; keys that were not recorded use the largest recorded
    curve of the same network.  Below the smallest recorded forward change the curve is held flat: that first point is
    the reference's own noise floor under an fp32-rounding-sized perturbation."""
    sens = load_json("sensitivity.json")[net]
    fwd = np.asarray(sens["fwd"], dtype=np.float64)
    names = [key] if key in sens else [k for k in sens if k != "fwd"]
    best = 0.0
    for k in names:
        g = np.asarray(sens[k], dtype=np.float64)
        order = np.argsort(fwd)
        v = float(np.exp(np.interp(np.log(max(fwd_err, 1e-30)), np.log(fwd[order]), np.log(g[order]))))
        best = max(best, v)
    return best


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-300))
