"""CPU: the oracle (oracle/spgan_oracle.py) against the golden vectors produced by running the real reference
(oracle/make_golden.py).  This is the pin that makes the oracle trustworthy on machines without /root/reference."""
import hashlib

import numpy as np
import pytest
import torch

import cases as K
import spgan_oracle as O
import synth


def test_bias_act_golden():
    g = K.load("bias_act.npz")
    for name, shape in (("4d", (2, 5, 7, 3)), ("2d", (3, 6))):
        x = synth.randn(K.SEED, "ba_x_" + name, shape)
        b = synth.randn(K.SEED, "ba_b_" + name, (shape[1],))
        go = synth.randn(K.SEED, "ba_go_" + name, shape)
        y = O.fused_leaky_relu(x, b)
        gx, gb = O.fused_leaky_relu_backward(go, y)
        assert K.rel_err(y, g["y_" + name]) < 1e-6
        assert K.rel_err(gx, g["gx_" + name]) < 1e-6
        assert K.rel_err(gb, g["gb_" + name]) < 1e-5


@pytest.mark.parametrize("case", K.UPFIRDN_CASES, ids=[c[0] for c in K.UPFIRDN_CASES])
def test_upfirdn2d_golden(case):
    name, shape, taps, gain, up, down, pad = case
    g = K.load("upfirdn2d.npz")
    k = O.make_kernel(taps) * np.float32(gain)
    x = synth.randn(K.SEED, "ufd_x_" + name, shape)
    p4 = (pad[0], pad[1], pad[0], pad[1])
    y = O.upfirdn2d(x, k, (up, up), (down, down), p4)
    go = synth.randn(K.SEED, "ufd_go_" + name, y.shape)
    gx = O.upfirdn2d_backward(go, k, (up, up), (down, down), p4, shape)
    assert K.rel_err(y, g["y_" + name]) < 1e-6
    assert K.rel_err(gx, g["gx_" + name]) < 1e-6


def test_grids_bit_exact_golden():
    g = K.load("grids.npz")
    cases = K.load_json("grid_cases.json")
    assert len(cases) >= 13
    for name, c in cases.items():
        grid = O.gen_sampling_grid(c["h"], c["h"], c["cp"])
        assert np.array_equal(grid.view(np.uint32), g[name].view(np.uint32)), name
        x0, y0, _, _ = O.gather_indices(grid, c["h"], c["h"])
        assert np.array_equal(x0, g[name + "_x0"]) and np.array_equal(y0, g[name + "_y0"]), name


def test_grid_checksum_of_checksums():
    meta = K.load_json("grid_checksum.json")
    outer = hashlib.sha256()
    for x_st in meta["x_st"]:
        for y_st in meta["y_st"]:
            cp = K.train_cp(x_st, y_st)
            for h in meta["sizes"]:
                outer.update(hashlib.sha256(O.gen_sampling_grid(h, h, cp).tobytes()).digest())
    assert outer.hexdigest() == meta["sha256_of_sha256"]


def test_gather_golden():
    g = K.load("gather.npz")
    for name, (B, C, h) in (("train", (2, 5, 17)), ("border", (1, 3, 11))):
        z = synth.randn(K.SEED, "gather_z_" + name, (B, C, h, h))
        y = O.grid_sample_border(z, g["grid_" + name])
        assert K.rel_err(y, g["y_" + name]) < 2e-6
        go = synth.randn(K.SEED, "gather_go_" + name, y.shape)
        assert K.rel_err(O.gather_surrogate_backward(go), g["gz_" + name]) < 1e-6


@pytest.mark.parametrize("case", K.MODCONV_CASES, ids=[c[0] for c in K.MODCONV_CASES])
def test_modconv_golden(case):
    name, cin, cout, k, demod, up, B, H = case
    g = K.load("modconv.npz")
    p = K.modconv_params(name, cin, cout, k)
    w = p["weight"].requires_grad_(True)
    x = synth.randn_t(K.SEED, "mc_x_" + name, (B, cin, H, H)).requires_grad_(True)
    s = synth.randn_t(K.SEED, "mc_s_" + name, (B, K.STYLE_DIM)).requires_grad_(True)
    blur = torch.from_numpy(O.make_kernel([1, 2, 1]) * 4) if up else None
    y = O.modulated_conv2d(x, s, w, p["modulation.weight"], p["modulation.bias"], demodulate=demod, upsample=up, blur_kernel=blur)
    assert K.rel_err(K.t2n(y), g["y_" + name]) < 1e-5
    go = synth.randn_t(K.SEED, "mc_go_" + name, y.shape)
    gx, gs, gw = torch.autograd.grad(y, [x, s, w], go)
    assert K.rel_err(K.t2n(gx), g["gx_" + name]) < 1e-5
    assert K.rel_err(K.t2n(gs), g["gs_" + name]) < 1e-5
    assert K.rel_err(K.t2n(gw), g["gw_" + name]) < 1e-5


@pytest.mark.parametrize("case", K.SPHERE_CASES, ids=[c[0] for c in K.SPHERE_CASES])
def test_sphere_modconv_golden(case):
    name, B, C, cout, h, cps = case
    g = K.load("sphere_modconv.npz")
    p = K.sphere_params(name, C + 3, cout)
    w = p["weight"].requires_grad_(True)
    x = synth.randn_t(K.SEED, "smc_x_" + name, (B, C, h, h)).requires_grad_(True)
    c = synth.randn_t(K.SEED, "smc_c_" + name, (B, 3, h, h))
    s = synth.randn_t(K.SEED, "smc_s_" + name, (B, K.STYLE_DIM)).requires_grad_(True)
    grid = torch.from_numpy(O.batch_sampling_grid(h, h, cps, B))
    y = O.sphere_modulated_conv2d(x, c, s, w, p["modulation.weight"], p["modulation.bias"], grid)
    assert K.rel_err(K.t2n(y), g["y_" + name]) < 1e-5
    go = synth.randn_t(K.SEED, "smc_go_" + name, y.shape)
    gx, gs, gw = torch.autograd.grad(y, [x, s, w], go)
    assert K.rel_err(K.t2n(gx), g["gx_" + name]) < 1e-5
    assert K.rel_err(K.t2n(gs), g["gs_" + name]) < 1e-5
    assert K.rel_err(K.t2n(gw), g["gw_" + name]) < 1e-5


def test_generator_golden():
    g = K.load("generator.npz")
    sd = K.generator_state_dict()
    torch.set_num_threads(8)
    with torch.no_grad():
        gl, lat, coords, cp, noises = K.generator_case("b1_p27", 1, 2, 7)
        y = O.generator_forward(sd, gl, lat, coords, cp, noises)
    assert K.rel_err(K.t2n(y), g["img_b1_p27"]) < 2e-4


def test_lattice_golden():
    ref = K.load_json("lattice.json")
    for key, (H, W) in (("384x768", (384, 768)), ("768x1536", (768, 1536))):
        plan = O.close_loop_plan(H, W)
        for k, v in ref[key].items():
            assert plan[k] == v, (key, k)
    mc = O.meta_coord_grid(ref["384x768"]["lat_h"], ref["384x768"]["lat_w"])
    assert hashlib.sha256(K.t2n(mc).tobytes()).hexdigest() == ref["meta_coords_384_sha256"]


def test_discriminator_golden():
    g = K.load("discriminator.npz")
    sd = K.discriminator_state_dict()
    img = synth.randn_t(K.SEED, "d_img", (2, 3, 101, 101)).clamp(-1, 1).requires_grad_(True)
    torch.set_num_threads(8)
    d, ac = O.discriminator_forward(sd, img)
    assert K.rel_err(K.t2n(d), g["d"]) < 1e-5 and K.rel_err(K.t2n(ac), g["ac"]) < 1e-5
    r1 = O.d_r1_penalty(d, img)
    assert abs(float(r1) - float(g["r1"])) < 1e-5 * abs(float(g["r1"]))


def test_second_order_ops_golden():
    g = K.load("second_order.npz")
    for name, up in (("plain", False), ("up", True)):
        p = K.module_params("so_" + name + "_", {"conv.weight": (1, 5, 6, 3, 3), "conv.modulation.weight": (6, K.STYLE_DIM),
                                                  "conv.modulation.bias": (6,), "activate.bias": (5,)})
        w = p["conv.weight"].requires_grad_(True)
        mw = p["conv.modulation.weight"].requires_grad_(True)
        x = synth.randn_t(K.SEED, "so_x_" + name, (2, 6, 7, 7)).requires_grad_(True)
        s = synth.randn_t(K.SEED, "so_s_" + name, (2, K.STYLE_DIM)).requires_grad_(True)
        oh = 11 if up else 5
        nz = synth.randn_t(K.SEED, "so_nz_" + name, (2, 1, oh, oh))
        blur = torch.from_numpy(O.make_kernel([1, 2, 1]) * 4) if up else None
        y = O.styled_conv(x, s, w, mw, p["conv.modulation.bias"], nz, torch.tensor([0.3]), p["activate.bias"], upsample=up,
                          blur_kernel=blur)
        assert K.rel_err(K.t2n(y), g["y_" + name]) < 1e-5
        n = synth.randn_t(K.SEED, "so_n_" + name, y.shape)
        gs, = torch.autograd.grad((y * n).sum(), s, create_graph=True)
        gw, gmw, gx = torch.autograd.grad(gs.pow(2).sum(), [w, mw, x])
        assert K.rel_err(K.t2n(gs), g["g_" + name]) < 1e-5
        assert K.rel_err(K.t2n(gw), g["gw_" + name]) < 1e-4
        assert K.rel_err(K.t2n(gmw), g["gmw_" + name]) < 1e-4
        assert K.rel_err(K.t2n(gx), g["gx_" + name]) < 1e-4
