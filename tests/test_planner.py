"""CPU: the host-side pass planner reproduces PyTorch's conv / conv_transpose / data-gradient results when its
passes are evaluated with the numpy restatement of the SpganConvPass contract (oracle/conv_pass_ref.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import conv_pass_ref
from spgan_b200.functional import ConvGeom, plan_passes

GEOMS = [
    ("k3", ConvGeom(3, 3), 7),
    ("k3_pad1", ConvGeom(3, 3, pad=1), 6),
    ("k7", ConvGeom(7, 7), 9),
    ("k1", ConvGeom(1, 1), 5),
    ("k3_s2", ConvGeom(3, 3, stride=2), 8),       # D: blur -> stride-2 3x3 (stylegan2discriminator.py:24-43)
    ("k3_s2_odd", ConvGeom(3, 3, stride=2), 9),
    ("k1_s2", ConvGeom(1, 1, stride=2), 8),       # D skip: blur -> stride-2 1x1
    ("k3_s3", ConvGeom(3, 3, stride=3), 9),       # spherical conv over the gathered taps
    ("convT_crop1", ConvGeom(3, 3, stride=2, transposed=True, crop=1), 5),   # models/ops.py:617-619
    ("convT_crop0", ConvGeom(3, 3, stride=2, transposed=True, crop=0), 4),
]


def _torch_base(x, w, geom):
    if geom.transposed:
        y = F.conv_transpose2d(x, w.transpose(0, 1), stride=geom.stride)
        c = geom.crop
        return y[:, :, c:y.shape[2] - c, c:y.shape[3] - c] if c else y
    return F.conv2d(x, w, stride=geom.stride, padding=geom.pad)


@pytest.mark.parametrize("name,geom,H", GEOMS, ids=[g[0] for g in GEOMS])
def test_forward_and_adjoint_passes(name, geom, H):
    rng = np.random.default_rng(5)
    B, C, O = 2, 3, 2
    x = torch.from_numpy(rng.standard_normal((B, C, H, H))).double().requires_grad_(True)
    w = torch.from_numpy(rng.standard_normal((O, C, geom.kh, geom.kw))).double()
    im = rng.standard_normal((B, C))
    om = rng.standard_normal((B, O))
    y_ref = _torch_base(x * torch.from_numpy(im)[:, :, None, None], w, geom) * torch.from_numpy(om)[:, :, None, None] * 0.7
    oh, ow = geom.out_size(H, H)
    assert (oh, ow) == tuple(y_ref.shape[2:])
    kk = geom.kh * geom.kw
    wf = w.numpy().reshape(-1)
    passes, covers = plan_passes(geom, False, (H, H), (oh, ow))
    y = np.zeros((B, O, oh, ow))
    written = np.zeros((oh, ow), dtype=int)
    for p in passes:
        conv_pass_ref.run_pass(p, y, x.detach().numpy(), wf, C * kk, kk, O, im, om, 0.7)
        for i in range(p["My"]):
            for j in range(p["Mx"]):
                written[i * p["out_stride"] + p["off_y"], j * p["out_stride"] + p["off_x"]] += 1
    assert written.max() == 1, "an output element is written by two passes"
    assert covers == bool((written == 1).all())
    np.testing.assert_allclose(y, y_ref.detach().numpy(), rtol=1e-10, atol=1e-10)

    # adjoint = data gradient: in_mul acts on the gradient's channels (O), out_mul on the input channels (C)
    g = torch.from_numpy(rng.standard_normal((B, O, oh, ow))).double()
    y_plain = _torch_base(x, w, geom)
    gx_ref, = torch.autograd.grad(y_plain, x, g * torch.from_numpy(om)[:, :, None, None])
    gx_ref = gx_ref * torch.from_numpy(im)[:, :, None, None] * 0.7
    passes, covers = plan_passes(geom, True, (oh, ow), (H, H))
    gx = np.zeros((B, C, H, H))
    written = np.zeros((H, H), dtype=int)
    for p in passes:
        conv_pass_ref.run_pass(p, gx, g.numpy(), wf, kk, C * kk, C, om, im, 0.7)
        for i in range(p["My"]):
            for j in range(p["Mx"]):
                written[i * p["out_stride"] + p["off_y"], j * p["out_stride"] + p["off_x"]] += 1
    assert written.max() <= 1
    assert covers == bool((written == 1).all())
    np.testing.assert_allclose(gx, gx_ref.numpy(), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name,geom,H", GEOMS, ids=[g[0] for g in GEOMS])
def test_packed_gemm_formulation_forward_adjoint_wgrad(name, geom, H):
    """The geometry the tcgen05 path is driven with (functional._phase_taps: common lattice, polyphase tap mapping)
    reproduces PyTorch when the packed operands are contracted the way spgan_conv_gemm (flat and im2col addressing) and
    spgan_conv_wgrad_gemm do, for every conv family member, its data gradient and its weight gradient."""
    from spgan_b200.functional import _phase_taps
    rng = np.random.default_rng(7)
    B, C, O = 2, 3, 2
    kk = geom.kh * geom.kw
    x = torch.from_numpy(rng.standard_normal((B, C, H, H))).double().requires_grad_(True)
    w = torch.from_numpy(rng.standard_normal((O, C, geom.kh, geom.kw))).double().requires_grad_(True)
    im = rng.standard_normal((B, C))
    y_ref = _torch_base(x * torch.from_numpy(im)[:, :, None, None], w, geom)
    oh, ow = y_ref.shape[2:]
    wn = w.detach().numpy().reshape(O, C, kk)

    def run(passes, src, mul, wsel_of, out_shape):
        step, pt, pl, Hl, Wl, mapped = _phase_taps(passes)
        if step == 1:
            Hl, Wl = max(Hl, pt + src.shape[2]), max(Wl, pl + src.shape[3])
        pack = conv_pass_ref.pack_lattice(src, None, step, pt, pl, Hl, Wl)
        if mul is not None:
            pack = pack * mul[None, :, None, None, :]
        outs = {}
        for mode in ("flat", "im2col"):
            y = np.zeros(out_shape)
            for p, taps in zip(passes, mapped):
                fn = conv_pass_ref.gemm_flat if mode == "flat" else conv_pass_ref.gemm_im2col
                t3 = [(ph, oy, ox) for ph, oy, ox, _ in taps]
                res = fn(pack, t3, [wsel_of(wi) for _, _, _, wi in taps], p["My"], p["Mx"])  # (B, My, Mx, N)
                for i in range(p["My"]):
                    Y = i * p["out_stride"] + p["off_y"]
                    if not 0 <= Y < out_shape[2]:
                        continue
                    for j in range(p["Mx"]):
                        X = j * p["out_stride"] + p["off_x"]
                        if 0 <= X < out_shape[3]:
                            y[:, :, Y, X] = res[:, i, j, :]
            outs[mode] = y
        return outs

    # forward
    passes, _ = plan_passes(geom, False, (H, H), (oh, ow))
    outs = run(passes, x.detach().numpy(), im, lambda wi: wn[:, :, wi], (B, O, oh, ow))
    for mode, y in outs.items():
        np.testing.assert_allclose(y, y_ref.detach().numpy(), rtol=1e-10, atol=1e-10, err_msg="forward " + mode)
    # data gradient
    g = rng.standard_normal((B, O, oh, ow))
    gx_ref, = torch.autograd.grad(_torch_base(x, w, geom), x, torch.from_numpy(g))
    passes, _ = plan_passes(geom, True, (oh, ow), (H, H))
    outs = run(passes, g, None, lambda wi: wn[:, :, wi].T, (B, C, H, H))
    for mode, y in outs.items():
        np.testing.assert_allclose(y, gx_ref.numpy(), rtol=1e-10, atol=1e-10, err_msg="adjoint " + mode)
    # weight gradient (as functional.conv_wgrad drives it)
    gw_ref, = torch.autograd.grad(_torch_base(x, w, geom), w, torch.from_numpy(g))
    passes, _ = plan_passes(geom, False, (H, H), (oh, ow))
    s_in, pt, pl, Hl, Wl, mapped = _phase_taps(passes)
    s_out = passes[0]["out_stride"]
    if s_in == 1:
        Hl, Wl = max(Hl, pt + H), max(Wl, pl + H)
    gpack = conv_pass_ref.pack_lattice(g, None, s_out, 0, 0, Hl, Wl)
    xpack = conv_pass_ref.pack_lattice(x.detach().numpy(), None, s_in, pt, pl, Hl, Wl)
    dw = np.zeros((O, C, kk))
    for p, taps in zip(passes, mapped):
        res = conv_pass_ref.wgrad_lattice(gpack, p["off_y"] * s_out + p["off_x"], xpack, [(ph, oy, ox) for ph, oy, ox, _ in taps])
        for (_, _, _, wi), m in zip(taps, res):
            dw[:, :, wi] += m
    np.testing.assert_allclose(dw.reshape(gw_ref.shape), gw_ref.numpy(), rtol=1e-10, atol=1e-10, err_msg="wgrad")
