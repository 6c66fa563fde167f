"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures, same seeded inputs.

Tolerances: bit-exact for sampling indices; <= 1e-5 relative-to-peak for the HBM-bound fp32 kernels and the exact
fp32 SIMT conv; <= 1e-4 for the bf16x3 tcgen05 conv (fp32-equivalent, north_star bound 1e-3); <= 3e-2 for the
plain-bf16 tcgen05 mode (the stated looser bound for bf16 paths); <= 4e-4 for the 2-MMA fp16 split (mode 3: fp16 hi+lo
activations x fp16 weights — one operand carries 11 significant bits, per-op error ~2^-12; forward only, gradients fall
back to bf16x3)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases as K
import spgan_oracle as O
import synth

pytestmark = pytest.mark.gpu

TOL = {0: 1e-5, 1: 1e-4, 2: 3e-2, 3: 4e-4}
MODES_16 = [1, 2, 3]


@pytest.fixture(scope="module")
def dev():
    import spgan_b200.lib as lib
    torch.cuda.set_device(0)
    lib.require_device()
    return torch.device("cuda:0")


def SF():
    import spgan_b200.functional as f
    return f


# ---------------------------------------------------------------------------------------------- K1
def test_bias_act_golden_forward_backward(dev):
    g = K.load("bias_act.npz")
    for name, shape in (("4d", (2, 5, 7, 3)), ("2d", (3, 6))):
        x = synth.randn_t(K.SEED, "ba_x_" + name, shape).to(dev).requires_grad_(True)
        b = synth.randn_t(K.SEED, "ba_b_" + name, (shape[1],)).to(dev).requires_grad_(True)
        go = synth.randn_t(K.SEED, "ba_go_" + name, shape).to(dev)
        y = SF().fused_leaky_relu(x, b)
        gx, gb = torch.autograd.grad(y, [x, b], go)
        assert K.rel_err(K.t2n(y), g["y_" + name]) < 1e-6
        assert K.rel_err(K.t2n(gx), g["gx_" + name]) < 1e-6
        assert K.rel_err(K.t2n(gb), g["gb_" + name]) < 1e-5


@pytest.mark.parametrize("shape", [(3, 7, 5, 9), (2, 8, 16, 16), (1, 3, 1, 1), (4, 512), (2, 6, 33, 35)])
@pytest.mark.parametrize("code", [(3, 0), (3, 1), (1, 0), (1, 1), (3, 2)])
def test_bias_act_all_codes_vs_oracle(dev, shape, code):
    act, grad = code
    x = synth.randn(1, "bx", shape)
    b = synth.randn(1, "bb", (shape[1],))
    r = synth.randn(1, "br", shape)
    want = O.fused_bias_act(x, b, r, act, grad, 0.2, 2 ** 0.5)
    got = SF().bias_act(torch.from_numpy(x).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(r).to(dev), act, grad, 0.2, 2 ** 0.5)
    assert np.array_equal(K.t2n(got), want) or K.rel_err(K.t2n(got), want) < 1e-7
    got = SF().bias_act(torch.from_numpy(x).to(dev), None, torch.from_numpy(r).to(dev), act, grad, 0.1, 1.0)
    assert K.rel_err(K.t2n(got), O.fused_bias_act(x, None, r, act, grad, 0.1, 1.0)) < 1e-7


def test_bias_act_empty_and_double_backward(dev):
    e = SF().bias_act(torch.zeros(0, 4, 3, 3, device=dev), torch.zeros(4, device=dev))
    assert e.shape == (0, 4, 3, 3)
    x = torch.randn(2, 3, 5, 5, device=dev, requires_grad=True)
    b = torch.randn(3, device=dev, requires_grad=True)
    y = SF().fused_leaky_relu(x, b)
    go = torch.randn_like(y).requires_grad_(True)
    gx, = torch.autograd.grad(y, x, go, create_graph=True)
    # d(gx)/d(go) applied to v = K1(grad=1) with the same gate
    v = torch.randn_like(gx)
    ggo, = torch.autograd.grad(gx, go, v)
    want = O.fused_bias_act(K.t2n(v), None, K.t2n(y), 3, 1, 0.2, 2 ** 0.5)
    assert K.rel_err(K.t2n(ggo), want) < 1e-6


@pytest.mark.parametrize("shape", [(3, 6, 19, 19), (4, 6, 19, 19), (2, 4, 8, 8), (4, 3, 5, 5), (8, 5, 101, 101)])
def test_noise_bias_act_vs_oracle(dev, shape):
    """Scalar plane kernel (element count not a multiple of 4) and the 128-bit flat kernel, whose float4s straddle plane,
    channel and sample boundaries on odd planes."""
    B, C, H, W = shape
    x = synth.randn_t(2, "nx", (B, C, H, W))
    nz = synth.randn_t(2, "nn", (B, 1, H, W))
    b = synth.randn_t(2, "nb", (C,))
    nw = torch.tensor([0.37])
    want = O.fused_leaky_relu(K.t2n(x + nw * nz), K.t2n(b))
    got = SF().noise_bias_act(x.to(dev), nz.to(dev), nw.to(dev), b.to(dev))
    assert K.rel_err(K.t2n(got), want) < 1e-6


# ---------------------------------------------------------------------------------------------- K2/K3
@pytest.mark.parametrize("case", K.UPFIRDN_CASES, ids=[c[0] for c in K.UPFIRDN_CASES])
def test_upfirdn2d_golden(dev, case):
    name, shape, taps, gain, up, down, pad = case
    g = K.load("upfirdn2d.npz")
    k = torch.from_numpy(O.make_kernel(taps) * np.float32(gain)).to(dev)
    x = synth.randn_t(K.SEED, "ufd_x_" + name, shape).to(dev).requires_grad_(True)
    y = SF().upfirdn2d(x, k, up=up, down=down, pad=pad)
    assert y.shape == g["y_" + name].shape
    go = synth.randn_t(K.SEED, "ufd_go_" + name, y.shape).to(dev)
    gx, = torch.autograd.grad(y, x, go)
    assert K.rel_err(K.t2n(y), g["y_" + name]) < 1e-6
    assert K.rel_err(K.t2n(gx), g["gx_" + name]) < 1e-6


def test_upfirdn2d_large_and_double_backward(dev):
    k = torch.from_numpy(O.make_kernel([1, 3, 3, 1])).to(dev)
    x = synth.randn_t(3, "ufl", (2, 7, 101, 101))
    y = SF().upfirdn2d(x.to(dev), k, pad=(2, 2))
    assert K.rel_err(K.t2n(y), O.upfirdn2d(K.t2n(x), K.t2n(k), pad=(2, 2, 2, 2))) < 1e-6
    xg = x.to(dev).requires_grad_(True)
    y = SF().upfirdn2d(xg, k, pad=(1, 1))
    go = torch.randn_like(y).requires_grad_(True)
    gx, = torch.autograd.grad(y, xg, go, create_graph=True)
    v = torch.randn_like(gx)
    ggo, = torch.autograd.grad(gx, go, v)  # adjoint of the adjoint = the forward op
    assert K.rel_err(K.t2n(ggo), O.upfirdn2d(K.t2n(v), K.t2n(k), pad=(1, 1, 1, 1))) < 1e-6
    # a delta kernel with no padding is the identity
    d = torch.ones(1, 1, device=dev)
    assert torch.equal(SF().upfirdn2d(x.to(dev), d), x.to(dev))


def _legacy(flag):
    import os
    if flag:
        os.environ["SPGAN_LEGACY_HBM_KERNELS"] = "1"
    else:
        os.environ.pop("SPGAN_LEGACY_HBM_KERNELS", None)


@pytest.mark.parametrize("shape,taps,pad", [
    ((2, 3, 13, 13), [1, 2, 1], (0, 0)),        # whole planes, tensor end not a multiple of 16 bytes (scalar tail)
    ((3, 5, 19, 19), [1, 3, 3, 1], (2, 2)),     # several planes per item, plane count not a multiple of the group
    ((5, 67, 21, 23), [1, 3, 3, 1], (1, 1)),    # more items than resident CTAs, ring of two stages wraps
    ((1, 2, 150, 131), [1, 3, 3, 1], (2, 1)),   # bands of one plane, asymmetric padding
    ((1, 3, 384, 384), [1, 2, 1], (0, 0)),      # configs[4] largest resolution: bands
    ((1, 2, 105, 105), [1, 2, 1], (2, 2)),      # gradient of the generator's blur
    ((2, 2, 9, 40), [1, 1], (1, 0)),            # 2x2 taps
    ((600, 1, 5, 7), [1, 3, 3, 1], (2, 2)),     # many tiny planes
    ((1, 2, 40, 400), [1, 3, 3, 1], (2, 2)),    # wider than a thread block: two balanced column passes
    ((1, 1, 30, 700), [1, 2, 1], (1, 1)),       # three column passes
], ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) and len(v) == 4 else None)
def test_upfirdn2d_streamed_kernel_vs_torch_and_legacy(dev, shape, taps, pad):
    """The bulk-copy staged FIR (csrc/upfirdn2d.cu: fir_stream_kernel) against an fp64 torch correlation and against the
    band / tiled kernels it replaces."""
    k = torch.from_numpy(O.make_kernel(taps)).to(dev)
    x = synth.randn_t(11, "fs_%s" % (shape,), shape).to(dev)
    _legacy(False)
    y = SF().upfirdn2d(x, k, pad=pad)
    _legacy(True)
    try:
        y_old = SF().upfirdn2d(x, k, pad=pad)
    finally:
        _legacy(False)
    B, C, H, W = shape
    xp = F.pad(x.double().reshape(B * C, 1, H, W), (pad[0], pad[1], pad[0], pad[1]))
    want = F.conv2d(xp, torch.flip(k.double(), [0, 1])[None, None]).reshape(B, C, *y.shape[2:])
    assert y.shape == want.shape
    assert K.rel_err(K.t2n(y), want.cpu().numpy()) < 1e-6
    assert K.rel_err(K.t2n(y), K.t2n(y_old)) < 1e-6
    # a view whose data pointer is not 16-byte aligned takes the fallback and still agrees
    flat = torch.zeros(x.numel() + 1, device=dev)
    flat[1:] = x.reshape(-1)
    y_off = SF().upfirdn2d(flat[1:].view(shape), k, pad=pad)
    assert K.rel_err(K.t2n(y_off), want.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("shape,taps,up,down,pad", [
    ((2, 3, 53, 53), [1, 3, 3, 1], 2, 1, (2, 1)),     # Upsample of ToRGB's skip (models/ops.py:32-55)
    ((2, 3, 17, 29), [1, 2, 1], 2, 1, (1, 0)),        # 3 taps, odd tap origin
    ((1, 2, 150, 131), [1, 3, 3, 1], 2, 1, (2, 1)),   # bands
    ((3, 4, 101, 101), [1, 3, 3, 1], 1, 2, (1, 1)),   # Downsample (models/ops.py:58-79)
    ((1, 2, 384, 384), [1, 3, 3, 1], 1, 2, (1, 1)),   # bands
    ((2, 2, 106, 106), [1, 3, 3, 1], 1, 2, (1, 2)),   # gradient of the up-sampling above
    ((5, 40, 12, 12), [1, 3, 3, 1], 2, 1, (2, 1)),    # several planes per item
], ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) and len(v) == 4 else None)
def test_upfirdn2d_streamed_up_down_vs_oracle_and_legacy(dev, shape, taps, up, down, pad):
    """Up / down by two through the bulk-copy staged kernel (fir_stream_kernel<4, 2, 1> / <4, 1, 2>) against the oracle
    and the polyphase kernel it replaces, forward and gradient."""
    k = torch.from_numpy(O.make_kernel(taps) * np.float32(up * up)).to(dev)
    x = synth.randn_t(12, "fsud_%s" % (shape,), shape).to(dev).requires_grad_(True)
    _legacy(False)
    y = SF().upfirdn2d(x, k, up=up, down=down, pad=pad)
    go = synth.randn_t(12, "fsud_go_%s" % (shape,), y.shape).to(dev)
    gx, = torch.autograd.grad(y, x, go)
    _legacy(True)
    try:
        y_old = SF().upfirdn2d(x, k, up=up, down=down, pad=pad)
        gx_old, = torch.autograd.grad(y_old, x, go)
    finally:
        _legacy(False)
    want = O.upfirdn2d(K.t2n(x), K.t2n(k), up=(up, up), down=(down, down), pad=(pad[0], pad[1], pad[0], pad[1]))
    assert y.shape == want.shape
    assert K.rel_err(K.t2n(y), want) < 1e-6
    assert K.rel_err(K.t2n(y), K.t2n(y_old)) < 1e-6
    assert K.rel_err(K.t2n(gx), K.t2n(gx_old)) < 1e-6


# ---------------------------------------------------------------------------------------------- gather
def test_gather_indices_bit_exact(dev):
    g = K.load("grids.npz")
    for name, c in K.load_json("grid_cases.json").items():
        h = c["h"]
        grid = torch.from_numpy(g[name]).to(dev)
        x0, y0, wx, wy = SF().sphere_gather_indices(grid, h, h)
        ox0, oy0, owx, owy = O.gather_indices(g[name], h, h)
        assert np.array_equal(K.t2n(x0), ox0) and np.array_equal(K.t2n(y0), oy0), name
        assert np.array_equal(K.t2n(x0).astype(np.int16), g[name + "_x0"]), name
        assert np.array_equal(K.t2n(wx), owx) and np.array_equal(K.t2n(wy), owy), name


def test_grid_assembly_on_device_bit_exact(dev):
    """Per-sample training grids assembled on the device from the factor tables equal the host-built grids bit for
    bit, across table growth (more than 16 distinct windows) and repeated windows."""
    from spgan_b200 import grids
    rng = np.random.RandomState(5)
    cache = grids.GridCache()
    for h in (35, 17, 53):
        for B in (1, 8, 40):
            cps = [K.train_cp(int(rng.randint(0, 10)), int(rng.randint(0, 140)), 35) for _ in range(B)]
            got = cache.batch(h, h, cps, B, dev)
            want = np.concatenate([grids.sampling_grid(h, h, cp) for cp in cps], 0)
            assert got.shape == want.shape
            assert np.array_equal(K.t2n(got).view(np.uint32), want.view(np.uint32)), (h, B)


def test_gather_golden_forward_and_surrogate_backward(dev):
    g = K.load("gather.npz")
    for name, (B, C, h) in (("train", (2, 5, 17)), ("border", (1, 3, 11))):
        z = synth.randn_t(K.SEED, "gather_z_" + name, (B, C, h, h)).to(dev).requires_grad_(True)
        grid = torch.from_numpy(g["grid_" + name]).to(dev)
        y = SF().sphere_gather(z, grid)
        go = synth.randn_t(K.SEED, "gather_go_" + name, y.shape).to(dev)
        gz, = torch.autograd.grad(y, z, go)
        assert K.rel_err(K.t2n(y), g["y_" + name]) < 2e-6
        assert K.rel_err(K.t2n(gz), g["gz_" + name]) < 1e-6


def test_gather_shared_grid_and_channel_chunks(dev):
    cp = K.test_cp(2, 7, 27)
    grid = O.gen_sampling_grid(23, 23, cp)
    z = synth.randn(4, "gz", (3, 37, 23, 23))
    want = O.grid_sample_border(z, np.repeat(grid, 3, 0))
    got = SF().sphere_gather_raw(torch.from_numpy(z).to(dev), torch.from_numpy(grid).to(dev))
    assert K.rel_err(K.t2n(got), want) < 2e-6


@pytest.mark.parametrize("B,C,h,w,shared", [(2, 5, 7, 9, False), (3, 37, 23, 23, True), (2, 3, 17, 17, False),
                                            (9, 20, 35, 35, True), (1, 2, 130, 70, True), (2, 1, 11, 11, False)])
def test_gather_streamed_kernel_vs_oracle_and_legacy(dev, B, C, h, w, shared):
    """The bulk-copy staged gather (csrc/sphere_gather.cu: sphere_gather_stream_kernel) against the oracle and against the
    L1-gather kernel it replaces (same corner arithmetic; the blend may contract its first product differently, so the
    two agree to an ulp, not bit for bit); also the coordinate-encoding variant and the strided output of the flat-concat
    buffer."""
    cp = K.test_cp(2, 7, 27)
    g1 = O.gen_sampling_grid(h, w, cp)
    grid = g1 if shared else np.concatenate([g1 + np.float32(0.003 * i) for i in range(B)], 0)
    z = synth.randn(5, "gsz%d" % C, (B, C, h, w))
    want = O.grid_sample_border(z, np.repeat(g1, B, 0) if shared else grid)
    zt, gt = torch.from_numpy(z).to(dev), torch.from_numpy(grid).to(dev)
    _legacy(False)
    got = SF().sphere_gather_raw(zt, gt)
    buf = torch.zeros(B, C + 4, 3 * h, 3 * w, device=dev)
    SF().sphere_gather_raw(zt, gt, out=buf, out_bstride=C + 4, out_coff=2)
    enc = SF().sphere_gather_raw(zt[:, :3].contiguous(), gt, encode=True) if C >= 3 else None
    _legacy(True)
    try:
        old = SF().sphere_gather_raw(zt, gt)
        enc_old = SF().sphere_gather_raw(zt[:, :3].contiguous(), gt, encode=True) if C >= 3 else None
    finally:
        _legacy(False)
    assert K.rel_err(K.t2n(got), want) < 2e-6
    assert K.rel_err(K.t2n(got), K.t2n(old)) < 5e-7
    assert torch.equal(buf[:, 2:2 + C], got) and float(buf[:, :2].abs().max()) == 0 and float(buf[:, 2 + C:].abs().max()) == 0
    if enc is not None:
        assert K.rel_err(K.t2n(enc), K.t2n(enc_old)) < 1e-6


# ---------------------------------------------------------------------------------------------- conv passes
def _ref_conv(x, w, geom, adjoint, out_hw, in_mul, out_mul, out_scale):
    x = x.double()
    w = w.double()
    if in_mul is not None:
        x = x * in_mul.double()[:, :, None, None]
    if not adjoint:
        if geom.transposed:
            y = F.conv_transpose2d(x, w.transpose(0, 1), stride=geom.stride)
            c = geom.crop
            y = y[:, :, c:y.shape[2] - c, c:y.shape[3] - c] if c else y
        else:
            y = F.conv2d(x, w, stride=geom.stride, padding=geom.pad)
    else:
        B = x.shape[0]
        probe = torch.zeros(B, w.shape[1], out_hw[0], out_hw[1], dtype=torch.float64, requires_grad=True)
        if geom.transposed:
            base = F.conv_transpose2d(probe, w.transpose(0, 1), stride=geom.stride)
            c = geom.crop
            base = base[:, :, c:base.shape[2] - c, c:base.shape[3] - c] if c else base
        else:
            base = F.conv2d(probe, w, stride=geom.stride, padding=geom.pad)
        y, = torch.autograd.grad(base, probe, x)
    if out_mul is not None:
        y = y * out_mul.double()[:, :, None, None]
    return y * out_scale


def _geoms():
    from spgan_b200.functional import ConvGeom
    return [
        ("k3", ConvGeom(3, 3), 19),
        ("k3_pad1", ConvGeom(3, 3, pad=1), 12),
        ("k7", ConvGeom(7, 7), 17),
        ("k1", ConvGeom(1, 1), 9),
        ("k3_s2", ConvGeom(3, 3, stride=2), 13),
        ("k1_s2", ConvGeom(1, 1, stride=2), 12),
        ("k3_s3", ConvGeom(3, 3, stride=3), 15),
        ("convT_crop1", ConvGeom(3, 3, stride=2, transposed=True, crop=1), 11),
        ("convT_crop0", ConvGeom(3, 3, stride=2, transposed=True, crop=0), 6),
    ]


@pytest.mark.parametrize("gi", range(9))
def test_conv_simt_forward_adjoint_wgrad_vs_torch(dev, gi):
    _conv_case(dev, gi, 0)


@pytest.mark.parametrize("gi", range(9))
@pytest.mark.parametrize("precision", MODES_16)
def test_conv_tcgen05_forward_adjoint_vs_torch(dev, gi, precision):
    _conv_case(dev, gi, precision)


def _conv_case(dev, gi, precision):
    name, geom, H = _geoms()[gi]
    B, C, Oc = 3, 70, 40  # C pads to 128 channels on the tensor path, Cout is not a multiple of 16
    x = synth.randn_t(7, "cx" + name, (B, C, H, H))
    w = synth.randn_t(7, "cw" + name, (Oc, C, geom.kh, geom.kw), 0.2)
    im = synth.randn_t(7, "cim" + name, (B, C), 0.3, 1.0)
    om = synth.randn_t(7, "com" + name, (B, Oc), 0.3, 1.0)
    want = _ref_conv(x, w, geom, False, None, im, om, 0.37)
    got = SF().conv_apply(x.to(dev), w.to(dev), geom, False, None, im.to(dev), om.to(dev), 0.37, precision=precision)
    assert got.shape == want.shape
    assert K.rel_err(K.t2n(got), want.numpy()) < TOL[precision], "forward"
    oh, ow = want.shape[2:]
    g = synth.randn_t(7, "cg" + name, (B, Oc, oh, ow))
    want = _ref_conv(g, w, geom, True, (H, H), om, im, 0.37)
    got = SF().conv_apply(g.to(dev), w.to(dev), geom, True, (H, H), om.to(dev), im.to(dev), 0.37, precision=precision)
    assert K.rel_err(K.t2n(got), want.numpy()) < TOL[precision], "adjoint"
    wd = w.double().requires_grad_(True)
    base = _ref_conv(x, wd, geom, False, None, im, om, 0.37)
    want_w, = torch.autograd.grad(base, wd, g.double())
    got_w = SF().conv_wgrad(g.to(dev), x.to(dev), tuple(w.shape), geom, im.to(dev), om.to(dev), 0.37, precision=precision)
    assert K.rel_err(K.t2n(got_w), want_w.numpy()) < (1e-5 if precision == 0 else TOL[precision]), "wgrad"


@pytest.mark.parametrize("precision", MODES_16)
def test_conv_tcgen05_tiles_and_epilogue(dev, precision):
    """Two N tiles with a ragged second tile, several M tiles, K padding, and every epilogue term at once."""
    from spgan_b200.functional import ConvGeom
    B, C, Oc, H = 5, 130, 300, 23
    geom = ConvGeom(3, 3)
    x = synth.randn_t(8, "tx", (B, C, H, H))
    w = synth.randn_t(8, "tw", (Oc, C, 3, 3), 0.1)
    im = synth.randn_t(8, "tim", (B, C), 0.3, 1.0)
    om = synth.randn_t(8, "tom", (B, Oc), 0.3, 1.0)
    nz = synth.randn_t(8, "tnz", (B, 1, H - 2, H - 2))
    nw = torch.tensor([0.41])
    bias = synth.randn_t(8, "tb", (Oc,))
    res = synth.randn_t(8, "tr", (B, Oc, H - 2, H - 2))
    y = _ref_conv(x, w, geom, False, None, im, om, 0.05) + (nw * nz).double() + bias.double()[None, :, None, None]
    y = F.leaky_relu(y, 0.2) * 2 ** 0.5 + res.double()
    got = SF().conv_apply(x.to(dev), w.to(dev), geom, False, None, im.to(dev), om.to(dev), 0.05, nz.to(dev), nw.to(dev),
                          bias.to(dev), (0.2, 2 ** 0.5), res.to(dev), precision)
    assert K.rel_err(K.t2n(got), y.numpy()) < TOL[precision]
    import spgan_b200.lib as lib
    assert lib.load().spgan_gemm_launch_count() > 0


def test_conv_tcgen05_matches_simt_at_layer_size_and_is_linear(dev):
    """Full-size property check (TS layer 7 shape, batch 4): tcgen05 bf16x3 == exact SIMT, and the op is linear in x."""
    from spgan_b200.functional import ConvGeom
    B, C, Oc, H = 4, 512, 512, 103
    geom = ConvGeom(3, 3)
    gen = torch.Generator(device="cpu").manual_seed(5)
    x1 = torch.randn(B, C, H, H, generator=gen).to(dev)
    x2 = torch.randn(B, C, H, H, generator=gen).to(dev)
    w = (torch.randn(Oc, C, 3, 3, generator=gen) * 0.05).to(dev)
    im = (torch.randn(B, C, generator=gen) * 0.3 + 1).to(dev)
    om = (torch.randn(B, Oc, generator=gen) * 0.3 + 1).to(dev)
    y1 = SF().conv_apply(x1, w, geom, in_mul=im, out_mul=om, out_scale=0.02, precision=1)
    y2 = SF().conv_apply(x2, w, geom, in_mul=im, out_mul=om, out_scale=0.02, precision=1)
    y12 = SF().conv_apply(2.5 * x1 + x2, w, geom, in_mul=im, out_mul=om, out_scale=0.02, precision=1)
    assert K.rel_err(K.t2n(y12), K.t2n(2.5 * y1 + y2)) < 1e-4
    ys = SF().conv_apply(x1[:1], w, geom, in_mul=im[:1], out_mul=om[:1], out_scale=0.02, precision=0)
    assert K.rel_err(K.t2n(y1[:1]), K.t2n(ys)) < 1e-4


def test_conv_tcgen05_partial_last_k_block(dev):
    """K per tap = 80 (a multiple of 16, not of 64): the last K block of every tap is partial (TMA zero fill, the MMAs of
    the empty K steps are skipped)."""
    from spgan_b200.functional import ConvGeom
    name, geom, H = "k3", ConvGeom(3, 3), 19
    B, C, Oc = 3, 70, 40
    x = synth.randn_t(7, "cx" + name, (B, C, H, H))
    w = synth.randn_t(7, "cw" + name, (Oc, C, 3, 3), 0.2)
    im = synth.randn_t(7, "cim" + name, (B, C), 0.3, 1.0)
    om = synth.randn_t(7, "com" + name, (B, Oc), 0.3, 1.0)
    want = _ref_conv(x, w, geom, False, None, im, om, 0.37)
    got = SF().conv_apply(x.to(dev), w.to(dev), geom, False, None, im.to(dev), om.to(dev), 0.37, precision=1, k_round=16)
    assert K.rel_err(K.t2n(got), want.numpy()) < TOL[1]


def test_wgrad_tcgen05_layer_size_split_k_vs_simt(dev):
    """Weight gradient at a real layer size (512 -> 259-like ragged Cin, several K chunks, two N tiles with a 3-column
    second tile) against the exact-fp32 SIMT kernel; the result is deterministic (fixed K-chunk summation order)."""
    from spgan_b200.functional import ConvGeom
    B, C, Oc, H = 4, 259, 256, 35
    geom = ConvGeom(7, 7)
    x = synth.randn_t(11, "wgx", (B, C, H, H)).to(dev)
    g = synth.randn_t(11, "wgg", (B, Oc, H - 6, H - 6)).to(dev)
    im = synth.randn_t(11, "wgim", (B, C), 0.3, 1.0).to(dev)
    om = synth.randn_t(11, "wgom", (B, Oc), 0.3, 1.0).to(dev)
    ref = SF().conv_wgrad(g, x, (Oc, C, 7, 7), geom, im, om, 0.02, precision=0)
    got = SF().conv_wgrad(g, x, (Oc, C, 7, 7), geom, im, om, 0.02, precision=1)
    again = SF().conv_wgrad(g, x, (Oc, C, 7, 7), geom, im, om, 0.02, precision=1)
    assert K.rel_err(K.t2n(got), K.t2n(ref)) < TOL[1]
    assert torch.equal(got, again)


@pytest.mark.parametrize("precision", [0, 1])
def test_polyphase_transposed_conv_and_fused_upblur(dev, precision):
    """Upsampling StyledConv tail: polyphase convT + fused interleave/FIR/noise/bias/act == convT -> Blur -> noise -> act."""
    from spgan_b200.functional import ConvGeom
    geom = ConvGeom(3, 3, stride=2, transposed=True, crop=1)
    B, C, Oc, H = 3, 70, 40, 11
    x = synth.randn_t(11, "px", (B, C, H, H))
    w = synth.randn_t(11, "pw", (Oc, C, 3, 3), 0.2)
    im = synth.randn_t(11, "pim", (B, C), 0.3, 1.0)
    om = synth.randn_t(11, "pom", (B, Oc), 0.3, 1.0)
    nz = synth.randn_t(11, "pnz", (B, 1, 2 * H - 3, 2 * H - 3))
    nw = torch.tensor([0.3])
    bias = synth.randn_t(11, "pb", (Oc,))
    k = torch.from_numpy(O.make_kernel([1, 2, 1]) * 4)
    z = _ref_conv(x, w, geom, False, None, im, om, 0.11).float()
    want = O.fused_leaky_relu(O.upfirdn2d(z.numpy(), k.numpy()) + (nw * nz).numpy(), bias.numpy())
    pp = SF().conv_apply(x.to(dev), w.to(dev), geom, in_mul=im.to(dev), out_mul=om.to(dev), out_scale=0.11,
                         precision=precision, polyphase=True)
    assert pp.shape == (B, Oc, 4, H, H)
    zz = torch.zeros(B, Oc, 2 * H - 1, 2 * H - 1, device=dev)
    for a in range(2):
        for b in range(2):
            sub = zz[:, :, a::2, b::2]
            sub.copy_(pp[:, :, a * 2 + b, :sub.shape[2], :sub.shape[3]])
    assert K.rel_err(K.t2n(zz), z.numpy()) < TOL[precision]
    got = SF().upblur_act(pp, k.to(dev), (2 * H - 1, 2 * H - 1), nz.to(dev), nw.to(dev), bias.to(dev))
    assert K.rel_err(K.t2n(got), want) < TOL[precision]


@pytest.mark.parametrize("H,Hq_extra,with_noise", [(53, 0, True), (28, 1, True), (6, 0, False), (2, 0, True), (100, 0, True)])
def test_upblur_act_bands_and_edges_vs_composition(dev, H, Hq_extra, with_noise):
    """Fused interleave + FIR + noise + bias + act against interleave -> upfirdn2d -> noise_bias_act, at sizes with
    several bands per plane, odd and even heights, padded polyphase planes and the single-row edge case."""
    B, C = 2, 5
    zh = 2 * H - 1
    Hq = H + Hq_extra
    pp = synth.randn_t(13, "ubpp%d" % H, (B, C, 4, Hq, Hq)).to(dev)
    k = torch.from_numpy(O.make_kernel([1, 2, 1]) * 4).to(dev)
    nz = synth.randn_t(13, "ubnz%d" % H, (B, 1, zh - 2, zh - 2)).to(dev) if with_noise else None
    nw = torch.tensor([0.7], device=dev) if with_noise else None
    bias = synth.randn_t(13, "ubb%d" % H, (C,)).to(dev)
    zz = torch.zeros(B, C, zh, zh, device=dev)
    for a in range(2):
        for b in range(2):
            sub = zz[:, :, a::2, b::2]
            sub.copy_(pp[:, :, a * 2 + b, :sub.shape[2], :sub.shape[3]])
    want = SF().noise_bias_act(SF().upfirdn2d(zz, k, pad=(0, 0)), nz, nw, bias)
    got = SF().upblur_act(pp, k, (zh, zh), nz, nw, bias)
    assert got.shape == want.shape
    assert K.rel_err(K.t2n(got), K.t2n(want)) < 1e-6


@pytest.mark.parametrize("precision,next_precision", [(1, 1), (3, 3), (1, 3), (3, 1)])
def test_chain_links_vs_module_path(dev, precision, next_precision):
    """Channels-last chain (csrc/chain.cu + the sinks of spgan_conv_gemm_ex) against the NCHW composition it replaces:
    packed input -> 4 parity GEMMs (NHWC planes) -> upblur_pack == pack(upblur_act(polyphase convT)); conv3 with the
    packed sink == pack(conv), its ToRGB partial sums + rgb_tail == the 1x1 modulated conv + bias + skip."""
    f = SF()
    from spgan_b200.functional import ConvGeom
    B, C, Oc, H = 3, 64, 128, 9
    x = synth.randn_t(21, "chx", (B, C, H, H)).to(dev)
    w_up = synth.randn_t(21, "chwu", (Oc, C, 3, 3), 0.2).to(dev)
    w_cv = synth.randn_t(21, "chwc", (Oc, Oc, 3, 3), 0.1).to(dev)
    w_rgb = synth.randn_t(21, "chwr", (3, Oc, 1, 1), 0.3).to(dev)
    s_up = synth.randn_t(21, "chsu", (B, C), 0.3, 1.0).to(dev)
    d_up = synth.randn_t(21, "chdu", (B, Oc), 0.3, 1.0).to(dev)
    s_cv = synth.randn_t(21, "chsc", (B, Oc), 0.3, 1.0).to(dev)
    d_cv = synth.randn_t(21, "chdc", (B, Oc), 0.3, 1.0).to(dev)
    s_nx = synth.randn_t(21, "chsn", (B, Oc), 0.3, 1.0).to(dev)
    s_rgb = synth.randn_t(21, "chsr", (B, Oc), 0.3, 1.0).to(dev)
    zh = 2 * H - 1
    nz1 = synth.randn_t(21, "chn1", (B, 1, zh - 2, zh - 2)).to(dev)
    nz2 = synth.randn_t(21, "chn2", (B, 1, zh - 4, zh - 4)).to(dev)
    nw = torch.tensor([0.3], device=dev)
    b1 = synth.randn_t(21, "chb1", (Oc,)).to(dev)
    b2 = synth.randn_t(21, "chb2", (Oc,)).to(dev)
    b_rgb = synth.randn_t(21, "chbr", (3,)).to(dev)
    skip = synth.randn_t(21, "chsk", (B, 3, zh - 4, zh - 4)).to(dev)
    k = torch.from_numpy(O.make_kernel([1, 2, 1]) * 4).to(dev)
    up = ConvGeom(3, 3, stride=2, transposed=True, crop=1)
    # --- module path
    pp_ref = f.conv_apply(x, w_up, up, in_mul=s_up, out_mul=d_up, out_scale=0.11, precision=precision, polyphase=True)
    h1 = f.upblur_act(pp_ref, k, (zh, zh), nz1, nw, b1)
    h2 = f.conv_apply(h1, w_cv, ConvGeom(3, 3), in_mul=s_cv, out_mul=d_cv, out_scale=0.07, noise=nz2, noise_w=nw, bias=b2,
                      act=(0.2, 2 ** 0.5), precision=next_precision)
    rgb_ref = f.conv_apply(h2, w_rgb, ConvGeom(1, 1), in_mul=s_rgb, out_scale=0.21, bias=b_rgb, residual=skip, precision=0)
    # --- chain
    a = f.chain_pack_input(x, s_up, precision)
    pp, zhw = f.chain_upconv(a, B, H, H, w_up, d_up, 0.11, precision)
    assert zhw == (zh, zh) and pp.shape == (B, 4, H, H, Oc)
    got_pp = pp.permute(0, 4, 1, 2, 3)
    mask = torch.ones_like(pp_ref, dtype=torch.bool)
    mask[:, :, 1, :, H - 1:] = False   # plane (0,1): columns X = 2j+1 < zw  ->  j < H-1
    mask[:, :, 2, H - 1:, :] = False
    mask[:, :, 3, H - 1:, :] = False
    mask[:, :, 3, :, H - 1:] = False
    assert torch.equal(torch.where(mask, got_pp, torch.zeros_like(got_pp)), torch.where(mask, pp_ref, torch.zeros_like(pp_ref)))
    a1, (oh, ow) = f.chain_upblur_pack(pp, zhw, k, nz1, nw, b1, s_cv, next_precision)
    assert (oh, ow) == (zh - 2, zh - 2)
    want_a1 = f.chain_pack_input(h1, s_cv, next_precision)
    tol16 = 2e-5  # a 1-ulp fp32 difference of the FIR sum can move the lo plane by one 2^-17 step
    assert K.rel_err(K.t2n(f.packed_to_float(a1, f._fmt(next_precision))), K.t2n(f.packed_to_float(want_a1, f._fmt(next_precision)))) < tol16
    rgb_w = (w_rgb.reshape(1, 3, Oc) * s_rgb.unsqueeze(1) * 0.21).contiguous()
    packed, rgb, y, (oh2, ow2) = f.chain_conv3(a1, B, oh, ow, w_cv, d_cv, 0.07, nz2, nw, b2, (0.2, 2 ** 0.5), next_precision,
                                              next_mul=s_nx, next_precision=precision, rgb_w=rgb_w, want_nchw=True)
    assert (oh2, ow2) == (zh - 4, zh - 4)
    assert K.rel_err(K.t2n(y), K.t2n(h2)) < 2e-5
    want_packed = f.chain_pack_input(y, s_nx, precision)
    assert torch.equal(packed.view(torch.int16), want_packed.view(torch.int16))
    got_rgb = f.rgb_tail(rgb[0], rgb[1], b_rgb, skip, B, oh2, ow2)
    assert K.rel_err(K.t2n(got_rgb), K.t2n(rgb_ref)) < 1e-5
    # sinks only (no fp32 tensor at all): same partial sums
    _, rgb2, y2, _ = f.chain_conv3(a1, B, oh, ow, w_cv, d_cv, 0.07, nz2, nw, b2, (0.2, 2 ** 0.5), next_precision, rgb_w=rgb_w)
    assert y2 is None and torch.equal(rgb2[0], rgb[0])


def test_conv_small_cout_paths(dev):
    """ToRGB-shaped (512 -> 3, 1x1, every epilogue term) and RGB-sphere-shaped (3 -> 3, 3x3 stride 3) convs."""
    from spgan_b200.functional import ConvGeom
    B, C, H = 3, 512, 29
    x = synth.randn_t(12, "sx", (B, C, H, H))
    w = synth.randn_t(12, "sw", (3, C, 1, 1))
    im = synth.randn_t(12, "sim", (B, C), 0.3, 1.0)
    bias = synth.randn_t(12, "sb", (3,))
    res = synth.randn_t(12, "sr", (B, 3, H, H))
    want = _ref_conv(x, w, ConvGeom(1, 1), False, None, im, None, 0.044) + bias.double()[None, :, None, None] + res.double()
    got = SF().conv_apply(x.to(dev), w.to(dev), ConvGeom(1, 1), in_mul=im.to(dev), out_scale=0.044, bias=bias.to(dev),
                          residual=res.to(dev))
    assert K.rel_err(K.t2n(got), want.numpy()) < 1e-5
    x2 = synth.randn_t(12, "sx2", (2, 3, 51, 51))
    w2 = synth.randn_t(12, "sw2", (3, 3, 3, 3))
    g3 = ConvGeom(3, 3, stride=3)
    want = F.leaky_relu(_ref_conv(x2, w2, g3, False, None, None, None, 0.19) + bias.double()[None, :, None, None], 0.01)
    got = SF().conv_apply(x2.to(dev), w2.to(dev), g3, out_scale=0.19, bias=bias.to(dev), act=(0.01, 1.0))
    assert K.rel_err(K.t2n(got), want.numpy()) < 1e-5


# ---------------------------------------------------------------------------------------------- modulated conv modules
def _modconv_module(name, cin, cout, k, demod, up, dev):
    from spgan_b200.models import ops
    from spgan_b200.generator import default_config
    m = ops.ModulatedConv2d(cin, cout, k, K.STYLE_DIM, demodulate=demod, upsample=up, no_zero_pad=True,
                            blur_kernel=[1, 2, 1], config=default_config(), side="ts")
    p = K.modconv_params(name, cin, cout, k)
    with torch.no_grad():
        m.weight.copy_(p["weight"])
        m.modulation.weight.copy_(p["modulation.weight"])
        m.modulation.bias.copy_(p["modulation.bias"])
    return m.to(dev)


@pytest.mark.parametrize("case", K.MODCONV_CASES, ids=[c[0] for c in K.MODCONV_CASES])
def test_modconv_module_golden_forward_and_grads(dev, case):
    name, cin, cout, k, demod, up, B, H = case
    g = K.load("modconv.npz")
    m = _modconv_module(name, cin, cout, k, demod, up, dev)
    x = synth.randn_t(K.SEED, "mc_x_" + name, (B, cin, H, H)).to(dev)
    s = synth.randn_t(K.SEED, "mc_s_" + name, (B, K.STYLE_DIM)).to(dev)
    with torch.no_grad():
        y0, _ = m(x, s)
    assert K.rel_err(K.t2n(y0), g["y_" + name]) < 1e-5
    x.requires_grad_(True)
    s.requires_grad_(True)
    y, _ = m(x, s)
    assert K.rel_err(K.t2n(y), g["y_" + name]) < 1e-5
    go = synth.randn_t(K.SEED, "mc_go_" + name, y.shape).to(dev)
    grads = torch.autograd.grad(y, [x, s, m.weight, m.modulation.weight, m.modulation.bias], go)
    for gname, got in zip(("gx", "gs", "gw", "gmw", "gmb"), grads):
        assert K.rel_err(K.t2n(got), g[gname + "_" + name]) < 2e-5, gname


@pytest.mark.parametrize("case", K.SPHERE_CASES, ids=[c[0] for c in K.SPHERE_CASES])
def test_sphere_modconv_module_golden_forward_and_grads(dev, case):
    name, B, C, cout, h, cps = case
    from spgan_b200.models import spgan_ops_gs
    from spgan_b200.generator import default_config
    g = K.load("sphere_modconv.npz")
    m = spgan_ops_gs.ModulatedConv2d(C + 3, cout, 3, K.STYLE_DIM, no_zero_pad=True, config=default_config(), side="ss",
                                     deal_coords=True)
    p = K.sphere_params(name, C + 3, cout)
    with torch.no_grad():
        m.weight.copy_(p["weight"])
        m.modulation.weight.copy_(p["modulation.weight"])
        m.modulation.bias.copy_(p["modulation.bias"])
    m = m.to(dev)
    x = synth.randn_t(K.SEED, "smc_x_" + name, (B, C, h, h)).to(dev)
    c = synth.randn_t(K.SEED, "smc_c_" + name, (B, 3, h, h)).to(dev)
    s = synth.randn_t(K.SEED, "smc_s_" + name, (B, K.STYLE_DIM)).to(dev)
    with torch.no_grad():
        y0, _ = m(x, s, coords=c.clone(), coords_partial=cps)
    assert K.rel_err(K.t2n(y0), g["y_" + name]) < 1e-5
    x.requires_grad_(True)
    s.requires_grad_(True)
    y, _ = m(x, s, coords=c.clone(), coords_partial=cps)
    assert K.rel_err(K.t2n(y), g["y_" + name]) < 1e-5
    go = synth.randn_t(K.SEED, "smc_go_" + name, y.shape).to(dev)
    gx, gs, gw = torch.autograd.grad(y, [x, s, m.weight], go)
    assert K.rel_err(K.t2n(gx), g["gx_" + name]) < 2e-5
    assert K.rel_err(K.t2n(gs), g["gs_" + name]) < 2e-5
    assert K.rel_err(K.t2n(gw), g["gw_" + name]) < 2e-5


@pytest.mark.parametrize("precision", MODES_16)
def test_sphere_tcgen05_fused_vs_oracle(dev, precision):
    """The fused producer + tcgen05 GEMM at tensor-path channel counts, train (per-sample grids) and test (shared)."""
    for B, C, Oc, h, cps in ((2, 61, 32, 17, [K.train_cp(7, 139, 17), K.train_cp(1, 20, 17)]),
                             (3, 125, 64, 23, K.test_cp(2, 7, 27))):
        x = synth.randn_t(9, "sfx%d" % B, (B, C, h, h))
        c = synth.randn_t(9, "sfc%d" % B, (B, 3, h, h))
        w = synth.randn_t(9, "sfw%d" % B, (1, Oc, C + 3, 3, 3))
        s = synth.randn_t(9, "sfs%d" % B, (B, C + 3), 0.3, 1.0)
        wm = O.modulated_weight(w, s, True)
        grid = torch.from_numpy(O.batch_sampling_grid(h, h, cps, B))
        xs = O.gather_t(x, grid)
        cs = O.encode_coords(O.gather_t(c, grid))
        inp = torch.cat([xs.reshape(1, B * C, 3 * h, 3 * h), cs.reshape(1, B * 3, 3 * h, 3 * h)], 1)
        want = F.leaky_relu(F.conv2d(inp, wm.reshape(B * Oc, C + 3, 3, 3), groups=B, stride=3).reshape(B, Oc, h, h), 0.01)
        scale = 1 / np.sqrt((C + 3) * 9)
        d = SF().demod_coefficients(w[0].to(dev), s.to(dev), scale)
        g1 = grid[:1] if isinstance(cps, dict) else grid
        got = SF().sphere_modconv_fused(x.to(dev), c.to(dev), g1.to(dev), w[0].to(dev), s.to(dev), d, scale,
                                        act=(0.01, 1.0), precision=precision)
        assert K.rel_err(K.t2n(got), want.numpy()) < TOL[precision]


@pytest.mark.parametrize("precision", MODES_16)
@pytest.mark.parametrize("B,C,Oc,h,extras", [(3, 70, 40, 17, False), (2, 256, 256, 23, True), (5, 61, 288, 12, True)])
def test_sphere_gather_producer_gemm_vs_pack_then_gemm(dev, precision, B, C, Oc, h, extras):
    """spgan_sphere_conv_gemm (the gather runs in the GEMM's producer warps, A tiles exist only in shared memory) against
    spgan_sphere_pack + spgan_conv_gemm (the operand it no longer writes): same operand bits, the K dimension is merely
    walked channel-block-major instead of tap-major, so only the fp32 accumulation order differs.  Covers tiles that span
    two samples, a ragged last M tile, one and two N tiles (the second ragged), the flat-concat table with the coordinate
    planes in the last group, bias + residual + activation."""
    f = SF()
    cp = K.test_cp(2, 7, 27)
    grid = torch.from_numpy(O.gen_sampling_grid(h, h, cp)).to(dev)
    x = synth.randn_t(31, "sgx%d" % B, (B, C, h, h)).to(dev)
    c = synth.randn_t(31, "sgc%d" % B, (B, 3, h, h)).to(dev)
    w = synth.randn_t(31, "sgw%d" % B, (Oc, C + 3, 3, 3), 0.2).to(dev)
    s = synth.randn_t(31, "sgs%d" % B, (B, C + 3), 0.3, 1.0).to(dev)
    d = synth.randn_t(31, "sgd%d" % B, (B, Oc), 0.3, 1.0).to(dev)
    bias = synth.randn_t(31, "sgb%d" % B, (Oc,)).to(dev) if extras else None
    res = synth.randn_t(31, "sgr%d" % B, (B, Oc, h, h)).to(dev) if extras else None
    outs = []
    for fused in (True, False):
        f.FUSED_SPHERE_GATHER = fused
        try:
            outs.append(f.sphere_modconv_fused(x, c, grid, w, s, d, 0.05, act=(0.01, 1.0), precision=precision, residual=res, bias=bias))
        finally:
            f.FUSED_SPHERE_GATHER = False
    assert K.rel_err(K.t2n(outs[0]), K.t2n(outs[1])) < 1e-5


def test_sphere_rgb_conv_golden(dev):
    from spgan_b200.models.spherenet import SphereConvBatchDiffFixBorderGNoGrad
    g = K.load("sphere_modconv.npz")
    m = SphereConvBatchDiffFixBorderGNoGrad(3, 3)
    p = K.module_params("srgb_", {"weight": (3, 3, 3, 3), "bias": (3,)})
    with torch.no_grad():
        m.weight.copy_(p["weight"])
        m.bias.copy_(p["bias"])
    m = m.to(dev)
    x = synth.randn_t(K.SEED, "srgb_x", (2, 3, 17, 17)).to(dev)
    cps = [K.train_cp(7, 139, 17), K.train_cp(1, 20, 17)]
    with torch.no_grad():
        y0 = m(x, cps)
    assert K.rel_err(K.t2n(y0), g["y_srgb"]) < 1e-5
    x.requires_grad_(True)
    y = m(x, cps)
    go = synth.randn_t(K.SEED, "srgb_go", y.shape).to(dev)
    gx, gw, gb = torch.autograd.grad(y, [x, m.weight, m.bias], go)
    assert K.rel_err(K.t2n(y), g["y_srgb"]) < 1e-5
    assert K.rel_err(K.t2n(gx), g["gx_srgb"]) < 2e-5
    assert K.rel_err(K.t2n(gw), g["gw_srgb"]) < 2e-5
    assert K.rel_err(K.t2n(gb), g["gb_srgb"]) < 2e-5


def test_equal_linear_vs_oracle(dev):
    x = synth.randn_t(3, "lx", (5, 512))
    w = synth.randn_t(3, "lw", (259, 512))
    b = synth.randn_t(3, "lb", (259,))
    want = O.equal_linear(x, w, b, lr_mul=1.0)
    got = SF().equal_linear(x.to(dev), w.to(dev), b.to(dev), 1 / np.sqrt(512), 1.0, False)
    assert K.rel_err(K.t2n(got), K.t2n(want)) < 1e-5
    w2 = synth.randn_t(3, "lw2", (512, 512), 100.0)
    b2 = synth.randn_t(3, "lb2", (512,))
    want = O.equal_linear(x, w2, b2, lr_mul=0.01, activation=True)
    got = SF().equal_linear(x.to(dev), w2.to(dev), b2.to(dev), 0.01 / np.sqrt(512), 0.01, True)
    assert K.rel_err(K.t2n(got), K.t2n(want)) < 1e-5
    xg = x.to(dev).requires_grad_(True)
    wg = w2.to(dev).requires_grad_(True)
    bg = b2.to(dev).requires_grad_(True)
    y = SF().equal_linear(xg, wg, bg, 0.01 / np.sqrt(512), 0.01, True)
    go = synth.randn_t(3, "lgo", (5, 512))
    gx, gw, gb = torch.autograd.grad(y, [xg, wg, bg], go.to(dev))
    xr, wr, br = x.clone().requires_grad_(True), w2.clone().requires_grad_(True), b2.clone().requires_grad_(True)
    rx, rw, rb = torch.autograd.grad(O.equal_linear(xr, wr, br, lr_mul=0.01, activation=True), [xr, wr, br], go)
    assert K.rel_err(K.t2n(gx), K.t2n(rx)) < 1e-5 and K.rel_err(K.t2n(gw), K.t2n(rw)) < 1e-5
    assert K.rel_err(K.t2n(gb), K.t2n(rb)) < 1e-5


def test_linear_second_order_vs_torch(dev):
    """Gradient-of-gradient through EqualLinear (R1 through the discriminator heads, path length through the mapping
    network): the dedicated weight-gradient kernel and its two adjoint maps against torch autograd of F.linear."""
    x = synth.randn_t(21, "l2x", (6, 40)).to(dev).requires_grad_(True)
    w = synth.randn_t(21, "l2w", (24, 40)).to(dev).requires_grad_(True)
    b = synth.randn_t(21, "l2b", (24,)).to(dev).requires_grad_(True)
    go = synth.randn_t(21, "l2g", (6, 24)).to(dev)
    def second(fn):
        y = fn(x, w, b)
        gx, gw = torch.autograd.grad((y * go).sum() + (y ** 2).sum(), [x, w], create_graph=True)
        loss = (gx ** 2).sum() + (gw ** 3).sum()
        return [gx, gw] + list(torch.autograd.grad(loss, [x, w, b], allow_unused=True))
    ours = second(lambda x, w, b: SF().equal_linear(x, w, b, 0.3, 1.0, False))
    ref = second(lambda x, w, b: F.linear(x, w * 0.3, b))
    for a, r in zip(ours, ref):
        if r is None:
            assert a is None or float(a.abs().max()) == 0.0
        else:
            assert K.rel_err(K.t2n(a), K.t2n(r)) < 2e-5


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("precision,tol", [(1, 1e-4), (2, 3e-2)])
def test_structure_chain_ops_vs_padded_path(precision, tol, fused):
    """K-segment operands of the structure chain (spgan_sphere_pack_seg / spgan_coord_taps_pack feeding the second K segment
    of spgan_conv_gemm_ex) against the same convs computed with every tap padded to 320 channels (sphere_modconv_fused,
    conv_apply), two position groups with different sampling grids."""
    import spgan_b200.functional as SF  # noqa: F811 (module, not the helper above)
    from spgan_b200 import grids
    dev = torch.device("cuda:0")
    G, Bg, C, H = 2, 3, 256, 17
    B = G * Bg
    cps = [{"p_x_st": 6 / 65, "p_x_ed": 42 / 65, "p_y_st": 12 / 48, "p_y_ed": 48 / 48, "circular_flag": False, "x_total": 65,
            "y_total": 48, "test_flag": True, "partial": 0.6667},
           {"p_x_st": 12 / 65, "p_x_ed": 48 / 65, "p_y_st": 30 / 48, "p_y_ed": 66 / 48, "circular_flag": True, "x_total": 65,
            "y_total": 48, "test_flag": True, "partial": 0.6667}]
    x = synth.randn_t(3, "ssc_x", (B, C, H, H)).to(dev)
    coords = synth.randn_t(3, "ssc_c", (B, 3, H, H)).to(dev)
    w_s = synth.randn_t(3, "ssc_ws", (256, C + 3, 3, 3)).to(dev)
    w_7 = synth.randn_t(3, "ssc_w7", (256, C + 3, 7, 7)).to(dev)
    w_sc = synth.randn_t(3, "ssc_wsc", (256, C, 1, 1), 0.05).to(dev)
    b_sc = synth.randn_t(3, "ssc_bsc", (256,), 0.1).to(dev)
    b_7 = synth.randn_t(3, "ssc_b7", (256,), 0.1).to(dev)
    s_s = synth.randn_t(3, "ssc_ss", (B, C + 3), 0.3, 1.0).to(dev)
    s_7 = synth.randn_t(3, "ssc_s7", (B, C + 3), 0.3, 1.0).to(dev)
    sc_s, sc_7 = 1 / np.sqrt((C + 3) * 9), 1 / np.sqrt((C + 3) * 49)
    d_s = SF.demod_coefficients(w_s, s_s, sc_s)
    d_7 = SF.demod_coefficients(w_7, s_7, sc_7)
    with torch.no_grad():
        # reference composition, one position group at a time (the flat-concat table is per call)
        want = []
        for i, cp in enumerate(cps):
            sl = slice(i * Bg, (i + 1) * Bg)
            grid = grids.GRID_CACHE.get(H, H, cp, dev)
            sph = SF.sphere_modconv_fused(x[sl], coords[sl], grid, w_s, s_s[sl].contiguous(), d_s[sl].contiguous(), sc_s,
                                          act=(0.01, 1.0), precision=precision)
            h = SF.conv_apply(x[sl].contiguous(), w_sc, SF.ConvGeom(1, 1), bias=b_sc, residual=sph, precision=precision)
            inp = torch.cat([h, SF.encode_coords(coords[sl])], 1)
            want.append(SF.conv_apply(inp, w_7, SF.ConvGeom(7, 7), in_mul=s_7[sl].contiguous(), out_mul=d_7[sl].contiguous(),
                                      out_scale=sc_7, bias=b_7, act=(0.2, 2 ** 0.5), precision=precision))
        want = torch.cat(want, 0)
        # chain: the spherical conv as ONE kernel (gather in the GEMM's producer warps) or as operand producer + GEMM
        xh, xp = SF.ss_input(x, precision)
        grid = grids.GRID_CACHE.group_grid(H, H, cps, dev)
        y_sc = SF.ss_shortcut(xp, B, H, H, w_sc, b_sc, precision)
        prev, SF.SS_FUSED_GATHER = SF.SS_FUSED_GATHER, fused
        try:
            launches = SF.lib.load().spgan_gemm_launch_count()
            a = SF.ss_sphere(xh, coords, grid, Bg, w_s, s_s, d_s, sc_s, (0.01, 1.0), y_sc, s_7[:, :C].contiguous(), precision)
            assert SF.lib.load().spgan_gemm_launch_count() == launches + 1
        finally:
            SF.SS_FUSED_GATHER = prev
        got, _, hw = SF.ss_conv_k(a, coords, B, H, H, w_7, s_7, d_7, sc_7, b_7, (0.2, 2 ** 0.5), precision, True)
        nh, pk, _ = SF.ss_conv_k(a, coords, B, H, H, w_7, s_7, d_7, sc_7, b_7, (0.2, 2 ** 0.5), precision, False)
    assert hw == (11, 11) and got.shape == want.shape
    err = K.rel_err(K.t2n(got), K.t2n(want))
    print("structure chain ops, precision %d, fused gather %s: %.2e" % (precision, fused, err))
    assert err < tol
    # the two sink sets of the 7x7 GEMM carry the same values: NHWC fp32 and the packed hi/lo planes
    assert torch.equal(nh.permute(0, 3, 1, 2), got)
    back = SF.packed_to_float(pk, 0).view(B, 11, 11, 256).permute(0, 3, 1, 2)
    assert K.rel_err(K.t2n(back), K.t2n(got)) < (2e-5 if precision == 1 else 1e-2)


@pytest.mark.parametrize("precision", [1, 2, 3])
@pytest.mark.parametrize("shape", [(3, 128, 256, 23, 3), (2, 64, 512, 31, 3), (5, 256, 256, 17, 7), (2, 96, 512, 19, 1)])
def test_gemm_cta_pair_kernel_bit_identical(precision, shape):
    """The cta_group::2 kernel (256 x 256 tiles over two SMs, weight tile split between the CTAs, multicast commits) issues the
    same products in the same K order as the single-CTA kernel: bit-identical outputs, im2col and flat A loads, odd numbers of
    M tiles (a half-empty last pair), epilogue terms included."""
    import spgan_b200.functional as SF  # noqa: F811
    dev = torch.device("cuda:0")
    B, C, O, H, k = shape
    x = synth.randn_t(5, "pair_x", (B, C, H, H)).to(dev)
    w = synth.randn_t(5, "pair_w", (O, C, k, k)).to(dev)
    s = synth.randn_t(5, "pair_s", (B, C), 0.3, 1.0).to(dev)
    if precision == 3:
        x = x.clamp(-8, 8)
    oh = H - k + 1
    nz = synth.randn_t(5, "pair_nz", (B, 1, oh, oh)).to(dev)
    nw = torch.tensor([0.3], device=dev)
    bias = synth.randn_t(5, "pair_b", (O,), 0.1).to(dev)
    d = SF.demod_coefficients(w, s, 0.05)
    outs = []
    try:
        for mode in (0, 2):
            SF.set_gemm_pair_mode(mode)
            before = SF.lib.load().spgan_gemm_launch_count()
            with torch.no_grad():
                outs.append(SF.conv_apply(x, w, SF.ConvGeom(k, k), in_mul=s, out_mul=d, out_scale=0.05, noise=nz, noise_w=nw, bias=bias,
                                          act=(0.2, 2 ** 0.5), precision=precision))
            assert SF.lib.load().spgan_gemm_launch_count() == before + 1
    finally:
        SF.set_gemm_pair_mode(1)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])
    want = F.conv2d((x * s[:, :, None, None]).double(), w.double()) * 0.05 * d.double()[:, :, None, None]
    want = F.leaky_relu(want + 0.3 * nz.double() + bias.double().view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    assert K.rel_err(K.t2n(outs[1]), K.t2n(want)) < TOL[precision]


@pytest.mark.parametrize("M,N,K", [(1, 512, 512), (8, 259, 512), (16, 3, 512), (33, 512, 4608), (64, 1, 512), (65, 512, 512), (8, 512, 500)])
def test_linear_small_and_general_kernels_vs_fp64(M, N, K, dev):
    """EqualLinear forward over both kernels (the small-batch kernel: M <= 64, K % 128 == 0; the general one otherwise),
    with and without the fused bias + leaky-ReLU epilogue."""
    x = synth.randn_t(9, "lin_x", (M, K)).to(dev)
    w = synth.randn_t(9, "lin_w", (N, K)).to(dev)
    b = synth.randn_t(9, "lin_b", (N,)).to(dev)
    for act in (False, True):
        with torch.no_grad():
            got = SF().equal_linear(x, w, b, 0.04, 0.5, act)
        want = x.double() @ (w.double() * 0.04).t() + b.double() * 0.5
        if act:
            want = F.leaky_relu(want, 0.2) * 2 ** 0.5
        assert K_rel(got, want) < 2e-6


def K_rel(a, b):
    return K.rel_err(K.t2n(a), K.t2n(b))


@pytest.mark.parametrize("shape", [(3, 6, 19, 19), (4, 6, 19, 19), (2, 3, 8, 8), (8, 5, 101, 101), (16, 512, 9, 9)])
def test_bias_act_backward_kernels_vs_torch(dev, shape):
    """First-order backward of fused_leaky_relu (models/custom_ops/fused_act.py:24-44) over both kernels: per-channel scalar
    kernel and the 128-bit flat kernel (float4s straddling plane / channel / sample boundaries, bias gradient through
    shared-memory bins)."""
    B, C, H, W = shape
    go = synth.randn_t(4, "bab_go", shape).to(dev)
    out = synth.randn_t(4, "bab_out", shape).to(dev)
    with torch.no_grad():
        gi, gb = SF().FusedLeakyReLUFunctionBackward.apply(go, out, 0.2, 2 ** 0.5)
    want = go.double() * torch.where(out > 0, 1.0, 0.2).double() * 2 ** 0.5
    assert K_rel(gi, want) < 1e-6
    assert K_rel(gb, want.sum(dim=(0, 2, 3))) < 2e-6
