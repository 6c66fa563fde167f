"""GPU parity of the module signatures north_star pins but configs/model/spgan.yaml never instantiates, against outputs of
the real reference (tests/golden/signatures.npz, oracle/make_golden_r2.py): the other three samplers with their TRUE input
gradients, the full-sphere SphereNet convs, models/spgan_ops.py's SphereModulatedConv2d (texture sampler, `batch * 256`
view), and the spatial-style (test-time style fusion) branch of ops.ModulatedConv2d.  Tolerances: 2e-6 for the samplers
(HBM-bound fp32 kernels), 1e-5 for compositions with the exact-fp32 SIMT conv (these modules are small: Cin / Cout < 16),
1e-4 where the bf16x3 tensor path runs."""
import numpy as np
import pytest
import torch

import cases as K
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import spgan_b200.lib as lib
    torch.cuda.set_device(0)
    lib.require_device()
    return torch.device("cuda:0")


def _fill(mod, tag, dev):
    with torch.no_grad():
        for n, p in mod.named_parameters():
            p.copy_(synth.randn_t(K.SEED, tag + n, p.shape, 1.0, 1.0 if n.endswith("modulation.bias") else 0.0))
    return mod.to(dev)


@pytest.mark.parametrize("name,cls_name", [("texture", "GridSamplerNewTexture"), ("nearest", "GridSampler"), ("bilinear", "GridSamplerNew")])
def test_samplers_forward_true_gradient_and_double_backward(dev, name, cls_name):
    from spgan_b200.models.spherenet import grid_generator as GG
    g = K.load("signatures.npz")
    z = synth.randn_t(K.SEED, "sig_z", (2, 5, 9, 11)).to(dev).requires_grad_(True)
    grid = torch.from_numpy(g["samp_grid"]).to(dev)
    go = synth.randn_t(K.SEED, "sig_go", (2, 5, 13, 17)).to(dev).requires_grad_(True)
    y = getattr(GG, cls_name)()(z, grid)
    assert K.rel_err(K.t2n(y), g["samp_y_" + name]) < 2e-6
    gz, = torch.autograd.grad(y, z, go, create_graph=True)
    assert K.rel_err(K.t2n(gz), g["samp_gz_" + name]) < 2e-6
    # the op is linear in z: d<gz, v>/d go is the forward applied to v
    v = synth.randn_t(K.SEED, "sig_v", (2, 5, 9, 11)).to(dev)
    ggo, = torch.autograd.grad((gz * v).sum(), go)
    with torch.no_grad():
        want = getattr(GG, cls_name)()(v, grid)
    assert K.rel_err(K.t2n(ggo), K.t2n(want)) < 2e-6


@pytest.mark.parametrize("name,cls_name,k,stride,shape", [("full", "SphereConv2d", (3, 3), 1, (2, 4, 10, 14)),
                                                           ("incre", "IncreIntervalSphereConv2d", (3, 3), 2, (2, 4, 12, 16))])
def test_full_sphere_convs_golden(dev, name, cls_name, k, stride, shape):
    import spgan_b200.models.spherenet as SN
    g = K.load("signatures.npz")
    m = _fill(getattr(SN, cls_name)(4, 6, kernel_size=k, stride=stride, scale=0.3), "sig_" + name + "_", dev)
    x = synth.randn_t(K.SEED, "sig_x_" + name, shape).to(dev).requires_grad_(True)
    y = m(x)
    assert K.rel_err(K.t2n(y), g["sc_y_" + name]) < 1e-5
    gg = synth.randn_t(K.SEED, "sig_g_" + name, tuple(y.shape)).to(dev)
    gx, gw, gb = torch.autograd.grad(y, [x, m.weight, m.bias], gg)
    assert K.rel_err(K.t2n(gx), g["sc_gx_" + name]) < 1e-5
    assert K.rel_err(K.t2n(gw), g["sc_gw_" + name]) < 1e-5
    assert K.rel_err(K.t2n(gb), g["sc_gb_" + name]) < 1e-5


def test_spgan_ops_sphere_modulated_conv_golden(dev):
    """models/spgan_ops.py:736-1379: texture sampler (true gradient) + the batch * 256 view."""
    from spgan_b200.generator import default_config
    from spgan_b200.models import spgan_ops
    g = K.load("signatures.npz")
    cfg = default_config()
    m = _fill(spgan_ops.SphereModulatedConv2d(256 + 3, 4, 3, K.STYLE_DIM, no_zero_pad=True, config=cfg, side="ss", deal_coords=True),
              "sig_smc_", dev)
    x = synth.randn_t(K.SEED, "sig_smc_x", (2, 256, 11, 11)).to(dev).requires_grad_(True)
    c = synth.randn_t(K.SEED, "sig_smc_c", (2, 3, 11, 11)).to(dev)
    s = synth.randn_t(K.SEED, "sig_smc_s", (2, K.STYLE_DIM)).to(dev).requires_grad_(True)
    cps = [K.train_cp(7, 139, 11), K.train_cp(1, 20, 11)]
    y, _ = m(x, s, coords=c.clone(), coords_partial=cps)
    assert K.rel_err(K.t2n(y), g["smc_y"]) < 1e-5
    gg = synth.randn_t(K.SEED, "sig_smc_g", tuple(y.shape)).to(dev)
    gx, gs, gw = torch.autograd.grad(y, [x, s, m.weight], gg)
    assert K.rel_err(K.t2n(gx), g["smc_gx"]) < 1e-5
    assert K.rel_err(K.t2n(gs), g["smc_gs"]) < 1e-5
    assert K.rel_err(K.t2n(gw), g["smc_gw"]) < 1e-5
    with torch.no_grad():
        y2, _ = m(x.detach(), s.detach(), coords=c.clone(), coords_partial=cps)
    assert K.rel_err(K.t2n(y2), g["smc_y"]) < 1e-5
    # any other feature width fails like the reference's view(1, batch * 256, ...)
    m2 = spgan_ops.SphereModulatedConv2d(8 + 3, 4, 3, K.STYLE_DIM, no_zero_pad=True, config=cfg, side="ss", deal_coords=True).to(dev)
    with pytest.raises(RuntimeError):
        m2(torch.zeros(2, 8, 11, 11, device=dev), s.detach(), coords=c, coords_partial=cps)


@pytest.mark.parametrize("name,up", [("plain", False), ("up", True)])
def test_spatial_style_branch_golden(dev, name, up):
    """models/ops.py:637-729: per-pixel style modulation of the activations, per-pixel demodulation estimate."""
    from spgan_b200.generator import default_config
    from spgan_b200.models import ops
    g = K.load("signatures.npz")
    m = _fill(ops.ModulatedConv2d(6, 5, 3, K.STYLE_DIM, upsample=up, no_zero_pad=True, blur_kernel=[1, 2, 1], config=default_config(),
                                  side="ts"), "sig_sp_" + name + "_", dev).eval()
    x = synth.randn_t(K.SEED, "sig_sp_x_" + name, (2, 6, 9, 9)).to(dev)
    st = synth.randn_t(K.SEED, "sig_sp_s_" + name, (2, K.STYLE_DIM, 11, 11)).to(dev)
    with torch.no_grad():
        y, _ = m(x, st)
    assert K.rel_err(K.t2n(y), g["sp_y_" + name]) < 1e-5
    # through StyledConv (the fused path must step aside for a spatial style)
    sc = ops.StyledConv(6, 5, 3, K.STYLE_DIM, upsample=up, blur_kernel=[1, 2, 1], no_zero_pad=True, config=default_config(), side="ts").to(dev).eval()
    with torch.no_grad():
        out, _ = sc(x, st, noise=torch.zeros(2, 1, y.shape[2], y.shape[3], device=dev))
    assert out.shape == y.shape and torch.isfinite(out).all()
