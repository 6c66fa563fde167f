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


@pytest.mark.parametrize("precision,tol", [(0, 2e-4), (1, 5e-4)])
@pytest.mark.parametrize("name,B,pos", [("b1_p27", 1, (2, 7)), ("b2_p59", 2, (5, 9))])
def test_generator_patch_golden(gen, name, B, pos, precision, tol):
    import spgan_b200.functional as SF
    g = K.load("generator.npz")
    gl, lat, coords, cp, noises = K.generator_case(name, B, *pos)
    SF.set_precision(precision)
    try:
        with torch.no_grad():
            img = gen(gl.cuda(), lat.cuda(), coords.cuda(), cp, noises=[n.cuda() for n in noises])
    finally:
        SF.set_precision(1)
    assert img.shape == (B, 3, 101, 101)
    assert K.rel_err(K.t2n(img), g["img_" + name]) < tol


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
    assert K.rel_err(img[:, :, 250:290, :], ref["strip"]) < 5e-4
    assert K.rel_err(img[:, :, :, 740:768], ref["col_seam"]) < 5e-4  # longitude seam columns
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
