"""GPU: the REFERENCE's own generator and close-loop manager run over the drop-in mirrors.

The reference tree is staged under oracle/_ref/ by oracle/build_ref.py (git-ignored, travels with the snapshot).  In a
subprocess (dropin.install() rebinds `models.*` in sys.modules) the reference's `InfinityGanGenerator`
(models/spgan/spgan.py:1278) is built on top of `spgan_b200.models.*`, moved to the GPU and (1) called on the golden
B = 2 patch case, (2) driven by the reference's `InfiniteGenerationManagerPatchCoordsCloseLoop.generate`
(test_managers/close_loop_infinite_generation.py:170-305) over a whole 384x768 panorama; both are compared with the
fixtures the unmodified reference produced on CPU (generator.npz, panorama_384.npz)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refimport  # noqa: E402

CODE = r'''
import json, os, sys
import numpy as np
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refrun
config, Gen = refrun.load(dropin=True, real_cuda=True)
import torch
import spgan_b200.lib as lib
import spgan_b200.models.ops as mirror_ops
import cases as K
import spgan_oracle as O
torch.cuda.set_device(0)
lib.require_device()
# SphereConditionalBlock.sc is a plain nn.Conv2d of the reference (models/spgan/spgan.py:141), not a mirrored module: keep
# cuDNN in true fp32 (torch's default lets it use TF32, error ~1e-3) so the comparison with the CPU fixtures is fp32 vs fp32
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
gen = refrun.synthetic_generator(config, Gen, K.load_json("generator_manifest.json")).cuda().eval()
assert type(gen.texture_synthesizer.convs[0]) is mirror_ops.StyledConv
out = {}
l0 = lib.launches()
g = K.load("generator.npz")
with torch.no_grad():
    for name, B, pos in (("b1_p27", 1, (2, 7)), ("b2_p59", 2, (5, 9))):
        gl, lat, coords, cp, noises = K.generator_case(name, B, *pos)
        img = gen(global_latent=gl.cuda(), local_latent=lat.cuda(), override_coords=coords.cuda().clone(),
                  coords_partial_override=cp, noises=[n.cuda() for n in noises], disable_dual_latents=True)["gen"]
        out["patch_" + name] = K.rel_err(K.t2n(img), g["img_" + name])
    mgr = refrun.manager(gen, config, "cuda", 384, 768)
    plan = O.close_loop_plan(384, 768)
    tv, _, _, _ = refrun.testing_vars(mgr, plan, "pano", device="cuda")
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        mgr.generate(tv, disable_pbar=True)
ref = K.load("panorama_384.npz")
img = K.t2n(tv.meta_img)
out["strip"] = K.rel_err(img[:, :, 250:290, :], ref["strip"])
out["seam"] = K.rel_err(img[:, :, :, 740:768], ref["col_seam"])
out["mean"] = abs(float(img.mean()) - float(ref["mean"])) / float(ref["std"])
out["launches"] = lib.launches() - l0
print("RESULT " + json.dumps(out))
'''


@pytest.mark.skipif(not refimport.available(), reason="reference tree not staged (run oracle/build_ref.py where /root/reference exists)")
def test_reference_generator_and_manager_run_over_dropin_on_gpu():
    out = subprocess.run([sys.executable, "-c", CODE % ROOT], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-4000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    print(res)
    assert res["launches"] > 4000, "the reference's modules must have run the C-ABI kernels"
    assert res["patch_b1_p27"] < 5e-4 and res["patch_b2_p59"] < 5e-4, res
    assert res["strip"] < 5e-4 and res["seam"] < 5e-4 and res["mean"] < 1e-3, res
