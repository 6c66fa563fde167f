"""CPU, build container only: the reference's own generator / discriminator construct on top of the mirrored op
modules (drop-in boundary).  Skipped where the reference tree is absent (the GPU box)."""
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present")


def test_reference_generator_and_discriminator_build_on_the_mirrors():
    import subprocess
    code = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import refimport
config = refimport.load_config()           # stubs easydict / lmdb / cuda-at-import; reference root on sys.path
import spgan_b200.dropin as dropin
names = dropin.install()
import models.ops, models.spgan_ops_gs
assert models.ops.__name__.startswith("spgan_b200."), models.ops.__name__
from models.spgan.spgan import InfinityGanGenerator
from models.stylegan2discriminator import StyleGan2Discriminator
import spgan_b200.models.ops as mirror_ops
g = InfinityGanGenerator(config)
d = StyleGan2Discriminator(config)
assert type(g.texture_synthesizer.convs[0]) is mirror_ops.StyledConv
assert type(g.structure_synthesizer.implicit_model.conv_stack[0].conv).__module__.endswith("spgan_ops_gs")
import json
manifest = json.load(open(%r))
sd = g.state_dict()
assert set(sd) == set(manifest), set(sd) ^ set(manifest)
assert all(list(sd[k].shape) == v for k, v in manifest.items())
print("G params", sum(p.numel() for p in g.parameters()), "D params", sum(p.numel() for p in d.parameters()))
''' % (os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"),
       os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generator_manifest.json"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "G params 39540776" in out.stdout and "D params 289" in out.stdout, out.stdout
