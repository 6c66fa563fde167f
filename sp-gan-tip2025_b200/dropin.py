"""Drop-in installation into a reference checkout.

    import spgan_b200.dropin as dropin
    dropin.install()                     # before `import models.spgan.spgan` / `models.stylegan2discriminator`

After this, `models.custom_ops`, `models.spherenet`, `models.ops`, `models.spgan_ops` and `models.spgan_ops_gs`
resolve to the mirrors of this package, while `models.spgan.*`, `models.stylegan2discriminator`, `models.losses`,
`train.py`, `test.py` and the test managers keep coming from the reference tree (they only compose the op modules).
The reference's own JIT build of its CUDA extensions (models/custom_ops/*.py:11-22) is never triggered.
"""
import importlib
import sys

MIRRORED = ("custom_ops", "spherenet", "ops", "spgan_ops", "spgan_ops_gs")


def install():
    from . import models as mirrors
    ref_models = importlib.import_module("models")  # the reference's package must be importable (its root on sys.path)
    for name in MIRRORED:
        # import order matters: custom_ops and spherenet first (ops imports them through the package-relative path)
        mod = importlib.import_module(mirrors.__name__ + "." + name)
        sys.modules["models." + name] = mod
        setattr(ref_models, name, mod)
    # submodules that the reference imports by their dotted path
    sys.modules["models.spherenet.grid_generator"] = importlib.import_module(mirrors.__name__ + ".spherenet.grid_generator")
    sys.modules["models.spherenet.sphere_conv2d"] = importlib.import_module(mirrors.__name__ + ".spherenet.sphere_conv2d")
    return [("models." + n) for n in MIRRORED]
