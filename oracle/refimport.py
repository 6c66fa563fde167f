"""TEST INFRASTRUCTURE ONLY (oracle/): import the real reference.

Installs the five in-process stubs of SURVEY.md §A.4 so that the reference (read-only, Python/PyTorch) can be
imported and run on CPU without CUDA, easydict, lmdb or matplotlib.  The reference tree is looked up at
$SPGAN_REFERENCE_ROOT, /root/reference (the build container) or oracle/_ref/reference (a staged copy made by
`oracle/build_ref.py`: git-ignored, so no reference source enters the history, but it travels to the GPU box with the
snapshot, where the parity tests and bench.py's reference arm run the real reference as the checker / CPU baseline).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(HERE, "_ref", "reference")


def _find_root():
    for cand in (os.environ.get("SPGAN_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if cand and os.path.isdir(os.path.join(cand, "models")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


class _AttrDict(dict):
    """Minimal stand-in for easydict.EasyDict: recursive attribute access."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            return _AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(_AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, _AttrDict._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def install(real_cuda=False):
    """Make `import models...`, `import coord_handler`, `import test_managers...` resolve to the reference.
    real_cuda=True keeps torch's CUDA entry points (the drop-in tests run the reference's generator and manager on the
    GPU over the mirrored op modules); the default replaces them so that everything stays on the CPU."""
    import torch
    import torch.utils.cpp_extension as cpp_ext

    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ed = types.ModuleType("easydict")
    ed.EasyDict = _AttrDict
    sys.modules.setdefault("easydict", ed)
    sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.imsave = lambda *a, **k: None
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    for name in ("cv2", "tensorboardX", "skimage"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    cpp_ext.load = lambda *a, **k: types.SimpleNamespace()
    if not real_cuda:
        torch.cuda.get_device_name = lambda *a, **k: "cpu-stub"
        torch.Tensor.cuda = lambda self, *a, **k: self
    return _AttrDict


def load_config(real_cuda=False):
    import yaml
    EasyDict = install(real_cuda)
    with open(os.path.join(REFERENCE_ROOT, "configs/model/spgan.yaml")) as f:
        config = EasyDict(yaml.safe_load(f))
    config.var = EasyDict()
    config.var.dataparallel = False
    return config
