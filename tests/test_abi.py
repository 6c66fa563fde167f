"""CPU: the C-ABI library loads and exports every symbol include/spgan_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "spgan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spgan_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    import spgan_b200.lib as lib
    handle = ctypes.CDLL(lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), "libspgan_b200.so does not export %s" % n


def test_binding_table_matches_header():
    import spgan_b200.lib as lib
    assert sorted(lib.SIGNATURES) == _header_symbols()


def test_load_and_version_without_gpu():
    import spgan_b200.lib as lib
    h = lib.load()
    assert h.spgan_abi_version() == 2
    assert isinstance(lib.last_error(), str)


def test_conv_pass_struct_layout():
    """ctypes mirror and the C struct agree on size (the library is compiled from the same header)."""
    import spgan_b200.lib as lib
    n = lib.MAX_TAPS
    expect = 4 * 14 + 4 * 3 * n  # 14 leading int32 + three tap arrays
    expect = (expect + 7) // 8 * 8 + 16 + 4 * 5  # two int64 strides (8-aligned), out_scale, act, alpha, gain, precision
    expect = (expect + 7) // 8 * 8 + 8  # out_cstride (8-aligned)
    assert ctypes.sizeof(lib.ConvPass) == expect


def test_product_has_no_cpu_fallback():
    import pytest
    import torch
    import spgan_b200.functional as SF
    with pytest.raises(RuntimeError):
        SF.fused_leaky_relu(torch.zeros(1, 2, 3, 3), torch.zeros(2))
    with pytest.raises(RuntimeError):
        SF.upfirdn2d(torch.zeros(1, 2, 3, 3), torch.ones(3, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sp-gan-tip2025_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "spgan_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
