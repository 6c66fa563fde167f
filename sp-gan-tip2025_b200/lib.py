"""ctypes binding of libspgan_b200.so (the C ABI declared in include/spgan_b200.h).

The library is built in-tree by `csrc/build.py` (nvcc, sm_100a).  Nothing here falls back to PyTorch: if the
library is missing, or a call returns non-zero, a RuntimeError is raised — the same failure mode as the
reference's `TORCH_CHECK`s in models/custom_ops/fused_bias_act_kernel.cu:52-99.
"""
import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libspgan_b200.so")
MAX_TAPS = 49

c_int, c_i64, c_f32, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class ConvPass(ctypes.Structure):
    """Mirror of `SpganConvPass` (include/spgan_b200.h)."""
    _fields_ = [
        ("B", ctypes.c_int32), ("Cin", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
        ("Cout", ctypes.c_int32), ("out_H", ctypes.c_int32), ("out_W", ctypes.c_int32),
        ("My", ctypes.c_int32), ("Mx", ctypes.c_int32),
        ("in_stride", ctypes.c_int32),
        ("out_stride", ctypes.c_int32), ("out_off_y", ctypes.c_int32), ("out_off_x", ctypes.c_int32),
        ("ntaps", ctypes.c_int32),
        ("tap_dy", ctypes.c_int32 * MAX_TAPS), ("tap_dx", ctypes.c_int32 * MAX_TAPS), ("tap_w", ctypes.c_int32 * MAX_TAPS),
        ("ws_o", ctypes.c_int64), ("ws_c", ctypes.c_int64),
        ("out_scale", ctypes.c_float),
        ("act", ctypes.c_int32),
        ("act_alpha", ctypes.c_float), ("act_gain", ctypes.c_float),
        ("precision", ctypes.c_int32),
        ("out_cstride", ctypes.c_int64),
    ]


class GemmIO(ctypes.Structure):
    """Mirror of `SpganGemmIO` (include/spgan_b200.h)."""
    _fields_ = [
        ("a_packed", c_vp), ("a_rows", ctypes.c_int64), ("kp", ctypes.c_int32), ("fmt", ctypes.c_int32),
        ("w_packed", c_vp), ("w_fmt", ctypes.c_int64), ("out_mul", c_vp), ("noise", c_vp), ("noise_w", c_vp), ("bias", c_vp), ("residual", c_vp),
        ("y", c_vp), ("y_layout", ctypes.c_int32), ("rgb_n", ctypes.c_int32), ("y_bstride", ctypes.c_int64),
        ("y_packed", c_vp), ("next_mul", c_vp), ("y_packed_rows", ctypes.c_int64), ("y_packed_cols", ctypes.c_int32),
        ("y_packed_fmt", ctypes.c_int32), ("rgb_w", c_vp), ("rgb_part", c_vp),
        ("residual_nhwc", c_vp), ("res_bstride", ctypes.c_int64),
        ("a2_packed", c_vp), ("a2_rows", ctypes.c_int64), ("w2_packed", c_vp), ("kp2", ctypes.c_int32), ("reserved0", ctypes.c_int32),
    ]


class SphereIn(ctypes.Structure):
    """Mirror of `SpganSphereIn` (include/spgan_b200.h)."""
    _fields_ = [("x_nhwc", c_vp), ("coords", c_vp), ("grid", c_vp), ("in_mul", c_vp), ("chan_map", c_vp),
                ("C", ctypes.c_int32), ("Cp", ctypes.c_int32), ("xg", c_vp), ("grid_group", ctypes.c_int32),
                ("cmap_ld", ctypes.c_int32)]


_PASS_P = ctypes.POINTER(ConvPass)

# name -> (restype, argtypes).  tests/test_abi.py checks this table against the header, symbol by symbol.
SIGNATURES = {
    "spgan_abi_version": (c_int, []),
    "spgan_last_error": (ctypes.c_char_p, []),
    "spgan_device_ok": (c_int, []),
    "spgan_bias_act": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_int, c_int, c_f32, c_f32, c_vp]),
    "spgan_bias_act_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_vp]),
    "spgan_noise_bias_act": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_vp]),
    "spgan_upblur_act": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_int, c_f32, c_f32, c_vp]),
    "spgan_upfirdn2d": (c_int, [c_vp, c_vp, c_vp, c_i64] + [c_int] * 12 + [c_vp]),
    "spgan_sphere_gather": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_i64, c_i64, c_int, c_vp]),
    "spgan_sphere_gather_plan": (c_int, [c_int] * 5 + [c_vp]),
    "spgan_upfirdn2d_plan": (c_int, [c_i64] + [c_int] * 10 + [c_vp]),
    "spgan_sphere_gather_indices": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp]),
    "spgan_sphere_gather_bwd": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp]),
    "spgan_grid_sample": (c_int, [c_vp, c_vp, c_vp] + [c_int] * 8 + [c_vp]),
    "spgan_grid_sample_bwd": (c_int, [c_vp, c_vp, c_vp] + [c_int] * 8 + [c_vp]),
    "spgan_sphere_grid_assemble": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, ctypes.c_double, c_vp]),
    "spgan_linear": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_f32, c_f32, c_int, c_f32, c_f32, c_vp]),
    "spgan_linear_wgrad": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_f32, c_vp]),
    "spgan_conv_pass": (c_int, [_PASS_P] + [c_vp] * 10),
    "spgan_demod": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_f32, c_vp]),
    "spgan_conv_wgrad": (c_int, [_PASS_P, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "spgan_plane_dot": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "spgan_pack_act": (c_int, [c_vp, c_vp, c_vp] + [c_int] * 11 + [c_vp]),
    "spgan_conv_wgrad_gemm_workspace": (c_i64, [_PASS_P]),
    "spgan_conv_wgrad_gemm": (c_int, [_PASS_P, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, ctypes.POINTER(ctypes.c_int32),
                                      c_int, c_vp, c_i64, c_int, c_vp]),
    "spgan_pack_weight": (c_int, [c_vp, c_vp, c_int, c_int, c_i64, c_i64, c_int, ctypes.POINTER(ctypes.c_int32), c_int, c_int, c_int, c_vp]),
    "spgan_nchw_to_nhwc": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    "spgan_sphere_pack": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "spgan_sphere_pack_seg_scratch": (c_i64, [c_int, c_int, c_int, c_int]),
    "spgan_sphere_concat_repack": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "spgan_sphere_pack_seg": (c_int, [c_vp] * 7 + [c_int] * 9 + [c_vp, c_vp]),
    "spgan_coord_taps_pack": (c_int, [c_vp, c_vp, c_vp] + [c_int] * 10 + [c_vp]),
    "spgan_conv_gemm": (c_int, [_PASS_P, c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "spgan_conv_gemm_ex": (c_int, [_PASS_P, ctypes.c_void_p, c_vp]),
    "spgan_conv_gemm_rgb_slots": (c_int, [_PASS_P, c_i64]),
    "spgan_sphere_conv_gemm": (c_int, [_PASS_P, ctypes.c_void_p, ctypes.c_void_p, c_vp]),
    "spgan_upblur_pack": (c_int, [c_vp] * 7 + [c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_i64, c_int, c_f32, c_f32, c_vp]),
    "spgan_rgb_tail": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_i64, c_int, c_i64, c_vp]),
    "spgan_gemm_launch_count": (c_i64, []),
    "spgan_set_option": (c_int, [c_int, c_int]),
    "spgan_mapping_chain": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_int, c_f32, c_f32, c_f32, c_f32, c_vp]),
    "spgan_modulation_layer_bytes": (c_int, []),
    "spgan_modulation_batch": (c_int, [c_vp, c_vp, c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_vp]),
    "spgan_ema_chunk_elems": (c_int, []),
    "spgan_ema_multi": (c_int, [c_vp, c_int, c_f32, c_f32, c_vp]),
    "spgan_minibatch_stddev": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp]),
}

_lib = None
_lock = threading.Lock()
_launches = 0  # kernel launches issued through this binding (bench.py reports it as `gpu_launches`)


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libspgan_b200.so is missing (%s). Build it with `python __graft_entry__.py build` or "
                "`python sp-gan-tip2025_b200/csrc/build.py`; this package has no CPU/PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.spgan_abi_version() != 2:
            raise RuntimeError("libspgan_b200.so ABI version %d, expected 2" % lib.spgan_abi_version())
        _lib = lib
    return _lib


def last_error():
    return load().spgan_last_error().decode("utf-8", "replace")


_hook = None  # optional profiler: _hook(name, None) -> token before the call, _hook(name, token) after it


def set_call_hook(fn):
    global _hook
    _hook = fn


def call(name, *args):
    """Invoke an int-returning entry point; non-zero -> RuntimeError carrying spgan_last_error()."""
    global _launches
    tok = _hook(name, None) if _hook is not None else None
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (code %d): %s" % (name, rc, last_error()))
    if _hook is not None:
        _hook(name, tok)
    _launches += 1


def launches():
    return _launches


def require_device():
    """Fail loudly unless the current CUDA device is a B200-class (sm_100) GPU."""
    if load().spgan_device_ok() != 1:
        raise RuntimeError("spgan_b200 needs an sm_100a GPU: " + last_error())
