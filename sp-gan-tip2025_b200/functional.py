"""Autograd-level operators over the C ABI (include/spgan_b200.h).

Every function here launches kernels of libspgan_b200.so on the current CUDA stream of the input tensors' device.
torch supplies device memory, streams and the autograd graph; none of the arithmetic of the path is done by torch
kernels in the forward direction.  The conv operator is written as a (bi)linear map with its adjoint so that
first- and second-order gradients (R1 through D, path-length regularisation through G; reference
models/losses.py:36-41, 60-78) compose out of the same three kernels, the way the reference composes them out of
cuDNN fwd / dgrad / wgrad.
"""
import ctypes
import math
from collections import OrderedDict
from dataclasses import dataclass

import torch

from . import lib
from .lib import ConvPass

# 0 = exact fp32 SIMT, 1 = bf16x3 split on tcgen05 (fp32-equivalent, default), 2 = plain bf16 on tcgen05,
# 3 = fp16x2 on tcgen05 (fp16 hi+lo activations x fp16 weights: 2 MMAs per product, relative error ~2^-12; forward only —
# gradients keep the bf16x3 split because their range does not fit fp16; activations beyond 65504 saturate, so callers
# pre-scale by a power of two, see TextureSynthesizer.act_scale)
_PRECISION = 1


def set_precision(p):
    global _PRECISION
    if p not in (0, 1, 2, 3):
        raise ValueError("precision must be 0 (fp32 SIMT), 1 (bf16x3 tcgen05), 2 (bf16 tcgen05) or 3 (fp16x2 tcgen05)")
    _PRECISION = p


def _fmt(precision):
    """Activation operand format of a precision mode: 0 = bf16 hi/lo planes, 1 = fp16 hi/lo planes."""
    return 1 if precision == 3 else 0


def _wfmt(precision):
    """Weight operand format of a precision mode."""
    return 1 if precision == 3 else 0


def get_precision():
    return _PRECISION


def set_gemm_pair_mode(mode):
    """0 = never use the CTA-pair (cta_group::2) GEMM kernel, 1 = where its tiling fills the SMs as well (default), 2 = wherever
    legal (Cout % 256 == 0).  Results are bit-identical either way (same products, same K order)."""
    if lib.load().spgan_set_option(1, int(mode)) != 0:  # not a kernel launch: bypass lib.call's launch counter
        raise RuntimeError("spgan_set_option failed: " + lib.last_error())


_PROFILE = None  # list of (start event, end event, algorithmic flops) while bench.py profiles the tcgen05 launches


def profile_gemm(on):
    """bench.py: bracket every tcgen05 GEMM launch with CUDA events on the launching stream.  profile_gemm(True) starts
    recording; profile_gemm(False) returns {"ms", "flops", "launches"} summed over the recorded launches."""
    global _PROFILE
    if on:
        _PROFILE = []
        return None
    rec, _PROFILE = _PROFILE or [], None
    torch.cuda.synchronize()
    shapes = {}
    for a, b, f, key in rec:
        ms, fl, n = shapes.get(key, (0.0, 0.0, 0))
        shapes[key] = (ms + a.elapsed_time(b), fl + f, n + 1)
    return {"ms": float(sum(v[0] for v in shapes.values())), "flops": float(sum(f for _, _, f, _ in rec)),
            "launches": len(rec), "shapes": shapes}


def profile_calls(on):
    """bench.py --profile-calls: bracket EVERY C-ABI call with CUDA events on the launching stream (warm, in-situ
    durations; ncu's per-launch times are cold-cache and serialised).  profile_calls(False) returns
    {entry point: (ms, calls)}."""
    global _CALLS
    if on:
        _CALLS = []

        def hook(name, tok):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            if tok is None:
                return e
            _CALLS.append((name, tok, e))
        lib.set_call_hook(hook)
        return None
    lib.set_call_hook(None)
    rec, _CALLS = _CALLS or [], None
    torch.cuda.synchronize()
    out = {}
    for name, a, b in rec:
        ms, n = out.get(name, (0.0, 0))
        out[name] = (ms + a.elapsed_time(b), n + 1)
    return out


_CALLS = None


def _timed_call(flops, name, *args):
    if _PROFILE is None:
        lib.call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.call(name, *args)
    e1.record()
    cp = args[0]._obj  # the SpganConvPass of this launch: label the shape for the per-layer breakdown
    key = "%s taps%d lattice%dx%dx%d cout%d %s" % (name.replace("spgan_conv_", ""), cp.ntaps, cp.B, cp.H, cp.W, cp.Cout,
                                                  "act" if cp.act else "plain")
    _PROFILE.append((e0, e1, flops, key))


def _gemm_call(flops, *args):
    _timed_call(flops, "spgan_conv_gemm", *args)


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _check_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError("%s: input must be a CUDA tensor (this package has no CPU path)" % who)
    if t.dtype != torch.float32:
        raise RuntimeError("%s: only float32 tensors are supported, got %s" % (who, t.dtype))


def _f32c(t, who):
    _check_cuda(t, who)
    return t.contiguous()


# =================================================================================================== K1 bias-act
def bias_act(x, bias=None, ref=None, act=3, grad=0, alpha=0.2, scale=2 ** 0.5):
    """fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale) — models/custom_ops/fused_bias_act.cpp:11-20.
    Empty tensors / None mean "absent", as in the reference."""
    x = _f32c(x, "fused_bias_act")
    out = torch.empty_like(x)
    b = bias if (bias is not None and bias.numel() > 0) else None
    r = ref if (ref is not None and ref.numel() > 0) else None
    step_b = 1
    if b is not None:
        b = _f32c(b, "fused_bias_act")
        for d in x.shape[2:]:
            step_b *= d
    if r is not None:
        r = _f32c(r, "fused_bias_act")
    with torch.cuda.device(x.device):
        lib.call("spgan_bias_act", _ptr(out), _ptr(x), _ptr(b), _ptr(r), x.numel(), step_b,
                 b.numel() if b is not None else 1, int(act), int(grad), float(alpha), float(scale), _stream(x))
    return out


class FusedLeakyReLUFunctionBackward(torch.autograd.Function):
    """models/custom_ops/fused_act.py:24-53 (grad_input and grad_bias in one kernel)."""

    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        go = _f32c(grad_output, "fused_leaky_relu backward")
        outc = _f32c(out, "fused_leaky_relu backward")
        grad_input = torch.empty_like(go)
        B, C = go.shape[0], go.shape[1]
        inner = go.numel() // max(B * C, 1)
        grad_bias = torch.empty(C, device=go.device, dtype=torch.float32)
        with torch.cuda.device(go.device):
            lib.call("spgan_bias_act_bwd", _ptr(grad_input), _ptr(grad_bias), _ptr(go), _ptr(outc), B, C, inner,
                     float(negative_slope), float(scale), _stream(go))
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gradgrad_out = bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(torch.autograd.Function):
    """models/custom_ops/fused_act.py:56-75."""

    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
        return grad_input, grad_bias, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """models/custom_ops/fused_act.py:91-101 (CUDA branch)."""
    _check_cuda(input, "fused_leaky_relu")
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)


def noise_bias_act(x, noise, noise_weight, bias, negative_slope=0.2, scale=2 ** 0.5):
    """NoiseInjection (models/ops.py:784) + FusedLeakyReLU in one pass; inference-only fusion (no autograd)."""
    x = _f32c(x, "noise_bias_act")
    B, C = x.shape[0], x.shape[1]
    inner = x.numel() // max(B * C, 1)
    out = torch.empty_like(x)
    nz = _f32c(noise, "noise_bias_act") if noise is not None else None
    if nz is not None and nz.numel() != B * inner:
        raise RuntimeError("noise_bias_act: noise must have shape (B, 1, H, W) matching the input")
    with torch.cuda.device(x.device):
        lib.call("spgan_noise_bias_act", _ptr(out), _ptr(x), _ptr(nz), _ptr(noise_weight) if nz is not None else _ptr(None),
                 _ptr(bias), B, C, inner, float(negative_slope), float(scale), _stream(x))
    return out


def upblur_act(pp, kernel, out_hw, noise, noise_weight, bias, negative_slope=0.2, scale=2 ** 0.5):
    """Fused tail of the upsampling StyledConv on polyphase planes pp (B, C, 4, Hq, Wq) -> (B, C, zh-2, zw-2):
    interleave + 3x3 FIR (Blur, pad 0) + noise + bias + leaky-ReLU (see spgan_upblur_act).  (zh, zw) = out_hw."""
    B, C, _, Hq, Wq = pp.shape
    zh, zw = out_hw
    out = torch.empty((B, C, zh - 2, zw - 2), device=pp.device, dtype=torch.float32)
    nz = _f32c(noise, "upblur_act") if noise is not None else None
    if nz is not None and nz.numel() != B * (zh - 2) * (zw - 2):
        raise RuntimeError("upblur_act: noise must have shape (B, 1, %d, %d)" % (zh - 2, zw - 2))
    if tuple(kernel.shape) != (3, 3):
        raise RuntimeError("upblur_act: only the 3x3 blur kernel is fused")
    with torch.cuda.device(pp.device):
        lib.call("spgan_upblur_act", _ptr(out), _ptr(pp), _ptr(_f32c(kernel, "upblur_act")), _ptr(nz),
                 _ptr(noise_weight) if nz is not None else _ptr(None), _ptr(bias), B, C, zh, zw, Hq, Wq,
                 float(negative_slope), float(scale), _stream(pp))
    return out


# =================================================================================================== K2/K3 upfirdn2d
def _upfirdn2d_raw(x4, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """x4: (planes..., H, W) contiguous; returns (planes, out_h, out_w)."""
    in_h, in_w = x4.shape[-2], x4.shape[-1]
    planes = x4.numel() // max(in_h * in_w, 1)
    kh, kw = kernel.shape
    out_h = (in_h * up_y + py0 + py1 - kh) // down_y + 1
    out_w = (in_w * up_x + px0 + px1 - kw) // down_x + 1
    out = torch.empty((planes, max(out_h, 0), max(out_w, 0)), device=x4.device, dtype=torch.float32)
    with torch.cuda.device(x4.device):
        lib.call("spgan_upfirdn2d", _ptr(out), _ptr(x4), _ptr(kernel), planes, in_h, in_w, kh, kw, up_x, up_y, down_x,
                 down_y, px0, px1, py0, py1, _stream(x4))
    return out


class UpFirDn2dBackward(torch.autograd.Function):
    """models/custom_ops/upfirdn2d.py:24-90."""

    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        up_x, up_y = up
        down_x, down_y = down
        g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1 = g_pad
        go = _f32c(grad_output, "upfirdn2d backward").reshape(-1, out_size[0], out_size[1])
        grad_input = _upfirdn2d_raw(go, grad_kernel, down_x, down_y, up_x, up_y, g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        grad_input = grad_input.view(in_size[0], in_size[1], in_size[2], in_size[3])
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        up_x, up_y = ctx.up
        down_x, down_y = ctx.down
        px0, px1, py0, py1 = ctx.pad
        ggi = _f32c(gradgrad_input, "upfirdn2d grad-grad").reshape(-1, ctx.in_size[2], ctx.in_size[3])
        ggo = _upfirdn2d_raw(ggi, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1)
        ggo = ggo.view(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0], ctx.out_size[1])
        return ggo, None, None, None, None, None, None, None, None


class UpFirDn2d(torch.autograd.Function):
    """models/custom_ops/upfirdn2d.py:93-147."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        batch, channel, in_h, in_w = input.shape
        ctx.in_size = input.shape
        x = _f32c(input, "upfirdn2d").reshape(-1, in_h, in_w)
        kernel = _f32c(kernel, "upfirdn2d")
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out_h = (in_h * up_y + pad_y0 + pad_y1 - kernel_h) // down_y + 1
        out_w = (in_w * up_x + pad_x0 + pad_x1 - kernel_w) // down_x + 1
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        g_pad_x0 = kernel_w - pad_x0 - 1
        g_pad_y0 = kernel_h - pad_y0 - 1
        g_pad_x1 = in_w * up_x - out_w * down_x + pad_x0 - up_x + 1
        g_pad_y1 = in_h * up_y - out_h * down_y + pad_y0 - up_y + 1
        ctx.g_pad = (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        out = _upfirdn2d_raw(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        return out.view(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad,
                                             ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """models/custom_ops/upfirdn2d.py:150-161 (CUDA branch)."""
    _check_cuda(input, "upfirdn2d")
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


# =================================================================================================== L7 linear
class _LinearFn(torch.autograd.Function):
    """y = x (W * w_scale)^T + bias * b_scale  (models/ops.py:213-218 without the activation)."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_scale, b_scale):
        xc, wc = _f32c(x, "equal_linear"), _f32c(weight, "equal_linear")
        ctx.save_for_backward(xc, wc)
        ctx.w_scale, ctx.b_scale, ctx.has_bias = w_scale, b_scale, bias is not None
        return _linear_raw(xc, wc, bias, w_scale, b_scale, 0, 0.0, 1.0)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _LinearFn.apply(g, w.t(), None, ctx.w_scale, 1.0)
        if ctx.needs_input_grad[1] and not _ONLY_DATA_GRADS:
            gw = _LinearWgradFn.apply(g, x, ctx.w_scale)
        if ctx.has_bias and ctx.needs_input_grad[2] and not _ONLY_DATA_GRADS:
            gb = g.sum(0) * ctx.b_scale
        return gx, gw, gb, None, None


class _LinearWgradFn(torch.autograd.Function):
    """dw = scale * g^T x (bilinear in g and x, so its own backward is two more linear maps: second-order through the
    discriminator's and the mapping network's linear layers, models/losses.py:36-41, 60-78)."""

    @staticmethod
    def forward(ctx, g, x, scale):
        gc, xc = _f32c(g, "equal_linear wgrad"), _f32c(x, "equal_linear wgrad")
        ctx.save_for_backward(gc, xc)
        ctx.scale = scale
        M, N = gc.shape
        K = xc.shape[1]
        dw = torch.empty((N, K), device=gc.device, dtype=torch.float32)
        with torch.cuda.device(gc.device):
            lib.call("spgan_linear_wgrad", _ptr(dw), _ptr(gc), _ptr(xc), M, N, K, float(scale), _stream(gc))
        return dw

    @staticmethod
    def backward(ctx, ggw):
        g, x = ctx.saved_tensors
        gg = gx = None
        if ctx.needs_input_grad[0]:  # d/dg: scale * x ggw^T  -> (M, N)
            gg = _LinearFn.apply(x, ggw, None, ctx.scale, 1.0)
        if ctx.needs_input_grad[1]:  # d/dx: scale * g ggw    -> (M, K)
            gx = _LinearFn.apply(g, ggw.t(), None, ctx.scale, 1.0)
        return gg, gx, None


def _linear_raw(x, w, bias, w_scale, b_scale, act, alpha, gain):
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), device=x.device, dtype=torch.float32)
    b = _f32c(bias, "equal_linear") if bias is not None else None
    with torch.cuda.device(x.device):
        lib.call("spgan_linear", _ptr(y), _ptr(x), _ptr(w), _ptr(b), M, N, K, float(w_scale), float(b_scale), int(act),
                 float(alpha), float(gain), _stream(x))
    return y


def equal_linear(x, weight, bias, scale, lr_mul=1.0, activation=False):
    """EqualLinear.forward (models/ops.py:213-218).  Under no_grad the bias + leaky-ReLU epilogue is fused."""
    _check_cuda(x, "equal_linear")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if not torch.is_grad_enabled() or not (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)):
        y = _linear_raw(_f32c(x2, "equal_linear"), _f32c(weight, "equal_linear"), bias, scale, lr_mul,
                        1 if activation else 0, 0.2, 2 ** 0.5)
    elif activation:
        y = _LinearFn.apply(x2, weight, None, scale, 1.0)
        y = fused_leaky_relu(y, bias * lr_mul)
    else:
        y = _LinearFn.apply(x2, weight, bias, scale, lr_mul)
    return y.reshape(*lead, weight.shape[0])


# =================================================================================================== conv geometry
@dataclass(frozen=True)
class ConvGeom:
    """A 2-D convolution family member: F.conv2d(stride, padding) or F.conv_transpose2d(stride) followed by a crop of
    `crop` pixels per side (models/ops.py:617-619).  Weights are always indexed (O, C, kh, kw)."""
    kh: int
    kw: int
    stride: int = 1
    pad: int = 0
    transposed: bool = False
    crop: int = 0

    def out_size(self, H, W):
        if self.transposed:
            return (H - 1) * self.stride + self.kh - 2 * self.crop, (W - 1) * self.stride + self.kw - 2 * self.crop
        return (H + 2 * self.pad - self.kh) // self.stride + 1, (W + 2 * self.pad - self.kw) // self.stride + 1


def _ceil_div(a, b):
    return -(-a // b)


def plan_passes(geom, adjoint, in_hw, out_hw):
    """Decompose the linear map (base conv, or its adjoint = data gradient) into lattice passes.
    Returns a list of dicts {My, Mx, in_stride, out_stride, off_y, off_x, taps=[(dy, dx, widx)]}; `covers` tells
    whether the passes write every output element (otherwise the caller zero-fills first)."""
    s = geom.stride
    kh, kw = geom.kh, geom.kw
    H_out, W_out = out_hw
    passes = []
    covers = True
    if not adjoint and not geom.transposed:
        taps = [(ky - geom.pad, kx - geom.pad, ky * kw + kx) for ky in range(kh) for kx in range(kw)]
        passes.append(dict(My=H_out, Mx=W_out, in_stride=s, out_stride=1, off_y=0, off_x=0, taps=taps))
    elif adjoint and geom.transposed:
        taps = [(ky - geom.crop, kx - geom.crop, ky * kw + kx) for ky in range(kh) for kx in range(kw)]
        passes.append(dict(My=H_out, Mx=W_out, in_stride=s, out_stride=1, off_y=0, off_x=0, taps=taps))
    else:
        for ay in range(s):
            for ax in range(s):
                if not adjoint:  # transposed conv forward: Y = s*i + ky - crop
                    ys = [(-(ky - geom.crop - ay) // s, ky) for ky in range(kh) if (ky - geom.crop - ay) % s == 0]
                    xs = [(-(kx - geom.crop - ax) // s, kx) for kx in range(kw) if (kx - geom.crop - ax) % s == 0]
                else:  # data gradient of a strided conv: Y = s*i + ky - pad
                    ys = [((ay - ky + geom.pad) // s, ky) for ky in range(kh) if (ay - ky + geom.pad) % s == 0]
                    xs = [((ax - kx + geom.pad) // s, kx) for kx in range(kw) if (ax - kx + geom.pad) % s == 0]
                My, Mx = _ceil_div(H_out - ay, s), _ceil_div(W_out - ax, s)
                if My <= 0 or Mx <= 0:
                    continue
                if not ys or not xs:
                    covers = False
                    continue
                taps = [(dy, dx, ky * kw + kx) for dy, ky in ys for dx, kx in xs]
                passes.append(dict(My=My, Mx=Mx, in_stride=1, out_stride=s, off_y=ay, off_x=ax, taps=taps))
    return passes, covers


def _fill_pass(p, B, Cin, H, W, Cout, out_H, out_W, ws_o, ws_c, out_scale, act, alpha, gain, precision, cstride=0):
    cp = ConvPass()
    cp.out_cstride = cstride
    cp.B, cp.Cin, cp.H, cp.W = B, Cin, H, W
    cp.Cout, cp.out_H, cp.out_W = Cout, out_H, out_W
    cp.My, cp.Mx = p["My"], p["Mx"]
    cp.in_stride, cp.out_stride = p["in_stride"], p["out_stride"]
    cp.out_off_y, cp.out_off_x = p["off_y"], p["off_x"]
    taps = p["taps"]
    if len(taps) > lib.MAX_TAPS:
        raise RuntimeError("conv: %d taps exceed the supported %d" % (len(taps), lib.MAX_TAPS))
    cp.ntaps = len(taps)
    for t, (dy, dx, wi) in enumerate(taps):
        cp.tap_dy[t], cp.tap_dx[t], cp.tap_w[t] = dy, dx, wi
    cp.ws_o, cp.ws_c = ws_o, ws_c
    cp.out_scale = out_scale
    cp.act, cp.act_alpha, cp.act_gain = act, alpha, gain
    cp.precision = precision
    return cp


# --------------------------------------------------------------------------------------------------- weight packing
_WEIGHT_CACHE = OrderedDict()
_WEIGHT_CACHE_MAX = 512
_WEIGHT_CACHE_MAX_BYTES = 4 << 30
_WEIGHT_CACHE_BYTES = 0


def _round_up(a, b):
    return _ceil_div(a, b) * b


def _weight_cache_get(key, version):
    hit = _WEIGHT_CACHE.get(key)
    if hit is not None and hit[2] == version:
        _WEIGHT_CACHE.move_to_end(key)
        return hit[0]
    return None


def _weight_cache_put(key, out, w):
    """One entry per (weight storage, geometry): a new parameter version OVERWRITES the stale pack instead of sitting next to it
    (eager training bumps the version every optimizer step).  Bounded by bytes as well as by entries."""
    global _WEIGHT_CACHE_BYTES
    old = _WEIGHT_CACHE.pop(key, None)
    if old is not None:
        _WEIGHT_CACHE_BYTES -= old[0].numel() * old[0].element_size()
    _WEIGHT_CACHE[key] = (out, w, w._version)  # keep `w` alive so the data_ptr cannot be recycled while the entry exists
    _WEIGHT_CACHE_BYTES += out.numel() * out.element_size()
    while _WEIGHT_CACHE and (len(_WEIGHT_CACHE) > _WEIGHT_CACHE_MAX or _WEIGHT_CACHE_BYTES > _WEIGHT_CACHE_MAX_BYTES):
        _, ev = _WEIGHT_CACHE.popitem(last=False)
        _WEIGHT_CACHE_BYTES -= ev[0].numel() * ev[0].element_size()


def _packed_weight(w, Cout, Cin, ws_o, ws_c, tap_w, Cp, merged, fmt=0):
    """16-bit hi/lo copy of the weight in [2][ntaps][Cout][Cp] (or merged-K) order, cached per weight version."""
    key = (w.data_ptr(), tuple(w.shape), Cout, Cin, ws_o, ws_c, tuple(tap_w), Cp, merged, fmt, w.device.index)
    hit = _weight_cache_get(key, w._version)
    if hit is not None:
        return hit
    ntaps = len(tap_w)
    out = torch.empty((2, ntaps, Cout, Cp), device=w.device, dtype=torch.bfloat16)
    arr = (ctypes.c_int32 * ntaps)(*tap_w)
    with torch.cuda.device(w.device):
        lib.call("spgan_pack_weight", _ptr(out), _ptr(w), Cout, Cin, ws_o, ws_c, ntaps, arr, Cp, int(merged), int(fmt), _stream(w))
    _weight_cache_put(key, out, w)
    return out


def clear_weight_cache():
    global _WEIGHT_CACHE_BYTES
    _WEIGHT_CACHE.clear()
    _WEIGHT_CACHE_BYTES = 0


_EPOCH = 0
_STYLE_EPOCH = 0


def memo_by_tensor(store, slot, t, build, extra=None):
    """build() memoised in the dict `store` per (slot, identity of tensor `t`, extra).  The entry keeps `t` alive, so its id
    cannot be recycled.  Several entries per slot: a panorama engine that runs lattice positions in groups of different sizes
    presents one style tensor per size, alternately (a single-entry memo would be recomputed on whichever stream comes next,
    and its readers on other streams would race with that)."""
    key = (slot, id(t), extra)
    hit = store.get(key)
    if hit is not None and hit[0] is t:
        return hit[1]
    while len(store) >= 64:
        del store[next(iter(store))]  # oldest first: never the entries of the run in flight
    val = build()
    store[key] = (t, val)
    return val


def epoch():
    return (_EPOCH, _STYLE_EPOCH)


def bump_style_epoch():
    """Invalidate only the memoised (modulation, demodulation) pairs: a CUDA-graph capture must contain the kernels that
    compute them (a memo hit during capture would freeze the styles of the capture-time latent into every replay)."""
    global _STYLE_EPOCH
    _STYLE_EPOCH += 1


def bump_epoch():
    """Invalidate everything memoised on parameter versions (packed weights here, the (s, d) modulation pairs of
    ModulatedConv2d).  A CUDA-graph replay updates parameters on the device without bumping their Python `_version`, so
    the training step calls this after every replay and before every capture."""
    global _EPOCH
    _EPOCH += 1
    clear_weight_cache()


# --------------------------------------------------------------------------------------------------- conv driver
def _tensor_path_ok(passes, Cin, Cout, precision):
    return precision != 0 and Cin >= 16 and Cout >= 16 and len({p["in_stride"] for p in passes}) == 1


def _phase_taps(passes):
    """Map every tap of every pass onto the polyphase lattice of the input: with s = in_stride and the taps shifted by
    (pt, pl) to be non-negative, input row i*s + dy + pt = s*(i + (dy+pt)//s) + (dy+pt)%s.  Returns
    (s, pt, pl, Hl, Wl, per-pass [(phase, oy, ox, widx)]) where (Hl, Wl) is the smallest lattice on which no tap of a valid
    lattice point wraps around a row."""
    s = passes[0]["in_stride"]
    dy_all = [t[0] for p in passes for t in p["taps"]]
    dx_all = [t[1] for p in passes for t in p["taps"]]
    pt, pl = max(0, -min(dy_all)), max(0, -min(dx_all))
    mapped, Hl, Wl = [], 1, 1
    for p in passes:
        cur = []
        for dy, dx, wi in p["taps"]:
            vy, vx = dy + pt, dx + pl
            cur.append(((vy % s) * s + (vx % s), vy // s, vx // s, wi))
        mapped.append(cur)
        Hl = max(Hl, p["My"] + max(t[1] for t in cur))
        Wl = max(Wl, p["Mx"] + max(t[2] for t in cur))
    return s, pt, pl, Hl, Wl, mapped


def conv_apply(x, w, geom, adjoint=False, out_hw=None, in_mul=None, out_mul=None, out_scale=1.0, noise=None,
               noise_w=None, bias=None, act=None, residual=None, precision=None, polyphase=False, k_round=64):
    """y = [act]( out_scale * out_mul[b,o] * L_w(in_mul[b,c] * x) + noise_w*noise + bias ) + residual, no autograd.

    L_w is the conv described by `geom` (adjoint=False) or its adjoint / data gradient (adjoint=True, `out_hw`
    required).  w is (O, C, kh, kw).  act = (alpha, gain) or None."""
    x = _f32c(x, "conv")
    w = _f32c(w, "conv")
    B, Cx, H, W = x.shape
    O, C = w.shape[0], w.shape[1]
    kk = geom.kh * geom.kw
    if tuple(w.shape[2:]) != (geom.kh, geom.kw):
        raise RuntimeError("conv: weight %s does not match a %dx%d kernel" % (tuple(w.shape), geom.kh, geom.kw))
    if not adjoint:
        if Cx != C:
            raise RuntimeError("conv: input has %d channels, weight expects %d" % (Cx, C))
        Cin, Cout, ws_o, ws_c = C, O, C * kk, kk
        oh, ow = geom.out_size(H, W)
    else:
        if Cx != O:
            raise RuntimeError("conv adjoint: input has %d channels, weight expects %d" % (Cx, O))
        Cin, Cout, ws_o, ws_c = O, C, kk, C * kk
        oh, ow = out_hw
    precision = _PRECISION if precision is None else precision
    if precision >= 3 and adjoint:
        precision = 1  # data gradients: fp16 has too little range
    fmt = _fmt(precision)
    passes, covers = plan_passes(geom, adjoint, (H, W), (oh, ow))
    if polyphase:
        # parity passes write dense planes (B, Cout, s*s, Hq, Wq), plane (off_y * s + off_x), instead of scattering
        # with stride s into the interleaved image: coalesced epilogue stores; the consumer interleaves on the fly
        sp = geom.stride
        if not (geom.transposed and not adjoint and sp == 2) or any(t is not None for t in (noise, bias, act, residual)):
            raise RuntimeError("conv: polyphase output is for the bare stride-2 transposed conv only")
        Hq, Wq = _ceil_div(oh, sp), _ceil_div(ow, sp)
        y = (torch.empty if covers else torch.zeros)((B, Cout, sp * sp, Hq, Wq), device=x.device, dtype=torch.float32)
        if y.numel() == 0 or not passes:
            return y
    else:
        y = (torch.empty if covers else torch.zeros)((B, Cout, max(oh, 0), max(ow, 0)), device=x.device, dtype=torch.float32)
        if y.numel() == 0 or not passes:
            return y
    a, g = (act if act is not None else (0.0, 1.0))
    act_on = 1 if act is not None else 0
    im = _f32c(in_mul, "conv") if in_mul is not None else None
    om = _f32c(out_mul, "conv") if out_mul is not None else None
    nz = _f32c(noise, "conv") if noise is not None else None
    if nz is not None:
        if noise_w is None:
            raise RuntimeError("conv: noise needs noise_w")
        want = (B, 1, max(oh, 0), max(ow, 0))
        if tuple(nz.shape) != want:
            # the reference's `image + weight * noise` broadcasts: accept what broadcasts to (B, 1, oh, ow), reject the rest
            # (the kernel indexes noise[b*oh*ow + pixel]; a smaller tensor would be read out of bounds)
            try:
                nz = nz.expand(want).contiguous()
            except RuntimeError:
                raise RuntimeError("conv: noise of shape %s does not broadcast to %s" % (tuple(nz.shape), want))
    nwt = _f32c(noise_w, "conv") if (noise is not None) else None
    bs = _f32c(bias, "conv") if bias is not None else None
    rs = _f32c(residual, "conv") if residual is not None else None
    if rs is not None and rs.shape != y.shape:
        raise RuntimeError("conv: residual shape %s != output shape %s" % (tuple(rs.shape), tuple(y.shape)))
    st = _stream(x)

    def target(p):
        """(pass geometry, output pointer, out_H, out_W, channel stride) for the dense or the polyphase layout."""
        if not polyphase:
            return p, _ptr(y), oh, ow, 0
        plane = (p["off_y"] * geom.stride + p["off_x"]) * Hq * Wq
        q = dict(p, out_stride=1, off_y=0, off_x=0)
        return q, ctypes.c_void_p(y.data_ptr() + 4 * plane), Hq, Wq, geom.stride * geom.stride * Hq * Wq

    with torch.cuda.device(x.device):
        if _tensor_path_ok(passes, Cin, Cout, precision):
            # one packed activation shared by all passes: pads cover every pass's tap reach; a strided conv reads the
            # polyphase planes of the input (spgan_pack_act with step = stride)
            step, pt, pl, Hl, Wl, mapped = _phase_taps(passes)
            if step == 1:
                Hl, Wl = max(Hl, pt + H), max(Wl, pl + W)
            # K per tap padded to whole 64-wide blocks.  The kernel accepts any multiple of 16 (partial last block), but
            # measured on the 259-channel 7x7 layers a 16-wide last block is slower than padding to 320: it saves 15 % of
            # the MMAs yet leaves a pipeline bubble per tap (its stage holds 3 MMAs, not enough to cover the next TMA load)
            Cp = _round_up(Cin, k_round)
            rows = B * Hl * Wl
            a_packed = torch.empty((2, step * step * rows, Cp), device=x.device, dtype=torch.bfloat16)
            lib.call("spgan_pack_act", _ptr(a_packed), _ptr(x), _ptr(im), B, Cin, H, W, Cp, pt, pl, Hl, Wl, step, fmt, st)
            for p, taps in zip(passes, mapped):
                q, yptr, o_h, o_w, cst = target(p)
                shifted = dict(q, in_stride=1, taps=[(ph * B * Hl + oy, ox, wi) for ph, oy, ox, wi in taps])
                cp = _fill_pass(shifted, B, Cin, Hl, Wl, Cout, o_h, o_w, ws_o, ws_c, out_scale, act_on, a, g, precision, cst)
                wp = _packed_weight(w, Cout, Cin, ws_o, ws_c, [t[2] for t in p["taps"]], Cp, False, _wfmt(precision))
                valid = min(p["My"], _ceil_div(oh - p["off_y"], p["out_stride"])) * min(p["Mx"], _ceil_div(ow - p["off_x"], p["out_stride"]))
                _gemm_call(2.0 * B * valid * Cout * Cin * len(p["taps"]), ctypes.byref(cp), yptr, _ptr(a_packed),
                           step * step * rows, Cp, _ptr(wp), _ptr(om), _ptr(nz), _ptr(nwt), _ptr(bs), _ptr(rs), st)
        else:
            for p in passes:
                q, yptr, o_h, o_w, cst = target(p)
                cp = _fill_pass(q, B, Cin, H, W, Cout, o_h, o_w, ws_o, ws_c, out_scale, act_on, a, g, 0, cst)
                lib.call("spgan_conv_pass", ctypes.byref(cp), yptr, _ptr(x), _ptr(w), _ptr(im), _ptr(om), _ptr(nz),
                         _ptr(nwt), _ptr(bs), _ptr(rs), st)
    return y


def conv_wgrad(g, x, w_shape, geom, in_mul=None, out_mul=None, out_scale=1.0, precision=None):
    """dw[o,c,ky,kx] = out_scale * sum_b <out_mul*g, L_{e(o,c,ky,kx)}(in_mul*x)> for the BASE conv of `geom`
    (x = base input (B, C, H, W), g = base output gradient (B, O, oh, ow))."""
    g = _f32c(g, "conv wgrad")
    x = _f32c(x, "conv wgrad")
    O, C, kh, kw = w_shape
    B, _, H, W = x.shape
    oh, ow = g.shape[2], g.shape[3]
    kk = kh * kw
    dw = torch.zeros(w_shape, device=x.device, dtype=torch.float32)
    if g.numel() == 0 or x.numel() == 0:
        return dw
    passes, _ = plan_passes(geom, False, (H, W), (oh, ow))
    im = _f32c(in_mul, "conv wgrad") if in_mul is not None else None
    om = _f32c(out_mul, "conv wgrad") if out_mul is not None else None
    precision = _PRECISION if precision is None else precision
    if precision >= 3:
        precision = 1  # weight gradients: fp16 has too little range
    if passes and _tensor_path_ok(passes, C, O, precision) and len({p["out_stride"] for p in passes}) == 1:
        # tcgen05: contraction over the flattened lattice rows of the same channels-last packs the forward GEMM reads
        s_in, pt, pl, Hl, Wl, mapped = _phase_taps(passes)
        s_out = passes[0]["out_stride"]
        if s_in == 1:
            Hl, Wl = max(Hl, pt + H), max(Wl, pl + W)
        Q = B * Hl * Wl
        Op, Cp = _round_up(O, 64), _round_up(C, 64)
        st = _stream(x)
        with torch.cuda.device(x.device):
            gp = torch.empty((2, s_out * s_out * Q, Op), device=x.device, dtype=torch.bfloat16)
            xp = torch.empty((2, s_in * s_in * Q, Cp), device=x.device, dtype=torch.bfloat16)
            lib.call("spgan_pack_act", _ptr(gp), _ptr(g), _ptr(om), B, O, oh, ow, Op, 0, 0, Hl, Wl, s_out, 0, st)
            lib.call("spgan_pack_act", _ptr(xp), _ptr(x), _ptr(im), B, C, H, W, Cp, pt, pl, Hl, Wl, s_in, 0, st)
            for p, taps in zip(passes, mapped):
                q = dict(p, in_stride=1, out_stride=1, off_y=0, off_x=0, taps=[(oy, ox, wi) for _, oy, ox, wi in taps])
                cp = _fill_pass(q, B, C, Hl, Wl, O, Hl, Wl, C * kk, kk, out_scale, 0, 0.0, 1.0, precision)
                phases = (ctypes.c_int32 * len(taps))(*[t[0] for t in taps])
                need = int(lib.load().spgan_conv_wgrad_gemm_workspace(ctypes.byref(cp)))
                ws = torch.empty((max(need, 1),), device=x.device, dtype=torch.float32)
                flops = 2.0 * B * min(p["My"], oh) * min(p["Mx"], ow) * O * C * len(taps)
                _timed_call(flops, "spgan_conv_wgrad_gemm", ctypes.byref(cp), _ptr(dw), _ptr(gp), s_out * s_out,
                            p["off_y"] * s_out + p["off_x"], Op, _ptr(xp), s_in * s_in, phases, Cp, _ptr(ws), need, 0, st)
        return dw
    with torch.cuda.device(x.device):
        for p in passes:
            cp = _fill_pass(p, B, C, H, W, O, oh, ow, C * kk, kk, out_scale, 0, 0.0, 1.0, 0)
            lib.call("spgan_conv_wgrad", ctypes.byref(cp), _ptr(dw), _ptr(g), _ptr(x), _ptr(im), _ptr(om), 1, _stream(x))
    return dw


_ONLY_DATA_GRADS = False


class only_data_grads:
    """Context manager for a `create_graph=True` backward whose caller consumes only DATA gradients (R1: d D(x) / dx,
    models/losses.py:36-41; path length: d <G(w), noise> / dw, :60-68).  autograd evaluates every output of a custom
    backward whose input `requires_grad`, so without it each conv / linear of the network would also compute — and record
    for double backward — a weight gradient that is thrown away.  Inside the context the weight / bias gradient slots of
    this package's Functions return None.  The second backward (through the recorded graph) is unaffected."""

    def __enter__(self):
        global _ONLY_DATA_GRADS
        self.prev, _ONLY_DATA_GRADS = _ONLY_DATA_GRADS, True

    def __exit__(self, *exc):
        global _ONLY_DATA_GRADS
        _ONLY_DATA_GRADS = self.prev
        return False


def plane_dot(a, b):
    """out[b, c] = sum over pixels of a[b, c] * b[b, c] in one pass (spgan_plane_dot): the style / demodulation gradients
    d s[b,c] = <x[b,c], dxs[b,c]> without the (B, C, H, W) product tensor.  No autograd: first-order backward only."""
    ac, bc = _f32c(a, "plane_dot"), _f32c(b, "plane_dot")
    if ac.shape != bc.shape or ac.dim() != 4:
        raise RuntimeError("plane_dot: operands must be two (B, C, H, W) tensors of equal shape")
    B, C, H, W = ac.shape
    out = torch.empty((B, C), device=ac.device, dtype=torch.float32)
    if B * C:
        with torch.cuda.device(ac.device):
            lib.call("spgan_plane_dot", _ptr(out), _ptr(ac), _ptr(bc), B * C, H * W, _stream(ac))
    return out


class _ConvFn(torch.autograd.Function):
    """y = out_scale * D(out_mul) L_w( D(in_mul) x ), differentiable to any order in x, w, in_mul, out_mul
    (the weight-gradient node itself is first-order only, which is all R1 / path-length need)."""

    @staticmethod
    def forward(ctx, x, w, in_mul, out_mul, geom, adjoint, out_hw, out_scale):
        y = conv_apply(x, w, geom, adjoint, out_hw, in_mul, out_mul, out_scale)
        ctx.geom, ctx.adjoint, ctx.out_scale = geom, adjoint, out_scale
        ctx.in_hw = (x.shape[2], x.shape[3])
        ctx.has_im, ctx.has_om = in_mul is not None, out_mul is not None
        ctx.save_for_backward(x, w, in_mul, out_mul, y if out_mul is not None else None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, w, in_mul, out_mul, y = ctx.saved_tensors
        need_x, need_w, need_im, need_om = ctx.needs_input_grad[:4]
        need_w = need_w and not _ONLY_DATA_GRADS
        gx = gw = gim = gom = None
        if need_x or need_im:
            # D(in_mul)^-1 applied lazily: the un-modulated data gradient serves both dx and d(in_mul)
            gx_un = _ConvFn.apply(g, w, out_mul, None, ctx.geom, not ctx.adjoint, ctx.in_hw, ctx.out_scale)
            if need_im:
                # first-order backward (no create_graph): one fused pass instead of a (B, C, H, W) product + reduction
                gim = plane_dot(x, gx_un) if not torch.is_grad_enabled() else (x * gx_un).sum(dim=(2, 3))
            if need_x:
                gx = gx_un * in_mul[:, :, None, None] if ctx.has_im else gx_un
        if need_w:
            if not ctx.adjoint:
                gw = _WgradFn.apply(g, x, in_mul, out_mul, tuple(w.shape), ctx.geom, ctx.out_scale)
            else:  # z = L^T(g'): <z, gz> = <L(gz), g'>, so the base input is the incoming gradient
                gw = _WgradFn.apply(x, g, out_mul, in_mul, tuple(w.shape), ctx.geom, ctx.out_scale)
        if need_om:
            gom = (plane_dot(g, y) if not torch.is_grad_enabled() else (g * y).sum(dim=(2, 3))) / out_mul
        return gx, gw, gim, gom, None, None, None, None


class _WgradFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, x, in_mul, out_mul, w_shape, geom, out_scale):
        return conv_wgrad(g, x, w_shape, geom, in_mul, out_mul, out_scale)

    @staticmethod
    def backward(ctx, ggw):
        raise NotImplementedError("spgan_b200: third-order derivative through the conv weight gradient is not implemented")


def conv2d(x, w, geom, in_mul=None, out_mul=None, out_scale=1.0):
    """Differentiable modulated / plain convolution (forward of models/ops.py:617, 634; 175)."""
    _check_cuda(x, "conv2d")
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, w, in_mul, out_mul)):
        return _ConvFn.apply(x, w, in_mul, out_mul, geom, False, None, out_scale)
    return conv_apply(x, w, geom, False, None, in_mul, out_mul, out_scale)


def demod_coefficients(w, s, scale, eps=1e-8):
    """d[b,o] = rsqrt(scale^2 * sum_c s[b,c]^2 sum_t w[o,c,t]^2 + eps)  (models/ops.py:603-604), no autograd."""
    w = _f32c(w, "demod")
    s = _f32c(s, "demod")
    O, C = w.shape[0], w.shape[1]
    taps = w.numel() // max(O * C, 1)
    B = s.shape[0]
    d = torch.empty((B, O), device=w.device, dtype=torch.float32)
    with torch.cuda.device(w.device):
        lib.call("spgan_demod", _ptr(d), _ptr(s), _ptr(w), B, C, O, taps, float(scale), float(eps), _stream(w))
    return d


# =================================================================================================== spherical gather
def sphere_gather_raw(z, grid, out=None, out_bstride=None, out_coff=0, encode=False):
    """F.grid_sample(z, grid, bilinear, border, align_corners=True) for a (Bg, 3H, 3W, 2) tap grid."""
    z = _f32c(z, "sphere_gather")
    grid = _f32c(grid, "sphere_gather")
    B, C, H, W = z.shape
    if grid.dim() != 4 or grid.shape[1] != 3 * H or grid.shape[2] != 3 * W or grid.shape[3] != 2 or grid.shape[0] not in (1, B):
        raise RuntimeError("sphere_gather: grid %s does not match input %s" % (tuple(grid.shape), tuple(z.shape)))
    if out is None:
        out = torch.empty((B, C, 3 * H, 3 * W), device=z.device, dtype=torch.float32)
        out_bstride = C
    with torch.cuda.device(z.device):
        lib.call("spgan_sphere_gather", _ptr(out), _ptr(z), _ptr(grid), B, C, H, W, grid.shape[0], out_bstride, out_coff,
                 1 if encode else 0, _stream(z))
    return out


def sphere_gather_indices(grid, H, W):
    """(x0, y0, wx, wy) exactly as the gather kernels compute them (bit-exactness tests)."""
    grid = _f32c(grid, "sphere_gather_indices")
    n = grid.numel() // 2
    x0 = torch.empty(n, device=grid.device, dtype=torch.int32)
    y0 = torch.empty_like(x0)
    wx = torch.empty(n, device=grid.device, dtype=torch.float32)
    wy = torch.empty_like(wx)
    with torch.cuda.device(grid.device):
        lib.call("spgan_sphere_gather_indices", _ptr(x0), _ptr(y0), _ptr(wx), _ptr(wy), _ptr(grid), n, H, W, _stream(grid))
    shp = grid.shape[:-1]
    return x0.view(shp), y0.view(shp), wx.view(shp), wy.view(shp)


class _BlockMeanFn(torch.autograd.Function):
    """grad_in = mean over each 3x3 block * 0.1 (grid_generator.py:615-623); linear, so its own backward is the
    adjoint (each block cell receives g * 0.1 / 9)."""

    @staticmethod
    def forward(ctx, go):
        go = _f32c(go, "sphere_gather backward")
        B, C, H3, W3 = go.shape
        H, W = H3 // 3, W3 // 3
        gi = torch.empty((B, C, H, W), device=go.device, dtype=torch.float32)
        with torch.cuda.device(go.device):
            lib.call("spgan_sphere_gather_bwd", _ptr(gi), _ptr(go), B * C, H, W, _stream(go))
        return gi

    @staticmethod
    def backward(ctx, gg):
        return (gg * (0.1 / 9.0)).repeat_interleave(3, dim=2).repeat_interleave(3, dim=3)


class GridSamplerFuncNoGrad(torch.autograd.Function):
    """models/spherenet/grid_generator.py:602-623: bilinear border gather with the SURROGATE backward.
    The guarded all_reduce of :621-622 is deliberately not reproduced."""

    @staticmethod
    def forward(ctx, z, grid):
        return sphere_gather_raw(z, grid)

    @staticmethod
    def backward(ctx, grad_output):
        return _BlockMeanFn.apply(grad_output), None


def sphere_gather(z, grid):
    _check_cuda(z, "sphere_gather")
    return GridSamplerFuncNoGrad.apply(z, grid)


GRID_SAMPLE_MODES = {"bilinear_border": 0, "texture": 1, "nearest_zeros": 2}


def _grid_sample_raw(z, grid, mode, backward_to=None):
    """Forward (backward_to=None): (B, C, IH, IW) sampled at grid (Bg, OH, OW, 2) -> (B, C, OH, OW).  Backward:
    z = grad_out (B, C, OH, OW), backward_to = (IH, IW) -> grad wrt the sampled image."""
    z = _f32c(z, "grid_sample")
    grid = _f32c(grid, "grid_sample")
    B, C = z.shape[0], z.shape[1]
    OH, OW = grid.shape[1], grid.shape[2]
    if grid.dim() != 4 or grid.shape[3] != 2 or grid.shape[0] not in (1, B):
        raise RuntimeError("grid_sample: grid %s does not match input %s" % (tuple(grid.shape), tuple(z.shape)))
    with torch.cuda.device(z.device):
        if backward_to is None:
            IH, IW = z.shape[2], z.shape[3]
            out = torch.empty((B, C, OH, OW), device=z.device, dtype=torch.float32)
            lib.call("spgan_grid_sample", _ptr(out), _ptr(z), _ptr(grid), B, C, IH, IW, OH, OW, grid.shape[0], mode, _stream(z))
        else:
            IH, IW = backward_to
            if tuple(z.shape[2:]) != (OH, OW):
                raise RuntimeError("grid_sample backward: gradient %s does not match grid %s" % (tuple(z.shape), tuple(grid.shape)))
            out = torch.empty((B, C, IH, IW), device=z.device, dtype=torch.float32)
            lib.call("spgan_grid_sample_bwd", _ptr(out), _ptr(z), _ptr(grid), B, C, IH, IW, OH, OW, grid.shape[0], mode, _stream(z))
    return out


class _GridSampleFn(torch.autograd.Function):
    """Linear in z: backward = transposed scatter, double backward = the forward again (what the reference gets from
    autograd through torch.gather, grid_sample_ops.py:43-53, and from _GridSample2dBackward, grid_sample_grad_fix.py:54-88).
    The grid is a constant of the model (built from coords_partial on the host): no gradient flows to it."""

    @staticmethod
    def forward(ctx, z, grid, mode):
        ctx.save_for_backward(grid)
        ctx.mode, ctx.in_hw = mode, (z.shape[2], z.shape[3])
        return _grid_sample_raw(z, grid, mode)

    @staticmethod
    def backward(ctx, go):
        grid, = ctx.saved_tensors
        return _GridSampleBwdFn.apply(go, grid, ctx.mode, ctx.in_hw), None, None


class _GridSampleBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, go, grid, mode, in_hw):
        ctx.save_for_backward(grid)
        ctx.mode = mode
        return _grid_sample_raw(go, grid, mode, backward_to=in_hw)

    @staticmethod
    def backward(ctx, gg):
        grid, = ctx.saved_tensors
        return _GridSampleFn.apply(gg, grid, ctx.mode), None, None, None


def grid_sample(z, grid, mode="bilinear_border"):
    """The reference's other samplers with TRUE input gradients: "bilinear_border" = F.grid_sample(bilinear, border,
    align_corners=True); "texture" = grid_sample_github (grid_sample_ops.py:5-55); "nearest_zeros" = the 'nearest' sampler of
    the full-sphere convs (grid_sample_grad_fix.py:29-48)."""
    _check_cuda(z, "grid_sample")
    return _GridSampleFn.apply(z, grid, GRID_SAMPLE_MODES[mode])


# =================================================================================================== spherical conv
def sphere_modconv_fused(x, coords, grid, w, in_mul, out_mul, out_scale, act=None, flat_concat=True, precision=None,
                         residual=None, bias=None):
    """Inference path of spgan_ops_gs.ModulatedConv2d (models/spgan_ops_gs.py:791-816): the gather, coordinate
    encoding, concat, modulation and bf16 split happen in one producer kernel (`spgan_sphere_pack`) whose output
    feeds the tcgen05 GEMM; the 9x gathered fp32 tensor of the reference never exists."""
    x = _f32c(x, "sphere_modconv")
    w = _f32c(w, "sphere_modconv")
    grid = _f32c(grid, "sphere_modconv")
    B, C, H, W = x.shape
    O, Ct = w.shape[0], w.shape[1]
    nc = 0 if coords is None else coords.shape[1]
    if Ct != C + nc or tuple(w.shape[2:]) != (3, 3):
        raise RuntimeError("sphere_modconv: weight %s does not match %d + %d channels, 3x3" % (tuple(w.shape), C, nc))
    precision = _PRECISION if precision is None else precision
    if precision == 0 or O < 16:
        return _sphere_modconv_simt(x, coords, grid, w, in_mul, out_mul, out_scale, act, flat_concat, residual, bias)
    Cp = _round_up(Ct, 64)
    rows = B * H * W
    st = _stream(x)
    a, g = (act if act is not None else (0.0, 1.0))
    y = torch.empty((B, O, H, W), device=x.device, dtype=torch.float32)
    fmt = _fmt(precision)
    with torch.cuda.device(x.device):
        xh = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
        lib.call("spgan_nchw_to_nhwc", _ptr(xh), _ptr(x), B, C, H, W, st)
        cc = _f32c(coords, "sphere_modconv") if coords is not None else None
        im = _f32c(in_mul, "sphere_modconv") if in_mul is not None else None
        cmap = _sphere_chan_map(B, C, nc, Cp, bool(flat_concat), x.device)
        wp = _packed_weight(w, O, Ct, Ct * 9, 9, list(range(9)), Cp, True, fmt)
        p = dict(My=H, Mx=W, in_stride=1, out_stride=1, off_y=0, off_x=0, taps=[(0, 0, 0)])
        cp = _fill_pass(p, B, Ct, H, W, O, H, W, Ct * 9, 9, out_scale, 1 if act is not None else 0, a, g, precision)
        om = _f32c(out_mul, "sphere_modconv") if out_mul is not None else None
        bs = _f32c(bias, "sphere_modconv") if bias is not None else None
        rs = _f32c(residual, "sphere_modconv") if residual is not None else None
        if rs is not None and rs.shape != y.shape:
            raise RuntimeError("sphere_modconv: residual shape %s != output shape %s" % (tuple(rs.shape), tuple(y.shape)))
        flops = 2.0 * B * H * W * O * Ct * 9
        if FUSED_SPHERE_GATHER and grid.shape[0] == 1 and H * W >= 128:
            # the gather is the A-operand producer of the GEMM itself: no [B*H*W][9*Cp] operand in HBM
            sphere_conv_gemm(cp, flops, st, xh, cc, grid, im, cmap, C, Cp, wp, fmt, out_mul=om, bias=bs, residual=rs, y=y)
            return y
        a_packed = torch.empty((2, rows, 9 * Cp), device=x.device, dtype=torch.bfloat16)
        lib.call("spgan_sphere_pack", _ptr(a_packed), _ptr(xh), _ptr(cc), _ptr(grid), _ptr(im), _ptr(cmap), B, C, H, W,
                 grid.shape[0], Cp, fmt, st)
        _gemm_call(flops, ctypes.byref(cp), _ptr(y), _ptr(a_packed), rows, 9 * Cp, _ptr(wp),
                   _ptr(om), _ptr(None), _ptr(None), _ptr(bs), _ptr(rs), st)
    return y


# Which spherical-conv implementation the no_grad path uses.  False (default): spgan_sphere_pack -> spgan_conv_gemm (the
# packer runs at full occupancy, the [B*H*W][9*Cp] operand takes a round trip through HBM).  True: spgan_sphere_conv_gemm,
# the gather inside the GEMM's producer warps (no operand in HBM).  Measured on the B200 at B = 32, 256 + 3 -> 256 channels,
# bf16x3: 35x35 0.48 ms vs 0.84 ms, 17x17 0.14 vs 0.47 ms — the eight producer warps of the persistent one-CTA-per-SM GEMM
# issue ~96 instructions per (pixel, tap, 64 channels) at 41 % issue utilisation against a 1573-cycle MMA stage, so the
# in-kernel producer is the bottleneck (profiles/r2_ncu_sphere_gemm.txt); it stays selectable and tested.
FUSED_SPHERE_GATHER = False


def sphere_conv_gemm(cp, flops, st, xh, coords, grid, in_mul, cmap, C, Cp, wp, fmt, **sinks):
    """spgan_sphere_conv_gemm: gather-producer GEMM.  `sinks` are SpganGemmIO fields (tensors or None)."""
    sin = lib.SphereIn()
    sin.x_nhwc, sin.coords, sin.grid = xh.data_ptr(), (coords.data_ptr() if coords is not None else None), grid.data_ptr()
    sin.in_mul, sin.chan_map = (in_mul.data_ptr() if in_mul is not None else None), cmap.data_ptr()
    sin.C, sin.Cp = C, Cp
    io = lib.GemmIO()
    io.kp, io.fmt, io.w_fmt, io.w_packed = 9 * Cp, fmt, fmt, wp.data_ptr()
    for k, v in sinks.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        setattr(io, k, v)
    _timed_call(flops, "spgan_sphere_conv_gemm", ctypes.byref(cp), ctypes.byref(sin), ctypes.byref(io), st)


_CHAN_MAPS = {}


def _sphere_chan_map(B, C, nc, Cp, flat_concat, device, group=None):
    """(B, Cp) uint32 table for spgan_sphere_pack: which gathered plane feeds channel k of group g.  With flat_concat
    the reference's (1, B*C) ++ (1, B*nc) concatenation under groups=B is reproduced (models/spgan_ops_gs.py:792-814):
    group g reads flat channels [g*Ct, (g+1)*Ct).  `group` (a divisor of B): the batch is a stack of independent generator
    calls of `group` samples each (grids.PositionGroup); the table is then block-diagonal, every block the table of one
    call."""
    group = B if group is None else int(group)
    key = (B, C, nc, Cp, flat_concat, str(device), group)
    t = _CHAN_MAPS.get(key)
    if t is not None:
        return t
    import numpy as np
    if group <= 0 or B % group:
        raise RuntimeError("sphere channel map: group %d does not divide the batch %d" % (group, B))
    Ct = C + nc
    m = np.full((group, Cp), 0xFFFFFFFF, dtype=np.uint64)
    g = np.arange(group)[:, None]
    k = np.arange(Ct)[None, :]
    if flat_concat:
        flat = g * Ct + k
        feat = flat < group * C
        bs = np.where(feat, flat // max(C, 1), (flat - group * C) // max(nc, 1))
        cs = np.where(feat, flat % max(C, 1), (flat - group * C) % max(nc, 1))
    else:
        feat = np.broadcast_to(k < C, (group, Ct))
        bs = np.broadcast_to(g, (group, Ct))
        cs = np.where(feat, k, k - C)
    blocks = []
    for i in range(B // group):
        blk = m.copy()
        blk[:, :Ct] = (np.where(feat, 0, 1).astype(np.uint64) << 31) | ((bs + i * group).astype(np.uint64) << 15) | cs.astype(np.uint64)
        blocks.append(blk)
    t = torch.from_numpy(np.concatenate(blocks, 0).astype(np.uint32).view(np.int32)).to(device)
    _CHAN_MAPS[key] = t
    return t


def sphere_concat_gather(x, coords, grid):
    """The reference's gathered + concatenated conv input, (1, B*C + B*nc, 3H, 3W) flat (models/spgan_ops_gs.py:791-813),
    coordinate channels encoded.  No autograd."""
    B, C, H, W = x.shape
    nc = 0 if coords is None else coords.shape[1]
    flat = torch.empty((1, B * (C + nc), 3 * H, 3 * W), device=x.device, dtype=torch.float32)
    sphere_gather_raw(x, grid, out=flat, out_bstride=C, out_coff=0)
    if nc:
        sphere_gather_raw(coords, grid, out=flat, out_bstride=nc, out_coff=B * C, encode=True)
    return flat


_SPHERE_GEOM = ConvGeom(3, 3, stride=3, pad=0)


def _sphere_modconv_simt(x, coords, grid, w, in_mul, out_mul, out_scale, act, flat_concat, residual, bias):
    B, C, H, W = x.shape
    Ct = w.shape[1]
    if flat_concat:
        inp = sphere_concat_gather(x, coords, grid).view(B, Ct, 3 * H, 3 * W)
    else:
        gx = sphere_gather_raw(x, grid)
        inp = gx if coords is None else torch.cat([gx, sphere_gather_raw(coords, grid, encode=True)], 1)
    return conv_apply(inp, w, _SPHERE_GEOM, False, None, in_mul, out_mul, out_scale, None, None, bias, act, residual, 0)


def encode_coords(c):
    """tanh / cos(pi.) / sin(pi.) on channels 0/1/2 (models/spgan_ops_gs.py:799-802); differentiable torch glue used
    only on the 3-channel coordinate planes of the training path."""
    return torch.stack([torch.tanh(c[:, 0]), torch.cos(c[:, 1] * math.pi), torch.sin(c[:, 2] * math.pi)], 1)


def sphere_modconv(x, coords, grid, w, in_mul, out_mul, out_scale, flat_concat=True, sampler="surrogate"):
    """Differentiable spherical modulated conv: gather -> flat concat -> stride-3 conv.  sampler = "surrogate": the live
    GridSamplerNewTextureNoGrad (bilinear/border forward, 3x3 block-mean * 0.1 backward); "texture": GridSamplerNewTexture
    of models/spgan_ops.py's SphereModulatedConv2d (grid_sample_github forward, true gradient)."""
    needs = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, w, in_mul, out_mul))
    if not needs and sampler == "surrogate":
        return sphere_modconv_fused(x, coords, grid, w, in_mul, out_mul, out_scale, None, flat_concat)
    B, C, H, W = x.shape
    Ct = w.shape[1]
    sample = sphere_gather if sampler == "surrogate" else (lambda t, g: grid_sample(t, g, "texture"))
    gx = sample(x, grid)
    if coords is not None:
        gc = encode_coords(sample(coords, grid))
        if flat_concat:
            inp = torch.cat([gx.reshape(1, B * C, 3 * H, 3 * W), gc.reshape(1, B * coords.shape[1], 3 * H, 3 * W)], 1)
            inp = inp.view(B, Ct, 3 * H, 3 * W)
        else:
            inp = torch.cat([gx, gc], 1)
    else:
        inp = gx
    if not needs:
        return conv_apply(inp, w, _SPHERE_GEOM, False, None, in_mul, out_mul, out_scale)
    return _ConvFn.apply(inp, w, in_mul, out_mul, _SPHERE_GEOM, False, None, out_scale)


# =================================================================================================== channels-last chain
# Inference path of the texture synthesiser (models/spgan/spgan.py:924-978) without the NCHW fp32 tensors between its
# convs: every GEMM epilogue / FIR tail writes the NEXT conv's packed operand, ToRGB is folded into the epilogue of the
# conv it reads.  See csrc/chain.cu and the sinks of spgan_conv_gemm_ex.
_UP_GEOM = ConvGeom(3, 3, stride=2, transposed=True, crop=1)
_C3_GEOM = ConvGeom(3, 3)


def _gemm_ex(cp, flops, st, **kw):
    io = lib.GemmIO()
    for k, v in kw.items():
        if isinstance(v, torch.Tensor) or v is None:
            v = v.data_ptr() if v is not None else None
        setattr(io, k, v)
    _timed_call(flops, "spgan_conv_gemm_ex", ctypes.byref(cp), ctypes.byref(io), st)


def packed_to_float(packed, fmt):
    """(2, rows, cols) 16-bit hi/lo planes -> fp32 (tests, diagnostics)."""
    v = packed.view(torch.float16 if fmt else torch.bfloat16)
    return v[0].float() + v[1].float()


def chain_pack_input(x, in_mul, precision):
    """NCHW fp32 -> packed operand (2, B*H*W, Cp) of a conv whose lattice is the image itself."""
    x = _f32c(x, "chain")
    B, C, H, W = x.shape
    Cp = _round_up(C, 64)
    a = torch.empty((2, B * H * W, Cp), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        lib.call("spgan_pack_act", _ptr(a), _ptr(x), _ptr(in_mul), B, C, H, W, Cp, 0, 0, H, W, 1, _fmt(precision), _stream(x))
    return a


def chain_upconv(a, B, H, W, w, out_mul, out_scale, precision):
    """Transposed 3x3 stride-2 conv cropped by 1 (models/ops.py:617-619) from the packed input `a` (2, B*H*W, Cp) to
    channels-last polyphase planes (B, 4, H, W, Cout) fp32: four parity GEMMs, each writing its plane."""
    O, C = w.shape[0], w.shape[1]
    Cp = a.shape[2]
    oh, ow = _UP_GEOM.out_size(H, W)
    passes, _ = plan_passes(_UP_GEOM, False, (H, W), (oh, ow))
    step, pt, pl, Hl, Wl, mapped = _phase_taps(passes)
    assert step == 1 and pt == 0 and pl == 0 and Hl <= H and Wl <= W
    pp = torch.empty((B, 4, H, W, O), device=a.device, dtype=torch.float32)
    fmt = _fmt(precision)
    st = _stream(a)
    with torch.cuda.device(a.device):
        for p, taps in zip(passes, mapped):
            q = dict(p, in_stride=1, out_stride=1, off_y=0, off_x=0, taps=[(oy, ox, wi) for _, oy, ox, wi in taps])
            cp = _fill_pass(q, B, C, H, W, O, H, W, C * 9, 9, out_scale, 0, 0.0, 1.0, precision)
            wp = _packed_weight(w, O, C, C * 9, 9, [t[2] for t in p["taps"]], Cp, False, _wfmt(precision))
            plane = p["off_y"] * 2 + p["off_x"]
            _gemm_ex(cp, 2.0 * B * p["My"] * p["Mx"] * O * C * len(taps), st, a_packed=a, a_rows=B * H * W, kp=Cp, fmt=fmt,
                     w_fmt=_wfmt(precision), w_packed=wp, out_mul=out_mul, y=pp.data_ptr() + 4 * plane * H * W * O, y_layout=1,
                     y_bstride=4 * H * W * O)
    return pp, (oh, ow)


def chain_upblur_pack(pp, out_hw, kernel, noise, noise_w, bias, next_mul, next_precision, negative_slope=0.2,
                      scale=2 ** 0.5):
    """pp (B, 4, Hq, Wq, C) channels-last planes -> the next conv's packed operand (2, B*oh*ow, C)."""
    B, _, Hq, Wq, C = pp.shape
    zh, zw = out_hw
    oh, ow = zh - 2, zw - 2
    out = torch.empty((2, B * oh * ow, C), device=pp.device, dtype=torch.bfloat16)
    nz = _f32c(noise, "upblur_pack") if noise is not None else None
    if nz is not None and nz.numel() != B * oh * ow:
        raise RuntimeError("upblur_pack: noise must have shape (B, 1, %d, %d)" % (oh, ow))
    with torch.cuda.device(pp.device):
        lib.call("spgan_upblur_pack", _ptr(out), _ptr(pp), _ptr(_f32c(kernel, "upblur_pack")), _ptr(nz),
                 _ptr(noise_w) if nz is not None else _ptr(None), _ptr(bias), _ptr(next_mul), B, C, zh, zw, Hq, Wq, C,
                 B * oh * ow, _fmt(next_precision), float(negative_slope), float(scale), _stream(pp))
    return out, (oh, ow)


def chain_conv3(a, B, H, W, w, out_mul, out_scale, noise, noise_w, bias, act, precision, next_mul=None,
                next_precision=None, rgb_w=None, want_nchw=False):
    """Unpadded 3x3 conv + noise + bias + leaky-ReLU from the packed input `a` (2, B*H*W, Cp).  Sinks: the next conv's
    packed operand (when next_precision is given), ToRGB partial sums (when rgb_w (B, 3, Cout) is given), an NCHW fp32
    tensor (want_nchw).  Returns (packed or None, (rgb_part, slots) or None, nchw or None, (oh, ow))."""
    O, C = w.shape[0], w.shape[1]
    Cp = a.shape[2]
    oh, ow = H - 2, W - 2
    fmt = _fmt(precision)
    taps = [(ky, kx, ky * 3 + kx) for ky in range(3) for kx in range(3)]
    p = dict(My=oh, Mx=ow, in_stride=1, out_stride=1, off_y=0, off_x=0, taps=taps)
    alpha, gain = act if act is not None else (0.0, 1.0)
    cp = _fill_pass(p, B, C, H, W, O, oh, ow, C * 9, 9, out_scale, 1 if act is not None else 0, alpha, gain, precision)
    st = _stream(a)
    nz = _f32c(noise, "chain_conv3") if noise is not None else None
    if nz is not None and nz.numel() != B * oh * ow:
        raise RuntimeError("chain_conv3: noise must have shape (B, 1, %d, %d)" % (oh, ow))
    kw = {}
    packed = rgb = y = None
    with torch.cuda.device(a.device):
        wp = _packed_weight(w, O, C, C * 9, 9, list(range(9)), Cp, False, _wfmt(precision))
        if next_precision is not None:
            packed = torch.empty((2, B * oh * ow, O), device=a.device, dtype=torch.bfloat16)
            kw.update(y_packed=packed, next_mul=next_mul, y_packed_rows=B * oh * ow, y_packed_cols=O,
                      y_packed_fmt=_fmt(next_precision))
        if rgb_w is not None:
            slots = int(lib.load().spgan_conv_gemm_rgb_slots(ctypes.byref(cp), B * H * W))
            part = torch.empty((slots, B, rgb_w.shape[1], oh * ow), device=a.device, dtype=torch.float32)
            kw.update(rgb_w=rgb_w, rgb_part=part, rgb_n=rgb_w.shape[1])
            rgb = (part, slots)
        if want_nchw:
            y = torch.empty((B, O, oh, ow), device=a.device, dtype=torch.float32)
            kw.update(y=y)
        _gemm_ex(cp, 2.0 * B * oh * ow * O * C * 9, st, a_packed=a, a_rows=B * H * W, kp=Cp, fmt=fmt, w_fmt=_wfmt(precision),
                 w_packed=wp,
                 out_mul=out_mul, noise=nz, noise_w=noise_w if nz is not None else None, bias=bias, **kw)
    return packed, rgb, y, (oh, ow)


def rgb_tail(part, slots, bias, skip, B, oh, ow):
    """ToRGB tail: fixed-order sum of the epilogue's partial sums + bias + upsampled skip -> (B, 3, oh, ow)."""
    n = part.shape[2]
    out = torch.empty((B, n, oh, ow), device=part.device, dtype=torch.float32)
    sk = _f32c(skip, "rgb_tail") if skip is not None else None
    if sk is not None and tuple(sk.shape) != tuple(out.shape):
        raise RuntimeError("rgb_tail: skip shape %s != %s" % (tuple(sk.shape), tuple(out.shape)))
    with torch.cuda.device(part.device):
        lib.call("spgan_rgb_tail", _ptr(out), _ptr(part), slots, _ptr(bias), _ptr(sk), B, n, oh * ow, _stream(part))
    return out


# =================================================================================================== structure chain
# Inference path of the structure synthesiser (models/spgan/spgan.py:79-254) with channels-last operands between its convs
# and the 256 + 3 channel split of csrc/structure.cu: per block
#   shortcut 1x1 GEMM (packed x -> NHWC fp32)  ->  spherical gather producer (NHWC x -> main + tail operand)  ->  spherical
#   GEMM (+ LeakyReLU + shortcut residual, writes the 7x7 conv's packed operand)  ->  coordinate tail operand  ->  7x7 GEMM
#   (+ bias + leaky-ReLU, writes the next block's NHWC gather source and packed shortcut operand, or the final NCHW tensor).
def _packed_weight_tail(w, c0, tap_w, kp2, fmt):
    """Second-segment weight [2][Cout][kp2], k2 = t*Cx + j <-> w[o, c0 + j, tap_w[t]]: the channels past c0 of every tap in
    one dense slab (pairs with spgan_sphere_pack_seg / spgan_coord_taps_pack).  Cached per weight version like
    _packed_weight; built with torch ops (a few KB, once per weight version)."""
    key = ("tail", w.data_ptr(), tuple(w.shape), c0, tuple(tap_w), kp2, fmt, w.device.index)
    hit = _weight_cache_get(key, w._version)
    if hit is not None:
        return hit
    O, C = w.shape[0], w.shape[1]
    wt = w.detach().reshape(O, C, -1)[:, c0:, :][:, :, list(tap_w)]  # (O, Cx, T)
    flat = wt.permute(0, 2, 1).reshape(O, -1)
    if flat.shape[1] > kp2:
        raise RuntimeError("weight tail: %d columns do not fit kp2=%d" % (flat.shape[1], kp2))
    v = torch.zeros((O, kp2), device=w.device, dtype=torch.float32)
    v[:, :flat.shape[1]] = flat
    if fmt:
        v = v.clamp(-65504.0, 65504.0)
        hi = v.half()
        lo = (v - hi.float()).half()
    else:
        hi = v.bfloat16()
        lo = (v - hi.float()).bfloat16()
    out = torch.stack([hi, lo]).contiguous().view(torch.bfloat16)
    _weight_cache_put(key, out, w)
    return out


SS_MAIN = 256  # feature channels per tap of the main K segment
# True: the spherical conv of the structure chain is ONE kernel (gather in the GEMM's producer warps, spgan_sphere_conv_gemm with
# the repacked source: no [B*H*W][9*256] operand in HBM); False (default): spgan_sphere_pack_seg writes the operand to HBM and
# spgan_conv_gemm_ex consumes it.  Measured on the B200 at B = 64, 256 + 3 -> 256 channels, bf16x3, in situ (bench.py
# --profile-calls): fused 1.08 / 0.66 / 0.50 / 0.35 ms at 35 / 29 / 23 / 17 pixels against 0.30 + 0.25 / 0.22 + 0.16 / 0.15 + 0.12 /
# 0.10 + 0.08 ms for producer + GEMM.  The persistent one-CTA-per-SM GEMM leaves the L1 ~30 KB (223 KB of the 256 KB are
# operand stages and the corner table), so the 128 KB a stage gathers comes from L2 at 8 warps' worth of loads in flight:
# ~5.8 us per stage against 0.9 us of MMAs.  The standalone producer runs 24 warps per SM with a full L1 (87 % hit rate).
# Both paths are tested against each other and against the padded formulation (tests/test_gpu_ops.py).
SS_FUSED_GATHER = False


def ss_input(x, precision):
    """NCHW fp32 local latent -> (NHWC fp32 gather source, packed un-modulated operand of the first shortcut conv)."""
    x = _f32c(x, "structure chain")
    B, C, H, W = x.shape
    xh = torch.empty((B, H, W, C), device=x.device, dtype=torch.float32)
    xp = torch.empty((2, B * H * W, C), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        lib.call("spgan_nchw_to_nhwc", _ptr(xh), _ptr(x), B, C, H, W, _stream(x))
        lib.call("spgan_pack_act", _ptr(xp), _ptr(x), _ptr(None), B, C, H, W, C, 0, 0, H, W, 1, _fmt(precision), _stream(x))
    return xh, xp


def ss_shortcut(xp, B, H, W, w, bias, precision):
    """1x1 conv + bias (nn.Conv2d(256, 256, 1), models/spgan/spgan.py:141) from the packed operand -> NHWC fp32."""
    O, C = w.shape[0], w.shape[1]
    y = torch.empty((B, H, W, O), device=xp.device, dtype=torch.float32)
    p = dict(My=H, Mx=W, in_stride=1, out_stride=1, off_y=0, off_x=0, taps=[(0, 0, 0)])
    cp = _fill_pass(p, B, C, H, W, O, H, W, C, 1, 1.0, 0, 0.0, 1.0, precision)
    with torch.cuda.device(xp.device):
        wp = _packed_weight(w, O, C, C, 1, [0], xp.shape[2], False, _wfmt(precision))
        _gemm_ex(cp, 2.0 * B * H * W * O * C, _stream(xp), a_packed=xp, a_rows=B * H * W, kp=xp.shape[2], fmt=_fmt(precision),
                 w_fmt=_wfmt(precision), w_packed=wp, bias=bias, y=y, y_layout=1)
    return y


def ss_sphere(xh, coords, grid, grid_group, w, in_mul, out_mul, out_scale, act, residual_nhwc, next_mul, precision):
    """Spherical modulated conv + LeakyReLU + shortcut residual (models/spgan_ops_gs.py:791-816, models/spgan/spgan.py:169)
    from the NHWC gather source; writes the NEXT (7x7) conv's packed operand (2, B*H*W, Cout), modulated by next_mul."""
    B, H, W, C = xh.shape
    O, Ct = w.shape[0], w.shape[1]
    nc = 0 if coords is None else coords.shape[1]
    Cm = SS_MAIN
    Cx = Ct - Cm
    if Ct != C + nc or Cx < 0 or Cx >= 32 or C % 64:
        raise RuntimeError("structure chain: %d + %d channels do not split as %d + tail" % (C, nc, Cm))
    kp2 = _round_up(9 * Cx, 64) if Cx else 0
    rows = B * H * W
    fmt = _fmt(precision)
    st = _stream(xh)
    G = grid.shape[0]
    if B % G or (B // G) != grid_group:
        raise RuntimeError("structure chain: %d grids for a batch of %d in groups of %d" % (G, B, grid_group))
    out = torch.empty((2, rows, O), device=xh.device, dtype=torch.bfloat16)
    alpha, gain = act
    if SS_FUSED_GATHER and C == 256 and kp2 == 64 and O <= 256 and O % 32 == 0 and H * W >= 128:
        # the gather runs in the producer warps of the GEMM itself (csrc/sphere_umma.cu, vectorised producer): neither the
        # reference's 9x gathered fp32 tensor nor the [B*H*W][9*256] 16-bit operand exists in HBM
        with torch.cuda.device(xh.device):
            cmap = _sphere_chan_map(B, C, nc, _round_up(Ct, 64), True, xh.device, group=grid_group)
            xg = torch.empty((int(lib.load().spgan_sphere_pack_seg_scratch(B, C, H, W)),), device=xh.device, dtype=torch.float32)
            lib.call("spgan_sphere_concat_repack", _ptr(xg), _ptr(xh), _ptr(coords), _ptr(cmap), B, C, H, W, cmap.shape[1], st)
            wp = _packed_weight(w, O, Cm, Ct * 9, 9, list(range(9)), Cm, True, _wfmt(precision))
            w2 = _packed_weight_tail(w, Cm, list(range(9)), kp2, _wfmt(precision))
            p = dict(My=H, Mx=W, in_stride=1, out_stride=1, off_y=0, off_x=0, taps=[(0, 0, 0)])
            cp = _fill_pass(p, B, Ct, H, W, O, H, W, Ct * 9, 9, out_scale, 1, alpha, gain, precision)
            sin = lib.SphereIn()
            sin.coords = coords.data_ptr() if coords is not None else None
            sin.grid, sin.in_mul, sin.chan_map = grid.data_ptr(), (in_mul.data_ptr() if in_mul is not None else None), cmap.data_ptr()
            sin.C, sin.Cp, sin.xg, sin.grid_group, sin.cmap_ld = C, Cm, xg.data_ptr(), grid_group, cmap.shape[1]
            io = lib.GemmIO()
            io.kp, io.fmt, io.w_fmt, io.w_packed, io.w2_packed, io.kp2 = 9 * Cm, fmt, _wfmt(precision), wp.data_ptr(), w2.data_ptr(), kp2
            io.out_mul = out_mul.data_ptr() if out_mul is not None else None
            io.residual_nhwc = residual_nhwc.data_ptr() if residual_nhwc is not None else None
            io.y_packed, io.next_mul = out.data_ptr(), (next_mul.data_ptr() if next_mul is not None else None)
            io.y_packed_rows, io.y_packed_cols, io.y_packed_fmt = rows, O, fmt
            _timed_call(2.0 * rows * O * Ct * 9, "spgan_sphere_conv_gemm", ctypes.byref(cp), ctypes.byref(sin), ctypes.byref(io), st)
        return out
    a = torch.empty((2, rows, 9 * Cm), device=xh.device, dtype=torch.bfloat16)
    a2 = torch.empty((2, rows, kp2), device=xh.device, dtype=torch.bfloat16) if kp2 else None
    with torch.cuda.device(xh.device):
        cmap = _sphere_chan_map(B, C, nc, _round_up(Ct, 64), True, xh.device, group=grid_group)
        need = int(lib.load().spgan_sphere_pack_seg_scratch(B, C, H, W))
        scratch = torch.empty((need,), device=xh.device, dtype=torch.float32) if need else None
        lib.call("spgan_sphere_pack_seg", _ptr(a), _ptr(a2), _ptr(xh), _ptr(coords), _ptr(grid), _ptr(in_mul), _ptr(cmap), B, C,
                 H, W, grid_group, Cm, cmap.shape[1], kp2, fmt, _ptr(scratch), st)
        wp = _packed_weight(w, O, Cm, Ct * 9, 9, list(range(9)), Cm, True, _wfmt(precision))
        w2 = _packed_weight_tail(w, Cm, list(range(9)), kp2, _wfmt(precision)) if kp2 else None
        p = dict(My=H, Mx=W, in_stride=1, out_stride=1, off_y=0, off_x=0, taps=[(0, 0, 0)])
        cp = _fill_pass(p, B, Ct, H, W, O, H, W, Ct * 9, 9, out_scale, 1, alpha, gain, precision)
        _gemm_ex(cp, 2.0 * rows * O * Ct * 9, st, a_packed=a, a_rows=rows, kp=9 * Cm, fmt=fmt, w_fmt=_wfmt(precision), w_packed=wp,
                 a2_packed=a2, a2_rows=rows if kp2 else 0, w2_packed=w2, kp2=kp2, out_mul=out_mul, residual_nhwc=residual_nhwc,
                 y_packed=out, next_mul=next_mul, y_packed_rows=rows, y_packed_cols=O, y_packed_fmt=fmt)
    return out


def ss_conv_k(a, coords, B, H, W, w, in_mul, out_mul, out_scale, bias, act, precision, last):
    """Unpadded k x k modulated conv + bias + leaky-ReLU (ConditionalBlock, models/spgan/spgan.py:100-101 + models/ops.py:853)
    whose input is [256 features (packed operand `a`, already modulated), 3 encoded coordinate planes (tail operand built
    here from the raw planes)].  Sinks: last=False -> (NHWC fp32, packed un-modulated operand) for the next block;
    last=True -> the NCHW fp32 structure latent."""
    O, Ct, kh, kw = w.shape
    Cm = a.shape[2]
    nc = Ct - Cm
    oh, ow = H - kh + 1, W - kw + 1
    T = kh * kw
    kp2 = _round_up(T * nc, 64)
    rows_out = B * oh * ow
    fmt = _fmt(precision)
    st = _stream(a)
    alpha, gain = act
    taps = [(ky, kx, ky * kw + kx) for ky in range(kh) for kx in range(kw)]
    p = dict(My=oh, Mx=ow, in_stride=1, out_stride=1, off_y=0, off_x=0, taps=taps)
    cp = _fill_pass(p, B, Ct, H, W, O, oh, ow, Ct * T, T, out_scale, 1, alpha, gain, precision)
    a2 = torch.empty((2, rows_out, kp2), device=a.device, dtype=torch.bfloat16)
    kw_sinks = {}
    y = packed = None
    if last:
        y = torch.empty((B, O, oh, ow), device=a.device, dtype=torch.float32)
        kw_sinks.update(y=y)
    else:
        y = torch.empty((B, oh, ow, O), device=a.device, dtype=torch.float32)
        packed = torch.empty((2, rows_out, O), device=a.device, dtype=torch.bfloat16)
        kw_sinks.update(y=y, y_layout=1, y_packed=packed, y_packed_rows=rows_out, y_packed_cols=O, y_packed_fmt=fmt)
    with torch.cuda.device(a.device):
        lib.call("spgan_coord_taps_pack", _ptr(a2), _ptr(coords), _ptr(in_mul), B, nc, H, W, kh, kw, Ct, Cm, kp2, fmt, st)
        wp = _packed_weight(w, O, Cm, Ct * T, T, list(range(T)), Cm, False, _wfmt(precision))
        w2 = _packed_weight_tail(w, Cm, list(range(T)), kp2, _wfmt(precision))
        _gemm_ex(cp, 2.0 * rows_out * O * Ct * T, st, a_packed=a, a_rows=B * H * W, kp=Cm, fmt=fmt, w_fmt=_wfmt(precision),
                 w_packed=wp, a2_packed=a2, a2_rows=rows_out, w2_packed=w2, kp2=kp2, out_mul=out_mul, bias=bias, **kw_sinks)
    return y, packed, (oh, ow)


# =================================================================================================== training-loop tails
_EMA_TABLES = {}


def ema_accumulate(dst_params, src_params, decay):
    """utils.py:86-94 (`accumulate`): dst = dst * decay + src * (1 - decay) for every parameter pair in ONE launch
    (spgan_ema_multi over a cached device table of (dst, src, count) chunks)."""
    dst_params, src_params = list(dst_params), list(src_params)
    if not dst_params:
        return
    dev = dst_params[0].device
    ptrs = tuple(p.data_ptr() for p in dst_params) + tuple(p.data_ptr() for p in src_params)
    hit = _EMA_TABLES.get(ptrs)
    if hit is None:
        chunk = int(lib.load().spgan_ema_chunk_elems())
        rows = []
        for d, s in zip(dst_params, src_params):
            if d.shape != s.shape or d.dtype != torch.float32 or s.dtype != torch.float32 or not d.is_contiguous() or not s.is_contiguous():
                raise RuntimeError("ema_accumulate: parameter pairs must be contiguous fp32 tensors of equal shape")
            n = d.numel()
            for off in range(0, n, chunk):
                rows.append((d.data_ptr() + 4 * off, s.data_ptr() + 4 * off, min(chunk, n - off)))
        table = torch.tensor(rows, dtype=torch.int64).to(dev)
        _EMA_TABLES.clear()  # one live model pair at a time; parameter storage is stable across optimizer steps
        hit = _EMA_TABLES[ptrs] = (table, len(rows))
    with torch.cuda.device(dev):
        lib.call("spgan_ema_multi", _ptr(hit[0]), hit[1], float(decay), float(1 - decay), _stream(dst_params[0]))


class _MinibatchStddevFn(torch.autograd.Function):
    """models/stylegan2discriminator.py:205-212 + the torch.cat: forward is one fused kernel pair (spgan_minibatch_stddev),
    backward is written in differentiable torch ops (R1 differentiates through it a second time)."""

    @staticmethod
    def forward(ctx, h, group, eps):
        hc = _f32c(h, "minibatch_stddev")
        B, C, H, W = hc.shape
        out = torch.empty((B, C + 1, H, W), device=hc.device, dtype=torch.float32)
        partial = torch.empty(((B // group) * 8,), device=hc.device, dtype=torch.float32)
        with torch.cuda.device(hc.device):
            lib.call("spgan_minibatch_stddev", _ptr(out), _ptr(partial), _ptr(hc), B, C, H * W, int(group), float(eps), _stream(hc))
        ctx.save_for_backward(h)
        ctx.group, ctx.eps = int(group), float(eps)
        return out

    @staticmethod
    def backward(ctx, go):
        h, = ctx.saved_tensors
        B, C, H, W = h.shape
        G = ctx.group
        M = B // G
        gh = go[:, :C]
        if ctx.needs_input_grad[0]:
            hv = h.view(G, M, C, H, W)
            d = hv - hv.mean(0, keepdim=True)
            sd = torch.sqrt((d * d).mean(0) + ctx.eps)
            gs = go[:, C].reshape(G, M, H * W).sum((0, 2))  # d loss / d (the per-sub-batch scalar)
            gh = gh + (gs.view(1, M, 1, 1, 1) / float(C * H * W * G) * d / sd.unsqueeze(0)).reshape(B, C, H, W)
        return gh, None, None


def minibatch_stddev(h, group, eps=1e-8):
    """(B, C, H, W) -> (B, C + 1, H, W): the feature map with its minibatch-stddev channel appended (stddev_feat = 1)."""
    _check_cuda(h, "minibatch_stddev")
    if h.shape[0] % group:
        raise RuntimeError("minibatch_stddev: batch %d is not a multiple of the group %d" % (h.shape[0], group))
    return _MinibatchStddevFn.apply(h, int(group), float(eps))


# =================================================================================================== style chain (f3)
def mapping_chain(z, weights, biases, w_scale, b_scale, alpha=0.2, gain=2 ** 0.5):
    """PixelNorm + len(weights) x (EqualLinear 512 -> 512 + fused leaky-ReLU) in one launch (csrc/style_chain.cu).  z: (B, 512)
    fp32, possibly a strided view (rows z.stride(0) floats apart, unit inner stride).  No autograd."""
    _check_cuda(z, "mapping_chain")
    if z.dim() != 2 or z.shape[1] != 512 or z.stride(1) != 1:
        raise RuntimeError("mapping_chain: z must be (B, 512) with unit inner stride, got %s / %s" % (tuple(z.shape), z.stride()))
    B = z.shape[0]
    n = len(weights)
    out = torch.empty((B, 512), device=z.device, dtype=torch.float32)
    wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in weights])
    bp = (ctypes.c_void_p * n)(*[(b.data_ptr() if b is not None else None) for b in biases])
    with torch.cuda.device(z.device):
        lib.call("spgan_mapping_chain", _ptr(out), _ptr(z), int(z.stride(0)), B, wp, bp, n, float(w_scale), float(b_scale),
                 float(alpha), float(gain), _stream(z))
    return out


_MOD_RECORD = "<QQQqqiiiiffff"


def modulation_table(records, device):
    """Device table for spgan_modulation_batch from a list of dicts (wm, bm, wsq tensors or None; s_off, d_off, Cin, Cout,
    style_sel, style_idx, m_scale, m_lr_mul, c_scale, eps)."""
    import struct
    size = int(lib.load().spgan_modulation_layer_bytes())
    if struct.calcsize(_MOD_RECORD) != size:
        raise RuntimeError("modulation_table: record layout mismatch (%d != %d bytes)" % (struct.calcsize(_MOD_RECORD), size))
    blob = b"".join(struct.pack(_MOD_RECORD, r["wm"].data_ptr(), r["bm"].data_ptr() if r["bm"] is not None else 0,
                                r["wsq"].data_ptr() if r["wsq"] is not None else 0, r["s_off"], r["d_off"], r["Cin"], r["Cout"],
                                r["style_sel"], r["style_idx"], r["m_scale"], r["m_lr_mul"], r["c_scale"], r["eps"]) for r in records)
    return torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)


def modulation_batch(table, n_layers, total_floats, styles, gl, B):
    """One launch for the (modulation, demodulation) pairs of every layer in `table`; returns the flat fp32 buffer."""
    ref = styles if styles is not None else gl
    out = torch.empty((int(total_floats),), device=ref.device, dtype=torch.float32)
    with torch.cuda.device(ref.device):
        lib.call("spgan_modulation_batch", _ptr(out), _ptr(table), int(n_layers), _ptr(styles),
                 int(styles.stride(0)) if styles is not None else 0, _ptr(gl), int(gl.stride(0)) if gl is not None else 0, int(B),
                 _stream(ref))
    return out
