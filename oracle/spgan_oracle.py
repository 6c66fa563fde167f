"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch CPU restatement (numpy float64/float32 + torch CPU fp32) of the SP-GAN conv hot path,
written by reading the reference and cited function by function (paths relative to /root/reference).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import this file.  The product package (`sp-gan-tip2025_b200/`) never imports it and has no CPU path.

Pinning: the reference has no tests or golden vectors, so this oracle is pinned against
outputs of the reference itself, executed in the build container by `oracle/make_golden.py`
(import of /root/reference with in-process stubs) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` re-checks this file against those fixtures on every run.

The dense arithmetic (conv2d / conv_transpose2d / linear / grid_sample) is PyTorch's in the reference
(third-party, torch==2.0.0+cu118 pinned in configs/env/environment.yml:236; this image has torch 2.11);
here the gather is restated explicitly in numpy following ATen's published formula
(torch/include/ATen/native/GridSampler.h:27-36 unnormalise, :58-60 clip) and the contractions use
torch CPU `F.conv2d` etc. with the reference's own call-site arguments.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = 2 ** 0.5


# --------------------------------------------------------------------------------------------
# K1  fused bias + activation            models/custom_ops/fused_bias_act_kernel.cu:18-49
# --------------------------------------------------------------------------------------------
def fused_bias_act(x, b=None, ref=None, act=3, grad=0, alpha=0.2, scale=SQRT2):
    """y = act(x + b[channel]) * scale with the kernel's `act*10+grad` table.

    x: ndarray (N, C, ...) ; b: (C,) or None ; ref: same shape as x or None.
    act 1 = linear, 3 = leaky-relu.  grad 0 = forward, 1 = first derivative (sign taken from `ref`),
    2 = second derivative (identically zero).  Bias index = (i / step_b) % size_b with step_b the
    product of the dims after dim 1 (fused_bias_act_kernel.cu:66-71).
    """
    x = np.asarray(x)
    v = x.astype(x.dtype, copy=True)
    if b is not None and np.size(b):
        shape = [1, -1] + [1] * (x.ndim - 2)
        v = v + np.asarray(b, dtype=x.dtype).reshape(shape)
    code = act * 10 + grad
    if code in (10, 11):
        y = v
    elif code in (12, 32):
        y = np.zeros_like(v)
    elif code == 30:
        y = np.where(v > 0, v, v * x.dtype.type(alpha))
    elif code == 31:
        r = np.asarray(ref)
        y = np.where(r > 0, v, v * x.dtype.type(alpha))
    else:  # kernel's `default:` falls through to linear
        y = v
    return (y * x.dtype.type(scale)).astype(x.dtype)


def fused_leaky_relu(x, bias, negative_slope=0.2, scale=SQRT2):
    """models/custom_ops/fused_act.py:56-64 (CUDA branch semantics: the slope argument is honoured;
    the reference's CPU fallback at :91-98 hard-codes 0.2, identical for every call in spgan.yaml)."""
    return fused_bias_act(x, bias, None, 3, 0, negative_slope, scale)


def fused_leaky_relu_backward(grad_out, out, negative_slope=0.2, scale=SQRT2):
    """models/custom_ops/fused_act.py:24-44 — grad_input and grad_bias (sum over all dims but 1)."""
    gi = fused_bias_act(grad_out, None, out, 3, 1, negative_slope, scale)
    dims = tuple(d for d in range(gi.ndim) if d != 1)
    return gi, gi.sum(axis=dims)


def fused_leaky_relu_t(x, bias, negative_slope=0.2, scale=SQRT2):
    """torch (differentiable) form of the same function, used by the model-level oracle."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    return F.leaky_relu(x + bias.view(shape), negative_slope) * scale


# --------------------------------------------------------------------------------------------
# K2/K3  upfirdn2d                       models/custom_ops/upfirdn2d_kernel.cu:49-105 (definition),
#                                        models/custom_ops/upfirdn2d.py:164-205 (native form)
# --------------------------------------------------------------------------------------------
def upfirdn2d(x, k, up=(1, 1), down=(1, 1), pad=(0, 0, 0, 0)):
    """x: (B, C, H, W) ndarray; k: (kh, kw); up/down = (x, y); pad = (x0, x1, y0, y1).

    Zero-stuff by `up`, pad (negative pad crops), correlate with the flipped kernel, keep every
    `down`-th sample.  Accumulates in the input dtype like the CUDA kernel.
    """
    x = np.asarray(x)
    k = np.asarray(k, dtype=x.dtype)
    up_x, up_y = up
    down_x, down_y = down
    px0, px1, py0, py1 = pad
    B, C, H, W = x.shape
    kh, kw = k.shape
    z = np.zeros((B, C, H * up_y, W * up_x), dtype=x.dtype)
    z[:, :, ::up_y, ::up_x] = x
    z = np.pad(z, ((0, 0), (0, 0), (max(py0, 0), max(py1, 0)), (max(px0, 0), max(px1, 0))))
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    fh = z.shape[2] - kh + 1
    fw = z.shape[3] - kw + 1
    out = np.zeros((B, C, max(fh, 0), max(fw, 0)), dtype=x.dtype)
    kf = k[::-1, ::-1]
    for i in range(kh):
        for j in range(kw):
            out += kf[i, j] * z[:, :, i:i + fh, j:j + fw]
    return np.ascontiguousarray(out[:, :, ::down_y, ::down_x])


def upfirdn2d_out_size(in_h, in_w, kh, kw, up, down, pad):
    """models/custom_ops/upfirdn2d.py:107-109."""
    out_h = (in_h * up[1] + pad[2] + pad[3] - kh) // down[1] + 1
    out_w = (in_w * up[0] + pad[0] + pad[1] - kw) // down[0] + 1
    return out_h, out_w


def upfirdn2d_grad_pads(in_h, in_w, kh, kw, up, down, pad):
    """g_pad of the backward op — models/custom_ops/upfirdn2d.py:116-121."""
    out_h, out_w = upfirdn2d_out_size(in_h, in_w, kh, kw, up, down, pad)
    gx0 = kw - pad[0] - 1
    gy0 = kh - pad[2] - 1
    gx1 = in_w * up[0] - out_w * down[0] + pad[0] - up[0] + 1
    gy1 = in_h * up[1] - out_h * down[1] + pad[2] - up[1] + 1
    return gx0, gx1, gy0, gy1


def upfirdn2d_backward(grad_out, k, up, down, pad, in_size):
    """models/custom_ops/upfirdn2d.py:24-56: same op with up/down swapped and the flipped kernel."""
    _, _, in_h, in_w = in_size
    kh, kw = np.asarray(k).shape
    g_pad = upfirdn2d_grad_pads(in_h, in_w, kh, kw, up, down, pad)
    return upfirdn2d(grad_out, np.asarray(k)[::-1, ::-1], up=down, down=up, pad=g_pad)


def make_kernel(k):
    """models/ops.py:23-28."""
    k = np.asarray(k, dtype=np.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return (k / k.sum()).astype(np.float32)


def upfirdn2d_t(x, k, up=1, down=1, pad=(0, 0)):
    """torch (differentiable) upfirdn2d for the model-level oracle (same definition)."""
    B, C, H, W = x.shape
    kh, kw = k.shape
    z = x.reshape(B * C, 1, H, 1, W, 1)
    z = F.pad(z, [0, up - 1, 0, 0, 0, up - 1])
    z = z.reshape(B * C, 1, H * up, W * up)
    p0, p1 = pad
    z = F.pad(z, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    z = z[:, :, max(-p0, 0): z.shape[2] - max(-p1, 0), max(-p0, 0): z.shape[3] - max(-p1, 0)]
    z = F.conv2d(z, torch.flip(k, [0, 1]).view(1, 1, kh, kw))
    z = z[:, :, ::down, ::down]
    return z.reshape(B, C, z.shape[2], z.shape[3])


# --------------------------------------------------------------------------------------------
# Spherical sampling pattern             models/spherenet/grid_generator.py:137-283, 303-352
# --------------------------------------------------------------------------------------------
def _tangent_kernel(x_total, y_total, kh=3, kw=3):
    """createKernel (grid_generator.py:303-323): tangent-plane tap offsets, float64."""
    d_lat = np.pi / x_total
    d_lon = 2 * np.pi / y_total
    rx = np.arange(-(kw // 2), kw // 2 + 1)
    ry = np.arange(-(kh // 2), kh // 2 + 1)
    ker_x = np.tan(rx * d_lon)
    ker_y = np.tan(ry * d_lat) / np.cos(ry * d_lon)
    return np.meshgrid(ker_x, ker_y)


def _min_max_norm(v):
    """grid_generator.py:349-352, start=-1."""
    return (v - np.min(v)) / (np.max(v) - np.min(v)) * 2 + (-1)


def patch_angular_ranges(h, w, cp):
    """lat/lon centre ranges of the patch — the plain branch, grid_generator.py:215-241.

    (`full_shape` / `pre_sample_mode` branches at :169-214 are never taken by spgan.yaml or the
    close-loop manager, which leaves `full_shape` commented out; not restated.)
    Training hard-codes partial 0.8 unless `test_flag` is set (:164-167).
    """
    partial = 0.8
    if cp.get("test_flag", False):
        partial = cp.get("partial", partial)
    x_st = cp["p_x_st"] * np.pi * partial
    x_ed = cp["p_x_ed"] * np.pi * partial
    y_st = cp["p_y_st"] * np.pi * 2
    y_ed = cp["p_y_ed"] * np.pi * 2
    if y_ed != 2 * np.pi:
        y_ed = y_ed % (np.pi * 2)
    lat_range = np.linspace(x_st, x_ed, h) - (np.pi / 2 * partial)
    if cp["circular_flag"]:
        y_ed = y_ed + 2 * np.pi
    lon_range = np.linspace(y_st, y_ed, w) - np.pi
    return lat_range, lon_range


def create_sampling_pattern(h, w, cp):
    """createSamplingPattern, stride 1, 3x3 — returns (1, 3h, 3w, 2) float64 = (lat, lon) in grid units."""
    kh = kw = 3
    ker_x, ker_y = _tangent_kernel(cp["x_total"], cp["y_total"], kh, kw)
    rho = np.sqrt(ker_x ** 2 + ker_y ** 2)
    rho[kh // 2][kw // 2] = 1e-8
    nu = np.arctan(rho)
    cos_nu, sin_nu = np.cos(nu), np.sin(nu)
    lat_range, lon_range = patch_angular_ranges(h, w, cp)

    lat = np.array([np.arcsin(cos_nu * np.sin(t) + ker_y * sin_nu * np.cos(t) / rho) for t in lat_range])
    lat_norm_c = _min_max_norm(lat_range)
    lat_off = np.empty_like(lat)
    for i in range(h):  # get_pattern (:325-335): offsets relative to the centre tap
        lat_off[i] = lat[i] - lat[i][kh // 2, kw // 2]
    lat_rows = np.stack([lat_norm_c[i] + lat_off[i] for i in range(h)])  # add_pattern_to_lat (:337-346)
    lat_norm = np.array([lat_rows for _ in lon_range]).transpose((1, 0, 2, 3))  # (H, W, 3, 3)

    lon = np.array([
        np.arctan(ker_x * sin_nu / (rho * np.cos(t) * cos_nu - ker_y * np.sin(t) * sin_nu))
        for t in lat_range])  # (H, 3, 3)
    lon_norm_c = _min_max_norm(lon_range)
    lon_norm = np.array([lon + c for c in lon_norm_c]).transpose((1, 0, 2, 3))  # (H, W, 3, 3)

    lat_g = (lat_norm / 2 + 0.5) * cp["x_total"]
    lon_g = (lon_norm / 2 + 0.5) * cp["y_total"]
    ll = np.stack((lat_g, lon_g)).transpose((1, 3, 2, 4, 0))  # (H, 3, W, 3, 2)
    return ll.reshape((1, h * kh, w * kw, 2))


def gen_sampling_grid(h, w, cp):
    """genSamplingPattern (models/spgan_ops_gs.py:410-428, models/spherenet/sphere_conv2d.py:147-165):
    the (1, 3h, 3w, 2) float32 grid in F.grid_sample convention, last dim = (x = lon, y = lat)."""
    p = create_sampling_pattern(h, w, cp)
    lat_grid = (p[:, :, :, 0] / cp["x_total"]) * 2 - 1
    lon_grid = (p[:, :, :, 1] / cp["y_total"]) * 2 - 1
    return np.stack((lon_grid, lat_grid), axis=-1).astype(np.float32)


def batch_sampling_grid(h, w, coords_partial, batch):
    """The per-batch grid of models/spgan_ops_gs.py:760-789: list → one grid per sample; dict with
    test_flag → one grid repeated over the batch."""
    if isinstance(coords_partial, (list, tuple)):
        return np.concatenate([gen_sampling_grid(h, w, cp) for cp in coords_partial], axis=0)
    g = gen_sampling_grid(h, w, coords_partial)
    return np.repeat(g, batch, axis=0)


# --------------------------------------------------------------------------------------------
# L1  bilinear gather, border padding, align_corners=True
#     models/spherenet/grid_generator.py:610-613 → ATen grid_sampler_2d
# --------------------------------------------------------------------------------------------
def gather_indices(grid, in_h, in_w):
    """Integer corner indices and fp32 weights, computed in float32 exactly as ATen does:
    ix = ((gx + 1) / 2) * (W - 1); clip to [0, W-1]; x0 = floor(ix); x1 = x0 + 1 (value clamped)."""
    g = np.asarray(grid, dtype=np.float32)
    one, two = np.float32(1), np.float32(2)
    ix = ((g[..., 0] + one) / two) * np.float32(in_w - 1)
    iy = ((g[..., 1] + one) / two) * np.float32(in_h - 1)
    ix = np.minimum(np.float32(in_w - 1), np.maximum(ix, np.float32(0)))
    iy = np.minimum(np.float32(in_h - 1), np.maximum(iy, np.float32(0)))
    x0f, y0f = np.floor(ix), np.floor(iy)
    wx1 = (ix - x0f).astype(np.float32)
    wy1 = (iy - y0f).astype(np.float32)
    x0 = x0f.astype(np.int32)
    y0 = y0f.astype(np.int32)
    return x0, y0, wx1, wy1


def grid_sample_border(z, grid):
    """z: (B, C, H, W) float32; grid: (B, Ho, Wo, 2) float32 → (B, C, Ho, Wo).
    Out-of-range corner taps (x0+1 == W) carry weight 0 after the clip, as in ATen."""
    z = np.asarray(z, dtype=np.float32)
    B, C, H, W = z.shape
    x0, y0, wx1, wy1 = gather_indices(grid, H, W)
    x1 = np.minimum(x0 + 1, W - 1)
    y1 = np.minimum(y0 + 1, H - 1)
    wx0 = np.float32(1) - wx1
    wy0 = np.float32(1) - wy1
    out = np.empty((B, C) + x0.shape[1:], dtype=np.float32)
    for b in range(B):
        zb = z[b]
        nw = zb[:, y0[b], x0[b]]
        ne = zb[:, y0[b], x1[b]]
        sw = zb[:, y1[b], x0[b]]
        se = zb[:, y1[b], x1[b]]
        out[b] = nw * (wx0[b] * wy0[b]) + ne * (wx1[b] * wy0[b]) + sw * (wx0[b] * wy1[b]) + se * (wx1[b] * wy1[b])
    return out


def gather_surrogate_backward(grad_out):
    """GridSamplerFuncNoGrad.backward (grid_generator.py:615-623): 3x3 block mean * 0.1.
    The guarded all_reduce at :621-622 never runs under the reference's launcher and is NOT part of
    the op (deliberate deviation)."""
    g = np.asarray(grad_out)
    B, C, H, W = g.shape
    return g.reshape(B, C, H // 3, 3, W // 3, 3).mean(axis=(3, 5)) * g.dtype.type(0.1)


class _GatherNoGrad(torch.autograd.Function):
    """torch wrapper so the model-level oracle reproduces the surrogate backward."""

    @staticmethod
    def forward(ctx, z, grid):
        return F.grid_sample(z, grid, align_corners=True, mode="bilinear", padding_mode="border")

    @staticmethod
    def backward(ctx, go):
        B, C, H, W = go.shape
        return go.contiguous().reshape(B, C, H // 3, 3, W // 3, 3).mean(dim=[3, 5]) * 0.1, None


def gather_t(z, grid_t):
    return _GatherNoGrad.apply(z, grid_t)


# --------------------------------------------------------------------------------------------
# Linear / modulated convolutions        models/ops.py:190-222, 580-640; models/spgan_ops_gs.py:700-816
# --------------------------------------------------------------------------------------------
def equal_linear(x, weight, bias, lr_mul=1.0, activation=False):
    """models/ops.py:190-222."""
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        return fused_leaky_relu_t(F.linear(x, weight * scale), bias * lr_mul)
    return F.linear(x, weight * scale, bias=bias * lr_mul)


def pixel_norm(x):
    """models/ops.py:13-20."""
    return x * torch.rsqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-8)


def modulated_weight(weight, style_mod, demodulate):
    """weight (1, O, I, k, k); style_mod (B, I) → (B, O, I, k, k).  models/ops.py:598-607."""
    _, O, I, k, _ = weight.shape
    B = style_mod.shape[0]
    scale = 1 / math.sqrt(I * k * k)
    w = scale * weight * style_mod.view(B, 1, I, 1, 1)
    if demodulate:
        demod = torch.rsqrt(w.pow(2).sum([2, 3, 4]) + 1e-8)
        w = w * demod.view(B, O, 1, 1, 1)
    return w


def modulated_conv2d(x, style, weight, mod_weight, mod_bias, demodulate=True, upsample=False,
                     blur_kernel=None, padding=0):
    """Plain StyleGAN2 modulated conv, no_zero_pad variant: models/ops.py:580-640.

    upsample: conv_transpose2d stride 2 → crop 1 px each side (:617-619) → Blur (:622).
    """
    B, I, H, W = x.shape
    _, O, _, k, _ = weight.shape
    s = equal_linear(style, mod_weight, mod_bias)  # bias_init=1 lives in the parameter
    w = modulated_weight(weight, s, demodulate)
    if upsample:
        wt = w.transpose(1, 2).reshape(B * I, O, k, k)
        out = F.conv_transpose2d(x.reshape(1, B * I, H, W), wt, padding=0, stride=2, groups=B)
        out = out[:, :, 1:-1, 1:-1]
        out = out.reshape(B, O, out.shape[2], out.shape[3])
        out = upfirdn2d_t(out, blur_kernel, pad=(0, 0))
    else:
        out = F.conv2d(x.reshape(1, B * I, H, W), w.reshape(B * O, I, k, k), padding=padding, groups=B)
        out = out.reshape(B, O, out.shape[2], out.shape[3])
    return out


def encode_coords(c):
    """tanh / cos(pi.) / sin(pi.) on channels 0/1/2 (coord_num_dir == 3):
    models/spgan_ops_gs.py:799-802, coord_handler.py:696-711."""
    return torch.stack([torch.tanh(c[:, 0]), torch.cos(c[:, 1] * np.pi), torch.sin(c[:, 2] * np.pi)], 1)


def sphere_modulated_conv2d(x, coords, style, weight, mod_weight, mod_bias, grid, demodulate=True):
    """spgan_ops_gs.ModulatedConv2d.forward with deal_coords=True (models/spgan_ops_gs.py:700-816):
    gather x and raw coords at `grid` (B, 3H, 3W, 2), encode the gathered coords, concat,
    grouped conv k=3 stride=3 padding=0."""
    B, C, H, W = x.shape
    _, O, I, k, _ = weight.shape
    s = equal_linear(style, mod_weight, mod_bias)
    w = modulated_weight(weight, s, demodulate)
    xs = gather_t(x, grid)
    cs = encode_coords(gather_t(coords, grid))
    inp = torch.cat([xs.reshape(1, B * C, 3 * H, 3 * W), cs.reshape(1, B * 3, 3 * H, 3 * W)], 1)
    # NB: the reference concatenates along dim 1 of the (1, B*C, ..) views, i.e. all samples' feature
    # channels first and then all samples' coord channels (models/spgan_ops_gs.py:792-813).  With
    # groups=B the g-th group therefore reads channels [g*259, (g+1)*259) of that concatenation.
    out = F.conv2d(inp, w.reshape(B * O, I, k, k), padding=0, groups=B, stride=(3, 3))
    return out.reshape(B, O, out.shape[2], out.shape[3])


def sphere_rgb_conv(x, weight, bias, grid):
    """SphereConvBatchDiffFixBorderGNoGrad.forward (models/spherenet/sphere_conv2d.py:167-205):
    gather → conv2d(W / sqrt(Cin*9), bias, stride 3) → LeakyReLU(0.01)."""
    scale = 1 / math.sqrt(weight.shape[1] * 9)
    y = F.conv2d(gather_t(x, grid), weight * scale, bias, stride=(3, 3), padding=0)
    return F.leaky_relu(y, 0.01)


def upsample_skip(skip, kernel):
    """Upsample(no_zero_pad=True).forward (models/spgan_ops.py:54-59): depthwise conv_transpose FIR ×2, crop 1."""
    B, C, H, W = skip.shape
    out = F.conv_transpose2d(skip.reshape(B * C, 1, H, W), kernel.view(1, 1, *kernel.shape), stride=2)
    return out.reshape(B, C, out.shape[2], out.shape[3])[:, :, 1:-1, 1:-1]


def center_crop(src, ref_h, ref_w):
    """ToRGB.align_spatial_size (models/spgan_ops.py:1551-1562) / ImplicitFunction._select_center (spgan.py:199-206)."""
    ph = (src.shape[2] - ref_h) // 2
    pw = (src.shape[3] - ref_w) // 2
    return src[:, :, ph:ph + ref_h, pw:pw + ref_w]


# --------------------------------------------------------------------------------------------
# Generator                              models/spgan/spgan.py (structure + texture synthesiser)
# --------------------------------------------------------------------------------------------
TS_UPSAMPLE = [True, False, True, False, True, False, True, False]  # spgan.py:433-450
TO_RGB_SRC = [1, 3, 5, 7]                                           # spgan.py:451-456
SP_CONV_AT = {3: 0, 5: 1, 7: 2}                                     # spgan.py:686-691


def _grid_t(h, w, coords_partial, batch):
    return torch.from_numpy(batch_sampling_grid(h, w, coords_partial, batch))


def structure_synthesizer(sd, global_latent, local_latent, coords, coords_partial, prefix="structure_synthesizer."):
    """ImplicitFunction.forward (spgan.py:229-254) over 4 x (SphereConditionalBlock :159-169, ConditionalBlock :111-119).
    `global_latent` is the raw (B, 512) latent column 0: spgan.yaml has no ss_mapping, so it is used as the style."""
    h = local_latent
    B = h.shape[0]
    for i in range(4):
        p = prefix + "implicit_model.conv_stack.%d." % (2 * i)
        c = center_crop(coords, h.shape[2], h.shape[3])
        grid = _grid_t(h.shape[2], h.shape[3], coords_partial, B)
        y = sphere_modulated_conv2d(h, c, global_latent, sd[p + "conv.conv.weight"],
                                    sd[p + "conv.conv.modulation.weight"], sd[p + "conv.conv.modulation.bias"], grid)
        y = F.leaky_relu(y, 0.01)  # nn.LeakyReLU() default slope, "LeakyReLU_n" (spgan_ops_gs.py:1085-1086)
        h = y + F.conv2d(h, sd[p + "sc.weight"], sd[p + "sc.bias"])
        p = prefix + "implicit_model.conv_stack.%d." % (2 * i + 1)
        c = encode_coords(center_crop(coords, h.shape[2], h.shape[3]))
        y = modulated_conv2d(torch.cat([h, c], 1), global_latent, sd[p + "conv.conv.weight"],
                             sd[p + "conv.conv.modulation.weight"], sd[p + "conv.conv.modulation.bias"])
        h = fused_leaky_relu_t(y, sd[p + "conv.activate.bias"])  # ss_disable_noise: no NoiseInjection
    return h


def mapping_network(sd, z, prefix="texture_synthesizer.mapping."):
    """PixelNorm + 8 x EqualLinear(lr_mul 0.01, fused lrelu): spgan.py:405-412."""
    h = pixel_norm(z)
    for i in range(1, 9):
        h = equal_linear(h, sd[prefix + "%d.weight" % i], sd[prefix + "%d.bias" % i], lr_mul=0.01, activation=True)
    return h


def texture_synthesizer(sd, styles, structure_latent, coords_partial, noises, prefix="texture_synthesizer.",
                        return_intermediates=False):
    """TextureSynthesizer.forward synthesis loop (spgan.py:924-978). styles: (B, 9, 512) w-space."""
    h = structure_latent
    B = h.shape[0]
    skip = None
    rgb_idx = 0
    inter = []
    for i in range(8):
        p = prefix + "convs.%d." % i
        blur = sd.get(p + "conv.blur.kernel")
        y = modulated_conv2d(h, styles[:, i], sd[p + "conv.weight"], sd[p + "conv.modulation.weight"],
                             sd[p + "conv.modulation.bias"], upsample=TS_UPSAMPLE[i], blur_kernel=blur)
        y = y + sd[p + "noise.weight"] * noises[i]  # NoiseInjection (ops.py:784)
        h = fused_leaky_relu_t(y, sd[p + "activate.bias"])
        if return_intermediates:
            inter.append(h)
        if i == TO_RGB_SRC[rgb_idx]:
            tgt = [3, 5, 7, 8][rgb_idx]
            if i in SP_CONV_AT:
                q = prefix + "sp_convs.%d." % SP_CONV_AT[i]
                grid = _grid_t(skip.shape[2], skip.shape[3], coords_partial, B)
                skip = sphere_rgb_conv(skip, sd[q + "weight"], sd[q + "bias"], grid)
            q = prefix + "to_rgbs.%d." % rgb_idx
            rgb = modulated_conv2d(h, styles[:, tgt], sd[q + "conv.weight"], sd[q + "conv.modulation.weight"],
                                   sd[q + "conv.modulation.bias"], demodulate=False) + sd[q + "bias"]
            if skip is not None:
                up = upsample_skip(skip, sd[q + "upsample.kernel"])
                rgb = rgb + center_crop(up, rgb.shape[2], rgb.shape[3])
            skip = rgb
            rgb_idx += 1
            if rgb_idx == 4:
                break
    if return_intermediates:
        return skip, inter
    return skip


def generator_forward(sd, global_latent, local_latent, coords, coords_partial, noises, inject_index=None):
    """InfinityGanGenerator.forward for `override_coords` inputs (spgan.py:1278-1420, test-time call of
    close_loop_infinite_generation.py:262-270): global_latent (B, 2, 512), local_latent (B, 256, 35, 35),
    coords (B, 3, 35, 35) raw meta-coord slice, noises = 8 tensors (B, 1, h, w).  Returns (B, 3, 101, 101)."""
    struct = structure_synthesizer(sd, global_latent[:, 0], local_latent, coords, coords_partial)
    w0 = mapping_network(sd, global_latent[:, 0])
    w1 = mapping_network(sd, global_latent[:, 1])
    n_latent = 9
    if inject_index is None or inject_index == n_latent:
        styles = w0.unsqueeze(1).repeat(1, n_latent, 1)
    else:  # style mixing (spgan.py:872-876)
        styles = torch.cat([w0.unsqueeze(1).repeat(1, inject_index, 1),
                            w1.unsqueeze(1).repeat(1, n_latent - inject_index, 1)], 1)
    return texture_synthesizer(sd, styles, struct, coords_partial, noises)


# --------------------------------------------------------------------------------------------
# Close-loop panorama lattice            test_managers/base_test_manager.py:86-121,
#                                        test_managers/close_loop_infinite_generation.py:170-305,428-460
# --------------------------------------------------------------------------------------------
TS_FEATURE_SIZES = [19, 17, 31, 29, 55, 53, 103, 101]      # calc_out_spatial_size(11)  
TS_FEATURE_SIZES_X2 = [35, 33, 63, 61, 119, 117, 231, 229]  # calc_out_spatial_size(22)
TEST_META_EXTRA_PAD = 3                                    # test_managers/global_config.py:1


def ts_out_sizes(in_size):
    """TextureSynthesizer.calc_out_spatial_size(return_list=True): upsample → 2n-3, plain → n-2 (ops.py:339-349)."""
    sizes = []
    n = in_size
    for up in TS_UPSAMPLE:
        n = n * 2 - 1 - 2 if up else n - 2
        sizes.append(n)
    return sizes


def ts_in_sizes(out_size):
    """TextureSynthesizer.calc_in_spatial_size(return_list=True) (ops.py:315-336)."""
    sizes = []
    n = out_size
    for up in TS_UPSAMPLE[::-1]:
        if up:
            v = n + 1 + 2
            n = (v if v % 2 == 0 else v + 1) // 2
        else:
            n = n + 2
        sizes.append(n)
    return sizes[::-1]


def close_loop_plan(target_h, target_w, ts_input=11, ss_unfold=12, patch=101):
    """Patch lattice of the close-loop manager (base_test_manager.py:86-121; close_loop..py:428-460, 46-48)."""
    out1 = np.array(ts_out_sizes(ts_input))
    out2 = np.array(ts_out_sizes(ts_input * 2))
    out_disp = out2 - out1
    in1 = np.array(ts_in_sizes(out1[-1]))
    in2 = np.array(ts_in_sizes(out2[-1]))
    in_disp = in2 - in1
    unit = out_disp[-1] // ts_input
    pix_step = (out1[-1] // unit) * unit
    lat_step = pix_step // unit
    infeat_step = lat_step * (in_disp // ts_input)
    outfeat_step = lat_step * (out_disp // ts_input)
    steps_h = math.ceil((target_h - out1[-1]) / pix_step) + TEST_META_EXTRA_PAD
    assert target_w % pix_step == 0
    steps_w_min = math.ceil(target_w / pix_step)
    steps_w = steps_w_min + 2
    meta_h = pix_step * (steps_h - 1) + out1[-1]
    meta_w = steps_w_min * pix_step
    noise_h = outfeat_step * (steps_h - 1) + out1
    noise_w = outfeat_step * steps_w_min
    lat_h = ts_in_sizes(meta_h)[0] + 2 * ss_unfold
    lat_w = meta_w // int(infeat_step[-1]) * 6  # latent_sampler.py:212-213
    return dict(pix_step=int(pix_step), lat_step=int(lat_step), outfeat_step=[int(v) for v in outfeat_step],
                out_sizes=[int(v) for v in out1], steps_h=int(steps_h), steps_w=int(steps_w),
                steps_w_min=int(steps_w_min), meta_h=int(meta_h), meta_w=int(meta_w),
                noise_h=[int(v) for v in noise_h], noise_w=[int(v) for v in noise_w],
                lat_h=int(lat_h), lat_w=int(lat_w), ts_input=ts_input, ss_unfold=ss_unfold)


def circular_slice(t, width, x_st, x_ed, y_st, y_ed):
    """circular_sample_width (close_loop_infinite_generation.py:307-331)."""
    if y_ed <= width:
        return t[:, :, x_st:x_ed, y_st:y_ed]
    if y_ed <= width * 2:
        if y_st < width:
            return torch.cat((t[:, :, x_st:x_ed, y_st:], t[:, :, x_st:x_ed, :y_ed % width]), dim=3)
        return t[:, :, x_st:x_ed, y_st % width:y_ed % width]
    return circular_slice(t, width, x_st, x_ed, y_st - width, y_ed - width)


def circular_assign(t, width, x_st, x_ed, y_st, y_ed, v):
    """_circular_assign_value_width (base_test_manager.py:305-325)."""
    if y_ed <= width:
        t[:, :, x_st:x_ed, y_st:y_ed] = v
    elif y_ed <= width * 2:
        if y_st < width:
            d = width - y_st
            t[:, :, x_st:x_ed, y_st:] = v[:, :, :, :d]
            t[:, :, x_st:x_ed, :y_ed % width] = v[:, :, :, d:]
        else:
            t[:, :, x_st:x_ed, y_st % width:y_ed % width] = v
    else:
        circular_assign(t, width, x_st, x_ed, y_st - width, y_ed - width, v)


def meta_coord_grid(height, width, cut_pt=3.0, const_x=45, const_y=140):
    """SphereCoordHandlerV3BatchDiff._creat_coord_grid for the test path (coord_handler.py:575-607, 620-627):
    x = arange(h)/(const_x-1), re-centred, *2-1, *cut_pt;  y = arange(w)/(const_y-1)*2-1; channels (x, y, y)."""
    x = torch.arange(height).type(torch.float32) / (const_x - 1)
    y = torch.arange(width).type(torch.float32) / (const_y - 1)
    x = x - (x[-1] - 1) / 2
    x = (x * 2 - 1) * cut_pt
    y = y * 2 - 1
    xt = x.view(-1, 1).repeat(1, width)
    yt = y.view(1, -1).repeat(height, 1)
    return torch.stack([xt, yt, yt], 0)


def patch_coords_partial(plan, ix, iy, meta_h, meta_w, iiter, partial=0.6667):
    """The per-patch dict of close_loop_infinite_generation.py:204-261 plus the slice cursors."""
    ss = plan["ss_unfold"]
    zx_st = ix * plan["lat_step"] + ss
    zy_st = iy * plan["lat_step"] + ss
    zx_ed = zx_st + plan["ts_input"]
    zy_ed = zy_st + plan["ts_input"]
    zx_st -= ss
    zy_st -= ss
    zx_ed += ss
    zy_ed += ss
    x_size = zx_ed - zx_st + 1
    y_size = zy_ed - zy_st + 1
    cursors = (zx_st, zx_ed, zy_st, zy_ed)
    if zy_ed > meta_w:  # get_circular_flag (:462-472)
        if zy_st < meta_w:
            circ, zy = True, zy_st
        else:
            circ, zy = False, zy_st % meta_w
    else:
        circ, zy = False, zy_st
    cp = {
        "p_x_st": zx_st / meta_h, "p_x_ed": (zx_st + x_size) / meta_h,
        "p_y_st": zy / meta_w, "p_y_ed": (zy + y_size) / meta_w,
        "circular_flag": circ, "x_total": meta_h, "y_total": meta_w, "test_flag": True,
        "start_flag": iiter == 0, "h_step": zx_st // 6, "w_step": zy // 6, "y_st": zy, "y_ed": zy_ed,
        "partial": partial,
    }
    return cp, cursors


def generate_panorama(sd, plan, global_latent, local_latent, noises, positions=None, forward=None):
    """InfiniteGenerationManagerPatchCoordsCloseLoop.generate (close_loop..py:170-305): row-major patch loop,
    later patches overwrite the overlap, longitude wraps.  `positions` optionally restricts the loop
    (bounded CPU-baseline samples); `forward` lets the bench swap the per-patch generator call."""
    B = global_latent.shape[0]
    meta = torch.zeros(B, 3, plan["meta_h"], plan["meta_w"])
    lat_h, lat_w = local_latent.shape[2:]
    coords_full = meta_coord_grid(lat_h, lat_w).unsqueeze(0).repeat(B, 1, 1, 1)
    fwd = forward or generator_forward
    idx = [(a, b) for a in range(plan["steps_h"]) for b in range(plan["steps_w"])]
    for it, (ix, iy) in enumerate(idx):
        if positions is not None and (ix, iy) not in positions:
            continue
        cp, (zx_st, zx_ed, zy_st, zy_ed) = patch_coords_partial(plan, ix, iy, lat_h, lat_w, it)
        cur_lat = circular_slice(local_latent, lat_w, zx_st, zx_ed, zy_st, zy_ed)
        cur_coords = circular_slice(coords_full, lat_w, zx_st, zx_ed, zy_st, zy_ed)
        cur_noises = []
        for l in range(8):
            fx, fy = ix * plan["outfeat_step"][l], iy * plan["outfeat_step"][l]
            s = plan["out_sizes"][l]
            cur_noises.append(circular_slice(noises[l], plan["noise_w"][l], fx, fx + s, fy, fy + s))
        patch = fwd(sd, global_latent, cur_lat.contiguous(), cur_coords.contiguous(), cp, cur_noises)
        px, py = ix * plan["pix_step"], iy * plan["pix_step"]
        circular_assign(meta, plan["meta_w"], px, px + 101, py, py + 101, patch.detach())
    return meta


# --------------------------------------------------------------------------------------------
# Discriminator                          models/stylegan2discriminator.py
# --------------------------------------------------------------------------------------------
def equal_conv2d(x, weight, bias=None, stride=1, padding=0):
    """EqualConv2d.forward (models/ops.py:172-182)."""
    scale = 1 / math.sqrt(weight.shape[1] * weight.shape[2] ** 2)
    return F.conv2d(x, weight * scale, bias=bias, stride=stride, padding=padding)


def d_conv_layer(sd, p, x, kernel_size, downsample=False, activate=True, bias=True):
    """ConvLayer (stylegan2discriminator.py:9-54): [Blur] -> EqualConv2d -> [FusedLeakyReLU].  `p` = key prefix."""
    i = 0
    if downsample:
        pd = (4 - 2) + (kernel_size - 1)
        k = sd[p + "0.kernel"]
        x = upfirdn2d_t(x, k, pad=((pd + 1) // 2, pd // 2))
        i = 1
        x = equal_conv2d(x, sd[p + "%d.weight" % i], None if activate or not bias else sd.get(p + "%d.bias" % i), stride=2, padding=0)
    else:
        x = equal_conv2d(x, sd[p + "%d.weight" % i], None if activate or not bias else sd.get(p + "%d.bias" % i),
                         padding=kernel_size // 2)
    if activate:
        x = fused_leaky_relu_t(x, sd[p + "%d.bias" % (i + 1)])
    return x


def discriminator_forward(sd, img, stddev_group=16):
    """StyleGan2Discriminator.forward (stylegan2discriminator.py:185-229) with coord_use_ac (spgan.yaml):
    returns (d_patch (B, 1), ac_coords_pred (B, 3))."""
    h = d_conv_layer(sd, "convs.0.", img, 1)
    for i in range(1, 6):  # ResBlocks 256->512->512->512->512->512 at 101, 50, 25, 12, 6 -> 3
        p = "convs.%d." % i
        out = d_conv_layer(sd, p + "conv1.", h, 3)
        out = d_conv_layer(sd, p + "conv2.", out, 3, downsample=True)
        skip = d_conv_layer(sd, p + "skip.", h, 1, downsample=True, activate=False, bias=False)
        h = (out + skip) / math.sqrt(2)
    batch, channel, height, width = h.shape
    group = min(batch, stddev_group)
    stddev = h.view(group, -1, 1, channel, height, width)
    stddev = torch.sqrt(stddev.var(0, unbiased=False) + 1e-8)
    stddev = stddev.mean([2, 3, 4], keepdims=True).squeeze(2)
    stddev = stddev.repeat(group, 1, height, width)
    h = torch.cat([h, stddev], 1)
    out = d_conv_layer(sd, "final_conv.", h, 3).view(batch, -1)
    d = equal_linear(equal_linear(out, sd["final_linear.0.weight"], sd["final_linear.0.bias"], activation=True),
                     sd["final_linear.1.weight"], sd["final_linear.1.bias"])
    ac = equal_linear(equal_linear(out, sd["coord_linear.0.weight"], sd["coord_linear.0.bias"], activation=True),
                      sd["coord_linear.1.weight"], sd["coord_linear.1.bias"])
    return d, ac


def d_r1_penalty(real_pred, real_img):
    """models/losses.py:36-41."""
    g, = torch.autograd.grad(real_pred.sum(), real_img, create_graph=True)
    return g.pow(2).reshape(g.shape[0], -1).sum(1).mean()


def styled_conv(x, style, weight, mod_weight, mod_bias, noise, noise_weight, act_bias, upsample=False, blur_kernel=None):
    """ops.StyledConv.forward (models/ops.py:853-863): modulated conv -> NoiseInjection -> FusedLeakyReLU."""
    y = modulated_conv2d(x, style, weight, mod_weight, mod_bias, upsample=upsample, blur_kernel=blur_kernel)
    if noise is not None:
        y = y + noise_weight * noise
    return fused_leaky_relu_t(y, act_bias)
