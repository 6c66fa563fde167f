/*
 * spgan_b200.h — C ABI of the B200-native SP-GAN conv hot path (libspgan_b200.so).
 *
 * The reference (chronos123/SP-GAN-TIP2025) has no FFI of its own for this path: its boundary is two
 * pybind11 ops plus PyTorch library calls.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference root).  Conventions, matching
 * models/custom_ops/fused_bias_act_kernel.cu:52-99 and upfirdn2d_kernel.cu:209-369:
 *   - every pointer is a DEVICE pointer to contiguous fp32 (int32 where stated); NULL = "absent tensor"
 *     (the reference passes an empty tensor for that);
 *   - outputs are caller-allocated and fully overwritten;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises;
 *   - return value 0 = success, non-zero = failure with a message in spgan_last_error() (the Python host
 *     raises RuntimeError, as the reference's TORCH_CHECK / CUDA errors do);
 *   - thread-safe as far as streams are: no global mutable state besides the per-thread error string.
 * There is no CPU fallback anywhere behind this ABI.
 */
#ifndef SPGAN_B200_H
#define SPGAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPGAN_ABI_VERSION 2
#define SPGAN_MAX_TAPS 49

int spgan_abi_version(void);
const char* spgan_last_error(void);
/* 1 when a CUDA device with compute capability 10.x is current, else 0 (message in spgan_last_error). */
int spgan_device_ok(void);

/* ---- K1: fused bias + activation -------------------------------------------------------------------
 * Replaces fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *   (models/custom_ops/fused_bias_act.cpp:11-20, fused_bias_act_kernel.cu:18-49).
 * out[i] = f(x[i] + bias[(i / step_b) % size_b]) * scale ; act*10+grad: 10/11 linear, 30 lrelu,
 * 31 lrelu gradient gated by sign of ref[i], 12/32 zero.  bias/ref may be NULL. */
int spgan_bias_act(float* out, const float* x, const float* bias, const float* ref, int64_t n, int64_t step_b,
                   int64_t size_b, int act, int grad, float alpha, float scale, void* stream);
/* Fused first-order backward of fused_leaky_relu (models/custom_ops/fused_act.py:24-44): grad_in = K1(act 3, grad 1)
 * and grad_bias[c] = sum over batch and pixels of grad_in.  x is (batch, channels, inner). grad_bias is overwritten. */
int spgan_bias_act_bwd(float* grad_in, float* grad_bias, const float* grad_out, const float* out_ref, int64_t batch,
                       int64_t channels, int64_t inner, float alpha, float scale, void* stream);
/* NoiseInjection + FusedLeakyReLU in one pass (models/ops.py:784 then models/custom_ops/fused_act.py:56-64):
 * out[b,c,p] = lrelu(x[b,c,p] + noise_w[0] * noise[b,p] + bias[c], alpha) * scale.  noise (batch, inner) and noise_w (1)
 * may be NULL together, bias (channels) may be NULL. */
int spgan_noise_bias_act(float* out, const float* x, const float* noise, const float* noise_w, const float* bias,
                         int64_t batch, int64_t channels, int64_t inner, float alpha, float scale, void* stream);

/* ---- K2/K3: upfirdn2d ---------------------------------------------------------------------------------
 * Replaces upfirdn2d_op.upfirdn2d(input[N,H,W,1], kernel[kh,kw], up_x, up_y, down_x, down_y, pad_x0..pad_y1)
 *   (models/custom_ops/upfirdn2d.cpp:12-22, upfirdn2d_kernel.cu:209-369) with minor == 1, i.e. `planes` = B*C
 *   independent (in_h, in_w) images.  out is (planes, out_h, out_w) with
 *   out_h = (in_h*up_y + pad_y0 + pad_y1 - kh) / down_y + 1 (same for w).  kernel is a DEVICE pointer. */
int spgan_upfirdn2d(float* out, const float* x, const float* kernel, int64_t planes, int in_h, int in_w, int kh, int kw,
                    int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                    void* stream);

/* Host-only: the launch plan spgan_upfirdn2d would use for a 16-byte aligned input of these sizes (no device work; callable
 * without a GPU).  plan[20] = { streamed variant K*100 + up*10 + down (0 = band / tiled / polyphase / generic kernel),
 *   planes per item, bands per plane, output rows per band, 8-row strips per band, thread groups, threads, floats per
 *   stage, interior columns, first interior column, border columns, border thread groups, first border thread, items, grid,
 *   out_h, out_w, stage capacity (floats), thread limit, strip rows }. */
int spgan_upfirdn2d_plan(int64_t planes, int in_h, int in_w, int kh, int kw, int up, int down, int pad_x0, int pad_x1,
                         int pad_y0, int pad_y1, int32_t* plan);

/* Fused tail of the upsampling StyledConv (models/ops.py:617-622 + 784 + fused_act.py:56-64): the cropped
 * conv_transpose2d(stride 2) output is held as four polyphase planes pp (planes, 4, Hq, Wq), plane index
 * (Y & 1) * 2 + (X & 1), element (Y >> 1, X >> 1); this kernel applies the 3x3 FIR of Blur (upfirdn2d, pad 0; `kernel`
 * is the un-flipped 3x3 buffer on the device), adds noise_w * noise (batch, out_h, out_w) and bias (channels), and the
 * leaky-ReLU * scale:  out (batch*channels, zh - 2, zw - 2) where (zh, zw) is the interleaved size. */
int spgan_upblur_act(float* out, const float* pp, const float* kernel, const float* noise, const float* noise_w,
                     const float* bias, int64_t batch, int64_t channels, int zh, int zw, int Hq, int Wq, float alpha,
                     float scale, void* stream);

/* ---- L1: spherical bilinear gather ----------------------------------------------------------------------
 * Replaces F.grid_sample(z, grid, mode='bilinear', padding_mode='border', align_corners=True) as called by
 *   GridSamplerFuncNoGrad.forward (models/spherenet/grid_generator.py:610-613) for 3x3 tap grids.
 * z (B, C, H, W); grid (Bg, 3H, 3W, 2) fp32 with Bg == B or Bg == 1 (test mode: one grid shared by the batch);
 * out plane (b, c) is written at out + ((b * out_bstride) + out_coff + c) * 9*H*W, which lets the caller lay
 * the result out exactly like the reference's flat (1, B*C + B*3, 3H, 3W) concatenation
 * (models/spgan_ops_gs.py:791-813).  encode != 0 applies the coordinate encoding of :799-802 to channels
 * 0/1/2 (tanh, cos(pi x), sin(pi x)) after sampling. */
int spgan_sphere_gather(float* out, const float* z, const float* grid, int B, int C, int H, int W, int grid_batch,
                        int64_t out_bstride, int64_t out_coff, int encode, void* stream);

/* Host-only: the launch plan spgan_sphere_gather would use for these sizes (no device work; callable without a GPU).
 * plan[12] = { streamed (0 = L1-gather fallback), channels per item, channel groups per sample, position slices,
 *   positions per slice, staged floats, dynamic shared-memory bytes, resident CTAs per SM, grid, items,
 *   shared-memory limit, threads }. */
int spgan_sphere_gather_plan(int B, int C, int H, int W, int encode, int32_t* plan);
/* Corner indices and weights exactly as ATen computes them (GridSampler.h:27-36, 58-60); used by the
 * bit-exactness tests.  grid (n, 2); x0,y0 int32 (n); wx,wy fp32 (n) = weight of the +1 corner. */
int spgan_sphere_gather_indices(int32_t* x0, int32_t* y0, float* wx, float* wy, const float* grid, int64_t n, int H,
                                int W, void* stream);
/* Surrogate backward of the gather (grid_generator.py:615-623): grad_in = mean over each 3x3 block * 0.1.
 * grad_out (planes, 3H, 3W) → grad_in (planes, H, W).  The guarded all_reduce of :621-622 is deliberately absent. */
int spgan_sphere_gather_bwd(float* grad_in, const float* grad_out, int64_t planes, int H, int W, void* stream);

/* General 2-D grid samplers with their true input gradients (samplers the spgan.yaml generator does not instantiate but
 * whose signatures the drop-in keeps): mode 0 = F.grid_sample(bilinear, border, align_corners=True) (GridSamplerNew,
 * models/spherenet/grid_generator.py:588-592); mode 1 = grid_sample_github, bilinear weights from the unclipped coordinate
 * and clamped corner indices (GridSamplerNewTexture, models/spherenet/grid_sample_ops.py:5-55); mode 2 =
 * F.grid_sample(nearest, zeros, align_corners=True) (GridSampler, models/spherenet/grid_sample_grad_fix.py:29-48).
 * z (B, C, IH, IW), grid (Bg, OH, OW, 2) with Bg in {1, B}, out (B, C, OH, OW).  spgan_grid_sample_bwd overwrites grad_z
 * (B, C, IH, IW) with the transposed scatter of grad_out (aten::grid_sampler_2d_backward's grad_input / autograd of
 * torch.gather); the op is linear in z, so its double backward is spgan_grid_sample itself. */
int spgan_grid_sample(float* out, const float* z, const float* grid, int B, int C, int IH, int IW, int OH, int OW,
                      int grid_batch, int mode, void* stream);
int spgan_grid_sample_bwd(float* grad_z, const float* grad_out, const float* grid, int B, int C, int IH, int IW, int OH,
                          int OW, int grid_batch, int mode, void* stream);

/* Per-sample training grids assembled on the device (replaces the per-sample, per-layer host rebuild of
 * models/spgan_ops_gs.py:767-781 -> models/spherenet/grid_generator.py:137-283).  The grid separates into a row factor
 * (lat_n (slots_x, H, 9) fp32 final latitude coordinate; lon (slots_x, H, 9) fp64 tangent-plane longitude offset) and a
 * column factor (nlon (slots_y, W) fp64 normalised column base); the host computes each factor once per distinct window
 * edge with numpy float64 and keeps it on the device.  ix / iy (B) int32 pick the slots of each sample.  The kernel redoes
 * the reference's remaining float64 add / divide / multiply sequence with round-to-nearest intrinsics: out
 * (B, 3H, 3W, 2) fp32 equals the host-built grid bit for bit. */
int spgan_sphere_grid_assemble(float* out, const float* lat_n, const double* lon, const double* nlon, const int32_t* ix,
                               const int32_t* iy, int B, int H, int W, double y_total, void* stream);

/* ---- L7: EqualLinear ------------------------------------------------------------------------------------
 * Replaces F.linear(x, W * w_scale, bias * b_scale) [+ fused_leaky_relu] (models/ops.py:213-218).
 * x (M, K), w (N, K), bias (N) or NULL, y (M, N).  act: 0 none, 1 = leaky-relu(alpha) * gain applied after bias. */
int spgan_linear(float* y, const float* x, const float* w, const float* bias, int M, int N, int K, float w_scale,
                 float b_scale, int act, float alpha, float gain, void* stream);

/* Weight gradient of the same layer (autograd of F.linear, models/ops.py:213-218): dw (N, K) = scale * g^T x with
 * g (M, N) the output gradient and x (M, K) the input, M = batch. */
int spgan_linear_wgrad(float* dw, const float* g, const float* x, int M, int N, int K, float scale, void* stream);

/* ---- L2-L6: convolution passes ---------------------------------------------------------------------------
 * One "pass" computes, for every sample b, output channel o and lattice point (i, j), 0<=i<My, 0<=j<Mx:
 *   acc = sum_c sum_t  w[o*ws_o + c*ws_c + tap_w[t]] * in_mul[b,c] * x[b, c, i*in_stride + tap_dy[t], j*in_stride + tap_dx[t]]
 *         (taps falling outside [0,H)x[0,W) contribute zero)
 *   v   = acc * out_scale * out_mul[b,o]  + noise_w[0]*noise[b, Y, X] + bias[o]
 *   v   = act ? (v > 0 ? v : v*alpha) * gain : v
 *   y[b, o, Y, X] = v + residual[b, o, Y, X],      Y = i*out_stride + out_off_y,  X = j*out_stride + out_off_x.
 * With tap lists this covers F.conv2d(groups=batch) k in {1,3,7} / stride {1,2,3} (models/ops.py:634,175;
 * models/spgan_ops_gs.py:814), F.conv_transpose2d(stride=2, groups=batch) as four parity passes
 * (models/ops.py:617), their data gradients, and the non-modulated convs of the discriminator.
 * in_mul (B, Cin) carries the style modulation, out_mul (B, Cout) the demodulation (models/ops.py:598-607):
 * the per-sample weight tensor is never materialised.  NULL pointers switch a term off. */
typedef struct SpganConvPass {
  int32_t B, Cin, H, W;           /* input tensor (B, Cin, H, W) */
  int32_t Cout, out_H, out_W;     /* output tensor (B, Cout, out_H, out_W) */
  int32_t My, Mx;                 /* lattice extent of this pass */
  int32_t in_stride;              /* input step per lattice step */
  int32_t out_stride, out_off_y, out_off_x;
  int32_t ntaps;
  int32_t tap_dy[SPGAN_MAX_TAPS], tap_dx[SPGAN_MAX_TAPS], tap_w[SPGAN_MAX_TAPS];
  int64_t ws_o, ws_c;             /* weight strides (elements) for output / input channel */
  float out_scale;
  int32_t act;                    /* 0 = none, 1 = leaky relu */
  float act_alpha, act_gain;
  int32_t precision;              /* 0 = fp32 SIMT, 1 = bf16x3 split on tcgen05 (fp32-equivalent), 2 = bf16 on tcgen05,
                                     3 = fp16x2 on tcgen05 (fp16 hi+lo activations x fp16 weights, relative error ~2^-12;
                                     activations beyond +-65504 saturate: the host pre-scales them by a power of two) */
  int64_t out_cstride;            /* elements between output channels; 0 = out_H*out_W (dense NCHW).  A larger stride
                                     lets the parity passes of a transposed conv write polyphase planes
                                     (B, Cout, s*s, Hq, Wq) instead of scattering with stride s (see spgan_upblur_act) */
} SpganConvPass;

int spgan_conv_pass(const SpganConvPass* p, float* y, const float* x, const float* w, const float* in_mul,
                    const float* out_mul, const float* noise, const float* noise_w, const float* bias,
                    const float* residual, void* stream);

/* Demodulation coefficients (models/ops.py:603-604): d[b,o] = rsqrt(scale^2 * sum_c s[b,c]^2 * sum_t w[o,c,t]^2 + eps).
 * w (Cout, Cin, taps), s (B, Cin), d (B, Cout). */
int spgan_demod(float* d, const float* s, const float* w, int B, int Cin, int Cout, int taps, float scale, float eps,
                void* stream);

/* Weight gradient of a conv pass: dw[o*ws_o + c*ws_c + tap_w[t]] (+)= sum_b sum_ij
 *   g[b,o,Y,X] * out_mul[b,o] * out_scale * in_mul[b,c] * x[b,c,i*in_stride+dy_t, j*in_stride+dx_t].
 * Replaces cuDNN wgrad of the same F.conv2d / F.conv_transpose2d calls.  accumulate != 0 adds into dw.  Partial sums over the
 * batch and over pixel splits are combined with atomicAdd: results are reproducible to fp32 rounding, not bit for bit (the
 * tcgen05 weight gradient below is deterministic). */
int spgan_conv_wgrad(const SpganConvPass* p, float* dw, const float* g, const float* x, const float* in_mul,
                     const float* out_mul, int accumulate, void* stream);

/* ---- tcgen05 / TMEM implicit-GEMM path (precision 1 = bf16x3 split, fp32-equivalent; 2 = plain bf16) ---------
 * The dense contractions run as  Y[p, o] = sum_t sum_k A[p + off_t, k] * Wp[t][o][k]  over the FLATTENED pixel index
 * p = (b*Hl + i)*Wl + j of a channels-last bf16 copy of the input ("packed activation"), off_t = dy_t*Wl + dx_t:
 * a tap shift is a constant row offset, so every A tile is one TMA box and the conv is a plain K-major GEMM on
 * tcgen05.mma with the accumulator in TMEM.  Lattice points whose window would wrap a row are computed and dropped
 * in the epilogue (the pass's My/Mx bounds).  Operands are stored as two bf16 planes, hi = bf16(v) and
 * lo = bf16(v - hi); precision 1 issues hi*hi + hi*lo + lo*hi (relative error ~2^-16), precision 2 only hi*hi.
 *
 * spgan_pack_act: x (B, C, H, W) fp32 [* in_mul (B, C)] -> out [2][B*Hl*Wl][Cp] bf16; image pixel (y, x) lands on
 *   lattice point (y + pad_y0, x + pad_x0) of the (Hl, Wl) lattice, everything else (borders, channels C..Cp-1,
 *   Cp a multiple of 16) is zero.  Carries the style modulation of models/ops.py:598-600 (applied to the
 *   activations instead of the weights).
 *   step s > 1 writes s*s polyphase planes, out [2][s*s][B*Hl*Wl][Cp]: lattice point (i, j) of phase py*s + px holds image
 *   pixel (i*s + py - pad_y0, j*s + px - pad_x0).  A strided conv (the discriminator's stride-2 convs, models/ops.py:175;
 *   the stride-3 conv behind the spherical gather, models/spgan_ops_gs.py:814) is then a stride-1 conv whose taps pick
 *   their phase plane.
 *   fmt selects the 16-bit planes: 0 = bf16 hi/lo (precision 1 and 2), 1 = fp16 hi/lo (precision 3; values saturate at
 *   +-65504). */
int spgan_pack_act(uint16_t* out, const float* x, const float* in_mul, int B, int C, int H, int W, int Cp, int pad_y0,
                   int pad_x0, int Hl, int Wl, int step, int fmt, void* stream);
/* spgan_pack_weight: w[o*ws_o + c*ws_c + tap_w[t]] fp32 -> out [2][ntaps][Cout][Cp] bf16 (merged == 0) or
 *   [2][1][Cout][ntaps*Cp] (merged != 0, k = t*Cp + c: the layout that pairs with spgan_sphere_pack). */
int spgan_pack_weight(uint16_t* out, const float* w, int Cout, int Cin, int64_t ws_o, int64_t ws_c, int ntaps,
                      const int32_t* tap_w, int Cp, int merged, int fmt, void* stream);
/* spgan_nchw_to_nhwc: (B, C, H, W) fp32 -> (B, H, W, C) fp32; staging copy that makes the spherical gather coalesced. */
int spgan_nchw_to_nhwc(float* out, const float* x, int B, int C, int H, int W, void* stream);
/* spgan_sphere_pack: the fused A-operand producer of the spherical modulated conv
 *   (models/spgan_ops_gs.py:791-813 + grid_generator.py:610-613): bilinear border gather of the features
 *   (x_nhwc, (B, H, W, C) fp32) and of the raw coords ((B, 3, H, W) fp32, may be NULL) at the 3x3 tap grid
 *   ((Bg, 3H, 3W, 2), Bg in {1, B}), coordinate encoding tanh / cos(pi.) / sin(pi.), style modulation
 *   in_mul (B, C + nc) (may be NULL), bf16 hi/lo split -> out [2][B*H*W][9*Cp], k = tap*Cp + channel.
 *   chan_map (B, Cp) uint32 on the device says which gathered plane feeds channel k of group g:
 *   bit 31 = coordinate plane, bits [15,31) = source sample, bits [0,15) = source channel, 0xFFFFFFFF = zero padding.
 *   The reference concatenates the gathered tensors as flat (1, B*C) ++ (1, B*nc) channel lists and then convolves
 *   with groups = B, so group g reads flat channels [g*(C+nc), (g+1)*(C+nc)) (for B > 1 that mixes samples); the host
 *   encodes exactly that mapping (or the per-sample concatenation) in the table. */
int spgan_sphere_pack(uint16_t* out, const float* x_nhwc, const float* coords, const float* grid, const float* in_mul,
                      const uint32_t* chan_map, int B, int C, int H, int W, int grid_batch, int Cp, int fmt, void* stream);
/* spgan_sphere_pack_seg: spgan_sphere_pack with the K layout of the structure synthesiser's chain: the first Cm channels of
 *   every group (Cm a multiple of 64, 256 for spgan.yaml) go to out [2][B*H*W][9*Cm], k = tap*Cm + channel, and the trailing
 *   Cx = C + nc - Cm channels (the coordinate planes, or the reference's flat-concat spill-over) of ALL nine taps go to the
 *   dense tail operand out2 [2][B*H*W][kp2], k2 = tap*Cx + j (columns 9*Cx..kp2-1 zero), which spgan_conv_gemm_ex consumes as
 *   its second K segment.  grid is (B / grid_group, 3H, 3W, 2): samples [i*grid_group, (i+1)*grid_group) share grid i, so
 *   several lattice positions of a panorama (close_loop_infinite_generation.py:185-261) run as one batch; chan_map rows are
 *   cmap_ld entries apart.  scratch: fp32 workspace of spgan_sphere_pack_seg_scratch(B, C, H, W) elements (may be NULL: slower
 *   scalar producer) that receives the flat-concat channels of every group as 16-byte aligned rows, so that the gather itself
 *   runs on 128-bit loads. */
int64_t spgan_sphere_pack_seg_scratch(int B, int C, int H, int W);
/* spgan_sphere_concat_repack: the reference's flat (1,B*C) ++ (1,B*3) concatenation made explicit once per layer:
 *   xg[(g*H*W + p)*264 + k] = the k-th channel group g reads (chan_map[g][k]: feature plane of some sample, or RAW coordinate
 *   plane), zero for k >= C + nc; 16-byte aligned rows, so that the gather of spgan_sphere_pack_seg / spgan_sphere_conv_gemm
 *   runs on aligned 128-bit loads.  C must be 256; xg holds spgan_sphere_pack_seg_scratch(B, C, H, W) floats. */
int spgan_sphere_concat_repack(float* xg, const float* x_nhwc, const float* coords, const uint32_t* chan_map, int B, int C,
                               int H, int W, int cmap_ld, void* stream);
int spgan_sphere_pack_seg(uint16_t* out, uint16_t* out2, const float* x_nhwc, const float* coords, const float* grid,
                          const float* in_mul, const uint32_t* chan_map, int B, int C, int H, int W, int grid_group, int Cm,
                          int cmap_ld, int kp2, int fmt, float* scratch, void* stream);
/* spgan_coord_taps_pack: the tail operand of an unpadded kh x kw conv whose last nc input channels are the encoded coordinate
 *   planes (ConditionalBlock, models/spgan/spgan.py:100-101: torch.cat([x, encode(coords)]) -> 7x7 StyledConv):
 *   out2 [2][B*My*Mx][kp2], row (b, y, x) of the My x Mx = (H-kh+1) x (W-kw+1) outputs, k2 = tap*nc + j,
 *   value = enc_j(coords[b, j, y+ty, x+tx]) * in_mul[b*mul_ld + mul_off + j]  (enc = tanh, cos(pi.), sin(pi.);
 *   coord_handler.py:696-711), 16-bit hi/lo split as in spgan_pack_act. */
int spgan_coord_taps_pack(uint16_t* out2, const float* coords, const float* in_mul, int B, int nc, int H, int W, int kh,
                          int kw, int mul_ld, int mul_off, int kp2, int fmt, void* stream);
/* spgan_conv_gemm: the tcgen05 kernel.  `p` is a conv pass whose (H, W) are the LATTICE dims (Hl, Wl) of the packed
 *   activation, Cin is ignored (K per tap = kp, a multiple of 16), in_stride must be 1, tap_w is ignored (the packed
 *   weight is already in tap order) and precision must be 1 or 2.  a_packed [2][a_rows][kp], w_packed
 *   [2][ntaps][Cout][kp].  a_rows = phases * B*H*W; a tap reads phase plane f by carrying f * B*H in its tap_dy (the row
 *   offset of a tap is tap_dy*W + tap_dx).  When the pass has lattice points that are not outputs (My < H or Mx < W, e.g.
 *   every unpadded 3x3 / 7x7 conv) the A tiles are fetched with TMA im2col loads over the (kp, W, H, 2*phases*B) view of
 *   a_packed: M then runs over the B*My*Mx outputs only, no MMA is spent on wrap-around points.  Epilogue terms as in
 *   spgan_conv_pass. */
int spgan_conv_gemm(const SpganConvPass* p, float* y, const uint16_t* a_packed, int64_t a_rows, int kp,
                    const uint16_t* w_packed, const float* out_mul, const float* noise, const float* noise_w,
                    const float* bias, const float* residual, void* stream);
/* spgan_conv_gemm_ex: the same kernel with every epilogue sink exposed.  Besides (or instead of) the NCHW fp32 tensor the
 *   epilogue can write, from the accumulator it already holds channels-last (TMEM lane = pixel, column = channel):
 *     - y with y_layout = 1: channels-last fp32, element (b, Y, X, o) at y[b*y_bstride + (Y*out_W + X)*Cout + o] (the
 *       parity passes of the transposed conv write their polyphase planes this way for spgan_upblur_pack);
 *     - y_packed: the NEXT conv's packed operand [2][y_packed_rows][y_packed_cols] (16-bit hi/lo planes in y_packed_fmt),
 *       row (b*out_H + Y)*out_W + X, already multiplied by that conv's style modulation next_mul (B, Cout): what
 *       spgan_pack_act would compute from the fp32 tensor, bit for bit, without the tensor or the pass;
 *     - rgb_w / rgb_part: ToRGB (1x1 modulated conv without demodulation, models/spgan_ops.py:1563-1586) folded in: rgb_w
 *       (B, rgb_n, Cout) = scale * w_rgb[j][o] * s_rgb[b][o]; every (N tile, epilogue half) writes its partial sum over
 *       its columns to slot 2*tile_n + half of rgb_part (slots, B, rgb_n, out_H*out_W), slots =
 *       spgan_conv_gemm_rgb_slots(p, a_rows); spgan_rgb_tail adds the slots in a fixed order (deterministic).
 *   y may be NULL when another sink is present (the last texture layer is consumed by ToRGB alone).  The extra sinks need
 *   Cout % 32 == 0 and exclude `residual`.  fmt / w_fmt are the 16-bit formats of a_packed / w_packed (0 = bf16 planes,
 *   1 = fp16 planes) and must match p->precision: (0, 0) for 1 and 2, (1, 1) for 3.  Mixed formats inside one MMA are
 *   rejected by the hardware (illegal instruction, measured), hence no bf16-activation x fp16-weight mode. */
typedef struct SpganGemmIO {
  const uint16_t* a_packed;
  int64_t a_rows;
  int32_t kp;
  int32_t fmt;
  const uint16_t* w_packed;
  int64_t w_fmt;
  const float* out_mul;
  const float* noise;
  const float* noise_w;
  const float* bias;
  const float* residual;
  float* y;
  int32_t y_layout;
  int32_t rgb_n;
  int64_t y_bstride;
  uint16_t* y_packed;
  const float* next_mul;
  int64_t y_packed_rows;
  int32_t y_packed_cols;
  int32_t y_packed_fmt;
  const float* rgb_w;
  float* rgb_part;
  const float* residual_nhwc;   /* channels-last residual (b, Y, X, o), added after the activation; needs the general sinks */
  int64_t res_bstride;          /* elements between samples of residual_nhwc; 0 = out_H*out_W*Cout */
  /* optional second K segment, Y[p, o] += sum_k a2_packed[p, k] * w2_packed[o, k]: a2_packed [2][a2_rows][kp2] is indexed by
   * the pass's own M row p (a lattice point in flat mode, the p-th output (b, i, j) in im2col mode), w2_packed [2][Cout][kp2],
   * kp2 a multiple of 64, same 16-bit formats as the main operands.  It carries the channels that do not fill a 64-wide K block
   * per tap (the 3 coordinate planes behind the 256 features of the structure synthesiser's 259-channel convs,
   * models/spgan/spgan.py:100, models/spgan_ops_gs.py:812) for ALL taps in one dense slab, so the main operand keeps
   * kp = 256 instead of padding every tap to 320. */
  const uint16_t* a2_packed;
  int64_t a2_rows;
  const uint16_t* w2_packed;
  int32_t kp2;
  int32_t reserved0;
} SpganGemmIO;
int spgan_conv_gemm_ex(const SpganConvPass* p, const SpganGemmIO* io, void* stream);
int spgan_conv_gemm_rgb_slots(const SpganConvPass* p, int64_t a_rows);

/* spgan_sphere_conv_gemm: the spherical modulated conv of models/spgan_ops_gs.py:791-816 as ONE kernel — the bilinear border
 *   gather at the 3x3 tangent taps (grid_generator.py:610-613), the coordinate encoding, the reference's flat-concat channel
 *   table and the style modulation run in the producer warps of the tcgen05 GEMM and write the A operand straight into
 *   swizzled shared memory: neither the reference's 9x gathered fp32 tensor nor spgan_sphere_pack's [B*H*W][9*Cp] 16-bit
 *   operand exists in HBM.  `in` describes the gather source exactly as spgan_sphere_pack does (x_nhwc (B, H, W, C) fp32,
 *   coords (B, 3, H, W) or NULL, grid (1, 3H, 3W, 2) shared by the batch, in_mul (B, C + nc) or NULL, chan_map (B, Cp)); `p`
 *   gives B, H, W (output lattice = input image), Cout, out_scale, activation and precision; `io` carries the packed weight
 *   (w_packed [2][1][Cout][9*Cp] in the merged layout of spgan_pack_weight, kp = 9*Cp, a_packed unused) and every epilogue
 *   term / sink of spgan_conv_gemm_ex.  Needs H*W >= 128 (smaller images: spgan_sphere_pack + spgan_conv_gemm). */
typedef struct SpganSphereIn {
  const float* x_nhwc;
  const float* coords;
  const float* grid;
  const float* in_mul;
  const uint32_t* chan_map;
  int32_t C;
  int32_t Cp;
  /* Optional: the REPACKED gather source of spgan_sphere_concat_repack (xg != NULL selects the vectorised producer: C = Cp = 256
   * main columns per tap, io->kp = 9*256 in the merged weight layout, the trailing C + nc - 256 channels of all taps in the
   * tail weight io->w2_packed with io->kp2 = 64, Cout <= 256).  x_nhwc is then unused; coords only says whether there are
   * coordinate planes; samples [i*grid_group, (i+1)*grid_group) share grid i; chan_map rows are cmap_ld entries apart. */
  const float* xg;
  int32_t grid_group;
  int32_t cmap_ld;
} SpganSphereIn;
int spgan_sphere_conv_gemm(const SpganConvPass* p, const SpganSphereIn* in, const SpganGemmIO* io, void* stream);

/* Channels-last tail of the upsampling StyledConv (same arithmetic as spgan_upblur_act): pp (batch, 4, Hq, Wq, channels)
 * fp32 polyphase planes in channels-last order -> interleave + 3x3 FIR + noise + bias + leaky-ReLU * scale, multiplied by the
 * next conv's style modulation next_mul (batch, channels) (may be NULL) and split into that conv's packed operand
 * out [2][out_rows][Cp] (Cp == channels, row (b*oh + oy)*ow + ox, fmt as in spgan_pack_act). */
int spgan_upblur_pack(uint16_t* out, const float* pp, const float* kernel, const float* noise, const float* noise_w,
                      const float* bias, const float* next_mul, int64_t batch, int channels, int zh, int zw, int Hq, int Wq,
                      int Cp, int64_t out_rows, int fmt, float alpha, float scale, void* stream);
/* out (batch, rgb_n, plane) = sum_s part[s] + bias[j] + skip (ToRGB's bias and upsampled skip, models/spgan_ops.py:1576-1585);
 * part (slots, batch, rgb_n, plane) from spgan_conv_gemm_ex; bias (rgb_n) and skip (batch, rgb_n, plane) may be NULL. */
int spgan_rgb_tail(float* out, const float* part, int slots, const float* bias, const float* skip, int64_t batch, int rgb_n,
                   int64_t plane, void* stream);
/* ---- mapping network and modulation chain --------------------------------------------------------------
 * spgan_mapping_chain: PixelNorm (models/ops.py:13-20) + n_layers x [EqualLinear 512 -> 512 + fused leaky-ReLU] (models/spgan/
 *   spgan.py:405-412, models/ops.py:190-222) in ONE kernel: a cluster of 8 CTAs per 8 latent rows, activations exchanged through
 *   distributed shared memory.  z (B rows, z_stride floats apart, 512 wide) -> out (B, 512).  weights / biases: HOST arrays of
 *   n_layers device pointers ((512, 512) and (512) or NULL).  y = lrelu(x (W w_scale)^T + b b_scale, alpha) * gain per layer.
 * spgan_modulation_batch: the modulation s = EqualLinear(style) and demodulation d = rsqrt(c_scale^2 sum_c s^2 sum_t W^2 + eps)
 *   (models/ops.py:598-604) of EVERY modulated conv of the generator in one launch.  `layers`: DEVICE array of n_layers records
 *   of spgan_modulation_layer_bytes() bytes: {const float* wm (Cin, 512); const float* bm (Cin) or NULL; const float* wsq
 *   (Cout, Cin) = sum over taps of W^2, or NULL for no demodulation; int64 s_off, d_off (float offsets into out of s (B, Cin)
 *   and d (B, Cout)); int32 Cin, Cout, style_sel, style_idx; float m_scale, m_lr_mul, c_scale, eps}.  style_sel 0 reads
 *   styles[b*styles_bstride + style_idx*512 + k], style_sel 1 reads gl[b*gl_bstride + k].  Cin <= 520. */
int spgan_mapping_chain(float* out, const float* z, int64_t z_stride, int B, const float* const* weights,
                        const float* const* biases, int n_layers, float w_scale, float b_scale, float alpha, float gain,
                        void* stream);
int spgan_modulation_layer_bytes(void);
int spgan_modulation_batch(float* out, const void* layers, int n_layers, const float* styles, int64_t styles_bstride,
                           const float* gl, int64_t gl_bstride, int B, void* stream);

/* ---- training-loop tails ------------------------------------------------------------------------------
 * spgan_ema_multi: the EMA `accumulate` of utils.py:86-94 over ALL parameters in one launch: table is a DEVICE array of
 *   nchunks records {float* dst; const float* src; int64_t n} (24 bytes each, n <= spgan_ema_chunk_elems()), one CTA per
 *   record: dst[i] = fma(alpha, src[i], dst[i] * decay) — torch's mul_(decay).add_(src, alpha=1-decay); the caller passes
 *   alpha = (float)(1 - decay) evaluated in double precision, as Python does for the reference. */
int spgan_ema_chunk_elems(void);
int spgan_ema_multi(const void* table, int nchunks, float decay, float alpha, void* stream);
/* spgan_minibatch_stddev: models/stylegan2discriminator.py:205-212 fused with its torch.cat: h (B, C, HW) -> out (B, C+1, HW),
 *   out[:, :C] = h, out[b, C, :] = mean over (c, p) of sqrt(var over the `group` samples {n*M + b % M} + eps), M = B / group
 *   (stddev_feat = 1).  partial: fp32 scratch of M * 8 elements.  Forward only; the host composes the gradient. */
int spgan_minibatch_stddev(float* out, float* partial, const float* h, int B, int C, int HW, int group, float eps, void* stream);

/* Tuning switches (process-wide).  key 1: use of the CTA-pair (tcgen05 cta_group::2, 256 x 256 tiles over two SMs) variant of
 * the GEMM kernel for Cout % 256 == 0: 0 = never, 1 = where its tiling fills the 148 SMs at least as well as the single-CTA
 * tiling (default), 2 = wherever legal.  Results do not depend on it (same products, same K order). */
int spgan_set_option(int key, int value);
/* Number of tcgen05 GEMM launches since load (the bench's gpu_launches evidence for the tensor path). */
int64_t spgan_gemm_launch_count(void);

/* ---- weight gradient on tcgen05 ---------------------------------------------------------------------------------
 * dW[t][o][c] = out_scale * sum_q G'[q][o] * X'[phase_t][q + off_t][c] over the flattened lattice q = (b*Hl + i)*Wl + j
 * (replaces cuDNN wgrad of models/ops.py:617, 634, 175 and models/spgan_ops_gs.py:814 under autograd).  Both operands are
 * spgan_pack_act outputs on one common (Hl, Wl) lattice: g_packed = pack of the output gradient with in_mul = the
 * demodulation and step = the pass's out_stride ([2][g_phases*Q][gp_cols]); x_packed = pack of the input with in_mul =
 * the style modulation and step = the pass's in_stride ([2][x_phases*Q][xp_cols]); Q = B*Hl*Wl.
 * `p` describes the pass on that lattice: (B, H, W) = (B, Hl, Wl), Cin, Cout, ntaps, tap_dy/tap_dx = non-negative lattice
 * offsets of each tap inside ITS phase plane of x_packed (tap_phase[t], NULL = 0), tap_w/ws_o/ws_c = where the tap lives
 * in dw, out_scale, precision 1 (bf16x3) or 2 (bf16).  This pass reads phase plane g_phase of g_packed.  workspace: fp32
 * scratch of at least spgan_conv_wgrad_gemm_workspace(p) elements (per-K-chunk partial tiles, summed in a fixed order:
 * results are deterministic).  accumulate != 0 adds to dw. */
int64_t spgan_conv_wgrad_gemm_workspace(const SpganConvPass* p);
int spgan_conv_wgrad_gemm(const SpganConvPass* p, float* dw, const uint16_t* g_packed, int g_phases, int g_phase,
                          int gp_cols, const uint16_t* x_packed, int x_phases, const int32_t* tap_phase, int xp_cols,
                          float* workspace, int64_t workspace_elems, int accumulate, void* stream);

/* Per-plane dot products: out[p] = sum_k a[p,k] * b[p,k]; planes x inner.  Used for the style / demodulation
 * gradients (d s[b,c] = <x[b,c], dxs[b,c]>). */
int spgan_plane_dot(float* out, const float* a, const float* b, int64_t planes, int64_t inner, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPGAN_B200_H */
