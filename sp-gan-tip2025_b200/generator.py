"""Generator composition over the mirrored op modules, with the reference's parameter names.

The reference's `models/spgan/spgan.py` is a *caller* of the hot path: it only wires op modules.
bench.py and the parity tests cannot import it on the GPU box (the reference tree is not there, and it needs
easydict / CUDA-at-import), so this file wires the same graph for `configs/model/spgan.yaml`
(patch 101, ts_input 11, 4 structure blocks, coords on every layer) out of `spgan_b200.models.*`.
State-dict keys equal the reference's `InfinityGanGenerator` keys, so checkpoints load either way.
"""
import random
from types import SimpleNamespace

import torch
from torch import nn

from . import functional as SF
from .models import ops
from .models import spgan_ops as sp_ops
from .models import spgan_ops_gs as sp_ops_gs
from .models.spherenet import SphereConvBatchDiffFixBorderGNoGrad


def default_config():
    """The fields of configs/model/spgan.yaml that the hot path reads."""
    tp = SimpleNamespace(
        patch_size=101, full_size=197, training_modality="patch", batch_size=16, partial=0.6667,
        global_latent_dim=512, local_latent_dim=256, n_mlp=8, channel_multiplier=2, mixing=0.9,
        use_ss=True, ss_n_layers=4, ss_unfold_radius=3, ss_coord_all_layers="each_layer", ss_disable_noise=True,
        ts_input_size=11, ts_no_zero_pad=True, coord_num_dir=3, coord_vert_cut_pt=3, coord_vert_sample_size=10,
        coord_hori_occupy_ratio=0.25, r1=10, path_regularize=2, path_batch_shrink=2, d_reg_every=16, g_reg_every=4,
        lr=0.002, coord_use_ac=True, coord_ac_w=1, coord_ac_vert_only=True, diversity_z_w=1, diversity_angular=True)
    return SimpleNamespace(train_params=tp, var=SimpleNamespace(dataparallel=False))


def encode_coords(c):
    """convert_idx_to_input_coords_ori (coord_handler.py:696-711) for coord_num_dir == 3."""
    return SF.encode_coords(c)


def center_crop(src, h, w):
    ph, pw = (src.shape[2] - h) // 2, (src.shape[3] - w) // 2
    if ph == 0 and pw == 0:
        return src
    return src[:, :, ph:ph + h, pw:pw + w]


class ShortcutConv(nn.Conv2d):
    """`self.sc = nn.Conv2d(256, 256, 1)` of SphereConditionalBlock (models/spgan/spgan.py:141), run through the conv
    kernels of this package; same parameter names (`weight`, `bias`) and default init as nn.Conv2d."""

    _GEOM = SF.ConvGeom(1, 1)

    def forward(self, x, residual=None):
        if ops._grad_needed(x, self.weight, self.bias):
            y = SF.conv2d(x, self.weight, self._GEOM) + self.bias.view(1, -1, 1, 1)
            return y if residual is None else y + residual
        return SF.conv_apply(x, self.weight, self._GEOM, bias=self.bias, residual=residual)


class SphereConditionalBlock(nn.Module):
    """models/spgan/spgan.py:122-169: spherical StyledConv (LeakyReLU 0.01, no noise) + 1x1 shortcut."""

    def __init__(self, idx, config):
        super().__init__()
        tp = config.train_params
        self.config = config
        self.deal_coords = tp.ss_coord_all_layers == "each_layer"
        in_channel = tp.local_latent_dim
        if tp.ss_coord_all_layers or idx == 0:
            in_channel += tp.coord_num_dir
        self.sc = ShortcutConv(tp.local_latent_dim, tp.local_latent_dim, kernel_size=1)
        self.conv = sp_ops_gs.StyledConv(in_channel=in_channel, out_channel=tp.local_latent_dim, kernel_size=3,
                                         disable_noise=True, style_dim=tp.global_latent_dim, no_zero_pad=True,
                                         config=config, activation="LeakyReLU_n", side="ss",
                                         deal_coords=self.deal_coords)

    def forward(self, x, cond, coords, coords_partial, noise=None, test_ids=None, calc_flops=False):
        if not self.deal_coords and self.config.train_params.ss_coord_all_layers:
            x = torch.cat([x, coords], 1)
        out, flops = self.conv(x, cond, noise=noise, coords=coords, coords_partial=coords_partial, test_ids=test_ids,
                               calc_flops=calc_flops)
        return self.sc(x, residual=out), flops


class ConditionalBlock(nn.Module):
    """models/spgan/spgan.py:79-119: concat encoded coords, 7x7 StyledConv (no padding, fused leaky-ReLU)."""

    def __init__(self, idx, config):
        super().__init__()
        tp = config.train_params
        self.config = config
        in_channel = tp.local_latent_dim
        if tp.ss_coord_all_layers or idx == 0:
            in_channel += tp.coord_num_dir
        self.conv = ops.StyledConv(in_channel=in_channel, out_channel=tp.local_latent_dim,
                                   kernel_size=tp.ss_unfold_radius * 2 + 1, style_dim=tp.global_latent_dim,
                                   no_zero_pad=True, disable_noise=tp.ss_disable_noise, config=config, side="ss")

    def forward(self, x, cond, coords, coords_partial, noise=None, test_ids=None, calc_flops=False):
        if self.config.train_params.ss_coord_all_layers:
            x = torch.cat([x, encode_coords(coords)], 1)
        return self.conv(x, cond, noise=noise, coords=coords, test_ids=test_ids, calc_flops=calc_flops)


class ImplicitFunction(nn.Module):
    """models/spgan/spgan.py:172-254."""

    def __init__(self, config):
        super().__init__()
        convs = []
        for i in range(config.train_params.ss_n_layers):
            convs.append(SphereConditionalBlock(idx=i, config=config))
            convs.append(ConditionalBlock(idx=i, config=config))
        self.conv_stack = nn.Sequential(*convs)
        self.global_mapping = None  # spgan.yaml has no ss_mapping

    use_chain = True

    def _chain_ok(self, global_latent, local_latent, coords, coords_partial, noises, test_ids, calc_flops):
        """The channels-last structure chain covers spgan.yaml's inference configuration: test-mode sampling grids (one
        dict, or a grids.PositionGroup of several lattice positions), 256 features + 3 coordinate planes on every layer,
        no noise in the structure convs, bf16 operand planes (precision 1 or 2)."""
        from .grids import PositionGroup
        if not self.use_chain or calc_flops or test_ids is not None or coords is None or not local_latent.is_cuda:
            return False
        if SF.get_precision() not in (1, 2) or self.global_mapping is not None:
            return False
        if not (isinstance(coords_partial, dict) or isinstance(coords_partial, PositionGroup)):
            return False
        if ops._grad_needed(global_latent, local_latent, coords, *self.parameters()):
            return False
        if len(self.conv_stack) % 2 or local_latent.shape[1] != SF.SS_MAIN:
            return False
        for i, blk in enumerate(self.conv_stack):
            sc = blk.conv
            if i % 2 == 0:
                if not (isinstance(blk, SphereConditionalBlock) and blk.deal_coords and sc.noise is None and sc.conv.deal_coords
                        and isinstance(sc.activate, nn.LeakyReLU) and not sc.upsample and sc.conv.padding == 0
                        and sc.conv.demodulate and sc.conv.out_channel == SF.SS_MAIN and sc.conv.in_channel == SF.SS_MAIN + 3):
                    return False
            else:
                c = sc.conv
                if not (isinstance(blk, ConditionalBlock) and sc.noise is None and not c.upsample and c.padding == 0 and c.demodulate
                        and c.out_channel == SF.SS_MAIN and c.in_channel == SF.SS_MAIN + 3
                        and blk.config.train_params.ss_coord_all_layers):
                    return False
        return True

    def _forward_chain(self, global_latent, local_latent, coords, coords_partial):
        from .grids import GRID_CACHE, PositionGroup
        B, _, H, W = local_latent.shape
        prec = SF.get_precision()
        if isinstance(coords_partial, PositionGroup):
            cps, group = list(coords_partial), coords_partial.group
        else:
            cps, group = [coords_partial], B
        if len(cps) * group != B:
            raise RuntimeError("structure chain: %d positions x %d samples for a batch of %d" % (len(cps), group, B))
        xh, xp = SF.ss_input(local_latent, prec)
        memo = self.__dict__.setdefault("_next_mul_cache", {})
        out = None
        n = len(self.conv_stack)
        for i in range(0, n, 2):
            sph, cnd = self.conv_stack[i], self.conv_stack[i + 1]
            cc = center_crop(coords, H, W).contiguous()
            s_s, w_s, d_s = sph.conv.conv._mod_demod(global_latent, B)
            s_c, w_c, d_c = cnd.conv.conv._mod_demod(global_latent, B)
            mul7 = SF.memo_by_tensor(memo, i, s_c, lambda: s_c[:, :SF.SS_MAIN].contiguous())
            grid = GRID_CACHE.group_grid(H, W, cps, local_latent.device)
            y_sc = SF.ss_shortcut(xp, B, H, W, sph.sc.weight, sph.sc.bias, prec)
            a = SF.ss_sphere(xh, cc, grid, group, w_s, s_s, d_s, sph.conv.conv.scale, (sph.conv.activate.negative_slope, 1.0),
                             y_sc, mul7, prec)
            act = cnd.conv.activate
            last = i + 2 >= n
            xh, xp, (H, W) = SF.ss_conv_k(a, cc, B, H, W, w_c, s_c, d_c, cnd.conv.conv.scale, act.bias,
                                          (act.negative_slope, act.scale), prec, last)
            out = xh
        return out

    def forward(self, global_latent, local_latent, coords, coords_partial, noises=None, test_ids=None, calc_flops=False):
        if self._chain_ok(global_latent, local_latent, coords, coords_partial, noises, test_ids, calc_flops):
            return self._forward_chain(global_latent, local_latent, coords, coords_partial), 0
        h = local_latent
        flops = 0
        for conv in self.conv_stack:
            coords = center_crop(coords, h.shape[2], h.shape[3])  # coords_partial is NOT re-cut (:199-206)
            h, cur = conv(h, global_latent, coords, coords_partial, noise=noises, test_ids=test_ids, calc_flops=calc_flops)
            flops += cur
        return h, flops


class StructureSynthesizer(nn.Module):
    """models/spgan/spgan.py:257-389, for inputs whose coords / coords_partial are supplied by the caller (the test
    managers' `override_coords` path, or the synthetic training sampler of bench.py)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.implicit_model = ImplicitFunction(config)

    def calc_out_spatial_size(self, in_spatial_size, return_list=False):
        tp = self.config.train_params
        return in_spatial_size - tp.ss_n_layers * tp.ss_unfold_radius * 2

    def forward(self, global_latent, local_latent, coords, coords_partial, noises=None, test_ids=None, calc_flops=False):
        return self.implicit_model(global_latent, local_latent, coords=coords, coords_partial=coords_partial,
                                   noises=noises, test_ids=test_ids, calc_flops=calc_flops)


class TextureSynthesizer(nn.Module):
    """models/spgan/spgan.py:392-978 for patch_size 101 / ts_input_size 11."""

    CONVS = [(512, True), (512, False), (512, True), (512, False), (512, True), (512, False), (None, True), (None, False)]
    TO_RGBS = [(1, 3), (3, 5), (5, 7), (7, 8)]
    I2J = {3: 0, 5: 1, 7: 2}

    def __init__(self, config):
        super().__init__()
        tp = config.train_params
        if tp.patch_size != 101 or tp.ts_input_size != 11:
            raise NotImplementedError("only patch_size 101 / ts_input_size 11 (configs/model/spgan.yaml) is wired")
        self.config = config
        self.global_latent_dim = tp.global_latent_dim
        self.local_latent_dim = tp.local_latent_dim
        blur_kernel = [1, 2, 1]
        layers = [ops.PixelNorm()]
        for _ in range(tp.n_mlp):
            layers.append(ops.EqualLinear(self.global_latent_dim, self.global_latent_dim, lr_mul=0.01,
                                          activation='fused_lrelu'))
        self.mapping = nn.Sequential(*layers)
        self.const_z = ops.ConstantInput(self.local_latent_dim)
        self.convs = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.sp_convs = nn.ModuleList()
        self.noises = nn.Module()
        self.num_layers = len(self.CONVS)
        self.n_latent = self.num_layers + 1
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f'noise_{layer_idx}', torch.randn(1, 1, 2 ** res, 2 ** res))
        in_ch = self.local_latent_dim
        specs = [(c if c is not None else 256 * tp.channel_multiplier, up) for c, up in self.CONVS]
        for i, (out_ch, up) in enumerate(specs):
            self.convs.append(ops.StyledConv(in_ch, out_ch, 3, self.global_latent_dim, upsample=up,
                                             blur_kernel=blur_kernel, no_zero_pad=tp.ts_no_zero_pad, config=config,
                                             side="ts"))
            if i in self.I2J:
                self.sp_convs.append(SphereConvBatchDiffFixBorderGNoGrad(3, 3))
            in_ch = out_ch
        for src, _ in self.TO_RGBS:
            self.to_rgbs.append(sp_ops.ToRGB(specs[src][0], self.global_latent_dim, upsample=True,
                                             no_zero_pad=tp.ts_no_zero_pad, blur_kernel=blur_kernel, config=config,
                                             side="ts"))

    def calc_in_spatial_size(self, out_spatial_size, return_list=False):
        sizes = []
        for conv in self.convs[::-1]:
            out_spatial_size = conv.calc_in_spatial_size(out_spatial_size)
            sizes.append(out_spatial_size)
        return sizes[::-1] if return_list else sizes[-1]

    def calc_out_spatial_size(self, in_spatial_size, return_list=False):
        sizes = []
        for conv in self.convs:
            in_spatial_size = conv.calc_out_spatial_size(in_spatial_size)
            sizes.append(in_spatial_size)
        return sizes if return_list else sizes[-1]

    use_fused_mapping = True

    def _map(self, z):
        """The mapping network on (B, 512) latents.  no_grad on the GPU: PixelNorm + the 8 EqualLinear / leaky-ReLU layers
        as ONE cluster kernel (csrc/style_chain.cu); otherwise the module-by-module path (autograd)."""
        lin = [m for m in self.mapping if isinstance(m, ops.EqualLinear)]
        if (self.use_fused_mapping and z.is_cuda and z.dim() == 2 and z.shape[1] == 512 and z.stride(1) == 1 and len(lin) <= 16
                and isinstance(self.mapping[0], ops.PixelNorm) and len(lin) == len(self.mapping) - 1
                and all(m.activation and tuple(m.weight.shape) == (512, 512) and m.scale == lin[0].scale and m.lr_mul == lin[0].lr_mul
                        for m in lin)
                and not ops._grad_needed(z, *self.mapping.parameters())):
            return SF.mapping_chain(z, [m.weight for m in lin], [m.bias for m in lin], lin[0].scale, lin[0].lr_mul)
        return self.mapping(z)

    def get_style(self, global_latent):
        return self._map(global_latent)

    def styles_for(self, global_latent, inject_index=None, inject_mask=None):
        """global_latent (B, 2, 512) -> (B, n_latent, 512) w-space styles with style mixing at `inject_index`
        (models/spgan/spgan.py:843-876).  `inject_mask` (n_latent,) is the same choice as a device tensor (1 = first
        latent, 0 = second): lets a captured CUDA graph mix at a different index on every replay."""
        if inject_mask is not None:
            w0 = self._map(global_latent[:, 0])
            w1 = self._map(global_latent[:, 1])
            m = inject_mask.view(1, self.n_latent, 1)
            return w0.unsqueeze(1) * m + w1.unsqueeze(1) * (1 - m)
        w0 = self._map(global_latent[:, 0])
        if inject_index is None or inject_index >= self.n_latent:
            return w0.unsqueeze(1).repeat(1, self.n_latent, 1)
        w1 = self._map(global_latent[:, 1])
        return torch.cat([w0.unsqueeze(1).repeat(1, inject_index, 1),
                          w1.unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)

    # Per-layer precision of the fused inference chain.  None = the default policy: the global mode for every layer, except
    # that under the fp32-equivalent global mode (1, bf16x3) the LAST conv (48 % of the generator's FLOPs) runs the 2-MMA
    # fp16 split (mode 3) — its rounding error is not amplified by any later layer, and every golden of the reference
    # (B = 1, 2, 32 patches, 384x768 and 768x1536 panoramas) still passes at the tolerance the all-bf16x3 path is held to
    # (5e-4; measured 3.2e-4 .. 3.8e-4 against 2.3e-4 .. 2.6e-4).  A list of 8 modes overrides it; STRICT = bf16x3 everywhere.
    layer_precision = None
    STRICT = (1, 1, 1, 1, 1, 1, 1, 1)
    FAST_TAIL = (1, 1, 1, 1, 1, 3, 3, 3)
    use_chain = True
    # Power-of-two divisors of each layer's INPUT operand, used only where that layer runs mode 3 (fp16 planes saturate at
    # 65504, and nothing bounds the activations of a GAN): the producer folds 1 / act_scale[i] into the style modulation it
    # already applies, the consumer folds act_scale[i] into its output scale — both exact.  calibrate_act_scales() sets
    # them from one forward so that the largest operand value maps to ~2^8 (256x headroom).
    act_scale = None

    def calibrate_act_scales(self, styles, structure_latent, coords_partial, noises):
        """One bf16x3 pass of the chain that records max |operand| of every layer -> power-of-two act_scale."""
        import math
        rec = []
        with torch.no_grad():
            self._forward_chain(styles, structure_latent, coords_partial, noises, modes=[1] * self.num_layers, record=rec)
        amax = torch.stack(rec).tolist()  # one host sync: calibration is not on the hot path
        self.act_scale = [2.0 ** math.ceil(math.log2(max(a, 1e-30) / 256.0)) for a in amax]
        return self.act_scale

    def _chain_modes(self):
        g = SF.get_precision()
        if self.layer_precision is not None:
            modes = list(self.layer_precision)
        else:
            modes = [g] * self.num_layers
            if g == 1:
                modes[-1] = 3
        if len(modes) != self.num_layers or any(m not in (1, 2, 3) for m in modes):
            return None
        return modes

    def _chain_ok(self, styles, structure_latent, noises, test_ids, calc_flops):
        if not self.use_chain or calc_flops or test_ids is not None or noises is None or not structure_latent.is_cuda:
            return False
        if SF.get_precision() == 0 or self._chain_modes() is None:
            return False
        if ops._grad_needed(styles, structure_latent, *self.parameters()):
            return False
        for i, conv in enumerate(self.convs):
            c = conv.conv
            if conv.noise is None or not c.demodulate or c.out_channel % 32 or c.kernel_size != 3 or not c.no_zero_pad:
                return False
            if c.upsample and (tuple(c.blur.kernel.shape) != (3, 3) or c.blur.zero_pad != (0, 0) or c.blur.use_replicate_pad):
                return False
            if c.upsample != (i % 2 == 0):
                return False
        return all(n is not None for n in noises)

    def _scaled_mul(self, i, s, modes):
        """Style modulation of layer i with 1 / act_scale[i] folded in when the layer runs mode 3."""
        if modes[i] != 3 or self.act_scale is None or self.act_scale[i] == 1.0:
            return s, 1.0
        cache = self.__dict__.setdefault("_scaled_mul_cache", {})
        k = self.act_scale[i]
        return SF.memo_by_tensor(cache, i, s, lambda: s * (1.0 / k), extra=k), k

    def _forward_chain(self, styles, structure_latent, coords_partial, noises, modes=None, record=None):
        """The synthesis loop with channels-last operands between the convs (csrc/chain.cu): per (upsampling conv, conv)
        pair 4 parity GEMMs -> FIR tail writing the conv's packed operand -> GEMM writing the next pair's packed operand
        and the ToRGB partial sums -> rgb tail.  No fp32 activation of the texture synthesiser is written to HBM."""
        B = structure_latent.shape[0]
        if modes is None:
            modes = self._chain_modes()
            if 3 in modes and self.act_scale is None:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("texture chain: fp16 layers need calibrate_act_scales() before a CUDA-graph capture")
                self.calibrate_act_scales(styles, structure_latent, coords_partial, noises)
        sd = [conv.conv._mod_demod(styles[:, i], B) for i, conv in enumerate(self.convs)]
        H, W = structure_latent.shape[2], structure_latent.shape[3]
        mul0, k_in = self._scaled_mul(0, sd[0][0], modes)
        a = SF.chain_pack_input(structure_latent, mul0, modes[0])
        skip = None

        def note(t):
            if record is not None:
                record.append(t.view(torch.bfloat16)[0].abs().max().float())
        note(a)
        for k in range(self.num_layers // 2):
            up, cv = self.convs[2 * k], self.convs[2 * k + 1]
            (s_u, w_u, d_u), (s_c, w_c, d_c) = sd[2 * k], sd[2 * k + 1]
            pp, zhw = SF.chain_upconv(a, B, H, W, w_u, d_u, up.conv.scale * k_in, modes[2 * k])
            mul_c, k_in = self._scaled_mul(2 * k + 1, s_c, modes)
            a, (H, W) = SF.chain_upblur_pack(pp, zhw, up.conv.blur.kernel, noises[2 * k], up.noise.weight, up.activate.bias,
                                             mul_c, modes[2 * k + 1], up.activate.negative_slope, up.activate.scale)
            note(a)
            last = 2 * k + 2 >= self.num_layers
            rgb_mod = self.to_rgbs[k]
            s_r, w_r, _ = rgb_mod.conv._mod_demod(styles[:, self.TO_RGBS[k][1]], B)
            rgb_cache = rgb_mod.__dict__.setdefault("_rgbw_cache", {})
            rgb_w = SF.memo_by_tensor(
                rgb_cache, 0, s_r, lambda: (w_r.reshape(1, w_r.shape[0], w_r.shape[1]) * s_r.unsqueeze(1) * rgb_mod.conv.scale).contiguous(),
                extra=(w_r._version, SF.epoch()))
            mul_n, k_next = (None, 1.0) if last else self._scaled_mul(2 * k + 2, sd[2 * k + 2][0], modes)
            a, rgb, _, (H, W) = SF.chain_conv3(a, B, H, W, w_c, d_c, cv.conv.scale * k_in, noises[2 * k + 1], cv.noise.weight,
                                               cv.activate.bias, (cv.activate.negative_slope, cv.activate.scale),
                                               modes[2 * k + 1], next_mul=mul_n,
                                               next_precision=None if last else modes[2 * k + 2], rgb_w=rgb_w)
            k_in = k_next
            if not last:
                note(a)
            i = 2 * k + 1
            if i in self.I2J:
                skip = self.sp_convs[self.I2J[i]](skip, coords_partial)
            res = None
            if skip is not None:
                res = rgb_mod.upsample(skip)
                if rgb_mod.no_zero_pad:
                    res = rgb_mod.align_spatial_size(res, target=torch.empty(1, 1, H, W, device="meta"))
            skip = SF.rgb_tail(rgb[0], rgb[1], rgb_mod.bias.view(-1), res, B, H, W)
        return skip

    def forward(self, styles, structure_latent, coords_partial, noises=None, test_ids=None, calc_flops=False):
        """Synthesis loop (models/spgan/spgan.py:924-978).  styles (B, 9, 512); noises: list of 8 (B, 1, h, w) or None."""
        if self._chain_ok(styles, structure_latent, noises, test_ids, calc_flops):
            return self._forward_chain(styles, structure_latent, coords_partial, noises), 0
        h = structure_latent
        skip = None
        flops = 0
        rgb_idx = 0
        for i, conv in enumerate(self.convs):
            nz = noises[i] if noises is not None else None
            h, cur = conv(h, styles[:, i], noise=nz, test_ids=test_ids, calc_flops=calc_flops)
            flops += cur
            if rgb_idx < len(self.TO_RGBS) and i == self.TO_RGBS[rgb_idx][0]:
                if i in self.I2J:
                    skip = self.sp_convs[self.I2J[i]](skip, coords_partial)
                skip, cur = self.to_rgbs[rgb_idx](h, styles[:, self.TO_RGBS[rgb_idx][1]], skip=skip, calc_flops=calc_flops)
                flops += cur
                rgb_idx += 1
        return skip, flops


class Generator(nn.Module):
    """InfinityGanGenerator (models/spgan/spgan.py:1180-1420) restricted to the generation / training data flow:
    forward(global_latent (B,2,512), local_latent (B,256,35,35), coords (B,3,35,35) raw meta coords,
    coords_partial (dict for test mode, list of B dicts for training), noises, inject_index) -> image (B,3,101,101)."""

    def __init__(self, config=None):
        super().__init__()
        self.config = config if config is not None else default_config()
        self.structure_synthesizer = StructureSynthesizer(self.config)
        self.texture_synthesizer = TextureSynthesizer(self.config)

    use_fused_modulation = True

    def _modulated_layers(self):
        """(module, latent source, style index) of every modulated conv, in execution order: the structure synthesiser's convs
        read the RAW global latent (column 0), texture conv i reads styles[:, i], ToRGB k reads styles[:, TO_RGBS[k][1]]."""
        ts = self.texture_synthesizer
        out = [(blk.conv.conv, 1, 0) for blk in self.structure_synthesizer.implicit_model.conv_stack]
        out += [(conv.conv, 0, i) for i, conv in enumerate(ts.convs)]
        out += [(rgb.conv, 0, ts.TO_RGBS[k][1]) for k, rgb in enumerate(ts.to_rgbs)]
        return out

    @torch.no_grad()
    def prepare_modulation(self, global_latent, styles):
        """Compute the (modulation, demodulation) pair of EVERY modulated conv for these latents in ONE launch
        (spgan_modulation_batch) and seed each module's memo with it, so that the per-layer
        `_mod_demod` calls of the following generator forwards are hits.  global_latent (B, 2, 512), styles (B, n_latent, 512).
        Returns False (and does nothing) when the configuration is outside the fused kernel's reach."""
        if not self.use_fused_modulation or not styles.is_cuda or global_latent.dim() != 3 or styles.dim() != 3:
            return False
        if styles.stride(2) != 1 or styles.stride(1) != 512 or global_latent.stride(2) != 1:
            return False
        layers = self._modulated_layers()
        if any(m.in_channel > 520 or m.modulation is None or tuple(m.modulation.weight.shape) != (m.in_channel, 512)
               for m, _, _ in layers):
            return False
        if ops._grad_needed(global_latent, styles, *self.parameters()):
            return False
        B = styles.shape[0]
        views = [(global_latent[:, 0] if sel else styles[:, idx]) for _, sel, idx in layers]
        keys = [m._md_key(v) for (m, _, _), v in zip(layers, views)]
        if all(k in m.__dict__.get("_md_cache", {}) for (m, _, _), k in zip(layers, keys)):
            return True
        tkey = (B, SF.epoch()[0]) + tuple(k[6:] for k in keys)
        tables = self.__dict__.setdefault("_mod_tables", {})  # one per batch size (position-group sizes of a panorama engine)
        hit = tables.get(tkey)
        if hit is None:
            if torch.cuda.is_current_stream_capturing():
                return False  # the table upload cannot be captured: the per-layer path computes the pairs instead
            recs, off = [], 0
            for m, sel, idx in layers:
                r = dict(wm=m.modulation.weight, bm=m.modulation.bias, wsq=m.weight_sq() if m.demodulate else None, s_off=off,
                         Cin=m.in_channel, Cout=m.out_channel, style_sel=sel, style_idx=idx, m_scale=m.modulation.scale,
                         m_lr_mul=m.modulation.lr_mul, c_scale=m.scale, eps=1e-8)
                off = -(-(off + B * m.in_channel) // 64) * 64  # every slice starts 256-byte aligned: consumers use 128-bit loads
                r["d_off"] = off
                if m.demodulate:
                    off = -(-(off + B * m.out_channel) // 64) * 64
                recs.append(r)
            hit = (tkey, SF.modulation_table(recs, styles.device), recs, off)
            while len(tables) >= 8:
                del tables[next(iter(tables))]
            tables[tkey] = hit
        _, table, recs, total = hit
        buf = SF.modulation_batch(table, len(recs), total, styles, global_latent, B)
        for (m, _, _), v, k, r in zip(layers, views, keys, recs):
            s_ = buf[r["s_off"]:r["s_off"] + B * r["Cin"]].view(B, r["Cin"])
            d_ = buf[r["d_off"]:r["d_off"] + B * r["Cout"]].view(B, r["Cout"]) if m.demodulate else None
            m._md_store(k, v, s_, d_)
        return True

    def forward(self, global_latent, local_latent, coords, coords_partial, noises=None, inject_index=None,
                test_ids=None, return_latents=False, styles=None):
        """`styles` (B, 9, 512): optional precomputed w-space styles (`texture_synthesizer.styles_for`); the panorama
        loop maps the global latent once per panorama batch instead of once per patch."""
        if global_latent.dim() == 2:
            global_latent = torch.stack([global_latent, global_latent], 1)
        ts = self.texture_synthesizer
        if styles is None and inject_index is None and self.training and self.config.train_params.mixing > 0:
            if random.random() < self.config.train_params.mixing:  # models/spgan/spgan.py:865-869
                inject_index = random.randint(1, ts.n_latent - 1)
        if styles is None:
            styles = ts.styles_for(global_latent, inject_index)
        if not torch.is_grad_enabled() and not self.training:
            self.prepare_modulation(global_latent, styles)  # one launch; no-op when the memo already holds these latents
        structure, _ = self.structure_synthesizer(global_latent[:, 0], local_latent, coords, coords_partial,
                                                  test_ids=test_ids)
        img, _ = ts(styles, structure, coords_partial, noises=noises, test_ids=test_ids)
        if return_latents:
            return img, styles, structure
        return img
