"""Data-parallel training step over the B200 hot path (host orchestration only).

The reference's loop (train.py:200-415) is out of scope as a subsystem; this module restates just enough of it to
*measure* the path under training load and to exercise first- and second-order gradients end to end:
D step (logistic + coord AC), lazy R1 (models/losses.py:36-41), G step (non-saturating + coord AC + diversity-z),
lazy path-length regularisation (models/losses.py:60-78), EMA.  One process per GPU; parameter gradients are averaged
with a bucketed NCCL all-reduce (the reference uses nn.DataParallel, train.py:809-816).  Inputs are synthetic
(SURVEY.md §8d): latents ~ N(0,1), "real" patches ~ clamp(N(0,1)), coordinate windows drawn like
coord_handler.py:907-921 from a 45 x 140 grid.
"""
import math

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import panorama
from .discriminator import Discriminator
from .generator import Generator, default_config


# ------------------------------------------------------------------------------------------------ losses (glue)
def d_logistic_loss(real_pred, fake_pred):
    return F.softplus(-real_pred).mean() + F.softplus(fake_pred).mean()


def g_nonsaturating_loss(fake_pred):
    return F.softplus(-fake_pred).mean()


def d_r1_loss(real_pred, real_img):
    grad_real, = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def path_lengths(fake_img, styles, noise=None):
    """calc_path_lengths (models/losses.py:60-68) w.r.t. the (B, n_latent, 512) styles."""
    if noise is None:
        noise = torch.randn_like(fake_img)
    noise = noise / math.sqrt(fake_img.shape[2] * fake_img.shape[3])
    grad, = torch.autograd.grad(outputs=(fake_img * noise).sum(), inputs=styles, create_graph=True)
    return torch.sqrt(grad.pow(2).mean([1, 2]))


def coord_ac_loss(pred, label):
    """coord_ac_vert_only (models/losses.py:85-86)."""
    return (pred[:, 0] - label[:, 0]).abs().mean()


def angular_similarity(a, b):
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    cos = (a * b).sum(1) / (a.norm(2, dim=1) * b.norm(2, dim=1))
    return 1 - torch.acos(cos) / np.pi


def diversity_z_loss(local_latent, structure_latent, eps=1e-5):
    """StructureSynthesizer.diversity_z_loss with diversity_angular (models/spgan/spgan.py:286-316)."""
    z = angular_similarity(local_latent[0::2], local_latent[1::2]).mean()
    x = angular_similarity(structure_latent[0::2], structure_latent[1::2]).mean()
    return 1 / (x / z + eps)


# ------------------------------------------------------------------------------------------------ synthetic inputs
class SyntheticSampler:
    """Latents, real patches and training coordinate windows (coord_handler.py:907-921, 1027-1038)."""

    GRID_X, GRID_Y, SIZE = 45, 140, 35

    def __init__(self, batch, device, seed=9000):
        self.batch, self.device = batch, device
        self.rng = np.random.RandomState(seed)
        self.gen = torch.Generator(device=device).manual_seed(seed)
        self.coord_canvas = panorama.meta_coords(self.GRID_X + self.SIZE, self.GRID_Y + self.SIZE, device)

    def latents(self, batch=None):
        b = batch or self.batch
        gl = torch.randn(b, 2, 512, device=self.device, generator=self.gen)
        lat = torch.randn(b, 256, self.SIZE, self.SIZE, device=self.device, generator=self.gen)
        return gl, lat

    def coords(self, batch=None):
        b = batch or self.batch
        x_st = self.rng.randint(0, 10, b)
        y_st = self.rng.randint(0, self.GRID_Y, b)
        S = self.SIZE
        coords = torch.stack([self.coord_canvas[:, x:x + S, y:y + S] for x, y in zip(x_st, y_st)]).contiguous()
        cps = [{"p_x_st": x / self.GRID_X, "p_x_ed": (x + S - 1) / self.GRID_X, "p_y_st": y / self.GRID_Y,
                "p_y_ed": (y + S - 1) / self.GRID_Y, "circular_flag": bool(y + S > self.GRID_Y), "x_total": self.GRID_X,
                "y_total": self.GRID_Y, "y_st": int(y), "y_ed": int(y + S), "partial": 0.6667} for x, y in zip(x_st, y_st)]
        ac = np.stack([(x_st / 9.0) * 2 - 1, np.cos(((y_st / (self.GRID_Y - 1)) * 2 - 1) * np.pi),
                       np.sin(((y_st / (self.GRID_Y - 1)) * 2 - 1) * np.pi)], 1)
        ac = torch.from_numpy(ac).float()
        if torch.device(self.device).type == "cuda":
            ac = ac.pin_memory()  # non-blocking upload: a pageable copy would stall the host behind the queued kernels
        return coords, cps, ac.to(self.device, non_blocking=True)

    def noises(self, batch=None):
        b = batch or self.batch
        return [torch.randn(b, 1, s, s, device=self.device, generator=self.gen) for s in (19, 17, 31, 29, 55, 53, 103, 101)]

    def real(self):
        img = torch.randn(self.batch, 3, 101, 101, device=self.device, generator=self.gen).clamp_(-1, 1)
        ac = torch.rand(self.batch, 3, device=self.device, generator=self.gen) * 2 - 1
        return img, ac


# ------------------------------------------------------------------------------------------------ gradient exchange
def allreduce_gradients(params, world, bucket_bytes=64 << 20):
    """Average parameter gradients over ranks: flatten into ~64 MB buckets (launch-latency sized, NVSwitch gives every
    rank full bandwidth), one NCCL all-reduce per bucket, unflatten.  No-op for world == 1."""
    if world <= 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    n = 0
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size, n
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n += 1
        bucket, size = [], 0
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n


def requires_grad(model, flag):
    for p in model.parameters():
        p.requires_grad_(flag)


@torch.no_grad()
def accumulate(ema, model, decay=0.999):
    """utils.py:86-94."""
    pe, pm = dict(ema.named_parameters()), dict(model.named_parameters())
    for k in pe:
        pe[k].mul_(decay).add_(pm[k], alpha=1 - decay)


# ------------------------------------------------------------------------------------------------ the step
class TrainStep:
    def __init__(self, batch, device, world=1, seed=9000, config=None, with_ema=True):
        self.config = config if config is not None else default_config()
        tp = self.config.train_params
        tp.batch_size = batch
        self.batch, self.device, self.world = batch, device, world
        torch.manual_seed(seed)
        self.G = Generator(self.config).to(device).train()
        self.D = Discriminator(self.config).to(device).train()
        self.G_ema = None
        if with_ema:
            self.G_ema = Generator(self.config).to(device).eval()
            self.G_ema.load_state_dict(self.G.state_dict())
        g_ratio = tp.g_reg_every / (tp.g_reg_every + 1)
        d_ratio = tp.d_reg_every / (tp.d_reg_every + 1)
        self.g_optim = torch.optim.Adam(self.G.parameters(), lr=tp.lr * g_ratio, betas=(0 ** g_ratio, 0.99 ** g_ratio))
        self.d_optim = torch.optim.Adam(self.D.parameters(), lr=tp.lr * d_ratio, betas=(0 ** d_ratio, 0.99 ** d_ratio))
        self.sampler = SyntheticSampler(batch, device, seed)
        self.mean_path_length = torch.zeros((), device=device)
        self.iter = 0

    def _fake(self, batch=None, need_latents=False):
        gl, lat = self.sampler.latents(batch)
        coords, cps, ac = self.sampler.coords(batch)
        noises = self.sampler.noises(batch)
        img, styles, structure = self.G(gl, lat, coords, cps, noises=noises, return_latents=True)
        return img, ac, styles, structure, lat

    def d_step(self, real=None):
        """`real` = (images (B,3,101,101), ac_coords (B,3)) from the caller's data loader, or None for synthetic."""
        requires_grad(self.G, False)
        requires_grad(self.D, True)
        with torch.no_grad():
            fake, fake_ac, _, _, _ = self._fake()
        real, real_ac = real if real is not None else self.sampler.real()
        fp, rp = self.D(fake), self.D(real)
        loss = d_logistic_loss(rp["d_patch"], fp["d_patch"])
        loss = loss + (coord_ac_loss(rp["ac_coords_pred"], real_ac) + coord_ac_loss(fp["ac_coords_pred"], fake_ac)) * \
            self.config.train_params.coord_ac_w
        self.D.zero_grad(set_to_none=True)
        loss.backward()
        allreduce_gradients(self.D.parameters(), self.world)
        self.d_optim.step()
        return loss.detach()

    def d_r1_step(self, real=None):
        tp = self.config.train_params
        requires_grad(self.D, True)
        real = real[0].detach().clone() if real is not None else self.sampler.real()[0]
        real.requires_grad_(True)
        rp = self.D(real)
        r1 = d_r1_loss(rp["d_patch"], real)
        self.D.zero_grad(set_to_none=True)
        (tp.r1 / 2 * r1 * tp.d_reg_every + 0 * rp["d_patch"][0]).sum().backward()
        allreduce_gradients(self.D.parameters(), self.world)
        self.d_optim.step()
        return r1.detach()

    def g_step(self):
        tp = self.config.train_params
        requires_grad(self.G, True)
        requires_grad(self.D, False)
        fake, fake_ac, _, structure, lat = self._fake()
        fp = self.D(fake)
        loss = g_nonsaturating_loss(fp["d_patch"]) + coord_ac_loss(fp["ac_coords_pred"], fake_ac) * tp.coord_ac_w
        if tp.diversity_z_w and self.batch % 2 == 0:
            loss = loss + diversity_z_loss(lat, structure) * tp.diversity_z_w
        self.G.zero_grad(set_to_none=True)
        loss.backward()
        allreduce_gradients(self.G.parameters(), self.world)
        self.g_optim.step()
        return loss.detach()

    def g_path_step(self):
        tp = self.config.train_params
        requires_grad(self.G, True)
        requires_grad(self.D, False)
        pb = max(1, self.batch // tp.path_batch_shrink)
        gl, lat = self.sampler.latents(pb)
        coords, cps, _ = self.sampler.coords(pb)
        noises = self.sampler.noises(pb)
        styles = self.G.texture_synthesizer.styles_for(gl, None)
        img = self.G(gl, lat, coords, cps, noises=noises, styles=styles)
        pl = path_lengths(img, styles)
        mean = self.mean_path_length + 0.01 * (pl.mean() - self.mean_path_length)
        penalty = (pl - mean).pow(2).mean()
        self.mean_path_length = mean.detach()
        self.G.zero_grad(set_to_none=True)
        (tp.path_regularize * tp.g_reg_every * penalty).backward()
        allreduce_gradients(self.G.parameters(), self.world)
        self.g_optim.step()
        return penalty.detach()

    def ema_step(self):
        if self.G_ema is not None:
            accumulate(self.G_ema, self.G)

    def step(self, lazy="schedule"):
        """One training iteration (train.py:200-415): D step, [R1 every d_reg_every], G step, [path-length every
        g_reg_every], EMA.  lazy="schedule" follows the reference's cadence (g_path_start ignored),
        "all" runs both regularisers, "none" skips them."""
        tp = self.config.train_params
        out = {"d": self.d_step()}
        if lazy == "all" or (lazy == "schedule" and self.iter % tp.d_reg_every == 0):
            out["r1"] = self.d_r1_step()
        out["g"] = self.g_step()
        if lazy == "all" or (lazy == "schedule" and self.iter % tp.g_reg_every == 0):
            out["path"] = self.g_path_step()
        self.ema_step()
        self.iter += 1
        return out
