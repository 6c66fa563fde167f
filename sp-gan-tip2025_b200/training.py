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


def discriminate_pair(D, fake, real):
    """D(fake), D(real) (train.py:262-263) as ONE pass over the interleaved batch [f0, r0, f1, r1, ...]: every conv of D is
    per-sample, and the minibatch-stddev feature groups samples {m, m + M, m + 2M, ...} (stylegan2discriminator.py:205-212:
    view(group, -1, ...)), so with M = 2 sub-batches the interleaving puts all fakes into one statistics group and all reals
    into the other — the same numbers as two separate calls, with half the launches and twice the rows per GEMM.  Falls back
    to two calls when the stddev grouping would not separate the two halves."""
    B = fake.shape[0]
    group = min(2 * B, getattr(D, "stddev_group", 0))
    if fake.shape != real.shape or group != B:
        return D(fake), D(real)
    both = torch.stack([fake, real], 1).reshape(2 * B, *fake.shape[1:])
    out = D(both)
    return {k: v[0::2] for k, v in out.items()}, {k: v[1::2] for k, v in out.items()}


def d_r1_loss(real_pred, real_img):
    from . import functional as SF
    with SF.only_data_grads():  # only d D(x) / dx is consumed here: no weight gradients in the create_graph pass
        grad_real, = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_img, create_graph=True)
    return grad_real.pow(2).reshape(grad_real.shape[0], -1).sum(1).mean()


def path_lengths(fake_img, styles, noise=None):
    """calc_path_lengths (models/losses.py:60-68) w.r.t. the (B, n_latent, 512) styles."""
    if noise is None:
        noise = torch.randn_like(fake_img)
    noise = noise / math.sqrt(fake_img.shape[2] * fake_img.shape[3])
    from . import functional as SF
    with SF.only_data_grads():
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
    """Latents, real patches and training coordinate windows (coord_handler.py:907-921, 1027-1038).

    Every sampled tensor lives in a STATIC device buffer that is refilled in place, and the coordinate windows are handed
    to the model as `grids.DeviceWindows` (the reference's list of dicts + device slot indices into pre-built factor
    tables).  Sampling therefore never blocks the host behind queued kernels, and the training step bodies can be
    captured in CUDA graphs: a replay reads whatever the buffers hold."""

    GRID_X, GRID_Y, SIZE = 45, 140, 35
    NOISE_SIZES = (19, 17, 31, 29, 55, 53, 103, 101)

    def __init__(self, batch, device, seed=9000):
        from . import grids
        self.batch, self.device = batch, torch.device(device)
        self.rng = np.random.RandomState(seed)
        self.gen = torch.Generator(device=device).manual_seed(seed)
        S = self.SIZE
        self.coord_canvas = panorama.meta_coords(self.GRID_X + S, self.GRID_Y + S, device)
        self._bufs = {}
        self.tables = None
        if self.device.type == "cuda":
            rows = [self._cp(x, 0) for x in range(10)]
            cols = [self._cp(0, y) for y in range(self.GRID_Y)]
            self.tables = grids.WindowTables(rows, cols, device)
            self.tables.prepare((35, 29, 23, 17, 53))
        self._ar = torch.arange(S, device=device)

    def _cp(self, x, y):
        S = self.SIZE
        return {"p_x_st": x / self.GRID_X, "p_x_ed": (x + S - 1) / self.GRID_X, "p_y_st": y / self.GRID_Y,
                "p_y_ed": (y + S - 1) / self.GRID_Y, "circular_flag": bool(y + S > self.GRID_Y), "x_total": self.GRID_X,
                "y_total": self.GRID_Y, "y_st": int(y), "y_ed": int(y + S), "partial": 0.6667}

    def _buf(self, name, shape, dtype=torch.float32):
        key = (name, tuple(shape))
        t = self._bufs.get(key)
        if t is None:
            t = self._bufs[key] = torch.zeros(shape, device=self.device, dtype=dtype)
        return t

    def _upload(self, dst, host):
        if self.device.type == "cuda":
            host = host.pin_memory()  # non-blocking: a pageable copy would stall the host behind the queued kernels
        dst.copy_(host, non_blocking=True)
        return dst

    def latents(self, batch=None, tag="a"):
        b = batch or self.batch
        gl = self._buf(tag + "_gl", (b, 2, 512)).normal_(generator=self.gen)
        lat = self._buf(tag + "_lat", (b, 256, self.SIZE, self.SIZE)).normal_(generator=self.gen)
        return gl, lat

    def coords(self, batch=None, tag="a"):
        b = batch or self.batch
        S = self.SIZE
        x_st = self.rng.randint(0, 10, b)
        y_st = self.rng.randint(0, self.GRID_Y, b)
        idx = self._upload(self._buf(tag + "_idx", (2, b), torch.int32), torch.from_numpy(np.stack([x_st, y_st]).astype(np.int32)))
        xs = idx[0].long()[:, None, None] + self._ar[None, :, None]
        ys = idx[1].long()[:, None, None] + self._ar[None, None, :]
        coords = self._buf(tag + "_coords", (b, 3, S, S))
        coords.copy_(self.coord_canvas[:, xs, ys].permute(1, 0, 2, 3))
        cps = [self._cp(int(x), int(y)) for x, y in zip(x_st, y_st)]
        if self.tables is not None:
            from . import grids
            cps = grids.DeviceWindows(cps, self.tables, idx[0], idx[1])
        ac = np.stack([(x_st / 9.0) * 2 - 1, np.cos(((y_st / (self.GRID_Y - 1)) * 2 - 1) * np.pi),
                       np.sin(((y_st / (self.GRID_Y - 1)) * 2 - 1) * np.pi)], 1)
        ac = self._upload(self._buf(tag + "_ac", (b, 3)), torch.from_numpy(ac).float())
        return coords, cps, ac

    def noises(self, batch=None, tag="a"):
        b = batch or self.batch
        return [self._buf(tag + "_nz%d" % i, (b, 1, s, s)).normal_(generator=self.gen) for i, s in enumerate(self.NOISE_SIZES)]

    def real(self, tag="a"):
        img = self._buf(tag + "_real", (self.batch, 3, 101, 101)).normal_(generator=self.gen).clamp_(-1, 1)
        ac = self._buf(tag + "_real_ac", (self.batch, 3)).uniform_(-1, 1, generator=self.gen)
        return img, ac

    def inject_mask(self, n_latent, mixing, tag="a"):
        """Style-mixing choice of spgan.py:865-869 as a device mask: 1 for the first latent, 0 for the second."""
        import random
        idx = n_latent
        if mixing > 0 and random.random() < mixing:
            idx = random.randint(1, n_latent - 1)
        m = torch.zeros(n_latent)
        m[:idx] = 1
        return self._upload(self._buf(tag + "_mask", (n_latent,)), m)


# ------------------------------------------------------------------------------------------------ gradient exchange
def _exchange_bucket(bucket, world):
    """One bucket: flatten, NCCL / gloo all-reduce (SUM), average, and ONE multi-tensor copy back into the gradients."""
    flat = torch.cat([g.reshape(-1) for g in bucket])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    views, off = [], 0
    for g in bucket:
        views.append(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    torch._foreach_copy_(bucket, views)


class OverlappedGradExchange:
    """Gradient all-reduce overlapped with the backward pass (what DistributedDataParallel's bucket hooks do).

    The parameters are assigned to ~`bucket_bytes` buckets in REVERSE registration order (backward reaches the last layers
    first).  A post-accumulate-grad hook counts the bucket's gradients; when the last one has arrived the bucket is exchanged
    on a SIDE stream (fork by event from the stream autograd runs on), while backward continues on the main stream; `finish()`
    exchanges whatever did not fill up (parameters without a gradient in this pass) and joins the side stream.  Inside a
    CUDA-graph capture the fork / join become graph edges, so the replayed step overlaps as well.  Before this the exchange
    ran after the whole backward: +3.5 ms on a 47 ms step at N = 2 (0.93 weak-scaling efficiency)."""

    def __init__(self, params, world, bucket_bytes=32 << 20):
        self.world = world
        self.params = [p for p in params]
        self.bucket_of, self.buckets = {}, []
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        for b, ps in enumerate(self.buckets):
            for p in ps:
                self.bucket_of[p] = b
        self.active = False
        self.side = None
        self.pending, self.done = [], []
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def begin(self):
        """Arm the hooks for the backward pass that follows (call after zero_grad, before backward)."""
        if self.world <= 1:
            return
        self.pending = [sum(1 for p in ps if p.requires_grad) for ps in self.buckets]
        self.done = [False] * len(self.buckets)
        if self.side is None:
            self.side = torch.cuda.Stream(device=self.params[0].device)
        self.active = True

    def _launch(self, b):
        grads = [p.grad for p in self.buckets[b] if p.grad is not None]
        self.done[b] = True
        if not grads:
            return
        main = torch.cuda.current_stream(grads[0].device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            _exchange_bucket(grads, self.world)

    def _hook(self, p):
        if not self.active:
            return
        b = self.bucket_of[p]
        self.pending[b] -= 1
        if self.pending[b] == 0 and not self.done[b]:
            self._launch(b)

    def finish(self):
        """Exchange the buckets that did not fill up and make the main stream wait for the side stream."""
        if self.world <= 1 or not self.active:
            return
        self.active = False
        for b in range(len(self.buckets)):
            if not self.done[b]:
                self._launch(b)
        torch.cuda.current_stream(self.params[0].device).wait_stream(self.side)


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
        _exchange_bucket(bucket, world)
        n += 1
        bucket, size = [], 0
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n


@torch.no_grad()
def sync_module_states(modules, world, src=0):
    """Make every rank start from rank `src`'s parameters and buffers (what DistributedDataParallel does at construction;
    nn.DataParallel, train.py:809-816, replicates one module, so the reference's replicas are equal by construction).
    Returns the number of broadcast tensors."""
    if world <= 1:
        return 0
    n = 0
    for m in modules:
        if m is None:
            continue
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src)
            n += 1
    return n


def replicas_in_sync(modules, world):
    """True when every rank holds the same parameters (max over ranks of a checksum == min over ranks)."""
    if world <= 1:
        return True
    acc = None
    for m in modules:
        for p in m.parameters():
            v = p.detach().double().sum() + p.detach().double().abs().sum()
            acc = v if acc is None else acc + v
    hi, lo = acc.clone(), acc.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return bool(hi == lo)


def requires_grad(model, flag):
    for p in model.parameters():
        p.requires_grad_(flag)


@torch.no_grad()
def accumulate(ema, model, decay=0.999):
    """utils.py:86-94."""
    pe, pm = dict(ema.named_parameters()), dict(model.named_parameters())
    dst = [pe[k] for k in pe]
    src = [pm[k].detach() for k in pe]
    if dst and dst[0].is_cuda:
        from . import functional as SF
        SF.ema_accumulate(dst, src, decay)  # ONE launch over all 115 parameter pairs (csrc/multi_tensor.cu)
        return
    torch._foreach_mul_(dst, decay)  # host-side tests (gloo, CPU): the torch multi-tensor ops
    torch._foreach_add_(dst, src, alpha=1 - decay)


# ------------------------------------------------------------------------------------------------ the step
class TrainStep:
    """One training iteration (train.py:200-415) over the B200 hot path.  Each part is split into `prepare_*` (sampling
    into static buffers, host-side) and a body that touches only device state; with `use_graphs=True` every body is
    captured in a CUDA graph after two eager runs and replayed afterwards — the B = 8 iteration is launch-bound
    (~6500 kernel launches for ~90 ms of kernels), a replay removes the Python / launch overhead."""

    def __init__(self, batch, device, world=1, seed=9000, config=None, with_ema=True, use_graphs=False, rank=0):
        """`seed` initialises the MODEL and is the same on every rank (the replicas are additionally broadcast from rank 0);
        the synthetic data stream of rank r is seeded with seed + 7919 * (r + 1), so the ranks see different samples."""
        self.config = config if config is not None else default_config()
        tp = self.config.train_params
        tp.batch_size = batch
        self.batch, self.device, self.world = batch, device, world
        self.use_graphs = use_graphs
        torch.manual_seed(seed)
        self.G = Generator(self.config).to(device).train()
        self.D = Discriminator(self.config).to(device).train()
        self.G_ema = None
        if with_ema:
            self.G_ema = Generator(self.config).to(device).eval()
            self.G_ema.load_state_dict(self.G.state_dict())
        g_ratio = tp.g_reg_every / (tp.g_reg_every + 1)
        d_ratio = tp.d_reg_every / (tp.d_reg_every + 1)
        cap = bool(use_graphs)
        self.g_optim = torch.optim.Adam(self.G.parameters(), lr=tp.lr * g_ratio, betas=(0 ** g_ratio, 0.99 ** g_ratio), capturable=cap)
        self.d_optim = torch.optim.Adam(self.D.parameters(), lr=tp.lr * d_ratio, betas=(0 ** d_ratio, 0.99 ** d_ratio), capturable=cap)
        sync_module_states([self.G, self.D, self.G_ema], world)
        # gradient exchange overlapped with backward (GPU, world > 1); the post-backward bucketed exchange otherwise
        self.overlap = world > 1 and torch.device(device).type == "cuda"
        self.g_xchg = OverlappedGradExchange(self.G.parameters(), world) if self.overlap else None
        self.d_xchg = OverlappedGradExchange(self.D.parameters(), world) if self.overlap else None
        self.rank = rank
        self.sampler = SyntheticSampler(batch, device, seed + 7919 * (rank + 1) if world > 1 else seed)
        self.mean_path_length = torch.zeros((), device=device)
        self.iter = 0
        self._graphs, self._eager_runs, self._inputs = {}, {}, {}

    # ---- sampling (host side, outside any graph) ----
    def _prepare_fake(self, tag, batch=None):
        ts = self.G.texture_synthesizer
        gl, lat = self.sampler.latents(batch, tag)
        coords, cps, ac = self.sampler.coords(batch, tag)
        noises = self.sampler.noises(batch, tag)
        mask = self.sampler.inject_mask(ts.n_latent, self.config.train_params.mixing, tag)
        return dict(gl=gl, lat=lat, coords=coords, cps=cps, ac=ac, noises=noises, mask=mask)

    def _real(self, tag, real):
        img, ac = self.sampler.real(tag)
        if real is not None:  # the caller's data loader: copied into the static buffers (H2D when it is host memory)
            img.copy_(real[0], non_blocking=True)
            ac.copy_(real[1], non_blocking=True)
        return img, ac

    def _fake(self, inp):
        styles = self.G.texture_synthesizer.styles_for(inp["gl"], inject_mask=inp["mask"])
        img, styles, structure = self.G(inp["gl"], inp["lat"], inp["coords"], inp["cps"], noises=inp["noises"],
                                        return_latents=True, styles=styles)
        return img, styles, structure

    def _backward(self, loss, model, xchg):
        """backward + gradient exchange: overlapped per bucket on a side stream (GPU, world > 1), else after the pass."""
        if xchg is None:
            loss.backward()
            allreduce_gradients(model.parameters(), self.world)
            return
        xchg.begin()
        try:
            loss.backward()
        finally:
            xchg.finish()

    # ---- graph plumbing ----
    def _run(self, name, body):
        if not self.use_graphs:
            return body()
        hit = self._graphs.get(name)
        if hit is None:
            n = self._eager_runs.get(name, 0)
            if n < 2:  # allocator, autotuned attributes and optimizer state settle in eager mode first
                self._eager_runs[name] = n + 1
                return body()
            from . import functional as SF
            SF.bump_epoch()  # every memoised pack / modulation must be recomputed INSIDE the capture
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = body()
            hit = self._graphs[name] = (graph, out)
        hit[0].replay()
        from . import functional as SF
        SF.bump_epoch()  # parameters changed on the device without a Python version bump
        return hit[1]

    # ---- the four parts ----
    def d_step(self, real=None):
        """`real` = (images (B,3,101,101), ac_coords (B,3)) from the caller's data loader, or None for synthetic."""
        requires_grad(self.G, False)
        requires_grad(self.D, True)
        inp = self._prepare_fake("d")
        real_img, real_ac = self._real("d", real)

        def body():
            with torch.no_grad():
                fake, _, _ = self._fake(inp)
            fp, rp = discriminate_pair(self.D, fake, real_img)
            loss = d_logistic_loss(rp["d_patch"], fp["d_patch"])
            loss = loss + (coord_ac_loss(rp["ac_coords_pred"], real_ac) + coord_ac_loss(fp["ac_coords_pred"], inp["ac"])) * \
                self.config.train_params.coord_ac_w
            self.D.zero_grad(set_to_none=True)
            self._backward(loss, self.D, self.d_xchg)
            self.d_optim.step()
            return loss.detach()
        return self._run("d", body)

    def d_r1_step(self, real=None):
        tp = self.config.train_params
        requires_grad(self.G, False)
        requires_grad(self.D, True)
        real_img, _ = self._real("r1", real)

        def body():
            x = real_img.detach().clone().requires_grad_(True)
            rp = self.D(x)
            r1 = d_r1_loss(rp["d_patch"], x)
            self.D.zero_grad(set_to_none=True)
            self._backward((tp.r1 / 2 * r1 * tp.d_reg_every + 0 * rp["d_patch"][0]).sum(), self.D, self.d_xchg)
            self.d_optim.step()
            return r1.detach()
        return self._run("r1", body)

    def g_step(self):
        tp = self.config.train_params
        requires_grad(self.G, True)
        requires_grad(self.D, False)
        inp = self._prepare_fake("g")

        def body():
            fake, _, structure = self._fake(inp)
            fp = self.D(fake)
            loss = g_nonsaturating_loss(fp["d_patch"]) + coord_ac_loss(fp["ac_coords_pred"], inp["ac"]) * tp.coord_ac_w
            if tp.diversity_z_w and self.batch % 2 == 0:
                loss = loss + diversity_z_loss(inp["lat"], structure) * tp.diversity_z_w
            self.G.zero_grad(set_to_none=True)
            self._backward(loss, self.G, self.g_xchg)
            self.g_optim.step()
            return loss.detach()
        return self._run("g", body)

    def g_path_step(self):
        tp = self.config.train_params
        requires_grad(self.G, True)
        requires_grad(self.D, False)
        pb = max(1, self.batch // tp.path_batch_shrink)
        inp = self._prepare_fake("p", pb)

        def body():
            styles = self.G.texture_synthesizer.styles_for(inp["gl"], inject_mask=inp["mask"])
            img = self.G(inp["gl"], inp["lat"], inp["coords"], inp["cps"], noises=inp["noises"], styles=styles)
            pl = path_lengths(img, styles)
            batch_mean = pl.mean()
            if self.world > 1:
                # the running mean is replica state: every rank folds in the GLOBAL batch mean (value), while the gradient
                # still flows through its own samples, as in models/losses.py:70-78
                glob = batch_mean.detach().clone()
                dist.all_reduce(glob, op=dist.ReduceOp.SUM)
                batch_mean = batch_mean + (glob / self.world - batch_mean.detach())
            mean = self.mean_path_length + 0.01 * (batch_mean - self.mean_path_length)
            penalty = (pl - mean).pow(2).mean()
            self.G.zero_grad(set_to_none=True)
            self._backward(tp.path_regularize * tp.g_reg_every * penalty, self.G, self.g_xchg)
            self.mean_path_length.copy_(mean.detach())  # in place: the running mean is state a graph replay must see
            self.g_optim.step()
            return penalty.detach()
        return self._run("path", body)

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
