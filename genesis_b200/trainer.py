"""Data-parallel training step of the engine: forward + loss assembly + backward + one gradient all-reduce +
fused optimiser, with GECO kept on the device.

Mirrors the caller's hot loop in the reference (train.py:215-263 and utils/geco.py:35-51) without its host
syncs.  One process per GPU; parameters and gradients live in flat fp32 arenas so the data-parallel exchange
is ONE all-reduce per step over NCCL (NVLink / NVSwitch), with the batch-mean `err` and `kl` appended to the
arena so every rank updates GECO's beta identically (replaces nn.DataParallel, train.py:153-155).

Checkpoints interchange with the reference: `TrainStep.checkpoint()` / `restore()` use train.py:410-416's layout
(`model_state_dict`, `optimiser_state_dict` in torch.optim's own format, `beta`, `iter_idx`, `err_ema`)."""
import torch
import torch.distributed as dist

from . import _lib, noise, ops


class GecoState(object):
    """utils/geco.py:19-51 as ONE device tensor {beta, err_ema, started} updated by one kernel (g2_geco_step_f32): no
    `.item()`, no host sync, replayable inside a CUDA graph."""

    def __init__(self, goal, step_size, device, alpha=0.99, beta_init=1.0, beta_min=1e-10, beta_max=1e10,
                 speedup=10.0):
        self.goal, self.step_size, self.alpha = float(goal), float(step_size), float(alpha)
        self.speedup = float(speedup) if speedup else 0.0
        self.beta_min, self.beta_max = float(beta_min), float(beta_max)
        self.vec = torch.tensor([float(beta_init), 0.0, 0.0], device=device)
        self.beta, self.err_ema, self.started = self.vec[0], self.vec[1], self.vec[2]     # 0-dim views

    def state(self):
        """What train.py:410-416 stores in a checkpoint ('beta', 'err_ema')."""
        return {'beta': self.beta.detach().clone(), 'err_ema': self.err_ema.detach().clone()}

    @torch.no_grad()
    def load_state(self, state):
        """Restore from a reference checkpoint dict (train.py:197-203)."""
        if 'beta' in state and state['beta'] is not None:
            self.beta.copy_(torch.as_tensor(state['beta'], dtype=torch.float32))
        if 'err_ema' in state and state['err_ema'] is not None:
            self.err_ema.copy_(torch.as_tensor(state['err_ema'], dtype=torch.float32))
            self.started.fill_(1.0)

    def step(self, err_kl, step_count, elbo, inv_world, update=True):
        _lib.call('g2_geco_step_f32', self.vec, err_kl, step_count, elbo, inv_world, self.goal, self.step_size, self.alpha,
                  self.speedup, self.beta_min, self.beta_max, 1 if update else 0)

    @torch.no_grad()
    def update(self, err):
        """Same update from a ready batch-mean `err` (0-dim device tensor); used by tests against utils/geco.py."""
        pair = torch.stack([err.detach().float().reshape(()), torch.zeros((), device=self.vec.device)])
        self.step(pair, None, None, 1.0)


class FlatArena(object):
    """Flat fp32 parameter and gradient arenas: [params | pad] and [grads | pad | err, kl, 0, 0].  Every parameter starts
    on a 256-byte boundary; `p.data` / `p.grad` become views, so autograd accumulates straight into the arena and the
    data-parallel exchange is one collective over `flat_g` (the 4 trailing floats carry the batch-mean err / kl so that
    every rank updates GECO identically)."""
    ALIGN = 64
    TAIL = 4

    def __init__(self, params):
        dev = params[0].device
        a = self.ALIGN
        self.n_params = sum(p.numel() for p in params)
        self.n_pad = sum((p.numel() + a - 1) // a * a for p in params)
        self.flat_p = torch.zeros(self.n_pad, device=dev)
        self.flat_g = torch.zeros(self.n_pad + self.TAIL, device=dev)
        self.offsets = []
        off = 0
        for p in params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            self.offsets.append(off)
            off += (k + a - 1) // a * a
        self.tail = self.flat_g[self.n_pad:]

    def exchange(self, err, kl, world, skip=()):
        """All-reduce(sum) of gradients + (err, kl) -- ONE collective over the whole arena, or, when some buckets were already
        reduced during the backward pass (`skip`: list of (start, end) element ranges), the remaining ranges.  Gradients and the
        tail are left as the SUM over ranks: the optimiser and the GECO kernel apply the 1/world scale."""
        self.tail[0].copy_(err)
        self.tail[1].copy_(kl)
        if world > 1:
            pos = 0
            for a, b in sorted(skip):
                if a > pos:
                    dist.all_reduce(self.flat_g[pos:a])
                pos = max(pos, b)
            dist.all_reduce(self.flat_g[pos:])


def shard_batch(x, rank, world):
    """Contiguous batch shard of rank `rank` (equal shards; the reference's DataParallel scatter, train.py:153-155)."""
    assert x.shape[0] % world == 0, 'global batch must divide by the number of ranks'
    per = x.shape[0] // world
    return x[rank * per:(rank + 1) * per]


OPTIMISERS = ('adam', 'rmsprop', 'sgd')          # train.py:171-176


class TrainStep(object):
    """step(x) = one optimisation step on the local shard `x` (float32 [B,3,H,W], device or pinned host).

    optimiser / lr, geco + g_* and beta / beta_warmup / train_iter take the values of the caller's flags (train.py:62-86).
    rank / noise_seed: every rank builds the model under the same torch.manual_seed so parameters match; the device noise
    generator is then re-seeded with noise_seed + rank, so the ranks draw INDEPENDENT eps / IC-SBP seeds for their shards (what
    nn.DataParallel's per-device generators give the reference; identical noise on every rank would correlate the global batch)."""

    def __init__(self, model, lr=1e-4, img_size=64, g_goal=0.5655, g_lr=1e-5, g_alpha=0.99, g_init=1.0,
                 g_min=1e-10, g_speedup=10.0, geco=True, world_size=1, rank=0, noise_seed=None, optimiser='adam',
                 beta=0.5, beta_warmup=False, train_iter=500000, overlap=True):
        if optimiser not in OPTIMISERS:
            raise ValueError('optimiser must be one of %s' % (OPTIMISERS,))
        if getattr(model, 'multi_gpu', False):
            raise ValueError('nn.DataParallel (--multi_gpu) is not supported: run one process per GPU (torchrun)')
        self.model = model
        self.world = world_size
        self.rank = rank
        self.lr = lr
        self.optimiser = optimiser
        params = [p for p in model.parameters() if p.requires_grad]
        self.arena = FlatArena(params)
        self.n_params, self.n_pad = self.arena.n_params, self.arena.n_pad
        self.flat_p, self.flat_g = self.arena.flat_p, self.arena.flat_g
        dev = self.flat_p.device
        self.flat_m = torch.zeros(self.n_pad, device=dev)          # Adam exp_avg | RMSprop square_avg | SGD momentum buffer
        self.flat_v = torch.zeros(self.n_pad, device=dev) if optimiser == 'adam' else None
        self.params = params
        self.step_count = torch.zeros((), device=dev)
        self.use_geco = bool(geco)
        self.fixed_beta, self.beta_warmup, self.train_iter = float(beta), bool(beta_warmup), int(train_iter)
        # reference train.py:159-169; without GECO the same object only carries err / kl / elbo bookkeeping
        self.geco = GecoState(g_goal * 3 * img_size ** 2, g_lr * (64 ** 2 / img_size ** 2), dev,
                              alpha=g_alpha, beta_init=g_init, beta_min=g_min, speedup=g_speedup)
        self.elbo = torch.zeros((), device=dev)
        self.x_dev = None
        self.graph = None
        self._tracking = False
        self.launches_per_step = None
        if noise_seed is not None:
            noise.seed_rank(noise_seed, rank, dev)
        # data-parallel overlap: gradient buckets are all-reduced on a communication stream as soon as the backward pass has
        # produced them (ops' gradient-ready tracking + autograd hooks), not after the whole backward
        self.overlap = world_size > 1 and overlap
        if self.overlap:
            self._setup_overlap()

    # ------------------------------------------------------------------------------------------ all-reduce overlap
    def _setup_overlap(self):
        total = self.n_pad
        cap = max(total // 4, 1 << 20)                     # about four collectives; a larger parameter is its own bucket
        self.buckets, self.bucket_of = [], {}
        start, size, members = 0, 0, []
        for p, off in zip(self.params, self.arena.offsets):
            k = (p.numel() + FlatArena.ALIGN - 1) // FlatArena.ALIGN * FlatArena.ALIGN
            if members and size + k > cap:
                self.buckets.append((start, off, members))
                start, size, members = off, 0, []
            members.append(p)
            size += k
        self.buckets.append((start, total, members))
        for bi, (_, _, members) in enumerate(self.buckets):
            for p in members:
                self.bucket_of[id(p)] = bi
        self.comm = torch.cuda.Stream(device=self.flat_p.device)
        for p in self.params:        # parameters whose gradient autograd accumulates itself (not written by a direct kernel)
            p.register_post_accumulate_grad_hook(self._on_grad_ready)

    def _begin_overlap(self):
        ops.reset_grad_tracking()
        ops.set_grad_ready_callback(self._on_grad_ready)
        self._pending = [len(m) for (_, _, m) in self.buckets]
        self._main = torch.cuda.current_stream()
        self._reported = set()
        self._fired = []

    def _on_grad_ready(self, p):
        if not self._tracking:
            return
        if id(p) in self._reported:
            raise RuntimeError('a parameter reported its gradient complete twice in one step (it is written both by a direct '
                               'kernel and by autograd): overlap of the gradient all-reduce is unsafe for this model; construct '
                               'TrainStep(overlap=False)')
        self._reported.add(id(p))
        bi = self.bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            a, b, _ = self.buckets[bi]
            self.comm.wait_stream(self._main)
            self.comm.wait_stream(torch.cuda.current_stream())      # a side-branch backward reports from its own stream
            for s_ in ops.all_side_streams(self.flat_p.device):
                self.comm.wait_stream(s_)
            with torch.cuda.stream(self.comm):
                dist.all_reduce(self.flat_g[a:b])
            self._fired.append((a, b))

    def _end_overlap(self):
        ops.set_grad_ready_callback(None)
        torch.cuda.current_stream().wait_stream(self.comm)
        return list(self._fired)

    # ------------------------------------------------------------------------------------------ loss
    def loss_terms(self, losses):
        """train.py:227-239 (if / elif: the stacked per-step lists win over a plain tensor of the same stage)."""
        err = losses['err'].mean(0)
        kl = err.new_zeros(())
        for plain, listed in (('kl_m', 'kl_m_k'), ('kl_l', 'kl_l_k')):
            if plain in losses and torch.is_tensor(losses[plain]):
                kl = kl + losses[plain].mean(0)
            elif listed in losses and len(losses[listed]):
                kl = kl + torch.stack(list(losses[listed]), dim=1).mean(0).sum()
        return err, kl

    def current_beta(self):
        """beta of the objective for THIS step (train.py:249-259)."""
        if self.use_geco:
            return self.geco.beta
        if self.beta_warmup:        # linear over the first 20 % of training; step_count == iter_idx before the update
            return (self.step_count * (self.fixed_beta / (0.2 * self.train_iter))).clamp(0.0, self.fixed_beta)
        return self.fixed_beta

    # ------------------------------------------------------------------------------------------ graph
    def _mutable_state(self):
        ts = [self.flat_p, self.flat_m, self.step_count, self.geco.vec, self.elbo]
        if self.flat_v is not None:
            ts.append(self.flat_v)
        ts += [b for b in self.model.buffers()]
        return ts

    def capture(self, x_example, warmup=3):
        """Capture forward + backward + all-reduce + GECO + optimiser of one step into a CUDA graph (static input buffer,
        static ELBO output).  The warm-up steps capture needs run on a snapshot: parameters, optimiser state, GECO, the
        BatchNorm buffers and the step counter are restored afterwards, so capture() does not train.  After this,
        step()/step_device() replay the graph: zero host work per step."""
        dev = self.flat_p.device
        self.x_static = torch.empty_like(x_example, device=dev)
        self.x_static.copy_(x_example)
        snapshot = [t.detach().clone() for t in self._mutable_state()]
        rng = torch.cuda.get_rng_state(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_eager(self.x_static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        lib = _lib.lib()
        l0 = lib.launches
        graph = torch.cuda.CUDAGraph()
        # Experiment (G2_MAIN_PRIORITY=1, off by default): capture the main chain on a HIGH-priority stream so that it wins the SMs
        # over the side streams whenever both have persistent one-CTA-per-SM kernels pending.  Measured: 8.26 -> 8.10 ... 8.25 ms at
        # N = 1, 8.39 -> 8.42 ms at N = 2 (the collectives then lose to the main chain): no consistent gain.
        import os
        cap_stream = torch.cuda.Stream(device=dev, priority=-1) if os.environ.get('G2_MAIN_PRIORITY', '0') == '1' else None
        with (torch.cuda.graph(graph, stream=cap_stream) if cap_stream is not None else torch.cuda.graph(graph)):
            self.elbo_static = self._step_eager(self.x_static)
        self.launches_per_step = lib.launches - l0
        with torch.no_grad():
            for t, s in zip(self._mutable_state(), snapshot):
                t.copy_(s)
            self.flat_g.zero_()
        torch.cuda.set_rng_state(rng, dev)
        torch.cuda.synchronize()
        self.graph = graph
        return self

    def step_device(self, x):
        """x already resident on the device.  Returns the (detached) ELBO scalar tensor of this step (a fresh tensor: the
        graph's static output is overwritten by the next replay)."""
        if self.graph is not None:
            if x is not self.x_static:
                self.x_static.copy_(x, non_blocking=True)
            self.graph.replay()
            return self.elbo_static.clone()
        return self._step_eager(x).clone()

    def _step_eager(self, x):
        # the arena gradients are pre-zeroed and re-zeroed by the fused optimiser kernel, so the conv / linear kernels may
        # accumulate weight and bias gradients straight into p.grad (no permute copy + AccumulateGrad add per parameter)
        ops.set_direct_grad(True)
        self._tracking = self.overlap
        if self.overlap:
            self._begin_overlap()
        reduced = ()
        try:
            recon, losses, stats, att, comp = self.model(x)
            err, kl = self.loss_terms(losses)
            loss = err + self.current_beta() * kl
            loss.backward()
            ops.join_grad_stream(self.flat_p.device)      # parameter-gradient kernels run on a side stream
        finally:
            ops.set_direct_grad(False)
            self._tracking = False
            if self.overlap:
                reduced = self._end_overlap()
        with torch.no_grad():
            # the buckets already reduced under the backward pass are skipped; the rest + (err, kl) go now
            self.arena.exchange(err.detach(), kl.detach(), self.world, skip=reduced)
            inv = 1.0 / self.world
            # GECO (or plain bookkeeping) + step counter + elbo in one scalar kernel; then the fused optimiser, which
            # scales the summed gradients by 1/world and re-zeroes them
            self.geco.step(self.arena.tail, self.step_count, self.elbo, inv, update=self.use_geco)
            if self.optimiser == 'adam':
                _lib.call('g2_adam_f32', self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.n_pad, self.lr,
                          0.9, 0.999, 1e-8, self.step_count, inv, 1)
            elif self.optimiser == 'rmsprop':
                _lib.call('g2_rmsprop_f32', self.flat_p, self.flat_g, self.flat_m, self.n_pad, self.lr, 0.99, 1e-8, inv, 1)
            else:
                _lib.call('g2_sgd_f32', self.flat_p, self.flat_g, self.flat_m, self.n_pad, self.lr, 0.9, self.step_count, inv, 1)
            self.arena.tail.zero_()
        return self.elbo

    def step(self, x):
        """Public entry point: x on the host (ideally pinned) or the device; returns the ELBO tensor."""
        if not x.is_cuda:
            if self.graph is not None:
                self.x_static.copy_(x, non_blocking=True)
                x = self.x_static
            else:
                dev = self.flat_p.device
                if self.x_dev is None or self.x_dev.shape != x.shape:
                    self.x_dev = torch.empty(x.shape, device=dev, dtype=torch.float32)
                self.x_dev.copy_(x, non_blocking=True)
                x = self.x_dev
        return self.step_device(x)

    # ------------------------------------------------------------------------------------------ checkpoints
    def _torch_optimiser(self):
        """A torch.optim object of the caller's kind over the same parameters (train.py:171-176): the source of the
        `param_groups` layout, so that the exported state loads into the reference's optimiser and vice versa."""
        if self.optimiser == 'adam':
            return torch.optim.Adam(self.params, self.lr)
        if self.optimiser == 'rmsprop':
            return torch.optim.RMSprop(self.params, self.lr)
        return torch.optim.SGD(self.params, self.lr, 0.9)

    def optimiser_state_dict(self):
        """The fused optimiser's state in torch.optim's own `state_dict()` format (per-parameter `step`, `exp_avg`,
        `exp_avg_sq` for Adam; `step`, `square_avg` for RMSprop; `momentum_buffer` for SGD)."""
        sd = self._torch_optimiser().state_dict()
        state = {}
        stepped = float(self.step_count.item()) > 0
        if stepped:
            for i, (p, off) in enumerate(zip(self.params, self.arena.offsets)):
                k = p.numel()
                m = self.flat_m[off:off + k].view_as(p).detach().clone()
                if self.optimiser == 'adam':
                    state[i] = {'step': self.step_count.detach().clone().cpu(), 'exp_avg': m,
                                'exp_avg_sq': self.flat_v[off:off + k].view_as(p).detach().clone()}
                elif self.optimiser == 'rmsprop':
                    state[i] = {'step': self.step_count.detach().clone().cpu(), 'square_avg': m}
                else:
                    state[i] = {'momentum_buffer': m}
        sd['state'] = state
        return sd

    @torch.no_grad()
    def load_optimiser_state_dict(self, sd):
        """Load a torch.optim state_dict (the `optimiser_state_dict` of a reference checkpoint, train.py:194) into the flat
        moment arenas and the device step counter."""
        state = sd.get('state', {})
        self.flat_m.zero_()
        if self.flat_v is not None:
            self.flat_v.zero_()
        step = 0.0
        key_m = {'adam': 'exp_avg', 'rmsprop': 'square_avg', 'sgd': 'momentum_buffer'}[self.optimiser]
        for i, (p, off) in enumerate(zip(self.params, self.arena.offsets)):
            st = state.get(i, state.get(str(i)))
            if not st:
                continue
            k = p.numel()
            if st.get(key_m) is not None:
                self.flat_m[off:off + k].copy_(st[key_m].reshape(-1))
            if self.optimiser == 'adam':
                self.flat_v[off:off + k].copy_(st['exp_avg_sq'].reshape(-1))
            step = max(step, float(st['step']) if 'step' in st else 1.0)
        self.step_count.fill_(step)

    def checkpoint(self, iter_idx=None):
        """The dict train.py:410-416 saves; loads in the reference's resume path (train.py:180-207) and in restore()."""
        ck = {'model_state_dict': self.model.state_dict(), 'optimiser_state_dict': self.optimiser_state_dict(),
              'beta': self.geco.beta.detach().clone() if self.use_geco else torch.tensor(self.fixed_beta),
              'iter_idx': int(self.step_count.item()) - 1 if iter_idx is None else iter_idx}
        if self.use_geco:
            ck['err_ema'] = self.geco.err_ema.detach().clone()
        return ck

    @torch.no_grad()
    def restore(self, ck):
        sd = dict(ck['model_state_dict'])
        for legacy in ('comp_vae.decoder_module.seq.0.pixel_coords.g_1', 'comp_vae.decoder_module.seq.0.pixel_coords.g_2'):
            sd.pop(legacy, None)                       # train.py:191-192
        self.model.load_state_dict(sd)                 # copies INTO the arena views
        self.load_optimiser_state_dict(ck['optimiser_state_dict'])
        nxt = ck.get('iter_idx', -1) + 1
        if self.optimiser == 'sgd' and nxt > 0:
            self.step_count.fill_(float(nxt))          # torch's SGD state carries no step: the checkpoint's iteration does
        if self.use_geco:
            self.geco.load_state(ck)
        return nxt
