"""Data-parallel training step of the engine: forward + loss assembly + backward + one gradient all-reduce +
fused Adam, with GECO kept on the device.

Mirrors the caller's hot loop in the reference (train.py:215-263 and utils/geco.py:35-51) without its host
syncs.  One process per GPU; parameters and gradients live in flat fp32 arenas so the data-parallel exchange
is ONE all-reduce per step over NCCL (NVLink / NVSwitch), with the batch-mean `err` and `kl` appended to the
arena so every rank updates GECO's beta identically (replaces nn.DataParallel, train.py:153-155)."""
import torch
import torch.distributed as dist

from . import _lib, ops


class GecoState(object):
    """utils/geco.py:19-51 as device tensors (no .item()): loss = err + beta*kl; err_ema; beta update."""

    def __init__(self, goal, step_size, device, alpha=0.99, beta_init=1.0, beta_min=1e-10, beta_max=1e10,
                 speedup=10.0):
        self.goal, self.step_size, self.alpha, self.speedup = float(goal), float(step_size), alpha, speedup
        self.beta = torch.tensor(beta_init, device=device)
        self.err_ema = torch.zeros((), device=device)
        self.started = torch.zeros((), device=device)
        self.beta_min, self.beta_max = beta_min, beta_max

    def state(self):
        """What train.py:410-416 stores in a checkpoint ('beta', 'err_ema')."""
        return {'beta': self.beta.detach().clone(), 'err_ema': self.err_ema.detach().clone()}

    @torch.no_grad()
    def load_state(self, state):
        """Restore from a reference checkpoint dict (train.py:197-203)."""
        if 'beta' in state:
            self.beta.copy_(torch.as_tensor(state['beta'], dtype=torch.float32))
        if 'err_ema' in state:
            self.err_ema.copy_(torch.as_tensor(state['err_ema'], dtype=torch.float32))
            self.started.fill_(1.0)

    @torch.no_grad()
    def update(self, err):
        ema = torch.where(self.started > 0, (1.0 - self.alpha) * err + self.alpha * self.err_ema, err)
        self.err_ema.copy_(ema)
        self.started.fill_(1.0)
        constraint = self.goal - self.err_ema
        rate = torch.where(constraint > 0, self.speedup * self.step_size, self.step_size) if self.speedup else self.step_size
        self.beta.copy_((torch.exp(rate * constraint) * self.beta).clamp(self.beta_min, self.beta_max))


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
        off = 0
        for p in params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            off += (k + a - 1) // a * a
        self.tail = self.flat_g[self.n_pad:]

    def exchange(self, err, kl, world):
        """ONE all-reduce(sum) of gradients + (err, kl); returns the global-batch means of err and kl.  Gradients are left
        as the SUM over ranks: the optimiser applies the 1/world scale."""
        self.tail[0].copy_(err)
        self.tail[1].copy_(kl)
        if world > 1:
            dist.all_reduce(self.flat_g)
        return self.tail[0] / world, self.tail[1] / world


def shard_batch(x, rank, world):
    """Contiguous batch shard of rank `rank` (equal shards; the reference's DataParallel scatter, train.py:153-155)."""
    assert x.shape[0] % world == 0, 'global batch must divide by the number of ranks'
    per = x.shape[0] // world
    return x[rank * per:(rank + 1) * per]


class TrainStep(object):
    """step(x) = one optimisation step on the local shard `x` (float32 [B,3,H,W], device or pinned host)."""

    def __init__(self, model, lr=1e-4, img_size=64, g_goal=0.5655, g_lr=1e-5, g_alpha=0.99, g_init=1.0,
                 g_min=1e-10, g_speedup=10.0, geco=True, world_size=1):
        self.model = model
        self.world = world_size
        self.lr = lr
        params = [p for p in model.parameters() if p.requires_grad]
        self.arena = FlatArena(params)
        self.n_params, self.n_pad = self.arena.n_params, self.arena.n_pad
        self.flat_p, self.flat_g = self.arena.flat_p, self.arena.flat_g
        dev = self.flat_p.device
        self.flat_m = torch.zeros(self.n_pad, device=dev)
        self.flat_v = torch.zeros(self.n_pad, device=dev)
        self.params = params
        self.step_count = torch.zeros((), device=dev)
        self.geco = None
        if geco:
            # reference train.py:159-169
            self.geco = GecoState(g_goal * 3 * img_size ** 2, g_lr * (64 ** 2 / img_size ** 2), dev,
                                  alpha=g_alpha, beta_init=g_init, beta_min=g_min, speedup=g_speedup)
        self.x_dev = None
        self.graph = None
        self.launches_per_step = None

    def loss_terms(self, losses):
        """train.py:227-239."""
        err = losses['err'].mean(0)
        kl = err.new_zeros(())
        for key in ('kl_l_k', 'kl_m_k'):
            if key in losses and len(losses[key]):
                kl = kl + torch.stack(list(losses[key]), dim=1).mean(0).sum()
        for key in ('kl_l', 'kl_m'):
            if key in losses and torch.is_tensor(losses[key]):
                kl = kl + losses[key].mean(0)
        return err, kl

    def capture(self, x_example, warmup=3):
        """Capture forward + backward + all-reduce + GECO + Adam of one step into a CUDA graph (static input
        buffer, static ELBO output).  After this, step()/step_device() replay the graph: zero host work per step."""
        dev = self.flat_p.device
        self.x_static = torch.empty_like(x_example, device=dev)
        self.x_static.copy_(x_example)
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
        with torch.cuda.graph(graph):
            self.elbo_static = self._step_eager(self.x_static)
        self.launches_per_step = lib.launches - l0
        self.graph = graph
        return self

    def step_device(self, x):
        """x already resident on the device.  Returns the (detached) ELBO scalar tensor."""
        if self.graph is not None:
            if x is not self.x_static:
                self.x_static.copy_(x, non_blocking=True)
            self.graph.replay()
            return self.elbo_static
        return self._step_eager(x)

    def _step_eager(self, x):
        # the arena gradients are pre-zeroed and re-zeroed by the fused Adam kernel, so the conv / linear kernels may
        # accumulate weight and bias gradients straight into p.grad (no permute copy + AccumulateGrad add per parameter)
        ops.set_direct_grad(True)
        try:
            recon, losses, stats, att, comp = self.model(x)
            err, kl = self.loss_terms(losses)
            beta = self.geco.beta if self.geco is not None else 1.0
            loss = err + beta * kl
            loss.backward()
            ops.join_grad_stream(self.flat_p.device)      # parameter-gradient kernels run on a side stream
        finally:
            ops.set_direct_grad(False)
        tail = self.arena.tail
        with torch.no_grad():
            gerr, gkl = self.arena.exchange(err.detach(), kl.detach(), self.world)      # ONE exchange per step
            if self.geco is not None:
                self.geco.update(gerr)
            self.step_count += 1
            elbo = (gerr + gkl).clone()
            _lib.call('g2_adam_f32', self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.n_pad, self.lr,
                      0.9, 0.999, 1e-8, self.step_count, 1.0 / self.world, 1)
            tail.zero_()
        return elbo

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
