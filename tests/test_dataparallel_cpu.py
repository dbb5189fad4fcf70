"""CPU, world_size 2 over gloo: the data-parallel host logic of genesis_b200/trainer.py (flat arenas, contiguous batch
shards, ONE all-reduce carrying gradients + err + kl, 1/world scaling) reproduces the single-process full-batch gradient.
The model here is a small per-sample-normalised torch network (no cross-sample statistics, like GENESIS-V2 / MONet), so
the result must be batch-partition invariant up to fp32 summation order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.GroupNorm(2, 8), nn.ReLU(), nn.Conv2d(8, 5, 1))


def _losses(model, x):
    y = model(x)
    err = (y ** 2).sum((1, 2, 3))                 # per-sample "err" [B]
    kl = [y[:, k].abs().sum((1, 2)) for k in range(3)]   # K per-sample KL terms
    return {'err': err, 'kl_l_k': kl}


def _loss_terms(losses):
    err = losses['err'].mean(0)
    kl = torch.stack(losses['kl_l_k'], 1).mean(0).sum()
    return err, kl


def _worker(rank, world, port, x, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from genesis_b200.trainer import FlatArena, shard_batch
    model = _model()
    arena = FlatArena(list(model.parameters()))
    xs = shard_batch(x, rank, world)
    err, kl = _loss_terms(_losses(model, xs))
    (err + 0.7 * kl).backward()
    arena.exchange(err.detach(), kl.detach(), world)
    gerr, gkl = arena.tail[0] / world, arena.tail[1] / world      # the tail carries the SUM over ranks of the batch means
    out[rank] = (arena.flat_g[:arena.n_pad].clone() / world, gerr.clone(), gkl.clone(),
                 [p.grad.data_ptr() for p in model.parameters()], arena.flat_g.data_ptr())
    dist.destroy_process_group()


def test_two_rank_gradient_equals_full_batch():
    torch.manual_seed(1)
    x = torch.rand(8, 3, 12, 12)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), x, out), nprocs=2, join=True)
    # single process, full batch
    from genesis_b200.trainer import FlatArena
    model = _model()
    arena = FlatArena(list(model.parameters()))
    err, kl = _loss_terms(_losses(model, x))
    (err + 0.7 * kl).backward()
    ref = arena.flat_g[:arena.n_pad]
    for rank in (0, 1):
        g, gerr, gkl, ptrs, base = out[rank]
        assert torch.allclose(g, ref, rtol=1e-5, atol=1e-6), (g - ref).abs().max()
        assert torch.allclose(gerr, err.detach(), rtol=1e-6) and torch.allclose(gkl, kl.detach(), rtol=1e-6)
        assert all(p >= base for p in ptrs)              # .grad tensors are views into the arena
    assert torch.equal(out[0][0], out[1][0])             # every rank ends with identical gradients


def test_arena_alignment_and_views():
    from genesis_b200.trainer import FlatArena
    model = _model()
    before = [p.detach().clone() for p in model.parameters()]
    arena = FlatArena(list(model.parameters()))
    assert arena.n_pad % FlatArena.ALIGN == 0 and arena.flat_g.numel() == arena.n_pad + FlatArena.TAIL
    for p, b in zip(model.parameters(), before):
        assert torch.equal(p.detach(), b)
        assert (p.data_ptr() - arena.flat_p.data_ptr()) % (4 * FlatArena.ALIGN) == 0
    arena.flat_p.mul_(2.0)                               # parameters are views: the arena IS the storage
    for p, b in zip(model.parameters(), before):
        assert torch.equal(p.detach(), 2 * b)


def test_shard_batch_rejects_ragged():
    import pytest
    from genesis_b200.trainer import shard_batch
    x = torch.zeros(6, 1)
    assert shard_batch(x, 1, 3).shape[0] == 2
    with pytest.raises(AssertionError):
        shard_batch(x, 0, 4)


def _noise_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from genesis_b200 import noise
    torch.manual_seed(0)                                 # every rank constructs the model under the same seed ...
    model = _model()
    noise.seed_rank(1234, rank, 'cpu')                   # ... and then takes its own noise stream (trainer.TrainStep does this)
    src = noise.NoiseMixin()
    like = next(model.parameters())
    out[rank] = (torch.cat([p.detach().flatten() for p in model.parameters()]), src._normal((4, 8), like), src._uniform((4, 8), like))
    dist.destroy_process_group()


def test_ranks_share_parameters_but_draw_independent_noise():
    """Data-parallel ranks must not replay the same eps / IC-SBP seed noise on every shard (correlated noise across the global
    batch): identical parameters, different noise -- the reference's nn.DataParallel replicas draw per-device noise too."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_noise_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    p0, n0, u0 = out[0]
    p1, n1, u1 = out[1]
    assert torch.equal(p0, p1)
    assert not torch.equal(n0, n1) and not torch.equal(u0, u1)
    assert abs(float((n0 * n1).mean())) < 0.5            # and not merely shifted copies
