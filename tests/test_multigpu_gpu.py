"""GPU, 2 ranks (skipped on a single-GPU box; run with `gpurun --gpus 2`): the data-parallel step of trainer.TrainStep.

SURVEY.md section 8e: images are independent given the parameters for GENESIS-V2 and MONet (GroupNorm / InstanceNorm / LayerNorm
are per sample), so the gradient the ranks hold after the ONE all-reduce over the flat arena must equal the gradient of a single
process on the full batch (up to fp32 summation order); GENESIS-V1's BatchNorm statistics are per rank (the reference's
nn.DataParallel semantics), so there each rank is compared with a single process on ITS shard and the all-reduced gradient with
the sum of the two.  Also: identical parameters on both ranks after optimiser steps, independent noise per rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _grads_full_batch(name, K, img, x, seed, full_B, rows=None):
    """Single-process reference: gradient SUM over the images of x (loss = sum over images of err + kl), flat arena order."""
    import full_size_props as FP
    from genesis_b200 import ops, trainer
    from test_oracle_golden import build_engine_model
    m, cfg = build_engine_model(name, K, img)
    m = m.cuda().train()
    ts = trainer.TrainStep(m, world_size=1, geco=False, beta=1.0)
    B = x.shape[0]
    m.set_noise_tape(FP.RowTape(seed, full_B, rows if rows is not None else list(range(B)), K))
    ops.set_direct_grad(True)
    try:
        out = m(x.cuda())
        err, kl = ts.loss_terms(out[1])
        ((err + kl) * B).backward()                    # sum over images
        ops.join_grad_stream(x.device if x.is_cuda else torch.device('cuda', torch.cuda.current_device()))
    finally:
        ops.set_direct_grad(False)
        m.set_noise_tape(None)
    torch.cuda.synchronize()
    return ts.flat_g[:ts.n_pad].detach().clone().cpu(), ts


def _worker(rank, world, port, name, K, img, x, seed, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import full_size_props as FP
    import util_parity as U
    from genesis_b200 import ops, trainer
    from test_oracle_golden import build_engine_model
    m, cfg = build_engine_model(name, K, img)
    m = m.cuda().train()
    ts = trainer.TrainStep(m, world_size=world, rank=rank, geco=False, beta=1.0, noise_seed=7, optimiser='sgd', lr=1e-3)
    xs = trainer.shard_batch(x, rank, world).cuda()
    per = xs.shape[0]
    rows = list(range(rank * per, (rank + 1) * per))
    m.set_noise_tape(FP.RowTape(seed, x.shape[0], rows, K))
    ops.set_direct_grad(True)
    try:
        o = m(xs)
        err, kl = ts.loss_terms(o[1])
        ((err + kl) * per).backward()                  # local SUM over the shard's images
        ops.join_grad_stream(xs.device)
    finally:
        ops.set_direct_grad(False)
        m.set_noise_tape(None)
    local = ts.flat_g[:ts.n_pad].detach().clone().cpu()
    ts.arena.exchange(err.detach(), kl.detach(), world)
    torch.cuda.synchronize()
    summed = ts.flat_g[:ts.n_pad].detach().clone().cpu()
    # two real optimiser steps with production noise: parameters must stay identical across ranks, noise must differ
    ts.flat_g.zero_()
    eps = torch.randn(4, device='cuda').cpu()
    for _ in range(2):
        ts.step(xs)
    torch.cuda.synchronize()
    # the same two steps with the gradient all-reduce NOT overlapped with the backward pass (one collective after it): the
    # bucketed, overlapped exchange must give the same parameters
    from genesis_b200 import noise
    m2, _ = build_engine_model(name, K, img)
    ts2 = trainer.TrainStep(m2.cuda().train(), world_size=world, rank=rank, geco=False, beta=1.0, noise_seed=7, overlap=False,
                            optimiser='sgd', lr=1e-3)
    torch.randn(4, device='cuda')
    for _ in range(2):
        ts2.step(xs)
    torch.cuda.synchronize()
    same = U.rel_l2(ts2.flat_p, ts.flat_p)
    fired = len(ts._fired)
    out[rank] = (local, summed, ts.flat_p.detach().clone().cpu(), eps, float(ts.elbo), same, fired, len(ts.buckets))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('name,K,img,B,gen', [('genesisv2', 4, 64, 4, 'stacks'), ('monet', 3, 64, 4, 'multid'), ('genesis', 3, 64, 4, 'multid')])
def test_two_rank_allreduced_gradient_equals_single_process(name, K, img, B, gen):
    import util_parity as U
    from genesis_b200.datasets import synth
    x = torch.from_numpy(synth.GENERATORS[gen](B, img, 21)[0])
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), name, K, img, x, 5, out), nprocs=2, join=True)
    (l0, s0, p0, e0, elbo0, same0, fired0, nb0), (l1, s1, p1, e1, elbo1, same1, fired1, nb1) = out[0], out[1]
    # overlapped == non-overlapped exchange, up to the run-to-run noise of the step (float-atomic reductions; for GENESIS-V2 an
    # IC-SBP argmax can move): a bucket sent before its gradients were complete would show as >= 1e-3
    assert same0 < 2e-4 and same1 < 2e-4, (same0, same1)
    assert fired0 >= 1 and fired0 == fired1, (fired0, nb0)   # at least one bucket went out under the backward pass
    assert torch.equal(s0, s1)                                   # one all-reduce: both ranks hold the same summed gradient
    assert torch.equal(p0, p1)                                   # ... and stay bit-identical through optimiser steps
    assert elbo0 == elbo1                                        # the ELBO / GECO inputs travel in the arena tail
    assert not torch.equal(e0, e1)                               # independent production noise per rank
    assert U.rel_l2(l0 + l1, s0) < 1e-6
    if name == 'genesis':
        # BatchNorm statistics are per rank: compare each rank with a single process on ITS shard
        for rank, local in ((0, l0), (1, l1)):
            ref, _ = _grads_full_batch(name, K, img, x[rank * 2:(rank + 1) * 2], 5, B, rows=[rank * 2, rank * 2 + 1])
            assert U.rel_l2(local, ref) < 2e-3, rank
    else:
        ref, _ = _grads_full_batch(name, K, img, x, 5, B)
        assert U.rel_l2(s0, ref) < 2e-3                          # per-sample norms: exact batch-partition invariance (8e)
