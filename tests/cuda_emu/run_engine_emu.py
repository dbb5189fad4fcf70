"""TEST INFRASTRUCTURE: one training step (forward + backward) of a plug-in on the CPU with EVERY kernel call routed to the CPU
emulation of the kernel sources (tests/cuda_emu/emu_lib.py), compared with the oracle on the same parameters, input and
noise: loss terms, masks and the gradient of every parameter.  The path exercised is the product's -- plug-in ->
genesis_b200.ops autograd Functions -> C-ABI argument marshalling -> kernel source -- minus the GPU.

    python tests/cuda_emu/run_engine_emu.py MODEL K B [key=value ...] [--fused-latent] [--direct] [--precision fp32|tf32]
                                            [--param-add name=value] [--gen multid]

Minutes per run (a 64x64 model is a few GFLOP of scalar C++); used by hand and by the opt-in test
tests/test_engine_emu.py (G2_RUN_EMU_MODELS=1).  The exact-fp32 SIMT kernels are the default; --precision tf32 runs the
tcgen05 kernels under the functional model of cuda_emu_sm100.h (slower)."""
import math
import os
import sys
import time

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cpu_ops_mock  # noqa: E402
import emu_lib  # noqa: E402
import util_parity as U  # noqa: E402
from genesis_b200 import ops  # noqa: E402
from oracle import functional as O  # noqa: E402
from oracle import models as M  # noqa: E402
from oracle import synth  # noqa: E402
from test_oracle_golden import build_engine_model  # noqa: E402


def parse(argv):
    model, K, B = argv[0], int(argv[1]), int(argv[2])
    over, flags, padd = {}, {'precision': 'fp32', 'gen': 'multid'}, {}
    it = iter(argv[3:])
    for a in it:
        if a == '--fused-latent':
            flags['fused'] = True
        elif a == '--direct':
            flags['direct'] = True
        elif a in ('--precision', '--gen'):
            flags[a[2:]] = next(it)
        elif a == '--param-add':
            k, v = next(it).split('=')
            padd[k] = float(v)
        else:
            k, v = a.split('=')
            over[k] = {'True': True, 'False': False}.get(v, v)
    return model, K, B, over, flags, padd


def main():
    model, K, B, over, flags, padd = parse(sys.argv[1:])
    mp = pytest.MonkeyPatch()
    emu = emu_lib.install(mp)
    ops.set_precision(flags['precision'])
    ops.set_fused_latent(bool(flags.get('fused')))
    m, cfg = build_engine_model(model, K, 64, **over)
    with torch.no_grad():
        for k, v in padd.items():
            dict(m.named_parameters())[k].add_(v)
    m.train()
    x = torch.from_numpy(synth.GENERATORS[flags['gen']](B, 64, 5)[0])
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    if flags.get('direct'):     # trainer.TrainStep's mode: kernels accumulate weight / bias gradients straight into pre-zeroed .grad
        for p_ in m.parameters():
            p_.grad = torch.zeros_like(p_)
        ops.set_side_streams(False)
        ops.set_direct_grad(True)
    m.set_noise_tape(O.NoiseTape(seed=3))
    t0 = time.time()
    out = m(x.as_subclass(cpu_ops_mock.AsCuda))
    t1 = time.time()
    losses = out[1]
    if model == 'vae':
        (losses['err'].mean(0) + losses['kl_l'].mean(0)).backward()
    else:
        U.engine_total_loss(losses).backward()
    t2 = time.time()
    ops.set_direct_grad(False)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
    fwd = M.vae_forward if model == 'vae' else M.FORWARD[model]
    ref = fwd(P, x, O.NoiseTape(seed=3), cfg, training=True)
    if model == 'vae':
        (ref['err'].mean(0) + ref['kl_l'].mean(0)).backward()
    else:
        M.total_loss(ref).backward()
    tf32 = flags['precision'] == 'tf32'
    names = sorted(set(emu.calls))
    print('%s K=%d B=%d %s %s: forward %.0f s, backward %.0f s, %d kernel calls of %d entry points' %
          (model, K, B, over, {k: v for k, v in flags.items() if k != 'gen'}, t1 - t0, t2 - t1, len(emu.calls), len(names)))
    e_err = U.rel_l2(losses['err'], ref['err'])
    print('  err rel-L2 %.2e' % e_err)
    assert e_err < (2e-3 if tf32 else 1e-5)
    for key in ('kl_l_k', 'kl_m_k'):
        if key in losses and len(losses[key]) and key in ref:
            a, b = torch.stack(list(losses[key]), 0), torch.stack(list(ref[key]), 0)
            d = (a.detach() - b.detach()).abs().max().item()
            print('  %s max |delta| %.2e' % (key, d))
            assert d < (5e-2 if tf32 else 2e-3) * (1 + b.detach().abs().max().item())
    for key in ('kl_m', 'kl_l'):
        if key in losses and torch.is_tensor(losses[key]) and key in ref:
            d = (losses[key].detach() - ref[key].detach()).abs().max().item()
            print('  %s max |delta| %.2e' % (key, d))
            assert d < (5e-2 if tf32 else 2e-3) * (1 + ref[key].detach().abs().max().item())
    if 'log_m_k' in ref and out[2] is not None and 'log_m_k' in out[2]:
        a, b = torch.stack(list(out[2]['log_m_k']), 0).detach(), torch.stack(list(ref['log_m_k']), 0).detach()
        assert a.shape == b.shape, (a.shape, b.shape)
        pad = b < -1e9
        d = ((a - b).abs() / (1 + b.abs()))[~pad].max().item()
        print('  log-masks max |delta|/(1+|ref|) %.2e; padded slots equal: %s' % (d, bool(((a < -1e9) == pad).all())))
        assert d < (5e-3 if tf32 else 2e-4) and ((a < -1e9) == pad).all()
    worst = U.compare_grads(m, P, tol=0.3 if tf32 else 1e-2)
    print('  worst per-tensor gradient rel-L2 %.2e (%s); whole-vector %.2e' % (worst[0], worst[1], U.global_grad_rel_l2(m, P)))
    print('  entry points:', ' '.join(n[3:] for n in names))
    print('OK')


if __name__ == '__main__':
    main()
