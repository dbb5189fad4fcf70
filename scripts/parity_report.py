"""Compact parity table: engine (CUDA) vs oracle (CPU) on identical parameters, inputs and noise, for every model and
precision mode.  Run on the GPU box:  python scripts/parity_report.py [model ...] [--modes tf32,tf32x3,fp32] > gpurun_out/parity.txt
  tf32   = plain TF32 tensor-core operands everywhere (round 1's product path)
  tf32x3 = the product default: 3xTF32 (ops.precise) in the GENESIS-V2 backbone / MONet attention UNet, plain TF32 elsewhere
  fp32   = exact-fp32 SIMT kernels everywhere
  tf32x3:fwd+dgrad+wgrad / ... = 3xTF32 also in the named backward contractions of the precise layers (experiments; the
           default is the forward contraction only -- measured to be what matters)
  tf32x3+dec / tf32x3+enc = additionally the GENESIS-V2 decoder / the MONet component encoder as precise layers"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402

import util_parity as U  # noqa: E402
from oracle import synth  # noqa: E402
from test_oracle_golden import build_engine_model  # noqa: E402
from genesis_b200 import ops  # noqa: E402

CASES = [('genesis', 5, 4, 64, 'multid', 11), ('genesisv2', 7, 4, 64, 'stacks', 11), ('genesisv2', 11, 2, 64, 'rooms', 12),
         ('monet', 7, 3, 64, 'multid', 11), ('monet', 3, 2, 128, 'multid', 11)]      # last = data seed (tests/test_v2_monet_gpu.py)
MODES = ['tf32', 'tf32x3', 'fp32']
_args = sys.argv[1:]
if '--modes' in _args:
    i = _args.index('--modes')
    MODES = _args[i + 1].split(',')
    _args = _args[:i] + _args[i + 2:]
if _args:
    CASES = [c for c in CASES if c[0] in _args]


def main():
    for model, K, B, img, gen, dseed in CASES:
        for prec in MODES:
            if model == 'genesis' and prec.startswith('tf32x3'):
                continue                      # GENESIS has no precise layers: tf32x3 == tf32
            ops.set_precision('fp32' if prec == 'fp32' else 'tf32')
            parts = prec.split(':')[1].split('+') if ':' in prec else ['fwd']
            prec_tag = prec
            ops.set_precise_parts('fwd' in parts, 'dgrad' in parts, 'wgrad' in parts)
            m, cfg = build_engine_model(model, K, img, seed=3)
            m = m.cuda().train()
            for attr in ('precise_unet', 'precise_backbone'):
                if hasattr(m, attr):
                    setattr(m, attr, prec.startswith('tf32x3'))
            for attr in ('precise_decoder', 'precise_comp_encoder'):
                if hasattr(m, attr):
                    setattr(m, attr, prec.startswith('tf32x3') and ('+dec' in prec or '+enc' in prec))
            if model == 'genesisv2':
                with torch.no_grad():
                    m.att_process.colour_head.gate.gate.fill_(0.3)
            sd0 = {k: v.clone() for k, v in m.state_dict().items()}
            x = torch.from_numpy(synth.GENERATORS[gen](B, img, dseed)[0])
            tape = U.make_tape(5)
            out, P = U.run_oracle(model, sd0, x, tape, cfg)
            recon, losses, stats, att, comp = U.run_engine(m, x, tape.rewound())
            print('== %s K=%d B=%d img=%d %s [%s]' % (model, K, B, img, gen, prec))
            if model == 'genesisv2':
                print('   oracle seed-pixel ReLU kink margin %.2e' % out['att']['seed_kink_margin'])
            print('   err rel %.2e | recon rel %.2e' % (U.rel_l2(losses['err'], out['err']), U.rel_l2(recon, out['recon'])))
            for key in ('kl_l_k', 'kl_m_k'):
                if key in out and key in losses:
                    a = torch.stack(list(losses[key]), 0).detach().cpu()
                    b = torch.stack(out[key], 0).detach()
                    print('   %s max abs %.2e (max |ref| %.2e)' % (key, (a - b).abs().max().item(), b.abs().max().item()))
            if 'kl_m' in out and 'kl_m' in losses:
                print('   kl_m rel %.2e' % U.rel_l2(losses['kl_m'], out['kl_m']))
            for key in ('log_m_k', 'log_m_r_k'):
                if key in out and key in stats:
                    a = torch.stack(list(stats[key]), 0).detach().cpu()
                    b = torch.stack(out[key], 0).detach()
                    d = (a - b).abs()
                    print('   %s max abs %.2e, max rel-to-(1+|ref|) %.2e, min ref %.1f' % (
                        key, d.max().item(), (d / (1 + b.abs())).max().item(), b.min().item()))
            rows = []
            gmax = max(p.grad.norm().item() for p in P.values() if torch.is_tensor(p) and p.grad is not None)
            for name, p in m.named_parameters():
                ref = P[name].grad
                if ref is None or p.grad is None:
                    continue
                e = (p.grad.detach().double().cpu() - ref.double()).norm().item() / max(ref.double().norm().item(), 2e-4 * gmax)
                rows.append((e, name, ref.norm().item()))
            rows.sort(reverse=True)
            num = sum((p.grad.detach().double().cpu() - P[n].grad.double()).pow(2).sum().item()
                      for n, p in m.named_parameters() if p.grad is not None and P[n].grad is not None)
            den = sum(P[n].grad.double().pow(2).sum().item() for n, p in m.named_parameters() if P[n].grad is not None)
            print('   GLOBAL grad rel-L2 %.2e' % ((num / den) ** 0.5))
            print('   worst grads: ' + '; '.join('%s %.2e (|ref| %.1e)' % (n, e, r) for e, n, r in rows[:4]))
            med = sorted(e for e, _, _ in rows)[len(rows) // 2]
            print('   median grad rel-L2 %.2e over %d tensors' % (med, len(rows)))


if __name__ == '__main__':
    main()
