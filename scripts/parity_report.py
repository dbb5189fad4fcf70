"""Compact parity table: engine (CUDA) vs oracle (CPU) on identical parameters, inputs and noise, for every model and
both precisions.  Run on the GPU box:  python scripts/parity_report.py > gpurun_out/parity.txt"""
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

CASES = [('genesis', 5, 4, 64, 'multid'), ('genesisv2', 7, 4, 64, 'stacks'), ('genesisv2', 11, 2, 64, 'rooms'),
         ('monet', 7, 3, 64, 'multid'), ('monet', 3, 2, 128, 'multid')]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] in sys.argv[1:]]


def main():
    for model, K, B, img, gen in CASES:
        for prec in ('tf32', 'fp32'):
            ops.set_precision(prec)
            m, cfg = build_engine_model(model, K, img, seed=3)
            m = m.cuda().train()
            if model == 'genesisv2':
                with torch.no_grad():
                    m.att_process.colour_head.gate.gate.fill_(0.3)
            sd0 = {k: v.clone() for k, v in m.state_dict().items()}
            x = torch.from_numpy(synth.GENERATORS[gen](B, img, 11)[0])
            tape = U.make_tape(5)
            out, P = U.run_oracle(model, sd0, x, tape, cfg)
            recon, losses, stats, att, comp = U.run_engine(m, x, tape.rewound())
            print('== %s K=%d B=%d img=%d %s [%s]' % (model, K, B, img, gen, prec))
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
