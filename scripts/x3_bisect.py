"""Debug: GENESIS-V2 K=11 rooms, gradients under the two 3xTF32 routes (in-kernel / pre-split), with and without a device
synchronisation after every C-ABI call, against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import util_parity as U
from oracle import synth
from test_oracle_golden import build_engine_model
from genesis_b200 import ops

model, K, B, img, gen = 'genesisv2', 11, 2, 64, 'rooms'
ops.set_precision('tf32')
res = {}
P = None
orig_call = ops._call
orig_fwd = ops._conv_fwd_common
SEL = {'mode': None}
def _fwd(ctx, x, w, b, stride, pad, act, transposed, bias_grad=True):
    if SEL['mode'] is not None:
        is6464 = tuple(x.shape[1:]) == (64, 64, 64) and tuple(w.shape) == (64, 64, 3, 3)
        ops._X3_INKERNEL = is6464 if SEL['mode'] == 'only6464' else not is6464
    return orig_fwd(ctx, x, w, b, stride, pad, act, transposed, bias_grad)
ops._conv_fwd_common = _fwd
TAGS = ('inkernel', 'presplit', 'only6464', 'allbut6464')
for tag in TAGS:
    ops._X3_INKERNEL = tag == 'inkernel'
    SEL['mode'] = tag if tag in ('only6464', 'allbut6464') else None
    m, cfg = build_engine_model(model, K, img, seed=3)
    m = m.cuda().train()
    with torch.no_grad():
        m.att_process.colour_head.gate.gate.fill_(0.3)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.from_numpy(synth.GENERATORS[gen](B, img, 11)[0])
    tape = U.make_tape(5)
    if P is None:
        out, P = U.run_oracle(model, sd0, x, tape, cfg)
    recon, losses, stats, att, comp = U.run_engine(m, x, U.make_tape(5))
    res[tag] = {n: p.grad.detach().double().cpu() for n, p in m.named_parameters() if p.grad is not None}
    res[tag + '/fwd'] = {'recon': recon.detach().double().cpu(), 'log_m': torch.stack(list(stats['log_m_k'])).detach().double().cpu()}
names = [n for n in res['inkernel'] if n in P and P[n].grad is not None and P[n].grad.norm() > 0]
for n in names:
    ref = P[n].grad.double()
    line = '%-40s' % n
    for tag in TAGS:
        line += ' %s %.3e' % (tag, ((res[tag][n] - ref).norm() / ref.norm()).item())
    line += ' | inkernel-presplit %.3e' % ((res['inkernel'][n] - res['presplit'][n]).norm() / ref.norm()).item()
    print(line)
for k in ('recon', 'log_m'):
    print(k, 'inkernel vs presplit max abs', (res['inkernel/fwd'][k] - res['presplit/fwd'][k]).abs().max().item())
g = res['inkernel']['seg_head.0.weight']; r = P['seg_head.0.weight'].grad.double(); q = res['presplit']['seg_head.0.weight']
d = (g - r)
print('seg_head.0.weight diff by output channel (top 8):', torch.topk(d.flatten(1).norm(dim=1), 8))
print('diff by input channel (top 8):', torch.topk(d.permute(1, 0, 2, 3).flatten(1).norm(dim=1), 8))
print('diff by tap:', d.permute(2, 3, 0, 1).flatten(2).norm(dim=2))
print('ref by tap:', r.permute(2, 3, 0, 1).flatten(2).norm(dim=2))

gb = res['inkernel']['seg_head.1.bias']; rb = P['seg_head.1.bias'].grad.double(); qb = res['presplit']['seg_head.1.bias']
print('seg_head.1.bias grad: ref', rb.tolist())
print('inkernel-ref', (gb - rb).tolist())
print('presplit-ref', (qb - rb).tolist())
