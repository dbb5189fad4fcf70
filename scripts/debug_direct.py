import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from test_direct_grad_gpu import grads_of
from test_oracle_golden import build_engine_model
from oracle import synth
model, K, img, B, gen = sys.argv[1], int(sys.argv[2]), 64, int(sys.argv[3]), 'multid'
m, cfg = build_engine_model(model, K, img)
m = m.cuda().train()
x = torch.from_numpy(synth.GENERATORS[gen](B, img, 5)[0]).cuda()
a, _ = grads_of(m, x, 11, False)
b, _ = grads_of(m, x, 11, False)
c, _ = grads_of(m, x, 11, True, 0.25)
def rel(u, v):
    return ((u - v).norm() / (v.norm() + 1e-20)).item()
print('%-55s %10s %10s %10s' % ('param', 'auto-auto', 'dir-auto', '|g|'))
for n in a:
    if a[n] is None: continue
    print('%-55s %10.2e %10.2e %10.2e' % (n, rel(b[n], a[n]), rel(c[n] - 0.25, a[n]), a[n].norm().item()))
