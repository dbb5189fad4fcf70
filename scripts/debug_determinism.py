import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from genesis_b200 import _lib, ops
from test_direct_grad_gpu import grads_of
from test_oracle_golden import build_engine_model
from oracle import synth
import scripts.conv_bench as cb

lib = _lib.lib()
dev = 'cuda'
# 1. kernel-level bitwise determinism
for name in ('c2_bdec_fwd70', 'c2_att64_fwd', 'c2_att64_dgrad', 'c2_att_up64_fwd', 'c2_enc64_s2', 'c3_unet8'):
    mode, N, H, W, Ci, Co, R, s, p = cb.SHAPES[name]
    N = 8
    Ho, Wo = cb.out_hw(mode, H, W, R, s, p)
    torch.manual_seed(0)
    x = torch.randn(N, H, W, Ci, device=dev); w = torch.randn(R * R, Co, Ci, device=dev) * 0.05; b = torch.randn(Co, device=dev)
    outs = []
    for i in range(3):
        out = torch.full((N, Ho, Wo, Co), float('nan'), device=dev)
        _lib.call('g2_conv_igemm_tf32', x, w, b, out, N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, 2)
        outs.append(out)
    torch.cuda.synchronize()
    print(name, 'conv bitwise equal:', bool((outs[0] == outs[1]).all() and (outs[0] == outs[2]).all()))
    if mode == 0:
        dims = (N, H, W, Ci, Ho, Wo, Co); g, t = x, outs[0]
    else:
        dims = (N, Ho, Wo, Co, H, W, Ci); g, t = outs[0], x
    wsb = lib.query('g2_conv_wgrad_tf32_workspace', *dims, R, R, s)
    if wsb > 0:
        ds = []
        for i in range(3):
            ws = torch.empty(wsb // 4, device=dev); dw = torch.empty(R, R, Ci, Co, device=dev)
            _lib.call('g2_conv_wgrad_tf32', g, t, dw, ws, *dims, R, R, s, p, 0)
            ds.append(dw)
        torch.cuda.synchronize()
        print(name, 'wgrad bitwise equal:', bool((ds[0] == ds[1]).all() and (ds[0] == ds[2]).all()))
# 2. model-level run-to-run difference per precision mode
for prec in ('fp32', 'tf32'):
    ops.set_precision(prec)
    m, cfg = build_engine_model('genesis', 3, 64)
    m = m.cuda().train()
    x = torch.from_numpy(synth.GENERATORS['multid'](4, 64, 5)[0]).cuda()
    a, _ = grads_of(m, x, 11, False); b, _ = grads_of(m, x, 11, False)
    worst = sorted(((((b[n] - a[n]).norm() / (a[n].norm() + 1e-20)).item(), n) for n in a if a[n] is not None and a[n].norm() > 1e-3), reverse=True)[:4]
    print(prec, 'run-to-run worst:', worst)
ops.set_precision('tf32')
