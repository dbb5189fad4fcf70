"""TEST INFRASTRUCTURE: run g2_conv_wgrad_tf32 of genesis_b200/csrc/wgrad_tc.cu under the CPU emulation on one stride-1 case and
compare with torch's conv weight gradient (integer-valued data: exact).

    python tests/cuda_emu/run_wgrad_emu.py "N Hg Wg Cg Ct R pad; ..." [halo]

`halo` sets G2_WGRAD_HALO=1 (the experimental resident-window kernel); without it the B200-validated tile kernel runs, which
calibrates the model of MN-major operands."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build_emu  # noqa: E402


def run_case(lib, case):
    N, Hg, Wg, Cg, Ct, R, pad = case
    Ht, Wt = Hg + 2 * pad - R + 1, Wg + 2 * pad - R + 1
    rng = np.random.RandomState(1)
    x = rng.randint(-3, 4, (N, Hg, Wg, Cg)).astype(np.float32)          # G, NHWC
    dy = rng.randint(-2, 3, (N, Ht, Wt, Ct)).astype(np.float32)         # T, NHWC
    w = torch.zeros(Ct, Cg, R, R, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(torch.from_numpy(x).double().permute(0, 3, 1, 2), w, None, padding=pad)
    (ref,) = torch.autograd.grad((y * torch.from_numpy(dy).double().permute(0, 3, 1, 2)).sum(), [w])      # [Ct, Cg, R, R]
    if os.environ.get('G2_WGRAD_HALO') == '1':       # the case must really take the new kernel (unsupported shapes fall through)
        plan = (ctypes.c_int * (16 + 7 * 12))()
        lib.g2_conv_wgrad_halo_plan(N, Hg, Wg, Cg, Ht, Wt, Ct, R, R, 1, plan)
        assert plan[0] == 1, 'shape not covered by the halo plan'
        print('halo plan: TH %d Wp %d groups %d cbs/CTA %d grid_y %d splits %d windows/CTA %d' %
              (plan[1], plan[2], plan[6], plan[7], plan[8], plan[9], plan[10]))
    ws_bytes = lib.g2_conv_wgrad_tf32_workspace(N, Hg, Wg, Cg, Ht, Wt, Ct, R, R, 1)
    assert ws_bytes > 0
    ws = np.full(ws_bytes // 4, np.nan, np.float32)
    dw = np.full((R, R, Cg, Ct), np.nan, np.float32)
    P = ctypes.c_void_p
    rc = lib.g2_conv_wgrad_tf32(x.ctypes.data_as(P), dy.ctypes.data_as(P), dw.ctypes.data_as(P), ws.ctypes.data_as(P),
                                N, Hg, Wg, Cg, Ht, Wt, Ct, R, R, 1, pad, 0, None)
    assert rc == 0, rc
    got = torch.from_numpy(dw).permute(3, 2, 0, 1).double()
    assert torch.isfinite(got).all(), 'unwritten outputs'
    err = (got - ref).abs().max().item()
    print('max abs err', err)
    assert err == 0.0
    print('OK', case)


def main():
    if len(sys.argv) > 2 and sys.argv[2] == 'halo':
        os.environ['G2_WGRAD_HALO'] = '1'
    lib = ctypes.CDLL(build_emu.build('wgrad_tc.cu'))
    lib.g2_conv_wgrad_tf32_workspace.restype = ctypes.c_long
    for c in sys.argv[1].split(';'):
        if c.strip():
            run_case(lib, tuple(int(a) for a in c.split()))


if __name__ == '__main__':
    main()
