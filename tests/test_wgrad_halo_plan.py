"""CPU: the host-side plan of the experimental halo weight-gradient kernel (genesis_b200/csrc/wgrad_tc.cu, namespace wgh;
G2_WGRAD_HALO=1, not yet run on a B200), replayed in numpy.

g2_conv_wgrad_halo_plan returns what the kernel receives: window rows TH, pitch Wp, G-window rows, the split of windows and
channel blocks over the grid, and the tap groups (first row, LBO in pixel rows, valid atoms, tap per atom).  The replay
performs the kernel's index arithmetic -- one zero-filled TMA box per (window, channel block) of G and per channel block of
T landing as flat pixel rows, slack rows behind the window zeroed, every atom j of a group reading rows
[off + j * lbo, off + j * lbo + 8 * ksteps) against T rows [0, 8 * ksteps), accumulation over the CTA's windows, per-split
partials reduced afterwards -- and must reproduce torch's conv weight gradient exactly on integer data.  It also checks
the resource limits the kernel relies on (TMEM columns, shared memory, slack covers every read, box dims)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

CASES = [  # N, Hg, Wg, Cg, Ct, R, pad          (stride 1; Ht = Hg + 2 pad - R + 1)
    (2, 64, 64, 64, 32, 5, 2),        # GENESIS mask decoder conv-transpose as a correlation: the largest wgrad of config c2
    (2, 70, 70, 32, 32, 3, 0),        # broadcast decoder (VALID 3x3)
    (2, 66, 66, 32, 32, 3, 0),
    (3, 32, 32, 64, 32, 5, 2),
    (3, 16, 16, 128, 64, 5, 2),
    (2, 32, 32, 64, 64, 3, 1),        # UNet block
    (1, 19, 23, 32, 32, 3, 1),        # ragged
    (1, 128, 128, 32, 32, 3, 1),      # MONet-128
    (80, 19, 23, 32, 32, 3, 1),       # more windows than CTAs: accumulation over a CTA's windows
]


def plan_of(N, Hg, Wg, Cg, Ct, R, pad):
    from genesis_b200 import _lib
    Ht, Wt = Hg + 2 * pad - R + 1, Wg + 2 * pad - R + 1
    buf = (ctypes.c_int * (16 + 7 * 12))()
    n = _lib.lib().query('g2_conv_wgrad_halo_plan', N, Hg, Wg, Cg, Ht, Wt, Ct, R, R, 1, ctypes.cast(buf, ctypes.c_void_p))
    v = list(buf)
    if not v[0]:
        return None
    keys = ['TH', 'Wp', 'g_rows', 'win_per_img', 'windows', 'ngroups', 'cbs_per_cta', 'grid_y', 'splits', 'wpc', 'slack_rows',
            'ksteps', 'g_buf', 't_blk', 'smem']
    d = dict(zip(keys, v[1:16]))
    d['groups'] = [dict(off=v[16 + 7 * g], lbo=v[17 + 7 * g], n=v[18 + 7 * g], taps=v[19 + 7 * g:23 + 7 * g]) for g in range(d['ngroups'])]
    assert n == 16 + 7 * d['ngroups']
    d['Ht'], d['Wt'] = Ht, Wt
    return d


@pytest.mark.parametrize('case', CASES, ids=[str(c) for c in CASES])
def test_wgrad_halo_plan_replay_matches_torch(case, built_lib):
    N, Hg, Wg, Cg, Ct, R, pad = case
    pl = plan_of(*case)
    assert pl is not None
    TH, Wp, g_rows, Ht, Wt = pl['TH'], pl['Wp'], pl['g_rows'], pl['Ht'], pl['Wt']
    # resource limits
    assert pl['smem'] <= 227 * 1024 and pl['g_buf'] % 1024 == 0 and pl['t_blk'] % 1024 == 0
    assert pl['ngroups'] * pl['cbs_per_cta'] * Ct <= 512
    assert Wp <= 256 and g_rows <= 256 and (TH * Wp) % 8 == 0 and pl['ksteps'] * 8 == TH * Wp
    assert pl['splits'] * pl['grid_y'] <= 148 and pl['splits'] * pl['wpc'] >= pl['windows'] > (pl['splits'] - 1) * pl['wpc']
    assert pl['grid_y'] * pl['cbs_per_cta'] >= Cg // 32
    assert sorted(t for g in pl['groups'] for t in g['taps'][:g['n']]) == list(range(R * R))      # every tap exactly once
    alloc_rows = pl['g_buf'] // 128
    assert alloc_rows >= g_rows * Wp + pl['slack_rows']

    rng = np.random.RandomState(1)
    x = rng.randint(-3, 4, (N, Hg, Wg, Cg)).astype(np.float64)          # G, NHWC
    dy = rng.randint(-2, 3, (N, Ht, Wt, Ct)).astype(np.float64)         # T, NHWC
    ws = np.zeros((pl['splits'], R * R, Cg, Ct))
    F_ = TH * Wp
    for split in range(pl['splits']):
        for win in range(split * pl['wpc'], min((split + 1) * pl['wpc'], pl['windows'])):
            n, h0 = win // pl['win_per_img'], (win % pl['win_per_img']) * TH
            # T box {32 ch, Wp, TH}: out-of-bounds columns / rows are zero
            T = np.zeros((F_, Ct))
            for r in range(TH):
                if h0 + r < Ht:
                    T[r * Wp:r * Wp + Wt] = dy[n, h0 + r]
            for cb in range(Cg // 32):
                buf = np.full((alloc_rows, 32), np.nan)                 # anything TMA / the zero fill does not write would poison the sums
                buf[g_rows * Wp:] = 0.0                                 # slack zeroed by the kernel's prologue
                for r in range(g_rows):
                    for c in range(Wp):
                        hh, ww = h0 - pad + r, c - pad
                        buf[r * Wp + c] = x[n, hh, ww, cb * 32:cb * 32 + 32] if (0 <= hh < Hg and 0 <= ww < Wg) else 0.0
                for g in pl['groups']:
                    for j in range(4):
                        lo = g['off'] + j * g['lbo']
                        assert lo + F_ <= alloc_rows                   # every atom (valid or not) stays inside the buffer
                        acc = buf[lo:lo + F_].T @ T                     # [32, Ct]; unused atoms are computed and dropped
                        assert np.isfinite(acc).all()
                        if j < g['n']:
                            ws[split, g['taps'][j], cb * 32:cb * 32 + 32] += acc
    dW = ws.sum(0)                                                      # wgrad_reduce_kernel
    w = torch.zeros(Ct, Cg, R, R, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), w, None, padding=pad)
    (ref,) = torch.autograd.grad((y * torch.from_numpy(dy).permute(0, 3, 1, 2)).sum(), [w])      # [Ct, Cg, R, R]
    got = dW.reshape(R, R, Cg, Ct).transpose(3, 2, 0, 1)
    np.testing.assert_array_equal(got, ref.numpy())


def test_wgrad_halo_plan_rejects_what_the_tile_kernel_keeps(built_lib):
    assert plan_of(2, 64, 64, 32, 32, 1, 0) is None          # 1x1: channel blocks are packed by the tile kernel instead
    assert plan_of(2, 8, 8, 32, 32, 3, 1) is None            # tiny maps
    assert plan_of(2, 16, 16, 128, 128, 5, 2) is None        # 7 groups x 128 columns exceed TMEM


def test_wgrad_halo_mma_work_estimate(built_lib):
    """MMA instructions per unit of useful work: the halo plan trades some padding (pitch, unused atoms) for R*S-fold less
    operand traffic; keep the overhead visible."""
    for case, bound in [((2, 64, 64, 64, 32, 5, 2), 1.25), ((2, 70, 70, 32, 32, 3, 0), 1.45)]:
        pl = plan_of(*case)
        N, Hg, Wg, Cg, Ct, R, pad = case
        issued = pl['windows'] * (Cg // 32) * pl['ngroups'] * pl['ksteps']
        ideal = N * pl['Ht'] * pl['Wt'] / 8 * (R * R * (Cg // 32) / 4)
        assert issued / ideal <= bound, (case, issued / ideal)
