"""CPU: the host-side launch plan of the halo implicit-GEMM kernel (genesis_b200/csrc/igemm_halo.cu), replayed in numpy.

g2_conv_halo_plan returns exactly what the kernel receives (tile rows, window pitch, chunking, flat tap offsets).
The replay below performs the kernel's index arithmetic -- TMA boxes with zero out-of-bounds fill landing as a flat
run of pixel rows, M tiles of 128 consecutive flat positions, taps as row shifts, the epilogue's validity mask --
and must reproduce torch's conv2d / conv_transpose2d exactly on integer data.  No GPU, no kernel launch."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

CASES = [  # mode, N, H, W, Ci, Co, R, stride, pad
    (0, 2, 72, 72, 32, 32, 3, 1, 0),       # broadcast decoder (VALID)
    (0, 2, 66, 66, 32, 32, 3, 1, 0),
    (1, 2, 68, 68, 32, 32, 3, 1, 0),       # its data gradient (full correlation)
    (1, 2, 64, 64, 32, 64, 5, 1, 2),       # GENESIS mask decoder 64x64
    (0, 2, 64, 64, 64, 32, 5, 1, 2),       # ... and its data gradient (two channel blocks)
    (1, 3, 16, 16, 64, 128, 5, 1, 2),      # small maps: several images per CTA
    (0, 5, 32, 32, 64, 64, 3, 1, 1),       # UNet block
    (0, 1, 128, 128, 32, 32, 3, 1, 1),     # MONet-128 UNet block
    (0, 1, 136, 136, 32, 32, 3, 1, 0),     # MONet-128 broadcast decoder
    (1, 2, 32, 32, 32, 64, 5, 2, 2),       # stride-2 conv-transpose: 4 sub-pixel classes
    (1, 3, 16, 16, 64, 64, 5, 2, 2),
    (1, 2, 32, 32, 32, 32, 3, 2, 1),
    (0, 3, 19, 23, 32, 32, 3, 1, 1),       # ragged sizes
]


def plan_of(case, cls):
    from genesis_b200 import _lib
    mode, N, H, W, Ci, Co, R, s, p = case
    if mode == 0:
        Ho, Wo = (H + 2 * p - R) // s + 1, (W + 2 * p - R) // s + 1
    else:
        Ho, Wo = (H - 1) * s - 2 * p + R + (s - 1), (W - 1) * s - 2 * p + R + (s - 1)
    buf = (ctypes.c_int * 96)()
    rc = _lib.lib().query('g2_conv_halo_plan', N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, cls, ctypes.cast(buf, ctypes.c_void_p))
    assert rc == 0, rc
    v = list(buf)
    keys = ['nclasses', 'TH', 'TNB', 'RH', 'ch_rows', 'nch', 'm', 'a_bytes', 'tiles_h', 'Wp', 'dh_min', 'dw_min', 'os', 'ph', 'pw',
            'Hv', 'Wv', 'ntaps', 'BN', 'smem', 'TW', 'tiles_w']
    d = dict(zip(keys, v[:22]))
    d['persistent'], d['resident'], d['pstages'], d['tmem_cols'] = v[22:26]
    d['toff'] = v[32:32 + d['ntaps']]
    d['widx'] = v[64:64 + d['ntaps']]
    d['Ho'], d['Wo'] = Ho, Wo
    return d


def replay(case, x_nhwc, wp, out):
    """x_nhwc [N,H,W,Ci], wp [R*R][Co][Ci] (the packed operand), out [N,Ho,Wo,Co] pre-filled with NaN."""
    mode, N, H, W, Ci, Co, R, s, p = case
    ncls = plan_of(case, 0)['nclasses']
    for cls in range(ncls):
        pl = plan_of(case, cls)
        TH, TNB, RH, Wp = pl['TH'], pl['TNB'], pl['RH'], pl['Wp']
        span_h = RH - TH
        alloc_pix = pl['a_bytes'] // 128
        assert pl['a_bytes'] % 1024 == 0 and pl['smem'] <= 227 * 1024 and pl['m'] * pl['BN'] <= 512
        assert pl['tmem_cols'] <= 512 and pl['persistent'] == {'1': 1, '0': 0}.get(os.environ.get('G2_HALO_PERSISTENT'), pl['persistent'])
        if pl['resident']:
            assert pl['pstages'] == pl['ntaps'] * (Ci // 32) and pl['pstages'] <= 64
        if TNB == 1:
            assert (pl['ch_rows'] * Wp * 128) % 1024 == 0 and pl['nch'] <= 8 and pl['nch'] * pl['ch_rows'] >= RH
        groups = (N + TNB - 1) // TNB
        for g in range(groups):
            n0 = g * TNB
            for th_i, tw_i in [(a, b) for a in range(pl['tiles_h']) for b in range(pl['tiles_w'])]:
                h0, w0 = th_i * TH, tw_i * pl['TW']
                rows_valid = min(TH, pl['Hv'] - h0)
                cols_valid = min(pl['TW'], pl['Wv'] - w0)
                imgs_valid = min(TNB, N - n0)
                f_last = ((imgs_valid - 1) * RH + rows_valid - 1) * Wp + cols_valid - 1
                m_tiles = f_last // 128 + 1
                assert m_tiles <= pl['m']
                for cb in range(Ci // 32):
                    A = np.full((alloc_pix, 32), np.nan, np.float64)       # shared-memory window (NaN = never written)

                    def box(dst_row, rows, imgs, hstart):
                        blk = np.zeros((imgs, rows, Wp, 32))
                        for i in range(imgs):
                            for r in range(rows):
                                h = hstart + r
                                if n0 + i >= N or h < 0 or h >= H:
                                    continue
                                for c in range(Wp):
                                    w = w0 + pl['dw_min'] + c
                                    if 0 <= w < W:
                                        blk[i, r, c] = x_nhwc[n0 + i, h, w, cb * 32:cb * 32 + 32]
                        A[dst_row:dst_row + imgs * rows * Wp] = blk.reshape(-1, 32)
                    if TNB > 1:
                        box(0, RH, TNB, h0 + pl['dh_min'])
                    else:
                        nch = min(pl['nch'], (rows_valid + span_h + pl['ch_rows'] - 1) // pl['ch_rows'])
                        for j in range(nch):
                            box(j * pl['ch_rows'] * Wp, pl['ch_rows'], 1, h0 + pl['dh_min'] + j * pl['ch_rows'])
                    acc = np.zeros((m_tiles * 128, Co)) if cb == 0 else acc
                    for tap in range(pl['ntaps']):
                        rows = A[pl['toff'][tap]:pl['toff'][tap] + m_tiles * 128]      # shifted operand, must stay in the allocation
                        assert rows.shape[0] == m_tiles * 128
                        with np.errstate(invalid='ignore'):
                            acc = acc + rows @ wp[pl['widx'][tap]][:, cb * 32:cb * 32 + 32].T.astype(np.float64)
                f = np.arange(m_tiles * 128)
                i = f // (RH * Wp)
                rem = f % (RH * Wp)
                hl, w = rem // Wp, rem % Wp
                valid = (i < imgs_valid) & (hl < rows_valid) & (w < cols_valid)
                oh, ow = (h0 + hl) * pl['os'] + pl['ph'], (w0 + w) * pl['os'] + pl['pw']
                for ff in f[valid]:
                    assert np.isnan(out[n0 + i[ff], oh[ff], ow[ff], 0]), 'output written twice'
                    out[n0 + i[ff], oh[ff], ow[ff]] = acc[ff]


@pytest.mark.parametrize('case', CASES)
def test_halo_plan_replay_matches_torch(case):
    mode, N, H, W, Ci, Co, R, s, p = case
    rng = np.random.RandomState(0)
    x = rng.randint(-3, 4, (N, Ci, H, W)).astype(np.float64)
    if mode == 0:
        w = rng.randint(-2, 3, (Co, Ci, R, R)).astype(np.float64)
        ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), None, stride=s, padding=p).numpy()
        wp = w.transpose(2, 3, 0, 1).reshape(R * R, Co, Ci)
    else:
        w = rng.randint(-2, 3, (Ci, Co, R, R)).astype(np.float64)
        ref = F.conv_transpose2d(torch.from_numpy(x), torch.from_numpy(w), None, stride=s, padding=p, output_padding=s - 1).numpy()
        wp = w.transpose(2, 3, 1, 0).reshape(R * R, Co, Ci)
    out = np.full((N, ref.shape[2], ref.shape[3], Co), np.nan)
    replay(case, x.transpose(0, 2, 3, 1), wp, out)
    assert not np.isnan(out).any(), 'unwritten outputs'
    np.testing.assert_array_equal(out.transpose(0, 3, 1, 2), ref)


def test_halo_plan_fits_two_ctas_per_sm_for_the_headline_layers():
    for case in CASES[:5]:
        pl = plan_of(case, 0)
        if pl['persistent']:      # one persistent CTA per SM: two windows + two accumulator sets
            assert pl['smem'] <= 227 * 1024 and pl['tmem_cols'] <= 512, pl
        else:
            assert pl['smem'] <= 112 * 1024 and pl['m'] * pl['BN'] <= 256, pl


@pytest.mark.skipif(os.environ.get('G2_HALO_PERSISTENT') in ('0', '1'), reason='already inside a fixed-mode child run')
@pytest.mark.parametrize('mode', ['0', '1'])
def test_halo_plan_replay_in_persistent_mode(mode):
    """(the default, G2_HALO_PERSISTENT unset, mixes the two kernels by a heuristic; the child runs pin one of them)
    The persistent kernel picks its tile geometry under a different budget (two windows, two accumulator
    sets, resident weights).  G2_HALO_PERSISTENT is read once per process, so the same replay runs in a child process with
    the switch on: g2_conv_halo_plan then reports the persistent geometry and the numpy replay must still be exact."""
    env = dict(os.environ, G2_HALO_PERSISTENT=mode)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-q', '-x', '-k', 'replay_matches_torch'],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
