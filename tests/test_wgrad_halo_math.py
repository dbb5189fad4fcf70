"""CPU: the flat "halo" formulation planned for the weight-gradient kernel, proven in numpy before
any CUDA is written.  For a stride-1 conv  y[n,oh,ow,b] = sum_tap x[n, oh+dh, ow+dw, a] w[tap][a][b]  the weight gradient is

    dW[tap][a][b] = sum_{n,oh,ow} x[n, oh+dh(tap), ow+dw(tap), a] * dy[n, oh, ow, b].

Per CTA window (TH output rows of one image): G = the zero-padded x window as a flat run of pixel rows with pitch
Wp = Wo + dw-span (exactly what the conv halo kernel's TMA box lands); T = the dy rows laid out with the SAME pitch, its
Wp - Wo pad columns ZERO (TMA out-of-bounds fill).  Then every tap is one matrix product over the flat positions,

    dW[tap] += G[toff(tap) : toff(tap) + F]^T @ T[0 : F],      toff = (dh - dh_min) * Wp + (dw - dw_min),  F = TH * Wp,

i.e. the K dimension of the MMA is the flat pixel index and a tap is a row shift of the resident window -- taps of one filter
row are 1 row (128 B) apart, which is what lets four of them form one M = 128 operand with LBO = 128 B."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


@pytest.mark.parametrize('H,W,R,pad,TH', [(12, 10, 3, 0, 4), (16, 16, 5, 2, 5), (9, 13, 3, 1, 3)])
def test_flat_window_weight_gradient(H, W, R, pad, TH):
    rng = np.random.RandomState(0)
    N, Ca, Cb = 2, 4, 3
    x = rng.randint(-3, 4, (N, Ca, H, W)).astype(np.float64)
    Ho, Wo = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    dy = rng.randint(-2, 3, (N, Cb, Ho, Wo)).astype(np.float64)
    w = torch.zeros(Cb, Ca, R, R, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(torch.from_numpy(x), w, None, padding=pad)
    (ref,) = torch.autograd.grad((y * torch.from_numpy(dy)).sum(), [w])         # [Cb, Ca, R, R]
    taps = [(r - pad, s - pad) for r in range(R) for s in range(R)]
    dh_min, dw_min = -pad, -pad
    span = R - 1
    Wp = Wo + span
    dW = np.zeros((R * R, Ca, Cb))
    xn, dyn = x.transpose(0, 2, 3, 1), dy.transpose(0, 2, 3, 1)                # NHWC
    for n in range(N):
        for h0 in range(0, Ho, TH):
            rows = min(TH, Ho - h0)
            # G: window rows [h0 + dh_min, h0 + rows - 1 + dh_max], cols [dw_min, Wo - 1 + dw_max], zero outside the image
            G = np.zeros(((rows + span) * Wp + span, Ca))                       # + span: the last tap's shift reads past the window end
            for r in range(rows + span):
                for c in range(Wp):
                    hh, ww = h0 + dh_min + r, dw_min + c
                    if 0 <= hh < H and 0 <= ww < W:
                        G[r * Wp + c] = xn[n, hh, ww]
            # T: dy rows with pitch Wp, pad columns zero
            T = np.zeros((rows * Wp, Cb))
            for r in range(rows):
                T[r * Wp:r * Wp + Wo] = dyn[n, h0 + r]
            Fl = rows * Wp
            for t, (dh, dw) in enumerate(taps):
                toff = (dh - dh_min) * Wp + (dw - dw_min)
                dW[t] += G[toff:toff + Fl].T @ T
    got = dW.reshape(R, R, Ca, Cb).transpose(3, 2, 0, 1)
    np.testing.assert_array_equal(got, ref.numpy())
