"""CPU: the closed-form backward that icsbp_bwd_kernel<KT> (csrc/v2.cu) evaluates -- reverse scan over the K-1 steps,
straight-through clamp, per-kernel d(alpha)/d(dist) and d(alpha)/d(log sigma), seed gradient scattered to the seed pixel --
restated in float64 torch with the kernel's own formulas and compared with autograd through the oracle's InstanceColouringSBP
(oracle/models.py:icsbp, pinned to the reference by the variant goldens) for the gaussian, laplacian and epanechnikov kernels.
The CUDA kernels themselves are compared with the same oracle on the GPU."""
import pytest
import torch

from oracle import models as M


def alpha_of(dist, inv_sigma, kt):
    if kt == 0:
        return torch.exp(-dist * inv_sigma)
    if kt == 1:
        return torch.exp(-dist.clamp(1e-10, 1e10).sqrt() * inv_sigma)
    return (1 - dist * inv_sigma).clamp_min(0)


def closed_form_backward(colour, log_sigma, seed_idx, dlog_m, kt):
    """colour [B,P,CD], seed_idx [K-1,B], dlog_m [K,B,P] -> dcolour [B,P,CD], dlog_sigma (scalar); mirrors icsbp_bwd_kernel."""
    K = dlog_m.shape[0]
    B, P, CD = colour.shape
    inv_sigma = 1.0 / log_sigma.exp()
    dcol = torch.zeros_like(colour)
    dls = colour.new_zeros(())
    R = dlog_m[K - 1].clone()
    ar = torch.arange(B)
    for k in range(K - 2, -1, -1):
        seed = colour[ar, seed_idx[k]]                      # [B,CD]
        df = colour - seed[:, None, :]
        dist = (df * df).sum(2)
        a = alpha_of(dist, inv_sigma, kt)
        ac = a.clamp(0.01, 0.99)
        G = dlog_m[k]
        da = G / ac - R / (1 - ac)
        R = R + G
        if kt == 0:
            dd = -da * a * inv_sigma
            dls = dls + (da * a * dist * inv_sigma).sum()
        elif kt == 1:
            d = dist.clamp(1e-10, 1e10).sqrt()
            dd = -da * a * inv_sigma * (0.5 / d)
            dls = dls + (da * a * d * inv_sigma).sum()
        else:
            on = (1 - dist * inv_sigma) > 0
            dd = torch.where(on, -da * inv_sigma, torch.zeros_like(da))
            dls = dls + torch.where(on, da * dist * inv_sigma, torch.zeros_like(da)).sum()
        g = 2 * df * dd[:, :, None]
        dcol = dcol + g
        dcol[ar, seed_idx[k]] -= g.sum(1)
    return dcol, dls


@pytest.mark.parametrize('kernel,kt', [('gaussian', 0), ('laplacian', 1), ('epanechnikov', 2)])
@pytest.mark.parametrize('K', [2, 5])
def test_closed_form_matches_autograd(kernel, kt, K):
    torch.manual_seed(kt * 10 + K)
    B, H, W, CD = 3, 12, 12, 8
    # colours spread so that alpha covers both clamp ends and, for epanechnikov, both sides of the relu
    scale = torch.tensor([0.03, 0.3, 1.5], dtype=torch.float64)[torch.randint(0, 3, (B, 1, H, W))]
    colour = (scale * torch.randn(B, CD, H, W, dtype=torch.float64)).requires_grad_(True)
    u = torch.rand(B, 1, H, W, dtype=torch.float64)
    sigma0 = {0: 1.0 / (K * 0.6931), 1: 1.0 / (K ** 0.5 * 0.6931), 2: 2.0 / K}[kt]
    log_sigma = torch.tensor(sigma0, dtype=torch.float64).log().requires_grad_(True)
    log_m_k, log_s_k, seeds, idxs = M.icsbp(colour, u, log_sigma, K - 1, kernel)
    log_m = torch.stack(log_m_k, 0)                                     # [K,B,1,H,W]
    dlog_m = torch.randn_like(log_m)
    gc, gs = torch.autograd.grad((log_m * dlog_m).sum(), [colour, log_sigma])
    col_nhwc = colour.detach().permute(0, 2, 3, 1).reshape(B, H * W, CD)
    dcol, dls = closed_form_backward(col_nhwc, log_sigma.detach(), torch.stack(idxs, 0), dlog_m.reshape(K, B, H * W), kt)
    ref = gc.permute(0, 2, 3, 1).reshape(B, H * W, CD)
    a = alpha_of(((col_nhwc - col_nhwc[torch.arange(B), idxs[0]][:, None]) ** 2).sum(2), 1 / log_sigma.detach().exp(), kt)
    assert (a < 0.01).any() and ((a > 0.05) & (a < 0.95)).any()                         # the draw exercises the clamp and the open range
    torch.testing.assert_close(dcol, ref, rtol=1e-9, atol=1e-10)
    torch.testing.assert_close(dls, gs, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize('kernel,kt', [('gaussian', 0), ('epanechnikov', 2)])
def test_closed_form_with_dynamic_K(kernel, kt):
    """dynamic_K: an image whose loop stopped after n masks gets gradient only through those n (the backward kernel runs its
    reverse scan with K = n_masks[b]; padded slots are constants)."""
    torch.manual_seed(5 + kt)
    H, W, CD, K = 16, 16, 8, 7
    colour = (0.3 * torch.randn(1, CD, H, W, dtype=torch.float64)).requires_grad_(True)
    u = torch.rand(1, 1, H, W, dtype=torch.float64)
    log_sigma = torch.tensor(3.0, dtype=torch.float64).log().requires_grad_(True)      # wide kernel: the scope empties quickly
    log_m_k, log_s_k, seeds, idxs = M.icsbp(colour, u, log_sigma, K - 1, kernel, dynamic_K=True)
    n = len(log_m_k)
    assert 1 < n < K                                                    # stopped early, with at least one real step
    log_m = torch.stack(log_m_k, 0)
    dlog_m = torch.randn_like(log_m)
    gc, gs = torch.autograd.grad((log_m * dlog_m).sum(), [colour, log_sigma])
    col_nhwc = colour.detach().permute(0, 2, 3, 1).reshape(1, H * W, CD)
    dcol, dls = closed_form_backward(col_nhwc, log_sigma.detach(), torch.stack(idxs, 0)[:n - 1], dlog_m.reshape(n, 1, H * W), kt)
    torch.testing.assert_close(dcol, gc.permute(0, 2, 3, 1).reshape(1, H * W, CD), rtol=1e-9, atol=1e-10)
    torch.testing.assert_close(dls, gs, rtol=1e-9, atol=1e-10)
