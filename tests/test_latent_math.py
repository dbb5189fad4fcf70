"""CPU: the closed-form forward / backward formulas that genesis_b200/csrc/latent.cu implements (LSTM cell, Gaussian head,
Monte-Carlo KL), transliterated to torch line by line, against torch autograd on the reference expressions
(torch.nn.LSTM cell arithmetic; blocks.to_sigma + rsample; genesis_config.py:328-336).  This pins the derivations; the CUDA
kernels themselves are compared with autograd on the GPU (tests/test_pending_next_round.py until validated)."""
import math

import torch

torch.manual_seed(0)
DT = torch.float64


def lstm_cell_fwd(gx, gh, c_prev):
    H = gx.shape[1] // 4
    gi, gf, gg, go = (gx + gh).split(H, dim=1)
    i, f, g, o = torch.sigmoid(gi), torch.sigmoid(gf), torch.tanh(gg), torch.sigmoid(go)
    c = (f * c_prev if c_prev is not None else 0) + i * g
    return o * torch.tanh(c), c


def lstm_cell_bwd(gx, gh, c_prev, c, dh, dc):
    H = gx.shape[1] // 4
    gi, gf, gg, go = (gx + gh).split(H, dim=1)
    i, f, g, o = torch.sigmoid(gi), torch.sigmoid(gf), torch.tanh(gg), torch.sigmoid(go)
    tc = torch.tanh(c)
    dcv = dc + dh * o * (1 - tc * tc)
    cp = c_prev if c_prev is not None else torch.zeros_like(c)
    dgates = torch.cat([dcv * g * i * (1 - i), dcv * cp * f * (1 - f), dcv * i * (1 - g * g), dh * tc * o * (1 - o)], dim=1)
    return dgates, dcv * f


def test_lstm_cell_formulas():
    B, H = 5, 12
    for with_state in (True, False):
        gx = torch.randn(B, 4 * H, dtype=DT, requires_grad=True)
        gh = torch.randn(B, 4 * H, dtype=DT, requires_grad=True)
        cp = torch.randn(B, H, dtype=DT, requires_grad=True) if with_state else None
        h, c = lstm_cell_fwd(gx, gh, cp)
        # reference arithmetic: torch's LSTM cell
        cell = torch.nn.LSTMCell(3, H).double()
        with torch.no_grad():
            cell.weight_ih.zero_(); cell.weight_hh.zero_(); cell.bias_hh.zero_(); cell.bias_ih.zero_()
        gates = (gx + gh).detach()
        i, f, g, o = gates.split(H, 1)
        c_ref = (torch.sigmoid(f) * cp.detach() if with_state else 0) + torch.sigmoid(i) * torch.tanh(g)
        h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
        assert torch.allclose(h, h_ref) and torch.allclose(c, c_ref)
        dh, dc = torch.randn(B, H, dtype=DT), torch.randn(B, H, dtype=DT)
        ins = [gx, gh] + ([cp] if with_state else [])
        grads = torch.autograd.grad((h * dh).sum() + (c * dc).sum(), ins)
        dgates, dcp = lstm_cell_bwd(gx.detach(), gh.detach(), cp.detach() if with_state else None, c.detach(), dh, dc)
        assert torch.allclose(grads[0], dgates, atol=1e-12) and torch.allclose(grads[1], dgates, atol=1e-12)
        if with_state:
            assert torch.allclose(grads[2], dcp, atol=1e-12)


def test_gauss_head_formulas():
    n = 64
    mu = torch.randn(n, dtype=DT, requires_grad=True)
    raw = (3 * torch.randn(n, dtype=DT)).requires_grad_(True)
    eps = torch.randn(n, dtype=DT)
    sigma = torch.nn.functional.softplus(raw + 0.5) + 1e-8          # blocks.to_sigma
    z = mu + sigma * eps
    dz, ds = torch.randn(n, dtype=DT), torch.randn(n, dtype=DT)
    gmu, graw = torch.autograd.grad((z * dz).sum() + (sigma * ds).sum(), [mu, raw])
    assert torch.allclose(gmu, dz)
    assert torch.allclose(graw, (dz * eps + ds) * torch.sigmoid(raw.detach() + 0.5), atol=1e-12)


def normal_log_prob(z, mu, sigma):
    return -((z - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - 0.5 * math.log(2 * math.pi)


def test_mc_kl_formulas():
    B, D = 4, 16
    for with_prior in (True, False):
        z, mu, pmu = (torch.randn(B, D, dtype=DT, requires_grad=True) for _ in range(3))
        sigma = (torch.rand(B, D, dtype=DT) + 0.2).requires_grad_(True)
        psigma = (torch.rand(B, D, dtype=DT) + 0.2).requires_grad_(True)
        lq = normal_log_prob(z, mu, sigma).sum(1)
        lp = normal_log_prob(z, pmu, psigma).sum(1) if with_prior else (-0.5 * z ** 2 - 0.5 * math.log(2 * math.pi)).sum(1)
        kl = lq - lp
        # forward as the kernel computes it (the log 2 pi terms cancel)
        t = (z - mu) / sigma
        kq = -0.5 * t * t - torch.log(sigma)
        kp = (-0.5 * ((z - pmu) / psigma) ** 2 - torch.log(psigma)) if with_prior else -0.5 * z * z
        assert torch.allclose(kl, (kq - kp).sum(1), atol=1e-12)
        dkl = torch.randn(B, dtype=DT)
        ins = [z, mu, sigma] + ([pmu, psigma] if with_prior else [])
        grads = torch.autograd.grad((kl * dkl).sum(), ins)
        g = dkl.view(B, 1)
        r, inv = (z - mu).detach(), 1 / sigma.detach() ** 2
        gz = -r * inv
        dmu = g * r * inv
        dsg = g * (r * r * inv / sigma.detach() - 1 / sigma.detach())
        if with_prior:
            pr, pinv = (z - pmu).detach(), 1 / psigma.detach() ** 2
            gz = gz + pr * pinv
            assert torch.allclose(grads[3], -g * pr * pinv, atol=1e-12)
            assert torch.allclose(grads[4], -g * (pr * pr * pinv / psigma.detach() - 1 / psigma.detach()), atol=1e-12)
        else:
            gz = gz + z.detach()
        assert torch.allclose(grads[0], g * gz, atol=1e-12)
        assert torch.allclose(grads[1], dmu, atol=1e-12) and torch.allclose(grads[2], dsg, atol=1e-12)
