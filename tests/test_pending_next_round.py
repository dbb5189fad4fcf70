"""Tests written at the end of round 1 AFTER the GPU budget was spent: they have not run on a B200 yet, so they are not
part of the `-m gpu` gate (a wrong expectation in an unvalidated test must not mask the validated suite).  Run them with

    G2_RUN_PENDING=1 python -m pytest tests/test_pending_next_round.py -q

on a GPU box; once green they move into the regular GPU files with the `gpu` marker."""
import os

import pytest
import torch

pytestmark = pytest.mark.skipif(os.environ.get('G2_RUN_PENDING') != '1' or not torch.cuda.is_available(),
                                reason='pending validation on a B200 (set G2_RUN_PENDING=1)')


def test_fused_adam_equals_torch_optim_adam():
    """g2_adam_f32 (flat arena, device step counter, grad scale, fused zero-grad) vs torch.optim.Adam (train.py:175, 263)."""
    from genesis_b200 import _lib
    torch.manual_seed(0)
    n = 4096 + 64
    p0 = torch.randn(n, device='cuda')
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    step = torch.zeros((), device='cuda')
    world = 2.0
    for it in range(5):
        g = torch.randn(n, device='cuda') * (10.0 ** (it - 2))
        ref.grad = g.clone()
        opt.step()
        gbuf = (g * world).clone()          # the arena holds the SUM over ranks; the kernel applies 1/world
        step += 1
        _lib.call('g2_adam_f32', p, gbuf, m, v, n, 1e-3, 0.9, 0.999, 1e-8, step, 1.0 / world, 1)
        torch.cuda.synchronize()
        assert gbuf.abs().max().item() == 0.0                         # zero-grad fused
        assert (p - ref.detach()).abs().max().item() <= 2e-6 * max(1.0, ref.detach().abs().max().item())


def test_lstm_first_step_matches_later_step_path():
    """holders.lstm_step with state=None (h_0 = 0 through the GEMM op) equals nn.LSTM's first step, and bias_hh / weight_hh
    receive their gradients through the op (weight_hh's is exactly zero)."""
    from genesis_b200 import holders as H
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(96, 128).cuda()
    x = torch.randn(8, 96, device='cuda')
    h, (h2, c) = H.lstm_step(x, None, lstm)
    out, (hn, cn) = lstm(x.view(1, 8, 96))
    assert (h - out[0]).abs().max().item() < 2e-3          # TF32 operands
    h.sum().backward()
    assert lstm.bias_hh_l0.grad is not None and lstm.weight_hh_l0.grad.abs().max().item() == 0.0
    assert (lstm.bias_hh_l0.grad - lstm.bias_ih_l0.grad).abs().max().item() < 1e-5
