"""CPU: the Python glue of genesis_b200/holders.py around the kernels -- LSTM stepping, the autoregressive prior, the Gaussian /
prior heads and the MC-KL -- with `ops.linear` replaced by torch's F.linear, against torch.nn.LSTM and the oracle.  (The
kernels themselves are checked on the GPU; this pins the host-side control flow, which is identical on both devices.)"""
import pytest
import torch
import torch.nn.functional as F

from genesis_b200 import holders as H
from genesis_b200 import ops
from oracle import functional as O


@pytest.fixture
def torch_linear(monkeypatch):
    def linear(x, w, b=None, act=None):
        y = F.linear(x, w, b)
        return {None: y, 'relu': F.relu(y), 'elu': F.elu(y)}[act] if act in (None, 'relu', 'elu') else y
    monkeypatch.setattr(ops, 'linear', linear)
    monkeypatch.setattr(ops, 'fused_latent', lambda: False)      # the ATen formulation = the contract of the fused latent kernels


def test_lstm_step_equals_nn_lstm(torch_linear):
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(40, 24)
    xs = torch.randn(3, 5, 40)
    ref, _ = lstm(xs)
    state, outs = None, []
    for t in range(3):
        h, state = H.lstm_step(xs[t], state, lstm)
        outs.append(h)
    assert torch.allclose(torch.stack(outs), ref, atol=1e-6)
    torch.stack(outs).sum().backward()
    g_hh = lstm.bias_hh_l0.grad.clone()
    lstm.zero_grad()
    lstm(xs)[0].sum().backward()
    assert torch.allclose(g_hh, lstm.bias_hh_l0.grad, atol=1e-5)           # every use of bias_hh goes through the op


def test_autoreg_prior_and_heads_equal_oracle(torch_linear):
    torch.manual_seed(1)
    lstm, lin = torch.nn.LSTM(16, 32), torch.nn.Linear(32, 32)
    z_k = [torch.randn(4, 16) for _ in range(4)]
    pmu, psig = H.autoreg_prior(z_k, lstm, lin)
    P = {'prior_lstm.' + k: v for k, v in lstm.state_dict().items()}
    P.update({'prior_linear.' + k: v for k, v in lin.state_dict().items()})
    omu, osig = O.autoreg_prior(z_k, P)
    assert len(pmu) == 3
    for k in range(3):          # the oracle returns K entries with None first
        assert torch.allclose(pmu[k], omu[k + 1], atol=1e-6) and torch.allclose(psig[k], osig[k + 1], atol=1e-6)
    lo, eps = torch.randn(4, 32), torch.randn(4, 16)
    z, mu, sigma = H.gauss_head(lo, eps)
    assert torch.equal(mu, lo[:, :16]) and torch.allclose(sigma, O.to_sigma(lo[:, 16:])) and torch.allclose(z, mu + sigma * eps)
    kl = H.mc_kl(z, mu, sigma, pmu[0], psig[0])
    assert torch.allclose(kl, O.mc_kl(z, mu, sigma, pmu[0], psig[0]), atol=1e-5)
    assert torch.allclose(H.mc_kl(z, mu, sigma), O.mc_kl(z, mu, sigma), atol=1e-5)
