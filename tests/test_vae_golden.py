"""BaselineVAE (BASELINE.json configs[0]) against the REAL reference: golden vectors from
oracle/make_golden.py --vae.  CPU: the engine's holder re-creates the reference's seeded parameters bit for bit (names and
checksums) and the oracle restatement reproduces outputs and gradient summaries."""
import os

import numpy as np
import torch

from oracle import models as M
from test_oracle_golden import build_engine_model, direction, tape_from_golden

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(os.path.join(HERE, 'golden', 'vae_b4.npz'), {}), (os.path.join(HERE, 'golden', 'vae_b2_broadcast.npz'), {'broadcast_decoder': True})]
PATH = CASES[0][0]


@pytest.mark.parametrize('path,over', CASES, ids=['deconv', 'broadcast'])
def test_vae_init_matches_reference_checksums(path, over):
    g = np.load(path)
    m, _ = build_engine_model('vae', 1, 64, **over)
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(g['param_names'])
    for n, (s, a) in zip(g['param_names'], g['param_sums']):
        t = sd[str(n)].double()
        assert abs(t.sum().item() - s) <= 1e-9 * max(1.0, abs(s)), n
        assert abs(t.abs().sum().item() - a) <= 1e-9 * max(1.0, a), n


@pytest.mark.parametrize('path,over', CASES, ids=['deconv', 'broadcast'])
def test_vae_oracle_matches_golden(path, over):
    g = np.load(path)
    m, cfg = build_engine_model('vae', 1, 64, **over)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in m.state_dict().items()}
    out = M.FORWARD['vae'](P, torch.from_numpy(g['x']), tape_from_golden(g), cfg, training=True)
    np.testing.assert_allclose(out['err'].detach().numpy(), g['err'], rtol=2e-6)
    np.testing.assert_allclose(out['kl_l'].detach().numpy(), g['kl_l'], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(out['recon'].detach().numpy(), g['recon'], atol=2e-6)
    np.testing.assert_allclose(out['z'].detach().numpy(), g['z'], atol=2e-6)
    M.total_loss(out).backward()
    gmax = max(float(s[0]) for s in g['grad_sums'])
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        gd = P[str(n)].grad.double().flatten()
        tol = 2e-4 * nrm + 1e-6 * gmax + 1e-7
        assert abs(gd.norm().item() - nrm) <= tol, (n, gd.norm().item(), nrm)
        assert abs((gd * direction(gd.numel(), i)).sum().item() - proj) <= 4 * tol, (n, proj)
