"""CPU: the oracle (oracle/models.py) reproduces the golden vectors generated from the REAL reference
(oracle/make_golden.py), with parameters re-created by the engine's holders from the same seed."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import functional as O
from oracle import models as M

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', '*.npz'))
              if not os.path.basename(p).startswith(('sample_', 'eval_', 'vae_', 'variant_')))   # sample_* / eval_*: tests/test_sample_golden.py, test_eval_golden.py


def load_plugin(model):
    import importlib
    return importlib.import_module('genesis_b200.model_configs.%s_config' % model)


def golden_case(path):
    g = np.load(path)
    model, K, img, B, gen = g['meta']
    return g, model, int(K), int(img), int(B)


def build_engine_model(model, K, img, seed=0, **over):
    cfg = M.make_cfg(model, K_steps=K, img_size=img, **over)
    torch.manual_seed(seed)
    return load_plugin(model).load(cfg), cfg


def tape_from_golden(g):
    rec = [(str(k), torch.from_numpy(g['noise_%d' % i])) for i, k in enumerate(g['noise_kinds'])]
    return O.NoiseTape(record=rec)


def direction(n, idx):
    i = torch.arange(n, dtype=torch.float64)
    return torch.cos(0.37 * i + 1.3 * idx + 0.1)


@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_engine_init_matches_reference_checksums(path):
    g, model, K, img, B = golden_case(path)
    try:
        m, _ = build_engine_model(model, K, img)
    except ModuleNotFoundError:
        pytest.skip('plug-in for %s not built yet' % model)
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(g['param_names'])
    for n, (s, a) in zip(g['param_names'], g['param_sums']):
        t = sd[str(n)].double()
        assert abs(t.sum().item() - s) <= 1e-9 * max(1.0, abs(s)), n
        assert abs(t.abs().sum().item() - a) <= 1e-9 * max(1.0, a), n


@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_golden(path):
    g, model, K, img, B = golden_case(path)
    if img > 64:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    try:
        m, cfg = build_engine_model(model, K, img)
    except ModuleNotFoundError:
        pytest.skip('plug-in for %s not built yet' % model)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone())
         for k, v in m.state_dict().items()}
    x = torch.from_numpy(g['x'])
    out = M.FORWARD[model](P, x, tape_from_golden(g), cfg, training=True)
    np.testing.assert_allclose(out['err'].detach().numpy(), g['err'], rtol=2e-6)
    np.testing.assert_allclose(out['recon'].detach().numpy(), g['recon'], atol=2e-6)
    np.testing.assert_allclose(torch.stack(out['log_m_k'], 0).detach().numpy(), g['log_m_k'], atol=2e-4, rtol=1e-5)
    for key in ('kl_l_k', 'kl_m_k'):
        if key in g.files:
            np.testing.assert_allclose(torch.stack(out[key], 0).detach().numpy(), g[key], atol=2e-4, rtol=1e-5)
    if 'kl_m' in g.files:
        np.testing.assert_allclose(out['kl_m'].detach().numpy(), g['kl_m'], rtol=1e-5)
    if 'log_m_r_k' in g.files:
        np.testing.assert_allclose(torch.stack(out['log_m_r_k'], 0).detach().numpy(), g['log_m_r_k'], atol=2e-5)
    M.total_loss(out).backward()
    gmax = max(float(s[0]) for s in g['grad_sums'])
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        gr = P[str(n)].grad
        if gr is None:
            assert nrm == 0.0, n
            continue
        gd = gr.double().flatten()
        tol = 2e-4 * nrm + 1e-6 * gmax + 1e-7
        if model == 'genesisv2' and str(n).startswith('encoder.down'):
            # One ReLU in encoder.down.1 sits within rounding of zero for the 'rooms' batch: the reference (which
            # evaluates feat_head K times) and the oracle land on different sides, which moves ~30 activation gradients
            # in down.0/down.1 by ~1e-4 absolute.  The fp64 oracle agrees with the fp32 oracle to 1e-6 on these tensors.
            tol += 5e-3 * nrm
        assert abs(gd.norm().item() - nrm) <= tol, (n, gd.norm().item(), nrm)
        assert abs((gd * direction(gd.numel(), i)).sum().item() - proj) <= tol * 4, (n, proj)
    if 'bn_names' in g.files:
        for n, (s, a) in zip(g['bn_names'], g['bn_sums']):
            t = out['bn_updates'][str(n)].double()
            assert abs(t.sum().item() - s) <= 1e-5 * max(1.0, abs(s)), n
            assert abs(t.abs().sum().item() - a) <= 1e-5 * max(1.0, a), n
