"""Evaluation-mode forward (train.py:479-546: model.eval(), no_grad) against the REAL reference: golden vectors from
oracle/make_golden.py --evals (one training-mode forward first, so the BatchNorm running statistics have moved, then the
eval forward on a second batch with a recorded noise tape).  CPU: the oracle restatement reproduces them."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import functional as O
from oracle import models as M
from oracle import synth

from test_oracle_golden import build_engine_model, tape_from_golden
from test_sample_golden import CASES

HERE = os.path.dirname(os.path.abspath(__file__))
EVALS = sorted(glob.glob(os.path.join(HERE, 'golden', 'eval_*.npz')))


@pytest.mark.parametrize('path', EVALS, ids=[os.path.basename(p)[:-4] for p in EVALS])
def test_oracle_eval_forward_matches_reference(path):
    g = np.load(path, allow_pickle=False)
    model, K, img, B, fwd, gen = (str(v) for v in g['meta'])
    K, img, B = int(K), int(img), int(B)
    m, cfg = build_engine_model(model, K, img)
    P = {k: v.clone() for k, v in m.state_dict().items()}
    x1 = torch.from_numpy(synth.GENERATORS[gen](B, img, 1)[0])
    with torch.no_grad():
        out = M.FORWARD[model](P, x1, O.NoiseTape(seed=2), cfg, training=True)
        P.update(out.get('bn_updates', {}))
        ev = M.FORWARD[model](P, torch.from_numpy(g['x']), tape_from_golden(g), cfg, training=False)
    np.testing.assert_allclose(ev['err'].numpy(), g['err'], rtol=3e-6)
    np.testing.assert_allclose(ev['recon'].numpy(), g['recon'], atol=3e-6)
    np.testing.assert_allclose(torch.stack(ev['log_m_k'], 0).numpy(), g['log_m_k'], atol=2e-4, rtol=1e-5)
    for key in ('kl_l_k', 'kl_m_k'):
        if key in g.files:
            np.testing.assert_allclose(torch.stack(ev[key], 0).numpy(), g[key], atol=2e-4, rtol=1e-5)
    if 'kl_m' in g.files:
        np.testing.assert_allclose(ev['kl_m'].numpy(), g['kl_m'], rtol=1e-5)
    assert fwd in CASES
