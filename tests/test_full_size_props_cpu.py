"""CPU: the size-independent property checks of tests/full_size_props.py, exercised at small sizes on the plug-ins' host logic
(kernel layer replaced by its torch contract, tests/cpu_ops_mock.py).  The same checks run at BASELINE.json's full sizes on the
GPU; running them here validates the checkers themselves and pins the identities for the host logic."""
import pytest
import torch

import cpu_ops_mock
import full_size_props as FP
from oracle import synth
from test_oracle_golden import build_engine_model

STD = {'genesis': lambda m: m.std.reshape(-1), 'genesisv2': lambda m: float(m.std), 'monet': lambda m: m.std.reshape(-1)}


def forward(m, x, tape):
    m.set_noise_tape(tape)
    with torch.no_grad():
        out = m(x.as_subclass(cpu_ops_mock.AsCuda))
    m.set_noise_tape(None)
    return out


@pytest.mark.parametrize('model,K,gen', [('genesis', 3, 'multid'), ('genesisv2', 4, 'stacks'), ('monet', 3, 'rooms')])
def test_identities_and_subset_invariance(monkeypatch, model, K, gen):
    from genesis_b200 import ops
    cpu_ops_mock.install(monkeypatch, ops)
    B, n = 4, 2
    m, cfg = build_engine_model(model, K, 64)
    x = torch.from_numpy(synth.GENERATORS[gen](B, 64, 11)[0])
    if model == 'genesis':          # BatchNorm: move the running statistics once, then compare in eval mode
        m.train()
        forward(m, x, FP.SubsetTape(1, B, B, K))
        m.eval()
    else:                           # per-sample norms: training mode is batch-partition invariant
        m.train()
    full = forward(m, x, FP.SubsetTape(5, B, B, K))
    FP.check_identities(model, x, full, STD[model](m))
    sub = forward(m, x[:n], FP.SubsetTape(5, B, n, K))
    FP.check_subset_invariance(full, sub, n, rtol=1e-5, atol=1e-5)
