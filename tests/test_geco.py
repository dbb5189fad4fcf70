"""The device-resident GECO of trainer.GecoState (one scalar kernel, g2_geco_step_f32: no .item(), CUDA-graph replayable) follows
the REAL reference class
utils/geco.py:19-51 step for step: same loss weight, EMA and beta trajectory, including the
first-step branch, the speed-up branch (constraint > 0) and the clamp.  Skipped where /root/reference is absent."""
import os
import sys

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='neither the reference checkout nor oracle/_ref is present')


@pytest.fixture(params=['emu', pytest.param('cuda', marks=pytest.mark.gpu)])
def device(request, monkeypatch):
    """'emu': the kernel source compiled for the CPU (tests/cuda_emu) behind genesis_b200._lib -- the default CPU suite;
    'cuda': the product library on the GPU."""
    if request.param == 'emu':
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cuda_emu'))
        import emu_lib
        emu_lib.install(monkeypatch)
        return 'cpu'
    return 'cuda'


@pytest.mark.parametrize('speedup', [10.0, None])
def test_geco_trajectory_equals_reference(speedup, device):
    ref_loader._setup()
    from utils.geco import GECO
    from genesis_b200 import trainer
    goal, lr = 0.5655 * 3 * 64 ** 2, 1e-5
    ref = GECO(goal, lr, alpha=0.99, beta_init=1.0, beta_min=1e-10, speedup=speedup)
    eng = trainer.GecoState(goal, lr, device, alpha=0.99, beta_init=1.0, beta_min=1e-10, speedup=speedup)
    g = torch.Generator().manual_seed(0)
    # err wanders across the goal so both rate branches and both signs of the constraint occur; one huge excursion hits the clamp
    errs = [goal * (1.0 + 0.4 * torch.randn((), generator=g).item()) for _ in range(60)] + [goal * 40.0] * 3 + [goal * 0.1] * 40
    for i, e in enumerate(errs):
        err, kl = torch.tensor(e), torch.tensor(123.4)
        beta_used = float(eng.beta)                 # the step's loss uses beta BEFORE the update, as geco.loss does
        loss_ref = ref.loss(err, kl)
        eng.update(err.to(device))
        assert float(loss_ref) == pytest.approx(e + beta_used * 123.4, rel=1e-6), i
        assert float(eng.err_ema) == pytest.approx(float(ref.err_ema), rel=1e-6), i
        assert float(eng.beta) == pytest.approx(float(ref.beta), rel=2e-5), i
    assert float(eng.beta) > 0
