"""CPU: checkpoint interchange with the reference.  A checkpoint written by the REAL
reference in train.py's format (train.py:405-416: model_state_dict, optimiser_state_dict, beta, iter_idx, err_ema) loads
into the engine's plug-in with strict key matching, including the legacy-key pop of train.py:191-192, and the engine's
state_dict loads back into the reference.  Skipped where /root/reference is absent (the GPU box)."""
import io

import pytest
import torch

from oracle import models as M
from oracle import ref_loader
from test_oracle_golden import load_plugin

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='reference checkout not present')


@pytest.mark.parametrize('model,K,img', [('genesis', 5, 64), ('genesisv2', 7, 64), ('monet', 7, 64), ('monet', 3, 128)])
def test_reference_checkpoint_round_trip(model, K, img):
    cfg = M.make_cfg(model, K_steps=K, img_size=img)
    ref = ref_loader.load_reference(model, cfg, seed=123)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4)
    buf = io.BytesIO()
    torch.save({'model_state_dict': ref.state_dict(), 'optimiser_state_dict': opt.state_dict(), 'beta': torch.tensor(0.37),
                'iter_idx': 1000, 'err_ema': torch.tensor(6789.0)}, buf)
    buf.seek(0)
    ckpt = torch.load(buf, map_location='cpu')
    sd = ckpt['model_state_dict']
    sd['comp_vae.decoder_module.seq.0.pixel_coords.g_1'] = torch.zeros(1)      # legacy buffers of older checkpoints
    sd.pop('comp_vae.decoder_module.seq.0.pixel_coords.g_1', None)              # train.py:191-192
    sd.pop('comp_vae.decoder_module.seq.0.pixel_coords.g_2', None)
    torch.manual_seed(0)
    eng = load_plugin(model).load(cfg)                                            # different seed: every tensor must be replaced
    missing, unexpected = eng.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    for k, v in ref.state_dict().items():
        assert torch.equal(eng.state_dict()[k], v), k
    # and back: the engine's state_dict is a valid reference checkpoint
    ref2 = ref_loader.load_reference(model, cfg, seed=7)
    ref2.load_state_dict(eng.state_dict(), strict=True)
    for (k, a), (_, b) in zip(ref.state_dict().items(), ref2.state_dict().items()):
        assert torch.equal(a, b), k


def test_geco_state_round_trip(monkeypatch):
    """beta / err_ema of a reference checkpoint (train.py:197-203, 410-416) restore into the device-resident GECO (its one
    kernel runs through the CPU emulation of the source here)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cuda_emu'))
    import emu_lib
    emu_lib.install(monkeypatch)
    from genesis_b200 import trainer
    g = trainer.GecoState(goal=0.5655 * 3 * 64 ** 2, step_size=1e-5, device='cpu')
    st = {'beta': torch.tensor(0.37), 'err_ema': torch.tensor(6789.0)}
    g.load_state(st)
    out = g.state()
    assert float(out['beta']) == pytest.approx(0.37) and float(out['err_ema']) == pytest.approx(6789.0)
    g.update(torch.tensor(7000.0))      # an update after restore uses the restored EMA, not the first-step branch
    assert float(g.err_ema) == pytest.approx(0.01 * 7000.0 + 0.99 * 6789.0)
