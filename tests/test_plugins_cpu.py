"""CPU: the HOST LOGIC of the plug-ins (genesis_b200/model_configs/*.py) and of holders.py -- everything between the Forge
`load(cfg)` boundary and the kernel calls -- with the kernel layer replaced by its plain-torch contract
(tests/cpu_ops_mock.py), against the oracle on identical parameters, inputs and noise: outputs, loss terms, BatchNorm buffer
updates and the gradient of every parameter (autograd runs through the stand-ins).  What this pins without a GPU: slot /
batch bookkeeping (k-major stacking), NHWC <-> NCHW edges, weight re-indexing of the full-map convs and MLPs, the first
broadcast-decoder layer split, noise draw order, KL / prior wiring, the K-th mask fix-up, sample().  The kernels are checked
against the same oracle on the GPU."""
import numpy as np
import pytest
import torch

import cpu_ops_mock
from oracle import functional as O
from oracle import models as M
from oracle import synth
from test_oracle_golden import build_engine_model


def run_plugin(monkeypatch, model, K, B, gen, seed=3, **over):
    from genesis_b200 import ops
    cpu_ops_mock.install(monkeypatch, ops)
    m, cfg = build_engine_model(model, K, 64, **over)
    m.train()
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.from_numpy(synth.GENERATORS[gen](B, 64, 5)[0])
    m.set_noise_tape(O.NoiseTape(seed=seed))
    out = m(x.as_subclass(cpu_ops_mock.AsCuda))
    m.set_noise_tape(None)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
    ref = M.FORWARD[model](P, x, O.NoiseTape(seed=seed), cfg, training=True)
    return m, out, P, ref


def check_grads(m, P, tol=2e-3):
    gmax = max(p.grad.norm().item() for p in P.values() if torch.is_tensor(p) and p.grad is not None)
    for n, p in m.named_parameters():
        ref = P[n].grad
        if ref is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, n
            continue
        assert p.grad is not None, n
        d = (p.grad.detach() - ref).norm().item()
        assert d <= tol * ref.norm().item() + 1e-5 * gmax, (n, d, ref.norm().item())


def stack(ts):
    return torch.stack([torch.as_tensor(t) for t in ts], 0).detach().numpy()


@pytest.mark.parametrize('K,B,gen', [(3, 3, 'multid'), (2, 2, 'rooms')])
def test_genesis_host_logic(monkeypatch, K, B, gen):
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'genesis', K, B, gen)
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(recon.detach().numpy(), ref['recon'].detach().numpy(), atol=1e-5)
    np.testing.assert_allclose(stack(stats['log_m_k']), stack(ref['log_m_k']), atol=1e-4)
    np.testing.assert_allclose(stack(losses['kl_m_k']), stack(ref['kl_m_k']), atol=1e-3, rtol=1e-4)
    np.testing.assert_allclose(stack(losses['kl_l_k']), stack(ref['kl_l_k']), atol=1e-3, rtol=1e-4)
    np.testing.assert_allclose(stack(comp['z_k']), stack(ref['comp']['z_k']), atol=1e-5)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P)
    for n, v in ref['bn_updates'].items():                      # BatchNorm running statistics moved as in the reference
        np.testing.assert_allclose(m.state_dict()[n].numpy(), v.numpy(), rtol=1e-4, atol=1e-6, err_msg=n)


def test_genesis_instance_norm_variant_host_logic(monkeypatch):
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'genesis', 3, 2, 'multid', enc_norm='in', dec_norm='in')
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(stack(stats['log_m_k']), stack(ref['log_m_k']), atol=1e-4)


def test_genesis_one_stage_variant_host_logic(monkeypatch):
    """two_stage=False: BroadcastDecoder on the mask latents, losses = {err, kl_m_k}, comp_stats None; also its sample()."""
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'genesis', 3, 2, 'multid', two_stage=False)
    assert comp is None and 'kl_l_k' not in losses
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(recon.detach().numpy(), ref['recon'].detach().numpy(), atol=1e-5)
    np.testing.assert_allclose(stack(losses['kl_m_k']), stack(ref['kl_m_k']), atol=1e-3, rtol=1e-4)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P)
    m.eval()
    m.set_noise_tape(O.NoiseTape(seed=6))
    img, st = m.sample(2, 3)
    m.set_noise_tape(None)
    with torch.no_grad():
        sref = M.SAMPLE['genesis']({k: v.detach() for k, v in m.state_dict().items()}, 2, O.NoiseTape(seed=6),
                                   M.make_cfg('genesis', K_steps=3, img_size=64, two_stage=False), training=False)
    np.testing.assert_allclose(img.numpy(), sref['image'].numpy(), atol=1e-5)


def test_monet_host_logic(monkeypatch):
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'monet', 3, 2, 'multid')
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(losses['kl_m'].detach().numpy(), ref['kl_m'].detach().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(recon.detach().numpy(), ref['recon'].detach().numpy(), atol=1e-5)
    np.testing.assert_allclose(stack(stats['log_m_k']), stack(ref['log_m_k']), atol=1e-4)
    np.testing.assert_allclose(stack(losses['kl_l_k']), stack(ref['kl_l_k']), atol=1e-3, rtol=1e-4)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P, tol=5e-3)


def test_genesisv2_host_logic(monkeypatch):
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'genesisv2', 4, 2, 'stacks')
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(recon.detach().numpy(), ref['recon'].detach().numpy(), atol=1e-5)
    np.testing.assert_allclose(stack(stats['log_m_k']), stack(ref['log_m_k']), atol=1e-4)
    np.testing.assert_allclose(stack(stats['log_m_r_k']), stack(ref['log_m_r_k']), atol=1e-4)
    np.testing.assert_allclose(stack(losses['kl_l_k']), stack(ref['kl_l_k']), atol=1e-3, rtol=1e-4)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P, tol=5e-3)


@pytest.mark.parametrize('over', [{}, {'broadcast_decoder': True}], ids=['deconv', 'broadcast'])
def test_vae_host_logic(monkeypatch, over):
    m, (recon, losses, stats, _, _), P, ref = run_plugin(monkeypatch, 'vae', 1, 3, 'multid', **over)
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(losses['kl_l'].detach().numpy(), ref['kl_l'].detach().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(recon.detach().numpy(), ref['recon'].detach().numpy(), atol=1e-5)
    (losses['err'].mean(0) + losses['kl_l'].mean(0)).backward()
    M.total_loss(ref).backward()
    check_grads(m, P)


@pytest.mark.parametrize('model,K', [('genesis', 4), ('genesisv2', 5), ('monet', 4)])
@pytest.mark.parametrize('training', [False, True])
def test_sample_host_logic(monkeypatch, model, K, training):
    """sample() (row a22) through the contract stand-ins vs the oracle's restatement; training=True exercises the per-slot
    BatchNorm decode of GENESIS (reference attention.py:61)."""
    from genesis_b200 import ops
    cpu_ops_mock.install(monkeypatch, ops)
    m, cfg = build_engine_model(model, K, 64)
    m.train(training)
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.set_noise_tape(O.NoiseTape(seed=5))
    img, stats = m.sample(3, K)
    m.set_noise_tape(None)
    with torch.no_grad():
        ref = M.SAMPLE[model](P, 3, O.NoiseTape(seed=5), cfg, training=training)
    np.testing.assert_allclose(img.numpy(), ref['image'].numpy(), atol=1e-5)
    np.testing.assert_allclose(stack(stats['log_m_k']), stack(ref['log_m_k']), atol=1e-4)
    np.testing.assert_allclose(stack(stats['x_k']), stack(ref['x_k']), atol=1e-5)


@pytest.mark.parametrize('model,K,gen,over', [('genesis', 3, 'rooms', dict(comp_prior=False)),
                                              ('genesisv2', 4, 'stacks', dict(autoreg_prior=False))])
def test_prior_flag_variants_host_logic(monkeypatch, model, K, gen, over):
    """comp_prior=False (standard-normal component prior) and GENESIS-V2 with autoreg_prior=False."""
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, model, K, 2, gen, **over)
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(stack(losses['kl_l_k']), stack(ref['kl_l_k']), atol=1e-3, rtol=1e-4)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P, tol=5e-3)


def test_genesis_comp_symmetric_variant_host_logic(monkeypatch):
    """comp_symmetric=True (reference genesis_config.py:101-120): gated conv component VAE with BatchNorm over the K*B slots."""
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'genesis', 3, 2, 'multid', comp_symmetric=True)
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(recon.detach().numpy(), ref['recon'].detach().numpy(), atol=1e-5)
    np.testing.assert_allclose(stack(losses['kl_l_k']), stack(ref['kl_l_k']), atol=1e-3, rtol=1e-4)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P)
    for n, v in ref['bn_updates'].items():
        np.testing.assert_allclose(m.state_dict()[n].numpy(), v.numpy(), rtol=1e-4, atol=1e-6, err_msg=n)


def test_genesis_comp_symmetric_sample_host_logic(monkeypatch):
    from genesis_b200 import ops
    cpu_ops_mock.install(monkeypatch, ops)
    m, cfg = build_engine_model('genesis', 3, 64, comp_symmetric=True)
    m.eval()
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.set_noise_tape(O.NoiseTape(seed=8))
    img, stats = m.sample(2, 3)
    m.set_noise_tape(None)
    with torch.no_grad():
        ref = M.SAMPLE['genesis'](P, 2, O.NoiseTape(seed=8), cfg, training=False)
    np.testing.assert_allclose(img.numpy(), ref['image'].numpy(), atol=1e-5)
    np.testing.assert_allclose(stack(stats['x_k']), stack(ref['x_k']), atol=1e-5)


@pytest.mark.parametrize('detach', [True, False])
def test_genesisv2_klm_loss_variant_host_logic(monkeypatch, detach):
    """klm_loss=True (reference genesisv2_config.py:171-176), with and without detach_mr_in_klm."""
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'genesisv2', 4, 2, 'stacks', klm_loss=True,
                                                              detach_mr_in_klm=detach)
    np.testing.assert_allclose(losses['kl_m'].detach().numpy(), ref['kl_m'].detach().numpy(), rtol=1e-4, atol=1e-3)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P, tol=5e-3)


def test_monet_scope_prior_variant_host_logic(monkeypatch):
    """prior_mode='scope' (reference monet_config.py:141-153), forward + backward and sample()."""
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'monet', 3, 2, 'multid', prior_mode='scope')
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(losses['kl_m'].detach().numpy(), ref['kl_m'].detach().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(stack(stats['log_m_r_k']), stack(ref['log_m_r_k']), atol=1e-4)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P, tol=5e-3)
    m.eval()
    m.set_noise_tape(O.NoiseTape(seed=4))
    img, st = m.sample(2, 3)
    m.set_noise_tape(None)
    with torch.no_grad():
        sref = M.SAMPLE['monet']({k: v.detach() for k, v in m.state_dict().items()}, 2, O.NoiseTape(seed=4),
                                 M.make_cfg('monet', K_steps=3, img_size=64, prior_mode='scope'), training=False)
    np.testing.assert_allclose(img.numpy(), sref['image'].numpy(), atol=1e-5)


@pytest.mark.parametrize('over', [dict(kernel='laplacian'), dict(kernel='epanechnikov'), dict(semiconv=False)],
                         ids=['laplacian', 'epanechnikov', 'nosemiconv'])
def test_genesisv2_attention_option_variants_host_logic(monkeypatch, over):
    """InstanceColouringSBP options (reference modules/attention.py:138-160, 195-205): kernel type and the plain 1x1 colour head."""
    m, (recon, losses, stats, att, comp), P, ref = run_plugin(monkeypatch, 'genesisv2', 4, 2, 'multid', **over)
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(stack(stats['log_m_k']), stack(ref['log_m_k']), atol=1e-4)
    np.testing.assert_allclose(att['colour'].detach().numpy(), ref['att']['colour'].detach().numpy(), atol=1e-5)
    assert (att['delta'] is None) == (ref['att']['delta'] is None)
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P, tol=5e-3)


@pytest.mark.parametrize('K,B,mult', [(6, 3, 16.0), (8, 1, 32.0)], ids=['batch-padded', 'single-truncated'])
def test_genesisv2_dynamic_K_host_logic(monkeypatch, K, B, mult):
    """dynamic_K (reference genesisv2_config.py:118-137): per-image uniform draws, early exit, -1e10 padding for batches, fewer
    slots for a single image.  log_sigma is widened so that the early exit is actually taken (as in the goldens)."""
    import math
    from genesis_b200 import ops
    cpu_ops_mock.install(monkeypatch, ops)
    m, cfg = build_engine_model('genesisv2', K, 64, dynamic_K=True)
    with torch.no_grad():
        m.att_process.log_sigma.add_(math.log(mult))
    m.train()
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.from_numpy(synth.GENERATORS['multid'](B, 64, 5)[0])
    m.set_noise_tape(O.NoiseTape(seed=3))
    recon, losses, stats, att, comp = m(x.as_subclass(cpu_ops_mock.AsCuda))
    m.set_noise_tape(None)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
    ref = M.FORWARD['genesisv2'](P, x, O.NoiseTape(seed=3), cfg, training=True)
    assert min(ref['n_masks']) < K                                    # the early exit is exercised
    assert len(stats['log_m_k']) == len(ref['log_m_k'])
    np.testing.assert_allclose(stack(stats['log_m_k']), stack(ref['log_m_k']), atol=1e-4)
    np.testing.assert_allclose(losses['err'].detach().numpy(), ref['err'].detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(stack(losses['kl_l_k']), stack(ref['kl_l_k']), atol=1e-3, rtol=1e-4)
    if B > 1:
        assert att is None and stats['log_s_k'] is None               # reference :121-122
    else:
        assert len(att['seeds']) == len(ref['att']['seeds']) and len(stats['log_s_k']) == len(ref['log_s_k'])
    import util_parity as U
    U.engine_total_loss(losses).backward()
    M.total_loss(ref).backward()
    check_grads(m, P, tol=5e-3)
