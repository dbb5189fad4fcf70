"""GPU: the parts of the engine beyond the three default models -- fused Adam vs torch.optim.Adam, the fused latent kernels vs
autograd, evaluation-mode forwards, the BaselineVAE plug-in, every non-default variant of SURVEY.md section 8 f.4 against the
goldens of the real reference, the IC-SBP kernel types, dynamic_K, the MN-major operand probe the halo weight-gradient kernel
relies on, and the size-independent identities + batch-partition invariance at BASELINE.json's full per-GPU sizes
(c2 B=64, c3 B=128, c5 B=64).  First validated on a B200 in round 2 (profiles/r02_pending_validation.txt)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


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


def test_fused_latent_ops_equal_autograd():
    """csrc/latent.cu through ops.lstm_cell / gauss_head / prior_head / mc_kl vs torch autograd on the ATen formulation."""
    from genesis_b200 import holders as H, ops
    torch.manual_seed(0)
    dev = 'cuda'

    def both(fn, inputs, n_out):
        outs = []
        for fused in (False, True):
            ops.set_fused_latent(fused)
            try:
                xs = [t.clone().requires_grad_(True) if t is not None else None for t in inputs]
                o = fn(*xs)
                o = o if isinstance(o, tuple) else (o,)
                w = [torch.randn_like(t) for t in o[:n_out]]
                torch.manual_seed(1)
                w = [torch.randn_like(t) for t in o[:n_out]]
                loss = sum((t * wt).sum() for t, wt in zip(o[:n_out], w))
                g = torch.autograd.grad(loss, [t for t in xs if t is not None])
                outs.append(([t.detach() for t in o[:n_out]], g))
            finally:
                ops.set_fused_latent(True)
        for a, b in zip(outs[0][0] + list(outs[0][1]), outs[1][0] + list(outs[1][1])):
            assert (a - b).abs().max().item() <= 2e-5 * max(1.0, a.abs().max().item())

    B, Hh, D = 6, 128, 64
    lstm = torch.nn.LSTM(96, Hh).to(dev)
    x = torch.randn(B, 96, device=dev)
    both(lambda xx: H.lstm_step(xx, None, lstm)[1], [x], 2)                                            # first step: (h, c)
    st = (torch.randn(B, Hh, device=dev), torch.randn(B, Hh, device=dev))
    both(lambda xx, h0, c0: H.lstm_step(xx, (h0, c0), lstm)[1], [x, st[0], st[1]], 2)
    lo, eps = torch.randn(B, 2 * D, device=dev) * 3, torch.randn(B, D, device=dev)
    both(lambda l: H.gauss_head(l, eps), [lo], 3)
    both(lambda l: H.prior_head(l), [lo], 2)
    z, mu, pmu = (torch.randn(B, D, device=dev) for _ in range(3))
    sg, psg = torch.rand(B, D, device=dev) + 0.2, torch.rand(B, D, device=dev) + 0.2
    both(lambda *a: H.mc_kl(*a), [z, mu, sg, pmu, psg], 1)
    both(lambda a, b, c: H.mc_kl(a, b, c), [z, mu, sg], 1)


@pytest.mark.parametrize('model,K,B,gen', [('genesis', 3, 4, 'multid'), ('genesisv2', 4, 3, 'stacks'), ('monet', 3, 2, 'multid')])
def test_fused_latent_model_equals_unfused(model, K, B, gen):
    """Whole-model forward + backward with ops.set_fused_latent(True) vs the validated ATen formulation."""
    import util_parity as U
    from genesis_b200 import ops
    from oracle import synth
    from test_oracle_golden import build_engine_model
    m, cfg = build_engine_model(model, K, 64)
    m = m.cuda().train()
    x = torch.from_numpy(synth.GENERATORS[gen](B, 64, 5)[0]).cuda()
    res = []
    for fused in (False, True):
        ops.set_fused_latent(fused)
        try:
            recon, losses, stats, att, comp = U.run_engine(m, x.cpu(), U.make_tape(11))
        finally:
            ops.set_fused_latent(True)
        res.append((losses['err'].detach().clone(), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}))
    assert (res[0][0] - res[1][0]).abs().max().item() <= 1e-4 * res[0][0].abs().max().item()
    gmax = max(g.norm().item() for g in res[0][1].values())
    for n, g in res[0][1].items():
        d = (res[1][1][n] - g).norm().item()
        assert d <= 2e-2 * g.norm().item() + 1e-3 * gmax, (n, d, g.norm().item())       # run-to-run TF32 noise is ~1e-3


@pytest.mark.skipif(os.environ.get('G2_EMU') == '1', reason='a probe of the HARDWARE; the CPU model only encodes the assumption it tests')
def test_mn_major_operand_with_row_shift_and_overlapping_atoms():
    """Feasibility probe for the halo layout of the weight-gradient kernel: an MN-major TF32 A operand (pixels are the K
    dimension, SWIZZLE_128B_BASE32B) that (1) starts at an arbitrary pixel row of a resident window and (2) takes its four
    32-channel M atoms from the SAME window shifted by one pixel each (LBO = 128 B: four filter taps of one row as one
    M = 128 operand).  If the swizzle is a function of the absolute shared-memory address, as it is for K-major operands
    (tests/test_umma_layouts_gpu.py), both hold."""
    import numpy as np
    from test_umma_layouts_gpu import N, desc, idesc, image, probe
    rng = np.random.RandomState(2)
    npix = 16                                              # K = 16 pixels = 2 MMAs of 8
    win = rng.randint(-3, 4, (npix + 16, 32)).astype(np.float32)          # one 32-channel window, more rows than one operand
    Bb = rng.randint(-3, 4, (N // 32, npix, 32)).astype(np.float32)
    b1 = np.concatenate([image(Bb[j], 8, 3) for j in range(N // 32)])
    refB = Bb.transpose(0, 2, 1).reshape(N, npix)                         # [n][pix]
    tileB = npix * 128
    a_img = image(win, 8, 3)                                              # rows swizzled by their ABSOLUTE index & 3
    for shift in (0, 1, 2, 5):
        # four atoms = the window shifted by 0, 1, 2, 3 pixels: D[j*32 + c][n] = sum_pix win[shift + j + pix][c] * B[n][pix]
        ref = np.concatenate([win[shift + j:shift + j + npix].T for j in range(4)], 0) @ refB.T
        got = probe(a_img, b1, desc(128, 512, 1), desc(tileB, 512, 1), idesc(N, 1, 1), npix // 8, 1024, 1024, a_off=shift * 128)
        np.testing.assert_allclose(got, ref, atol=1e-3, err_msg='shift %d' % shift)


@pytest.mark.parametrize('name', ['eval_genesis_k5', 'eval_genesisv2_k7', 'eval_monet_k7'])
def test_engine_eval_forward_matches_reference(name):
    """Engine model.eval() forward vs the reference's (tests/golden/eval_*.npz; oracle side: tests/test_eval_golden.py)."""
    import numpy as np
    from oracle import functional as O, synth
    from test_oracle_golden import build_engine_model, tape_from_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name + '.npz'), allow_pickle=False)
    model, K, img, B, fwd, gen = (str(v) for v in g['meta'])
    K, img, B = int(K), int(img), int(B)
    m, cfg = build_engine_model(model, K, img)
    m = m.cuda().train()
    m.set_noise_tape(O.NoiseTape(seed=2))
    with torch.no_grad():
        m(torch.from_numpy(synth.GENERATORS[gen](B, img, 1)[0]).cuda())
    m.eval()
    m.set_noise_tape(tape_from_golden(g))
    with torch.no_grad():
        recon, losses, stats, att, comp = m(torch.from_numpy(g['x']).cuda())
    m.set_noise_tape(None)
    np.testing.assert_allclose(losses['err'].cpu().numpy(), g['err'], rtol=1e-4)
    np.testing.assert_allclose(recon.cpu().numpy(), g['recon'], atol=3e-3)
    np.testing.assert_allclose(torch.stack(list(stats['log_m_k']), 0).cpu().numpy(), g['log_m_k'], atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize('name,over', [('vae_b4', {}), ('vae_b2_broadcast', {'broadcast_decoder': True})])
def test_vae_engine_matches_reference_golden(name, over):
    """BaselineVAE plug-in (genesis_b200/model_configs/vae_config.py) forward + backward vs tests/golden/vae_*.npz."""
    import numpy as np
    from test_oracle_golden import build_engine_model, direction, tape_from_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name + '.npz'))
    m, cfg = build_engine_model('vae', 1, 64, **over)
    m = m.cuda().train()
    m.set_noise_tape(tape_from_golden(g))
    recon, losses, stats, _, _ = m(torch.from_numpy(g['x']).cuda())
    (losses['err'].mean(0) + losses['kl_l'].mean(0)).backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(losses['err'].detach().cpu().numpy(), g['err'], rtol=1e-4)
    np.testing.assert_allclose(losses['kl_l'].detach().cpu().numpy(), g['kl_l'], rtol=1e-3, atol=5e-3)
    np.testing.assert_allclose(recon.detach().cpu().numpy(), g['recon'], atol=2e-3)
    gmax = max(float(s[0]) for s in g['grad_sums'])
    params = dict(m.named_parameters())
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        gd = params[str(n)].grad.detach().double().cpu().flatten()
        tol = 1e-2 * nrm + 1e-4 * gmax
        assert abs(gd.norm().item() - nrm) <= tol, (n, gd.norm().item(), nrm)
        assert abs((gd * direction(gd.numel(), i)).sum().item() - proj) <= 4 * tol, (n, proj)


def test_variant_genesis_instance_norm_engine_matches_reference():
    """GENESIS with enc_norm = dec_norm = 'in' (gated layers with InstanceNorm): engine vs tests/golden/variant_genesis_k3_in.npz."""
    import numpy as np
    import util_parity as U
    from test_oracle_golden import build_engine_model, direction, tape_from_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'variant_genesis_k3_in.npz'))
    m, cfg = build_engine_model('genesis', 3, 64, enc_norm='in', dec_norm='in')
    m = m.cuda().train()
    recon, losses, stats, att, comp = U.run_engine(m, torch.from_numpy(g['x']), tape_from_golden(g))
    np.testing.assert_allclose(losses['err'].detach().cpu().numpy(), g['err'], rtol=1e-4)
    np.testing.assert_allclose(recon.detach().cpu().numpy(), g['recon'], atol=2e-3)
    np.testing.assert_allclose(torch.stack(stats['log_m_k'], 0).detach().cpu().numpy(), g['log_m_k'], atol=5e-3, rtol=1e-3)
    gmax = max(float(s[0]) for s in g['grad_sums'])
    params = dict(m.named_parameters())
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        gd = params[str(n)].grad.detach().double().cpu().flatten()
        assert abs(gd.norm().item() - nrm) <= 2e-2 * nrm + 1e-4 * gmax, (n, gd.norm().item(), nrm)


def test_variant_genesis_one_stage_engine_matches_reference():
    """GENESIS with two_stage=False: engine vs tests/golden/variant_genesis_k3_onestage.npz."""
    import numpy as np
    import util_parity as U
    from test_oracle_golden import build_engine_model, tape_from_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'variant_genesis_k3_onestage.npz'))
    m, cfg = build_engine_model('genesis', 3, 64, two_stage=False)
    m = m.cuda().train()
    recon, losses, stats, att, comp = U.run_engine(m, torch.from_numpy(g['x']), tape_from_golden(g))
    assert comp is None
    np.testing.assert_allclose(losses['err'].detach().cpu().numpy(), g['err'], rtol=1e-4)
    np.testing.assert_allclose(recon.detach().cpu().numpy(), g['recon'], atol=2e-3)
    np.testing.assert_allclose(torch.stack(losses['kl_m_k'], 0).detach().cpu().numpy(), g['kl_m_k'], atol=5e-3, rtol=1e-3)
    gmax = max(float(s[0]) for s in g['grad_sums'])
    params = dict(m.named_parameters())
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        gd = params[str(n)].grad.detach().double().cpu().flatten()
        assert abs(gd.norm().item() - nrm) <= 2e-2 * nrm + 1e-4 * gmax, (n, gd.norm().item(), nrm)


def test_variant_genesis_comp_symmetric_engine_matches_reference():
    """GENESIS with comp_symmetric=True: engine vs tests/golden/variant_genesis_k3_symmetric.npz."""
    import numpy as np
    import util_parity as U
    from test_oracle_golden import build_engine_model, tape_from_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'variant_genesis_k3_symmetric.npz'))
    m, cfg = build_engine_model('genesis', 3, 64, comp_symmetric=True)
    m = m.cuda().train()
    recon, losses, stats, att, comp = U.run_engine(m, torch.from_numpy(g['x']), tape_from_golden(g))
    np.testing.assert_allclose(losses['err'].detach().cpu().numpy(), g['err'], rtol=1e-4)
    np.testing.assert_allclose(recon.detach().cpu().numpy(), g['recon'], atol=2e-3)
    np.testing.assert_allclose(torch.stack(losses['kl_l_k'], 0).detach().cpu().numpy(), g['kl_l_k'], atol=1e-2, rtol=1e-3)
    gmax = max(float(s[0]) for s in g['grad_sums'])
    params = dict(m.named_parameters())
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        gd = params[str(n)].grad.detach().double().cpu().flatten()
        assert abs(gd.norm().item() - nrm) <= 2e-2 * nrm + 1e-4 * gmax, (n, gd.norm().item(), nrm)


@pytest.mark.parametrize('name,over', [('variant_genesisv2_k4_klm', dict(klm_loss=True)),
                                       ('variant_genesisv2_k4_klm_nodetach', dict(klm_loss=True, detach_mr_in_klm=False)),
                                       ('variant_genesisv2_k4_noprior', dict(autoreg_prior=False)),
                                       ('variant_genesis_k3_nocompprior', dict(comp_prior=False)),
                                       ('variant_monet_k4_scope', dict(prior_mode='scope')),
                                       ('variant_genesisv2_k4_laplacian', dict(kernel='laplacian')),
                                       ('variant_genesisv2_k4_epanechnikov', dict(kernel='epanechnikov')),
                                       ('variant_genesisv2_k4_nosemiconv', dict(semiconv=False))])
def test_flag_variants_engine_matches_reference(name, over):
    """Engine vs the reference goldens of the flag variants (ops.mask_kl for klm_loss; prior flags)."""
    import numpy as np
    import util_parity as U
    from test_oracle_golden import build_engine_model, tape_from_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name + '.npz'))
    model, K, img, B, gen = (str(v) for v in g['meta'])
    m, cfg = build_engine_model(model, int(K), int(img), **over)
    m = m.cuda().train()
    recon, losses, stats, att, comp = U.run_engine(m, torch.from_numpy(g['x']), tape_from_golden(g))
    np.testing.assert_allclose(losses['err'].detach().cpu().numpy(), g['err'], rtol=1e-4)
    if 'kl_m' in g.files:
        np.testing.assert_allclose(losses['kl_m'].detach().cpu().numpy(), g['kl_m'], rtol=1e-3, atol=1e-2)
    gmax = max(float(s[0]) for s in g['grad_sums'])
    params = dict(m.named_parameters())
    worst = 0.0
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        if params[str(n)].grad is None:
            assert nrm == 0.0, n
            continue
        gd = params[str(n)].grad.detach().double().cpu().flatten()
        worst = max(worst, abs(gd.norm().item() - nrm) / (nrm + 1e-3 * gmax))
    assert worst <= 0.3, worst          # V2 per-tensor TF32 tolerance of DESIGN.md section 5


@pytest.mark.parametrize('kernel', ['gaussian', 'laplacian', 'epanechnikov'])
@pytest.mark.parametrize('K', [2, 6])
def test_icsbp_kernel_types_match_oracle(kernel, K):
    """icsbp_fwd/bwd_kernel<KT> (csrc/v2.cu) against the oracle's InstanceColouringSBP (pinned to the reference by the
    variant goldens) on the same colours / uniform draws: seeds identical, log-masks, d colour, d log_sigma."""
    from genesis_b200 import ops
    from oracle import models as M
    torch.manual_seed(K)
    B, S, CD = 3, 32, 8
    scale = torch.tensor([0.03, 0.3, 1.5])[torch.randint(0, 3, (B, S, S, 1))]
    colour = scale * torch.randn(B, S, S, CD)
    u = torch.rand(B, 1, S, S)
    sigma0 = {'gaussian': 1.0 / (K * 0.6931), 'laplacian': 1.0 / (K ** 0.5 * 0.6931), 'epanechnikov': 2.0 / K}[kernel]
    ls = torch.tensor(sigma0, dtype=torch.float64).log()
    c64 = colour.double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ls64 = ls.clone().requires_grad_(True)
    log_m_k, log_s_k, seeds, idxs = M.icsbp(c64, u.double(), ls64, K - 1, kernel)
    ref = torch.stack(log_m_k, 0)
    w = torch.randn(ref.shape, dtype=torch.float64)
    (ref * w).sum().backward()
    cg = colour.cuda().requires_grad_(True)
    lsg = ls.cuda().requires_grad_(True)
    log_m, log_s, idx = ops.icsbp(cg, u.cuda(), lsg, K, kernel)
    assert torch.equal(idx.cpu().long(), torch.stack(idxs, 0))
    torch.testing.assert_close(log_m.detach().cpu().double(), ref.detach(), rtol=1e-4, atol=2e-4)
    (log_m * w.float().cuda()).sum().backward()
    gref = c64.grad.permute(0, 2, 3, 1)
    err = (cg.grad.cpu().double() - gref).norm() / gref.norm()
    assert err < 1e-4, err
    assert abs(lsg.grad.item() - ls64.grad.item()) <= 1e-4 * abs(ls64.grad.item()) + 1e-3


@pytest.mark.parametrize('name', ['variant_genesisv2_k6_dynamic_b3', 'variant_genesisv2_k8_dynamic_b1'])
def test_dynamic_K_engine_matches_reference(name):
    """GENESIS-V2 dynamic_K: engine (icsbp_fwd_kernel<KT, true>, per-image mask counts in the backward) vs the reference goldens
    (batch: -1e10 padded masks; single image: fewer slots).  fp32 engine precision: the early-exit threshold is a hard
    comparison of a pixel count with 20."""
    import numpy as np
    import util_parity as U
    from genesis_b200 import ops
    from test_oracle_golden import build_engine_model, tape_from_golden
    from test_variants_golden import apply_param_add, overrides
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name + '.npz'))
    model, K, img, B, gen = (str(v) for v in g['meta'])
    m, cfg = build_engine_model(model, int(K), int(img), **overrides(g))
    apply_param_add(m, g)
    m = m.cuda().train()
    prev = ops.get_precision()
    ops.set_precision('fp32')
    try:
        recon, losses, stats, att, comp = U.run_engine(m, torch.from_numpy(g['x']), tape_from_golden(g))
    finally:
        ops.set_precision(prev)
    lm = torch.stack(list(stats['log_m_k']), 0).detach().cpu().numpy()
    assert lm.shape == g['log_m_k'].shape                         # same number of slots
    np.testing.assert_array_equal(lm < -1e9, g['log_m_k'] < -1e9)   # same padded slots
    np.testing.assert_allclose(np.where(lm < -1e9, 0, lm), np.where(g['log_m_k'] < -1e9, 0, g['log_m_k']), atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(losses['err'].detach().cpu().numpy(), g['err'], rtol=1e-4)
    gmax = max(float(s[0]) for s in g['grad_sums'])
    params = dict(m.named_parameters())
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        if params[str(n)].grad is None:
            assert nrm == 0.0, n
            continue
        gd = params[str(n)].grad.detach().double().cpu().flatten()
        assert abs(gd.norm().item() - nrm) <= 2e-2 * nrm + 1e-3 * gmax, (n, gd.norm().item(), nrm)


FULL_SIZE = [  # model, K, img, B (BASELINE.json configs c2, c3, c5 per GPU), generator, images of the subset run
    ('genesis', 5, 64, 64, 'multid', 8),
    ('genesisv2', 7, 64, 128, 'stacks', 8),
    ('monet', 7, 128, 64, 'multid', 4),
]


@pytest.mark.parametrize('model,K,img,B,gen,n', FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_identities_and_batch_partition_invariance(model, K, img, B, gen, n):
    """BASELINE.json's full per-GPU sizes, where the CPU oracle would take minutes: the size-independent properties of
    tests/full_size_props.py (checked on the CPU at small sizes by tests/test_full_size_props_cpu.py).  Identities under the
    default TF32 path; batch-partition invariance under the exact-fp32 path (tile shapes, hence summation orders, depend on
    the batch)."""
    import full_size_props as FP
    from genesis_b200 import ops
    from oracle import synth
    from test_oracle_golden import build_engine_model
    m, cfg = build_engine_model(model, K, img)
    m = m.cuda()
    x = torch.from_numpy(synth.GENERATORS[gen](B, img, 11)[0]).cuda()
    std = float(m.std) if model == 'genesisv2' else m.std.reshape(-1)

    def forward(xb, tape):
        m.set_noise_tape(tape)
        with torch.no_grad():
            out = m(xb)
        torch.cuda.synchronize()
        m.set_noise_tape(None)
        return out

    if model == 'genesis':
        m.train()
        forward(x, FP.SubsetTape(1, B, B, K))       # move the BatchNorm running statistics
        m.eval()
    else:
        m.train()
    full = forward(x, FP.SubsetTape(5, B, B, K))
    FP.check_identities(model, x, full, std, tol=5.0)
    prev = ops.get_precision()
    ops.set_precision('fp32')
    try:
        full32 = forward(x, FP.SubsetTape(5, B, B, K))
        sub32 = forward(x[:n], FP.SubsetTape(5, B, n, K))
    finally:
        ops.set_precision(prev)
    FP.check_subset_invariance(full32, sub32, n, rtol=2e-4, atol=2e-4)
    # training step at full size: every parameter receives a finite gradient
    import util_parity as U
    m.train()
    m.set_noise_tape(FP.SubsetTape(7, B, B, K))
    m.zero_grad(set_to_none=True)
    U.engine_total_loss(m(x)[1]).backward()
    torch.cuda.synchronize()
    m.set_noise_tape(None)
    for name, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name


def test_genesis_single_slot_matches_the_real_reference():
    """GENESIS with K_steps = 1 (reference genesis_config.py:94-96, 161-166, 226-228): no attention process (the core is built
    -- it consumes the seeded generator -- but is not part of the model), one all-ones mask, standard-normal component prior,
    `kl_m` = 0, att_stats None, sample() raises NotImplementedError.  Compared with the reference itself (oracle/_ref)."""
    import util_parity as U
    from oracle import models as M, ref_loader
    from genesis_b200.datasets import synth
    from genesis_b200.model_configs import genesis_config as plugin
    if not ref_loader.available():
        pytest.skip('neither the reference checkout nor oracle/_ref is present')
    cfg = M.make_cfg('genesis', K_steps=1, img_size=64)
    x = torch.from_numpy(synth.multid(3, 64, 4)[0])
    ref, out = U.run_reference('genesis', cfg, x, U.make_tape(9), seed=3)
    torch.manual_seed(3)
    eng = plugin.load(M.make_cfg('genesis', K_steps=1, img_size=64))
    assert [k for k, _ in eng.state_dict().items()] == [k for k, _ in ref.state_dict().items()]
    for (k, a), (_, b) in zip(eng.state_dict().items(), ref.state_dict().items()):
        assert torch.equal(a, b), k                       # seeded initialisation is bit-identical (same generator consumption)
    eng = eng.cuda().train()
    recon, losses, stats, att, comp = U.run_engine(eng, x, U.make_tape(9))
    assert att is None and out[3] is None
    assert U.rel_l2(losses['err'], out[1]['err']) < 1e-4
    assert float(losses['kl_m']) == 0.0 and float(out[1]['kl_m']) == 0.0
    assert (losses['kl_l_k'][0].cpu() - out[1]['kl_l_k'][0]).abs().max().item() < 1e-2
    assert U.rel_l2(recon, out[0]) < 1e-3
    assert float(stats['log_m_k'][0].abs().max()) == 0.0 and float(stats['log_s_k'][0].max()) == -1e10
    U.compare_grads_with_module(eng, ref, tol=1e-2)
    with pytest.raises(NotImplementedError):
        eng.sample(2)


def test_x_loss_pixel_wise_matches_the_real_reference():
    """Genesis.x_loss(..., pixel_wise=True) (reference genesis_config.py:273-286): values and the gradients w.r.t. the component
    means and the log-masks under a per-pixel upstream gradient."""
    from oracle import ref_loader
    from genesis_b200.model_configs import genesis_config as plugin
    if not ref_loader.available():
        pytest.skip('neither the reference checkout nor oracle/_ref is present')
    ref_loader._setup()
    from models.genesis_config import Genesis as RefGenesis
    torch.manual_seed(0)
    K, B, Hh = 4, 3, 16
    x = torch.rand(B, 3, Hh, Hh)
    xr = [torch.rand(B, 3, Hh, Hh, requires_grad=True) for _ in range(K)]
    lm = torch.log_softmax(torch.randn(B, 1, Hh, Hh, K), dim=4)
    lmk = [lm[..., k].clone().requires_grad_(True) for k in range(K)]
    std = torch.tensor([0.7, 0.5, 0.9, 0.7]).view(1, 1, 1, 1, K)
    w = torch.randn(B, 3, Hh, Hh)
    ref = RefGenesis.x_loss(x, lmk, xr, std, pixel_wise=True)
    (ref * w).sum().backward()
    xr_g = [t.detach().cuda().requires_grad_(True) for t in xr]
    lm_g = [t.detach().cuda().requires_grad_(True) for t in lmk]
    got = plugin.Genesis.x_loss(x.cuda(), lm_g, xr_g, std.cuda(), pixel_wise=True)
    assert got.shape == ref.shape
    (got * w.cuda()).sum().backward()
    torch.testing.assert_close(got.cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
    for a, b in zip(xr_g + lm_g, xr + lmk):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-4, atol=1e-5)
    # and the summed form is consistent with it
    summed = plugin.Genesis.x_loss(x.cuda(), lm_g, xr_g, std.cuda())
    torch.testing.assert_close(summed, got.sum((1, 2, 3)), rtol=1e-5, atol=1e-3)
