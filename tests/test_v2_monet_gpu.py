"""GPU: GENESIS-V2 and MONet engines vs the golden vectors generated from the reference and vs the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import synth

import util_parity as U
from test_oracle_golden import build_engine_model, golden_case, tape_from_golden, direction

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, 'golden', 'genesisv2_*.npz')) + glob.glob(os.path.join(HERE, 'golden', 'monet*.npz')))

# Stated tolerances.  Forward: `err` rel 1e-4; KL terms abs 1e-2 + rel 1e-3; log masks
# |d| <= LM_TOL * (1 + |ref|).  Backward, per model and precision: GLOBAL = rel-L2 error of the whole gradient vector,
# TENSOR = worst per-parameter rel-L2 (denominator floored at 2e-4 of the largest gradient norm).
# 'tf32' = the PRODUCT path: tcgen05 TF32 operands / fp32 accumulate, with the layers that amplify operand rounding (GENESIS-V2:
# UNet backbone, seg / feat heads, z_head, first decoder layer; MONet: attention UNet, component encoder) run as 3xTF32
# (ops.precise -- fp32-level accuracy on the same tensor-core kernels).  'fp32' = exact-fp32 SIMT kernels everywhere.
# Round 1 ran plain TF32 everywhere and needed TENSOR 0.3 / GLOBAL 6e-2 (measured worst 0.16); measured now
# (profiles/r02_parity_x3.txt): GENESIS-V2 worst tensor 1.6e-2, global 1.4e-3; MONet worst 8.3e-3, global 1.4e-3 -- the level
# of the exact-fp32 path, whose own distance to the reference is set by the conditioning of the K-1 recurrent InstanceNorm UNet
# passes on flat synthetic images (the reference's fp32 gradients differ from its fp64 ones by 6.5e-4 per tensor there).
ERR_RTOL, KL_ATOL = 1e-4, 1e-2
LM_TOL = {'tf32': 5e-3, 'fp32': 2e-4}
# GENESIS-V2 on the exact-fp32 path: 12 repeats of the K = 11 case on a B200 gave a whole-vector distance of 2.1e-5 ... 9.2e-5
# (gpurun_out/r02_v2_flaky2.txt) -- the float atomics of the masked pooling / fp32 weight-gradient kernels change the last bits from
# run to run and the IC-SBP model amplifies them -- so the bound is 3e-4, not the 1e-4 a single lucky run suggests.
GLOBAL_TOL = {('genesisv2', 'tf32'): 5e-3, ('genesisv2', 'fp32'): 3e-4, ('monet', 'tf32'): 5e-3, ('monet', 'fp32'): 5e-3}
TENSOR_TOL = {('genesisv2', 'tf32'): 2e-2, ('genesisv2', 'fp32'): 2e-3, ('monet', 'tf32'): 2e-2, ('monet', 'fp32'): 2e-2}


@pytest.fixture(params=['tf32', 'fp32'])
def precision(request):
    from genesis_b200 import ops
    ops.set_precision(request.param)
    yield request.param
    ops.set_precision('tf32')


def _stack(lst):
    return torch.stack(list(lst), 0).detach().cpu().numpy()


@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_matches_reference_golden(path, precision):
    g, model, K, img, B = golden_case(path)
    gtol = TENSOR_TOL[(model, precision)]
    m, cfg = build_engine_model(model, K, img)
    m = m.cuda().train()
    recon, losses, stats, att, comp = U.run_engine(m, torch.from_numpy(g['x']), tape_from_golden(g))
    np.testing.assert_allclose(losses['err'].detach().cpu().numpy(), g['err'], rtol=ERR_RTOL)
    np.testing.assert_allclose(recon.detach().cpu().numpy(), g['recon'], atol=2e-3)
    lt = max(LM_TOL[precision], 1e-3)          # golden log masks carry the reference's own fp32 noise
    np.testing.assert_allclose(_stack(stats['log_m_k']), g['log_m_k'], atol=lt, rtol=lt)
    np.testing.assert_allclose(_stack(stats['log_m_r_k']), g['log_m_r_k'], atol=lt, rtol=lt)
    np.testing.assert_allclose(_stack(losses['kl_l_k']), g['kl_l_k'], atol=KL_ATOL, rtol=1e-3)
    if 'kl_m' in g.files:
        np.testing.assert_allclose(losses['kl_m'].detach().cpu().numpy(), g['kl_m'], rtol=2e-3, atol=KL_ATOL)
    gmax = max(float(s[0]) for s in g['grad_sums'])
    params = dict(m.named_parameters())
    for i, (n, (nrm, proj)) in enumerate(zip(g['grad_names'], g['grad_sums'])):
        p = params[str(n)]
        if p.grad is None:
            assert nrm == 0.0, n
            continue
        gd = p.grad.detach().double().cpu().flatten()
        tol = gtol * nrm + 2e-4 * gmax
        assert abs(gd.norm().item() - nrm) <= tol, (n, gd.norm().item(), nrm)
        assert abs((gd * direction(gd.numel(), i)).sum().item() - proj) <= 4 * tol, (n, proj)


# (model, K, B, img, generator, data seed).  The K=11 rooms batch uses data seed 12: with seed 11 channel 9 of seg_head at one
# IC-SBP seed pixel sits 7e-6 from the ReLU kink (oracle att['seed_kink_margin']); the gradient of seg_head.0.weight /
# seg_head.1.bias is discontinuous there and implementations that differ by 1e-6 in the UNet output (the two 3xTF32 routes,
# both 5e-6 from fp64, gpurun_out/r02_x3_bisect.txt) land 14 % apart on those two tensors.  The guard below keeps the comparison
# well-posed for every case.
CASES = [('genesisv2', 7, 4, 64, 'stacks', 11), ('genesisv2', 3, 5, 64, 'rooms', 11), ('genesisv2', 11, 2, 64, 'rooms', 12),
         ('monet', 7, 3, 64, 'multid', 11), ('monet', 2, 2, 64, 'stacks', 11), ('monet', 3, 2, 128, 'multid', 11)]
KINK_MARGIN = 1e-4


@pytest.mark.parametrize('model,K,B,img,gen,dseed', CASES)
def test_matches_oracle(model, K, B, img, gen, dseed, precision):
    gtol = TENSOR_TOL[(model, precision)]
    m, cfg = build_engine_model(model, K, img, seed=3)
    m = m.cuda().train()
    if model == 'genesisv2':
        # the SemiConv gate is initialised to 0, which zeroes every gradient of seg_head / colour_head: open it
        with torch.no_grad():
            m.att_process.colour_head.gate.gate.fill_(0.3)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.from_numpy(synth.GENERATORS[gen](B, img, dseed)[0])
    tape = U.make_tape(5)
    out, P = U.run_oracle(model, sd0, x, tape, cfg)
    recon, losses, stats, att, comp = U.run_engine(m, x, tape.rewound())
    if model == 'genesisv2':
        assert out['att']['seed_kink_margin'] > KINK_MARGIN, 'ill-posed case: a seed pixel sits on a ReLU kink'
        # discrete seed choice first: identical argmax indices
        ref_idx = torch.stack(out['att']['seed_idx'], 0)
        assert torch.equal(att['seed_idx'].cpu().long(), ref_idx), 'IC-SBP seeds differ'
    assert U.rel_l2(losses['err'], out['err']) < ERR_RTOL
    assert U.rel_l2(recon, out['recon']) < 1e-3
    a = torch.stack(list(losses['kl_l_k']), 0).detach().cpu()
    b = torch.stack(out['kl_l_k'], 0).detach()
    assert (a - b).abs().max().item() < KL_ATOL + 1e-3 * b.abs().max().item()
    if model == 'monet':
        assert U.rel_l2(losses['kl_m'], out['kl_m']) < 2e-3
    for k in range(K):
        for key in ('log_m_k', 'log_m_r_k'):
            ref = out[key][k].detach()
            d = (stats[key][k].detach().cpu() - ref).abs() / (1 + ref.abs())
            assert d.max().item() < LM_TOL[precision], (key, k, d.max().item())
        assert U.rel_l2(stats['x_r_k'][k], out['x_r_k'][k]) < 1e-3
        assert U.rel_l2(comp['z_k'][k], out['comp']['z_k'][k]) < 1e-3
    for key in ('log_m_k', 'log_m_r_k'):      # reference utils/misc.py:258-270
        s = torch.stack(list(stats[key]), 0).exp().sum(0)
        assert (s - 1).abs().max().item() < 1e-3
    worst = U.compare_grads(m, P, gtol, floor_frac=2e-4)
    glob_err = U.global_grad_rel_l2(m, P)
    print('worst grad rel-L2', worst, 'global', glob_err)
    assert glob_err < GLOBAL_TOL[(model, precision)], glob_err


@pytest.mark.parametrize('model,K', [('genesisv2', 5), ('monet', 4)])
def test_eval_sample_and_state_dict(model, K):
    m, cfg = build_engine_model(model, K, 64, seed=1)
    m = m.cuda().eval()
    x = torch.from_numpy(synth.multid(2, 64, 3)[0]).cuda()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    tape = U.make_tape(9)
    with torch.no_grad():
        m.set_noise_tape(tape)
        recon, losses, stats, att, comp = m(x)
        m.set_noise_tape(None)
    out, _ = U.run_oracle(model, sd0, x.cpu(), tape.rewound(), cfg, training=False)
    assert U.rel_l2(losses['err'], out['err']) < ERR_RTOL
    assert U.rel_l2(recon, out['recon']) < 1e-3
    img, st = m.sample(3)
    assert img.shape == (3, 3, 64, 64) and torch.isfinite(img).all()
    assert len(st['log_m_k']) == K
    s = torch.stack(list(st['log_m_k']), 0).exp().sum(0)
    assert (s - 1).abs().max().item() < 1e-3
    m2, _ = build_engine_model(model, K, 64, seed=7)
    m2.load_state_dict(m.state_dict())
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1.cpu(), v2.cpu())


def test_mask_kl_kernel_matches_torch():
    """g2_mask_kl_{fwd,bwd}: MONet.kl_m_loss + log-softmax of the reconstructed mask logits."""
    from genesis_b200 import ops
    from oracle import models as M
    torch.manual_seed(0)
    K, B, H = 5, 3, 16
    x = torch.rand(B, 3, H, H)
    dec = torch.randn(K, B, 4, H, H)
    dec[:, :, :3] = torch.sigmoid(dec[:, :, :3])
    dec[0, :, 3] -= 14.0                                   # drives some reconstructed masks under the 1e-5 floor
    lm = torch.log_softmax(3 * torch.randn(K, B, 1, H, H), dim=0)
    lm[1, 0] = -30.0
    std = torch.full((K,), 0.7)
    dec_g = dec.clone().cuda().requires_grad_(True)
    lm_g = lm.clone().cuda().requires_grad_(True)
    err, kl, recon, lmr = ops.monet_loss(x.cuda(), dec_g, lm_g, std.cuda())
    w = torch.linspace(0.5, 1.5, B).cuda()
    ((err + 2.0 * kl) * w).sum().backward()
    dec_c = dec.clone().requires_grad_(True)
    lm_c = lm.clone().requires_grad_(True)
    lm_k = list(lm_c.unbind(0))
    lmr_k = M.mask_recon_log_softmax([dec_c[k, :, 3:] for k in range(K)])
    kl_ref = M.monet_kl_m(lm_k, lmr_k)
    from oracle import functional as O
    err_ref = O.mixture_nll(x, lm_k, [dec_c[k, :, :3] for k in range(K)], std)
    ((err_ref + 2.0 * kl_ref) * w.cpu()).sum().backward()
    assert U.rel_l2(kl, kl_ref) < 1e-5
    assert U.rel_l2(err, err_ref) < 1e-5
    assert U.rel_l2(lmr, torch.stack(lmr_k, 0)) < 1e-5
    assert U.rel_l2(dec_g.grad, dec_c.grad) < 1e-4
    assert U.rel_l2(lm_g.grad, lm_c.grad) < 1e-4
