"""CPU: the plain-SIMT kernels that were written without a GPU at hand, executed on the CPU (tests/cuda_emu: the kernel source
compiled unchanged by g++, one OS thread per CUDA thread, real barriers and warp shuffles) and compared with numpy / torch /
the oracle: the fused latent kernels (csrc/latent.cu) and the IC-SBP kernels with
their kernel-type and dynamic_K instantiations (csrc/v2.cu).  The extern "C" entry points are emulated too, so argument
checks, grid sizes and kernel selection are covered.  This validates indexing, tile edges, barrier placement and scan order --
not performance, and not the product library (which is only ever run on a GPU)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cuda_emu'))
import build_emu  # noqa: E402

from oracle import models as M  # noqa: E402

P = ctypes.c_void_p


def ptr(a):
    return None if a is None else a.ctypes.data_as(P)


@pytest.fixture(scope='module')
def emu():
    libs = {}

    def get(name):
        if name not in libs:
            libs[name] = ctypes.CDLL(build_emu.build(name))
        return libs[name]
    return get


# ------------------------------------------------------------------------------------------------ fused latent kernels
def test_latent_lstm_cell(emu):
    lib = emu('latent.cu')
    torch.manual_seed(0)
    B, H = 5, 12
    gx = torch.randn(B, 4 * H, requires_grad=True)
    gh = torch.randn(B, 4 * H, requires_grad=True)
    for with_prev in (True, False):
        cp = torch.randn(B, H, requires_grad=True) if with_prev else None
        i, f, g, o = torch.chunk(gx + gh, 4, dim=1)
        c_ref = torch.sigmoid(i) * torch.tanh(g) + (torch.sigmoid(f) * cp if with_prev else 0)
        h_ref = torch.sigmoid(o) * torch.tanh(c_ref)
        h = np.empty((B, H), np.float32)
        c = np.empty((B, H), np.float32)
        cpn = cp.detach().numpy().copy() if with_prev else None
        assert lib.g2_lstm_cell_fwd_f32(ptr(gx.detach().numpy()), ptr(gh.detach().numpy()), ptr(cpn), ptr(h), ptr(c), B, H, None) == 0
        np.testing.assert_allclose(h, h_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(c, c_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
        dh, dc = torch.randn(B, H), torch.randn(B, H)
        grads = torch.autograd.grad((h_ref * dh).sum() + (c_ref * dc).sum(), [gx] + ([cp] if with_prev else []))
        dg = np.empty((B, 4 * H), np.float32)
        dcp = np.empty((B, H), np.float32) if with_prev else None
        assert lib.g2_lstm_cell_bwd_f32(ptr(gx.detach().numpy()), ptr(gh.detach().numpy()), ptr(cpn), ptr(c), ptr(dh.numpy()),
                                        ptr(dc.numpy()), ptr(dg), ptr(dcp), B, H, None) == 0
        np.testing.assert_allclose(dg, grads[0].numpy(), rtol=1e-4, atol=1e-6)
        if with_prev:
            np.testing.assert_allclose(dcp, grads[1].numpy(), rtol=1e-4, atol=1e-6)


def test_latent_heads_and_kl(emu):
    lib = emu('latent.cu')
    torch.manual_seed(1)
    B, D = 7, 10
    lo = (2 * torch.randn(B, 2 * D)).requires_grad_(True)
    eps = torch.randn(B, D)
    mu_ref, raw = torch.chunk(lo, 2, dim=1)
    sig_ref = torch.nn.functional.softplus(raw + 0.5) + 1e-8
    z_ref = mu_ref + sig_ref * eps
    z, mu, sig = (np.empty((B, D), np.float32) for _ in range(3))
    assert lib.g2_gauss_head_fwd_f32(ptr(lo.detach().numpy()), ptr(eps.numpy()), ptr(z), ptr(mu), ptr(sig), B, D, None) == 0
    np.testing.assert_allclose(z, z_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sig, sig_ref.detach().numpy(), rtol=1e-5, atol=1e-7)
    dz, dmu, dsig = torch.randn(B, D), torch.randn(B, D), torch.randn(B, D)
    (g_ref,) = torch.autograd.grad((z_ref * dz).sum() + (mu_ref * dmu).sum() + (sig_ref * dsig).sum(), [lo])
    dlo = np.empty((B, 2 * D), np.float32)
    assert lib.g2_gauss_head_bwd_f32(ptr(lo.detach().numpy()), ptr(eps.numpy()), ptr(dz.numpy()), ptr(dmu.numpy()), ptr(dsig.numpy()),
                                     ptr(dlo), B, D, None) == 0
    np.testing.assert_allclose(dlo, g_ref.numpy(), rtol=1e-4, atol=1e-6)
    # prior head
    for use_tanh in (1, 0):
        a, b = torch.chunk(lo, 2, dim=1)
        pm_ref = torch.tanh(a) if use_tanh else a
        ps_ref = torch.sigmoid(b + 4) + 1e-4
        pm, ps = np.empty((B, D), np.float32), np.empty((B, D), np.float32)
        assert lib.g2_prior_head_fwd_f32(ptr(lo.detach().numpy()), ptr(pm), ptr(ps), B, D, use_tanh, None) == 0
        np.testing.assert_allclose(pm, pm_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ps, ps_ref.detach().numpy(), rtol=1e-5, atol=1e-7)
        (g_ref,) = torch.autograd.grad((pm_ref * dz).sum() + (ps_ref * dsig).sum(), [lo])
        assert lib.g2_prior_head_bwd_f32(ptr(pm), ptr(ps), ptr(dz.numpy()), ptr(dsig.numpy()), ptr(dlo), B, D, use_tanh, None) == 0
        np.testing.assert_allclose(dlo, g_ref.numpy(), rtol=1e-4, atol=1e-6)
    # Monte-Carlo KL, D not a multiple of the warp size
    from oracle import functional as O
    D2 = 45
    zz, m, pm = (torch.randn(B, D2, requires_grad=True) for _ in range(3))
    s = (torch.rand(B, D2) + 0.3).requires_grad_(True)
    ps = (torch.rand(B, D2) + 0.2).requires_grad_(True)
    for prior in (True, False):
        kl_ref = O.mc_kl(zz, m, s, pm if prior else None, ps if prior else None)
        kl = np.empty(B, np.float32)
        args = [t.detach().numpy() for t in (zz, m, s)] + ([pm.detach().numpy(), ps.detach().numpy()] if prior else [None, None])
        assert lib.g2_mc_kl_fwd_f32(*[ptr(a) for a in args], ptr(kl), B, D2, None) == 0
        np.testing.assert_allclose(kl, kl_ref.detach().numpy(), rtol=1e-4, atol=1e-4)
        dkl = torch.randn(B)
        grads = torch.autograd.grad((kl_ref * dkl).sum(), [zz, m, s] + ([pm, ps] if prior else []))
        outs = [np.empty((B, D2), np.float32) for _ in range(5)]
        assert lib.g2_mc_kl_bwd_f32(*[ptr(a) for a in args], ptr(dkl.numpy()), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]),
                                    ptr(outs[3]) if prior else None, ptr(outs[4]) if prior else None, B, D2, None) == 0
        for got, ref in zip(outs, grads):
            np.testing.assert_allclose(got, ref.numpy(), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------ IC-SBP
def icsbp_inputs(B, S, seed):
    torch.manual_seed(seed)
    scale = torch.tensor([0.03, 0.3, 1.5])[torch.randint(0, 3, (B, S, S, 1))]
    colour = (scale * torch.randn(B, S, S, 8)).contiguous()
    u = torch.rand(B, 1, S, S)
    return colour, u


@pytest.mark.parametrize('kernel,kt', [('gaussian', 0), ('laplacian', 1), ('epanechnikov', 2)])
@pytest.mark.parametrize('S,K', [(16, 4), (40, 3)])          # 40 x 40 = 1600 pixels: two pixels per thread
def test_icsbp_kernels(emu, kernel, kt, S, K):
    lib = emu('v2.cu')
    B = 2
    colour, u = icsbp_inputs(B, S, 3 + kt)
    sigma0 = {0: 1.0 / (K * 0.6931), 1: 1.0 / (K ** 0.5 * 0.6931), 2: 2.0 / K}[kt]
    ls = torch.tensor(sigma0).log()
    c_ref = colour.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ls_ref = ls.clone().requires_grad_(True)
    log_m_k, log_s_k, seeds, idxs = M.icsbp(c_ref, u, ls_ref, K - 1, kernel)
    ref = torch.stack(log_m_k, 0)                                             # [K,B,1,S,S]
    Pn = S * S
    log_m, log_s = np.empty((K, B, Pn), np.float32), np.empty((K, B, Pn), np.float32)
    idx = np.empty((K - 1, B), np.int32)
    lsn = np.array([ls.item()], np.float32)
    if kt == 0:     # the validated gaussian entry point and the kernel-type entry point must agree
        assert lib.g2_icsbp_fwd_f32(ptr(colour.numpy()), ptr(u.numpy()), ptr(lsn), ptr(log_m), ptr(log_s), ptr(idx), B, Pn, K, 8, None) == 0
        first = log_m.copy()
    assert lib.g2_icsbp_kernel_fwd_f32(ptr(colour.numpy()), ptr(u.numpy()), ptr(lsn), ptr(log_m), ptr(log_s), ptr(idx), B, Pn, K, 8, kt, None) == 0
    if kt == 0:
        np.testing.assert_array_equal(first, log_m)
    np.testing.assert_array_equal(idx, torch.stack(idxs, 0).numpy())
    np.testing.assert_allclose(log_m.reshape(ref.shape), ref.detach().numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(log_s.reshape(ref.shape), torch.stack(log_s_k[:K], 0).detach().numpy(), rtol=1e-4, atol=1e-4)
    w = torch.randn(ref.shape)
    gc, gs = torch.autograd.grad((ref * w).sum(), [c_ref, ls_ref])
    dcol, dsig = np.empty((B, Pn, 8), np.float32), np.empty(B, np.float32)
    assert lib.g2_icsbp_kernel_bwd_f32(ptr(colour.numpy()), ptr(lsn), ptr(idx), ptr(w.numpy()), ptr(dcol), ptr(dsig), B, Pn, K, 8, kt, None) == 0
    gref = gc.permute(0, 2, 3, 1).reshape(B, Pn, 8).numpy()
    assert np.linalg.norm(dcol - gref) <= 2e-4 * np.linalg.norm(gref)
    assert abs(dsig.sum() - gs.item()) <= 2e-4 * abs(gs.item()) + 1e-4


@pytest.mark.parametrize('kernel,kt', [('gaussian', 0), ('epanechnikov', 2)])
def test_icsbp_dynamic_K(emu, kernel, kt):
    """Early exit per image, -1e10 padding, n_masks, and the backward's per-image scan length."""
    lib = emu('v2.cu')
    B, S, K = 3, 16, 7
    colour, u = icsbp_inputs(B, S, 11 + kt)
    colour = colour * 0.3
    sigmas = {0: 3.0, 2: 6.0}
    ls = torch.tensor(sigmas[kt]).log()                                       # wide kernel: the scope empties quickly
    Pn = S * S
    c_ref = colour.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ls_ref = ls.clone().requires_grad_(True)
    per = [M.icsbp(c_ref[b:b + 1], u[b:b + 1], ls_ref, K - 1, kernel, dynamic_K=True) for b in range(B)]
    n_ref = [len(r[0]) for r in per]
    assert min(n_ref) < K                                                     # the early exit is taken
    log_m, log_s = np.empty((K, B, Pn), np.float32), np.empty((K, B, Pn), np.float32)
    idx = np.empty((K - 1, B), np.int32)
    n = np.empty(B, np.int32)
    lsn = np.array([ls.item()], np.float32)
    assert lib.g2_icsbp_dynamic_fwd_f32(ptr(colour.numpy()), ptr(u.numpy()), ptr(lsn), ptr(log_m), ptr(log_s), ptr(idx), ptr(n),
                                        B, Pn, K, 8, kt, None) == 0
    assert n.tolist() == n_ref
    w = torch.randn(K, B, Pn)
    loss = 0
    for b in range(B):
        for k in range(K):
            if k < n_ref[b]:
                np.testing.assert_allclose(log_m[k, b], per[b][0][k].detach().numpy().reshape(-1), rtol=1e-4, atol=1e-4)
                loss = loss + (per[b][0][k].reshape(-1) * w[k, b]).sum()
            else:
                assert (log_m[k, b] == -1e10).all()
        steps = min(n_ref[b], K - 1)                                          # the step that stopped the loop still recorded its seed
        assert idx[:steps, b].tolist() == [int(i) for i in per[b][3]][:steps]
        assert (idx[steps:, b] == -1).all()
    gc, gs = torch.autograd.grad(loss, [c_ref, ls_ref])
    dcol, dsig = np.empty((B, Pn, 8), np.float32), np.empty(B, np.float32)
    assert lib.g2_icsbp_dynamic_bwd_f32(ptr(colour.numpy()), ptr(lsn), ptr(idx), ptr(n), ptr(w.numpy()), ptr(dcol), ptr(dsig),
                                        B, Pn, K, 8, kt, None) == 0
    gref = gc.permute(0, 2, 3, 1).reshape(B, Pn, 8).numpy()
    assert np.linalg.norm(dcol - gref) <= 2e-4 * np.linalg.norm(gref)
    assert abs(dsig.sum() - gs.item()) <= 2e-4 * abs(gs.item()) + 1e-4


# ------------------------------------------------------------------------------------------------ optimiser, scans, losses
def test_fused_adam_matches_torch_optim(emu):
    """g2_adam_f32 (flat arena, step counter read from memory, gradient scale, fused zero-grad) vs torch.optim.Adam."""
    lib = emu('pointwise.cu')
    lib.g2_adam_f32.argtypes = [P, P, P, P, ctypes.c_long, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, P,
                                ctypes.c_float, ctypes.c_int, P]
    torch.manual_seed(0)
    n = 1024 + 64
    p0 = torch.randn(n)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.numpy().copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    step = np.zeros(1, np.float32)
    world = 2.0
    for it in range(4):
        g = torch.randn(n) * (10.0 ** (it - 2))
        ref.grad = g.clone()
        opt.step()
        gbuf = (g * world).numpy().copy()           # the arena holds the SUM over ranks; the kernel applies 1 / world
        step += 1
        assert lib.g2_adam_f32(ptr(p), ptr(gbuf), ptr(m), ptr(v), n, 1e-3, 0.9, 0.999, 1e-8, ptr(step), 1.0 / world, 1, None) == 0
        assert np.abs(gbuf).max() == 0.0
        assert np.abs(p - ref.detach().numpy()).max() <= 2e-6 * max(1.0, ref.detach().abs().max().item())


@pytest.mark.parametrize('kind', ['rmsprop', 'sgd'])
def test_fused_rmsprop_sgd_match_torch_optim(emu, kind):
    """g2_rmsprop_f32 / g2_sgd_f32 vs torch.optim.RMSprop(lr) / torch.optim.SGD(lr, 0.9) (train.py:171-176)."""
    lib = emu('pointwise.cu')
    lib.g2_rmsprop_f32.argtypes = [P, P, P, ctypes.c_long, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, P]
    lib.g2_sgd_f32.argtypes = [P, P, P, ctypes.c_long, ctypes.c_float, ctypes.c_float, P, ctypes.c_float, ctypes.c_int, P]
    torch.manual_seed(0)
    n = 512 + 64
    p0 = torch.randn(n)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.RMSprop([ref], 1e-3) if kind == 'rmsprop' else torch.optim.SGD([ref], 1e-3, 0.9)
    p, m = p0.numpy().copy(), np.zeros(n, np.float32)
    step = np.zeros(1, np.float32)
    for it in range(5):
        g = torch.randn(n) * (10.0 ** (it - 2))
        ref.grad = g.clone()
        opt.step()
        gbuf = (g * 2.0).numpy().copy()
        step += 1
        if kind == 'rmsprop':
            assert lib.g2_rmsprop_f32(ptr(p), ptr(gbuf), ptr(m), n, 1e-3, 0.99, 1e-8, 0.5, 1, None) == 0
        else:
            assert lib.g2_sgd_f32(ptr(p), ptr(gbuf), ptr(m), n, 1e-3, 0.9, ptr(step), 0.5, 1, None) == 0
        assert np.abs(gbuf).max() == 0.0
        assert np.abs(p - ref.detach().numpy()).max() <= 3e-6 * max(1.0, ref.detach().abs().max().item()), it


@pytest.mark.parametrize('nl_is_K', [True, False])
def test_sbp_scan(emu, nl_is_K):
    lib = emu('pointwise.cu')
    lib.g2_sbp_scan_fwd_f32.argtypes = [P, P, P, ctypes.c_long, ctypes.c_int, ctypes.c_int, P]
    lib.g2_sbp_scan_bwd_f32.argtypes = [P, P, P, ctypes.c_long, ctypes.c_int, ctypes.c_int, P]
    import cpu_ops_mock
    torch.manual_seed(2)
    K, BP = 5, 300
    nl = K if nl_is_K else K - 1
    lg = (2 * torch.randn(nl, BP)).requires_grad_(True)
    m_ref, s_ref = cpu_ops_mock.sbp_scan(lg, K)
    log_m, log_s = np.empty((K, BP), np.float32), np.empty((nl + 1, BP), np.float32)
    assert lib.g2_sbp_scan_fwd_f32(ptr(lg.detach().numpy()), ptr(log_m), ptr(log_s), BP, K, nl, None) == 0
    np.testing.assert_allclose(log_m, m_ref.detach().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(log_s, s_ref.numpy(), rtol=1e-5, atol=1e-5)
    w = torch.randn(K, BP)
    (g_ref,) = torch.autograd.grad((m_ref * w).sum(), [lg])
    dlg = np.empty((nl, BP), np.float32)
    assert lib.g2_sbp_scan_bwd_f32(ptr(lg.detach().numpy()), ptr(w.numpy()), ptr(dlg), BP, K, nl, None) == 0
    np.testing.assert_allclose(dlg, g_ref.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('cs', [1, 4], ids=['plain-logits', 'packed-decoder-plane'])
def test_mask_kl_both_layouts(emu, cs):
    """The stand-alone mask KL of ops.mask_kl: logits as their own [K,B,1,P] tensor (MONet prior_mode='scope') or as plane 3 of
    the packed decoder output [K,B,4,P] (GENESIS-V2 klm_loss)."""
    lib = emu('loss.cu')
    import cpu_ops_mock
    torch.manual_seed(3)
    K, B, Hh = 4, 3, 10
    Pn = Hh * Hh
    lm = torch.log_softmax(2 * torch.randn(K, B, 1, Hh, Hh), dim=0).requires_grad_(True)
    dec = torch.randn(K, B, cs, Hh, Hh, requires_grad=True)
    kl_ref = cpu_ops_mock.mask_kl(lm, dec, detach=False)
    decn = dec.detach().numpy()
    off = 3 * Pn if cs == 4 else 0
    lg_ptr = ctypes.c_void_p(decn.ctypes.data + 4 * off)
    lmr, kl = np.empty((K, B, Pn), np.float32), np.empty(B, np.float32)
    assert lib.g2_mask_kl_fwd_f32(ptr(lm.detach().numpy()), lg_ptr, ptr(lmr), ptr(kl), K, B, Pn, 1, cs, None) == 0
    np.testing.assert_allclose(kl, kl_ref.detach().numpy(), rtol=1e-4, atol=1e-4)
    plane = dec.detach()[:, :, 3:] if cs == 4 else dec.detach()
    np.testing.assert_allclose(lmr.reshape(K, B, 1, Hh, Hh), torch.log_softmax(plane, dim=0).numpy(), rtol=1e-5, atol=1e-5)
    gkl = torch.randn(B)
    g_lm, g_dec = torch.autograd.grad((kl_ref * gkl).sum(), [lm, dec])
    dlm = np.empty((K, B, Pn), np.float32)
    ddec = np.zeros_like(decn)
    dlg_ptr = ctypes.c_void_p(ddec.ctypes.data + 4 * off)
    assert lib.g2_mask_kl_bwd_f32(ptr(lm.detach().numpy()), lg_ptr, ptr(gkl.numpy()), ptr(dlm), dlg_ptr, K, B, Pn, 1, cs, 1, cs, 0,
                                  None) == 0
    np.testing.assert_allclose(dlm.reshape(g_lm.shape), g_lm.numpy(), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(ddec, g_dec.numpy(), rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize('seed,deg,K', [(0, False, 7), (2, True, 5)])
def test_seg_metrics_kernel(emu, seed, deg, K):
    """g2_seg_metrics (one confusion matrix per image -> ARI / segmentation covering / argmax map) vs the oracle, which is
    pinned to the reference's utils/misc functions (tests/test_metrics.py)."""
    from oracle import metrics as OM
    from test_metrics import random_case
    lib = emu('metrics.cu')
    log_m, inst = random_case(seed, K=K, degenerate=deg)
    _, B, H, _ = log_m.shape
    Pn = H * H
    lm = np.ascontiguousarray(log_m.reshape(K, B, Pn).astype(np.float32))
    ins = np.ascontiguousarray(inst.reshape(B, Pn).astype(np.int64))
    want = OM.metrics(lm, ins)
    seg = np.empty((B, Pn), np.int64)
    out = np.empty((B, 8), np.float64)
    assert lib.g2_seg_metrics(ptr(lm), None, ptr(ins), ptr(seg), ptr(out), B, Pn, K, None) == 0
    for j, key in enumerate(('ari', 'ari_fg', 'msc', 'msc_fg', 'msc_scaled', 'msc_fg_scaled')):
        np.testing.assert_allclose(out[:, j], want[key], atol=1e-12, err_msg=key)
    np.testing.assert_array_equal(seg, want['instance_seg'].reshape(B, Pn))


# ------------------------------------------------------------------------------------------------ tcgen05 / TMA kernels
def _run_child(script, cases, mode, env_extra):
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, EMU_TIMEOUT_S='20', **env_extra)
    cmd = [sys.executable, os.path.join(here, 'cuda_emu', script), ';'.join(cases)] + ([mode] if mode else [])
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count('OK') == len(cases), r.stdout[-2000:]


HALO_CASES = [  # mode N H W Ci Co R stride pad act
    '0 2 18 20 32 32 3 1 1 1',        # 3x3, resident weights in the persistent variant
    '1 3 16 16 64 64 5 1 2 0',        # 5x5 conv-transpose, two channel blocks, several images per CTA, weight ring
    '0 2 19 23 32 32 3 1 1 0',        # ragged: bottom-edge tiles
    '1 2 16 16 32 64 5 2 2 1',        # stride-2 conv-transpose: four sub-pixel classes
    '0 2 16 16 32 256 3 1 1 1',       # two output-channel blocks: bias reload between items
]


def test_halo_conv_kernel_calibrates_the_model():
    """conv_halo_kernel is validated on a B200 (tests/test_halo_gpu.py); under the functional model of mbarriers, TMA and
    tcgen05 (tests/cuda_emu/cuda_emu_sm100.h) it must give the same exact answers -- this is what entitles the model to say
    anything about the kernels below."""
    _run_child('run_halo_emu.py', HALO_CASES[:4], None, {})


def test_persistent_halo_kernel_under_the_model():
    """conv_halo_persistent_kernel (G2_HALO_PERSISTENT=1, never run on a GPU): exact results with two CTAs walking several
    items each -- double-buffered windows, alternating accumulator sets, weight ring and resident weights, bias reload."""
    _run_child('run_halo_emu.py', HALO_CASES, 'persistent', {'G2_HALO_PERSISTENT_CTAS': '2'})


def test_persistent_halo_kernel_3xtf32_under_the_model():
    """3xTF32 inside the persistent kernel: units (item, channel block) issued in pairs P12(u) P12(u+1) P3(u) P3(u+1), the
    rewrite warps turning each window into x_lo in place between its passes, resident and streamed [w_hi | w_lo] packs, two CTAs
    walking several units each; and the one-item kernel's in-kernel split (G2_HALO_X3_PERSISTENT=0) on the same cases."""
    _run_child('run_halo_emu.py', HALO_CASES, 'x3', {'G2_HALO_PERSISTENT_CTAS': '2', 'G2_HALO_X3_PERSISTENT': '1', 'G2_HALO_X3_RESIDENT': '1'})
    _run_child('run_halo_emu.py', HALO_CASES[:2], 'x3', {'G2_HALO_PERSISTENT_CTAS': '2', 'G2_HALO_X3_PERSISTENT': '1', 'G2_HALO_X3_RESIDENT': '0'})
    _run_child('run_halo_emu.py', HALO_CASES[:3], 'x3', {'G2_HALO_X3_PERSISTENT': '0'})


WGRAD_CASES = [  # N Hg Wg Cg Ct R pad
    '4 18 18 32 32 3 1',              # 3x3: three-tap groups, the unused fourth atom reads the zeroed slack
    '3 20 20 64 32 5 2',              # 5x5: row groups + a column group + a single, two channel blocks per CTA, 2 windows per CTA
    '5 16 16 128 64 5 2',             # grid_y = 4, ten windows per CTA
    '1 19 23 32 64 3 1',              # ragged
    '2 22 22 32 32 3 0',              # VALID
]


def test_wgrad_tile_kernel_calibrates_the_mn_major_model():
    """wgrad_tc_kernel is validated on a B200 (tests/test_tc_gpu.py): MN-major operands, SWIZZLE_128B_ATOM_32B boxes."""
    _run_child('run_wgrad_emu.py', WGRAD_CASES[:2], None, {})


def test_halo_wgrad_kernel_under_the_model():
    """wgrad_halo_kernel (G2_WGRAD_HALO=1, never run on a GPU): taps as LBO-strided M atoms of one resident window, exact under
    the model's rule that the swizzle is a function of the absolute shared-memory address (the rule the B200 probe in
    tests/test_pending_next_round.py checks on the hardware)."""
    _run_child('run_wgrad_emu.py', WGRAD_CASES, 'halo', {'G2_WGRAD_HALO_CTAS': '4'})
