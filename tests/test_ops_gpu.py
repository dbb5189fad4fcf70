"""GPU: every operator of the C ABI against a plain torch fp32/fp64 CPU evaluation of the same op,
forward and backward, on the layer shapes the hot path uses."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def _ops():
    from genesis_b200 import ops
    return ops


@pytest.fixture(autouse=True)
def exact_fp32_kernels():
    """These tests pin the exact-fp32 SIMT kernels; the tensor-core path is covered by test_tc_gpu.py and
    by test_ops_tf32 below."""
    ops = _ops()
    ops.set_precision('fp32')
    yield
    ops.set_precision('tf32')


def close(a, b, rtol=2e-4, atol=None, name=''):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = b.abs().max().item() + 1e-30
    err = (a - b).abs().max().item()
    tol = rtol * scale if atol is None else atol + rtol * scale
    assert err <= tol, '%s: max abs err %.3e > tol %.3e (scale %.3e)' % (name, err, tol, scale)


def grads(outs, inputs, seed=0):
    """d(sum_i <out_i, probe_i>)/d inputs with deterministic probes."""
    g = torch.Generator().manual_seed(seed)
    loss = 0
    for o in outs:
        probe = torch.randn(o.shape, generator=g, dtype=torch.float64).to(o.dtype).to(o.device)
        loss = loss + (o * probe).sum()
    return torch.autograd.grad(loss, inputs, allow_unused=True)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


CONV_CASES = [  # N,H,W,Ci,Co,R,stride,pad,act
    (2, 16, 16, 3, 64, 5, 1, 2, None),      # sylvester encoder layer 0 (Ci=3: scalar loads)
    (2, 16, 16, 32, 64, 5, 2, 2, None),     # stride-2 gated conv
    (3, 16, 16, 64, 128, 5, 1, 2, None),
    (2, 16, 16, 4, 32, 3, 2, 1, 'elu'),     # component encoder layer 0
    (2, 8, 8, 32, 64, 3, 2, 1, 'relu'),
    (2, 14, 14, 32, 32, 3, 1, 0, 'elu'),    # broadcast decoder VALID conv
    (1, 12, 12, 2, 32, 3, 1, 0, None),      # coordinate map conv
    (2, 16, 16, 64, 64, 3, 1, 1, None),     # UNet block conv (no bias tested below)
    (5, 10, 10, 24, 40, 3, 1, 1, None),     # ragged channel counts
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv2d(case):
    ops = _ops()
    N, H, W, Ci, Co, R, s, p, act = case
    torch.manual_seed(1)
    x = torch.randn(N, Ci, H, W, dtype=torch.float64)
    w = torch.randn(Co, Ci, R, R, dtype=torch.float64) * 0.1
    b = torch.randn(Co, dtype=torch.float64)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = F.conv2d(xr, wr, br, stride=s, padding=p)
    ref = {'elu': F.elu, 'relu': F.relu, None: lambda t: t}[act](ref)
    xg = nhwc(x).float().to(DEV).requires_grad_(True)
    wg = w.float().to(DEV).requires_grad_(True)
    bg = b.float().to(DEV).requires_grad_(True)
    out = ops.conv2d(xg, wg, bg, s, p, act)
    close(nchw(out), ref, name='fwd')
    gr = grads([ref], [xr, wr, br])
    gg = grads([nchw(out)], [xg, wg, bg])
    close(nchw(gg[0]), gr[0], name='dx')
    close(gg[1], gr[1], name='dw')
    close(gg[2], gr[2], name='db')


CONVT_CASES = [  # N,H,W,Ci,Co,R,stride,pad,act
    (2, 8, 8, 64, 128, 5, 1, 2, None),      # sylvester decoder s1 (gated: 2*64)
    (2, 8, 8, 64, 64, 5, 2, 2, None),       # s2, output_padding 1
    (3, 16, 16, 32, 64, 5, 2, 2, None),
    (2, 4, 4, 66, 64, 5, 2, 2, None),       # GENESIS-V2 decoder first layer (66 channels)
    (2, 6, 6, 16, 8, 3, 2, 1, 'relu'),
]


@pytest.mark.parametrize('case', CONVT_CASES)
def test_conv_transpose2d(case):
    ops = _ops()
    N, H, W, Ci, Co, R, s, p, act = case
    torch.manual_seed(2)
    x = torch.randn(N, Ci, H, W, dtype=torch.float64)
    w = torch.randn(Ci, Co, R, R, dtype=torch.float64) * 0.1
    b = torch.randn(Co, dtype=torch.float64)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = F.conv_transpose2d(xr, wr, br, stride=s, padding=p, output_padding=s - 1)
    ref = {'elu': F.elu, 'relu': F.relu, None: lambda t: t}[act](ref)
    xg = nhwc(x).float().to(DEV).requires_grad_(True)
    wg = w.float().to(DEV).requires_grad_(True)
    bg = b.float().to(DEV).requires_grad_(True)
    out = ops.conv_transpose2d(xg, wg, bg, s, p, act)
    close(nchw(out), ref, name='fwd')
    gr = grads([ref], [xr, wr, br])
    gg = grads([nchw(out)], [xg, wg, bg])
    close(nchw(gg[0]), gr[0], name='dx')
    close(gg[1], gr[1], name='dw')
    close(gg[2], gr[2], name='db')


def test_conv_no_bias():
    ops = _ops()
    torch.manual_seed(3)
    x = torch.randn(2, 8, 12, 12, dtype=torch.float64)
    w = torch.randn(16, 8, 3, 3, dtype=torch.float64)
    out = ops.conv2d(nhwc(x).float().to(DEV), w.float().to(DEV), None, 1, 1, None)
    close(nchw(out), F.conv2d(x, w, None, padding=1))


@pytest.mark.parametrize('M,N,K,act', [(64, 128, 256, None), (64, 512, 16384, None), (320, 32768, 64, None),
                                        (320, 256, 1024, 'elu'), (7, 12, 33, None), (64, 512, 320, None),
                                        (448, 256, 4096, 'relu')])
def test_linear(M, N, K, act):
    ops = _ops()
    torch.manual_seed(4)
    x = torch.randn(M, K, dtype=torch.float64)
    w = torch.randn(N, K, dtype=torch.float64) / K ** 0.5
    b = torch.randn(N, dtype=torch.float64)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = {'elu': F.elu, 'relu': F.relu, None: lambda t: t}[act](xr @ wr.t() + br)
    xg, wg, bg = (t.float().to(DEV).requires_grad_(True) for t in (x, w, b))
    out = ops.linear(xg, wg, bg, act)
    close(out, ref, name='fwd')
    gr = grads([ref], [xr, wr, br])
    gg = grads([out], [xg, wg, bg])
    for a, b_, n in zip(gg, gr, ('dx', 'dw', 'db')):
        close(a, b_, name=n)


def _gate_ref(y, C, hn, gn):
    return hn(y[:, :C]) * torch.sigmoid(gn(y[:, C:]))


@pytest.mark.parametrize('N,C,H', [(4, 32, 16), (6, 64, 8), (3, 128, 4)])
def test_batchnorm_gate(N, C, H):
    ops = _ops()
    torch.manual_seed(5)
    y = torch.randn(N, 2 * C, H, H, dtype=torch.float64) * 2 + 0.5
    hn = torch.nn.BatchNorm2d(C).double()
    gn = torch.nn.BatchNorm2d(C).double()
    for bn in (hn, gn):
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.uniform_(-0.5, 0.5)
    yr = y.clone().requires_grad_(True)
    ref = _gate_ref(yr, C, hn, gn)
    params = [hn.weight, hn.bias, gn.weight, gn.bias]
    gr = grads([ref], [yr] + params)
    yg = nhwc(y).float().to(DEV).requires_grad_(True)
    pg = [p.detach().float().to(DEV).requires_grad_(True) for p in params]
    bufs = [torch.zeros(C, device=DEV), torch.ones(C, device=DEV), torch.zeros(C, device=DEV), torch.ones(C, device=DEV)]
    out = ops.norm_post(yg, pg[0], pg[1], pg[2], pg[3], *bufs, mode=ops.NORM_BATCH, post=ops.POST_GATE, training=True)
    close(nchw(out), ref, name='fwd')
    gg = grads([nchw(out)], [yg] + pg)
    close(nchw(gg[0]), gr[0], name='dy', rtol=5e-4)
    for a, b, n in zip(gg[1:], gr[1:], ('dgh', 'dbh', 'dgg', 'dbg')):
        close(a, b, name=n, rtol=5e-4)
    close(bufs[0], hn.running_mean, name='rm_h', rtol=1e-5)
    close(bufs[1], hn.running_var, name='rv_h', rtol=1e-5)
    close(bufs[2], gn.running_mean, name='rm_g', rtol=1e-5)
    close(bufs[3], gn.running_var, name='rv_g', rtol=1e-5)
    # eval mode uses the running statistics
    hn.eval(); gn.eval()
    ref_e = _gate_ref(y, C, hn, gn)
    out_e = ops.norm_post(yg.detach(), pg[0], pg[1], pg[2], pg[3], *bufs, mode=ops.NORM_BATCH, post=ops.POST_GATE,
                          training=False)
    close(nchw(out_e), ref_e, name='eval')


def test_gate_without_norm():
    ops = _ops()
    torch.manual_seed(6)
    y = torch.randn(5, 64, 3, 3, dtype=torch.float64)
    yr = y.clone().requires_grad_(True)
    ref = yr[:, :32] * torch.sigmoid(yr[:, 32:])
    yg = nhwc(y).float().to(DEV).requires_grad_(True)
    out = ops.norm_post(yg, mode=ops.NORM_NONE, post=ops.POST_GATE)
    close(nchw(out), ref)
    close(nchw(grads([nchw(out)], [yg])[0]), grads([ref], [yr])[0], name='dy')


@pytest.mark.parametrize('mode,N,C,H', [('gn', 3, 64, 16), ('gn', 2, 128, 4), ('in', 3, 32, 16), ('in', 2, 64, 8)])
def test_group_instance_norm_relu(mode, N, C, H):
    ops = _ops()
    torch.manual_seed(7)
    y = torch.randn(N, C, H, H, dtype=torch.float64) * 1.5 + 0.3
    norm = (torch.nn.GroupNorm(8, C) if mode == 'gn' else torch.nn.InstanceNorm2d(C, affine=True)).double()
    norm.weight.data.uniform_(0.5, 1.5)
    norm.bias.data.uniform_(-0.5, 0.5)
    yr = y.clone().requires_grad_(True)
    ref = F.relu(norm(yr))
    gr = grads([ref], [yr, norm.weight, norm.bias])
    yg = nhwc(y).float().to(DEV).requires_grad_(True)
    wg = norm.weight.detach().float().to(DEV).requires_grad_(True)
    bg = norm.bias.detach().float().to(DEV).requires_grad_(True)
    out = ops.norm_post(yg, wg, bg, mode=ops.NORM_GROUP if mode == 'gn' else ops.NORM_INSTANCE, post=ops.POST_RELU, groups=8)
    close(nchw(out), ref, name='fwd')
    gg = grads([nchw(out)], [yg, wg, bg])
    close(nchw(gg[0]), gr[0], name='dy', rtol=5e-4)
    close(gg[1], gr[1], name='dgamma', rtol=5e-4)
    close(gg[2], gr[2], name='dbeta', rtol=5e-4)


@pytest.mark.parametrize('K,nl', [(5, 5), (7, 6), (2, 2), (3, 2)])
def test_sbp_scan(K, nl):
    ops = _ops()
    torch.manual_seed(8)
    logits = torch.randn(nl, 3, 1, 8, 8, dtype=torch.float64) * 3
    lr = logits.clone().requires_grad_(True)
    log_s = [torch.zeros_like(lr[0])]
    log_m = []
    for k in range(K - 1):
        log_m.append(log_s[-1] + F.logsigmoid(lr[k]))
        log_s.append(log_s[-1] + F.logsigmoid(-lr[k]))
    log_m.append(log_s[-1])
    if nl == K:
        log_s.append(log_s[-1] + F.logsigmoid(-lr[K - 1]))
    ref_m, ref_s = torch.stack(log_m), torch.stack(log_s)
    lg = logits.float().to(DEV).requires_grad_(True)
    m, s = ops.sbp_scan(lg, K)
    close(m, ref_m, atol=1e-5, name='log_m')
    close(s, ref_s, atol=1e-5, name='log_s')
    assert (m.exp().sum(0) - 1).abs().max().item() < 1e-4      # check_log_masks invariant
    close(grads([m], [lg])[0], grads([ref_m], [lr])[0], name='dlogits')


def test_comp_pack_and_layout():
    ops = _ops()
    torch.manual_seed(9)
    K, B, H = 3, 2, 8
    x = torch.rand(B, 3, H, H)
    lm = torch.randn(K, B, 1, H, H)
    lmg = lm.to(DEV).requires_grad_(True)
    out = ops.comp_pack(x.to(DEV), lmg)
    ref = torch.cat([lm.reshape(K * B, 1, H, H), x.repeat(K, 1, 1, 1)], dim=1)
    close(nchw(out), ref, atol=0, rtol=0)
    out32 = ops.comp_pack(x.to(DEV), lmg, 32)
    close(nchw(out32)[:, :4], ref, atol=0, rtol=0)
    assert out32[..., 4:].abs().max().item() == 0
    xp = ops.to_nhwc_padded(x.to(DEV), 32)
    close(xp[..., :3], nhwc(x), atol=0, rtol=0)
    assert xp[..., 3:].abs().max().item() == 0
    g = grads([out], [lmg])[0]
    assert g.shape == lm.shape
    t = torch.randn(3, 5, 4, 6)
    close(ops.to_nhwc(t.to(DEV)), nhwc(t), atol=0, rtol=0)
    close(ops.to_nchw(nhwc(t).to(DEV)), t, atol=0, rtol=0)


@pytest.mark.parametrize('act', [None, 'elu', 'relu'])
def test_bcast_add_act(act):
    ops = _ops()
    torch.manual_seed(10)
    a = torch.randn(6, 32, dtype=torch.float64)
    m = torch.randn(49, 32, dtype=torch.float64)
    ar, mr = a.clone().requires_grad_(True), m.clone().requires_grad_(True)
    ref = {'elu': F.elu, 'relu': F.relu, None: lambda t: t}[act](ar[:, None, :] + mr[None])
    ag, mg = a.float().to(DEV).requires_grad_(True), m.float().to(DEV).requires_grad_(True)
    out = ops.bcast_add_act(ag, mg, act)
    close(out, ref, name='fwd')
    for x_, y_, n in zip(grads([out], [ag, mg]), grads([ref], [ar, mr]), ('da', 'dm')):
        close(x_, y_, name=n)


@pytest.mark.parametrize('Cin,nout,nsig', [(32, 3, 3), (32, 4, 0), (32, 1, 0), (64, 4, 0)])
def test_out1x1(Cin, nout, nsig):
    ops = _ops()
    torch.manual_seed(11)
    h = torch.randn(3, Cin, 8, 8, dtype=torch.float64)
    w = torch.randn(nout, Cin, 1, 1, dtype=torch.float64) * 0.3
    b = torch.randn(nout, dtype=torch.float64)
    hr, wr, br = (t.clone().requires_grad_(True) for t in (h, w, b))
    ref = F.conv2d(hr, wr, br)
    if nsig:
        ref = torch.cat([torch.sigmoid(ref[:, :nsig]), ref[:, nsig:]], 1)
    hg = nhwc(h).float().to(DEV).requires_grad_(True)
    wg, bg = w.float().to(DEV).requires_grad_(True), b.float().to(DEV).requires_grad_(True)
    out = ops.out1x1(hg, wg, bg, nsig)
    close(out, ref, name='fwd')
    gr = grads([ref], [hr, wr, br])
    gg = grads([out], [hg, wg, bg])
    close(nchw(gg[0]), gr[0], name='dh')
    close(gg[1], gr[1], name='dw')
    close(gg[2], gr[2], name='db')


@pytest.mark.parametrize('K,B,H,softmax', [(5, 3, 16, False), (7, 2, 8, True), (1, 2, 8, False), (11, 2, 8, True)])
def test_mixture_nll(K, B, H, softmax):
    from oracle import functional as O
    ops = _ops()
    torch.manual_seed(12)
    x = torch.rand(B, 3, H, H, dtype=torch.float64)
    xr = torch.rand(K, B, 3, H, H, dtype=torch.float64)
    lm = torch.randn(K, B, 1, H, H, dtype=torch.float64) * 2
    std = torch.full((K,), 0.7, dtype=torch.float64)
    std[0] = 0.5
    xrr, lmr = xr.clone().requires_grad_(True), lm.clone().requires_grad_(True)
    logm = F.log_softmax(lmr, dim=0) if softmax else lmr
    ref_err = O.mixture_nll(x, list(logm.unbind(0)), list(xrr.unbind(0)), std)
    ref_recon = (logm.exp() * xrr).sum(0)
    xg = x.float().to(DEV)
    xrg, lmg = xr.float().to(DEV).requires_grad_(True), lm.float().to(DEV).requires_grad_(True)
    err, recon, lm_out = ops.mixture_nll(xg, xrg, lmg, std.float().to(DEV), softmax)
    close(err, ref_err, rtol=2e-5, name='err')
    close(recon, ref_recon, rtol=1e-4, name='recon')
    if softmax:
        close(lm_out, logm, atol=1e-5, name='log_softmax')
    gr = grads([ref_err], [xrr, lmr])
    gg = grads([err], [xrg, lmg])
    close(gg[0], gr[0], name='dxr', rtol=5e-4)
    close(gg[1], gr[1], name='dlm', rtol=5e-4)


@pytest.mark.parametrize('kind,case', [('conv', c) for c in CONV_CASES[1:3] + CONV_CASES[4:6] + CONV_CASES[7:8]]
                         + [('convT', c) for c in CONVT_CASES[:3]])
def test_ops_tf32(kind, case):
    """Same operators through the tcgen05 path (forward + data gradient on tensor cores): TF32 tolerance."""
    ops = _ops()
    ops.set_precision('tf32')
    N, H, W, Ci, Co, R, s, p, act = case
    torch.manual_seed(21)
    x = torch.randn(N, Ci, H, W, dtype=torch.float64)
    w = torch.randn((Co, Ci, R, R) if kind == 'conv' else (Ci, Co, R, R), dtype=torch.float64) * 0.1
    b = torch.randn(Co, dtype=torch.float64)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    if kind == 'conv':
        ref = F.conv2d(xr, wr, br, stride=s, padding=p)
    else:
        ref = F.conv_transpose2d(xr, wr, br, stride=s, padding=p, output_padding=s - 1)
    ref = {'elu': F.elu, 'relu': F.relu, None: lambda t: t}[act](ref)
    xg = nhwc(x).float().to(DEV).requires_grad_(True)
    wg, bg = w.float().to(DEV).requires_grad_(True), b.float().to(DEV).requires_grad_(True)
    out = (ops.conv2d if kind == 'conv' else ops.conv_transpose2d)(xg, wg, bg, s, p, act)
    close(nchw(out), ref, name='fwd', rtol=3e-3)
    gr = grads([ref], [xr, wr, br])
    gg = grads([nchw(out)], [xg, wg, bg])
    close(nchw(gg[0]), gr[0], name='dx', rtol=3e-3)
    close(gg[1], gr[1], name='dw', rtol=3e-3)
    close(gg[2], gr[2], name='db', rtol=3e-3)
