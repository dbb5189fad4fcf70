"""CPU: genesis_b200.ops autograd Functions on CPU tensors with their C-ABI calls routed to the emulated kernel sources
(tests/cuda_emu/emu_lib.py), against the plain-torch contract of each op (tests/cpu_ops_mock.py): argument marshalling, strides,
workspaces and saved tensors of the Functions themselves, which the kernel-level emulation tests do not see."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cuda_emu'))
import cpu_ops_mock as R  # noqa: E402
import emu_lib  # noqa: E402


@pytest.fixture
def ops(monkeypatch):
    from genesis_b200 import ops as _ops
    emu_lib.install(monkeypatch)
    prev = _ops.get_precision()
    _ops.set_precision('fp32')
    yield _ops
    _ops.set_precision(prev)


def grads(out, ins, w):
    return torch.autograd.grad((out * w).sum(), ins, allow_unused=True)


@pytest.mark.parametrize('mode,post', [(1, 0), (2, 1), (3, 1), (1, 2)], ids=['batch-gate', 'instance-relu', 'group-relu', 'batch-none'])
def test_norm_post(ops, mode, post):
    torch.manual_seed(mode * 3 + post)
    N, Hh, C = 3, 6, 16
    Cy = 2 * C if post == 0 else C
    y = torch.randn(N, Hh, Hh, Cy, requires_grad=True)
    half = C if post == 0 else Cy
    g0, b0 = (torch.rand(half) + 0.5).requires_grad_(True), torch.randn(half, requires_grad=True)
    g1 = (torch.rand(Cy - half) + 0.5).requires_grad_(True) if post == 0 else None
    b1 = torch.randn(Cy - half, requires_grad=True) if post == 0 else None
    rm = [torch.zeros(half), torch.ones(half), torch.zeros(Cy - half), torch.ones(Cy - half)]
    rm2 = [t.clone() for t in rm]
    kw = dict(mode=mode, post=post, groups=4, training=True)
    out = ops.norm_post(y, g0, b0, g1, b1, rm[0], rm[1], rm[2] if post == 0 else None, rm[3] if post == 0 else None, **kw)
    ref = R.norm_post(y, g0, b0, g1, b1, rm2[0], rm2[1], rm2[2] if post == 0 else None, rm2[3] if post == 0 else None, **kw)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
    if mode == 1:
        torch.testing.assert_close(rm[0], rm2[0], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(rm[1], rm2[1], rtol=1e-5, atol=1e-6)
    w = torch.randn_like(ref)
    ins = [t for t in (y, g0, b0, g1, b1) if t is not None]
    for a, b in zip(grads(out, ins, w), grads(ref, ins, w)):
        torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-4)


def test_norm_post_direct_accumulation(ops):
    """G2_NORM_DIRECT: in direct-gradient mode the finalize kernel adds the affine-parameter gradients to param.grad."""
    torch.manual_seed(0)
    N, Hh, C = 2, 5, 8
    y = torch.randn(N, Hh, Hh, 2 * C, requires_grad=True)
    ps = [torch.nn.Parameter(torch.rand(C) + 0.5), torch.nn.Parameter(torch.randn(C)), torch.nn.Parameter(torch.rand(C) + 0.5),
          torch.nn.Parameter(torch.randn(C))]
    base = [torch.randn(C) for _ in ps]
    for p_, b_ in zip(ps, base):
        p_.grad = b_.clone()
    w = torch.randn(N, Hh, Hh, C)
    ref = R.norm_post(y, *[p_.detach().clone().requires_grad_(True) for p_ in ps], mode=2, post=0)
    ops.set_direct_grad(True)
    ops.set_norm_direct(True)
    try:
        out = ops.norm_post(y, *ps, mode=2, post=0)
        (out * w).sum().backward()
    finally:
        ops.set_direct_grad(False)
        ops.set_norm_direct(False)
    y2 = y.detach().clone().requires_grad_(True)
    qs = [p_.detach().clone().requires_grad_(True) for p_ in ps]
    gref = torch.autograd.grad((R.norm_post(y2, *qs, mode=2, post=0) * w).sum(), [y2] + qs)
    torch.testing.assert_close(y.grad, gref[0], rtol=2e-3, atol=2e-4)
    for p_, b_, g_ in zip(ps, base, gref[1:]):
        torch.testing.assert_close(p_.grad, b_ + g_, rtol=2e-3, atol=2e-4)        # added to what was there


@pytest.mark.parametrize('kernel', ['gaussian', 'laplacian', 'epanechnikov'])
def test_icsbp_function(ops, kernel):
    torch.manual_seed(2)
    B, S, K = 2, 16, 4
    colour = (0.4 * torch.randn(B, S, S, 8)).requires_grad_(True)
    u = torch.rand(B, 1, S, S)
    ls = torch.tensor(0.4).log().requires_grad_(True)
    log_m, log_s, idx = ops.icsbp(colour, u, ls, K, kernel)
    rm, rs, ridx = R.icsbp(colour, u, ls, K, kernel)
    assert torch.equal(idx.long(), ridx.long())
    torch.testing.assert_close(log_m, rm, rtol=1e-4, atol=1e-4)
    w = torch.randn_like(rm)
    for a, c in zip(grads(log_m, [colour, ls], w), grads(rm, [colour, ls], w)):
        torch.testing.assert_close(a, c.to(a.dtype), rtol=2e-3, atol=2e-4)


def test_mask_kl_function_plain_and_packed(ops):
    torch.manual_seed(3)
    K, B, Hh = 3, 2, 8
    lm = torch.log_softmax(torch.randn(K, B, 1, Hh, Hh), 0).requires_grad_(True)
    for cs, detach in ((4, True), (4, False), (1, False)):
        dec = torch.randn(K, B, cs, Hh, Hh, requires_grad=True)
        out = ops.mask_kl(lm, dec, detach)
        ref = R.mask_kl(lm, dec, detach)
        torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)
        g = torch.randn(B)
        for a, c in zip(grads(out, [lm, dec], g), grads(ref, [lm, dec], g)):
            if c is None:
                assert a is None or a.abs().max() == 0
            else:
                torch.testing.assert_close(a, c, rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize('direct', [False, True])
def test_gated_conv_bias_gradient_fused_into_the_norm_backward(ops, direct):
    """conv2d(bias_grad=False) + norm_post(conv_bias=b): the conv's bias gradient (column sums of dy) comes out of the norm
    backward-apply kernel -- returned to autograd, or accumulated into b.grad in direct-gradient mode."""
    torch.manual_seed(5)
    N, Hh, C = 2, 6, 8
    x = torch.randn(N, Hh, Hh, 4, requires_grad=True)
    w = (0.3 * torch.randn(2 * C, 4, 3, 3)).requires_grad_(True)
    b = torch.randn(2 * C, requires_grad=True)
    g0, b0 = (torch.rand(C) + 0.5).requires_grad_(True), torch.randn(C, requires_grad=True)
    g1, b1 = (torch.rand(C) + 0.5).requires_grad_(True), torch.randn(C, requires_grad=True)
    wgt = torch.randn(N, Hh, Hh, C)

    def fwd(mod, fused):
        y = mod.conv2d(x, w, b, 1, 1, bias_grad=not fused)
        return mod.norm_post(y, g0, b0, g1, b1, mode=2, post=0, conv_bias=b if fused else None)
    ref = torch.autograd.grad((fwd(R, False) * wgt).sum(), [x, w, b])
    if direct:
        b.grad = torch.full_like(b, 0.5)
        ops.set_direct_grad(True)
        try:
            out = fwd(ops, True)
            got = torch.autograd.grad((out * wgt).sum(), [x, w, b], allow_unused=True)
        finally:
            ops.set_direct_grad(False)
        assert got[2] is None
        torch.testing.assert_close(b.grad - 0.5, ref[2], rtol=2e-3, atol=2e-4)
    else:
        got = torch.autograd.grad((fwd(ops, True) * wgt).sum(), [x, w, b])
        torch.testing.assert_close(got[2], ref[2], rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(got[0], ref[0], rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(got[1], ref[1], rtol=2e-3, atol=2e-4)


def test_conv_activation_bias_gradient_fused(ops):
    """conv + bias + ELU: activation backward and bias gradient in one kernel (g2_act_bwd_bias_f32)."""
    torch.manual_seed(6)
    x = torch.randn(3, 7, 7, 4, requires_grad=True)
    w = (0.3 * torch.randn(8, 4, 3, 3)).requires_grad_(True)
    b = torch.randn(8, requires_grad=True)
    wgt = torch.randn(3, 5, 5, 8)
    ref = torch.autograd.grad((R.conv2d(x, w, b, 1, 0, 'elu') * wgt).sum(), [x, w, b])
    got = torch.autograd.grad((ops.conv2d(x, w, b, 1, 0, 'elu') * wgt).sum(), [x, w, b])
    for a, c in zip(got, ref):
        torch.testing.assert_close(a, c, rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize('transposed', [False, True])
@pytest.mark.parametrize('Hh', [16, 8])
def test_precise_conv_3xtf32(ops, transposed, Hh):
    """ops.precise(): the forward contraction as 3xTF32 -- (x_hi, w_hi) + (x_hi, w_lo) + (x_lo, w_hi) -- at fp32-level accuracy,
    against torch in float64, with the plain TF32 result of the same kernel as the contrast.  16x16 maps take the halo kernel's
    in-kernel split (the raw window is x_hi because the tensor core truncates; the epilogue warps rewrite it in place to x_lo
    between the passes; only the weight is split on the host); 8x8 maps (< 256 pixels: tile kernel) take the [hi|hi|lo]
    channel-concatenated form.  The emulation truncates MMA operands as the hardware was measured to do."""
    ops.set_precision('tf32')
    torch.manual_seed(8)
    N, Ci, Co = 2, 32, 32
    x = torch.randn(N, Hh, Hh, Ci, requires_grad=True)
    w = (0.2 * torch.randn(Ci, Co, 3, 3) if transposed else 0.2 * torch.randn(Co, Ci, 3, 3)).requires_grad_(True)
    b = torch.randn(Co, requires_grad=True)
    fn_o = ops.conv_transpose2d if transposed else ops.conv2d
    fn_r = R.conv_transpose2d if transposed else R.conv2d
    calls = []
    import genesis_b200._lib as L
    orig = L._LIB.call
    L._LIB.call = lambda name, *a: (calls.append((name, a)), orig(name, *a))[1]
    try:
        with ops.precise():
            out = fn_o(x, w, b, 1, 1, 'relu')
    finally:
        L._LIB.call = orig
    names = [n for n, _ in calls]
    if Hh == 16:
        assert names.count('g2_conv_halo_x3_tf32') == 1 and names.count('g2_split_tf32_f32') == 2      # only w -> hi, lo
    else:
        convs = [a for n, a in calls if n in ('g2_conv_halo_tf32', 'g2_conv_igemm_tf32')]
        assert len(convs) == 1 and convs[0][7] == 3 * Ci and names.count('g2_split_tf32_f32') == 3
    plain = fn_o(x, w, b, 1, 1, 'relu')
    ref = fn_r(x.double(), w.double(), b.double(), 1, 1, 'relu')
    e_x3 = (out.double() - ref).abs().max().item()
    e_plain = (plain.double() - ref).abs().max().item()
    assert e_x3 < 2e-5 and e_plain > 20 * e_x3, (e_x3, e_plain)        # 3xTF32 is >= 20x closer than plain TF32
    g = torch.randn_like(out)
    ref32 = fn_r(x, w, b, 1, 1, 'relu')
    for a, c in zip(grads(out, [x, w, b], g), grads(ref32, [x, w, b], g)):
        torch.testing.assert_close(a, c, rtol=2e-2, atol=2e-2)            # backward = plain TF32 tensor-core path


def test_split_tf32_kernel_is_exact(ops):
    """hi + lo == x exactly, hi has a 10-bit mantissa, the three layouts of g2_split_tf32_f32."""
    import genesis_b200._lib as L
    torch.manual_seed(9)
    x = torch.randn(6, 8) * torch.tensor([1e-6, 1e-3, 1.0, 1e3, 1e6, 1.0, 1.0, 1.0])
    out3, out2, hi, lo = torch.empty(6, 24), torch.empty(6, 16), torch.empty(6, 8), torch.empty(6, 8)
    for o, mode in ((out3, 0), (out2, 1), (hi, 2), (lo, 3)):
        L.call('g2_split_tf32_f32', x, o, 6, 8, mode)
    assert torch.equal(hi + lo, x)
    assert (hi.view(torch.int32) & 0x1FFF).abs().max().item() == 0
    assert ((x - hi).abs() <= x.abs() * 2.0 ** -11).all()
    assert torch.equal(out3, torch.cat([hi, hi, lo], 1)) and torch.equal(out2, torch.cat([hi, lo], 1))


@pytest.mark.parametrize('Cin,nout,nsig,NP', [(32, 3, 3, (3, 9, 7)), (32, 4, 0, (2, 20, 33)), (64, 1, 0, (2, 5, 6))])
def test_out1x1_head_kernels(ops, Cin, nout, nsig, NP):
    """The 1x1 output head (out1x1_fwd / out1x1_bwd with cooperative row stores / head_wgrad with four rows per load) against
    torch, at pixel counts that are not multiples of the kernels' 16-pixel strides."""
    import torch.nn.functional as F
    torch.manual_seed(21)
    N, Hh, Ww = NP
    h = torch.randn(N, Hh, Ww, Cin, requires_grad=True)
    w = (0.3 * torch.randn(nout, Cin, 1, 1)).requires_grad_(True)
    b = torch.randn(nout, requires_grad=True)
    out = ops.out1x1(h, w, b, nsig)
    ref = F.conv2d(h.permute(0, 3, 1, 2), w, b)
    if nsig:
        ref = torch.cat([torch.sigmoid(ref[:, :nsig]), ref[:, nsig:]], 1)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
    g = torch.randn_like(ref)
    for a, c in zip(grads(out, [h, w, b], g), grads(ref, [h, w, b], g)):
        torch.testing.assert_close(a, c, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize('N,J', [(5, 8), (16, 40), (37, 12)])
def test_sum_dim0_kernel(ops, N, J):
    import genesis_b200._lib as L
    torch.manual_seed(22)
    x = torch.randn(N, J)
    out = torch.empty(J)
    L.call('g2_sum_dim0_f32', x, out, N, J)
    torch.testing.assert_close(out, x.sum(0), rtol=1e-5, atol=1e-5)
