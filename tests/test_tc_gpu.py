"""GPU: the tcgen05 / TMA TF32 implicit-GEMM kernel, called through the C ABI, against float64 torch
convolutions.  With operands pre-rounded to TF32 the tensor-core products are exact, so the comparison is
tight (fp32 accumulation order only); with raw fp32 operands the error is the TF32 rounding (~1e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def tf32_round(t):
    """Round-to-nearest-even to 10 mantissa bits (what the tensor map's TFLOAT32 load produces)."""
    i = t.float().contiguous().view(torch.int32)
    r = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return r.view(torch.float32)


def run_conv(mode, N, H, W, Ci, Co, R, s, p, preround, seed=0):
    from genesis_b200 import _lib
    torch.manual_seed(seed)
    x = torch.randn(N, Ci, H, W)
    w = torch.randn((Co, Ci, R, R) if mode == 0 else (Ci, Co, R, R)) * 0.1
    b = torch.randn(Co)
    if preround:
        x, w = tf32_round(x), tf32_round(w)
    if mode == 0:
        ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p)
        wp = w.permute(2, 3, 0, 1).reshape(R * R, Co, Ci)
    else:
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=s, padding=p, output_padding=s - 1)
        wp = w.permute(2, 3, 1, 0).reshape(R * R, Co, Ci)
    Ho, Wo = ref.shape[2], ref.shape[3]
    lib = _lib.lib()
    assert lib.query('g2_conv_tf32_supported', N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode) == 1
    xg = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.full((N, Ho, Wo, Co), float('nan'), device=DEV)
    _lib.call('g2_conv_igemm_tf32', xg, wp.contiguous().to(DEV), b.to(DEV), out, N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, 0)
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).double().cpu()
    assert torch.isfinite(got).all(), 'unwritten / non-finite outputs'
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    return err


CASES = [  # mode, N, H, W, Ci, Co, R, stride, pad
    (0, 2, 16, 16, 32, 64, 5, 1, 2),
    (0, 2, 64, 64, 32, 64, 5, 1, 2),
    (0, 2, 32, 32, 64, 128, 5, 1, 2),
    (0, 3, 16, 16, 64, 128, 5, 1, 2),
    (0, 2, 64, 64, 64, 64, 3, 1, 1),       # UNet block
    (0, 2, 128, 128, 32, 32, 3, 1, 1),     # MONet-128 UNet block (TW = 128)
    (0, 2, 70, 70, 32, 32, 3, 1, 0),       # broadcast decoder, flat tiles
    (0, 2, 68, 68, 32, 32, 3, 1, 0),
    (0, 2, 66, 66, 32, 32, 3, 1, 0),       # -> 64: 2-D tiles
    (0, 1, 136, 136, 32, 32, 3, 1, 0),     # MONet-128 broadcast decoder
    (0, 2, 64, 64, 32, 64, 5, 2, 2),       # stride-2 conv (parity planes)
    (0, 2, 32, 32, 64, 128, 5, 2, 2),
    (0, 2, 32, 32, 32, 64, 3, 2, 1),
    (1, 2, 16, 16, 64, 128, 5, 1, 2),      # conv-transpose s1
    (1, 2, 32, 32, 32, 64, 5, 1, 2),
    (1, 2, 16, 16, 64, 64, 5, 2, 2),       # conv-transpose s2 (sub-pixel classes)
    (1, 2, 32, 32, 32, 64, 5, 2, 2),
    (1, 2, 32, 32, 64, 64, 5, 2, 2),
    (1, 2, 32, 32, 32, 32, 3, 2, 1),
    (0, 5, 8, 8, 64, 64, 3, 1, 1),         # small maps: several images per 128-pixel tile
    (0, 11, 4, 4, 128, 128, 3, 1, 1),
    (0, 6, 16, 16, 32, 64, 3, 2, 1),       # -> 8x8
    (1, 7, 4, 4, 128, 64, 5, 2, 2),        # GENESIS-V2 decoder layer 1 (66 -> 128 padded channels), 4x4 class grid
    (1, 5, 8, 8, 64, 64, 5, 2, 2),
    (0, 3, 72, 72, 32, 64, 5, 1, 2),       # non power-of-two width with padding: 8x16 tiles
    (1, 2, 68, 68, 32, 32, 3, 1, 0),       # data-gradient of a VALID conv (70x70 output)
]


@pytest.mark.parametrize('case', CASES)
def test_conv_tf32_exact_on_prerounded_operands(case):
    err = run_conv(*case, preround=True)
    assert err < 2e-5, err


@pytest.mark.parametrize('case', CASES[:3] + CASES[10:11] + CASES[15:16])
def test_conv_tf32_precision_on_fp32_operands(case):
    err = run_conv(*case, preround=False)
    assert err < 3e-3, err


def test_unsupported_shapes_are_reported():
    from genesis_b200 import _lib
    lib = _lib.lib()
    assert lib.query('g2_conv_tf32_supported', 2, 64, 64, 3, 64, 64, 64, 5, 5, 1, 2, 0) == 0      # Ci = 3
    assert lib.query('g2_conv_tf32_supported', 2, 6, 6, 64, 6, 6, 64, 3, 3, 1, 1, 0) == 0         # 6x6 < 128 pixels, not 2^k
    assert lib.query('g2_conv_tf32_supported', 2, 2, 2, 64, 2, 2, 64, 3, 3, 1, 1, 0) == 0


@pytest.mark.parametrize('M,N,K', [(64, 512, 16384), (320, 32768, 64), (448, 256, 4096), (100, 128, 96), (320, 256, 1024)])
@pytest.mark.parametrize('act', [0, 2])
def test_gemm_tf32(M, N, K, act):
    """g2_gemm_tf32_ws (what ops.linear calls): exact operands -> float64 reference within fp32 accumulation error; the
    split-K shapes (few tiles, K >= 512) are reduced in a fixed order, so two runs are BITWISE equal (no float atomics);
    bias + activation fused in both the split and the unsplit form."""
    from genesis_b200 import _lib
    torch.manual_seed(1)
    a = tf32_round(torch.randn(M, K))
    w = tf32_round(torch.randn(N, K) / K ** 0.5)
    b = torch.randn(N)
    ref = a.double() @ w.double().t() + b.double()
    if act == 2:
        ref = torch.where(ref > 0, ref, torch.expm1(ref))
    ws_bytes = _lib.lib().query('g2_gemm_tf32_workspace', M, N, K)
    if (M, N, K) == (64, 512, 16384):
        assert ws_bytes > 0                                  # the long-reduction shape is split
    ws = torch.empty(max(ws_bytes // 4, 1), device=DEV) if ws_bytes else None
    ad, wd, bd = a.to(DEV), w.to(DEV), b.to(DEV)
    outs = []
    for _ in range(2):
        out = torch.full((M, N), float('nan'), device=DEV)
        _lib.call('g2_gemm_tf32_ws', ad, wd, bd, out, ws, M, N, K, act)
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    got = outs[0].double().cpu()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() / ref.abs().max().item() < 2e-5


def test_gemm_tf32_unsplit_entry():
    """g2_gemm_tf32 (no workspace -> never split): one fp32 accumulation chain over the whole reduction."""
    from genesis_b200 import _lib
    torch.manual_seed(1)
    M, N, K = 448, 256, 4096
    a = tf32_round(torch.randn(M, K))
    w = tf32_round(torch.randn(N, K) / K ** 0.5)
    b = torch.randn(N)
    ref = a.double() @ w.double().t() + b.double()
    out = torch.full((M, N), float('nan'), device=DEV)
    _lib.call('g2_gemm_tf32', a.to(DEV), w.to(DEV), b.to(DEV), out, M, N, K)
    torch.cuda.synchronize()
    assert (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item() < 5e-5


WGRAD_CASES = [  # kind, N, H, W, Ci, Co, R, stride, pad   (conv: x[N,H,W,Ci] -> y; convT: x[N,H,W,Ci] -> y upsampled)
    ('conv', 2, 16, 16, 32, 64, 5, 1, 2),
    ('conv', 3, 32, 32, 64, 128, 5, 1, 2),
    ('conv', 2, 64, 64, 32, 64, 5, 1, 2),
    ('conv', 2, 64, 64, 64, 64, 3, 1, 1),
    ('conv', 2, 70, 70, 32, 32, 3, 1, 0),      # VALID, 8x8 chunks with ragged edges
    ('conv', 2, 66, 66, 32, 32, 3, 1, 0),
    ('conv', 2, 64, 64, 32, 64, 5, 2, 2),      # stride 2: parity planes of x
    ('conv', 2, 32, 32, 32, 64, 3, 2, 1),
    ('convT', 2, 16, 16, 64, 128, 5, 1, 2),
    ('convT', 2, 16, 16, 64, 64, 5, 2, 2),     # stride 2: parity planes of dy
    ('convT', 2, 32, 32, 32, 64, 5, 2, 2),
    ('conv', 9, 4, 4, 128, 128, 3, 1, 1),     # small maps: several images per chunk
    ('conv', 5, 16, 16, 32, 64, 3, 2, 1),
    ('convT', 6, 4, 4, 128, 64, 5, 2, 2),    # GENESIS-V2 decoder layer 1 (66 -> 128 padded channels)
    ('conv', 3, 8, 8, 64, 64, 3, 1, 1),
]


@pytest.mark.parametrize('case', WGRAD_CASES)
def test_wgrad_tf32_exact_on_prerounded_operands(case):
    from genesis_b200 import _lib
    kind, N, H, W, Ci, Co, R, s, p = case
    torch.manual_seed(3)
    x = tf32_round(torch.randn(N, Ci, H, W))
    if kind == 'conv':
        w = torch.zeros(Co, Ci, R, R, dtype=torch.float64, requires_grad=True)
        y = F.conv2d(x.double(), w, None, stride=s, padding=p)
    else:
        w = torch.zeros(Ci, Co, R, R, dtype=torch.float64, requires_grad=True)
        y = F.conv_transpose2d(x.double(), w, None, stride=s, padding=p, output_padding=s - 1)
    dy = tf32_round(torch.randn(y.shape))
    (ref,) = torch.autograd.grad((y * dy.double()).sum(), [w])
    Ho, Wo = y.shape[2], y.shape[3]
    xg = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    dyg = dy.permute(0, 2, 3, 1).contiguous().to(DEV)
    lib = _lib.lib()
    if kind == 'conv':      # g = x (a = Ci), t = dy (b = Co); packed [R,S,Ci,Co]
        dims = (N, H, W, Ci, Ho, Wo, Co)
        g, t, outT = xg, dyg, 0
        dwp = torch.full((R, R, Ci, Co), float('nan'), device=DEV)
    else:                   # g = dy (a = Co), t = x (b = Ci); packed [R,S,Ci,Co] via outT
        dims = (N, Ho, Wo, Co, H, W, Ci)
        g, t, outT = dyg, xg, 1
        dwp = torch.full((R, R, Ci, Co), float('nan'), device=DEV)
    ws_bytes = lib.query('g2_conv_wgrad_tf32_workspace', *dims, R, R, s)
    assert ws_bytes > 0
    ws = torch.empty(ws_bytes // 4, device=DEV)
    _lib.call('g2_conv_wgrad_tf32', g, t, dwp, ws, *dims, R, R, s, p, outT)
    torch.cuda.synchronize()
    got = (dwp.permute(3, 2, 0, 1) if kind == 'conv' else dwp.permute(2, 3, 0, 1)).double().cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err
