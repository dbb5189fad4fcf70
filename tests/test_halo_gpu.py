"""GPU: the halo implicit-GEMM kernel (igemm_halo.cu) through the C ABI against float64 torch convolutions.
Operands are pre-rounded to TF32 so the tensor-core products are exact and only the fp32 accumulation order differs."""
import pytest
import torch
import torch.nn.functional as F

from test_halo_plan import CASES as PLAN_CASES
from test_tc_gpu import tf32_round

pytestmark = pytest.mark.gpu
DEV = 'cuda'

CASES = PLAN_CASES + [
    (0, 7, 70, 70, 32, 32, 3, 1, 0),
    (1, 5, 66, 66, 32, 32, 3, 1, 0),
    (1, 9, 32, 32, 32, 64, 5, 1, 2),
    (0, 4, 64, 64, 32, 64, 5, 1, 2),       # attention encoder layer (after channel padding)
    (0, 3, 32, 32, 64, 128, 5, 1, 2),
    (0, 3, 64, 64, 128, 64, 3, 1, 1),      # UNet up block: 4 channel blocks
    (0, 2, 64, 64, 64, 128, 1, 1, 0),      # 1x1 head
    (1, 3, 32, 32, 64, 64, 5, 2, 2),       # V2 decoder
    (0, 2, 128, 128, 64, 32, 3, 1, 1),     # MONet-128 UNet up block
    (0, 37, 16, 16, 64, 64, 3, 1, 1),      # odd image count with several images per CTA
]


def run(case, act=0, bias=True, halo=True):
    from genesis_b200 import _lib
    mode, N, H, W, Ci, Co, R, s, p = case
    torch.manual_seed(1)
    x = tf32_round(torch.randn(N, Ci, H, W))
    w = tf32_round(torch.randn((Co, Ci, R, R) if mode == 0 else (Ci, Co, R, R)) * 0.1)
    b = torch.randn(Co) if bias else None
    bd = b.double() if bias else None
    if mode == 0:
        ref = F.conv2d(x.double(), w.double(), bd, stride=s, padding=p)
        wp = w.permute(2, 3, 0, 1).reshape(R * R, Co, Ci)
    else:
        ref = F.conv_transpose2d(x.double(), w.double(), bd, stride=s, padding=p, output_padding=s - 1)
        wp = w.permute(2, 3, 1, 0).reshape(R * R, Co, Ci)
    if act == 1:
        ref = F.relu(ref)
    elif act == 2:
        ref = F.elu(ref)
    Ho, Wo = ref.shape[2], ref.shape[3]
    lib = _lib.lib()
    prev = lib.query('g2_conv_halo_enable', 1 if halo else 0)
    try:
        if halo:
            assert lib.query('g2_conv_halo_supported', N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode) == 1
        xg = x.permute(0, 2, 3, 1).contiguous().to(DEV)
        out = torch.full((N, Ho, Wo, Co), float('nan'), device=DEV)
        name = 'g2_conv_halo_tf32' if halo else 'g2_conv_igemm_tf32'
        _lib.call(name, xg, wp.contiguous().to(DEV), b.to(DEV) if bias else None, out, N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, act)
        torch.cuda.synchronize()
    finally:
        lib.query('g2_conv_halo_enable', prev)
    got = out.permute(0, 3, 1, 2).double().cpu()
    assert torch.isfinite(got).all(), 'unwritten / non-finite outputs'
    return (got - ref).abs().max().item() / ref.abs().max().item(), got


@pytest.mark.parametrize('case', CASES)
def test_halo_exact_on_prerounded_operands(case):
    err, _ = run(case)
    assert err < 2e-5, err


@pytest.mark.parametrize('act,bias', [(1, True), (2, True), (0, False), (1, False)])
def test_halo_epilogue_variants(act, bias):
    for case in (CASES[0], CASES[3], CASES[9]):
        err, _ = run(case, act=act, bias=bias)
        assert err < 2e-5, (case, err)


def test_halo_agrees_with_the_tile_kernel():
    for case in (CASES[1], CASES[4], CASES[10]):
        _, a = run(case, halo=True)
        _, b = run(case, halo=False)
        assert (a - b).abs().max().item() < 1e-4 * b.abs().max().item()


def test_halo_routing_switch():
    from genesis_b200 import _lib
    lib = _lib.lib()
    prev = lib.query('g2_conv_halo_enable', 0)
    assert lib.query('g2_conv_halo_supported', 2, 72, 72, 32, 70, 70, 32, 3, 3, 1, 0, 0) == 0
    lib.query('g2_conv_halo_enable', 1)
    assert lib.query('g2_conv_halo_supported', 2, 72, 72, 32, 70, 70, 32, 3, 3, 1, 0, 0) == 1
    assert lib.query('g2_conv_halo_supported', 2, 64, 64, 32, 32, 32, 64, 5, 5, 2, 2, 0) == 0     # stride-2 conv: parity-plane kernel
    lib.query('g2_conv_halo_enable', prev)
