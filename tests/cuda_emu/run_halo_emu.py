"""TEST INFRASTRUCTURE: run g2_conv_halo_tf32 of genesis_b200/csrc/igemm_halo.cu under the CPU emulation on one case and compare
with torch (integer-valued data: every product and sum is exact in fp32, so the comparison is exact).

    python tests/cuda_emu/run_halo_emu.py "mode N H W Ci Co R stride pad act; ..." [persistent]

Runs in its own process: the kernel's environment switches are read once per process, and a protocol deadlock aborts."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build_emu  # noqa: E402


def run_case(lib, case):
    mode, N, H, W, Ci, Co, R, s, p, act = case
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.randint(-3, 4, (N, Ci, H, W)).astype(np.float32))
    w = torch.from_numpy(rng.randint(-2, 3, (Co, Ci, R, R) if mode == 0 else (Ci, Co, R, R)).astype(np.float32))
    b = torch.from_numpy(rng.randint(-4, 5, (Co,)).astype(np.float32))
    if mode == 0:
        ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p)
        wp = w.permute(2, 3, 0, 1).reshape(R * R, Co, Ci)
    else:
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=s, padding=p, output_padding=s - 1)
        wp = w.permute(2, 3, 1, 0).reshape(R * R, Co, Ci)
    if act == 1:
        ref = F.relu(ref)
    Ho, Wo = ref.shape[2], ref.shape[3]
    xg = np.ascontiguousarray(x.permute(0, 2, 3, 1).numpy())
    wpn = np.ascontiguousarray(wp.numpy())
    out = np.full((N, Ho, Wo, Co), np.nan, np.float32)
    P = ctypes.c_void_p
    assert lib.g2_conv_halo_supported(N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode) == 1
    rc = lib.g2_conv_halo_tf32(xg.ctypes.data_as(P), wpn.ctypes.data_as(P), b.numpy().ctypes.data_as(P), out.ctypes.data_as(P),
                               N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, act, None)
    assert rc == 0, rc
    got = torch.from_numpy(out).permute(0, 3, 1, 2).double()
    assert torch.isfinite(got).all(), 'unwritten outputs'
    err = (got - ref).abs().max().item()
    print('max abs err', err)
    assert err == 0.0
    print('OK', case)


def run_case_x3(lib, case):
    """g2_conv_halo_x3_tf32 (3xTF32, weight pack [w_hi | w_lo]) on operands that NEED their low parts (13 significant bits):
    against float64 the result must be at the 3xTF32 level (1e-6), far from what a missing / misplaced low-part pass leaves (3e-4)."""
    mode, N, H, W, Ci, Co, R, s, p, act = case
    rng = np.random.RandomState(1)
    x = torch.from_numpy((rng.randint(-3, 4, (N, Ci, H, W)) + rng.randint(0, 4, (N, Ci, H, W)) * 2.0 ** -11).astype(np.float32))
    wshape = (Co, Ci, R, R) if mode == 0 else (Ci, Co, R, R)
    w = torch.from_numpy((rng.randint(-2, 3, wshape) + rng.randint(0, 4, wshape) * 2.0 ** -11).astype(np.float32))
    b = torch.from_numpy(rng.randint(-4, 5, (Co,)).astype(np.float32))
    if mode == 0:
        ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p)
        wp = w.permute(2, 3, 0, 1).reshape(R * R, Co, Ci)
    else:
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=s, padding=p, output_padding=s - 1)
        wp = w.permute(2, 3, 1, 0).reshape(R * R, Co, Ci)
    if act == 1:
        ref = F.relu(ref)
    w_hi = ((wp.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    wpn = np.ascontiguousarray(torch.cat([w_hi, wp - w_hi], dim=2).numpy())          # [R*R][Co][2*Ci]
    Ho, Wo = ref.shape[2], ref.shape[3]
    xg = np.ascontiguousarray(x.permute(0, 2, 3, 1).numpy())
    out = np.full((N, Ho, Wo, Co), np.nan, np.float32)
    P = ctypes.c_void_p
    rc = lib.g2_conv_halo_x3_tf32(xg.ctypes.data_as(P), wpn.ctypes.data_as(P), b.numpy().ctypes.data_as(P), out.ctypes.data_as(P),
                                  N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, act, None)
    assert rc == 0, rc
    got = torch.from_numpy(out).permute(0, 3, 1, 2).double()
    assert torch.isfinite(got).all(), 'unwritten outputs'
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print('max rel err', err)
    assert err < 5e-6, err
    print('OK', case)


def main():
    mode = sys.argv[2] if len(sys.argv) > 2 else ''
    if mode in ('persistent', 'x3'):
        os.environ['G2_HALO_PERSISTENT'] = '1'
    lib = ctypes.CDLL(build_emu.build('igemm_halo.cu'))
    for c in sys.argv[1].split(';'):
        if c.strip():
            (run_case_x3 if mode.startswith('x3') else run_case)(lib, tuple(int(a) for a in c.split()))


if __name__ == '__main__':
    main()
