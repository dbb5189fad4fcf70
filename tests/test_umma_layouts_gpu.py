"""GPU: pins the tcgen05 shared-memory descriptor conventions the production kernels rely on, measured with the
g2_debug_umma_probe entry point (one CTA, caller-built smem images, caller-built descriptors):
  1. K-major TF32 operands: 128-byte rows, 16-byte chunks XOR (row & 7) (SWIZZLE_128B, layout type 2),
     SBO = distance between 8-row groups; the swizzle is a function of the ABSOLUTE smem address, so a
     descriptor may start at any 128-byte row and SBO may be a non-dense pitch (halo tiles) with base_offset 0.
  2. MN-major TF32 operands only work with layout type 1 (SWIZZLE_128B_BASE32B): 128-byte rows whose 32-byte
     chunks are XORed with (row & 3), LBO = distance between 32-element atoms along M/N, SBO = 512 B.
     Layout type 2 with MN-major silently yields zeros."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
N = 64


def desc(lbo, sbo, layout):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def idesc(n, a_mn=0, b_mn=0):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | (8 << 24)


def image(mat, chunk_floats, mask):
    rows = mat.shape[0]
    nch = 32 // chunk_floats
    img = np.zeros((rows, nch, chunk_floats), np.float32)
    m = mat.reshape(rows, nch, chunk_floats)
    for r in range(rows):
        for c in range(nch):
            img[r, c ^ (r & mask)] = m[r, c]
    return img.reshape(-1)


def probe(a_img, b_img, ad, bd, idc, nk, a_k, b_k, a_off=0):
    from genesis_b200 import _lib
    D = torch.full((128, N), float('nan'), device='cuda')
    a, b = torch.from_numpy(a_img).cuda(), torch.from_numpy(b_img).cuda()
    _lib.probe().call('g2_debug_umma_probe', a, b, D, a.numel() * 4, b.numel() * 4, ad, bd, idc, N, nk, a_k, b_k, a_off, 0, 0)
    torch.cuda.synchronize()
    return D.cpu().numpy()


def test_k_major_sw128_and_address_based_swizzle():
    rng = np.random.RandomState(0)
    B = rng.randint(-3, 4, (N, 32)).astype(np.float32)
    A = rng.randint(-3, 4, (136, 32)).astype(np.float32)
    for shift in (0, 1, 3, 8):
        got = probe(image(A, 4, 7), image(B, 4, 7), desc(16, 1024, 2), desc(16, 1024, 2), idesc(N), 4, 32, 32, a_off=shift * 128)
        np.testing.assert_allclose(got, A[shift:shift + 128] @ B.T, atol=1e-3)
    # 16x8-pixel tile inside a 12-pixel-wide halo: SBO = halo pitch, start = shifted tap
    H = rng.randint(-3, 4, (240, 32)).astype(np.float32)
    for dh, dw in ((0, 0), (1, 2), (2, 4), (0, 3)):
        start = dh * 12 + dw
        rows = np.array([start + th * 12 + tw for th in range(16) for tw in range(8)])
        got = probe(image(H, 4, 7), image(B, 4, 7), desc(16, 12 * 128, 2), desc(16, 1024, 2), idesc(N), 4, 32, 32, a_off=start * 128)
        np.testing.assert_allclose(got, H[rows] @ B.T, atol=1e-3)


def test_mn_major_tf32_needs_the_32b_base_layout():
    rng = np.random.RandomState(1)
    npix = 16
    Ab = rng.randint(-3, 4, (4, npix, 32)).astype(np.float32)
    Bb = rng.randint(-3, 4, (N // 32, npix, 32)).astype(np.float32)
    ref = Ab.transpose(0, 2, 1).reshape(128, npix) @ Bb.transpose(0, 2, 1).reshape(N, npix).T
    tile = npix * 128
    a1 = np.concatenate([image(Ab[j], 8, 3) for j in range(4)])
    b1 = np.concatenate([image(Bb[j], 8, 3) for j in range(N // 32)])
    got = probe(a1, b1, desc(tile, 512, 1), desc(tile, 512, 1), idesc(N, 1, 1), npix // 8, 1024, 1024)
    np.testing.assert_allclose(got, ref, atol=1e-3)
    a2 = np.concatenate([image(Ab[j], 4, 7) for j in range(4)])
    b2 = np.concatenate([image(Bb[j], 4, 7) for j in range(N // 32)])
    got = probe(a2, b2, desc(tile, 1024, 2), desc(tile, 1024, 2), idesc(N, 1, 1), npix // 8, 1024, 1024)
    assert np.count_nonzero(got) == 0          # documented behaviour: layout type 2 + MN-major TF32 -> zeros


def test_tf32_mma_operand_conversion_of_raw_fp32_bits():
    """What tcgen05.mma kind::tf32 does with fp32 bit patterns in shared memory whose low 13 mantissa bits are set (operands the
    TMA did not round): TRUNCATION (the low bits are ignored) or round-to-nearest.  The in-kernel 3xTF32 split relies on the
    answer (lo = x - hi must use the same hi the tensor core uses); recorded in profiles/r02_tf32_operand_conversion.txt."""
    ulp = 2.0 ** -10                                    # TF32 spacing in [1, 2)
    A = np.zeros((128, 32), np.float32)
    for r in range(128):
        k = r % 8
        A[r, 0] = (1.0 + k * ulp / 8.0) * (-1.0 if (r // 8) % 2 else 1.0)     # 1 + k/8 ulp, both signs
    B = np.zeros((N, 32), np.float32)
    B[0, 0] = 1.0
    got = probe(image(A, 4, 7), image(B, 4, 7), desc(16, 1024, 2), desc(16, 1024, 2), idesc(N), 4, 32, 32)[:, 0]
    trunc = np.sign(A[:, 0]) * 1.0
    rn = np.sign(A[:, 0]) * np.where((np.arange(128) % 8) >= 4, 1.0 + ulp, 1.0)
    is_trunc = np.array_equal(got, trunc)
    is_rn = np.allclose(got, rn, atol=0) or np.array_equal(np.where((np.arange(128) % 8) == 4, trunc, got),
                                                           np.where((np.arange(128) % 8) == 4, trunc, rn))
    print('TF32_OPERAND_CONVERSION', 'truncate' if is_trunc else 'round-to-nearest' if is_rn else 'other', got[:8].tolist())
    assert is_trunc or is_rn, got[:16]
