"""Run UMMA descriptor experiments through g2_debug_umma_probe and print which hypotheses hold."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from genesis_b200 import _lib

DEV = 'cuda'


def desc(lbo, sbo, layout=2):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def idesc(N, a_mn=0, b_mn=0, M=128):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def rows_image(mat, row_offset=0, total_rows=None):
    """mat [rows, 32] float32 -> smem image with 128-byte rows, 16B chunks XOR-swizzled by (absolute row & 7)."""
    rows = mat.shape[0]
    total = total_rows or (rows + row_offset)
    img = np.zeros((total, 8, 4), np.float32)
    m = mat.reshape(rows, 8, 4)
    for r in range(rows):
        ar = r + row_offset
        for c in range(8):
            img[ar, c ^ (ar & 7)] = m[r, c]
    return img.reshape(-1)


def probe(a_img, b_img, ad, bd, idc, N, nk, a_k, b_k, a_off=0, b_off=0, auto=0):
    D = torch.full((128, N), float('nan'), device=DEV)
    a = torch.from_numpy(a_img).to(DEV); b = torch.from_numpy(b_img).to(DEV)
    _lib.probe().call('g2_debug_umma_probe', a, b, D, a.numel() * 4, b.numel() * 4, ad, bd, idc, N, nk, a_k, b_k, a_off, b_off, auto)
    torch.cuda.synchronize()
    return D.cpu().numpy()


def report(name, got, ref):
    ok = np.allclose(got, ref, atol=1e-3)
    nz = np.count_nonzero(got)
    print('%-50s %s  (nonzero %d/%d, max|got| %.1f, max|ref| %.1f)' % (name, 'OK' if ok else 'MISMATCH', nz, got.size, np.abs(got).max(), np.abs(ref).max()))
    return ok


rng = np.random.RandomState(0)
N = 64
# E1: K-major sanity
A = rng.randint(-3, 4, (128, 32)).astype(np.float32); B = rng.randint(-3, 4, (N, 32)).astype(np.float32)
got = probe(rows_image(A), rows_image(B), desc(16, 1024), desc(16, 1024), idesc(N), N, 4, 32, 32)
report('E1 K-major SW128 (production fwd config)', got, A @ B.T)

# E2: MN-major: A blocks [4][npix][32], B blocks [N/32][npix][32]
npix = 16
Ab = rng.randint(-3, 4, (4, npix, 32)).astype(np.float32)        # [blk][pixel][ch]
Bb = rng.randint(-3, 4, (N // 32, npix, 32)).astype(np.float32)
Amat = Ab.transpose(0, 2, 1).reshape(128, npix)                    # [m = blk*32+ch][k = pixel]
Bmat = Bb.transpose(0, 2, 1).reshape(N, npix)
ref = Amat @ Bmat.T
a_img = np.concatenate([rows_image(Ab[j]) for j in range(4)])
b_img = np.concatenate([rows_image(Bb[j]) for j in range(N // 32)])
tile = npix * 128
for (nm, lbo, sbo) in (('LBO=tile,SBO=1024', tile, 1024), ('LBO=1024,SBO=tile', 1024, tile)):
    got = probe(a_img, b_img, desc(lbo, sbo), desc(lbo, sbo), idesc(N, 1, 1), N, npix // 8, 1024, 1024)
    report('E2 MN-major A,B  ' + nm, got, ref)
# E2c: MN-major A with K-major B (B [N][npix] rows of... needs K=npix floats per row: use npix=32 -> 128B rows)
npix = 32
Ab = rng.randint(-3, 4, (4, npix, 32)).astype(np.float32)
Amat = Ab.transpose(0, 2, 1).reshape(128, npix)
Bk = rng.randint(-3, 4, (N, npix)).astype(np.float32)
a_img = np.concatenate([rows_image(Ab[j]) for j in range(4)])
tile = npix * 128
for (nm, lbo, sbo) in (('LBO=tile,SBO=1024', tile, 1024), ('LBO=1024,SBO=tile', 1024, tile)):
    got = probe(a_img, rows_image(Bk), desc(lbo, sbo), desc(16, 1024), idesc(N, 1, 0), N, npix // 8, 1024, 32)
    report('E2c MN-major A, K-major B  ' + nm, got, Amat @ Bk.T)
# E2d: K-major A, MN-major B
Ak = rng.randint(-3, 4, (128, npix)).astype(np.float32)
Bb = rng.randint(-3, 4, (N // 32, npix, 32)).astype(np.float32)
Bmat = Bb.transpose(0, 2, 1).reshape(N, npix)
b_img = np.concatenate([rows_image(Bb[j]) for j in range(N // 32)])
for (nm, lbo, sbo) in (('LBO=tile,SBO=1024', tile, 1024), ('LBO=1024,SBO=tile', 1024, tile)):
    got = probe(rows_image(Ak), b_img, desc(16, 1024), desc(lbo, sbo), idesc(N, 0, 1), N, npix // 8, 32, 1024)
    report('E2d K-major A, MN-major B  ' + nm, got, Ak @ Bmat.T)

# E3: K-major A with a row-shifted start (halo reuse): A has 128+8 rows, start at row j
Abig = rng.randint(-3, 4, (136, 32)).astype(np.float32)
B = rng.randint(-3, 4, (N, 32)).astype(np.float32)
for j in (1, 3, 8):
    for auto in (0, 1):
        got = probe(rows_image(Abig), rows_image(B), desc(16, 1024), desc(16, 1024), idesc(N), N, 4, 32, 32, a_off=j * 128, auto=auto)
        report('E3 K-major A shifted by %d rows, base_offset %s' % (j, 'auto' if auto else '0'), got, Abig[j:j + 128] @ B.T)
# E4: K-major A, 8-row groups with SBO = 12*128 (tile 16x8 inside a 12-wide halo), shifted start
halo_w = 12
H = rng.randint(-3, 4, (20 * halo_w, 32)).astype(np.float32)
for (dh, dw) in ((0, 0), (1, 2), (2, 4), (0, 3)):
    start = dh * halo_w + dw
    rows = np.array([start + th * halo_w + tw for th in range(16) for tw in range(8)])
    for auto in (0, 1):
        got = probe(rows_image(H), rows_image(B), desc(16, halo_w * 128), desc(16, 1024), idesc(N), N, 4, 32, 32, a_off=start * 128, auto=auto)
        report('E4 halo tile 16x8 pitch 12, shift (%d,%d), base_offset %s' % (dh, dw, 'auto' if auto else '0'), got, H[rows] @ B.T)
