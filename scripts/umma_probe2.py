import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scripts.umma_probe import desc, idesc, probe, report, rng, N  # noqa (runs probe 1 as well)


def rows_image32(mat, mode):
    """[rows,32] float32 -> 128-byte rows; 32-byte chunks XORed with (row & 3) [mode 'x3'], ((row>>1)&3) ['x3s'], none."""
    rows = mat.shape[0]
    img = np.zeros((rows, 4, 8), np.float32)
    m = mat.reshape(rows, 4, 8)
    for r in range(rows):
        for c in range(4):
            x = {'x3': r & 3, 'x3s': (r >> 1) & 3, 'none': 0}[mode]
            img[r, c ^ x] = m[r, c]
    return img.reshape(-1)


print('---- MN-major with layout_type = 1 (SWIZZLE_128B_BASE32B)')
npix = 16
Ab = rng.randint(-3, 4, (4, npix, 32)).astype(np.float32)
Bb = rng.randint(-3, 4, (N // 32, npix, 32)).astype(np.float32)
Amat = Ab.transpose(0, 2, 1).reshape(128, npix)
Bmat = Bb.transpose(0, 2, 1).reshape(N, npix)
ref = Amat @ Bmat.T
tile = npix * 128
for mode in ('x3', 'x3s', 'none'):
    a_img = np.concatenate([rows_image32(Ab[j], mode) for j in range(4)])
    b_img = np.concatenate([rows_image32(Bb[j], mode) for j in range(N // 32)])
    for (nm, lbo, sbo) in (('LBO=tile,SBO=512', tile, 512), ('LBO=512,SBO=tile', 512, tile), ('LBO=tile,SBO=1024', tile, 1024)):
        got = probe(a_img, b_img, desc(lbo, sbo, 1), desc(lbo, sbo, 1), idesc(N, 1, 1), N, npix // 8, 1024, 1024)
        report('MN/MN layout1 img=%s %s' % (mode, nm), got, ref)
