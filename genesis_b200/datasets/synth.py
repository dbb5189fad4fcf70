"""Synthetic image batches shaped like the reference's datasets: the inputs of bench.py, of the
synthetic Forge data config (genesis_b200/datasets/synthetic_config.py) and of the parity tests.  Pure numpy, seeded, no
files needed (the real datasets cannot be downloaded here).

multid : uniform background + 1-4 filled shapes, colours from the 5-level-per-channel palette of
         scripts/generate_multid.py:32-34,55-70 (the dSprites masks themselves are not available).
stacks : ShapeStacks-like -- 2-colour vertical gradient + 2-6 stacked axis-aligned rectangles.
rooms  : GQN-rooms-like -- sky / wall / floor bands + 1-3 shaded blobs.
Returns float32 [B,3,H,W] in [0,1] and int64 instance labels [B,1,H,W]
(contract: datasets/multid_config.py:131-144)."""
import numpy as np


def _palette(rng):
    return rng.randint(0, 5, size=3).astype(np.float32) * (63.75 / 255.0)


def multid(batch, img=64, seed=1):
    rng = np.random.RandomState(seed)
    x = np.zeros((batch, 3, img, img), np.float32)
    inst = np.zeros((batch, 1, img, img), np.int64)
    yy, xx = np.mgrid[0:img, 0:img].astype(np.float32)
    for b in range(batch):
        x[b] = _palette(rng)[:, None, None]
        for obj in range(rng.randint(1, 5)):
            cy, cx = rng.uniform(0.15, 0.85, 2) * img
            r = rng.uniform(0.08, 0.22) * img
            kind = rng.randint(3)
            if kind == 0:
                m = (np.abs(yy - cy) < r) & (np.abs(xx - cx) < r)
            elif kind == 1:
                m = ((yy - cy) / r) ** 2 + ((xx - cx) / (0.7 * r)) ** 2 < 1
            else:
                m = (np.abs(yy - cy) + np.abs(xx - cx)) < 1.2 * r
            col = _palette(rng)
            x[b][:, m] = col[:, None]
            inst[b, 0][m] = obj + 1
    return x, inst


def stacks(batch, img=64, seed=1):
    rng = np.random.RandomState(seed)
    x = np.zeros((batch, 3, img, img), np.float32)
    inst = np.zeros((batch, 1, img, img), np.int64)
    ramp = np.linspace(0, 1, img, dtype=np.float32)[None, :, None]
    for b in range(batch):
        c0, c1 = rng.uniform(0.2, 0.9, (2, 3, 1, 1)).astype(np.float32)
        x[b] = c0 * (1 - ramp) + c1 * ramp
        top = img - 4
        cx = rng.uniform(0.35, 0.65) * img
        for obj in range(rng.randint(2, 7)):
            h = int(rng.uniform(0.08, 0.16) * img)
            w = int(rng.uniform(0.10, 0.30) * img)
            cx += rng.uniform(-0.06, 0.06) * img
            y0, y1 = max(top - h, 0), top
            x0, x1 = int(max(cx - w / 2, 0)), int(min(cx + w / 2, img))
            x[b, :, y0:y1, x0:x1] = rng.uniform(0, 1, (3, 1, 1))
            inst[b, 0, y0:y1, x0:x1] = obj + 1
            top = y0
    return x, inst


def rooms(batch, img=64, seed=1):
    rng = np.random.RandomState(seed)
    x = np.zeros((batch, 3, img, img), np.float32)
    inst = np.zeros((batch, 1, img, img), np.int64)
    yy, xx = np.mgrid[0:img, 0:img].astype(np.float32)
    for b in range(batch):
        h1, h2 = sorted(rng.randint(img // 6, 5 * img // 6, 2))
        for (a, e) in ((0, h1), (h1, h2), (h2, img)):
            x[b, :, a:e] = rng.uniform(0.1, 0.9, (3, 1, 1))
        for obj in range(rng.randint(1, 4)):
            cy, cx = rng.uniform(0.3, 0.9, 2) * img
            r = rng.uniform(0.06, 0.15) * img
            d2 = ((yy - cy) ** 2 + (xx - cx) ** 2) / (r * r)
            m = d2 < 1
            shade = (1 - 0.5 * d2[m]).astype(np.float32)
            col = rng.uniform(0, 1, 3).astype(np.float32)
            x[b][:, m] = col[:, None] * shade[None]
            inst[b, 0][m] = obj + 1
    return np.clip(x, 0, 1), inst


GENERATORS = {'multid': multid, 'stacks': stacks, 'rooms': rooms}
