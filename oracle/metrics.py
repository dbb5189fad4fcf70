"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the reference's segmentation metrics.

average_ari: utils/misc.py:101-114 (per image: argmax over slots, optional foreground filter, sklearn's
adjusted_rand_score restated through its pair confusion matrix).  average_segcover: utils/misc.py:173-235 (per ground-truth
label the best IoU over predicted labels; mean and size-weighted mean).  Pinned against the real functions imported from
/root/reference in tests/test_metrics.py."""
import numpy as np


def contingency(gt, pred, na=32, nk=16):
    c = np.zeros((na, nk), np.float64)
    ok = (gt >= 0) & (gt < na) & (pred >= 0) & (pred < nk)
    np.add.at(c, (gt[ok], pred[ok]), 1.0)
    return c


def ari_from_contingency(c):
    n = c.sum()
    n_c, n_k = c.sum(1), c.sum(0)
    sumsq = (c ** 2).sum()
    tp = sumsq - n
    fp = (c @ n_k).sum() - sumsq
    fn = (c.T @ n_c).sum() - sumsq
    tn = n ** 2 - fp - fn - sumsq
    if fn == 0 and fp == 0:
        return 1.0
    return 2.0 * (tp * tn - fn * fp) / ((tp + fn) * (fn + tn) + (tp + fp) * (fp + tn))


def segcover_from_contingency(c, ignore_background):
    col = c.sum(0)
    scores, scaled, scaling, nlab = 0.0, 0.0, 0.0, 0
    for a in range(1 if ignore_background else 0, c.shape[0]):
        row = c[a].sum()
        if row == 0:
            continue
        union = row + col - c[a]
        iou = np.where(union > 0, c[a] / np.maximum(union, 1), 0.0)
        best = iou.max()
        scores += best; scaled += row * best; scaling += row; nlab += 1
    return scores / max(nlab, 1), scaled / max(scaling, 1.0)


def metrics(log_m, inst):
    """log_m [K,B,P] float, inst [B,P] int -> dict of per-image arrays."""
    K, B, P = log_m.shape
    pred = np.argmax(log_m, axis=0)
    out = {k: np.zeros(B) for k in ('ari', 'ari_fg', 'msc', 'msc_fg', 'msc_scaled', 'msc_fg_scaled')}
    for b in range(B):
        c = contingency(inst[b], pred[b])
        out['ari'][b] = ari_from_contingency(c)
        out['ari_fg'][b] = ari_from_contingency(c[1:])
        out['msc'][b], out['msc_scaled'][b] = segcover_from_contingency(c, False)
        out['msc_fg'][b], out['msc_fg_scaled'][b] = segcover_from_contingency(c, True)
    out['instance_seg'] = pred
    return out
