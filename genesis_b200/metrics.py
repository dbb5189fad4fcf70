"""Segmentation metrics of the reference's evaluation loop (train.py:479-589) on the device.

`average_ari` / `average_segcover` keep the signatures and return conventions of utils/misc.py:101-114 and :173-235, but
one kernel launch (g2_seg_metrics: a confusion matrix per image in shared memory, then the closed forms) replaces the
per-image numpy / sklearn loops and their device->host copies.  `segmentation_metrics` returns everything at once."""
import torch

from . import _lib


def _run(log_m_k=None, pred=None, instances=None, want_seg=False):
    inst = instances.reshape(instances.shape[0], -1).contiguous().long()
    B, P = inst.shape
    if not inst.is_cuda:
        raise RuntimeError('genesis_b200.metrics runs on CUDA tensors only; there is no CPU path')
    out = torch.empty(B, 8, dtype=torch.float64, device=inst.device)
    seg = torch.empty(B, P, dtype=torch.int64, device=inst.device) if want_seg else None
    if log_m_k is not None:
        lm = log_m_k if torch.is_tensor(log_m_k) else torch.stack(list(log_m_k), 0)
        K = lm.shape[0]
        lm = lm.detach().reshape(K, B, P).contiguous().float()
        _lib.call('g2_seg_metrics', lm, None, inst, seg, out, B, P, K)
    else:
        pr = pred.reshape(B, P).contiguous().long()
        _lib.call('g2_seg_metrics', None, pr, inst, seg, out, B, P, 16)
    # out[:, 7] = pixels that entered the confusion matrix: the kernel holds ground-truth labels 0..31 and predicted labels 0..15
    # (the reference's sklearn / np.unique code takes any labels); a dropped pixel must be an error, not a silently wrong metric
    if bool((out[:, 7] != P).any()):
        raise ValueError('segmentation metrics: an instance label >= 32 (or < 0) or a predicted label >= 16; remap the labels '
                         'to dense ids first')
    return out, seg


def segmentation_metrics(log_m_k, instances):
    """log_m_k: list of K [B,1,H,W] log-masks (or a stacked [K,B,1,H,W] tensor); instances [B,1,H,W] integer labels.
    Returns a dict of per-image float64 tensors (ari, ari_fg, msc, msc_fg, msc_scaled, msc_fg_scaled) and `instance_seg`
    [B,1,H,W] int64 (argmax over slots)."""
    out, seg = _run(log_m_k=log_m_k, instances=instances, want_seg=True)
    names = ('ari', 'ari_fg', 'msc', 'msc_fg', 'msc_scaled', 'msc_fg_scaled')
    d = {n: out[:, i] for i, n in enumerate(names)}
    d['instance_seg'] = seg.view(instances.shape[0], 1, *instances.shape[-2:])
    return d


def average_ari(log_m_k, instances, foreground_only=False):
    """utils/misc.py:101-114 -> (mean ARI over the batch, list of per-image ARI)."""
    out, _ = _run(log_m_k=log_m_k, instances=instances)
    ari = out[:, 1 if foreground_only else 0]
    vals = ari.tolist()
    return sum(vals) / len(vals), vals


def average_segcover(segA, segB, ignore_background=False):
    """utils/misc.py:173-235: covering of segA (ground truth) by segB (prediction), both [B,1,H,W] integer maps.
    -> (mean over the batch of the unweighted covering, mean of the size-weighted covering), as float32 tensors."""
    assert segA.shape == segB.shape and segA.shape[1] == 1
    out, _ = _run(pred=segB, instances=segA)
    i = 1 if ignore_background else 0
    return out[:, 2 + i].mean(0).float(), out[:, 4 + i].mean(0).float()
