"""Segmentation metrics.  CPU: the oracle restatement (oracle/metrics.py) equals the REAL
reference functions utils/misc.average_ari / average_segcover (skipped where /root/reference is absent).  GPU: the
g2_seg_metrics kernel behind genesis_b200.metrics equals the oracle, through the reference's own call signatures."""
import numpy as np
import pytest
import torch

from oracle import metrics as OM
from oracle import ref_loader


def random_case(seed, B=5, K=7, H=64, n_obj=4, degenerate=False):
    rng = np.random.RandomState(seed)
    inst = np.zeros((B, H, H), np.int64)
    for b in range(B):
        for o in range(1, 1 + rng.randint(0 if degenerate else 1, n_obj + 1)):
            y, x, h, w = rng.randint(0, H - 8), rng.randint(0, H - 8), rng.randint(4, 30), rng.randint(4, 30)
            inst[b, y:y + h, x:x + w] = o
    # predicted masks: noisy version of the ground truth spread over K slots
    logits = rng.randn(K, B, H, H).astype(np.float32)
    for b in range(B):
        perm = rng.permutation(K)
        for o in range(n_obj + 1):
            logits[perm[o % K], b][inst[b] == o] += 2.5
    if degenerate:
        logits[:, 0] = 0.0
        logits[2, 0] = 1.0          # image 0: a single predicted segment
        inst[1] = 0                 # image 1: background only (empty foreground)
    log_m = torch.log_softmax(torch.from_numpy(logits), dim=0).numpy()
    return log_m, inst


@pytest.mark.skipif(not ref_loader.available(), reason='reference checkout not present')
@pytest.mark.parametrize('seed,deg', [(0, False), (1, False), (2, True)])
def test_oracle_metrics_equal_reference(seed, deg):
    ref_loader._setup()
    from utils import misc
    log_m, inst = random_case(seed, degenerate=deg)
    K, B, H, _ = log_m.shape
    log_m_k = [torch.from_numpy(log_m[k]).unsqueeze(1) for k in range(K)]
    inst_t = torch.from_numpy(inst).unsqueeze(1)
    got = OM.metrics(log_m.reshape(K, B, -1), inst.reshape(B, -1))
    for fg, key in ((False, 'ari'), (True, 'ari_fg')):
        if deg and fg:
            continue            # sklearn rejects the empty foreground of image 1; the kernel defines it as 1.0
        mean, per = misc.average_ari(log_m_k, inst_t, foreground_only=fg)
        np.testing.assert_allclose(got[key], np.array(per), atol=1e-12)
    ins_seg = torch.argmax(torch.cat(log_m_k, 1), 1, True)
    np.testing.assert_array_equal(got['instance_seg'].reshape(B, 1, H, H), ins_seg.numpy())
    for ib, (k1, k2) in ((False, ('msc', 'msc_scaled')), (True, ('msc_fg', 'msc_fg_scaled'))):
        a, s = misc.average_segcover(inst_t, ins_seg, ignore_background=ib)
        assert abs(got[k1].mean() - float(a)) < 1e-6 and abs(got[k2].mean() - float(s)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('seed,deg,K', [(0, False, 7), (3, False, 11), (2, True, 5)])
def test_device_metrics_equal_oracle(seed, deg, K):
    from genesis_b200 import metrics
    log_m, inst = random_case(seed, K=K, degenerate=deg)
    _, B, H, _ = log_m.shape
    want = OM.metrics(log_m.reshape(K, B, -1), inst.reshape(B, -1))
    log_m_k = [torch.from_numpy(log_m[k]).unsqueeze(1).cuda() for k in range(K)]
    inst_t = torch.from_numpy(inst).unsqueeze(1).cuda()
    got = metrics.segmentation_metrics(log_m_k, inst_t)
    for key in ('ari', 'ari_fg', 'msc', 'msc_fg', 'msc_scaled', 'msc_fg_scaled'):
        np.testing.assert_allclose(got[key].cpu().numpy(), want[key], atol=1e-12, err_msg=key)
    np.testing.assert_array_equal(got['instance_seg'].cpu().numpy().reshape(B, -1), want['instance_seg'].reshape(B, -1))
    # the reference's call signatures
    mean, per = metrics.average_ari(log_m_k, inst_t, foreground_only=True)
    assert abs(mean - want['ari_fg'].mean()) < 1e-12 and len(per) == B
    a, s = metrics.average_segcover(inst_t, got['instance_seg'], ignore_background=True)
    assert abs(float(a) - want['msc_fg'].mean()) < 1e-6 and abs(float(s) - want['msc_fg_scaled'].mean()) < 1e-6
    with pytest.raises(RuntimeError):
        metrics.average_ari([t.cpu() for t in log_m_k], inst_t.cpu())
