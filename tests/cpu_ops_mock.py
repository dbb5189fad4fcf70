"""TEST INFRASTRUCTURE: a plain-torch stand-in for the CONTRACT of genesis_b200.ops (same argument meaning, layouts and return
values, computed with ATen on the CPU), used by tests/test_plugins_cpu.py to run the plug-ins' and holders' host logic --
slot bookkeeping, layouts, weight re-indexing, noise order, loss assembly, BatchNorm buffer updates -- without a GPU and
compare it with the oracle.  It is NOT a fallback of the product: nothing under genesis_b200/ imports it, and the product
ops raise on CPU tensors.  The kernels themselves are checked on the GPU against the oracle."""
import contextlib

import torch
import torch.nn.functional as F

from oracle import functional as O

NORM_NONE, NORM_BATCH, NORM_INSTANCE, NORM_GROUP = 0, 1, 2, 3
POST_GATE, POST_RELU, POST_NONE = 0, 1, 2


def _act(y, act):
    return {None: y, 'none': y, 'relu': F.relu(y), 'elu': F.elu(y)}[act] if act in (None, 'none', 'relu', 'elu') else y


def to_nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def to_nhwc_padded(x, cp):
    y = x.detach().permute(0, 2, 3, 1)
    return F.pad(y, (0, cp - y.shape[3])).contiguous()


def conv2d(x, w, b=None, stride=1, pad=0, act=None, bias_grad=True):      # bias_grad: where the kernel layer computes db (autograd here)
    ci = w.shape[1]
    y = F.conv2d(x[..., :ci].permute(0, 3, 1, 2), w, b, stride=stride, padding=pad)
    return _act(y, act).permute(0, 2, 3, 1).contiguous()


def conv_transpose2d(x, w, b=None, stride=1, pad=0, act=None, bias_grad=True):
    ci = w.shape[0]
    y = F.conv_transpose2d(x[..., :ci].permute(0, 3, 1, 2), w, b, stride=stride, padding=pad, output_padding=stride - 1)
    return _act(y, act).permute(0, 2, 3, 1).contiguous()


def linear(x, w, b=None, act=None):
    return _act(F.linear(x, w, b), act)


def _norm(t, mode, w, b, rm, rv, groups, training, eps, momentum):
    if mode == NORM_NONE:
        return t
    if mode == NORM_BATCH:
        return F.batch_norm(t, rm, rv, w, b, training, momentum, eps)
    if mode == NORM_INSTANCE:
        return F.instance_norm(t, weight=w, bias=b, eps=eps)
    return F.group_norm(t, groups, w, b, eps)


def norm_post(y, g0=None, b0=None, g1=None, b1=None, rm0=None, rv0=None, rm1=None, rv1=None,
              mode=NORM_NONE, post=POST_GATE, groups=1, training=True, eps=1e-5, momentum=0.1, conv_bias=None):
    t = y.permute(0, 3, 1, 2)
    if post == POST_GATE:
        h, g = torch.chunk(t, 2, dim=1)
        h = _norm(h, mode, g0, b0, rm0, rv0, groups, training, eps, momentum)
        g = _norm(g, mode, g1, b1, rm1, rv1, groups, training, eps, momentum)
        out = h * torch.sigmoid(g)
    else:
        out = _norm(t, mode, g0, b0, rm0, rv0, groups, training, eps, momentum)
        if post == POST_RELU:
            out = F.relu(out)
    return out.permute(0, 2, 3, 1).contiguous()


def sbp_scan(logits, K):
    """g2_sbp_scan_fwd_f32: logits [nl,...] (nl = K or K-1) -> log_m [K,...], log_s [nl+1,...]."""
    nl = logits.shape[0]
    s = torch.zeros_like(logits[0])
    log_m, log_s = [None] * K, [s]
    for k in range(nl):
        if k < K - 1:
            log_m[k] = s + F.logsigmoid(logits[k])
        elif k == K - 1:
            log_m[k] = s
        s = s + F.logsigmoid(-logits[k])
        log_s.append(s)
    if nl == K - 1:
        log_m[K - 1] = s
    return torch.stack(log_m, 0), torch.stack(log_s, 0).detach()


def comp_pack(x, log_m, cp=4):
    K, B = log_m.shape[0], log_m.shape[1]
    t = torch.cat([log_m.reshape(K * B, 1, *x.shape[2:]), x.repeat(K, 1, 1, 1)], dim=1).permute(0, 2, 3, 1)
    return F.pad(t, (0, cp - t.shape[3])).contiguous()


def bcast_add_act(a, m, act=None):
    return _act(a.unsqueeze(1) + m.unsqueeze(0), act)


def out1x1(h, w, b=None, nsig=0):
    y = F.conv2d(h.permute(0, 3, 1, 2), w, b)
    if nsig:
        y = torch.cat([torch.sigmoid(y[:, :nsig]), y[:, nsig:]], dim=1)
    return y.contiguous()


def mixture_nll(x, xr, lm, std, softmax=False):
    K = xr.shape[0]
    if softmax:
        lm = torch.log_softmax(lm, dim=0)
    logp = lm + O.normal_log_prob(x.unsqueeze(0), xr, std.view(K, 1, 1, 1, 1))
    err = -torch.log(torch.exp(logp).sum(0)).sum(dim=(1, 2, 3))
    recon = (lm.exp() * xr).sum(0)
    return err, recon.detach(), (lm.detach() if softmax else x.new_zeros(0))


def mixture_nll_packed(x, dec, lm, std, softmax):
    return mixture_nll(x, dec[:, :, :3], dec[:, :, 3:] if softmax else lm, std, softmax)


def monet_loss(x, dec, lm, std):
    err, recon, _ = mixture_nll(x, dec[:, :, :3], lm, std, False)
    K, B = lm.shape[0], lm.shape[1]
    lmr = torch.log_softmax(dec[:, :, 3:], dim=0)
    q = lm.exp().clamp_min(1e-5).permute(1, 2, 3, 4, 0).reshape(-1, K)
    p = lmr.exp().clamp_min(1e-5).permute(1, 2, 3, 4, 0).reshape(-1, K)
    kl = O.categorical_kl(q, p).view(B, -1).sum(dim=1)
    return err, kl, recon, lmr.detach()


def mask_kl(lm, dec, detach=True):
    K, B = lm.shape[0], lm.shape[1]
    lmr = torch.log_softmax(dec[:, :, 3:] if dec.shape[2] == 4 else dec, dim=0)
    if detach:
        lmr = lmr.detach()
    q = lm.exp().clamp_min(1e-5).permute(1, 2, 3, 4, 0).reshape(-1, K)
    p = lmr.exp().clamp_min(1e-5).permute(1, 2, 3, 4, 0).reshape(-1, K)
    return O.categorical_kl(q, p).view(B, -1).sum(dim=1)


def down2(x):
    return F.interpolate(x.permute(0, 3, 1, 2), scale_factor=0.5, mode='nearest').permute(0, 2, 3, 1).contiguous()


def up2(x):
    return F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode='nearest').permute(0, 2, 3, 1).contiguous()


def icsbp(colour, u, log_sigma, K, kernel='gaussian'):
    from oracle import models as M
    log_m_k, log_s_k, seeds, idxs = M.icsbp(colour.permute(0, 3, 1, 2), u, log_sigma, K - 1, kernel)
    return torch.stack(log_m_k, 0), torch.stack(log_s_k[:K], 0).detach(), torch.stack(idxs, 0).int()


def icsbp_dynamic(colour, u, log_sigma, K, kernel='gaussian'):
    """g2_icsbp_dynamic_*: per image, early exit, -1e10 padding up to K, n_masks [B]."""
    from oracle import models as M
    B, Hh, Ww = colour.shape[0], colour.shape[1], colour.shape[2]
    lm, ls, idx, n = [], [], [], []
    for b in range(B):
        log_m_k, log_s_k, seeds, idxs = M.icsbp(colour[b:b + 1].permute(0, 3, 1, 2), u[b:b + 1], log_sigma, K - 1, kernel, True)
        nb = len(log_m_k)
        pad = torch.full((1, 1, Hh, Ww), -1e10, dtype=colour.dtype)
        lm.append(torch.stack(log_m_k + [pad] * (K - nb), 0))
        ls.append(torch.stack(log_s_k + [log_s_k[-1]] * (K - nb), 0).detach())
        ii = [int(i) for i in idxs][:K - 1]
        idx.append(torch.tensor(ii + [-1] * (K - 1 - len(ii)), dtype=torch.int32))
        n.append(nb)
    return torch.cat(lm, 1), torch.cat(ls, 1), torch.stack(idx, 1), torch.tensor(n, dtype=torch.int32)


def masked_pool(f, log_m):
    m = log_m.exp().flatten(2)                                   # [K,B,P]
    num = torch.einsum('kbp,bpc->kbc', m, f.flatten(1, 2))
    return num, m.sum(2)


class _Stream(object):
    def wait_stream(self, other):
        pass


_STREAM = _Stream()


def install(monkeypatch, ops):
    """Replace the kernel-backed entry points of genesis_b200.ops (and the CUDA stream calls the plug-ins make) with the
    stand-ins above for the duration of a test."""
    g = globals()
    for name in ('to_nhwc', 'to_nchw', 'to_nhwc_padded', 'conv2d', 'conv_transpose2d', 'linear', 'norm_post', 'sbp_scan',
                 'comp_pack', 'bcast_add_act', 'out1x1', 'mixture_nll', 'mixture_nll_packed', 'monet_loss', 'down2', 'up2',
                 'icsbp', 'icsbp_dynamic', 'masked_pool', 'mask_kl'):
        monkeypatch.setattr(ops, name, g[name])
    monkeypatch.setattr(ops, 'get_precision', lambda: 'fp32')
    monkeypatch.setattr(ops, 'fused_latent', lambda: False)      # the ATen formulation in holders.py is the contract of the fused latent kernels
    monkeypatch.setattr(ops, 'side_streams_enabled', lambda: False)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: _STREAM)
    monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())


class AsCuda(torch.Tensor):
    """A CPU tensor that answers is_cuda = True, to pass the plug-ins' `no CPU path` guard in this test only."""

    @property
    def is_cuda(self):
        return True
