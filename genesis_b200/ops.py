"""Autograd operators over the C ABI of libgenesis_b200.so.

Every heavy operator of the hot path (conv, conv-transpose, linear, norm+gate/ReLU, stick-breaking scan,
broadcast-decoder layers, mixture likelihood) is a torch.autograd.Function whose forward and backward
enqueue hand-written sm_100a kernels on the current CUDA stream.  Activations are NHWC fp32.
There is no CPU or library fallback: a non-CUDA tensor raises."""
import os

import torch
from torch.autograd import Function

from . import _lib

ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
NORM_NONE, NORM_BATCH, NORM_INSTANCE, NORM_GROUP = 0, 1, 2, 3
POST_GATE, POST_RELU, POST_NONE = 0, 1, 2
ACTS = {None: 0, 'none': 0, 'relu': 1, 'elu': 2}


_PRECISION = {'mode': 'tf32'}


def set_precision(mode):
    """'tf32': tcgen05 tensor-core kernels wherever the shape fits (product default);
    'fp32': exact-fp32 SIMT kernels everywhere (on-device cross-check)."""
    assert mode in ('tf32', 'fp32')
    _PRECISION['mode'] = mode


def get_precision():
    return _PRECISION['mode']


def _call(name, *args):
    _lib.call(name, *args)


_SIDE = {}
_SIDE_ON = {'on': os.environ.get('G2_SIDE_STREAMS', '1') != '0', 'grad': os.environ.get('G2_GRAD_STREAM', '1') != '0'}


def set_side_streams(on):
    """False: run everything on the current stream (clean per-kernel timing in bench.py's profiled steps)."""
    _SIDE_ON['on'] = bool(on)


def side_streams_enabled():
    return _SIDE_ON['on']


def side_stream(device, idx=0):
    """Auxiliary CUDA streams per device for work that does not feed the critical path: idx 0 = the prior / KL branch of
    the forward (and its backward), idx 1 = parameter-gradient kernels in direct-gradient mode."""
    dev = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    key = (dev, idx)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


class _GradStream(object):
    """Direct-gradient mode: weight / bias gradient kernels write into param.grad and nothing reads them before the
    optimiser, so they are enqueued on side stream 1 (fork after the operands are ready; trainer.TrainStep joins once
    after backward).  The data-gradient chain -- the critical path of the backward pass -- never waits for them."""

    def __init__(self, *operands):
        self.cur = torch.cuda.current_stream()
        self.side = side_stream(operands[0].device, 1) if (_SIDE_ON['on'] and _SIDE_ON['grad']) else self.cur
        if self.side is not self.cur:
            self.side.wait_stream(self.cur)
            for t in operands:
                t.record_stream(self.side)
        self.ctx = torch.cuda.stream(self.side)

    def __enter__(self):
        return self.ctx.__enter__()

    def __exit__(self, *a):
        return self.ctx.__exit__(*a)


def join_grad_stream(device):
    """Make the current stream wait for all parameter-gradient kernels queued on the gradient side stream."""
    if (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), 1) in _SIDE:
        torch.cuda.current_stream().wait_stream(side_stream(device, 1))


def _tc_ok(N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode):
    if _PRECISION['mode'] != 'tf32':
        return False
    return _lib.lib().query('g2_conv_tf32_supported', N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode) == 1


def _wgrad_any(g, t, dims, R, S, stride, pad, outT, like):
    """dW[r,s,a,b] = sum g[.., a] * t[.., b] on the tensor cores when supported, else fp32 SIMT.
    Returns the packed gradient [R,S,Cg,Ct] (outT=0) or [R,S,Ct,Cg] (outT=1)."""
    N, Hg, Wg, Cg, Ht, Wt, Ct = dims
    dwp = _new(like, R, S, Ct, Cg) if outT else _new(like, R, S, Cg, Ct)
    ws_bytes = 0
    if _PRECISION['mode'] == 'tf32':
        ws_bytes = _lib.lib().query('g2_conv_wgrad_tf32_workspace', N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride)
    if ws_bytes > 0:
        ws = _new(like, ws_bytes // 4)
        _call('g2_conv_wgrad_tf32', g, t, dwp, ws, N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, pad, outT)
    else:
        _call('g2_conv_wgrad_f32', g, t, dwp, N, Hg, Wg, Cg, Ht, Wt, Ct, R, S, stride, pad, outT)
    return dwp


def _conv_tc(x, pack, bias, out, dims, R, S, stride, pad, mode, act):
    """Tensor-core implicit GEMM through the entry point that will run it: the halo kernel (activation window resident,
    taps via shifted descriptors) when it covers the problem, else the tile kernel -- same contract, separate profile rows."""
    N, Hi, Wi, Ci, Ho, Wo, Co = dims
    if _lib.lib().query('g2_conv_halo_supported', N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode) == 1:
        _call('g2_conv_halo_tf32', x, pack, bias, out, N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode, act)
    else:
        _call('g2_conv_igemm_tf32', x, pack, bias, out, N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode, act)


def _conv_any(x, w_t, bias, out, dims, R, S, stride, pad, mode, act, perm_tc, perm_simt_wT):
    """Run mode-0/1 implicit GEMM on the tensor cores when supported, else the fp32 SIMT kernel.
    w_t: weight in torch layout; perm_tc: permutation giving [R,S,Cout_op,Cred_op];
    perm_simt_wT: (permutation giving the packed SIMT weight, wT flag)."""
    N, Hi, Wi, Ci, Ho, Wo, Co = dims
    if _tc_ok(N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode):
        wp = w_t.permute(*perm_tc).contiguous()
        _conv_tc(x, wp, bias, out, dims, R, S, stride, pad, mode, act)
    else:
        perm, wT = perm_simt_wT
        wp = w_t.permute(*perm).contiguous()
        _call('g2_conv_igemm_f32', x, wp, bias, None, out, N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode, wT, act)


def _new(like, *shape, dtype=torch.float32):
    return torch.empty(shape, device=like.device, dtype=dtype)


def _c(t):
    """contiguous and 16-byte aligned (TMA and float4 accesses need it; flat-arena views may not be)."""
    if not t.is_contiguous():
        return t.contiguous()
    if t.data_ptr() % 16 != 0:
        return t.clone()
    return t


def _act_bwd(dout, out, act):
    if act == ACT_NONE:
        return dout
    dpre = torch.empty_like(dout)
    _call('g2_act_bwd_f32', dout, out, dpre, dout.numel(), act)
    return dpre


def _colsum(x2d_rows, C, like):
    out = _new(like, C)
    rows = x2d_rows.numel() // C
    if C > 1024 and C % 4 == 0:
        _call('g2_sum_dim0_f32', x2d_rows, out, rows, C)
    else:
        _call('g2_colsum_f32', x2d_rows, out, rows, C, 0)
    return out


# ----------------------------------------------------------------------------------------- layout
class _Layout(Function):
    @staticmethod
    def forward(ctx, x, to_nchw):
        x = _c(x)
        ctx.to_nchw = to_nchw
        if to_nchw:
            N, H, W, C = x.shape
            y = _new(x, N, C, H, W)
        else:
            N, C, H, W = x.shape
            y = _new(x, N, H, W, C)
        _call('g2_layout_f32', x, y, N, C, H * W, 1 if to_nchw else 0)
        return y

    @staticmethod
    def backward(ctx, dy):
        return _Layout.apply(dy, not ctx.to_nchw), None


def to_nhwc(x):
    return _Layout.apply(x, False)


def to_nchw(x):
    return _Layout.apply(x, True)


# ----------------------------------------------------------------------------------------- conv
_DIRECT = {'on': False}


def set_direct_grad(on):
    """True (set by trainer.TrainStep, which owns flat, pre-zeroed gradient arenas): weight / bias gradients of the conv
    and linear operators are ACCUMULATED by the kernels straight into `param.grad` (torch layout) and the autograd
    Functions return None for them -- no permute copy and no AccumulateGrad add per parameter.
    False (drop-in mode under the reference's train.py): gradients are returned to autograd as usual."""
    _DIRECT['on'] = bool(on)


_NORM_DIRECT = {'on': os.environ.get('G2_NORM_DIRECT', '1') == '1'}     # validated + timed on a B200 (round 2): +1.5 % images/s


def set_norm_direct(on):
    _NORM_DIRECT['on'] = bool(on)


def _direct(p):
    return (_DIRECT['on'] and p is not None and p.is_leaf and p.requires_grad and p.grad is not None
            and p.grad.is_contiguous() and p.grad.data_ptr() % 16 == 0)


# Gradient-ready tracking (data-parallel overlap, trainer.TrainStep with world_size > 1): in direct-gradient mode the autograd
# Functions write parameter gradients themselves, so they also know WHEN a parameter's gradient is complete: every forward use is
# counted, every backward that has enqueued its gradient kernels counts down, and at zero the callback fires -- the trainer then
# all-reduces the bucket the parameter belongs to while the rest of the backward pass is still running.
_GRAD_TRACK = {'cb': None, 'uses': {}}


def set_grad_ready_callback(cb):
    _GRAD_TRACK['cb'] = cb
    _GRAD_TRACK['uses'] = {}


def reset_grad_tracking():
    _GRAD_TRACK['uses'] = {}


def all_side_streams(device):
    """Every auxiliary stream created for `device` so far (a collective must be ordered after all of them)."""
    dev = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    return [s_ for (d, _), s_ in _SIDE.items() if d == dev]


def _track_use(*params):
    if _GRAD_TRACK['cb'] is None or not _DIRECT['on'] or not torch.is_grad_enabled():
        return
    uses = _GRAD_TRACK['uses']
    for p in params:
        if p is not None and _direct(p):
            uses[id(p)] = uses.get(id(p), 0) + 1


def _track_done(*params):
    cb = _GRAD_TRACK['cb']
    if cb is None or not _DIRECT['on']:
        return
    uses = _GRAD_TRACK['uses']
    for p in params:
        if p is not None and id(p) in uses:
            uses[id(p)] -= 1
            if uses[id(p)] == 0:
                del uses[id(p)]
                cb(p)


def _pad_dim(w, dim, size):
    if w.shape[dim] == size:
        return w
    shape = list(w.shape)
    shape[dim] = size - w.shape[dim]
    return torch.cat([w, w.new_zeros(shape)], dim=dim)


def _bias_grad(dpre, b, C):
    """Bias gradient = column sum of dpre [rows, C]; accumulated into b.grad in direct mode (returns None)."""
    rows = dpre.numel() // C
    if _direct(b) and not (C > 1024 and C % 4 == 0):
        _call('g2_colsum_f32', dpre, b.grad, rows, C, 1)
        return None
    return _colsum(dpre, C, dpre)


def _conv_fwd_common(ctx, x, w, b, stride, pad, act, transposed, bias_grad=True):
    """Shared forward of Conv2d (mode 0) / ConvTranspose2d (mode 1, output_padding = stride-1).  The weight may have fewer
    input channels than x (x zero-padded to a 32-channel k-block): the missing channels are zero in the packs."""
    x = _c(x)
    N, H, W, Cx = x.shape
    if transposed:
        Ci, Co, R, S = w.shape
        op = stride - 1
        Ho = (H - 1) * stride - 2 * pad + R + op
        Wo = (W - 1) * stride - 2 * pad + S + op
    else:
        Co, Ci, R, S = w.shape
        Ho = (H + 2 * pad - R) // stride + 1
        Wo = (W + 2 * pad - S) // stride + 1
    assert Ci <= Cx, (x.shape, w.shape)
    mode_f, mode_d = (1, 0) if transposed else (0, 1)
    wd = w.detach()
    out = _new(x, N, Ho, Wo, Co)
    pb = None
    if _X3['on'] and _X3['fwd'] and Cx % 32 == 0 and _tc_ok(N, H, W, 3 * Cx, Ho, Wo, Co, R, S, stride, pad, mode_f):
        # precise layer (ops.precise): the FORWARD contraction as 3xTF32 -- [x_hi|x_hi|x_lo] against [w_hi|w_lo|w_hi] over a 3x
        # longer reduction, fp32-level accuracy.  Measured on B200 (profiles/r02_parity_x3.txt): the gradient errors of the
        # UNet models come from the forward rounding alone, so the backward contractions below stay plain TF32.
        wp = _pad_dim(wd, 0 if transposed else 1, Cx)
        w_hi, w_lo = _w_hi_lo(wp)
        cin_dim = 0 if transposed else 1
        if _X3_INKERNEL and _lib.lib().query('g2_conv_halo_supported', N, H, W, Cx, Ho, Wo, Co, R, S, stride, pad, mode_f) == 1:
            # in-kernel split (igemm_halo.cu): the kernel derives x_lo from its resident window -- no [hi|hi|lo] copy of the
            # activation in HBM; only the (tiny) weight is split here
            pa2 = _new(x, R * S, Co, 2 * Cx)
            _call('g2_pack_conv_weight_f32', _c(torch.cat([w_hi, w_lo], dim=cin_dim)), pa2, None, Co, 2 * Cx, 2 * Cx, R * S,
                  1 if transposed else 0)
            _call('g2_conv_halo_x3_tf32', x, pa2, b, out, N, H, W, Cx, Ho, Wo, Co, R, S, stride, pad, mode_f, act)
            if _X3_DEBUG:       # cross-check against the pre-split route (scripts/parity_report.py, G2_X3_DEBUG=1)
                ref = torch.empty_like(out)
                pa3 = _new(x, R * S, Co, 3 * Cx)
                _call('g2_pack_conv_weight_f32', _c(torch.cat([w_hi, w_lo, w_hi], dim=cin_dim)), pa3, None, Co, 3 * Cx, 3 * Cx,
                      R * S, 1 if transposed else 0)
                _conv_tc(_split(x, Cx, 0), pa3, b, ref, (N, H, W, 3 * Cx, Ho, Wo, Co), R, S, stride, pad, mode_f, act)
                d = (out.double() - ref.double())
                msg = 'in-kernel vs pre-split rel-L2 %.2e' % (d.norm().item() / max(ref.double().norm().item(), 1e-30))
                if not transposed and act == ACT_NONE:
                    import torch.nn.functional as _F
                    t64 = _F.conv2d(x[..., :Ci].permute(0, 3, 1, 2).double(), wd.double(), None if b is None else b.detach().double(),
                                    stride=stride, padding=pad).permute(0, 2, 3, 1)
                    nn = t64.norm().item()
                    msg += ' | vs fp64: in-kernel %.2e pre-split %.2e' % ((out.double() - t64).norm().item() / nn,
                                                                          (ref.double() - t64).norm().item() / nn)
                    pa1 = _new(x, R * S, Co, Cx)
                    _call('g2_pack_conv_weight_f32', _c(wp), pa1, None, Co, Cx, Cx, R * S, 0)
                    _conv_tc(x, pa1, b, ref, (N, H, W, Cx, Ho, Wo, Co), R, S, stride, pad, mode_f, act)
                    msg += ' plain-tf32 %.2e' % ((ref.double() - t64).norm().item() / nn)
                print('X3DEBUG', (N, H, W, Cx, Ho, Wo, Co, R, S, stride, pad, mode_f, act), msg)
        else:
            w3 = torch.cat([w_hi, w_lo, w_hi], dim=cin_dim)
            pa3 = _new(x, R * S, Co, 3 * Cx)
            _call('g2_pack_conv_weight_f32', _c(w3), pa3, None, Co, 3 * Cx, 3 * Cx, R * S, 1 if transposed else 0)
            _conv_tc(_split(x, Cx, 0), pa3, b, out, (N, H, W, 3 * Cx, Ho, Wo, Co), R, S, stride, pad, mode_f, act)
        if ctx.needs_input_grad[0] and _tc_ok(N, Ho, Wo, Co, H, W, Cx, R, S, stride, pad, mode_d):
            pb = _new(x, R * S, Cx, Co)
            _call('g2_pack_conv_weight_f32', _c(wd), _new(x, R * S, Co, Cx), pb, Co, Ci, Cx, R * S, 1 if transposed else 0)
    elif _tc_ok(N, H, W, Cx, Ho, Wo, Co, R, S, stride, pad, mode_f):
        pa = _new(x, R * S, Co, Cx)
        if ctx.needs_input_grad[0] and _tc_ok(N, Ho, Wo, Co, H, W, Cx, R, S, stride, pad, mode_d):
            pb = _new(x, R * S, Cx, Co)
        _call('g2_pack_conv_weight_f32', _c(wd), pa, pb, Co, Ci, Cx, R * S, 1 if transposed else 0)
        _conv_tc(x, pa, b, out, (N, H, W, Cx, Ho, Wo, Co), R, S, stride, pad, mode_f, act)
    else:
        wp = _pad_dim(wd, 0 if transposed else 1, Cx)
        perm = (2, 3, 0, 1) if transposed else (2, 3, 1, 0)          # [R,S,Cred,Cout]
        _call('g2_conv_igemm_f32', x, wp.permute(*perm).contiguous(), b, None, out, N, H, W, Cx, Ho, Wo, Co, R, S, stride, pad,
              mode_f, 0, act)
    ctx.save_for_backward(x, wd, out if act != ACT_NONE else None, pb)
    ctx.params = (w, b)
    ctx.cfg = (stride, pad, act, b is not None and bias_grad, transposed)
    _track_use(w, b if bias_grad else None)
    return out


def _conv_bwd_common(ctx, dout):
    x, wd, out, pb = ctx.saved_tensors
    w_param, b_param = ctx.params
    stride, pad, act, has_b, transposed = ctx.cfg
    N, H, W, Cx = x.shape
    if transposed:
        Ci, Co, R, S = wd.shape
    else:
        Co, Ci, R, S = wd.shape
    dout = _c(dout)
    _, Ho, Wo, _ = dout.shape
    mode_d = 0 if transposed else 1
    dx = dw = db = None
    if act != ACT_NONE and has_b and ctx.needs_input_grad[2] and Co % 4 == 0 and Co <= 1024:
        # conv + bias + activation: activation backward and bias gradient in ONE pass over dout (no column-sum re-read of dpre)
        dpre = torch.empty_like(dout)
        if _direct(b_param):
            _call('g2_act_bwd_bias_f32', dout, out, dpre, b_param.grad, dout.numel() // Co, Co, act)
        else:
            db = torch.zeros(Co, device=dout.device, dtype=torch.float32)
            _call('g2_act_bwd_bias_f32', dout, out, dpre, db, dout.numel() // Co, Co, act)
        has_b = False
    else:
        dpre = _act_bwd(dout, out, act)
    if ctx.needs_input_grad[0]:
        dx = torch.empty_like(x)
        if pb is not None:
            _conv_tc(dpre, pb, None, dx, (N, Ho, Wo, Co, H, W, Cx), R, S, stride, pad, mode_d, ACT_NONE)
        else:   # data gradient = the other mode with the reduction over Co: SIMT pack [R,S,Cx,Co] + wT (or the TC pack)
            wp = _pad_dim(wd, 0 if transposed else 1, Cx)
            perm = (2, 3, 0, 1) if transposed else (2, 3, 1, 0)      # [R,S,Cx,Co]
            _conv_any(dpre, wp, None, dx, (N, Ho, Wo, Co, H, W, Cx), R, S, stride, pad, mode_d, ACT_NONE, perm, (perm, 1))
    if ctx.needs_input_grad[1]:
        # conv: g = x (a = Cx), t = dpre (b = Co);  conv-transpose: g = dpre (a = Co), t = x (b = Cx)
        if transposed:
            g, t, dims = dpre, x, (N, Ho, Wo, Co, H, W, Cx)
            strides, lims = (1, R * S, Co * R * S), (Co, Ci)         # torch [Ci,Co,R,S]: a = Co, b = Ci
        else:
            g, t, dims = x, dpre, (N, H, W, Cx, Ho, Wo, Co)
            strides, lims = (1, R * S, Ci * R * S), (Ci, Co)         # torch [Co,Ci,R,S]: a = Ci, b = Co
        ws_bytes = 0
        if _PRECISION['mode'] == 'tf32':
            ws_bytes = _lib.lib().query('g2_conv_wgrad_tf32_workspace', *dims, R, S, stride)
        if ws_bytes > 0 and _direct(w_param):
            with _GradStream(g, t):
                ws = _new(x, ws_bytes // 4)
                _call('g2_conv_wgrad_tf32_to', g, t, w_param.grad, ws, *dims, R, S, stride, pad, *strides, *lims, 1)
                if has_b and ctx.needs_input_grad[2] and _direct(b_param):
                    _bias_grad(dpre, b_param, Co)
                    has_b = False
        elif ws_bytes > 0:
            ws = _new(x, ws_bytes // 4)
            dw = torch.empty_like(wd)
            _call('g2_conv_wgrad_tf32_to', g, t, dw, ws, *dims, R, S, stride, pad, *strides, *lims, 0)
        else:
            if transposed:
                dwp = _new(x, R, S, Cx, Co)
                _call('g2_conv_wgrad_f32', g, t, dwp, *dims, R, S, stride, pad, 1)
                dw = dwp.permute(2, 3, 0, 1)[:Ci].contiguous()
            else:
                dwp = _new(x, R, S, Cx, Co)
                _call('g2_conv_wgrad_f32', g, t, dwp, *dims, R, S, stride, pad, 0)
                dw = dwp.permute(3, 2, 0, 1)[:, :Ci].contiguous()
    if has_b and ctx.needs_input_grad[2]:
        db = _bias_grad(dpre, b_param, Co)
    _track_done(w_param, b_param if ctx.cfg[3] else None)
    return dx, dw, db, None, None, None, None


class _Conv(Function):
    """nn.Conv2d on NHWC activations; w in torch layout [Co,Ci,R,S] (Ci may be smaller than x's zero-padded channels)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, act, bias_grad=True):
        return _conv_fwd_common(ctx, x, w, b, stride, pad, act, False, bias_grad)

    @staticmethod
    def backward(ctx, dout):
        return _conv_bwd_common(ctx, dout)


def conv2d(x, w, b=None, stride=1, pad=0, act=None, bias_grad=True):
    """bias_grad=False: the bias gradient is produced elsewhere (norm_post(conv_bias=b) fuses it into the norm backward)."""
    if _x3_conv_ok(x, w, stride, pad, False):
        return _ConvX3.apply(x, w, b, stride, pad, ACTS[act], False, bias_grad)
    return _Conv.apply(x, w, b, stride, pad, ACTS[act], bias_grad)


class _ConvT(Function):
    """nn.ConvTranspose2d (output_padding = stride-1) on NHWC activations; w torch layout [Ci,Co,R,S]."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, act, bias_grad=True):
        return _conv_fwd_common(ctx, x, w, b, stride, pad, act, True, bias_grad)

    @staticmethod
    def backward(ctx, dout):
        return _conv_bwd_common(ctx, dout)


def conv_transpose2d(x, w, b=None, stride=1, pad=0, act=None, bias_grad=True):
    if _x3_conv_ok(x, w, stride, pad, True):
        return _ConvX3.apply(x, w, b, stride, pad, ACTS[act], True, bias_grad)
    return _ConvT.apply(x, w, b, stride, pad, ACTS[act], bias_grad)


# ----------------------------------------------------------------------------------------- tf32x3 ("precise") layers
# GENESIS-V2's and MONet's UNets (per-sample norms over flat images, K-1 recurrent passes) amplify the 2^-11 operand rounding of a
# plain TF32 contraction into 5-15 % per-tensor gradient errors.  Inside `with ops.precise():` convolutions and linears run as
# 3xTF32: every operand is split into hi = tf32(x) and lo = x - hi, and a * b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi is evaluated as
# ONE tensor-core contraction over a 3x longer reduction ([hi|hi|lo] channel blocks against [w_hi|w_lo|w_hi]) -- fp32-level
# accuracy (2^-21) on the same tcgen05 kernels, forward, data gradient and weight gradient.
import contextlib as _contextlib

_X3 = {'on': False, 'fwd': True, 'dgrad': False, 'wgrad': False}     # measured: forward rounding is what matters
_X3_INKERNEL = os.environ.get('G2_X3_INKERNEL', '1') != '0'          # halo kernel derives x_lo from its resident window
_X3_DEBUG = os.environ.get('G2_X3_DEBUG', '0') == '1'


@_contextlib.contextmanager
def precise(on=True, backward=None):
    """Layers created inside run their forward contraction as 3xTF32; backward=True also their data- and weight-gradient
    contractions (small latent heads whose backward GEMMs sit on a sensitive gradient path)."""
    prev = dict(_X3)
    _X3['on'] = bool(on) and _PRECISION['mode'] == 'tf32'
    if backward is not None:
        _X3['dgrad'] = _X3['wgrad'] = bool(backward)
    try:
        yield
    finally:
        _X3.update(prev)


def set_precise_parts(fwd=True, dgrad=False, wgrad=False):
    """Which of the three contractions of a precise layer use 3xTF32 (experiments: the others stay plain TF32)."""
    _X3.update(fwd=bool(fwd), dgrad=bool(dgrad), wgrad=bool(wgrad))


def _split(x2d_like, C, mode):
    """x viewed as [rows, C] -> [hi|hi|lo] (mode 0), [hi|lo] (1), hi (2), lo (3) along the last dim."""
    x = _c(x2d_like)
    rows = x.numel() // C
    oc = {0: 3 * C, 1: 2 * C, 2: C, 3: C}[mode]
    out = _new(x, *x.shape[:-1], oc)
    _call('g2_split_tf32_f32', x, out, rows, C, mode)
    return out


def _w_hi_lo(w):
    w = _c(w)
    n = w.numel()
    if n % 4 != 0:       # tiny odd-sized weights: plain torch (bit-identical definition of hi / lo)
        hi = (w.view(torch.int32) + 0x1000).bitwise_and(-8192).view(torch.float32)
        return hi, w - hi
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _call('g2_split_tf32_f32', w, hi, n // 4, 4, 2)
    _call('g2_split_tf32_f32', w, lo, n // 4, 4, 3)
    return hi, lo


class _ConvX3(Function):
    """Conv2d / ConvTranspose2d (output_padding = stride-1) on NHWC activations as 3xTF32 contractions; same contract as _Conv."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, act, transposed, bias_grad):
        x = _c(x)
        N, H, W, Cx = x.shape
        if transposed:
            Ci, Co, R, S = w.shape
            Ho = (H - 1) * stride - 2 * pad + R + stride - 1
            Wo = (W - 1) * stride - 2 * pad + S + stride - 1
        else:
            Co, Ci, R, S = w.shape
            Ho = (H + 2 * pad - R) // stride + 1
            Wo = (W + 2 * pad - S) // stride + 1
        mode_f = 1 if transposed else 0
        wd = _pad_dim(w.detach(), 0 if transposed else 1, Cx)           # zero rows for the padded input channels
        w_hi, w_lo = _w_hi_lo(wd)
        out = _new(x, N, Ho, Wo, Co)
        cin_dim = 0 if transposed else 1
        if _X3['fwd']:
            x3 = _split(x, Cx, 0)
            w3 = torch.cat([w_hi, w_lo, w_hi], dim=cin_dim)
            pa = _new(x, R * S, Co, 3 * Cx)
            _call('g2_pack_conv_weight_f32', _c(w3), pa, None, Co, 3 * Cx, 3 * Cx, R * S, 1 if transposed else 0)
            _conv_tc(x3, pa, b, out, (N, H, W, 3 * Cx, Ho, Wo, Co), R, S, stride, pad, mode_f, act)
        else:
            pa = _new(x, R * S, Co, Cx)
            _call('g2_pack_conv_weight_f32', _c(wd), pa, None, Co, Cx, Cx, R * S, 1 if transposed else 0)
            _conv_tc(x, pa, b, out, (N, H, W, Cx, Ho, Wo, Co), R, S, stride, pad, mode_f, act)
        ctx.save_for_backward(x, w_hi, w_lo, out if act != ACT_NONE else None)
        ctx.params = (w, b)
        ctx.cfg = (stride, pad, act, b is not None and bias_grad, transposed, Ci)
        ctx.parts = (_X3['dgrad'], _X3['wgrad'])
        _track_use(w, b if bias_grad else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w_hi, w_lo, out = ctx.saved_tensors
        x3_dgrad, x3_wgrad = ctx.parts
        w_param, b_param = ctx.params
        stride, pad, act, has_b, transposed, Ci = ctx.cfg
        N, H, W, Cx = x.shape
        if transposed:
            _, Co, R, S = w_hi.shape
        else:
            Co, _, R, S = w_hi.shape
        dout = _c(dout)
        _, Ho, Wo, _ = dout.shape
        mode_d = 0 if transposed else 1
        dx = dw = db = None
        if act != ACT_NONE and has_b and ctx.needs_input_grad[2] and Co % 4 == 0:
            dpre = torch.empty_like(dout)
            if _direct(b_param):
                _call('g2_act_bwd_bias_f32', dout, out, dpre, b_param.grad, dout.numel() // Co, Co, act)
            else:
                db = torch.zeros(Co, device=dout.device, dtype=torch.float32)
                _call('g2_act_bwd_bias_f32', dout, out, dpre, db, dout.numel() // Co, Co, act)
            has_b = False
        else:
            dpre = _act_bwd(dout, out, act)
        co_dim = 1 if transposed else 0
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if x3_dgrad:
                d3 = _split(dpre, Co, 0)                                         # [hi|hi|lo] over the reduction (Co)
                w3 = torch.cat([w_hi, w_lo, w_hi], dim=co_dim)                   # 3*Co output channels
                pb = _new(x, R * S, Cx, 3 * Co)
                _call('g2_pack_conv_weight_f32', _c(w3), _new(x, R * S, 3 * Co, Cx), pb, 3 * Co, Cx, Cx, R * S, 1 if transposed else 0)
                _conv_tc(d3, pb, None, dx, (N, Ho, Wo, 3 * Co, H, W, Cx), R, S, stride, pad, mode_d, ACT_NONE)
            else:
                wd = w_hi + w_lo
                pb = _new(x, R * S, Cx, Co)
                _call('g2_pack_conv_weight_f32', _c(wd), _new(x, R * S, Co, Cx), pb, Co, Cx, Cx, R * S, 1 if transposed else 0)
                _conv_tc(dpre, pb, None, dx, (N, Ho, Wo, Co, H, W, Cx), R, S, stride, pad, mode_d, ACT_NONE)
        if ctx.needs_input_grad[1]:
            # conv: g = x (a = Cx), t = dpre (b = Co);  conv-transpose: g = dpre (a = Co), t = x (b = Cx)
            if transposed:
                g, t, Cg, Ct, dims = dpre, x, Co, Cx, (N, Ho, Wo, None, H, W, None)
            else:
                g, t, Cg, Ct, dims = x, dpre, Cx, Co, (N, H, W, None, Ho, Wo, None)
            use3 = x3_wgrad
            mult = 2 if use3 else 1
            d = (dims[0], dims[1], dims[2], mult * Cg, dims[4], dims[5], mult * Ct)
            ws_bytes = _lib.lib().query('g2_conv_wgrad_tf32_workspace', *d, R, S, stride)
            stream_ctx = _GradStream(g, t) if _direct(w_param) else _contextlib.nullcontext()
            with stream_ctx:
                if ws_bytes > 0:
                    g2 = _split(g, Cg, 1) if use3 else g
                    t2 = _split(t, Ct, 1) if use3 else t
                    dwp = _new(x, R, S, mult * Cg, mult * Ct)
                    ws = _new(x, ws_bytes // 4)
                    _call('g2_conv_wgrad_tf32', g2, t2, dwp, ws, *d, R, S, stride, pad, 0)
                    if use3:     # [hi|lo] x [hi|lo]: hh + hl + lh (ll is below fp32 resolution)
                        dwp = dwp[:, :, :Cg, :Ct] + dwp[:, :, :Cg, Ct:] + dwp[:, :, Cg:, :Ct]
                else:            # shape outside the tensor-core tiles: exact fp32 SIMT
                    dwp = _new(x, R, S, Cg, Ct)
                    _call('g2_conv_wgrad_f32', g, t, dwp, dims[0], dims[1], dims[2], Cg, dims[4], dims[5], Ct, R, S, stride, pad, 0)
                # dwp [R,S,a,b] -> torch layout: conv [Co,Ci,R,S] (a = Cx, b = Co);  conv-transpose [Ci,Co,R,S] (a = Co, b = Cx)
                dwt = dwp.permute(3, 2, 0, 1)
                dwt = dwt[:Ci] if transposed else dwt[:, :Ci]
                if _direct(w_param):
                    w_param.grad.add_(dwt)
                else:
                    dw = dwt.contiguous()
                if has_b and ctx.needs_input_grad[2] and _direct(b_param):
                    _bias_grad(dpre, b_param, Co)
                    has_b = False
        if has_b and ctx.needs_input_grad[2]:
            db = _bias_grad(dpre, b_param, Co)
        _track_done(w_param, b_param if ctx.cfg[3] else None)
        return dx, dw, db, None, None, None, None, None


def _x3_conv_ok(x, w, stride, pad, transposed):
    """Full 3xTF32 route (_ConvX3: also the backward contractions; experiments, scripts/parity_report.py): TF32 mode, inside
    `precise()`, a backward part requested, and the tensor-core kernels cover the 3x-reduction shapes.  The default precise
    layer (forward only) goes through _Conv / _ConvT."""
    if not _X3['on'] or _PRECISION['mode'] != 'tf32' or not (_X3['dgrad'] or _X3['wgrad']):
        return False
    N, H, W, Cx = x.shape
    if transposed:
        Ci, Co, R, S = w.shape
        Ho = (H - 1) * stride - 2 * pad + R + stride - 1
        Wo = (W - 1) * stride - 2 * pad + S + stride - 1
    else:
        Co, Ci, R, S = w.shape
        Ho = (H + 2 * pad - R) // stride + 1
        Wo = (W + 2 * pad - S) // stride + 1
    mf, md = (1, 0) if transposed else (0, 1)
    return (Cx % 32 == 0 and _tc_ok(N, H, W, 3 * Cx, Ho, Wo, Co, R, S, stride, pad, mf)
            and _tc_ok(N, Ho, Wo, 3 * Co, H, W, Cx, R, S, stride, pad, md))


# ----------------------------------------------------------------------------------------- linear
class _Linear(Function):
    """y = act(x @ w.T + b);  x [M,K], w [N,K].  Large products run on the tensor cores (TF32); the many small ones
    that the tensor-core tile does not cover on the exact-fp32 SIMT GEMM, which takes transposed operands in place (no
    transpose copies), fuses the activation and can accumulate the weight gradient straight into `w.grad`."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        x = _c(x)
        wd = _c(w.detach())
        M, K = x.shape
        N = wd.shape[0]
        y = _new(x, M, N)
        ctx.x3 = _X3['on'] and (_X3['dgrad'] or _X3['wgrad'])
        if _X3['on'] and _X3['fwd'] and _gemm_tc_ok(M, N, 3 * K):
            w_hi, w_lo = _w_hi_lo(wd)
            _gemm_tc(_split(x, K, 0), torch.cat([w_hi, w_lo, w_hi], dim=1), b, y, M, N, 3 * K, act)      # 3xTF32
        elif _X3['on'] and _X3['fwd']:
            _call('g2_gemm_f32', x, wd, b, y, M, N, K, K, K, N, 0, 1, act, 0)                              # exact fp32
        elif _gemm_tc_ok(M, N, K):
            _gemm_tc(x, wd, b, y, M, N, K, act)      # deterministic split-K; bias + activation fused
        else:
            _call('g2_gemm_f32', x, wd, b, y, M, N, K, K, K, N, 0, 1, act, 0)
        ctx.save_for_backward(x, wd, y if act != ACT_NONE else None)
        ctx.params = (w, b)
        ctx.cfg = (act, b is not None)
        _track_use(w, b)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        w_param, b_param = ctx.params
        act, has_b = ctx.cfg
        M, K = x.shape
        N = w.shape[0]
        dpre = _act_bwd(_c(dy), y, act)
        dx = dw = db = None
        x3 = getattr(ctx, 'x3', False)
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if x3 and _gemm_tc_ok(M, K, 3 * N):
                wt_hi, wt_lo = _w_hi_lo(w.t().contiguous())
                _gemm_tc(_split(dpre, N, 0), torch.cat([wt_hi, wt_lo, wt_hi], dim=1), None, dx, M, K, 3 * N)
            elif x3:
                _call('g2_gemm_f32', dpre, w, None, dx, M, K, N, N, K, K, 0, 0, ACT_NONE, 0)
            elif _gemm_tc_ok(M, K, N):
                _gemm_tc(dpre, w.t().contiguous(), None, dx, M, K, N)
            else:
                _call('g2_gemm_f32', dpre, w, None, dx, M, K, N, N, K, K, 0, 0, ACT_NONE, 0)
        if ctx.needs_input_grad[1]:
            if _direct(w_param):
                with _GradStream(dpre, x):
                    if x3 and _gemm_tc_ok(N, K, 3 * M):
                        dwt = torch.empty_like(w)
                        xt_hi, xt_lo = _w_hi_lo(x.t().contiguous())
                        _gemm_tc(_split(dpre.t().contiguous(), M, 0), torch.cat([xt_hi, xt_lo, xt_hi], dim=1), None, dwt, N, K, 3 * M)
                        w_param.grad.add_(dwt)
                    elif x3:
                        _call('g2_gemm_f32', dpre, x, None, w_param.grad, N, K, M, N, K, K, 1, 0, ACT_NONE, 1)
                    elif _gemm_tc_ok(N, K, M):
                        dwt = torch.empty_like(w)
                        _gemm_tc(dpre.t().contiguous(), x.t().contiguous(), None, dwt, N, K, M)
                        w_param.grad.add_(dwt)
                    else:
                        _call('g2_gemm_f32', dpre, x, None, w_param.grad, N, K, M, N, K, K, 1, 0, ACT_NONE, 1)
                    if has_b and ctx.needs_input_grad[2] and _direct(b_param):
                        _bias_grad(dpre, b_param, N)
                        has_b = False
            elif x3 and _gemm_tc_ok(N, K, 3 * M):
                dw = torch.empty_like(w)
                xt_hi, xt_lo = _w_hi_lo(x.t().contiguous())
                _gemm_tc(_split(dpre.t().contiguous(), M, 0), torch.cat([xt_hi, xt_lo, xt_hi], dim=1), None, dw, N, K, 3 * M)
            elif x3:
                dw = torch.empty_like(w)
                _call('g2_gemm_f32', dpre, x, None, dw, N, K, M, N, K, K, 1, 0, ACT_NONE, 0)
            elif _gemm_tc_ok(N, K, M):
                dw = torch.empty_like(w)
                _gemm_tc(dpre.t().contiguous(), x.t().contiguous(), None, dw, N, K, M)
            else:
                dw = torch.empty_like(w)
                _call('g2_gemm_f32', dpre, x, None, dw, N, K, M, N, K, K, 1, 0, ACT_NONE, 0)
        if has_b and ctx.needs_input_grad[2]:
            db = _bias_grad(dpre, b_param, N)
        _track_done(w_param, b_param)
        return dx, dw, db, None


def _gemm_tc_ok(m_rows, n_out, k_red):
    """g2_gemm_tf32 needs the reduction dim % 32 == 0 and the output width in {32,64,128} or % 64 == 0.  Measured on
    B200: even for the LSTM-step sized products (a few MFLOP) the tensor-core kernel (~6-10 us, latency-bound) beats the
    un-pipelined SIMT GEMM (20-70 us), so every supported shape goes to it."""
    return (_PRECISION['mode'] == 'tf32' and k_red % 32 == 0 and k_red >= 64
            and (n_out in (32, 64, 128) or n_out % 64 == 0))


def _gemm_tc(a, w, bias, out, M, N, K, act=ACT_NONE):
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) on the tensor cores; split-K partials go to a workspace and are reduced in a
    fixed order (bitwise reproducible, no float atomics)."""
    ws_bytes = _lib.lib().query('g2_gemm_tf32_workspace', M, N, K)
    ws = _new(a, ws_bytes // 4) if ws_bytes > 0 else None
    _call('g2_gemm_tf32_ws', a, w, bias, out, ws, M, N, K, act)


_ZERO_ROWS = {}


def _zeros_row(like, n):
    key = (like.device, n)
    if key not in _ZERO_ROWS:
        _ZERO_ROWS[key] = torch.zeros(n, device=like.device, dtype=torch.float32)
    return _ZERO_ROWS[key]


def linear(x, w, b=None, act=None):
    return _Linear.apply(x, w, b, ACTS[act])


# ----------------------------------------------------------------------------------------- norm + post
class _NormPost(Function):
    """y [N,H,W,Cy] -> post(norm(y)).  post = gate (Cy = 2C, out C channels) or ReLU.
    params: (g0,b0) affine of the first half / whole, (g1,b1) affine of the second (gate) half."""

    @staticmethod
    def forward(ctx, y, g0, b0, g1, b1, rm0, rv0, rm1, rv1, mode, post, groups, training, eps, momentum, conv_bias=None):
        y = _c(y)
        ctx.conv_bias = conv_bias
        N, H, W, Cy = y.shape
        HW = H * W
        C = Cy // 2 if post == POST_GATE else Cy
        half = C if post == POST_GATE else Cy
        out = _new(y, N, H, W, C)
        if mode == NORM_NONE:
            _call('g2_norm_apply_f32', y, None, None, out, N, HW, C, 0, post)
            ctx.save_for_backward(y)
            ctx.cfg = (mode, post, groups, half)
            _track_use(conv_bias)
            return out
        Ns = 1 if mode == NORM_BATCH else N
        sums = None
        if not (mode == NORM_BATCH and not training):
            sums = _new(y, N, Cy, 2, dtype=torch.float64)
            _call('g2_norm_stats_f32', y, sums, N, HW, Cy)
        mean, rstd, scale, shift = (_new(y, Ns, Cy) for _ in range(4))
        _call('g2_norm_finalize_f32', sums, g0, b0, g1, b1, rm0, rv0, rm1, rv1, mean, rstd, scale, shift,
              N, HW, Cy, half, mode, groups, 1 if training else 0, eps, momentum)
        sn = 0 if mode == NORM_BATCH else Cy
        _call('g2_norm_apply_f32', y, scale, shift, out, N, HW, C, sn, post)
        ctx.save_for_backward(y, scale, shift, mean, rstd, g0, g1)
        ctx.cfg = (mode, post, groups, half)
        ctx.params = (g0, b0, g1, b1)
        if _NORM_DIRECT['on']:
            _track_use(g0, b0, g1, b1)
        _track_use(conv_bias)
        return out

    @staticmethod
    def backward(ctx, dout):
        mode, post, groups, half = ctx.cfg
        dout = _c(dout)
        cb = ctx.conv_bias
        want_cb = cb is not None and ctx.needs_input_grad[15]
        dcb = None

        def conv_bias_buffer(Cy):
            # the bias gradient of the producing convolution = column sums of dy, accumulated by the backward-apply kernel
            if _direct(cb):
                return cb.grad, None
            buf = torch.zeros(Cy, device=dout.device, dtype=torch.float32)
            return buf, buf
        if mode == NORM_NONE:
            (y,) = ctx.saved_tensors
            N, H, W, Cy = y.shape
            C = dout.shape[3]
            dy = torch.empty_like(y)
            if want_cb:
                buf, dcb = conv_bias_buffer(Cy)
                _call('g2_norm_bwd_apply_bias_f32', y, dout, None, None, None, None, None, None, dy, buf, N, H * W, C, 0, post)
            else:
                _call('g2_norm_bwd_apply_f32', y, dout, None, None, None, None, None, None, dy, N, H * W, C, 0, post)
            _track_done(cb)
            return (dy,) + (None,) * 14 + (dcb,)
        y, scale, shift, mean, rstd, g0, g1 = ctx.saved_tensors
        N, H, W, Cy = y.shape
        HW = H * W
        C = dout.shape[3]
        Ns = scale.shape[0]
        sn = 0 if mode == NORM_BATCH else Cy
        sums2 = _new(y, N, Cy, 2, dtype=torch.float64)
        _call('g2_norm_bwd_stats_f32', y, dout, scale, shift, mean, rstd, sums2, N, HW, C, sn, post)
        m1, m2 = _new(y, Ns, Cy), _new(y, Ns, Cy)
        params = [p_ for p_ in ctx.params if p_ is not None]
        if _NORM_DIRECT['on'] and params and all(_direct(p_) for p_ in params):
            # direct-gradient mode (experimental switch): the finalize kernel adds the affine-parameter gradients to param.grad
            # itself -- no four small tensors and no AccumulateGrad `add_` per norm layer.  The same stream orders repeated
            # uses of one layer (MONet's recurrent UNet); nothing else writes these .grad tensors.
            pg0, pb0, pg1, pb1 = ((p_.grad if p_ is not None else None) for p_ in ctx.params)
            _call('g2_norm_bwd_finalize_acc_f32', sums2, g0, g1, m1, m2, pg0, pb0, pg1, pb1, N, HW, Cy, half, mode, groups, 1)
            dg0 = db0 = dg1 = db1 = None
        else:
            dg0 = _new(y, half) if g0 is not None else None
            db0 = _new(y, half) if g0 is not None else None
            dg1 = _new(y, Cy - half) if g1 is not None else None
            db1 = _new(y, Cy - half) if g1 is not None else None
            _call('g2_norm_bwd_finalize_f32', sums2, g0, g1, m1, m2, dg0, db0, dg1, db1, N, HW, Cy, half, mode, groups)
        dy = torch.empty_like(y)
        if want_cb:
            buf, dcb = conv_bias_buffer(Cy)
            _call('g2_norm_bwd_apply_bias_f32', y, dout, scale, shift, mean, rstd, m1, m2, dy, buf, N, HW, C, sn, post)
        else:
            _call('g2_norm_bwd_apply_f32', y, dout, scale, shift, mean, rstd, m1, m2, dy, N, HW, C, sn, post)
        _track_done(*ctx.params, cb)
        return (dy, dg0, db0, dg1, db1) + (None,) * 10 + (dcb,)


def norm_post(y, g0=None, b0=None, g1=None, b1=None, rm0=None, rv0=None, rm1=None, rv1=None,
              mode=NORM_NONE, post=POST_GATE, groups=1, training=True, eps=1e-5, momentum=0.1, conv_bias=None):
    """conv_bias: the bias parameter of the convolution that produced y (called with bias_grad=False); its gradient -- the column
    sums of dy -- is then accumulated by the norm backward pass itself instead of a separate pass over dy."""
    return _NormPost.apply(y, g0, b0, g1, b1, rm0, rv0, rm1, rv1, mode, post, groups, training, eps, momentum, conv_bias)


# ----------------------------------------------------------------------------------------- SBP scan
class _SBPScan(Function):
    """logits [nl,B,...] -> log_m [K,B,...], log_s [nl+1,B,...] (log_s is a statistic: no gradient)."""

    @staticmethod
    def forward(ctx, logits, K):
        logits = _c(logits)
        nl = logits.shape[0]
        BP = logits[0].numel()
        log_m = _new(logits, K, *logits.shape[1:])
        log_s = _new(logits, nl + 1, *logits.shape[1:])
        _call('g2_sbp_scan_fwd_f32', logits, log_m, log_s, BP, K, nl)
        ctx.save_for_backward(logits)
        ctx.K = K
        ctx.mark_non_differentiable(log_s)
        return log_m, log_s

    @staticmethod
    def backward(ctx, dlog_m, _dlog_s):
        (logits,) = ctx.saved_tensors
        nl = logits.shape[0]
        d = torch.empty_like(logits)
        _call('g2_sbp_scan_bwd_f32', logits, _c(dlog_m), d, logits[0].numel(), ctx.K, nl)
        return d, None


def sbp_scan(logits, K):
    return _SBPScan.apply(logits, K)


# ----------------------------------------------------------------------------------------- comp pack
class _CompPack(Function):
    """x [B,3,H,W] (NCHW), log_m [K,B,1,H,W] -> [K*B,H,W,4] NHWC with channel 0 = log_m, 1..3 = x."""

    @staticmethod
    def forward(ctx, x, log_m, cp):
        x, log_m = _c(x), _c(log_m)
        K, B = log_m.shape[0], log_m.shape[1]
        H, W = x.shape[2], x.shape[3]
        out = _new(x, K * B, H, W, cp)
        _call('g2_comp_pack_f32', x, log_m, out, K, B, H * W, cp)
        ctx.shape = log_m.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        return None, dout[..., 0].reshape(ctx.shape).contiguous(), None


def comp_pack(x, log_m, cp=4):
    """cp = 4, or 32 (zero-padded) so the first encoder conv runs on the tensor cores."""
    return _CompPack.apply(x, log_m, cp)


def to_nhwc_padded(x, cp):
    """NCHW input image -> NHWC with channels zero-padded to cp (no gradient: x is data)."""
    x = _c(x.detach())
    N, C, H, W = x.shape
    y = _new(x, N, H, W, cp)
    _call('g2_nhwc_pad_f32', x, y, N, C, H * W, cp)
    return y


# ----------------------------------------------------------------------------------------- broadcast add
class _BcastAddAct(Function):
    """out[n,p,c] = act(a[n,c] + m[p,c])."""

    @staticmethod
    def forward(ctx, a, m, act):
        a, m = _c(a), _c(m)
        N, C = a.shape
        P = m.shape[0]
        out = _new(a, N, P, C)
        _call('g2_bcast_add_act_f32', a, m, out, N, P, C, act)
        ctx.save_for_backward(out)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        N, P, C = out.shape
        dpre = _act_bwd(_c(dout), out, ctx.act)
        da = dm = None
        if ctx.needs_input_grad[0]:
            da = _new(out, N, C)
            _call('g2_seg_colsum_f32', dpre, da, N, P, C)
        if ctx.needs_input_grad[1]:
            dm = _new(out, P, C)
            _call('g2_sum_dim0_f32', dpre, dm, N, P * C)
        return da, dm, None


def bcast_add_act(a, m, act=None):
    return _BcastAddAct.apply(a, m, ACTS[act])


# ----------------------------------------------------------------------------------------- 1x1 output head
class _Out1x1(Function):
    """h [N,H,W,Cin] NHWC -> out [N,nout,H,W] NCHW, sigmoid on the first nsig channels; w [nout,Cin,1,1]."""

    @staticmethod
    def forward(ctx, h, w, b, nsig):
        h = _c(h)
        N, H, W, Cin = h.shape
        nout = w.shape[0]
        w2 = w.detach().reshape(nout, Cin).contiguous()
        out = _new(h, N, nout, H, W)
        _call('g2_out1x1_fwd_f32', h, w2, b, out, N, H * W, Cin, nout, nsig)
        ctx.save_for_backward(h, w2, out)
        ctx.params = (w, b)
        ctx.cfg = (nsig, b is not None)
        if Cin % 32 == 0:
            _track_use(w, b)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, w2, out = ctx.saved_tensors
        nsig, has_b = ctx.cfg
        N, H, W, Cin = h.shape
        nout = w2.shape[0]
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        dpre4 = _new(h, N, H, W, 4)
        _call('g2_out1x1_bwd_f32', _c(dout), out, w2, dh, dpre4, N, H * W, Cin, nout, nsig)
        dw = db = None
        w_param, b_param = ctx.params
        if ctx.needs_input_grad[1] and Cin % 32 == 0 and _direct(w_param) and (not has_b or _direct(b_param)):
            # direct-gradient mode: head weight / bias gradients accumulate into param.grad on the gradient side stream
            with _GradStream(h, dpre4):
                dw4 = _new(h, 4, Cin)
                _call('g2_head_wgrad_f32', h, dpre4, dw4, N * H * W, Cin)
                w_param.grad.view(nout, Cin).add_(dw4[:nout])
                if has_b and ctx.needs_input_grad[2]:
                    b_param.grad.add_(_colsum(dpre4, 4, dpre4)[:nout])
            _track_done(w_param, b_param)
            return dh, None, None, None
        if ctx.needs_input_grad[1]:
            dw4 = _new(h, 4, Cin)
            if Cin % 32 == 0:
                _call('g2_head_wgrad_f32', h, dpre4, dw4, N * H * W, Cin)
            else:
                _call('g2_conv_wgrad_f32', h, dpre4, dw4, N, H, W, Cin, H, W, 4, 1, 1, 1, 0, 1)
            dw = dw4[:nout].reshape(nout, Cin, 1, 1).contiguous()
        if has_b and ctx.needs_input_grad[2]:
            db = _colsum(dpre4, 4, dpre4)[:nout].contiguous()
        return dh, dw, db, None


def out1x1(h, w, b=None, nsig=0):
    return _Out1x1.apply(h, w, b, nsig)


# ----------------------------------------------------------------------------------------- mixture loss
class _Mixture(Function):
    """Genesis.x_loss fused with recon (and log-softmax of mask logits when softmax=True).
    x [B,3,H,W], xr [K,B,3,H,W], lm [K,B,1,H,W], std [K] -> err [B], recon [B,3,H,W],
    log-softmax masks [K,B,1,H,W] (empty when softmax=False)."""

    @staticmethod
    def forward(ctx, x, xr, lm, std, softmax):
        x, xr, lm, std = _c(x), _c(xr), _c(lm), _c(std)
        K, B = lm.shape[0], lm.shape[1]
        P = x.shape[2] * x.shape[3]
        err = _new(x, B)
        recon = torch.empty_like(x)
        lse = torch.empty_like(x)
        lm_out = torch.empty_like(lm) if softmax else None
        _call('g2_mixture_fwd_f32', x, xr, lm, std, err, recon, lse, lm_out, K, B, P, 1 if softmax else 0, 3, 1)
        ctx.save_for_backward(x, xr, lm_out if softmax else lm, std, lse)
        ctx.softmax = softmax
        if not softmax:
            lm_out = _new(x, 0)
        ctx.mark_non_differentiable(recon, lm_out)
        return err, recon, lm_out

    @staticmethod
    def backward(ctx, gerr, _grecon, _glm):
        x, xr, lm, std, lse = ctx.saved_tensors
        K, B = lm.shape[0], lm.shape[1]
        P = x.shape[2] * x.shape[3]
        dxr = torch.empty_like(xr)
        dlm = torch.empty_like(lm)
        _call('g2_mixture_bwd_f32', x, xr, lm, std, lse, _c(gerr), dxr, dlm, K, B, P, 1 if ctx.softmax else 0, 3, 1, 1)
        return None, dxr, dlm, None, None


def mixture_nll(x, xr, lm, std, softmax=False):
    return _Mixture.apply(x, xr, lm, std, softmax)


class _MixturePixelwise(Function):
    """Genesis.x_loss(pixel_wise=True): per pixel and channel loss [B,3,H,W] = -log sum_k exp(log m_k + log N(x; xr_k, std_k))."""

    @staticmethod
    def forward(ctx, x, xr, lm, std):
        x, xr, lm, std = _c(x), _c(xr), _c(lm), _c(std)
        K, B = lm.shape[0], lm.shape[1]
        P = x.shape[2] * x.shape[3]
        err = _new(x, B)
        recon = torch.empty_like(x)
        lse = torch.empty_like(x)
        _call('g2_mixture_fwd_f32', x, xr, lm, std, err, recon, lse, None, K, B, P, 0, 3, 1)
        ctx.save_for_backward(x, xr, lm, std, lse)
        return -lse

    @staticmethod
    def backward(ctx, gpix):
        x, xr, lm, std, lse = ctx.saved_tensors
        K, B = lm.shape[0], lm.shape[1]
        P = x.shape[2] * x.shape[3]
        dxr = torch.empty_like(xr)
        dlm = torch.empty_like(lm)
        _call('g2_mixture_bwd_pix_f32', x, xr, lm, std, lse, _c(gpix), dxr, dlm, K, B, P, 3, 1, 1)
        return None, dxr, dlm, None


def mixture_nll_pixelwise(x, xr, lm, std):
    return _MixturePixelwise.apply(x, xr, lm, std)


class _MixturePacked(Function):
    """Mixture likelihood on a packed decoder output dec [K,B,4,H,W] (planes 0-2 = x_r, plane 3 = mask logit).
    mode 'softmax': masks = log_softmax over K of plane 3 (GENESIS-V2, genesisv2_config.py:213-223); gradients flow to
    all 4 planes.  mode 'given': masks = lm [K,B,1,H,W] (MONet, monet_config.py:92-105); gradients flow to planes 0-2
    and to lm, plane 3 gets zero here (its gradient comes from the mask KL)."""

    @staticmethod
    def forward(ctx, x, dec, lm, std, softmax):
        x, dec, std = _c(x), _c(dec), _c(std)
        K, B = dec.shape[0], dec.shape[1]
        P = x.shape[2] * x.shape[3]
        err = _new(x, B)
        recon = torch.empty_like(x)
        lse = torch.empty_like(x)
        if softmax:
            lm_out = _new(x, K, B, 1, x.shape[2], x.shape[3])
            _call('g2_mixture_fwd_f32', x, dec, dec.data_ptr() + 12 * P, std, err, recon, lse, lm_out, K, B, P, 1, 4, 4)
            ctx.save_for_backward(x, dec, lm_out, std, lse)
        else:
            lm = _c(lm)
            lm_out = _new(x, 0)
            _call('g2_mixture_fwd_f32', x, dec, lm, std, err, recon, lse, None, K, B, P, 0, 4, 1)
            ctx.save_for_backward(x, dec, lm, std, lse)
        ctx.softmax = softmax
        ctx.mark_non_differentiable(recon, lm_out)
        return err, recon, lm_out

    @staticmethod
    def backward(ctx, gerr, _grecon, _glm):
        x, dec, lm, std, lse = ctx.saved_tensors
        K, B = dec.shape[0], dec.shape[1]
        P = x.shape[2] * x.shape[3]
        if ctx.softmax:
            ddec = torch.empty_like(dec)
            _call('g2_mixture_bwd_f32', x, dec, lm, std, lse, _c(gerr), ddec, ddec.data_ptr() + 12 * P, K, B, P, 1, 4, 1, 4)
            return None, ddec, None, None, None
        ddec = torch.zeros_like(dec)
        dlm = torch.empty_like(lm)
        _call('g2_mixture_bwd_f32', x, dec, lm, std, lse, _c(gerr), ddec, dlm, K, B, P, 0, 4, 1, 1)
        return None, ddec, dlm, None, None


def mixture_nll_packed(x, dec, lm, std, softmax):
    return _MixturePacked.apply(x, dec, lm, std, softmax)


# ----------------------------------------------------------------------------------------- UNet resampling
class _Resample(Function):
    @staticmethod
    def forward(ctx, x, up):
        x = _c(x)
        N, H, W, C = x.shape
        Ho, Wo = (2 * H, 2 * W) if up else (H // 2, W // 2)
        y = _new(x, N, Ho, Wo, C)
        _call('g2_resample_f32', x, y, N, Ho, Wo, C, 1 if up else 0)
        ctx.up = up
        ctx.in_hw = (H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        N, _, _, C = dy.shape
        H, W = ctx.in_hw
        dx = _new(dy, N, H, W, C)
        _call('g2_resample_f32', dy, dx, N, H, W, C, 3 if ctx.up else 2)
        return dx, None


def down2(x):
    """F.interpolate(scale_factor=0.5, mode='nearest') on NHWC (reference unet.py:77-78)."""
    return _Resample.apply(x, False)


def up2(x):
    """F.interpolate(scale_factor=2.0, mode='nearest') on NHWC (reference unet.py:88-89)."""
    return _Resample.apply(x, True)


# ----------------------------------------------------------------------------------------- IC-SBP
class _ICSBP(Function):
    """colour [B,H,W,8] NHWC, u [B,1,H,W], log_sigma [] -> log_m [K,B,1,H,W], log_s [K,B,1,H,W], seed_idx [K-1,B]."""

    @staticmethod
    def forward(ctx, colour, u, log_sigma, K, kernel='gaussian'):
        colour, u = _c(colour), _c(u)
        B, H, W, CD = colour.shape
        P = H * W
        log_m = _new(colour, K, B, 1, H, W)
        log_s = _new(colour, K, B, 1, H, W)
        idx = torch.empty((K - 1, B), device=colour.device, dtype=torch.int32)
        ls = log_sigma.detach().reshape(1).float().contiguous()
        kt = ICSBP_KERNELS[kernel]
        if kt == 0:
            _call('g2_icsbp_fwd_f32', colour, u, ls, log_m, log_s, idx, B, P, K, CD)
        else:
            _call('g2_icsbp_kernel_fwd_f32', colour, u, ls, log_m, log_s, idx, B, P, K, CD, kt)
        ctx.save_for_backward(colour, ls, idx)
        ctx.K, ctx.kt = K, kt
        ctx.ls_dtype = log_sigma.dtype
        ctx.mark_non_differentiable(log_s, idx)
        return log_m, log_s, idx

    @staticmethod
    def backward(ctx, dlog_m, _dls, _didx):
        colour, ls, idx = ctx.saved_tensors
        B, H, W, CD = colour.shape
        dcol = torch.empty_like(colour)
        dsig = _new(colour, B)
        if ctx.kt == 0:
            _call('g2_icsbp_bwd_f32', colour, ls, idx, _c(dlog_m), dcol, dsig, B, H * W, ctx.K, CD)
        else:
            _call('g2_icsbp_kernel_bwd_f32', colour, ls, idx, _c(dlog_m), dcol, dsig, B, H * W, ctx.K, CD, ctx.kt)
        return dcol, None, dsig.sum().reshape(()).to(ctx.ls_dtype), None, None


ICSBP_KERNELS = {'gaussian': 0, 'laplacian': 1, 'epanechnikov': 2}       # reference modules/attention.py:146-153


class _ICSBPDynamic(Function):
    """dynamic_K form of _ICSBP: also returns n_masks [B] (int32); log_m slots k >= n_masks[b] hold -1e10."""

    @staticmethod
    def forward(ctx, colour, u, log_sigma, K, kernel):
        colour, u = _c(colour), _c(u)
        B, H, W, CD = colour.shape
        log_m = _new(colour, K, B, 1, H, W)
        log_s = _new(colour, K, B, 1, H, W)
        idx = torch.empty((K - 1, B), device=colour.device, dtype=torch.int32)
        n = torch.empty((B,), device=colour.device, dtype=torch.int32)
        ls = log_sigma.detach().reshape(1).float().contiguous()
        kt = ICSBP_KERNELS[kernel]
        _call('g2_icsbp_dynamic_fwd_f32', colour, u, ls, log_m, log_s, idx, n, B, H * W, K, CD, kt)
        ctx.save_for_backward(colour, ls, idx, n)
        ctx.K, ctx.kt = K, kt
        ctx.ls_dtype = log_sigma.dtype
        ctx.mark_non_differentiable(log_s, idx, n)
        return log_m, log_s, idx, n

    @staticmethod
    def backward(ctx, dlog_m, _dls, _didx, _dn):
        colour, ls, idx, n = ctx.saved_tensors
        B, H, W, CD = colour.shape
        dcol = torch.empty_like(colour)
        dsig = _new(colour, B)
        _call('g2_icsbp_dynamic_bwd_f32', colour, ls, idx, n, _c(dlog_m), dcol, dsig, B, H * W, ctx.K, CD, ctx.kt)
        return dcol, None, dsig.sum().reshape(()).to(ctx.ls_dtype), None, None


def icsbp_dynamic(colour, u, log_sigma, K, kernel='gaussian'):
    return _ICSBPDynamic.apply(colour, u, log_sigma, K, kernel)


def icsbp(colour, u, log_sigma, K, kernel='gaussian'):
    return _ICSBP.apply(colour, u, log_sigma, K, kernel)


# ----------------------------------------------------------------------------------------- masked pooling
class _MaskedPool(Function):
    """f [B,H,W,C] NHWC, log_m [K,B,1,H,W] -> num [K,B,C] = sum_p m f, msum [K,B] = sum_p m."""

    @staticmethod
    def forward(ctx, f, log_m):
        f, log_m = _c(f), _c(log_m)
        B, H, W, C = f.shape
        K = log_m.shape[0]
        num = _new(f, K, B, C)
        msum = _new(f, K, B)
        _call('g2_masked_pool_fwd_f32', f, log_m, num, msum, B, H * W, C, K)
        ctx.save_for_backward(f, log_m)
        return num, msum

    @staticmethod
    def backward(ctx, dnum, dmsum):
        f, log_m = ctx.saved_tensors
        B, H, W, C = f.shape
        K = log_m.shape[0]
        df = torch.empty_like(f)
        dlm = torch.empty_like(log_m)
        _call('g2_masked_pool_bwd_f32', f, log_m, _c(dnum), _c(dmsum), df, dlm, B, H * W, C, K)
        return df, dlm


def masked_pool(f, log_m):
    return _MaskedPool.apply(f, log_m)


# ----------------------------------------------------------------------------------------- MONet loss head
class _MonetLoss(Function):
    """MONet reconstruction + mask-KL terms on the packed decoder output (reference monet_config.py:85-107).
    x [B,3,H,W]; dec [K,B,4,H,W] (planes 0-2 = x_r after the pixel-bound sigmoid, plane 3 = mask logits);
    lm [K,B,1,H,W] inferred log masks; std [K].  Returns err [B], kl_m [B], recon, log_m_r [K,B,1,H,W]."""

    @staticmethod
    def forward(ctx, x, dec, lm, std):
        x, dec, lm, std = _c(x), _c(dec), _c(lm), _c(std)
        K, B = dec.shape[0], dec.shape[1]
        P = x.shape[2] * x.shape[3]
        err, kl = _new(x, B), _new(x, B)
        recon, lse = torch.empty_like(x), torch.empty_like(x)
        lmr = torch.empty_like(lm)
        _call('g2_mixture_fwd_f32', x, dec, lm, std, err, recon, lse, None, K, B, P, 0, 4, 1)
        _call('g2_mask_kl_fwd_f32', lm, dec.data_ptr() + 12 * P, lmr, kl, K, B, P, 1, 4)
        ctx.save_for_backward(x, dec, lm, std, lse)
        ctx.mark_non_differentiable(recon, lmr)
        return err, kl, recon, lmr

    @staticmethod
    def backward(ctx, gerr, gkl, _grecon, _glmr):
        x, dec, lm, std, lse = ctx.saved_tensors
        K, B = dec.shape[0], dec.shape[1]
        P = x.shape[2] * x.shape[3]
        ddec = torch.empty_like(dec)
        dlm = torch.empty_like(lm)
        _call('g2_mixture_bwd_f32', x, dec, lm, std, lse, _c(gerr), ddec, dlm, K, B, P, 0, 4, 1, 1)
        _call('g2_mask_kl_bwd_f32', lm, dec.data_ptr() + 12 * P, _c(gkl), dlm, ddec.data_ptr() + 12 * P, K, B, P,
              1, 4, 1, 4, 1)
        return None, ddec, dlm, None


def monet_loss(x, dec, lm, std):
    return _MonetLoss.apply(x, dec, lm, std)


# ----------------------------------------------------------------------------------------- stand-alone mask KL
class _MaskKL(Function):
    """MONet.kl_m_loss (monet_config.py:157-170) between given log-masks lm [K,B,1,H,W] and the log-softmax over K of mask
    logits -> kl [B].  The logits are plane 3 of a packed decoder output dec [K,B,4,H,W] (GENESIS-V2 `klm_loss`,
    genesisv2_config.py:171-176) or a plain [K,B,1,H,W] tensor (MONet prior_mode='scope': the stick-breaking log-masks,
    which the log-softmax leaves unchanged because they already sum to one).
    detach=True (`detach_mr_in_klm`, the V2 default): no gradient to the logits."""

    @staticmethod
    def forward(ctx, lm, logits, detach):
        lm, logits = _c(lm), _c(logits)
        K, B = logits.shape[0], logits.shape[1]
        cs = logits.shape[2]
        assert cs in (1, 4)
        P = logits.shape[3] * logits.shape[4]
        off = 12 * P if cs == 4 else 0
        kl = _new(lm, B)
        lmr = torch.empty_like(lm)
        _call('g2_mask_kl_fwd_f32', lm, logits.data_ptr() + off, lmr, kl, K, B, P, 1, cs)
        ctx.save_for_backward(lm, logits)
        ctx.detach = detach
        return kl

    @staticmethod
    def backward(ctx, gkl):
        lm, logits = ctx.saved_tensors
        K, B = logits.shape[0], logits.shape[1]
        cs = logits.shape[2]
        P = logits.shape[3] * logits.shape[4]
        off = 12 * P if cs == 4 else 0
        dlm = torch.empty_like(lm)
        dlg = torch.zeros_like(logits)
        _call('g2_mask_kl_bwd_f32', lm, logits.data_ptr() + off, _c(gkl), dlm, dlg.data_ptr() + off, K, B, P, 1, cs, 1, cs, 0)
        return dlm, (None if ctx.detach else dlg), None


def mask_kl(lm, dec, detach=True):
    return _MaskKL.apply(lm, dec, detach)


# ----------------------------------------------------------------------------------------- fused latent path
# One kernel each for the LSTM cell, the Gaussian head (to_sigma + rsample), the prior head (tanh / to_prior_sigma) and the
# Monte-Carlo KL, forward and backward (csrc/latent.cu).  set_fused_latent(False) selects the ATen formulation in holders.py
# (the op contract; tests compare the two).
import os as _os
_FUSED = {'on': _os.environ.get('G2_FUSED_LATENT', '1') == '1'}      # default on (validated on a B200 in round 2: +5 % images/s)


def set_fused_latent(on):
    _FUSED['on'] = bool(on)


def fused_latent():
    return _FUSED['on']


class _LSTMCell(Function):
    """gx, gh [B,4H] (the two gate GEMMs incl. biases), c_prev [B,H] or None -> h, c."""

    @staticmethod
    def forward(ctx, gx, gh, c_prev):
        gx, gh = _c(gx), _c(gh)
        B, H = gx.shape[0], gx.shape[1] // 4
        cp = _c(c_prev) if c_prev is not None else None
        h, c = _new(gx, B, H), _new(gx, B, H)
        _call('g2_lstm_cell_fwd_f32', gx, gh, cp, h, c, B, H)
        ctx.save_for_backward(gx, gh, cp, c)
        return h, c

    @staticmethod
    def backward(ctx, dh, dc):
        gx, gh, cp, c = ctx.saved_tensors
        B, H = c.shape
        dg = torch.empty_like(gx)
        dcp = torch.empty_like(c) if cp is not None else None
        _call('g2_lstm_cell_bwd_f32', gx, gh, cp, c, _c(dh) if dh is not None else None, _c(dc) if dc is not None else None,
              dg, dcp, B, H)
        return dg, dg, dcp


def lstm_cell(gx, gh, c_prev=None):
    return _LSTMCell.apply(gx, gh, c_prev)


class _GaussHead(Function):
    """lo [B,2D] = (mu | raw), eps [B,D] -> z, mu, sigma (sigma = softplus(raw + 0.5) + 1e-8, z = mu + sigma * eps)."""

    @staticmethod
    def forward(ctx, lo, eps):
        lo, eps = _c(lo), _c(eps)
        B, D = lo.shape[0], lo.shape[1] // 2
        z, mu, sigma = _new(lo, B, D), _new(lo, B, D), _new(lo, B, D)
        _call('g2_gauss_head_fwd_f32', lo, eps, z, mu, sigma, B, D)
        ctx.save_for_backward(lo, eps)
        return z, mu, sigma

    @staticmethod
    def backward(ctx, dz, dmu, dsigma):
        lo, eps = ctx.saved_tensors
        B, D = eps.shape
        dlo = torch.empty_like(lo)
        _call('g2_gauss_head_bwd_f32', lo, eps, _c(dz) if dz is not None else None, _c(dmu) if dmu is not None else None,
              _c(dsigma) if dsigma is not None else None, dlo, B, D)
        return dlo, None


def gauss_head(lo, eps):
    return _GaussHead.apply(lo, eps)


class _PriorHead(Function):
    """lo [B,2D] = (a | b) -> pmu = tanh(a) (or a), psigma = sigmoid(b + 4) + 1e-4."""

    @staticmethod
    def forward(ctx, lo, use_tanh):
        lo = _c(lo)
        B, D = lo.shape[0], lo.shape[1] // 2
        pmu, psig = _new(lo, B, D), _new(lo, B, D)
        _call('g2_prior_head_fwd_f32', lo, pmu, psig, B, D, 1 if use_tanh else 0)
        ctx.save_for_backward(pmu, psig)
        ctx.use_tanh = use_tanh
        return pmu, psig

    @staticmethod
    def backward(ctx, dpmu, dpsig):
        pmu, psig = ctx.saved_tensors
        B, D = pmu.shape
        dlo = _new(pmu, B, 2 * D)
        _call('g2_prior_head_bwd_f32', pmu, psig, _c(dpmu) if dpmu is not None else None,
              _c(dpsig) if dpsig is not None else None, dlo, B, D, 1 if ctx.use_tanh else 0)
        return dlo, None


def prior_head(lo, use_tanh=True):
    return _PriorHead.apply(lo, use_tanh)


class _MCKL(Function):
    """kl[b] = sum_d log N(z;mu,sigma) - log N(z;pmu,psigma)   (pmu None: standard-normal prior)."""

    @staticmethod
    def forward(ctx, z, mu, sigma, pmu, psigma):
        z, mu, sigma = _c(z), _c(mu), _c(sigma)
        pmu = _c(pmu) if pmu is not None else None
        psigma = _c(psigma) if psigma is not None else None
        B, D = z.shape
        kl = _new(z, B)
        _call('g2_mc_kl_fwd_f32', z, mu, sigma, pmu, psigma, kl, B, D)
        ctx.save_for_backward(z, mu, sigma, pmu, psigma)
        return kl

    @staticmethod
    def backward(ctx, dkl):
        z, mu, sigma, pmu, psigma = ctx.saved_tensors
        B, D = z.shape
        dz, dmu, dsig = torch.empty_like(z), torch.empty_like(z), torch.empty_like(z)
        dpmu = torch.empty_like(z) if pmu is not None else None
        dpsig = torch.empty_like(z) if pmu is not None else None
        _call('g2_mc_kl_bwd_f32', z, mu, sigma, pmu, psigma, _c(dkl), dz, dmu, dsig, dpmu, dpsig, B, D)
        return dz, dmu, dsig, dpmu, dpsig


def mc_kl(z, mu, sigma, pmu=None, psigma=None):
    return _MCKL.apply(z, mu, sigma, pmu, psigma)
