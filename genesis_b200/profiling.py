"""Per-entry-point device timing of the C-ABI calls with CUDA events on the launching stream, plus the
ALGORITHMIC flops / bytes of each call (the roofline numerators of DESIGN.md section 5)."""
import collections

import torch

from . import _lib


def algo_cost(name, a):
    """(flops, bytes) a call must perform / move at minimum, from its arguments (pointer args included)."""
    f = b = 0
    if name in ('g2_conv_igemm_f32', 'g2_conv_igemm_tf32', 'g2_conv_halo_tf32', 'g2_conv_halo_x3_tf32'):
        o = 5 if name.endswith('f32') and not name.endswith('tf32') else 4
        N, Hi, Wi, Ci, Ho, Wo, Co, R, S, stride, pad, mode = a[o:o + 12]
        if mode == 0:
            f = 2.0 * N * Ho * Wo * Co * Ci * R * S
        else:
            f = 2.0 * N * Hi * Wi * Ci * Co * R * S
        b = 4.0 * (N * Hi * Wi * Ci + N * Ho * Wo * Co + R * S * Ci * Co)
    elif name in ('g2_conv_wgrad_f32', 'g2_conv_wgrad_tf32', 'g2_conv_wgrad_tf32_to'):
        o = 3 if name == 'g2_conv_wgrad_f32' else 4
        N, Hg, Wg, Cg, Ht, Wt, Ct, R, S = a[o:o + 9]
        f = 2.0 * N * Ht * Wt * Cg * Ct * R * S
        b = 4.0 * (N * Hg * Wg * Cg + N * Ht * Wt * Ct + R * S * Cg * Ct)
    elif name in ('g2_gemm_f32', 'g2_gemm_tf32', 'g2_gemm_tf32_ws'):
        M, N, K = a[5:8] if name == 'g2_gemm_tf32_ws' else a[4:7]    # (A, B, bias, C, [ws,] M, N, K, ...)
        f = 2.0 * M * N * K
        b = 4.0 * (M * K + N * K + M * N)
    elif name == 'g2_norm_stats_f32':
        N, HW, C = a[2:5]
        b = 4.0 * N * HW * C
    elif name == 'g2_norm_apply_f32':
        N, HW, C, sn, post = a[4:9]
        b = 4.0 * N * HW * C * (3 if post == 0 else 2)
    elif name == 'g2_norm_bwd_stats_f32':
        N, HW, C, sn, post = a[7:12]
        b = 4.0 * N * HW * C * (3 if post == 0 else 2)
    elif name in ('g2_norm_bwd_apply_f32', 'g2_norm_bwd_apply_bias_f32'):
        N, HW, C, sn, post = a[10:15] if name.endswith('bias_f32') else a[9:14]
        b = 4.0 * N * HW * C * (5 if post == 0 else 3)
    elif name == 'g2_mixture_fwd_f32':
        K, B, P = a[8:11]
        b = 4.0 * B * P * ((3 + 4 * K) + 6)
    elif name == 'g2_mixture_bwd_f32':
        K, B, P = a[8:11]
        b = 4.0 * B * P * ((6 + 4 * K) + 4 * K)
    elif name == 'g2_act_bwd_f32':
        b = 12.0 * a[3]
    elif name == 'g2_act_bwd_bias_f32':
        b = 12.0 * a[4] * a[5]
    elif name == 'g2_colsum_f32':
        b = 4.0 * a[2] * a[3]
    elif name == 'g2_head_wgrad_f32':
        b = 4.0 * a[3] * (a[4] + 4)
        f = 2.0 * a[3] * a[4] * 4
    elif name == 'g2_sbp_scan_fwd_f32':
        BP, K, nl = a[3:6]
        b = 4.0 * BP * (nl + K + nl + 1)
    elif name == 'g2_sbp_scan_bwd_f32':
        BP, K, nl = a[3:6]
        b = 4.0 * BP * (nl + K + nl)
    elif name == 'g2_comp_pack_f32':
        K, B, P, cp = a[3:7]
        b = 4.0 * B * P * (3 + K + K * cp)
    elif name == 'g2_nhwc_pad_f32':
        N, C, P, cp = a[2:6]
        b = 4.0 * N * P * (C + cp)
    elif name == 'g2_layout_f32':
        N, C, P = a[2:5]
        b = 8.0 * N * C * P
    elif name == 'g2_resample_f32':
        N, Ho, Wo, C, mode = a[2:7]
        # (Ho, Wo) is the OUTPUT map: 0 = nearest x0.5 (read 1 of 4 inputs), 1 = nearest x2, 2 = adjoint of x0.5, 3 = adjoint of x2
        b = 4.0 * N * Ho * Wo * C * {0: 2.0, 1: 1.25, 2: 1.25, 3: 5.0}[mode]
    elif name in ('g2_icsbp_fwd_f32', 'g2_icsbp_kernel_fwd_f32', 'g2_icsbp_dynamic_fwd_f32'):
        o = 7 if name == 'g2_icsbp_dynamic_fwd_f32' else 6
        B, P, K, CD = a[o:o + 4]
        b = 4.0 * B * P * (CD + 1 + 2 * K)
    elif name in ('g2_icsbp_bwd_f32', 'g2_icsbp_kernel_bwd_f32', 'g2_icsbp_dynamic_bwd_f32'):
        o = 7 if name == 'g2_icsbp_dynamic_bwd_f32' else 6
        B, P, K, CD = a[o:o + 4]
        b = 4.0 * B * P * (2 * CD + K)
    elif name == 'g2_masked_pool_fwd_f32':
        B, P, C, K = a[4:8]
        b = 4.0 * B * P * (C + K)
    elif name == 'g2_masked_pool_bwd_f32':
        B, P, C, K = a[6:10]
        b = 4.0 * B * P * (2 * C + 2 * K)
    elif name in ('g2_mask_kl_fwd_f32',):
        K, B, P = a[4:7]
        b = 4.0 * B * P * 3 * K
    elif name in ('g2_mask_kl_bwd_f32',):
        K, B, P = a[5:8]
        b = 4.0 * B * P * 4 * K
    elif name == 'g2_pack_conv_weight_f32':
        Co, Ci, Cp, RS = a[3:7]
        b = 4.0 * RS * Co * (Ci + 2 * Cp)
    elif name in ('g2_rmsprop_f32', 'g2_sgd_f32'):
        b = 20.0 * a[3]
    elif name == 'g2_bcast_add_act_f32':
        N, P, C = a[3:6]
        b = 4.0 * N * P * C
    elif name in ('g2_seg_colsum_f32',):
        N, P, C = a[2:5]
        b = 4.0 * N * P * C
    elif name == 'g2_sum_dim0_f32':
        b = 4.0 * a[2] * a[3]
    elif name == 'g2_out1x1_fwd_f32':
        N, P, Cin, nout = a[4:8]
        b = 4.0 * N * P * (Cin + nout)
        f = 2.0 * N * P * Cin * nout
    elif name == 'g2_out1x1_bwd_f32':
        N, P, Cin, nout = a[5:9]
        b = 4.0 * N * P * (Cin + 2 * nout + 4)
        f = 2.0 * N * P * Cin * nout
    elif name == 'g2_adam_f32':
        b = 28.0 * a[4]
    return f, b


class Profiler(object):
    """with Profiler() as prof: ...; prof.table() -> rows sorted by device time."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        self._lib = _lib.lib()
        self._orig = self._lib.call
        lib = self._lib

        def call(name, *args):
            s = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            s.record()
            self._orig(name, *args)
            e.record()
            shape = tuple(x for x in args if isinstance(x, int))
            self.records.append((name, shape, s, e, algo_cost(name, args)))

        lib.call = call
        return self

    def __exit__(self, *exc):
        self._lib.call = self._orig
        torch.cuda.synchronize()
        return False

    def table(self, by_shape=False):
        agg = collections.OrderedDict()
        for name, shape, s, e, (f, b) in self.records:
            key = (name, shape) if by_shape else name
            r = agg.setdefault(key, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
            r['ms'] += s.elapsed_time(e)
            r['calls'] += 1
            r['flops'] += f
            r['bytes'] += b
        rows = [dict(key=k, **v) for k, v in agg.items()]
        rows.sort(key=lambda r: -r['ms'])
        return rows
