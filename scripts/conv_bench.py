"""Micro-benchmark of the TF32 implicit-GEMM entry points on the conv shapes of the BASELINE configs, through the C ABI.

    python scripts/conv_bench.py [out.json] [--wgrad] [--ab] [--x3] [--reps 10] [--only name,name]     (--x3: the 3xTF32 entry point; TF/s counts the useful third)

CUDA events on the launch stream, 2 warm-up calls, each timed call preceded by an L2 flush (256 MB memset) that is
outside the event pair.  Prints TF/s (2*MACs / time) per shape.  Used under ncu for the roofline `traffic` figure."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from genesis_b200 import _lib  # noqa: E402

# name: (mode, N, H, W, Ci, Co, R, stride, pad)   mode 0 = conv-like gather, 1 = conv-transpose-like
SHAPES = {
    'c2_bdec_fwd70': (0, 320, 72, 72, 32, 32, 3, 1, 0),
    'c2_bdec_fwd64': (0, 320, 66, 66, 32, 32, 3, 1, 0),
    'c2_bdec_dgrad': (1, 320, 68, 68, 32, 32, 3, 1, 0),
    'c2_att64_fwd': (1, 320, 64, 64, 32, 64, 5, 1, 2),
    'c2_att64_dgrad': (0, 320, 64, 64, 64, 32, 5, 1, 2),
    'c2_att32_fwd': (1, 320, 32, 32, 32, 64, 5, 1, 2),
    'c2_att16_fwd': (1, 320, 16, 16, 64, 128, 5, 1, 2),
    'c2_att_up64_fwd': (1, 320, 32, 32, 32, 64, 5, 2, 2),
    'c2_att_up64_dgrad': (0, 320, 64, 64, 64, 32, 5, 2, 2),
    'c2_enc64_s2': (0, 64, 64, 64, 32, 64, 5, 2, 2),
    'c3_unet64': (0, 128, 64, 64, 64, 64, 3, 1, 1),
    'c3_unet32': (0, 128, 32, 32, 64, 64, 3, 1, 1),
    'c3_unet8': (0, 128, 8, 8, 128, 128, 3, 1, 1),
    'c3_dec_up64': (1, 896, 32, 32, 64, 64, 5, 2, 2),
    'c3_dec_up64_dgrad': (0, 896, 64, 64, 64, 64, 5, 2, 2),
    'c5_unet128': (0, 64, 128, 128, 32, 32, 3, 1, 1),
    'c5_bdec_fwd': (0, 224, 136, 136, 32, 32, 3, 1, 0),
}


def out_hw(mode, H, W, R, s, p):
    if mode == 0:
        return (H + 2 * p - R) // s + 1, (W + 2 * p - R) // s + 1
    return (H - 1) * s - 2 * p + R + (s - 1), (W - 1) * s - 2 * p + R + (s - 1)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    reps = 10
    only = None
    for i, a in enumerate(sys.argv):
        if a == '--reps':
            reps = int(sys.argv[i + 1])
            args = [x for x in args if x != sys.argv[i + 1]]
        if a == '--only':
            only = sys.argv[i + 1].split(',')
            args = [x for x in args if x != sys.argv[i + 1]]
    wgrad = '--wgrad' in sys.argv
    ab = '--ab' in sys.argv
    out_path = args[0] if args else None
    dev = torch.device('cuda', 0)
    lib = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, (mode, N, H, W, Ci, Co, R, s, p) in SHAPES.items():
        if only and name not in only:
            continue
        Ho, Wo = out_hw(mode, H, W, R, s, p)
        torch.manual_seed(0)
        x = torch.randn(N, H, W, Ci, device=dev)
        w = torch.randn(R * R, Co, Ci, device=dev) * 0.05
        b = torch.randn(Co, device=dev)
        out = torch.empty(N, Ho, Wo, Co, device=dev)
        flops = 2.0 * N * (Ho * Wo if mode == 0 else H * W) * Co * Ci * R * R
        if mode == 0 and s == 2:
            pass
        if mode == 1 and s == 2:
            flops = 2.0 * N * H * W * Co * Ci * R * R
        if lib.query('g2_conv_tf32_supported', N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode) != 1:
            print('%-20s unsupported' % name)
            continue

        x3 = '--x3' in sys.argv and lib.query('g2_conv_halo_supported', N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode) == 1
        w2 = torch.cat([w, w * 1e-4], dim=2).contiguous() if x3 else None          # [w_hi | w_lo] pack of the 3xTF32 entry point

        def run():
            if x3:
                _lib.call('g2_conv_halo_x3_tf32', x, w2, b, out, N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, 0)
            else:
                _lib.call('g2_conv_igemm_tf32', x, w, b, out, N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, 0)

        io_bytes = 4.0 * (x.numel() + out.numel() + w.numel())
        for halo in ((0, 1) if ab else (1,)):
            lib.query('g2_conv_halo_enable', halo)
            routed = lib.query('g2_conv_halo_supported', N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode)
            ms = []
            for i in range(reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ms.append(e0.elapsed_time(e1))
            ms.sort()
            med = ms[len(ms) // 2]
            row = dict(name=name, kind='conv', kernel='halo' if routed else 'tile', shape=[mode, N, H, W, Ci, Co, R, s, p], ms=med,
                       ms_min=ms[0], tflops=flops / med / 1e9, gbs=io_bytes / med / 1e6)
            rows.append(row)
            print('%-20s %-4s %8.3f ms (min %.3f) %7.1f TF/s  %7.0f GB/s algorithmic I/O' % (
                name, row['kernel'], med, ms[0], row['tflops'], row['gbs']))
        lib.query('g2_conv_halo_enable', 1)
        if wgrad and s == 1 or wgrad and mode == 0:
            # weight gradient of the same layer: g = input-side tensor, t = output-side tensor
            if mode == 0:
                dims = (N, H, W, Ci, Ho, Wo, Co)
                g, t, outT = x, out, 0
            else:
                dims = (N, Ho, Wo, Co, H, W, Ci)
                g, t, outT = out, x, 1
            ws_bytes = lib.query('g2_conv_wgrad_tf32_workspace', *dims, R, R, s)
            if ws_bytes > 0:
                ws = torch.empty(ws_bytes // 4, device=dev)
                dwp = torch.empty(R, R, Ci, Co, device=dev)
                ms = []
                for i in range(reps + 2):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _lib.call('g2_conv_wgrad_tf32', g, t, dwp, ws, *dims, R, R, s, p, outT)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 2:
                        ms.append(e0.elapsed_time(e1))
                ms.sort()
                med = ms[len(ms) // 2]
                rows.append(dict(name=name + '_wgrad', kind='wgrad', shape=[mode, N, H, W, Ci, Co, R, s, p], ms=med, ms_min=ms[0],
                                 tflops=flops / med / 1e9, gbs=4.0 * (x.numel() + out.numel()) / med / 1e6))
                print('%-20s %8.3f ms (min %.3f) %7.1f TF/s   wgrad' % (name, med, ms[0], flops / med / 1e9))
        del x, w, out
    if out_path:
        json.dump(rows, open(out_path, 'w'), indent=1)




def timeline(name):
    """Per-CTA phase durations (SM clocks) of one halo launch: where a CTA's lifetime goes."""
    import ctypes
    mode, N, H, W, Ci, Co, R, s, p = SHAPES[name]
    Ho, Wo = out_hw(mode, H, W, R, s, p)
    dev = torch.device('cuda', 0)
    lib = _lib.lib()
    x = torch.randn(N, H, W, Ci, device=dev); w = torch.randn(R * R, Co, Ci, device=dev) * 0.05; b = torch.randn(Co, device=dev)
    out = torch.empty(N, Ho, Wo, Co, device=dev)
    buf = torch.zeros(1 << 14, 64, dtype=torch.int64, device=dev)
    for i in range(3):
        if i == 2:
            lib.query('g2_conv_halo_debug', ctypes.c_void_p(buf.data_ptr()))
        _lib.call('g2_conv_igemm_tf32', x, w, b, out, N, H, W, Ci, Ho, Wo, Co, R, R, s, p, mode, 0)
    torch.cuda.synchronize()
    lib.query('g2_conv_halo_debug', None)
    t = buf.cpu()
    t = t[t[:, 0] > 0].double()
    names = ['prologue', 'wait first operands', 'MMA issue (all taps)', 'drain to accumulators complete', 'epilogue']
    d = [t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4]]
    print('%s: %d CTAs, lifetime mean %.0f clk' % (name, t.shape[0], (t[:, 5] - t[:, 0]).mean()))
    for n_, v in zip(names, d):
        print('   %-32s mean %8.0f  p10 %8.0f  p90 %8.0f clk' % (n_, v.mean(), v.kthvalue(max(1, int(0.1 * len(v))))[0], v.kthvalue(max(1, int(0.9 * len(v))))[0]))
    ntap = R * R if s == 1 else 9
    ntap = min(ntap, 28)
    w_ = t[:, 8:8 + 2 * ntap:2]
    i_ = t[:, 9:9 + 2 * ntap:2]
    okc = (w_ > 0).all(1) & (i_ > 0).all(1)
    if okc.any():
        w_, i_, base = w_[okc], i_[okc], t[okc, 2:3]
        print('   per tap (clk since first operands landed): weight-ready ' + ' '.join('%d' % v for v in (w_ - base).mean(0).tolist()))
        print('   per tap: MMAs issued                                    ' + ' '.join('%d' % v for v in (i_ - base).mean(0).tolist()))
    # concurrency: CTAs per SM over the kernel
    span = (t[:, 5].max() - t[:, 0].min())
    print('   kernel span (max end - min start, mixed SM clocks) %.0f clk; sum lifetimes / (148 * span) = %.2f CTAs resident per SM'
          % (span, (t[:, 5] - t[:, 0]).sum() / (148 * span)))


if '--timeline' in sys.argv:
    for nm in sys.argv[sys.argv.index('--timeline') + 1].split(','):
        timeline(nm)
elif __name__ == '__main__':
    main()
