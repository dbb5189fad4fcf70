"""tcgen05.mma kind::tf32 cost table on a B200 (g2_debug_umma_rate): clocks per M=128 x N x K=8 instruction with both operands in
shared memory, vs N, the A start-row alignment (0 = same row, 8 = aligned moves, 1/3/71 = the halo kernels' tap shifts) and the
number of resident CTAs per SM.  python scripts/umma_rate.py > gpurun_out/r02_umma_rate.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from genesis_b200 import _lib  # noqa: E402


def rate(N, shift, n_acc, ctas, b_tiles=1, n_mma=4096, a_rows=512):
    out = torch.zeros(148 * ctas, dtype=torch.int64, device='cuda')
    for _ in range(2):
        _lib.probe().call('g2_debug_umma_rate', out, N, n_mma, shift, n_acc, a_rows, ctas, b_tiles)
    torch.cuda.synchronize()
    o = out.double()
    return o.mean().item() / n_mma, o.max().item() / n_mma


def main():
    print('clocks per tcgen05.mma.kind::tf32 M=128 x N x K=8 (math floor N/2 x 128*8/.. = N*0.54 clk at 1.1 PF nominal), mean over CTAs')
    print('%-5s %-6s %-6s %-5s %-8s %10s %10s %12s' % ('N', 'shift', 'n_acc', 'ctas', 'b_tiles', 'clk/MMA', 'max', 'B/clk/SM'))
    for N in (32, 64, 128, 256):
        for shift in (0, 1, 71):
            for ctas in (1, 2):
                n_acc = max(1, min(4, 256 // N))
                for b_tiles in (1, 4):
                    if N * 128 * b_tiles + 512 * 128 > 98 * 1024:
                        continue
                    m, mx = rate(N, shift, n_acc, ctas, b_tiles)
                    byts = (4096 + 32 * N) * ctas / m
                    print('%-5d %-6d %-6d %-5d %-8d %10.1f %10.1f %12.1f' % (N, shift, n_acc, ctas, b_tiles, m, mx, byts))
    print('-- one accumulator (dependent chain) vs rotating accumulators, N=32, shift 1')
    for n_acc in (1, 2, 4, 8):
        m, mx = rate(32, 1, n_acc, 1)
        print('n_acc %d: %.1f clk/MMA' % (n_acc, m))


    print('-- two issuing threads in ONE CTA (warps 0 and 1, own accumulators): clocks per MMA of the CTA')
    for N in (32, 64):
        for ctas in (1, 2):
            m, mx = rate(N, 1, -2, ctas)
            print('N %d ctas/SM %d: %.1f clk/MMA per CTA  (one issuer: %.1f)' % (N, ctas, m, rate(N, 1, 2, ctas)[0]))


if __name__ == '__main__':
    main()
