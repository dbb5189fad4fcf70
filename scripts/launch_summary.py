"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table: launches, summed device time, share; ATen share.
    python scripts/launch_summary.py gpurun_out/r02_launches.csv [header text] > profiles/r02_launches_summary.txt"""
import collections
import csv
import sys


def main(src, note=''):
    rows = list(csv.reader(open(src, errors='replace')))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[start]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    mi = hdr.index('Metric Name')
    agg = collections.OrderedDict()
    n = 0
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[mi] != 'gpu__time_duration.sum':
            continue
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', '')) * scale
        n += 1
    tot = sum(a[1] for a in agg.values())
    aten = [(k, a) for k, a in agg.items() if 'at::' in k or 'cub::' in k or 'nccl' in k.lower()]
    at_n, at_us = sum(a[0] for _, a in aten), sum(a[1] for _, a in aten)
    if note:
        print(note)
    print('%d launches, summed device time %.2f ms (per-launch times under ncu are cold-cache and serialised: compare SHARES); '
          'ATen / cub kernels: %d launches = %.2f ms (%.1f %%)\n' % (n, tot / 1e3, at_n, at_us / 1e3, 100 * at_us / tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print('%-112s %6d %10.1f us %5.1f%%' % (k[:112], a[0], a[1], 100 * a[1] / tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
