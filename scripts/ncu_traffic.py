"""Turn an ncu CSV (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum over the conv kernels of
one eager training step) into profiles/<name>.json: per kernel symbol the launches, summed device time and the MEAN DRAM
bytes per launch -- the `roofline.traffic` figure bench.py reports beside the algorithmic bytes / flops.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
        -k regex:'conv_halo|conv_tc|wgrad_tc' -s <launches of the warm-up steps> -c <launches of one step> \\
        --csv --log-file gpurun_out/traffic.csv python scripts/profile_step.py
    python scripts/ncu_traffic.py gpurun_out/traffic.csv profiles/r01_traffic.json"""
import collections
import csv
import json
import sys

ENTRY = {'conv_halo_kernel': 'g2_conv_halo_tf32', 'conv_halo_persistent_kernel': 'g2_conv_halo_tf32', 'conv_tc_kernel': 'g2_conv_igemm_tf32',
         'wgrad_tc_kernel': 'g2_conv_wgrad_tf32_to', 'wgrad_halo_kernel': 'g2_conv_wgrad_tf32_to'}


def main(src, dst, steps=1):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[start]
    ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        d = per.setdefault(r[0], {'kernel': r[ki]})
        v = float(r[vi].replace(',', ''))
        unit = r[ui]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6}.get(unit, 1)
        d[r[mi]] = v * scale
    out = {}
    for d in per.values():
        sym = next((e for k, e in ENTRY.items() if k in d['kernel']), None)
        if sym is None:
            continue
        o = out.setdefault(sym, {'launches': 0, 'ns': 0.0, 'dram_bytes': 0.0})
        o['launches'] += 1
        o['ns'] += d.get('gpu__time_duration.sum', 0.0)
        o['dram_bytes'] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    for o in out.values():
        o['launches'] /= steps; o['ns'] /= steps; o['dram_bytes'] /= steps          # per ONE step
        o['dram_bytes_per_launch'] = o['dram_bytes'] / max(1, o['launches'])
        o['us_per_launch_under_ncu'] = o['ns'] / 1e3 / max(1, o['launches'])
    out['_how'] = ('ncu dram__bytes_read.sum + dram__bytes_write.sum over the conv / wgrad kernel launches of eager training steps '
                   '(scripts/profile_step.py: %d steps captured, figures per ONE step; cold, serialised)' % steps)
    json.dump(out, open(dst, 'w'), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
