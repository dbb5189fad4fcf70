"""cuobjdump -sass genesis_b200/lib/libgenesis_b200.so | python scripts/sass_summary.py > profiles/rNN_sass_tcgen05_tma.txt
Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md): UTCHMMA = tcgen05.mma,
LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit."""
import collections
import re
import subprocess
import sys

PAT = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|UTMAPF[.\w]*|UTCATOMSWS[.\w]*)')


def main():
    fn, counts = None, collections.OrderedDict()
    for line in sys.stdin:
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
        elif fn:
            for mm in PAT.finditer(line):
                counts[fn][mm.group(1)] += 1
    print('# SASS evidence of the built libgenesis_b200.so, per kernel (cuobjdump -sass | python scripts/sass_summary.py).')
    print('# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit')
    tot = collections.Counter()
    for fn, c in counts.items():
        if c:
            name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()
            print('%-90s %s' % (name[:90], ' '.join('%s=%d' % kv for kv in sorted(c.items()))))
            tot.update(c)
    print('\nTOTAL ' + ' '.join('%s=%d' % kv for kv in sorted(tot.items())))


if __name__ == '__main__':
    main()
