"""CPU: include/genesis_b200.h against the extern "C" definitions in genesis_b200/csrc/*.cu.

ctypes marshals by the HEADER (genesis_b200/_lib.py), most sources do not include it, and C linkage has no mangling, so a
definition whose parameters drift from the header (order, type, count) would link and mis-marshal silently.  This parses
both sides and requires, for every exported g2_* function: same return type, same number of parameters, same
pointer-ness and base type per position (g2_stream_t == cudaStream_t), and the same parameter NAMES (names are how a
swapped pair of ints would show)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def strip_comments(text):
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    return re.sub(r'//[^\n]*', ' ', text)


def norm_param(p):
    p = p.replace('__restrict__', ' ').replace('cudaStream_t', 'g2_stream_t').strip()
    ptr = p.count('*')
    toks = p.replace('*', ' ').split()
    name = toks[-1]
    base = ' '.join(t for t in toks[:-1] if t != 'const')
    return base, ptr, name


def parse(text, terminator):
    out = {}
    for m in re.finditer(r'(?:^|\n)\s*(?:extern\s+"C"\s+)?(int|long)\s+(g2_\w+)\s*\(([^)]*)\)\s*' + terminator, strip_comments(text)):
        args = m.group(3).strip()
        params = [] if args in ('', 'void') else [norm_param(a) for a in args.split(',')]
        out[m.group(2)] = (m.group(1), params)
    return out


def test_definitions_match_the_header():
    header = parse(open(os.path.join(ROOT, 'include', 'genesis_b200.h')).read(), ';')
    defs = {}
    for path in sorted(glob.glob(os.path.join(ROOT, 'genesis_b200', 'csrc', '*.cu'))):
        for name, sig in parse(open(path).read(), r'\{').items():
            assert name not in defs, name + ' defined twice'
            defs[name] = (os.path.basename(path),) + sig
    assert len(header) >= 55
    missing = sorted(set(header) - set(defs))
    assert not missing, 'declared but not defined: %s' % missing
    for name, (ret, params) in header.items():
        src, dret, dparams = defs[name]
        assert ret == dret, (name, src, ret, dret)
        assert len(params) == len(dparams), '%s (%s): header has %d parameters, definition %d' % (name, src, len(params), len(dparams))
        for i, (h, d) in enumerate(zip(params, dparams)):
            assert h[:2] == d[:2], '%s (%s) parameter %d: header %r, definition %r' % (name, src, i, h, d)
            assert h[2] == d[2], '%s (%s) parameter %d is named %r in the header and %r in the definition' % (name, src, i, h[2], d[2])
    # exported g2_* definitions that the header does not declare would be unreachable through the ABI
    extra = sorted(n for n in set(defs) - set(header) if not n.startswith('g2_debug_'))
    assert not extra, 'defined but not declared: %s' % extra
