"""CPU: static check of every C-ABI call site in the package against include/genesis_b200.h.

The ctypes layer checks the argument COUNT at call time, but a call site that only runs on the GPU (or only in a
non-default variant) would first fail there.  This walks the package's source with `ast`, finds every
`_call('g2_…', …)` / `_lib.call('g2_…', …)` / `.query('g2_…', …)` and compares the number of positional arguments with the
prototype (the stream is appended by the wrapper for compute entry points; query entry points take none).  Call sites that
splat a tuple (`*dims`) are counted through the tuple's known length where it is a literal in the same function, otherwise
they are listed as unchecked and must stay few."""
import ast
import glob
import os

from genesis_b200 import _lib

ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'genesis_b200')


def call_sites():
    for path in sorted(glob.glob(os.path.join(ROOT, '**', '*.py'), recursive=True)):
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            if not isinstance(node, ast.Call) or not node.args:
                continue
            f = node.func
            fname = f.id if isinstance(f, ast.Name) else f.attr if isinstance(f, ast.Attribute) else None
            if fname not in ('_call', 'call', 'query'):
                continue
            a0 = node.args[0]
            if not (isinstance(a0, ast.Constant) and isinstance(a0.value, str) and a0.value.startswith('g2_')):
                continue
            yield path, node.lineno, fname, a0.value, node.args[1:]


def test_every_call_site_matches_its_prototype():
    protos = _lib.parse_header()
    checked, unchecked, names = 0, [], set()
    for path, line, fname, name, args in call_sites():
        where = '%s:%d %s' % (os.path.relpath(path, ROOT), line, name)
        assert name in protos, where + ' is not declared in include/genesis_b200.h'
        names.add(name)
        if any(isinstance(a, ast.Starred) for a in args):
            unchecked.append(where)
            continue
        sig = protos[name]
        has_stream = sig[-1][2] == 'stream'
        want = len(sig) - 1 if has_stream else len(sig)
        if fname == 'query':
            assert not has_stream, where + ': query() on a compute entry point'
        else:
            assert has_stream, where + ': call() on a host-only query entry point'
        assert len(args) == want, '%s: %d arguments, prototype takes %d' % (where, len(args), want)
        checked += 1
    assert checked >= 60, checked
    assert len(unchecked) <= 16, unchecked


def test_splatted_call_sites_by_hand():
    """The conv / wgrad call sites splat `dims` (7 ints) and stride tuples; pin their arithmetic here."""
    protos = _lib.parse_header()
    # g2_conv_wgrad_tf32_to(g, t, dw, ws, *dims[7], R, S, stride, pad, *strides[3], *lims[2], accumulate)
    assert len(protos['g2_conv_wgrad_tf32_to']) - 1 == 4 + 7 + 4 + 3 + 2 + 1
    # g2_conv_wgrad_tf32(g, t, dwp, ws, *dims[7], R, S, stride, pad, outT)
    assert len(protos['g2_conv_wgrad_tf32']) - 1 == 4 + 7 + 5
    # g2_conv_wgrad_f32(g, t, dwp, *dims[7], R, S, stride, pad, outT)
    assert len(protos['g2_conv_wgrad_f32']) - 1 == 3 + 7 + 5
    # g2_conv_wgrad_tf32_workspace(*dims[7], R, S, stride) -- host-only
    assert len(protos['g2_conv_wgrad_tf32_workspace']) == 10
