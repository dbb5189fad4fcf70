"""Count the ATen operators (~ CUDA kernel launches outside our C ABI) and the C-ABI calls of one training step, on the CPU
emulation of the kernels (no GPU needed): python scripts/aten_census.py MODEL K [--fused-latent].
View-only operators (reshape / chunk / permute / ...) are not counted: they launch nothing."""
import collections
import os
import sys

import pytest
import torch
from torch.utils._python_dispatch import TorchDispatchMode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'cuda_emu')):
    sys.path.insert(0, p)

import cpu_ops_mock  # noqa: E402
import emu_lib  # noqa: E402
import util_parity as U  # noqa: E402
from genesis_b200 import ops  # noqa: E402
from oracle import functional as O  # noqa: E402
from oracle import synth  # noqa: E402
from test_oracle_golden import build_engine_model  # noqa: E402

VIEWS = {'view', '_unsafe_view', 'reshape', 'chunk', 'split', 'split_with_sizes', 'slice', 'select', 'unbind', 'permute', 'transpose',
         't', 'expand', 'as_strided', 'detach', 'alias', 'unsqueeze', 'squeeze', 'narrow', 'unfold', 'view_as', 'lift_fresh',
         '_to_copy', 'empty', 'empty_like', 'empty_strided', 'new_empty', 'new_empty_strided', 'is_same_size', 'sym_size', 'stride', 'size'}


class Census(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.ops = collections.Counter()
        self.sites = collections.Counter()
        self.trace = set(a[8:].split(',')[0] for a in sys.argv if a.startswith('--trace=')) | set(sum([a[8:].split(',') for a in sys.argv if a.startswith('--trace=')], []))

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = func.overloadpacket.__name__
        if name not in VIEWS:
            self.ops[name] += 1
            if name in self.trace:
                import traceback
                fr = [f for f in traceback.extract_stack() if 'genesis_b200' in f.filename]
                key = '%s:%d' % (os.path.basename(fr[-1].filename), fr[-1].lineno) if fr else 'autograd engine'
                self.sites[(name, key)] += 1
        return func(*args, **(kwargs or {}))


def main():
    model, K = sys.argv[1], int(sys.argv[2])
    mp = pytest.MonkeyPatch()
    emu = emu_lib.install(mp)
    torch.Tensor.is_cuda = property(lambda self: True)      # the plug-ins and holders branch on it
    ops.set_precision(os.environ.get('CENSUS_PRECISION', 'fp32'))
    ops.set_fused_latent('--fused-latent' in sys.argv)
    sys.argv = [a for a in sys.argv]
    m, cfg = build_engine_model(model, K, 64)
    m.train()
    x = torch.from_numpy(synth.GENERATORS['multid'](1, 64, 5)[0])
    m.set_noise_tape(O.NoiseTape(seed=3))
    for p_ in m.parameters():
        p_.grad = torch.zeros_like(p_)
    ops.set_side_streams(False)
    ops.set_direct_grad(True)
    with Census() as c:
        out = m(x.as_subclass(cpu_ops_mock.AsCuda))
        U.engine_total_loss(out[1]).backward()
    ops.set_direct_grad(False)
    print('%s K=%d %s: %d ATen operators, %d C-ABI calls' % (model, K, ' '.join(sys.argv[3:]) or 'default', sum(c.ops.values()), len(emu.calls)))
    print('  top ATen:', ', '.join('%s x%d' % kv for kv in c.ops.most_common(14)))
    for (name, site), n in c.sites.most_common(24):
        print('    %-8s %-40s x%d' % (name, site, n))


if __name__ == '__main__':
    main()
