"""ORACLE / test infrastructure -- import the REAL reference: the live checkout in the build container, or the
byte-identical copy under oracle/_ref (oracle/build_ref.py; git-ignored, shipped to the GPU box) where /root/reference does
not exist.  Used by oracle/make_golden.py, by tests that skip when neither is present, and by bench.py's CPU reference arm.  The reference needs
`attrdict`, `forge`, `tensorflow`, `simplejson`: the stand-ins under
genesis_b200/compat are put on sys.path.  Noise is replayed from a NoiseTape by patching
torch.distributions.normal._standard_normal and Tensor.uniform_ for the duration of a forward."""
import contextlib
import os
import sys

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _default_root():
    env = os.environ.get('GENESIS_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isdir('/root/reference'):
        return '/root/reference'
    return os.path.join(_REPO, 'oracle', '_ref')


REF_ROOT = _default_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'genesis_config.py'))


def _setup():
    compat = os.path.join(_REPO, 'genesis_b200', 'compat')
    for p in (REF_ROOT, compat):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_reference(model, cfg_dict, seed=0):
    """Returns the reference nn.Module for model in {genesis, genesisv2, monet, vae}."""
    _setup()
    import importlib
    mod = importlib.import_module('models.%s_config' % model)
    from attrdict import AttrDict
    torch.manual_seed(seed)
    return mod.load(AttrDict(dict(cfg_dict)))


@contextlib.contextmanager
def replay_noise(tape):
    import torch.distributions.normal as tdn
    import torch.distributions.utils as tdu
    orig_sn = tdn._standard_normal
    orig_uniform = torch.Tensor.uniform_
    orig_normal = torch.normal

    def normal(mean, std, *a, **k):         # Normal.sample() -> torch.normal(loc, scale): the sample() paths
        if torch.is_tensor(mean) and torch.is_tensor(std):
            return mean + std * tape.normal(tuple(mean.shape), mean.dtype).to(mean.device)
        return orig_normal(mean, std, *a, **k)

    def sn(shape, dtype, device):
        return tape.normal(tuple(shape), dtype).to(device)

    def uni(self, *a, **k):
        self.copy_(tape.uniform(tuple(self.shape), self.dtype))
        return self

    tdn._standard_normal = sn
    tdu._standard_normal = sn
    torch.Tensor.uniform_ = uni
    torch.normal = normal
    try:
        yield
    finally:
        tdn._standard_normal = orig_sn
        tdu._standard_normal = orig_sn
        torch.Tensor.uniform_ = orig_uniform
        torch.normal = orig_normal
