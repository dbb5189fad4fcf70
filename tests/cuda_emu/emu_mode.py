"""TEST INFRASTRUCTURE: run the GPU tests on the CPU (G2_EMU=1 in the environment, picked up by tests/conftest.py).

Every C-ABI call goes to the CPU emulation of the kernel sources (emu_lib.EmuLib), 'cuda' devices are rewritten to 'cpu' by a
TorchFunctionMode, `.cuda()` is the identity, tensors answer is_cuda = True (the product refuses CPU tensors), and the few
torch.cuda entry points the package and the tests touch (streams, synchronize, is_available) are inert stand-ins.  Meant for
checking GPU tests and GPU-only code paths BEFORE spending GPU time on them; it is slow (a 64x64 model step takes minutes)
and proves nothing about the hardware or about performance."""
import contextlib
import os
import sys

import torch
from torch.overrides import TorchFunctionMode

HERE = os.path.dirname(os.path.abspath(__file__))
_STATE = {}


def _is_cuda(d):
    if isinstance(d, str):
        return d.startswith('cuda')
    if isinstance(d, torch.device):
        return d.type == 'cuda'
    return False


class HostTensor(torch.Tensor):
    """What `.cpu()` returns in emulation mode: the one kind of tensor that answers is_cuda = False, so that the product's
    `CPU tensors are rejected` behaviour stays testable."""

    @property
    def is_cuda(self):
        return False


class CpuAsCuda(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if _is_cuda(kwargs.get('device')):
            kwargs['device'] = 'cpu'
        name = getattr(func, '__name__', '')
        if name == 'cuda' and args and isinstance(args[0], torch.Tensor):
            return args[0].as_subclass(torch.Tensor) if isinstance(args[0], HostTensor) else args[0]
        if name == 'cpu' and args and isinstance(args[0], torch.Tensor) and type(args[0]) is torch.Tensor:
            return args[0].as_subclass(HostTensor)
        if name in ('to', 'pin_memory') and args and isinstance(args[0], torch.Tensor):
            if name == 'pin_memory':
                return args[0]
            args = tuple('cpu' if _is_cuda(a) else a for a in args)
        if name == 'record_stream':
            return None
        return func(*args, **kwargs)


class _Stream(object):
    cuda_stream = 0

    def wait_stream(self, other):
        pass

    def synchronize(self):
        pass


def enable():
    if _STATE:
        return
    sys.path.insert(0, HERE)
    import emu_lib
    from genesis_b200 import _lib, ops
    _lib._LIB = emu_lib.EmuLib()
    ops.set_side_streams(False)
    stream = _Stream()
    torch.cuda.is_available = lambda: True
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: stream
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.device_count = lambda: 1
    torch.cuda.current_device = lambda: 0
    torch.cuda.is_current_stream_capturing = lambda: False
    torch.cuda.graphs.is_current_stream_capturing = lambda: False
    torch.cuda._lazy_init = lambda: None
    torch.Tensor.is_cuda = property(lambda self: True)
    mode = CpuAsCuda()
    mode.__enter__()
    _STATE['mode'] = mode
