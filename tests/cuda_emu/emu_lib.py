"""TEST INFRASTRUCTURE: a stand-in for genesis_b200._lib._Lib that dispatches every C-ABI call to the CPU emulation of the .cu
file defining it (tests/cuda_emu/build_emu.py), so that genesis_b200.ops -- the autograd Functions, their argument
marshalling, strides, workspaces and saved tensors -- can be exercised on CPU tensors against torch autograd.  Installed by
tests only (monkeypatching genesis_b200._lib._LIB); the product library is never involved."""
import ctypes
import glob
import os
import re

import torch

import build_emu
from genesis_b200 import _lib


def _definitions():
    out = {}
    for path in sorted(glob.glob(os.path.join(build_emu.CSRC, '*.cu'))):
        text = open(path).read()
        for m in re.finditer(r'(?:^|\n)\s*(?:extern\s+"C"\s+)?(?:int|long)\s+(g2_\w+)\s*\(', text):
            out.setdefault(m.group(1), os.path.basename(path))
    return out


class EmuLib(object):
    def __init__(self):
        self.protos = _lib.parse_header()
        self.where = _definitions()
        self.libs, self._fn = {}, {}
        self.launches = 0
        self.calls = []

    def _get(self, name):
        if name not in self._fn:
            src = self.where[name]
            if src == 'igemm_halo.cu':      # one instance of the halo kernel's switches: the object igemm_tc.cu links it into
                src = 'igemm_tc.cu'
            if src not in self.libs:
                self.libs[src] = ctypes.CDLL(build_emu.build(src))
            fn = getattr(self.libs[src], name)
            fn.restype = _lib.RESTYPES.get(name, ctypes.c_int)
            fn.argtypes = [t for t, _, _ in self.protos[name]]
            self._fn[name] = fn
        return self._fn[name]

    def query(self, name, *args):
        return self._get(name)(*args)

    def call(self, name, *args):
        sig = self.protos[name]
        if len(args) != len(sig) - 1:
            raise TypeError('%s expects %d arguments (+stream), got %d' % (name, len(sig) - 1, len(args)))
        conv = []
        for a, (t, is_ptr, an) in zip(args, sig):
            if is_ptr:
                if a is None:
                    conv.append(None)
                elif torch.is_tensor(a):
                    if not a.is_contiguous():
                        raise RuntimeError('%s: argument %s must be contiguous' % (name, an))
                    conv.append(a.data_ptr())
                else:
                    conv.append(int(a))
            else:
                if torch.is_tensor(a):
                    raise TypeError('%s: tensor passed for scalar argument %s' % (name, an))
                conv.append(a)
        conv.append(None)
        rc = self._get(name)(*conv)
        self.launches += 1
        self.calls.append(name)
        if rc != 0:
            raise RuntimeError('%s failed with code %d' % (name, rc))


def install(monkeypatch):
    """Route genesis_b200.ops to the emulation for the duration of a test; returns the EmuLib (records the calls made)."""
    import contextlib

    from genesis_b200 import ops
    emu = EmuLib()
    monkeypatch.setattr(_lib, '_LIB', emu)
    monkeypatch.setattr(ops, 'side_streams_enabled', lambda: False)

    class _Stream(object):
        cuda_stream = 0

        def wait_stream(self, other):
            pass
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
    return emu
