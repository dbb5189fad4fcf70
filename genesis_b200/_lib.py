"""ctypes binding of libgenesis_b200.so.  Prototypes are parsed from include/genesis_b200.h so the
Python side cannot drift from the C ABI.  There is NO fallback: if the library is missing the import
of any op raises."""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'genesis_b200.h')
LIB_PATH = os.path.join(HERE, 'lib', 'libgenesis_b200.so')

_CTYPES = {'int': ctypes.c_int, 'long': ctypes.c_long, 'float': ctypes.c_float, 'double': ctypes.c_double,
           'size_t': ctypes.c_size_t, 'g2_stream_t': ctypes.c_void_p, 'uint64_t': ctypes.c_uint64,
           'int64_t': ctypes.c_int64, 'unsigned': ctypes.c_uint}


RESTYPES = {}


def parse_header(path=HEADER):
    """-> {name: [(ctype, is_pointer, argname), ...]} for every `int g2_*(...)` prototype."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    text = re.sub(r'//[^\n]*', ' ', text)
    protos = {}
    for m in re.finditer(r'\b(int|long)\s+(g2_\w+)\s*\(([^)]*)\)\s*;', text):
        name, args = m.group(2), m.group(3).strip()
        RESTYPES[name] = ctypes.c_long if m.group(1) == 'long' else ctypes.c_int
        sig = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    sig.append((ctypes.c_void_p, True, a.split('*')[-1].strip()))
                else:
                    toks = a.replace('const', ' ').split()
                    sig.append((_CTYPES[toks[0]], False, toks[-1]))
        protos[name] = sig
    return protos


class _Lib(object):
    def __init__(self, lib_path=LIB_PATH, header=HEADER):
        if not os.path.exists(lib_path):
            raise RuntimeError('%s is not built (%s). Run `python -m genesis_b200.build`; '
                               'there is no fallback path.' % (os.path.basename(lib_path), lib_path))
        self.cdll = ctypes.CDLL(lib_path)
        self.protos = parse_header(header)
        self.launches = 0
        self._fn = {}
        for name, sig in self.protos.items():
            fn = getattr(self.cdll, name)          # raises AttributeError if a declared symbol is missing
            fn.restype = RESTYPES.get(name, ctypes.c_int)
            fn.argtypes = [t for t, _, _ in sig]
            self._fn[name] = fn

    def query(self, name, *args):
        """Plain host-side query entry point (no stream, no launch); returns the int result."""
        return self._fn[name](*args)

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
                    if not a.is_cuda:
                        raise RuntimeError('%s: argument %s must be a CUDA tensor' % (name, an))
                    if not a.is_contiguous():
                        raise RuntimeError('%s: argument %s must be contiguous' % (name, an))
                    if a.device.index != torch.cuda.current_device():
                        # the launch goes to the CURRENT device's stream: a tensor of another GPU would be a silent fault
                        raise RuntimeError('%s: argument %s lives on cuda:%d but the current device is cuda:%d'
                                           % (name, an, a.device.index, torch.cuda.current_device()))
                    conv.append(a.data_ptr())
                else:
                    conv.append(int(a))
            else:
                conv.append(a)
        conv.append(torch.cuda.current_stream().cuda_stream)
        rc = self._fn[name](*conv)
        self.launches += 1
        if rc != 0:
            raise RuntimeError('%s failed with code %d (%s)' % (
                name, rc, 'bad argument' if rc == -1 else 'unsupported' if rc == -2 else 'cudaError'))


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def call(name, *args):
    lib().call(name, *args)


_PROBE = None


def probe():
    """TEST INFRASTRUCTURE: the tcgen05 probe library (tests/probe/, built by build.build_probe) behind the same marshalling."""
    global _PROBE
    if _PROBE is None:
        root = os.path.dirname(HERE)
        _PROBE = _Lib(os.path.join(root, 'tests', 'probe', 'libgenesis_b200_probe.so'),
                      os.path.join(root, 'tests', 'probe', 'genesis_b200_probe.h'))
    return _PROBE
