"""Minimal Forge experiment-tools surface the hot path's callers use (reference:
forge/forge/experiment_tools.py:246-282 `load`, :444-457 `fprint`).  `load(path, cfg)` imports a
config file BY PATH (module name = file basename) and calls its module-level `load(cfg)`."""
from __future__ import print_function
import importlib.util
import os
import sys

from . import flags as _flags  # noqa: F401  (train.py:142 reads fet._flags.FLAGS)

FPRINT_FILE = None


def import_by_path(path):
    path = os.path.abspath(path)
    name = os.path.splitext(os.path.basename(path))[0]
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules.setdefault(name, module)
    spec.loader.exec_module(module)
    return module


def load(conf_path, *args, **kwargs):
    return import_by_path(conf_path).load(*args, **kwargs)


def parse_flags():
    _flags.FLAGS._parse_flags()
    return _flags.FLAGS


def fprint(*args, **kwargs):
    printonly = kwargs.pop('printonly', False)
    print(*args, **kwargs)
    if FPRINT_FILE is not None and not printonly:
        with open(FPRINT_FILE, 'a') as f:
            print(*args, file=f)
