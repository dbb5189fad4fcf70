"""genesis_b200 -- B200-native engine for the GENESIS / GENESIS-V2 / MONet per-slot inference-and-decode
hot path, behind the reference's Forge `load(cfg)` plug-in API.

  genesis_b200/csrc/           hand-written sm_100a CUDA kernels + the C ABI (include/genesis_b200.h)
  genesis_b200/_lib.py         ctypes binding (no fallback: fails loudly when the .so is missing)
  genesis_b200/ops.py          autograd Functions over the C ABI
  genesis_b200/model_configs/  plug-in files mirroring reference models/*_config.py (flags + load(cfg))
  genesis_b200/compat/         Forge / attrdict stand-ins the reference's callers import
"""
import os
import sys

_COMPAT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'compat')


def enable_compat():
    """Put the Forge / attrdict / simplejson / tensorboardX stand-ins on sys.path (idempotent)."""
    if _COMPAT not in sys.path:
        sys.path.append(_COMPAT)


__version__ = '0.1.0'
