"""CPU: the C-ABI library builds, loads and exports every symbol include/genesis_b200.h declares."""
import ctypes
import os

from genesis_b200 import _lib


# host-only queries / switches: no launch, no stream
QUERIES = ('g2_abi_version', 'g2_conv_tf32_supported', 'g2_conv_wgrad_tf32_workspace', 'g2_conv_halo_enable',
           'g2_conv_halo_supported', 'g2_conv_halo_plan', 'g2_conv_halo_debug', 'g2_conv_wgrad_halo_plan', 'g2_gemm_tf32_workspace')


def test_header_parses():
    protos = _lib.parse_header()
    assert len(protos) >= 20
    for name, sig in protos.items():
        assert name.startswith('g2_')
        if name not in QUERIES:
            assert sig[-1][2] == 'stream', name     # every compute entry point takes the stream last


def test_library_exports_all_declared_symbols(built_lib):
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.parse_header():
        assert hasattr(cdll, name), 'missing export: ' + name
    assert cdll.g2_abi_version() >= 1


def test_no_torch_types_in_abi():
    text = open(_lib.HEADER).read()
    for bad in ('at::', 'torch::', 'Tensor', 'c10::'):
        assert bad not in text


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, 'genesis_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, os.path.join(dirpath, f)
