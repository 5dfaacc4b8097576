"""The C-ABI library loads on a box without a GPU and exports every symbol that
include/cae_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from cnn_autoencoder_b200 import _cabi as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'cae_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(cae_[a-z0-9_]+)\s*\(', txt)))


def test_header_and_binding_agree():
    assert _declared() == sorted(C.SYMBOLS)


def test_library_exports_every_symbol():
    lib = ctypes.CDLL(C.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert C.lib().cae_abi_version() == 1


def test_struct_layout_matches_header():
    # sizes implied by the header on LP64: cae_tensor 24 B, cae_conv_desc 152 B
    assert ctypes.sizeof(C.Tensor) == 24
    assert ctypes.sizeof(C.ConvDesc) == 24 + 3 * 24 + 16 + 6 * 4 + 8
    assert ctypes.sizeof(C.EbTables) == 8 + 8 + 8 + 8 + 8 + 40 + 8


def test_error_text_is_reported():
    d = C.ConvDesc()
    d.kind = 99
    rc = C.lib().cae_conv_igemm(ctypes.byref(d), None)
    assert rc != 0 and b'bad kind' in C.lib().cae_last_error()
