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
    assert C.lib().cae_abi_version() == C.ABI_VERSION == 4


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors against the header itself: sizes and a few offsets printed by a C
    program compiled with the header (gcc is part of the image)."""
    import subprocess
    src = tmp_path / 'layout.c'
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "cae_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(cae_tensor), sizeof(cae_conv_desc),
         sizeof(cae_eb_tables), sizeof(cae_head_desc), sizeof(cae_quant_fuse),
         offsetof(cae_conv_desc, quant), offsetof(cae_head_desc, w_stem),
         offsetof(cae_quant_fuse, y_q_planar), sizeof(cae_act_grad_desc),
         offsetof(cae_act_grad_desc, skip), offsetof(cae_act_grad_desc, db));
  return 0;
}
''')
    exe = tmp_path / 'layout'
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include')
    subprocess.run(['gcc', '-I', inc, str(src), '-o', str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True,
                                          check=True).stdout.split()]
    want = [ctypes.sizeof(C.Tensor), ctypes.sizeof(C.ConvDesc), ctypes.sizeof(C.EbTables),
            ctypes.sizeof(C.HeadDesc), ctypes.sizeof(C.QuantFuse), C.ConvDesc.quant.offset,
            C.HeadDesc.w_stem.offset, C.QuantFuse.y_q_planar.offset, ctypes.sizeof(C.ActGradDesc),
            C.ActGradDesc.skip.offset, C.ActGradDesc.db.offset]
    assert got == want
    assert ctypes.sizeof(C.Tensor) == 24


def test_error_text_is_reported():
    d = C.ConvDesc()
    d.kind = 99
    rc = C.lib().cae_conv_igemm(ctypes.byref(d), None)
    assert rc != 0 and b'bad kind' in C.lib().cae_last_error()


def test_graft_entry_build_runs():
    """The driver's build check: __graft_entry__.build() compiles (incrementally) and imports."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    g = importlib.import_module('__graft_entry__')
    g.build()


def test_new_entry_points_refuse_bad_arguments_without_touching_the_device():
    """Argument checks of the round-2 entries run before any CUDA call: null pointers, channel
    counts the kernels do not cover and a short workspace come back as error codes with a
    message (no GPU needed)."""
    L = C.lib()
    none = C.Tensor(None, C.FMT_NONE, 0, 0, 0)
    planar = C.Tensor(1 << 20, C.FMT_F16_PLANAR, 16, 0, 0)
    assert L.cae_act_grad(None, None) != 0 and b'cae_act_grad' in L.cae_last_error()
    d = C.ActGradDesc()
    d.n, d.h, d.w, d.c = 1, 8, 8, 128
    assert L.cae_act_grad(ctypes.byref(d), None) != 0 and b'null pointer' in L.cae_last_error()
    d.g, d.dz, d.out, d.skip = planar, planar, none, planar
    assert L.cae_act_grad(ctypes.byref(d), None) != 0 and b'residual' in L.cae_last_error()
    assert L.cae_conv_wgrad_workspace_bytes() > 0
    rc = L.cae_conv_wgrad(C.CONV_S1, 1, 16, 16, 128, 128, planar, planar, 0, None, None, None, 0, None)
    assert rc != 0 and b'null pointer' in L.cae_last_error()
    rc = L.cae_conv_wgrad(C.CONV_S1, 1, 16, 16, 300, 128, planar, planar, 0, 1 << 20, None, 1 << 20,
                          L.cae_conv_wgrad_workspace_bytes(), None)
    assert rc != 0 and b'256 channels' in L.cae_last_error()
    rc = L.cae_conv_wgrad(C.CONV_S1, 1, 16, 16, 128, 128, planar, planar, 0, 1 << 20, None, 1 << 20, 16, None)
    assert rc != 0 and b'workspace' in L.cae_last_error()
    assert L.cae_pack_proj_weights(64, 3, 1 << 20, None, 1 << 20, None) != 0 and b'c_in = 128' in L.cae_last_error()
    assert L.cae_image_from_proj(None, 1, 8, 8, 3, None, 0, 0, None, None, None) != 0
    assert L.cae_ssim_u8(1 << 20, 1 << 20, 1, 4, 4, 3, 1 << 20, None) != 0 and b'7x7' in L.cae_last_error()
    assert L.cae_ssim_gauss_planes_f32(1 << 20, 1 << 20, 1, 8, 8, 255.0, 1 << 20, 1 << 20, None) != 0
    assert L.cae_tiles_download_u8_banded(None, 1, 64, 3, None, None, 64, 64, None, None) != 0
