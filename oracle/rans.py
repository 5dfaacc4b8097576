"""ctypes front-end of ``oracle/rans_oracle.c``.  TEST INFRASTRUCTURE ONLY.

Mirrors the three CompressAI entry points the reference reaches through
``EntropyBottleneck.update/compress/decompress``
(``/root/reference/src/models/tasks/_autoencoders.py:502,549-551,568-572``):
``pmf_to_quantized_cdf``, ``RansEncoder.encode_with_indexes`` and
``RansDecoder.decode_with_indexes``.  Parity unpinned.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle_rans.so')
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'rans_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE, '-B'])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        i32p = ctypes.POINTER(ctypes.c_int32)
        lib.oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        lib.oracle_pmf_to_quantized_cdf.argtypes = [
            ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_int,
            ctypes.POINTER(ctypes.c_uint32)]
        lib.oracle_rans_encode.restype = ctypes.c_long
        lib.oracle_rans_encode.argtypes = [
            i32p, i32p, ctypes.c_long, i32p, ctypes.c_int, i32p, i32p,
            ctypes.POINTER(ctypes.c_uint8), ctypes.c_long]
        lib.oracle_rans_decode.restype = ctypes.c_int
        lib.oracle_rans_decode.argtypes = [
            ctypes.POINTER(ctypes.c_uint8), ctypes.c_long, i32p, ctypes.c_long,
            i32p, ctypes.c_int, i32p, i32p, i32p]
        _lib = lib
    return _lib


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def pmf_to_quantized_cdf(pmf, precision=16):
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    cdf = np.zeros(pmf.shape[0] + 1, dtype=np.uint32)
    rc = _load().oracle_pmf_to_quantized_cdf(
        pmf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), pmf.shape[0], precision,
        cdf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    if rc != 0:
        raise ValueError(f'pmf_to_quantized_cdf domain error {rc}')
    return cdf


def encode_with_indexes(symbols, indexes, cdfs, cdf_sizes, offsets):
    sym, psym = _i32(symbols)
    idx, pidx = _i32(indexes)
    cdf, pcdf = _i32(cdfs)
    siz, psiz = _i32(cdf_sizes)
    off, poff = _i32(offsets)
    n = sym.shape[0]
    cap = 4 * (4 * n + 16)
    out = np.empty(cap, dtype=np.uint8)
    nb = _load().oracle_rans_encode(psym, pidx, n, pcdf, cdf.shape[1], psiz, poff,
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), cap)
    if nb < 0:
        raise RuntimeError(f'oracle_rans_encode failed ({nb})')
    return out[:nb].tobytes()


def decode_with_indexes(encoded, indexes, cdfs, cdf_sizes, offsets):
    enc = np.frombuffer(bytes(encoded), dtype=np.uint8)
    idx, pidx = _i32(indexes)
    cdf, pcdf = _i32(cdfs)
    siz, psiz = _i32(cdf_sizes)
    off, poff = _i32(offsets)
    out = np.empty(idx.shape[0], dtype=np.int32)
    rc = _load().oracle_rans_decode(
        enc.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), enc.shape[0], pidx, idx.shape[0],
        pcdf, cdf.shape[1], psiz, poff, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    if rc != 0:
        raise RuntimeError(f'oracle_rans_decode failed ({rc})')
    return out
