"""Minimal zarr-v2 directory array (``.zarray`` + one file per chunk, C order,
``.`` separator): just enough for the tile loops to read and write whole-slide
arrays when the ``zarr`` package is absent (it is not installed in the build
image).  The on-disk layout is the zarr v2 one, the compressor entry is the
codec's ``get_config()`` (R:530-673 ``codec_id`` 'cae' / 'cae_bn'), so a real
zarr installation with the codecs registered opens these arrays directly.
"""
import json
import os

import numpy as np


class DirArray:
    def __init__(self, path, shape=None, chunks=None, dtype=None, compressor=None, fill_value=0,
                 mode='r'):
        self.path = path
        meta_file = os.path.join(path, '.zarray')
        if mode == 'w':
            os.makedirs(path, exist_ok=True)
            self.shape = tuple(int(s) for s in shape)
            self.chunks = tuple(int(c) for c in chunks)
            self.dtype = np.dtype(dtype)
            self.fill_value = fill_value
            self._codec = compressor
            meta = dict(zarr_format=2, shape=list(self.shape), chunks=list(self.chunks),
                        dtype=self.dtype.str, fill_value=fill_value, order='C', filters=None,
                        compressor=compressor.get_config() if compressor is not None else None,
                        dimension_separator='.')
            # written aside and renamed: the other ranks of a sharded job open the array while
            # rank 0 (re)creates it and must never see a half-written file
            tmp = meta_file + '.%d.tmp' % os.getpid()
            with open(tmp, 'w') as f:
                json.dump(meta, f, indent=1, default=str)
            os.replace(tmp, meta_file)
        else:
            with open(meta_file) as f:
                meta = json.load(f)
            self.shape = tuple(meta['shape'])
            self.chunks = tuple(meta['chunks'])
            self.dtype = np.dtype(meta['dtype'])
            self.fill_value = meta.get('fill_value', 0) or 0
            self._codec = compressor
            self.compressor_config = meta.get('compressor')

    @property
    def grid(self):
        return tuple(-(-s // c) for s, c in zip(self.shape, self.chunks))

    def chunk_file(self, idx):
        return os.path.join(self.path, '.'.join(str(i) for i in idx))

    def chunk_slices(self, idx):
        return tuple(slice(i * c, min((i + 1) * c, s))
                     for i, c, s in zip(idx, self.chunks, self.shape))

    # raw (already encoded) chunk bytes
    def write_encoded(self, idx, *parts):
        """Write one chunk file from byte-like parts (header, payload, ...) without joining them."""
        tmp = self.chunk_file(idx) + '.partial'
        with open(tmp, 'wb') as f:
            for part in parts:
                f.write(part)
        os.replace(tmp, self.chunk_file(idx))

    def read_encoded(self, idx):
        with open(self.chunk_file(idx), 'rb') as f:
            return f.read()

    # uncompressed arrays (compressor None): full chunk, edge chunks padded with fill_value
    def write_chunk(self, idx, arr):
        full = np.full(self.chunks, self.fill_value, dtype=self.dtype)
        full[tuple(slice(0, s) for s in arr.shape)] = arr
        self.write_encoded(idx, full.tobytes())

    def read_chunk(self, idx):
        raw = self.read_encoded(idx)
        full = np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks)
        sl = self.chunk_slices(idx)
        return full[tuple(slice(0, s.stop - s.start) for s in sl)]

    def remove_chunks(self, threads=1):
        """Delete every chunk file of the array (what ``overwrite=True`` of the reference's
        ``to_zarr`` calls does before writing, compress.py:123-128); metadata stays."""
        names = [os.path.join(self.path, f) for f in os.listdir(self.path) if not f.startswith('.')]
        native_remove([n for n in names if os.path.isfile(n)], threads)
        return len(names)

    def nbytes_stored(self):
        return sum(os.path.getsize(os.path.join(self.path, f)) for f in os.listdir(self.path)
                   if not f.startswith('.'))


def padded_tile(src, y0, x0, ps, fill_value=0):
    """ps x ps x C tile at (y0, x0) of an H x W x C array-like, edge tiles padded to
    the full chunk with the fill value (what zarr v2 hands the chunk codec)."""
    h, w, c = src.shape
    tile = np.asarray(src[y0:min(y0 + ps, h), x0:min(x0 + ps, w)])
    if tile.shape[0] == ps and tile.shape[1] == ps:
        return np.ascontiguousarray(tile)
    out = np.full((ps, ps, c), fill_value, dtype=tile.dtype)
    out[:tile.shape[0], :tile.shape[1]] = tile
    return out


# --------------------------------------------------------------------------
# Native batch I/O of the tile loops (csrc/host_io.cpp through the C ABI)
# --------------------------------------------------------------------------

def _paths_blob(paths):
    return b''.join(os.fsencode(p) + b'\0' for p in paths)


def native_gather(src, ps, tile_yx, dst, threads):
    """dst[k] = ps x ps x C tile tile_yx[k] of the C-contiguous uint8 H x W x C ndarray ``src``
    (edge tiles zero filled), by native threads (``cae_tiles_gather_u8``)."""
    from . import _cabi as C
    tile_yx = np.ascontiguousarray(tile_yx, dtype=np.int32)
    h, w, c = src.shape
    C.check(C.lib().cae_tiles_gather_u8(src.ctypes.data, h, w, c, ps, tile_yx.ctypes.data,
                                        tile_yx.shape[0], dst.ctypes.data, threads))


def can_native_gather(src):
    return (isinstance(src, np.ndarray) and src.dtype == np.uint8 and src.ndim == 3
            and src.flags.c_contiguous)


def native_write(paths, headers, payload, payload_off, threads):
    """File k = headers[k] (uint8 n x hdr_len, or None) + payload[payload_off[k]:payload_off[k+1]]
    (``cae_files_write``: written as .partial, then renamed)."""
    from . import _cabi as C
    off = np.ascontiguousarray(payload_off, dtype=np.int64)
    hdr_len = 0 if headers is None else headers.shape[1]
    if headers is not None:
        headers = np.ascontiguousarray(headers, dtype=np.uint8)
    C.check(C.lib().cae_files_write(_paths_blob(paths), len(paths),
                                    headers.ctypes.data if headers is not None else None, hdr_len,
                                    payload.ctypes.data, off.ctypes.data, threads))


def native_remove(paths, threads):
    """Unlink ``paths`` (missing files are skipped) by native threads (``cae_files_remove``)."""
    from . import _cabi as C
    if len(paths):
        C.check(C.lib().cae_files_remove(_paths_blob(paths), len(paths), threads))


def native_read(paths, hdr_len, threads, alloc=None, slices=1, on_slice=None):
    """Read files whole: returns (headers uint8 n x hdr_len, payload uint8 1-D, int64 offsets
    n + 1).  ``alloc(nbytes)`` provides the payload buffer (a pinned one for uploads).  With
    ``slices`` > 1 the files are read in that many runs and ``on_slice(payload, lo, hi)`` is
    called after each with the byte range that has just arrived (the caller starts its upload
    while the next run is read)."""
    from . import _cabi as C
    n = len(paths)
    enc = [os.fsencode(p) for p in paths]
    blob = b'\0'.join(enc) + b'\0' if n else b''
    sizes = np.empty(n, dtype=np.int64)
    C.check(C.lib().cae_files_stat(blob, n, sizes.ctypes.data, threads))
    if (sizes < hdr_len).any():
        bad = paths[int(np.argmax(sizes < hdr_len))]
        raise FileNotFoundError('chunk file %s is missing or truncated' % bad)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(sizes - hdr_len, out=off[1:])
    headers = np.empty((n, max(hdr_len, 1)), dtype=np.uint8)
    payload = alloc(int(off[-1])) if alloc is not None else np.empty(int(off[-1]), dtype=np.uint8)
    slices = max(1, min(int(slices), n))
    if slices == 1 or on_slice is None:
        C.check(C.lib().cae_files_read(blob, n, headers.ctypes.data, hdr_len, payload.ctypes.data,
                                       off.ctypes.data, threads))
        if on_slice is not None:
            on_slice(payload, 0, int(off[-1]))
        return headers[:, :hdr_len], payload, off
    ends = np.cumsum([len(e) + 1 for e in enc])                      # byte offsets into the blob
    for k in range(slices):
        a, b = k * n // slices, (k + 1) * n // slices
        if a == b:
            continue
        sub = blob[int(ends[a - 1]) if a else 0:int(ends[b - 1])]
        # payload_off is read relative to the payload base: pass the slice of the offset table
        C.check(C.lib().cae_files_read(sub, b - a, headers[a:].ctypes.data, hdr_len,
                                       payload.ctypes.data, off[a:].ctypes.data, threads))
        on_slice(payload, int(off[a]), int(off[b]))
    return headers[:, :hdr_len], payload, off
