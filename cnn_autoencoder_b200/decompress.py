"""Whole-slide decompression tile loop, mirroring ``src/decompress.py`` of the
reference (``decompress_fn_impl`` :24-37, ``decompress_image`` :40-140,
``decompress`` :143-180).

Native threads read up to ``coder_tiles`` chunk files into one pinned buffer, all their
rANS streams are decoded concurrently on the device, and the synthesis transform with its
fused ``*255 -> clip -> uint8 -> HWC`` epilogue runs over ``batch_tiles`` tiles at a time
while native threads write the previous batch.  (Few or irregular chunks take the general
path: host C++ coder on a thread pool.)  As in
``compress.py`` the chunk grid is sharded across processes by contiguous range;
no collective.  The reconstruction goes to ``<output>/<decomp_group>/<group>/0``
like the reference (:81-96); chunks are stored raw (the reference's Blosc-zlib
compressor comes from ``numcodecs``, which is optional here).
"""
import argparse
import os
import struct
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _autoencoders as AE
from ._entropy import decode_symbols
from ._store import DirArray, native_read, native_write
from . import _slide
from .compress import _dist_info, default_workers, load_model, shard_range


def decompress_image(input_filename, output_filename, destination_format='zarr',
                     data_group='0/0', decomp_group='decompressed', checkpoint=None,
                     progress_bar=False, gpu=False, *, rank=None, world_size=None,
                     batch_tiles=16, workers=None, coder_tiles=None):
    """Same positional signature as the reference (``decompress.py:40-47``).  Returns a
    dict of counters (tiles, pixels, seconds, device_decoded)."""
    if not torch.cuda.is_available():
        raise RuntimeError('decompress_image needs a CUDA device (no CPU fallback)')
    if checkpoint is None or (isinstance(checkpoint, str) and not len(checkpoint)):
        raise ValueError('a checkpoint is required to run the synthesis transform')
    rank, world_size = _dist_info(rank, world_size)
    workers = workers or default_workers()
    src = DirArray(os.path.join(input_filename, data_group) if data_group else input_filename,
                   mode='r')
    cfg = src.compressor_config or {}
    codec_id = cfg.get('id')
    if codec_id not in ('cae', 'cae_bn'):
        raise ValueError('array compressor %r is not a CAE codec' % codec_id)

    model = load_model(checkpoint)
    fact_ent = model['fact_ent'].module
    decoder = model['decoder']
    level = decoder.module.rec_level
    cdf, sizes, offs = fact_ent._host_tables()
    med = fact_ent._medians().reshape(1, -1, 1, 1)
    C_bn = fact_ent.channels
    c_img = decoder.module.synthesis_track[-1].model[-1].out_channels \
        if hasattr(decoder.module.synthesis_track[-1], 'model') else 3

    gy, gx = src.grid[0], src.grid[1]
    if codec_id == 'cae':
        H, W = src.shape[0], src.shape[1]
        ps = src.chunks[0]
    else:
        ps = src.chunks[0] * 2 ** level
        H, W = src.shape[0] * 2 ** level, src.shape[1] * 2 ** level

    component = '%s/%s' % (decomp_group, data_group) if len(decomp_group) else data_group
    comp_r = '/'.join(component.split('/')[:-1]) + '/0'
    in_memory = isinstance(output_filename, np.ndarray)
    out_path = None if in_memory else os.path.join(output_filename, comp_r)
    want_png = in_memory or 'zarr' not in destination_format
    if in_memory:
        # extension: reconstruct straight into a caller-owned H x W x c uint8 array (page-locked
        # memory makes the device -> host tile copies asynchronous, see _slide.pin_array)
        if output_filename.shape != (H, W, c_img) or output_filename.dtype != np.uint8:
            raise ValueError('destination array must be uint8 of shape %r' % ((H, W, c_img),))
        canvas = output_filename
    elif want_png:
        canvas = np.zeros((H, W, c_img), dtype=np.uint8)
    elif rank == 0:
        dst = DirArray(out_path, shape=(H, W, c_img), chunks=(ps, ps, c_img), dtype=np.uint8,
                       compressor=None, mode='w')
        if world_size == 1:
            dst.remove_chunks(workers)       # overwrite=True (decompress.py:92-96), see compress_image
    else:
        while not os.path.exists(os.path.join(out_path, '.zarray')):
            time.sleep(0.05)
        dst = DirArray(out_path, mode='r')

    tiles = [(i, j) for i in range(gy) for j in range(gx)]
    mine = [tiles[k] for k in shard_range(len(tiles), rank, world_size)]
    if coder_tiles is None:
        coder_tiles = _slide.default_schedule(len(mine), batch_tiles, decode=True)
    stats = dict(tiles=len(mine), pixels=0, seconds=0.0, device_decoded=0,
                 t_read=0.0, t_decode=0.0, t_gpu=0.0, t_write_wait=0.0)
    t_start = time.perf_counter()
    out_image = output_filename if isinstance(output_filename, np.ndarray) else None
    if (codec_id == 'cae' and ps % (2 ** level) == 0 and (not want_png or out_image is not None)
            and len(mine) >= fact_ent.GPU_CODER_MIN_STREAMS
            and not os.environ.get('CAE_NO_SLIDE_ENGINE')):
        tc = _slide.tile_codec(model, ps, c_img, batch_tiles)
        if len(mine) >= batch_tiles:
            tc.warm(encode=False, decode=True)
        t_start = time.perf_counter()
        st = _slide._Stats(want_trace=bool(os.environ.get('CAE_SLIDE_TRACE')))
        _slide.decompress_tiles(tc, mine, src.chunk_file, workers, coder_tiles, st, H, W,
                                out_chunk_path=None if out_image is not None else dst.chunk_file,
                                out_image=out_image)
        stats.update(st)
        stats['engine'] = 'slide'
        stats['seconds'] = time.perf_counter() - t_start
        return stats
    if not isinstance(coder_tiles, (int, np.integer)):
        coder_tiles = max(int(g) for g in coder_tiles)     # a group schedule (_slide.group_sizes)
    pool = ThreadPoolExecutor(max_workers=workers)

    def read_tile(idx):
        buf = src.read_encoded((idx[0], idx[1], 0))
        h, w = struct.unpack('>QQ', buf[:16])
        lh, lw = (h // 2 ** level, w // 2 ** level) if codec_id == 'cae' else (h, w)
        if codec_id == 'cae_bn':
            # zarr hands the caller only the part of an edge chunk that lies inside the array
            # (decompress.py:51-58): the synthesis transform runs on the cropped latent
            vh = min(lh, src.shape[0] - idx[0] * src.chunks[0])
            vw = min(lw, src.shape[1] - idx[1] * src.chunks[1])
        else:
            vh, vw = lh, lw
        return idx, (lh, lw, vh, vw), buf[16:]

    def decode_tile(item):
        idx, (lh, lw, vh, vw), stream = item
        return decode_symbols(stream, C_bn, lh * lw, cdf, sizes, offs).reshape(C_bn, lh, lw)

    writes = []

    def put_tile(idx, tile):
        y0, x0 = idx[0] * ps, idx[1] * ps
        tile = tile[:min(ps, H - y0), :min(ps, W - x0)]
        if want_png:
            canvas[y0:y0 + tile.shape[0], x0:x0 + tile.shape[1]] = tile
        else:
            dst.write_chunk((idx[0], idx[1], 0), tile)
        return tile.shape[0] * tile.shape[1]

    def run_group(batch):
        """batch: up to ``coder_tiles`` items (idx, (lh, lw), stream bytes) of one latent shape.
        All streams are entropy-decoded in one device call when there are enough of them (a
        small remainder goes to the host coder threads), then the synthesis transform runs
        over ``batch_tiles`` tiles at a time and the chunk writes go to the thread pool."""
        lh, lw, vh, vw = batch[0][1]
        t0 = time.perf_counter()
        if len(batch) >= fact_ent.GPU_CODER_MIN_STREAMS:
            sym = fact_ent.decode_streams_gpu([b[2] for b in batch], lh * lw)
            stats['device_decoded'] += len(batch)
        else:
            sym = torch.from_numpy(np.stack(list(pool.map(decode_tile, batch)))).pin_memory()
            sym = sym.cuda(non_blocking=True)
        sym = sym.reshape(len(batch), C_bn, lh, lw)[:, :, :vh, :vw]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        stats['t_decode'] += t1 - t0
        prev = None
        for k0 in range(0, len(batch) + batch_tiles, batch_tiles):
            cur = None
            if k0 < len(batch):
                y_q = sym[k0:k0 + batch_tiles].float() + med
                _, _, u8 = decoder(y_q, as_uint8='only')
                host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
                host.copy_(u8, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
                cur = (k0, host, done)
            if prev is not None:                 # hand batch k-1 to the writers while k runs
                p0, phost, pdone = prev
                pdone.synchronize()
                img = phost.numpy()
                for k in range(img.shape[0]):
                    writes.append(pool.submit(put_tile, batch[p0 + k][0], img[k]))
            prev = cur
        stats['t_gpu'] += time.perf_counter() - t1

    pins = {}

    def out_pin(shape, slot):
        # two alternating page-locked output buffers per shape (allocating them per batch would
        # stall the other CUDA calls of the process)
        key = (shape, slot)
        if key not in pins:
            pins[key] = torch.empty(shape, dtype=torch.uint8, pin_memory=True)
        return pins[key]

    read_pin = [None]

    def read_alloc(nbytes):
        if read_pin[0] is None or read_pin[0].numel() < nbytes:
            read_pin[0] = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
        return read_pin[0][:nbytes].numpy()

    def run_native(part):
        """The fast path for one read batch: chunk files -> pinned buffer (native threads) ->
        device -> all streams decoded in one call -> synthesis transform per ``batch_tiles`` ->
        pinned buffer -> chunk files (native threads).  Returns False when the batch does not
        qualify (mixed tile shapes, too few streams): the caller then takes the general path."""
        n = len(part)
        if n < fact_ent.GPU_CODER_MIN_STREAMS or want_png:
            return False
        t0 = time.perf_counter()
        paths = [src.chunk_file((i, j, 0)) for i, j in part]
        hdr, payload, off = native_read(paths, 16, workers, alloc=read_alloc)
        hw = hdr.copy().view('>u8').reshape(n, 2).astype(np.int64)
        if (hw != hw[0]).any() or (off % 4).any():
            return False
        h, w = int(hw[0, 0]), int(hw[0, 1])
        lh, lw = (h // 2 ** level, w // 2 ** level) if codec_id == 'cae' else (h, w)
        stats['t_read'] += time.perf_counter() - t0
        t1 = time.perf_counter()
        words = torch.from_numpy(payload).view(torch.int32).cuda(non_blocking=True)
        sym = fact_ent.decode_streams_device(words, off // 4, lh * lw).reshape(n, C_bn, lh, lw)
        stats['device_decoded'] += n
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        stats['t_decode'] += t2 - t1
        tile_bytes = ps * ps * c_img
        prev = None
        for k0 in range(0, n + batch_tiles, batch_tiles):
            cur = None
            if k0 < n:
                y_q = sym[k0:k0 + batch_tiles].float() + med
                _, _, u8 = decoder(y_q, as_uint8='only')
                host = out_pin(tuple(u8.shape), (k0 // batch_tiles) & 1)
                host.copy_(u8, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
                cur = (k0, host, done)
            if prev is not None:                 # write batch k-1 while batch k runs on the GPU
                p0, phost, pdone = prev
                pdone.synchronize()
                img = phost.numpy()
                m = img.shape[0]
                if img.shape[1] != ps or img.shape[2] != ps:
                    return False                 # (cannot happen for the 'cae' codec: full chunks)
                for k in range(m):               # zero what lies beyond the image (edge chunks)
                    i, j = part[p0 + k]
                    vh, vw = min(ps, H - i * ps), min(ps, W - j * ps)
                    if vh < ps:
                        img[k, vh:] = 0
                    if vw < ps:
                        img[k, :, vw:] = 0
                    stats['pixels'] += vh * vw
                native_write([dst.chunk_file((i, j, 0)) for i, j in part[p0:p0 + m]], None,
                             img.reshape(-1), np.arange(m + 1, dtype=np.int64) * tile_bytes, workers)
            prev = cur
        stats['t_gpu'] += time.perf_counter() - t2
        return True

    groups = {}
    for r0 in range(0, len(mine), coder_tiles):
        part = mine[r0:r0 + coder_tiles]
        if codec_id == 'cae' and run_native(part):
            continue
        t0 = time.perf_counter()
        items = list(pool.map(read_tile, part))
        stats['t_read'] += time.perf_counter() - t0
        for item in items:
            g = groups.setdefault(item[1], [])
            g.append(item)
            if len(g) == coder_tiles:
                run_group(g)
                groups[item[1]] = []
    for g in groups.values():
        if g:
            run_group(g)
    t0 = time.perf_counter()
    for f in writes:
        stats['pixels'] += f.result()
    stats['t_write_wait'] = time.perf_counter() - t0
    pool.shutdown()
    torch.cuda.synchronize()
    if want_png and not in_memory:
        from PIL import Image
        fn_out = output_filename.split(destination_format)[0] + destination_format
        Image.fromarray(canvas if c_img != 1 else canvas[..., 0]).save(fn_out)
    stats['seconds'] = time.perf_counter() - t_start
    return stats


def decompress(args):
    """CLI driver (``decompress.py:143-180``)."""
    inputs = args.data_dir if isinstance(args.data_dir, (list, tuple)) else [args.data_dir]
    fmt = args.destination_format if args.destination_format.startswith('.') \
        else '.' + args.destination_format
    for fn in inputs:
        if fmt.lower() in args.output_dir.lower():
            out = args.output_dir
        else:
            base = os.path.basename(fn.rstrip('/')).split('.zarr')[0]
            out = os.path.join(args.output_dir, base + fmt)
        st = decompress_image(input_filename=fn, output_filename=out, destination_format=fmt,
                              data_group=args.data_group,
                              decomp_group=args.task_label_identifier or 'decompressed',
                              checkpoint=args.checkpoint, gpu=True, batch_tiles=args.batch_tiles)
        print('Decompressed %s into %s: %d tiles, %.1f MP/s' % (
            fn, out, st['tiles'], st['pixels'] / 1e6 / max(st['seconds'], 1e-9)))


def _parser():
    ap = argparse.ArgumentParser(description='B200 CAE whole-slide decompression')
    ap.add_argument('-chk', '--checkpoint', required=True)
    ap.add_argument('-dd', '--data-dir', nargs='+', required=True)
    ap.add_argument('-o', '--output-dir', required=True)
    ap.add_argument('-df', '--destination-format', default='zarr')
    ap.add_argument('-dg', '--data-group', default='0/0')
    ap.add_argument('-tli', '--task-label-identifier', default='decompressed')
    ap.add_argument('-g', '--gpu', action='store_true')
    ap.add_argument('--batch-tiles', type=int, default=16)
    return ap


if __name__ == '__main__':
    decompress(_parser().parse_args())
