"""Throughput of the tile loops (compress_image -> decompress_image) on a synthetic
tissue-like slide: configs[2]/[3] of BASELINE.json scaled by --size.  Reports MP/s of each
direction, bpp and PSNR (vs. the source), and where the time goes (GPU vs. host entropy coder).

    python tools/slidebench.py --size 8192 [--arch A] [--batch-tiles 32]
    torchrun --nproc-per-node 2 tools/slidebench.py --size 16384      # sharded by chunk range
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import cae_oracle as O  # noqa: E402  (synthetic tiles + PSNR definition only)
from cnn_autoencoder_b200 import compress, decompress, _store  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=8192)
    ap.add_argument('--ps', type=int, default=512)
    ap.add_argument('--arch', default='A')
    ap.add_argument('--batch-tiles', type=int, default=32)
    ap.add_argument('--workers', type=int, default=None)
    ap.add_argument('--coder-tiles', type=int, default=1024)
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    n = -(-args.size // args.ps)
    # a few distinct tiles repeated over the grid keep generation cheap; content is per (i%4, j%4)
    bank = {(i, j): O.synth_tissue_tile(i, j, ps=args.ps, seed=3) for i in range(4) for j in range(4)}
    slide = np.empty((args.size, args.size, 3), dtype=np.uint8)
    for i in range(n):
        for j in range(n):
            t = bank[(i % 4, j % 4)]
            y0, x0 = i * args.ps, j * args.ps
            h, w = min(args.ps, args.size - y0), min(args.ps, args.size - x0)
            slide[y0:y0 + h, x0:x0 + w] = t[:h, :w]
    chk = O.make_checkpoint(O.NAMED_ARCHS[args.arch], seed=1234)
    base = os.environ.get('SLIDEBENCH_DIR') or tempfile.mkdtemp(prefix='slide_')
    comp, rec = os.path.join(base, 'c.zarr'), os.path.join(base, 'r.zarr')
    kw = dict(rank=rank, world_size=world, batch_tiles=args.batch_tiles, workers=args.workers,
              coder_tiles=args.coder_tiles)
    compress.compress_image('CAE', chk, slide[:args.ps * 2, :args.ps * 2], os.path.join(base, 'w.zarr'),
                            patch_size=args.ps, rank=0, world_size=1, batch_tiles=4)   # warm-up
    t0 = time.perf_counter()
    cs = compress.compress_image('CAE', chk, slide, comp, patch_size=args.ps, **kw)
    t1 = time.perf_counter()
    ds = decompress.decompress_image(comp, rec, checkpoint=chk, **kw)
    t2 = time.perf_counter()
    out = dict(rank=rank, world=world, size=args.size, tiles=cs['tiles'],
               compress_MPps=round(cs['pixels'] / 1e6 / (t1 - t0), 1),
               decompress_MPps=round(ds['pixels'] / 1e6 / (t2 - t1), 1),
               both_MPps=round(cs['pixels'] / 1e6 / (t2 - t0), 1),
               device_coded=cs.get('device_coded'), device_decoded=ds.get('device_decoded'),
               compress_phases={k: round(v, 3) for k, v in cs.items() if k.startswith('t_')},
               decompress_phases={k: round(v, 3) for k, v in ds.items() if k.startswith('t_')},
               bytes=cs['bytes'], bpp=round(8 * cs['bytes'] / max(cs['pixels'], 1), 4),
               host_cores=os.cpu_count())
    if world == 1:
        arr = _store.DirArray(os.path.join(rec, 'decompressed/0/0'), mode='r')
        se, cnt = 0.0, 0
        for i in range(0, n, max(1, n // 4)):
            for j in range(0, n, max(1, n // 4)):
                sl = arr.chunk_slices((i, j, 0))
                d = slide[sl[0], sl[1]].astype(np.float64) - arr.read_chunk((i, j, 0)).astype(np.float64)
                se += float((d ** 2).sum()); cnt += d.size
        out['psnr_sampled_dB'] = round(20 * np.log10(255) - 10 * np.log10(se / cnt), 3)
    print(json.dumps(out))
    if not os.environ.get('SLIDEBENCH_DIR'):
        shutil.rmtree(base, ignore_errors=True)


if __name__ == '__main__':
    main()
