"""Host-side phase trace of compress_image -> decompress_image steps (CAE_SLIDE_TRACE=1):
when the symbols are ready, coded, on the host, written; when the files are read, decoded, the
last batch issued, the GPU done.

    python tools/micro/trace_slide.py [--no-trace] [chunks=4096] [schedule ...]

A schedule is ``coder[/decoder]``; each side is one group size or a comma-separated group
schedule (``_slide.group_sizes``), e.g. ``4096`` or ``4096,3072,1024/1024,3072,4096``.  Every
schedule runs 4 steps on the same slide; the last step's times and traces are printed.
"""
import os
import shutil
import sys
import time

IN_FLIGHT = 1
if '--in-flight' in sys.argv:         # two steps in flight (jobs.SlideJobs), plain timings
    sys.argv.remove('--in-flight')
    IN_FLIGHT = 2
    sys.argv.append('--no-trace') if '--no-trace' not in sys.argv else None
if '--no-trace' in sys.argv:          # plain timings: the trace marks synchronise the helper threads
    sys.argv.remove('--no-trace')
else:
    os.environ['CAE_SLIDE_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from oracle import cae_oracle as O  # noqa: E402
from cnn_autoencoder_b200 import compress as CMP, decompress as DEC, _slide  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096


def sched(a):
    v = [int(x) for x in a.split(',')]
    return v[0] if len(v) == 1 else v


schedules = []
for a in sys.argv[2:] or ['auto']:
    cd = a.split('/')
    schedules.append(tuple(None if x == 'auto' else sched(x) for x in (cd[0], cd[-1])))
chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
rows = T // bench.GX
H, W = rows * bench.PS, bench.GX * bench.PS
slide = bench.aligned_empty(H * W * 3).reshape(H, W, 3)
tile = bench.slide_tile(O, 0, 0)
for i in range(rows):
    for j in range(bench.GX):
        slide[i * 512:(i + 1) * 512, j * 512:(j + 1) * 512] = tile if (i + j) % 7 else bench.slide_tile(O, i % 3, j % 5)
pin = _slide.pin_array(slide)
recon = bench.aligned_empty(H * W * 3).reshape(H, W, 3)
recon[:] = 0
pin_r = _slide.pin_array(recon)
work = '/dev/shm/cae_trace'
shutil.rmtree(work, ignore_errors=True)
os.makedirs(work)
def pipelined(G, GD, steps=6):
    """`steps` steps with the compress call of step k + 1 beside the decompress call of step k."""
    from cnn_autoencoder_b200.jobs import SlideJobs
    stores = [work + '/p%d.zarr' % k for k in range(2)]
    # (CUDA graphs are captured on first use, and a capture does not tolerate another thread's
    # launches: one step alone first)
    CMP.compress_image('CAE', chk, slide, stores[0], patch_size=512, gpu=True, batch_tiles=32, coder_tiles=G)
    DEC.decompress_image(stores[0], recon, checkpoint=chk, gpu=True, batch_tiles=32, coder_tiles=GD)
    with SlideJobs() as jobs:
        def comp(k, prev):
            if prev is not None:
                prev.result()          # (compress_image empties the store it overwrites by itself)
            return CMP.compress_image('CAE', chk, slide, stores[k % 2], patch_size=512, gpu=True,
                                      batch_tiles=32, coder_tiles=G)

        def dec(k, fc):
            fc.result()
            return DEC.decompress_image(stores[k % 2], recon, checkpoint=chk, gpu=True, batch_tiles=32,
                                        coder_tiles=GD)
        for rep in range(2):
            for st in stores:
                shutil.rmtree(st, ignore_errors=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fds = []
            for k in range(steps):
                fc = jobs.submit('compress', comp, k, fds[k - 2] if k >= 2 else None)
                fds.append(jobs.submit('decompress', dec, k, fc))
            out = [f.result() for f in fds]
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / steps
    print('chunks %d  coder groups %s  decoder groups %s  TWO STEPS IN FLIGHT: %.4f s per step '
          '(%.0f MP/s; store removal inside)' % (T, G, GD, dt, T * 512 * 512 / dt / 1e6), flush=True)


for G, GD in schedules:
    if IN_FLIGHT == 2:
        pipelined(G, GD)
        continue
    best = None
    for it in range(4):
        shutil.rmtree(work + '/s.zarr', ignore_errors=True)      # a fresh store per step (untimed)
        t0 = time.perf_counter()
        cs = CMP.compress_image('CAE', chk, slide, work + '/s.zarr', patch_size=512, gpu=True,
                                batch_tiles=32, coder_tiles=G)
        t1 = time.perf_counter()
        ds = DEC.decompress_image(work + '/s.zarr', recon, checkpoint=chk, gpu=True, batch_tiles=32,
                                  coder_tiles=GD)
        t2 = time.perf_counter()
        if it and (best is None or t2 - t0 < sum(best)):
            best = (t1 - t0, t2 - t1)
    print('chunks %d  coder groups %s  decoder groups %s' % (T, G, GD))
    print('last: compress %.4f s  decompress %.4f s   best step: %.4f + %.4f = %.4f s' %
          (t1 - t0, t2 - t1, best[0], best[1], sum(best)))
    print('compress trace', cs.get('trace'))
    print('decompress trace', ds.get('trace'))
    print({k: v for k, v in cs.items() if k != 'trace'})
    print({k: v for k, v in ds.items() if k != 'trace'}, flush=True)
shutil.rmtree(work, ignore_errors=True)
