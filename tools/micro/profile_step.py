"""cProfile of one compress_image + decompress_image step of the WSI bench (host-side view)."""
import cProfile, os, pstats, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from oracle import cae_oracle as O
from cnn_autoencoder_b200 import compress as CMP, decompress as DEC, _slide

T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
rows = T // bench.GX
H, W = rows * bench.PS, bench.GX * bench.PS
slide = bench.aligned_empty(H * W * 3).reshape(H, W, 3)
for i in range(rows):
    for j in range(bench.GX):
        slide[i * 512:(i + 1) * 512, j * 512:(j + 1) * 512] = bench.slide_tile(O, i, j)
pin = _slide.pin_array(slide)
work = '/dev/shm/cae_prof'
shutil.rmtree(work, ignore_errors=True); os.makedirs(work)
kw = dict(batch_tiles=32, coder_tiles=T)
def step():
    cs = CMP.compress_image('CAE', chk, slide, work + '/s.zarr', patch_size=512, gpu=True, **kw)
    ds = DEC.decompress_image(work + '/s.zarr', work + '/r.zarr', checkpoint=chk, gpu=True, **kw)
    return cs, ds
step(); step(); step()
torch.cuda.synchronize()
t0 = time.perf_counter(); cs, ds = step(); torch.cuda.synchronize(); print('step wall', time.perf_counter() - t0, cs, ds)
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
shutil.rmtree(work, ignore_errors=True)
