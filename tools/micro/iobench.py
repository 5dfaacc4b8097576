"""Host-side ceilings of the tile loops on this box: strided tile DMA vs contiguous copies,
chunk-file writes / reads on tmpfs (native threads)."""
import ctypes, os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cnn_autoencoder_b200 import _cabi as C, _slide
from cnn_autoencoder_b200._store import native_write, native_read

T, PS, GX = 2048, 512, 64
H, W = T // GX * PS, GX * PS
pin = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
pin.numpy()[:] = 7
slide = pin.numpy()
dev = torch.empty((T, PS, PS, 3), dtype=torch.uint8, device='cuda')
yx = np.array([(i, j) for i in range(T // GX) for j in range(GX)], dtype=np.int32)
L = C.lib()
st = torch.cuda.current_stream()
sp = ctypes.c_void_p(st.cuda_stream)
def timeit(name, fn, nbytes, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f'{name:44s} {dt*1e3:8.1f} ms  {nbytes/dt/1e9:7.1f} GB/s', flush=True)
nb = T * PS * PS * 3
def up():
    for k0 in range(0, T, 32):
        C.check(L.cae_tiles_upload_u8(slide.ctypes.data, H, W, 3, PS, yx[k0:k0+32].ctypes.data, 32, dev[k0:].data_ptr(), sp))
timeit('tile upload, 2-D DMA per tile (1536 B rows)', up, nb)
flat_d = dev.view(-1)
flat_h = pin.view(-1)
timeit('contiguous H2D, one copy', lambda: flat_d.copy_(flat_h, non_blocking=True), nb)
def up_rows():
    # one contiguous copy per tile ROW of the slide (512 x W x 3 bytes)
    for i in range(T // GX):
        flat_d[i * PS * W * 3:(i + 1) * PS * W * 3].copy_(flat_h[i * PS * W * 3:(i + 1) * PS * W * 3], non_blocking=True)
timeit('contiguous H2D, one copy per tile row', up_rows, nb)
def down():
    for k0 in range(0, T, 32):
        C.check(L.cae_tiles_download_u8(dev[k0:].data_ptr(), 32, PS, 3, yx[k0:k0+32].ctypes.data, slide.ctypes.data, H, W, sp))
timeit('tile download, 2-D DMA per tile', down, nb)
timeit('contiguous D2H, one copy', lambda: flat_h.copy_(flat_d, non_blocking=True), nb)
s2 = torch.cuda.Stream()
def both():
    flat_d.copy_(flat_h, non_blocking=True)
    with torch.cuda.stream(s2):
        out_h.copy_(dev2, non_blocking=True)
    s2.synchronize()
dev2 = torch.empty_like(flat_d); out_h = torch.empty_like(flat_h).pin_memory()
timeit('contiguous H2D + D2H concurrently', both, 2 * nb)
# files
work = '/dev/shm/cae_iobench'
shutil.rmtree(work, ignore_errors=True); os.makedirs(work)
paths = [f'{work}/{i}.{j}.0' for i, j in yx]
small = np.full(T * 132000, 3, dtype=np.uint8)
off_s = np.arange(T + 1, dtype=np.int64) * 132000
hdr = np.zeros((T, 16), dtype=np.uint8)
for th in (4, 8, 16):
    timeit(f'write {T} stream files of 132 KB, {th} threads', lambda: native_write(paths, hdr, small, off_s, th), small.size, reps=2)
timeit(f'read  {T} stream files, 16 threads', lambda: native_read(paths, 16, 16), small.size, reps=2)
big = pin.numpy().reshape(-1)
off_b = np.arange(T + 1, dtype=np.int64) * (PS * PS * 3)
paths2 = [f'{work}/r{i}.{j}.0' for i, j in yx]
for th in (4, 8, 16):
    timeit(f'write {T} raw chunk files of 786 KB, {th} threads', lambda: native_write(paths2, None, big, off_b, th), big.size, reps=2)
shutil.rmtree(work, ignore_errors=True)
