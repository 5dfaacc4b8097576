"""Copy-only ceiling of the WSI bench on this box: every rank moves the bytes one step of
bench.py moves (2048 tiles of 512 x 512 x 3 up as strided tile DMA, the same down, plus the
stream bytes both ways), no kernels, H2D and D2H on separate streams.  Run under torchrun at
N = 1, 2, 4, 8; rank 0 prints one JSON line with the aggregate GB/s and the MP/s an infinitely
fast GPU would reach through these copies."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
from cnn_autoencoder_b200 import _cabi as C

rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
T, PS, GX = 2048, 512, 64
H, W = T // GX * PS, GX * PS
HUGE = '--hugepages' in sys.argv
if HUGE:
    # transparent huge pages under the page-locked slide (fewer IOMMU / DMA mappings): 2 MB aligned
    # anonymous memory, madvise(MADV_HUGEPAGE) before the first touch, then cudaHostRegister
    from cnn_autoencoder_b200 import _slide
    libc = ctypes.CDLL(None, use_errno=True)
    def huge(nbytes):
        raw = np.empty(nbytes + (2 << 20), dtype=np.uint8)
        off = (-raw.ctypes.data) % (2 << 20)
        a = raw[off:off + nbytes]
        rc = libc.madvise(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(nbytes), 14)   # MADV_HUGEPAGE
        a[:] = 3
        return a, rc
    src_np, rc1 = huge(H * W * 3)
    dst_np, rc2 = huge(H * W * 3)
    pins = [_slide.pin_array(src_np), _slide.pin_array(dst_np)]
    src, dst = torch.from_numpy(src_np).view(H, W, 3), torch.from_numpy(dst_np).view(H, W, 3)
    if rank == 0:
        thp = open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip()
        anon = [l for l in open('/proc/self/smaps_rollup') if 'AnonHuge' in l]
        print('madvise rc', rc1, rc2, 'THP', thp, anon, file=sys.stderr)
else:
    src = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory(); src.fill_(3)
    dst = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
streams_h = torch.empty(272_000_000, dtype=torch.uint8).pin_memory()
streams_d = torch.empty(272_000_000, dtype=torch.uint8, device='cuda')
dev_in = torch.empty((T, PS, PS, 3), dtype=torch.uint8, device='cuda')
dev_out = torch.empty_like(dev_in)
yx = np.array([(i, j) for i in range(T // GX) for j in range(GX)], dtype=np.int32)
L = C.lib()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
def up():
    for k0 in range(0, T, 32):
        C.check(L.cae_tiles_upload_u8(src.numpy().ctypes.data, H, W, 3, PS, yx[k0:k0 + 32].ctypes.data, 32,
                                      dev_in[k0:].data_ptr(), ctypes.c_void_p(s_in.cuda_stream)))
def down():
    for k0 in range(0, T, 32):
        C.check(L.cae_tiles_download_u8(dev_out[k0:].data_ptr(), 32, PS, 3, yx[k0:k0 + 32].ctypes.data,
                                        dst.numpy().ctypes.data, H, W, ctypes.c_void_p(s_out.cuda_stream)))
SEQ = '--sequential' in sys.argv
def step():
    if SEQ:
        # the order of compress_image -> decompress_image: tiles up + streams down, THEN streams
        # up + tiles down (the two directions of a step are not in flight together)
        up()
        with torch.cuda.stream(s_out):
            streams_h.copy_(streams_d, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
        with torch.cuda.stream(s_in):
            streams_d.copy_(streams_h, non_blocking=True)
        down()
        s_in.synchronize(); s_out.synchronize()
        return
    up(); down()
    with torch.cuda.stream(s_in):
        streams_d.copy_(streams_h, non_blocking=True)
    with torch.cuda.stream(s_out):
        streams_h.copy_(streams_d, non_blocking=True)
    s_in.synchronize(); s_out.synchronize()
def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
for _ in range(2):
    step()
barrier()
t0 = time.perf_counter()
K = 5
for _ in range(K):
    step()
barrier()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    nbytes = 2 * (T * PS * PS * 3 + streams_h.numel())
    print(json.dumps(dict(what='copy-only ceiling of the WSI step, directions ' + ('one after the other (as compress -> decompress)' if SEQ else 'concurrent'), n_gpus=world, s_per_step=round(dt.item() / K, 4),
                          aggregate_gb_s=round(world * nbytes * K / dt.item() / 1e9, 1),
                          per_gpu_gb_s_each_way=round(nbytes / 2 * K / dt.item() / 1e9, 1),
                          ceiling_mp_s=round(world * T * PS * PS * K / dt.item() / 1e6, 1),
                          cpus=os.cpu_count())))
if world > 1:
    dist.destroy_process_group()
