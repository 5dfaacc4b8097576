"""The batched whole-slide engine (``_slide.py``) behind compress_image / decompress_image:
config 3 in miniature -- a synthetic tissue slide with ragged edge chunks, enough chunks for the
device coder, through the public tile loops; checked chunk by chunk against the oracle codec."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ARCH = dict(channels_org=3, channels_net=32, channels_bn=16, compression_level=3,
            act_layer_type='LeakyReLU')
PS = 64


def _slide(gy, gx, crop):
    from oracle import cae_oracle as O
    s = np.concatenate([np.concatenate([O.synth_tissue_tile(i, j, ps=PS, seed=2) for j in range(gx)],
                                       axis=1) for i in range(gy)], axis=0)
    return np.ascontiguousarray(s[:s.shape[0] - crop[0], :s.shape[1] - crop[1]])


@pytest.mark.parametrize('pinned', [True, False])
def test_slide_engine_roundtrip_against_oracle(tmp_path, pinned):
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import compress, decompress, _store, _slide as S
    chk = O.make_checkpoint(ARCH, seed=31)
    oracle = O.OracleModel(chk)
    slide = _slide(12, 13, (24, 40))                # 156 chunks, edge chunks of 40 and 24 px
    handle = None
    if pinned:
        pin = torch.empty(slide.shape, dtype=torch.uint8).pin_memory()
        pin.numpy()[:] = slide
        slide = pin.numpy()
        assert S.is_pinned_array(slide)
    out = str(tmp_path / 'slide.zarr')
    st = compress.compress_image('CAE', chk, slide, out, patch_size=PS, batch_tiles=16, coder_tiles=96)
    assert st['engine'] == 'slide' and st['tiles'] == 156 and st['device_coded'] == 156
    arr = _store.DirArray(os.path.join(out, '0/0'), mode='r')
    assert arr.compressor_config['id'] == 'cae' and arr.grid == (12, 13, 1)
    lh = PS // 8
    agree = total = 0
    for (i, j) in [(0, 0), (3, 7), (11, 12), (11, 0), (0, 12), (5, 5)]:
        tile = _store.padded_tile(slide, i * PS, j * PS, PS)
        ref = oracle.codec_encode(tile)
        got = arr.read_encoded((i, j, 0))
        assert got[:16] == ref[:16]                  # '>QQ' (ps, ps): zarr pads edge chunks
        s_ref = oracle.fact_ent.decompress([ref[16:]], (lh, lh))
        s_got = oracle.fact_ent.decompress([got[16:]], (lh, lh))
        agree += int((s_ref == s_got).sum()); total += s_ref.numel()
        if torch.equal(s_ref, s_got):
            assert got == ref                        # identical symbols => identical bytes
    assert agree / total >= 0.999
    # decompress: chunk files, and straight into a caller-owned array
    rec_dir = str(tmp_path / 'rec.zarr')
    ds = decompress.decompress_image(out, rec_dir, checkpoint=chk, batch_tiles=16, coder_tiles=96)
    assert ds['engine'] == 'slide' and ds['device_decoded'] == 156
    assert ds['pixels'] == slide.shape[0] * slide.shape[1]
    rec = _store.DirArray(os.path.join(rec_dir, 'decompressed/0/0'), mode='r')
    full = np.zeros_like(slide)
    for i in range(12):
        for j in range(13):
            full[rec.chunk_slices((i, j, 0))[:2]] = rec.read_chunk((i, j, 0))
    canvas = torch.zeros(slide.shape, dtype=torch.uint8).pin_memory().numpy()
    decompress.decompress_image(out, canvas, checkpoint=chk, batch_tiles=16, coder_tiles=96)
    assert np.array_equal(canvas, full)
    ref_full = np.zeros_like(slide)
    for i in range(12):
        for j in range(13):
            tile = _store.padded_tile(slide, i * PS, j * PS, PS)
            r_tile = oracle.codec_decode(oracle.codec_encode(tile))
            sl = rec.chunk_slices((i, j, 0))
            ref_full[sl[:2]] = r_tile[:sl[0].stop - sl[0].start, :sl[1].stop - sl[1].start]
    assert abs(O.psnr_u8(slide, full) - O.psnr_u8(slide, ref_full)) <= 0.05
    # the raw chunk files of edge chunks are zero beyond the image, like zarr's fill value
    edge = np.frombuffer(rec.read_encoded((11, 12, 0)), dtype=np.uint8).reshape(PS, PS, 3)
    assert not edge[40:].any() and not edge[:, 24:].any()


def test_slide_engine_matches_the_general_tile_loop(tmp_path, monkeypatch):
    """Same chunk files from the batched engine and from the per-batch general path."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import compress, _store
    chk = O.make_checkpoint(ARCH, seed=5)
    slide = _slide(12, 12, (0, 0))
    a, b = str(tmp_path / 'a.zarr'), str(tmp_path / 'b.zarr')
    sa = compress.compress_image('CAE', chk, slide, a, patch_size=PS, batch_tiles=16)
    monkeypatch.setenv('CAE_NO_SLIDE_ENGINE', '1')
    sb = compress.compress_image('CAE', chk, slide, b, patch_size=PS, batch_tiles=16)
    assert sa.get('engine') == 'slide' and sb.get('engine') is None
    A, B = (_store.DirArray(os.path.join(p, '0/0'), mode='r') for p in (a, b))
    for i in range(12):
        for j in range(12):
            assert A.read_encoded((i, j, 0)) == B.read_encoded((i, j, 0))


def test_coder_group_schedule_writes_the_same_files(tmp_path):
    """A tapered group schedule (small last group in compress, small first group in decompress:
    ``_slide.group_sizes``) only moves the coder calls: same chunk files, same reconstruction."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import compress, decompress, _store
    chk = O.make_checkpoint(ARCH, seed=5)
    slide = _slide(12, 13, (24, 40))
    a, b = str(tmp_path / 'a.zarr'), str(tmp_path / 'b.zarr')
    sa = compress.compress_image('CAE', chk, slide, a, patch_size=PS, batch_tiles=16, coder_tiles=156)
    sb = compress.compress_image('CAE', chk, slide, b, patch_size=PS, batch_tiles=16,
                                 coder_tiles=(64, 48, 16))
    assert sa['engine'] == 'slide' and sb['engine'] == 'slide' and sb['device_coded'] == 156
    A, B = (_store.DirArray(os.path.join(p, '0/0'), mode='r') for p in (a, b))
    for i in range(12):
        for j in range(13):
            assert A.read_encoded((i, j, 0)) == B.read_encoded((i, j, 0))
    ca = torch.zeros(slide.shape, dtype=torch.uint8).pin_memory().numpy()
    cb = torch.zeros(slide.shape, dtype=torch.uint8).pin_memory().numpy()
    decompress.decompress_image(a, ca, checkpoint=chk, batch_tiles=16, coder_tiles=156)
    ds = decompress.decompress_image(b, cb, checkpoint=chk, batch_tiles=16, coder_tiles=(16, 48, 64))
    assert ds['engine'] == 'slide' and ds['device_decoded'] == 156
    assert np.array_equal(ca, cb)


def test_two_slides_in_flight_give_the_same_files_and_pixels(tmp_path):
    """``jobs.SlideJobs``: a compress_image call of one slide beside the decompress_image call of
    another (two threads, two main streams, one model) -- same chunk files and the same
    reconstruction as the calls one after the other."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import compress, decompress, _store
    from cnn_autoencoder_b200.jobs import SlideJobs
    chk = O.make_checkpoint(ARCH, seed=5)
    slides = [_slide(12, 13, (24, 40)), _slide(12, 13, (24, 40))[::-1].copy()]
    kw = dict(patch_size=PS, batch_tiles=16)

    def pinned_zeros(shape):
        return torch.zeros(shape, dtype=torch.uint8).pin_memory().numpy()
    # one after the other
    ref_dirs = [str(tmp_path / ('ref%d.zarr' % k)) for k in range(2)]
    ref_rec = [pinned_zeros(s.shape) for s in slides]
    for k in range(2):
        compress.compress_image('CAE', chk, slides[k], ref_dirs[k], coder_tiles=(96, 60), **kw)
        decompress.decompress_image(ref_dirs[k], ref_rec[k], checkpoint=chk, batch_tiles=16, coder_tiles=96)
    # six steps, two in flight
    dirs = [str(tmp_path / ('s%d.zarr' % k)) for k in range(2)]
    rec = [pinned_zeros(s.shape) for s in slides]
    with SlideJobs() as jobs:
        def comp(k, prev):
            if prev is not None:
                prev.result()
            return compress.compress_image('CAE', chk, slides[k % 2], dirs[k % 2], coder_tiles=(96, 60), **kw)

        def dec(k, fc):
            fc.result()
            return decompress.decompress_image(dirs[k % 2], rec[k % 2], checkpoint=chk, batch_tiles=16,
                                               coder_tiles=96)
        fds = []
        for k in range(6):
            fc = jobs.submit('compress', comp, k, fds[k - 2] if k >= 2 else None)
            fds.append(jobs.submit('decompress', dec, k, fc))
        stats = [f.result() for f in fds]
    torch.cuda.synchronize()
    assert all(st['engine'] == 'slide' and st['device_decoded'] == 156 for st in stats)
    for k in range(2):
        assert np.array_equal(rec[k], ref_rec[k])
        A, B = (_store.DirArray(os.path.join(p, '0/0'), mode='r') for p in (ref_dirs[k], dirs[k]))
        for i in range(12):
            for j in range(13):
                assert A.read_encoded((i, j, 0)) == B.read_encoded((i, j, 0))


def test_device_roundtrip_and_phase_record():
    from oracle import cae_oracle as O
    import cnn_autoencoder_b200 as M
    from cnn_autoencoder_b200 import _slide as S
    chk = O.make_checkpoint(ARCH, seed=9)
    model = M.autoencoder_from_state_dict(chk, gpu=True, train=False)
    tc = S.TileCodec(model, PS, 3, batch=16)
    x = torch.from_numpy(np.stack([O.synth_tissue_tile(0, j, ps=PS, seed=4) for j in range(40)])).cuda()
    out = torch.empty_like(x)
    nbytes = S.device_roundtrip(tc, x, out, coder_tiles=32)
    torch.cuda.synchronize()
    # against the eager kernels run batch by batch
    ref = tc.decode_eager(tc.encode_eager(x))
    assert torch.equal(out, ref) and nbytes > 40 * 16
    rec = S.phase_times(tc, x, 32)
    assert rec['roundtrip_exact'] and rec['chunks'] == 32


def test_metrics_match_the_reference_definitions():
    """PSNR / RMSE / bpp as src/test_cae.py:57-73 defines them (with a widening cast), on the GPU."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import metrics
    g = np.random.default_rng(4)
    x = g.integers(0, 256, size=(3, 70, 50, 3), dtype=np.uint8)
    x_r = np.clip(x.astype(np.int32) + g.integers(-9, 10, size=x.shape), 0, 255).astype(np.uint8)
    assert abs(metrics.psnr(x, x_r) - O.psnr_u8(x, x_r)) < 1e-9
    d = x.astype(np.float64) - x_r.astype(np.float64)
    assert abs(metrics.rmse(torch.from_numpy(x).cuda(), torch.from_numpy(x_r).cuda()) - np.sqrt((d ** 2).mean())) < 1e-9
    per = metrics.sse_u8(x, x_r, per_image=True).cpu().numpy()
    assert per.tolist() == [(d[i] ** 2).sum() for i in range(3)]
    assert metrics.psnr(x, x) == float('inf')
    assert metrics.bpp(1000, 100, 80) == O.bpp(1000, 100, 80)


def test_sse_kernel_gives_the_psnr_numerator():
    import ctypes
    from cnn_autoencoder_b200 import _cabi as C
    g = torch.Generator().manual_seed(3)
    a = torch.randint(0, 256, (5, 3 * 100 * 77), dtype=torch.uint8, generator=g)
    b = torch.randint(0, 256, (5, 3 * 100 * 77), dtype=torch.uint8, generator=g)
    want = ((a.double() - b.double()) ** 2).sum(1)
    sse = torch.zeros(5, dtype=torch.int64, device='cuda')
    ad, bd = a.cuda(), b.cuda()
    C.check(C.lib().cae_sse_u8(ad.data_ptr(), bd.data_ptr(), 5, a.shape[1], sse.data_ptr(),
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(sse.cpu().double(), want)
    # unaligned views take the scalar path
    sse.zero_()
    C.check(C.lib().cae_sse_u8(ad[0, 1:].data_ptr(), bd[0, 1:].data_ptr(), 1, a.shape[1] - 1,
                               sse.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert sse[0].item() == int(((a[0, 1:].double() - b[0, 1:].double()) ** 2).sum().item())


def test_save_as_bottleneck_tile_loop_with_ragged_edges(tmp_path):
    """'-sbn' (compress.py:50-62, 103-113 of the reference; 'cae_bn' decode R:653-673): the stored
    array is the latent, edge chunks zero-padded to the full latent chunk as zarr does before the
    codec sees them, header = full chunk.  Checked chunk by chunk against the oracle."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import compress, decompress, _store
    chk = O.make_checkpoint(ARCH, seed=17)
    oracle = O.OracleModel(chk)
    ps = 128
    slide = np.concatenate([np.concatenate([O.synth_tissue_tile(i, j, ps=ps, seed=6) for j in range(3)],
                                           axis=1) for i in range(2)], axis=0)
    slide = np.ascontiguousarray(slide[:200, :304])          # edge chunks of 72 and 48 px (multiples of 8)
    out = str(tmp_path / 'lat.zarr')
    st = compress.compress_image('CAE', chk, slide, out, patch_size=ps, save_as_bottleneck=True, batch_tiles=4)
    assert st['tiles'] == 6
    arr = _store.DirArray(os.path.join(out, '0/0'), mode='r')
    assert arr.compressor_config['id'] == 'cae_bn'
    assert arr.shape == (16 + 9, 16 + 16 + 6, 16) and arr.chunks == (16, 16, 16) and arr.dtype == np.float32
    agree = total = 0
    for i in range(2):
        for j in range(3):
            tile = slide[i * ps:(i + 1) * ps, j * ps:(j + 1) * ps]
            y = oracle.encoder(O.to_float_chw(tile))                  # true-size edge tile (map_blocks)
            full = torch.zeros(1, 16, 16, 16)
            full[:, :, :y.shape[2], :y.shape[3]] = y                  # zarr's chunk padding, fill 0
            ref = oracle.fact_ent.compress(full)[0]
            got = arr.read_encoded((i, j, 0))
            assert got[:16] == (16).to_bytes(8, 'big') * 2
            s_ref = oracle.fact_ent.decompress([ref], (16, 16))
            s_got = oracle.fact_ent.decompress([got[16:]], (16, 16))
            agree += int((s_ref == s_got).sum()); total += s_ref.numel()
    assert agree / total >= 0.999
    rec_dir = str(tmp_path / 'rec.zarr')
    ds = decompress.decompress_image(out, rec_dir, checkpoint=chk, batch_tiles=4)
    assert ds['pixels'] == 200 * 304
    rec = _store.DirArray(os.path.join(rec_dir, 'decompressed/0/0'), mode='r')
    assert rec.shape == (200, 304, 3)
    full_img, ref_img = np.zeros_like(slide), np.zeros_like(slide)
    for i in range(2):
        for j in range(3):
            sl = rec.chunk_slices((i, j, 0))
            full_img[sl[:2]] = rec.read_chunk((i, j, 0))
            tile = slide[sl[:2]]
            y = oracle.encoder(O.to_float_chw(tile))
            y_q, _ = oracle.fact_ent(y)
            x_r, _ = oracle.decoder(y_q)
            ref_img[sl[:2]] = O.to_uint8_hwc(x_r[0][0])
    assert abs(O.psnr_u8(slide, full_img) - O.psnr_u8(slide, ref_img)) <= 0.05


def test_save_as_bottleneck_refuses_unsupported_edges_up_front(tmp_path):
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import compress
    chk = O.make_checkpoint(ARCH, seed=17)
    slide = np.zeros((200, 300, 3), dtype=np.uint8)            # 300 - 256 = 44: not a multiple of 8
    with pytest.raises(ValueError, match='edge tiles'):
        compress.compress_image('CAE', chk, slide, str(tmp_path / 'x.zarr'), patch_size=128,
                                save_as_bottleneck=True)
    assert not os.path.exists(str(tmp_path / 'x.zarr' / '0' / '0' / '0.0.0'))


@pytest.mark.gpu
def test_banded_tile_copies_equal_the_per_tile_copies():
    """cae_tiles_upload_u8_banded / cae_tiles_download_u8_banded (one 2-D copy per run of tiles in
    a tile row + device re-tiling) against the per-tile strided copies, on a ragged slide with
    runs of different lengths, an isolated tile and edge tiles."""
    import ctypes
    import numpy as np
    import torch
    from cnn_autoencoder_b200 import _cabi as C
    rng = np.random.default_rng(3)
    H, W, c, ps = 5 * 64 + 20, 7 * 64 + 33, 3, 64
    img = torch.from_numpy(rng.integers(0, 256, size=(H, W, c), dtype=np.uint8)).pin_memory()
    tiles = [(0, 0), (0, 1), (0, 2), (1, 4), (2, 1), (2, 2), (2, 3), (2, 4), (2, 5), (2, 6), (2, 7),
             (5, 0), (5, 1), (3, 7), (4, 2), (4, 3)]
    yx = np.ascontiguousarray(np.array(tiles, dtype=np.int32))
    n = len(tiles)
    L = C.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    a = torch.full((n, ps, ps, c), 7, dtype=torch.uint8, device='cuda')
    b = torch.full((n, ps, ps, c), 9, dtype=torch.uint8, device='cuda')
    scratch = torch.empty(n * ps * ps * c, dtype=torch.uint8, device='cuda')
    C.check(L.cae_tiles_upload_u8(img.data_ptr(), H, W, c, ps, yx.ctypes.data, n, a.data_ptr(), st))
    C.check(L.cae_tiles_upload_u8_banded(img.data_ptr(), H, W, c, ps, yx.ctypes.data, n, b.data_ptr(),
                                         scratch.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    out1 = torch.zeros((H, W, c), dtype=torch.uint8).pin_memory()
    out2 = torch.zeros((H, W, c), dtype=torch.uint8).pin_memory()
    C.check(L.cae_tiles_download_u8(a.data_ptr(), n, ps, c, yx.ctypes.data, out1.data_ptr(), H, W, st))
    C.check(L.cae_tiles_download_u8_banded(a.data_ptr(), n, ps, c, yx.ctypes.data, out2.data_ptr(), H, W,
                                           scratch.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(out1, out2)
    for i, j in tiles:
        y0, x0 = i * ps, j * ps
        assert torch.equal(out2[y0:y0 + ps, x0:x0 + ps], img[y0:y0 + ps, x0:x0 + ps])


@pytest.mark.gpu
def test_test_image_scores_like_the_oracle_metrics(tmp_path):
    """cnn_autoencoder_b200.test_cae.test_image (src/test_cae.py:91-163): compress -> decompress ->
    the six metrics on the device, against the float64 restatements on the host arrays."""
    import numpy as np
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import test_cae as T
    from cnn_autoencoder_b200.decompress import decompress_image
    chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=12)
    img = O.synth_natural(1, 3, 330, 420, seed=4)[0].permute(1, 2, 0).contiguous().numpy()
    out = T.test_image(chk, img, patch_size=128, temp_output_filename=str(tmp_path / 'temp.zarr'))
    rec = np.zeros_like(img)
    decompress_image(str(tmp_path / 'temp.zarr'), rec, checkpoint=chk, gpu=True)
    assert set(T.metric_fun) <= set(out) and 'execution_time' in out and 'evaluation_time' in out
    assert abs(out['psnr'] - O.psnr_u8(img, rec)) < 1e-6
    assert abs(out['dist'] - float(np.sqrt(((img.astype(np.float64) - rec) ** 2).mean()))) < 1e-6
    assert abs(out['ssim'] - O.ssim_u8(img, rec)) < 1e-9
    assert abs(out['ms-ssim'] - O.ms_ssim_u8(img, rec)) < 2e-5
    assert abs(out['delta_cielab'] - O.delta_cielab_u8(img, rec)) <= 2e-4 * max(1.0, out['delta_cielab'])
    assert out['rate'] > 0
