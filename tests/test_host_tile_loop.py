"""Host-side logic of the tile loops: chunk-range sharding (also across two gloo ranks),
the zarr-v2 directory array, edge-tile padding, and the reference-facing signatures."""
import inspect
import json
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cnn_autoencoder_b200 import _store, compress, decompress


def test_shard_ranges_partition_the_chunk_grid():
    for n in (1, 7, 64, 9604):
        for g in (1, 2, 4, 8):
            parts = [list(compress.shard_range(n, k, g)) for k in range(g)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_wsi_chunk_grid_matches_survey():
    # 50 000 x 50 000 slide, 512 px chunks -> 98 x 98 = 9 604 chunks, edge chunks 336 px
    H = W = 50000
    gy, gx = -(-H // 512), -(-W // 512)
    assert (gy, gx, gy * gx) == (98, 98, 9604)
    assert H - (gy - 1) * 512 == 336


def test_dir_array_roundtrip_and_metadata(tmp_path):
    class Cfg:
        codec_id = 'cae'

        def get_config(self):
            return dict(id='cae', checkpoint='x.pth', gpu=True)

    a = _store.DirArray(str(tmp_path / 'a'), shape=(5, 7, 3), chunks=(4, 4, 3), dtype=np.uint8,
                        compressor=None, mode='w')
    assert a.grid == (2, 2, 1)
    src = np.arange(5 * 7 * 3, dtype=np.uint8).reshape(5, 7, 3)
    for i in range(2):
        for j in range(2):
            sl = a.chunk_slices((i, j, 0))
            a.write_chunk((i, j, 0), src[sl])
    b = _store.DirArray(str(tmp_path / 'a'), mode='r')
    out = np.zeros_like(src)
    for i in range(2):
        for j in range(2):
            out[b.chunk_slices((i, j, 0))] = b.read_chunk((i, j, 0))
    assert np.array_equal(out, src)
    c = _store.DirArray(str(tmp_path / 'c'), shape=(8, 8, 3), chunks=(4, 4, 3), dtype=np.uint8,
                        compressor=Cfg(), mode='w')
    c.write_encoded((1, 0, 0), b'abc')
    meta = json.load(open(tmp_path / 'c' / '.zarray'))
    assert meta['zarr_format'] == 2 and meta['compressor']['id'] == 'cae'
    assert meta['chunks'] == [4, 4, 3] and meta['dtype'] == '|u1'
    assert c.read_encoded((1, 0, 0)) == b'abc' and c.nbytes_stored() == 3
    view = compress.open_source(str(tmp_path / 'a'), data_group='')
    assert np.array_equal(view[slice(1, 5), slice(2, 7)], src[1:5, 2:7])


def test_edge_tiles_are_padded_like_zarr_chunks():
    img = np.full((10, 6, 3), 9, dtype=np.uint8)
    t = _store.padded_tile(img, 8, 4, 4)
    assert t.shape == (4, 4, 3)
    assert (t[:2, :2] == 9).all() and (t[2:] == 0).all() and (t[:, 2:] == 0).all()
    assert _store.padded_tile(img, 0, 0, 4).flags['C_CONTIGUOUS']


def test_reference_signatures_are_kept():
    # compress.py:29-36 and decompress.py:40-47 of the reference
    ci = list(inspect.signature(compress.compress_image).parameters)
    assert ci[:11] == ['codec', 'checkpoint', 'input_filename', 'output_filename', 'patch_size',
                       'source_format', 'data_group', 'data_axes', 'progress_bar',
                       'save_as_bottleneck', 'gpu']
    di = list(inspect.signature(decompress.decompress_image).parameters)
    assert di[:8] == ['input_filename', 'output_filename', 'destination_format', 'data_group',
                      'decomp_group', 'checkpoint', 'progress_bar', 'gpu']
    with pytest.raises(ValueError):
        compress.compress_image('Blosc', None, np.zeros((4, 4, 3), np.uint8), '/tmp/x')


def _gloo_worker(rank, world, port, n_chunks, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = torch.zeros(n_chunks, dtype=torch.int32)
    for i in compress.shard_range(n_chunks, rank, world):
        mine[i] = rank + 1
    dist.all_reduce(mine)            # test-only check; the data path itself has no collective
    if rank == 0:
        torch.save(mine, out)
    dist.destroy_process_group()


def test_two_ranks_cover_every_chunk_exactly_once(tmp_path):
    out = str(tmp_path / 'cover.pt')
    mp.spawn(_gloo_worker, args=(2, 29517, 97, out), nprocs=2, join=True)
    cover = torch.load(out)
    assert (cover > 0).all()
    assert (cover[:48] == 1).all() and (cover[48:] == 2).all()


def test_model_mirror_builds_on_cpu_and_refuses_to_run():
    import cnn_autoencoder_b200 as M
    from cnn_autoencoder_b200 import _cabi
    m = M.setup_modules(channels_org=3, channels_net=16, channels_bn=8, compression_level=2,
                        act_layer_type='LeakyReLU', use_residual=True)
    assert set(m) == {'encoder', 'decoder', 'fact_ent'}
    keys = set(m['encoder'].state_dict())
    assert 'analysis_track.0.res_model.0.weight' in keys and 'analysis_track.1.model.0.weight' in keys
    m['encoder'].eval()
    with pytest.raises(_cabi.CaeError):
        m['encoder'](torch.zeros(1, 3, 16, 16))
    g = M.setup_modules(channels_org=3, channels_net=16, channels_bn=8, compression_level=2,
                        act_layer_type='GDN')
    gk = set(g['encoder'].state_dict())
    assert {'analysis_track.0.model.1.beta', 'analysis_track.0.model.1.gamma',
            'analysis_track.0.model.1.beta_reparam.pedestal',
            'analysis_track.0.model.1.gamma_reparam.lower_bound.bound'} <= gk
    with pytest.raises(ValueError):      # 16 is not a multiple of groups=3: torch refuses, as in the reference
        M.setup_modules(channels_org=3, channels_net=16, channels_bn=8, compression_level=2,
                        act_layer_type='LeakyReLU', groups=True)
    d = M.setup_modules(channels_org=4, channels_net=4, channels_bn=4, compression_level=2,
                        act_layer_type='LeakyReLU', groups=True)
    assert d['encoder'].analysis_track[0].model[0].groups == 4
    with pytest.raises(ValueError):
        M.Analyzer(act_layer_type='LeakyRelU')      # the reference's own typo default is rejected


def test_multiscale_synthesizer_train_mode_equals_oracle():
    """The colour heads of multiscale_analysis=True exist with the reference's state-dict keys
    and the train()-mode (torch) forward reproduces the oracle on CPU."""
    import torch
    import cnn_autoencoder_b200 as M
    from oracle import cae_oracle as O
    arch = dict(channels_org=3, channels_net=16, channels_bn=8, compression_level=3,
                act_layer_type='LeakyReLU', multiscale_analysis=True, bias=True)
    chk = O.make_checkpoint(arch, seed=6)
    assert 'color_layers.0.0.weight' in chk['decoder'] and 'color_layers.1.0.bias' in chk['decoder']
    model = M.autoencoder_from_state_dict(chk, gpu=False, train=True)
    y = torch.randn(1, 8, 4, 6)
    with torch.no_grad():
        x_r, _ = model['decoder'](y)
        ref, _ = O.OracleModel(chk).decoder(y)
    for a, b in zip(x_r, ref):
        assert torch.equal(a, b)


def test_native_tile_gather_and_chunk_files(tmp_path):
    """csrc/host_io.cpp through the C ABI (no GPU): tiles cut out of a slide with zero-filled
    edges equal ``padded_tile``; chunk files written / read back by native threads keep the
    header / payload split byte for byte."""
    import numpy as np
    from cnn_autoencoder_b200 import _store as S
    rng = np.random.default_rng(0)
    src = rng.integers(0, 255, (700, 900, 3), dtype=np.uint8)
    ps = 256
    yx = np.array([[0, 0], [2, 3], [1, 2], [2, 0], [0, 3]], dtype=np.int32)
    dst = np.full((len(yx), ps, ps, 3), 7, np.uint8)
    assert S.can_native_gather(src) and not S.can_native_gather(src[:, ::2])
    S.native_gather(src, ps, yx, dst, 4)
    for k, (ty, tx) in enumerate(yx):
        assert np.array_equal(dst[k], S.padded_tile(src, ty * ps, tx * ps, ps)), k
    paths = [str(tmp_path / f'c.{k}') for k in range(6)]
    payload = rng.integers(0, 255, 4000, dtype=np.uint8)
    off = np.array([0, 16, 16, 900, 2777, 3999, 4000])          # includes an empty payload
    hdr = rng.integers(0, 255, (6, 16), dtype=np.uint8)
    S.native_write(paths, hdr, payload, off, 3)
    assert not list(tmp_path.glob('*.partial'))
    h2, p2, o2 = S.native_read(paths, 16, 3)
    assert np.array_equal(h2, hdr) and np.array_equal(p2, payload) and np.array_equal(o2, off)
    with open(paths[3], 'rb') as f:
        assert f.read() == hdr[3].tobytes() + payload[900:2777].tobytes()
    import pytest
    with pytest.raises(FileNotFoundError):
        S.native_read(paths + [str(tmp_path / 'missing')], 16, 2)


def test_engine_plans_the_fused_head_where_it_applies():
    """Host-side planning only (no kernels): the first unit is handed to the fused head kernel
    for plain and residual nets reading a raw image, and not where its conditions fail."""
    import torch
    import cnn_autoencoder_b200 as M
    from cnn_autoencoder_b200 import _cabi as C, _ops
    from oracle import cae_oracle as O

    def span(arch, fmt=C.FMT_U8_HWC, keep=()):
        chk = O.make_checkpoint(dict(channels_org=3, channels_net=32, channels_bn=16, **arch), seed=1)
        model = M.autoencoder_from_state_dict(chk, gpu=False, train=False)
        ex = model['encoder'].module._executor()
        t = torch.zeros(1, 8, 8, 3, dtype=torch.uint8)
        x = _ops.Act(t, fmt, 1, 3, 8, 8)
        return ex._head_match(0, x, keep, C.FMT_F32_NCHW), len(ex.steps)

    assert span(dict(compression_level=3, act_layer_type='LeakyReLU'))[0] == 2
    assert span(dict(compression_level=3, act_layer_type='LeakyReLU', use_residual=True))[0] == 3
    assert span(dict(compression_level=3, act_layer_type='ReLU', batch_norm=True))[0] == 2
    # stride-2 consumer wants the parity-split layout: not covered by the head kernel
    assert span(dict(compression_level=2, act_layer_type='LeakyReLU'))[0] == 0
    # GDN units have no stride-1 stem; planar (already converted) inputs are not raw images
    assert span(dict(compression_level=3, act_layer_type='GDN'))[0] == 0
    assert span(dict(compression_level=3, act_layer_type='LeakyReLU'), fmt=C.FMT_F16_PLANAR)[0] == 0
    # a caller that wants the stem's output kept gets the separate kernels
    assert span(dict(compression_level=3, act_layer_type='LeakyReLU'), keep=(1,))[0] == 0


def test_source_axes_are_mapped_like_the_reference():
    """data_axes (compress.py:89-101 of the reference): a TCZYX source is read as Y x X x C with
    index 0 of the other axes; three-dimensional sources are YXC already."""
    import numpy as np
    from cnn_autoencoder_b200.compress import as_yxc
    rng = np.random.default_rng(0)
    a = rng.integers(0, 255, size=(2, 3, 2, 20, 30), dtype=np.uint8)       # T C Z Y X
    v = as_yxc(a, 'TCZYX')
    assert v.shape == (20, 30, 3)
    want = a[0, :, 0].transpose(1, 2, 0)
    assert np.array_equal(v[3:11, 5:30], want[3:11, 5:30])
    assert np.array_equal(v[0:20, 0:30, :], want)
    b = rng.integers(0, 255, size=(4, 5, 3, 1), dtype=np.uint8)            # Y X C Z
    assert np.array_equal(as_yxc(b, 'YXCZ')[1:3, 0:5], b[1:3, 0:5, :, 0])
    c = rng.integers(0, 255, size=(6, 7, 3), dtype=np.uint8)
    assert as_yxc(c, 'TCZYX') is c
    import pytest
    with pytest.raises(ValueError):
        as_yxc(a, 'YXC')


def test_coder_group_schedules():
    """``_slide.group_sizes``: one size = equal groups; a sequence = a schedule whose last entry
    repeats; sizes are whole batches; every tile is in exactly one group."""
    from cnn_autoencoder_b200._slide import group_sizes
    assert group_sizes(8192, 4096, 32) == [4096, 4096]
    assert group_sizes(8192, 8192, 32) == [8192]
    assert group_sizes(8192, 100000, 32) == [8192]
    assert group_sizes(8192, (4096, 3072, 1024), 32) == [4096, 3072, 1024]
    assert group_sizes(8192, [1024, 3072], 32) == [1024, 3072, 3072, 1024]
    assert group_sizes(156, 96, 16) == [96, 60]
    assert group_sizes(156, (64, 48, 16), 16) == [64, 48, 16, 16, 12]
    assert group_sizes(100, 10, 32) == [32, 32, 32, 4]        # never below one batch
    assert group_sizes(5, 1024, 16) == [5]
    for bad in (0, (), (16, 0), -3):
        with pytest.raises(ValueError):
            group_sizes(64, bad, 16)


def test_chunk_files_are_removed_like_overwrite_true(tmp_path):
    """``cae_files_remove`` / ``DirArray.remove_chunks``: what ``to_zarr(overwrite=True)`` of the
    reference (compress.py:123-128) does to the chunks of an array that exists already."""
    arr = _store.DirArray(str(tmp_path / 'a'), shape=(8, 8, 3), chunks=(4, 4, 3), dtype=np.uint8, mode='w')
    for i in range(2):
        for j in range(2):
            arr.write_chunk((i, j, 0), np.full((4, 4, 3), i + j, dtype=np.uint8))
    paths = [arr.chunk_file((i, j, 0)) for i in range(2) for j in range(2)]
    assert all(os.path.exists(p) for p in paths)
    _store.native_remove(paths[:2] + [str(tmp_path / 'a' / 'missing.0.0')], 3)     # missing: skipped
    assert [os.path.exists(p) for p in paths] == [False, False, True, True]
    assert arr.remove_chunks(2) == 2
    assert not any(os.path.exists(p) for p in paths)
    assert os.path.exists(os.path.join(arr.path, '.zarray'))                        # metadata stays
    _store.native_remove([], 2)


def test_native_read_in_slices_reports_byte_ranges(tmp_path):
    """``native_read(slices=k, on_slice=...)``: same result as one run; the callback sees
    consecutive byte ranges that cover the payload (the upload of a range starts while the next
    run of files is read)."""
    g = np.random.default_rng(3)
    n = 37
    sizes = g.integers(0, 900, size=n)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    payload = g.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    hdr = g.integers(0, 256, size=(n, 16), dtype=np.uint8)
    paths = [str(tmp_path / ('%d.%d.0' % (k // 6, k % 6))) for k in range(n)]
    _store.native_write(paths, hdr, payload, off, 4)
    h1, p1, o1 = _store.native_read(paths, 16, 3)
    seen = []
    h2, p2, o2 = _store.native_read(paths, 16, 3, slices=4,
                                    on_slice=lambda buf, lo, hi: seen.append((lo, hi, bytes(buf[lo:hi]))))
    assert np.array_equal(h1, hdr) and np.array_equal(p1, payload) and np.array_equal(o1, off)
    assert np.array_equal(h2, hdr) and np.array_equal(p2, payload) and np.array_equal(o2, off)
    assert len(seen) == 4 and seen[0][0] == 0 and seen[-1][1] == off[-1]
    assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    assert b''.join(s[2] for s in seen) == payload.tobytes()           # each range had arrived
    seen.clear()
    _store.native_read(paths[:2], 16, 3, slices=8, on_slice=lambda buf, lo, hi: seen.append((lo, hi)))
    assert seen[0][0] == 0 and seen[-1][1] == off[2]


def test_default_coder_schedule():
    """``_slide.default_schedule`` (what ``coder_tiles=None`` means): one group for small shards,
    two for a shard, the last group of compress_image a quarter (the measured optimum at 8192
    chunks), groups capped at 8192 tiles; every tile in exactly one group."""
    from cnn_autoencoder_b200._slide import default_schedule, group_sizes
    assert default_schedule(156, 16) == [156] and default_schedule(156, 16, decode=True) == [156]
    assert default_schedule(4096, 32) == [2048, 2048]
    assert default_schedule(8192, 32) == [6144, 2048]
    assert default_schedule(8192, 32, decode=True) == [4096, 4096]
    assert default_schedule(16384, 32) == [8192, 6144, 2048]
    for n in (1, 31, 2047, 2048, 6143, 6144, 9604, 11000, 40000, 100001):
        for dec in (False, True):
            sizes = group_sizes(n, default_schedule(n, 32, decode=dec), 32)
            assert sum(sizes) == n and max(sizes) <= 8192 + 32 and min(sizes) > 0
            assert all(g % 32 == 0 for g in sizes[:-1])


def test_slide_jobs_need_a_gpu():
    """No CPU fallback: the two-slides-in-flight runner refuses to start without a CUDA device."""
    from cnn_autoencoder_b200.jobs import SlideJobs
    if torch.cuda.is_available():
        pytest.skip('CUDA present: covered by tests/test_gpu_slide.py')
    with pytest.raises(RuntimeError):
        SlideJobs()
