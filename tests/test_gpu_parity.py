"""Parity of the CUDA path (through the public API and the C ABI) against the CPU
oracle and the committed reference outputs.  Runs on the B200 box.

Bars (BASELINE.json north_star): quantized latents agree on >= 99.9 % of symbols
with flips only at rounding boundaries; decoded PSNR within 0.05 dB; estimated
bpp within 0.5 %; integer histograms and entropy-coded streams bit-exact given
identical symbols.
"""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
SMALL = sorted(glob.glob(os.path.join(GOLDEN, 'transforms_*.pt')))
NAMED = sorted(glob.glob(os.path.join(GOLDEN, 'named_*.pt')))


def _load(p):
    return torch.load(p, map_location='cpu', weights_only=False)


def _model(chk):
    import cnn_autoencoder_b200 as M
    return M.autoencoder_from_state_dict(chk, gpu=True, train=False)


def _symbol_report(y_gpu, y_ref, medians):
    """fraction of agreeing symbols, and the largest distance of a flipped value from its
    rounding boundary (flips may only happen where y - median is within eps of k + 0.5)."""
    med = medians.reshape(1, -1, 1, 1)
    s_gpu = torch.round(y_gpu - med)
    s_ref = torch.round(y_ref - med)
    flips = s_gpu != s_ref
    agree = 1.0 - flips.float().mean().item()
    frac = (y_ref - med) - torch.floor(y_ref - med)
    dist = (frac - 0.5).abs()[flips]
    return agree, (dist.max().item() if flips.any() else 0.0), int(flips.sum().item())


def _psnr(a, b):
    from oracle import cae_oracle as O
    return O.psnr_u8(a, b)


def _to_u8(x_r):
    return (x_r * 255.0).clip(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().cpu().numpy()


@pytest.mark.parametrize('path', SMALL, ids=[os.path.basename(p)[11:-3] for p in SMALL])
def test_small_archs_against_reference_outputs(path):
    g = _load(path)
    model = _model(g['checkpoint'])
    x = g['x_u8'].float() / 255.0
    y = model['encoder'](x.cuda()).cpu()
    assert y.shape == g['y'].shape
    med = g['checkpoint']['fact_ent']['quantiles'][:, 0, 1]
    agree, boundary, nflip = _symbol_report(y, g['y'], med)
    # tiny tensors (<= 1.5 k symbols): allow 2 boundary flips
    assert nflip <= max(2, 1e-3 * y.numel()), (agree, nflip)
    assert boundary < 0.02
    assert torch.allclose(y, g['y'], atol=2e-2, rtol=2e-2)
    x_r, fx_brg = model['decoder'](g['y_q'].cuda())
    assert len(x_r) == len(fx_brg) == g['arch']['compression_level']
    assert all(v is None for v in x_r[1:])
    img = g['x_u8'].permute(0, 2, 3, 1).numpy()
    d = abs(_psnr(img, _to_u8(x_r[0])) - _psnr(img, _to_u8(g['x_r'])))
    assert d <= 0.05, d
    assert torch.allclose(x_r[0].cpu(), g["x_r"], atol=1e-2, rtol=2e-2)
    # the uint8 epilogue is the truncating cast of the fp32 output
    _, _, u8 = model['decoder'](g['y_q'].cuda(), as_uint8=True)
    assert np.array_equal(u8.cpu().numpy(), _to_u8(x_r[0]))
    # uint8 HWC input path == fp32 NCHW input path
    y2 = model['encoder'](g['x_u8'].permute(0, 2, 3, 1).contiguous().cuda()).cpu()
    assert torch.equal(y2, y)


@pytest.mark.parametrize('path', NAMED, ids=[os.path.basename(p)[6:-3] for p in NAMED])
def test_named_archs_against_reference_outputs(path):
    from oracle import cae_oracle as O
    g = _load(path)
    chk = O.make_checkpoint(g['arch'], seed=g['seed'])
    model = _model(chk)
    y = model['encoder']((g['x_u8'].float() / 255.0).cuda()).cpu()
    med = chk['fact_ent']['quantiles'][:, 0, 1]
    agree, boundary, nflip = _symbol_report(y, g['y'], med)
    assert nflip <= max(2, 1e-3 * y.numel()), (agree, nflip)
    assert boundary < 0.02
    y_q, p_y = model['fact_ent'](g['y'].cuda())
    assert torch.equal(y_q.cpu(), g['y_q'])
    assert torch.allclose(p_y.cpu(), g['p_y'], rtol=2e-4, atol=1e-12)
    x_r, _ = model['decoder'](g['y_q'].cuda())
    img = g['x_u8'].permute(0, 2, 3, 1).numpy()
    d = abs(_psnr(img, _to_u8(x_r[0])) - _psnr(img, _to_u8(g['x_r'])))
    assert d <= 0.05, d


PARITY_LOG = os.environ.get('CAE_PARITY_LOG')     # measured gate values, one JSON line per case


@pytest.mark.parametrize('name,n,size', [('A', 4, 256), ('A_res', 2, 256), ('B', 2, 256),
                                         ('A', 1, 512), ('A_res', 1, 512), ('B', 1, 512)])
def test_full_size_pipeline_against_oracle(name, n, size):
    """configs[1]/[2] shapes: whole pipeline vs the oracle on the same seeded inputs."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200.pipeline import CodecPipeline
    torch.set_num_threads(os.cpu_count() or 8)
    chk = O.make_checkpoint(O.NAMED_ARCHS[name], seed=1234)
    model = _model(chk)
    oracle = O.OracleModel(chk)
    x_u8 = O.synth_natural(n, 3, size, size, seed=1)
    out = CodecPipeline(model)(x_u8.permute(0, 2, 3, 1).contiguous().cuda())
    ref = oracle.forward(x_u8.float() / 255.0)
    med = chk['fact_ent']['quantiles'][:, 0, 1]
    agree, boundary, nflip = _symbol_report(out['y'].cpu(), ref['y'], med)
    assert agree >= 0.999, (agree, nflip)
    assert boundary < 0.02, boundary
    img = x_u8.permute(0, 2, 3, 1).numpy()
    ref_u8 = _to_u8(ref['x_r'][0])
    d_psnr = abs(_psnr(img, out['x_r_u8'].cpu().numpy()) - _psnr(img, ref_u8))
    assert d_psnr <= 0.05, d_psnr
    bpp_ref = O.rate_loss(x_u8.float(), ref['p_y']).item()
    if PARITY_LOG:
        import json
        with open(PARITY_LOG, 'a') as f:
            f.write(json.dumps(dict(net=name, n=n, size=size, symbols=int(out['y'].numel()),
                                    flipped=nflip, flip_pct=round(100 * (1 - agree), 5),
                                    max_flip_distance_from_boundary=round(boundary, 6),
                                    abs_dpsnr_db=round(d_psnr, 6), est_bpp=round(out['bpp'].item(), 5),
                                    abs_dbpp_pct=round(100 * abs(out['bpp'].item() - bpp_ref) / bpp_ref, 6),
                                    gates='flip_pct <= 0.1, boundary < 0.02, dPSNR <= 0.05 dB, dbpp <= 0.5 %')) + '\n')
    assert abs(out['bpp'].item() - bpp_ref) <= 0.005 * bpp_ref
    # histogram == bincount of the symbols the kernel itself produced (integer work: exact)
    sym, hist, bits = model['fact_ent'].module.symbols_hist_rate(out['y'])
    tb = model['fact_ent'].module._device_tables()
    s = sym.cpu().numpy()
    for c in (0, s.shape[1] // 2, s.shape[1] - 1):
        idx = np.clip(s[:, c].reshape(-1) - tb['lut_min'], 0, tb['lut_len'] - 1)
        assert np.array_equal(np.bincount(idx, minlength=tb['lut_len']), hist[c].cpu().numpy())
    assert hist.sum().item() == s.size


def test_entropy_coder_streams_bit_exact_given_identical_symbols():
    from oracle import cae_oracle as O
    chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
    model = _model(chk)
    oracle = O.OracleModel(chk)
    g = torch.Generator().manual_seed(3)
    y = torch.randn(2, 48, 16, 24, generator=g) * 9
    y[0, 0, 0, 0] = 4000.2
    y[1, 47, 15, 23] = -77777.0
    mine = model['fact_ent'].module.compress(y.cuda())
    ref = oracle.fact_ent.compress(y)
    assert mine == ref
    back = model['fact_ent'].module.decompress(mine, size=(16, 24))
    assert torch.equal(back.cpu(), oracle.fact_ent.decompress(ref, (16, 24)))
    assert torch.equal(back.cpu(), torch.round(y))


def test_cae_codec_roundtrip_against_oracle_codec():
    from oracle import cae_oracle as O
    import cnn_autoencoder_b200 as M
    arch = dict(channels_org=3, channels_net=32, channels_bn=16, compression_level=3,
                act_layer_type='LeakyReLU')
    chk = O.make_checkpoint(arch, seed=77)
    codec = M.ConvolutionalAutoencoder(checkpoint=chk, gpu=True)
    oracle = O.OracleModel(chk)
    tile = O.synth_tissue_tile(1, 2, ps=128, seed=2)
    enc = codec.encode(tile)
    enc_ref = oracle.codec_encode(tile)
    assert enc[:16] == enc_ref[:16]
    # the oracle decodes the product's stream into the product's symbols (format parity)
    sym = oracle.fact_ent.decompress([enc[16:]], (16, 16))
    sym_ref = oracle.fact_ent.decompress([enc_ref[16:]], (16, 16))
    assert (sym == sym_ref).float().mean().item() >= 0.999
    assert abs(len(enc) - len(enc_ref)) <= 0.005 * len(enc_ref) + 8
    rec = codec.decode(enc)
    rec_ref = oracle.codec_decode(enc_ref)
    assert rec.shape == tile.shape and rec.dtype == np.uint8
    assert abs(O.psnr_u8(tile, rec) - O.psnr_u8(tile, rec_ref)) <= 0.05
    out = np.empty_like(tile)
    assert codec.decode(enc, out=out) is not None and np.array_equal(out, rec)


def test_cae_bn_codec_roundtrip():
    from oracle import cae_oracle as O
    import cnn_autoencoder_b200 as M
    chk = O.make_checkpoint(O.NAMED_ARCHS['M'], seed=9)
    model = _model(chk)
    codec = M.ConvolutionalAutoencoderBottleneck(channels_bn=16, fact_ent=model['fact_ent'].module,
                                                 gpu=True)
    cfg = codec.get_config()
    assert cfg['id'] == 'cae_bn' and cfg['channels_bn'] == 16
    lat = (np.random.default_rng(0).normal(0, 5, size=(8, 12, 16))).astype(np.float32)
    enc = codec.encode(lat)
    dec = codec.decode(enc)
    assert dec.shape == lat.shape and dec.dtype == np.float32
    assert np.array_equal(dec, np.round(lat))
    oracle = O.OracleModel(chk)
    y = torch.from_numpy(lat).permute(2, 0, 1).unsqueeze(0)
    assert enc[16:] == oracle.fact_ent.compress(y)[0]


def test_step_closure_and_criterion_match_oracle():
    from oracle import cae_oracle as O
    import cnn_autoencoder_b200 as M
    arch = O.NAMED_ARCHS['A']
    chk = O.make_checkpoint(arch, seed=1234)
    model = _model(chk)
    oracle = O.OracleModel(chk)
    x_u8 = O.synth_natural(2, 3, 128, 128, seed=8)
    x = x_u8.float() / 255.0
    fwd = M.decorate_trainable_modules(trainable_modules=[],
                                       enabled_modules=['encoder', 'decoder', 'fact_ent'])
    out = fwd(x.cuda(), model)
    assert set(out) == {'x_r', 'fx_brg', 'y', 'y_q', 'p_y', 't_pred', 't_aux_pred', 's_pred',
                        's_aux_pred'}
    crit = M.setup_loss('RateMSE', distortion_lambda=0.01)
    loss = crit(inputs=x.cuda(), outputs=out, net=model)
    ref = oracle.forward(x)
    want = O.general_loss(x, ref, oracle.fact_ent, distortion_lambda=0.01)
    assert abs(loss['rate_loss'].item() - want['rate_loss'].item()) <= 0.005 * want['rate_loss'].item()
    assert abs(loss['dist'][0].item() - want['dist'][0].item()) <= 0.02 * want['dist'][0].item()
    assert torch.allclose(loss['entropy_loss'].cpu(), want['entropy_loss'], rtol=1e-4)


def test_no_cpu_fallback():
    from oracle import cae_oracle as O
    import cnn_autoencoder_b200 as M
    from cnn_autoencoder_b200 import _cabi
    chk = O.make_checkpoint(O.NAMED_ARCHS['M'], seed=1)
    model = M.autoencoder_from_state_dict(chk, gpu=False, train=False)
    with pytest.raises(_cabi.CaeError):
        model['encoder'](torch.zeros(1, 1, 32, 32))
    with pytest.raises(_cabi.CaeError):
        model['fact_ent'](torch.zeros(1, 16, 4, 4))


def test_ragged_and_edge_sizes():
    """Edge chunks of a slide (336 px) and non-multiples of the CTA tile."""
    from oracle import cae_oracle as O
    arch = dict(channels_org=3, channels_net=32, channels_bn=16, compression_level=3,
                act_layer_type='LeakyReLU')
    chk = O.make_checkpoint(arch, seed=21)
    model = _model(chk)
    oracle = O.OracleModel(chk)
    for (h, w) in ((336, 336), (8, 8), (40, 264), (16, 8)):
        x_u8 = O.synth_natural(1, 3, h, w, seed=h + w)
        x = x_u8.float() / 255.0
        y = model['encoder'](x.cuda()).cpu()
        ref = oracle.forward(x)
        med = chk['fact_ent']['quantiles'][:, 0, 1]
        agree, boundary, nflip = _symbol_report(y, ref['y'], med)
        assert nflip <= max(2, 1e-3 * y.numel()) and boundary < 0.02, (h, w, agree)
        x_r, _ = model['decoder'](ref['y_q'].cuda())
        assert torch.allclose(x_r[0].cpu(), ref['x_r'][0], atol=4e-3, rtol=1e-2), (h, w)


def test_tile_loop_roundtrip_and_sharding(tmp_path):
    """configs[2]: compress_image -> decompress_image over a small synthetic slide with edge
    chunks, two shards written by two calls (rank 0/2 and 1/2), checked chunk by chunk against
    the oracle codec."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import compress, decompress, _store
    arch = dict(channels_org=3, channels_net=32, channels_bn=16, compression_level=3,
                act_layer_type='LeakyReLU')
    chk = O.make_checkpoint(arch, seed=31)
    oracle = O.OracleModel(chk)
    ps = 128
    slide = np.concatenate([np.concatenate([O.synth_tissue_tile(i, j, ps=ps, seed=2)
                                            for j in range(3)], axis=1) for i in range(2)], axis=0)
    slide = slide[:200, :300]                       # ragged: edge chunks of 72 and 44 px
    out = str(tmp_path / 'slide.zarr')
    stats = [compress.compress_image('CAE', chk, slide, out, patch_size=ps, rank=r, world_size=2,
                                     batch_tiles=4) for r in (0, 1)]
    assert sum(s['tiles'] for s in stats) == 6
    arr = _store.DirArray(os.path.join(out, '0/0'), mode='r')
    assert arr.compressor_config['id'] == 'cae' and arr.grid == (2, 3, 1)
    for i in range(2):
        for j in range(3):
            tile = _store.padded_tile(slide, i * ps, j * ps, ps)
            ref = oracle.codec_encode(tile)
            got = arr.read_encoded((i, j, 0))
            assert got[:16] == ref[:16]
            s_ref = oracle.fact_ent.decompress([ref[16:]], (16, 16))
            s_got = oracle.fact_ent.decompress([got[16:]], (16, 16))
            assert (s_ref == s_got).float().mean().item() >= 0.999
    rec_dir = str(tmp_path / 'rec.zarr')
    for r in (0, 1):
        decompress.decompress_image(out, rec_dir, checkpoint=chk, rank=r, world_size=2,
                                    batch_tiles=4)
    rec = _store.DirArray(os.path.join(rec_dir, 'decompressed/0/0'), mode='r')
    full = np.zeros_like(slide)
    for i in range(2):
        for j in range(3):
            full[rec.chunk_slices((i, j, 0))[:2]] = rec.read_chunk((i, j, 0))
    ref_full = np.zeros_like(slide)
    for i in range(2):
        for j in range(3):
            tile = _store.padded_tile(slide, i * ps, j * ps, ps)
            r_tile = oracle.codec_decode(oracle.codec_encode(tile))
            sl = rec.chunk_slices((i, j, 0))
            ref_full[sl[:2]] = r_tile[:sl[0].stop - sl[0].start, :sl[1].stop - sl[1].start]
    assert abs(O.psnr_u8(slide, full) - O.psnr_u8(slide, ref_full)) <= 0.05
    bpp = O.bpp(arr.nbytes_stored(), *slide.shape[:2])
    bpp_ref = O.bpp(sum(len(oracle.codec_encode(_store.padded_tile(slide, i * ps, j * ps, ps)))
                        for i in range(2) for j in range(3)), *slide.shape[:2])
    assert abs(bpp - bpp_ref) <= 0.005 * bpp_ref


@pytest.mark.parametrize('residual', [False, True])
def test_gdn_archs_against_oracle(residual):
    """act_layer_type='GDN' (R:29-30): GDN in the analysis track, IGDN in the synthesis track."""
    from oracle import cae_oracle as O
    arch = dict(channels_org=3, channels_net=32, channels_bn=16, compression_level=3,
                act_layer_type='GDN', use_residual=residual)
    chk = O.make_checkpoint(arch, seed=44)
    g = torch.Generator().manual_seed(2)
    for part in ('encoder', 'decoder'):          # make beta / gamma non-trivial
        for k, v in chk[part].items():
            if k.endswith('.gamma'):
                v.add_(torch.rand(v.shape, generator=g) * 0.05)
            elif k.endswith('.beta'):
                v.add_(torch.rand(v.shape, generator=g) * 0.3)
    model = _model(chk)
    oracle = O.OracleModel(chk)
    x_u8 = O.synth_natural(2, 3, 64, 96, seed=6)
    x = x_u8.float() / 255.0
    ref = oracle.forward(x)
    y = model['encoder'](x.cuda()).cpu()
    med = chk['fact_ent']['quantiles'][:, 0, 1]
    agree, boundary, nflip = _symbol_report(y, ref['y'], med)
    assert nflip <= max(2, 2e-3 * y.numel()) and boundary < 0.03, (agree, nflip, boundary)
    x_r, _ = model['decoder'](ref['y_q'].cuda())
    img = x_u8.permute(0, 2, 3, 1).numpy()
    assert abs(_psnr(img, _to_u8(x_r[0])) - _psnr(img, _to_u8(ref['x_r'][0]))) <= 0.05
    # random-init IGDN chains amplify (|x_r| reaches the hundreds): compare in relative L2
    rel = (x_r[0].cpu() - ref['x_r'][0]).norm() / ref['x_r'][0].norm()
    assert rel < 2e-2, rel


def test_device_entropy_coder_is_byte_identical_to_host_coder():
    """cae_rans_encode_batch / cae_rans_decode_batch vs. the host coder (itself byte-identical
    to the oracle, tests/test_host_coder.py): same bytes per stream, escapes included."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import _entropy as E
    chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
    fe = _model(chk)['fact_ent'].module
    n, c, hw = 70, 48, 24 * 20
    g = torch.Generator().manual_seed(8)
    sym = torch.round(torch.randn(n, c, hw, generator=g) * 7).int()
    idx = torch.randint(0, sym.numel(), (400,), generator=g)
    sym.view(-1)[idx] = torch.randint(-90000, 90000, (400,), generator=g, dtype=torch.int32)
    sym[3, 0, 0] = 2 ** 27
    sym[3, 47, hw - 1] = -(2 ** 27)
    streams = fe.encode_symbols_gpu(sym.cuda())           # reciprocal-multiply table path
    cdf, sizes, offs = fe._host_tables()
    for k in range(n):
        assert streams[k] == E.encode_symbols(sym[k].numpy(), cdf, sizes, offs), k
    os.environ['CAE_RANS_NO_TABLE'] = '1'                  # plain 64-bit division path
    try:
        assert fe.encode_symbols_gpu(sym.cuda()) == streams
    finally:
        del os.environ['CAE_RANS_NO_TABLE']
    back = fe.decode_streams_gpu(streams, hw).cpu()
    assert torch.equal(back, sym)
    # through the module API (>= GPU_CODER_MIN_STREAMS tiles): compress / decompress
    y = torch.randn(64, 48, 8, 8, generator=g) * 6
    fe.GPU_CODER_MIN_STREAMS = 32           # instance override: force the device coder
    s2 = fe.compress(y.cuda())
    assert len(s2) == 64
    assert s2[5] == O.OracleModel(chk).fact_ent.compress(y[5:6])[0]
    assert torch.equal(fe.decompress(s2, (8, 8)).cpu(), torch.round(y))


def test_graphed_pipeline_matches_eager():
    """The CUDA-graph replay of the pipeline (what bench.py times) returns exactly what the
    eager call returns, also after the static input buffer is refilled."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200.pipeline import CodecPipeline
    chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=7)
    model = _model(chk)
    pipe = CodecPipeline(model)
    x0 = O.synth_natural(4, 3, 64, 96, seed=3).permute(0, 2, 3, 1).contiguous().cuda()
    x1 = O.synth_natural(4, 3, 64, 96, seed=4).permute(0, 2, 3, 1).contiguous().cuda()
    static = x0.clone()
    g = pipe.graphed(static)
    assert g.launches >= 8
    for x in (x0, x1, x0):
        static.copy_(x)
        got = g.replay()
        want = pipe(x)
        torch.cuda.synchronize()
        assert torch.equal(got['x_r_u8'], want['x_r_u8'])
        assert torch.equal(got['y_q'], want['y_q'])
        assert torch.equal(got['hist'], want['hist'])
        assert got['bits'].item() == want['bits'].item()


def test_fused_quantizer_matches_standalone():
    """The quantizer fused into the latent layer's epilogue (cae_conv_desc.quant) against the
    stand-alone cae_eb_quantize pass on the same latent: symbols, y_q, histogram and the
    reconstruction are identical, the rate agrees to float summation order."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200.pipeline import CodecPipeline
    for arch, shape in (('A', (3, 64, 96)), ('A', (2, 40, 72)), ('B', (2, 64, 64))):
        chk = O.make_checkpoint(O.NAMED_ARCHS[arch], seed=11)
        model = _model(chk)
        pipe = CodecPipeline(model)
        n, h, w = shape
        x = O.synth_natural(n, 3, h, w, seed=5).permute(0, 2, 3, 1).contiguous().cuda()
        pipe.fuse_quantizer = True
        fused = pipe(x)
        eb = model['fact_ent'].module
        req = eb.quant_request(want_sym=True)
        y = model['encoder'](x, quant=req)
        assert req.done, 'the latent layer of the named nets is a tensor-core layer'
        pipe.fuse_quantizer = False
        plain = pipe(x)
        sym_ref, hist_ref, rate_ref = eb.symbols_hist_rate(y)
        torch.cuda.synchronize()
        assert torch.equal(fused['y'], plain['y'])
        assert torch.equal(fused['y_q'], plain['y_q'])
        assert torch.equal(req.sym, sym_ref)
        assert torch.equal(fused['hist'], plain['hist']) and torch.equal(req.hist, hist_ref)
        assert int(fused['hist'].sum()) == fused['y'].numel()
        assert abs(fused['bits'].item() - plain['bits'].item()) <= 1e-5 * abs(plain['bits'].item())
        assert torch.equal(fused['x_r_u8'], plain['x_r_u8'])
        assert int(req.status.item()) == 0


def test_multiscale_colour_heads_match_oracle():
    """multiscale_analysis=True (R:417-429): the reflect-padded colour heads on the intermediate
    scales, run on the kept planar tensors, against the oracle (pinned to the reference for this
    switch in tests/test_oracle_vs_reference.py)."""
    from oracle import cae_oracle as O
    arch = dict(channels_org=3, channels_net=32, channels_bn=16, compression_level=3,
                act_layer_type='LeakyReLU', multiscale_analysis=True, bias=True)
    chk = O.make_checkpoint(arch, seed=5)
    model = _model(chk)
    om = O.OracleModel(chk)
    y_q = torch.round(torch.randn(2, 16, 8, 12, generator=torch.Generator().manual_seed(2)) * 4)
    x_r, fx_brg = model['decoder'](y_q.cuda())
    ref, _ = om.decoder(y_q)
    assert len(x_r) == 3 and all(t is not None for t in x_r)
    for got, want in zip(x_r, ref):
        assert got.shape == want.shape
        assert torch.allclose(got.cpu(), want, atol=2e-2, rtol=2e-2), (got.cpu() - want).abs().max()
    # the codec path is unchanged by the switch
    _, _, u8 = model['decoder'](y_q.cuda(), as_uint8='only')
    assert u8.shape == (2, 64, 96, 3)


@pytest.mark.parametrize('shape,c_org', [((2, 64, 96), 3), ((1, 40, 56), 1), ((2, 40, 72), 4),
                                         ((1, 256, 256), 3)])
def test_fused_head_matches_the_two_kernel_path(shape, c_org):
    """cae_conv_head (stem + first stride-2 layer in one launch) against the unfused pair of
    kernels on the same weights: same fp16 rounding points, so the latents agree to fp16
    accumulation-order noise and both sit within the oracle tolerance."""
    from oracle import cae_oracle as O
    arch = dict(channels_org=c_org, channels_net=32, channels_bn=16, compression_level=3,
                act_layer_type='LeakyReLU', bias=True)
    chk = O.make_checkpoint(arch, seed=21)
    model = _model(chk)
    n, h, w = shape
    x = O.synth_natural(n, c_org, h, w, seed=9)
    x_in = x.permute(0, 2, 3, 1).contiguous().cuda() if c_org == 3 else (x.float() / 255.0).cuda()
    enc = model['encoder']
    ex = enc.module._executor()
    ex.fuse_head = True
    y_fused = enc(x_in).clone()
    assert ex.last_calls[0][0] == 'head'
    ex.fuse_head = False
    y_plain = enc(x_in).clone()
    assert ex.last_calls[0][0] != 'head'
    ref = O.OracleModel(chk).encoder(x.float() / 255.0)
    assert torch.allclose(y_fused.cpu(), ref, atol=2e-2, rtol=2e-2)
    assert torch.allclose(y_fused, y_plain, atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize('shape,c_org', [((2, 8, 8), 3), ((3, 16, 24), 3), ((1, 20, 12), 3),
                                         ((2, 7, 5), 1), ((2, 32, 32), 2)])
def test_projection_fusion_matches_the_two_kernel_path(shape, c_org):
    """cae_conv_desc.proj (the last 128-channel transposed layer projected onto the image
    layer's taps in its epilogue + cae_image_from_proj) against the two tensor-core layers run
    one after the other, and against the oracle.  shape = latent n x h x w (image 8x larger)."""
    from oracle import cae_oracle as O
    arch = dict(O.NAMED_ARCHS['A'], channels_org=c_org)
    chk = O.make_checkpoint(arch, seed=31)
    model = _model(chk)
    n, h, w = shape
    g = torch.Generator().manual_seed(3)
    y_q = torch.round(torch.randn(n, arch['channels_bn'], h, w, generator=g) * 3.0)
    dec = model['decoder']
    ex = dec.module._executor()
    last = len(ex.steps) - 1
    ex.fuse_proj = True
    x_f, _, u8_f = dec(y_q.cuda(), as_uint8=True)
    assert ex.last_calls[last][0] == 'image_from_proj'
    x_f, u8_f = x_f[0].clone(), u8_f.clone()
    ex.fuse_proj = False
    x_p, _, u8_p = dec(y_q.cuda(), as_uint8=True)
    assert ex.last_calls[last][0] != 'image_from_proj'
    ex.fuse_proj = True
    ref = O.OracleModel(chk).decoder(y_q)[0][0]
    # the records are fp16: one extra rounding of each of the up to four partial sums
    assert torch.allclose(x_f, x_p[0], atol=2e-3, rtol=2e-3), (x_f - x_p[0]).abs().max()
    assert torch.allclose(x_f.cpu(), ref, atol=4e-3, rtol=1e-2), (x_f.cpu() - ref).abs().max()
    assert np.array_equal(u8_f.cpu().numpy(), _to_u8(x_f))
    d = (u8_f.int() - u8_p.int()).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 0.05


def test_cta_pair_form_matches_the_single_cta_kernels(tmp_path):
    """The opt-in cta_group::2 form of the 128-channel layers (CAE_IGEMM_PAIR_MMA: M = 256 across
    a CTA pair, each CTA holding half of the weights, resident when they fit) against the default
    single-CTA kernels: same operands and accumulation order, so latent and image agree to fp16
    noise.  Knobs are read once per process, hence the child process."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle import cae_oracle as O
import cnn_autoencoder_b200 as M
chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=5)
model = M.autoencoder_from_state_dict(chk, gpu=True, train=False)
x = O.synth_natural(3, 3, 96, 160, seed=2).permute(0, 2, 3, 1).contiguous().cuda()
y = model['encoder'](x)
y_q = torch.round(y)
x_r, _, u8 = model['decoder'](y_q, as_uint8=True)
torch.cuda.synchronize()
torch.save(dict(y=y.cpu(), x_r=x_r[0].cpu(), u8=u8.cpu()), sys.argv[1])
''' % root
    outs = {}
    for name, env in (('single', {}), ('pair', {'CAE_DEBUG': '1', 'CAE_IGEMM_PAIR_MMA': '1'})):
        path = str(tmp_path / (name + '.pt'))
        e = dict(os.environ)
        e.pop('CAE_IGEMM_PAIR_MMA', None)
        e.update(env)
        subprocess.run([sys.executable, '-c', script, path], check=True, env=e, timeout=300)
        outs[name] = torch.load(path)
    a, b = outs['single'], outs['pair']
    assert torch.allclose(a['y'], b['y'], atol=2e-3, rtol=2e-3), (a['y'] - b['y']).abs().max()
    # y_q = round(y) may differ in a symbol sitting on a rounding boundary; compare the image loosely
    d = (a['u8'].int() - b['u8'].int()).abs()
    assert float((d > 1).float().mean()) < 1e-3


def test_training_mode_bottleneck_kernels_match_autograd():
    """``cae_eb_train_fwd`` / ``cae_eb_train_bwd`` against the same model written with torch
    autograd ops (CompressAI's formulation): outputs, d/dy and the gradient of
    every density parameter, with noise fixed through ``noise_fn``."""
    from cnn_autoencoder_b200._entropy import EntropyBottleneck
    torch.manual_seed(11)
    eb = EntropyBottleneck(6).cuda().train()
    with torch.no_grad():
        eb._factor1.add_(torch.randn_like(eb._factor1) * 0.4)
        eb._factor2.add_(torch.randn_like(eb._factor2) * 0.4)
        eb._matrix2.add_(torch.randn_like(eb._matrix2) * 0.3)
    y0 = (torch.randn(3, 6, 9, 7, device='cuda') * 6)
    y0[0, 0, 0, 0] = 400.0                       # far tail: likelihood on the bound
    noise = (torch.rand(6, 1, 3 * 63, device='cuda') - 0.5)
    eb.noise_fn = lambda v: noise
    w_lik = torch.randn(3, 6, 9, 7, device='cuda')
    w_hat = torch.randn(3, 6, 9, 7, device='cuda')

    def run(kernel):
        eb.zero_grad()
        y = y0.clone().requires_grad_(True)
        if kernel:
            y_hat, lik = eb(y)
        else:
            y_hat, lik = eb._forward_torch(y, training=True)
        loss = (torch.log2(lik) * w_lik).sum() + (y_hat * w_hat).sum()
        loss.backward()
        return y_hat.detach(), lik.detach(), y.grad.clone(), {n: p.grad.clone() for n, p in eb.named_parameters()
                                                             if p.grad is not None}
    a = run(True)
    b = run(False)
    assert torch.allclose(a[0], b[0], atol=1e-6)
    assert torch.allclose(a[1], b[1], rtol=2e-5, atol=1e-12)
    assert torch.allclose(a[2], b[2], rtol=2e-4, atol=1e-6)
    assert set(a[3]) == set(b[3]) and len(a[3]) >= 14
    for k in a[3]:
        scale = b[3][k].abs().max().item() + 1e-12
        assert (a[3][k] - b[3][k]).abs().max().item() <= 2e-4 * scale + 1e-7, k


def test_ssim_and_delta_e_kernels_match_the_oracle():
    """cae_ssim_u8 / cae_delta_e_u8 (metrics.ssim, metrics.delta_cielab: compute_ssim and
    compute_deltaCIELAB of src/test_cae.py on the device) against the float64 restatement of the
    scikit-image algorithms, on a tissue-like image and its noisy / shifted versions, ragged
    sizes included."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import metrics
    rng = np.random.default_rng(1)
    for h, w in ((64, 64), (71, 45), (200, 333)):
        x = O.synth_natural(1, 3, h, w, seed=h)[0].permute(1, 2, 0).contiguous().numpy()
        noisy = np.clip(x.astype(np.int32) + rng.integers(-20, 21, size=x.shape), 0, 255).astype(np.uint8)
        dark = (x * 0.8).astype(np.uint8)
        for y in (x, noisy, dark):
            got, want = metrics.ssim(x, y), O.ssim_u8(x, y)
            assert abs(got - want) < 1e-9, (h, w, got, want)
            got, want = metrics.delta_cielab(x, y), O.delta_cielab_u8(x, y)
            assert abs(got - want) <= 2e-4 * max(1.0, want), (h, w, got, want)
    # batch form
    xb = O.synth_natural(3, 3, 48, 40, seed=9).permute(0, 2, 3, 1).contiguous()
    yb = (xb.float() * 0.9).to(torch.uint8)
    per = metrics.ssim(xb.cuda(), yb.cuda(), per_image=True).cpu().numpy()
    for i in range(3):
        assert abs(per[i] - O.ssim_u8(xb[i].numpy(), yb[i].numpy())) < 1e-9


def test_ms_ssim_kernels_match_the_oracle():
    """metrics.ms_ssim (compute_ms_ssim of src/test_cae.py:46-50 on the device: Gaussian-window
    moments, average pooling with the odd-size padding, five scales) against the float64
    restatement of pytorch_msssim's algorithm; even and odd sizes."""
    from oracle import cae_oracle as O
    from cnn_autoencoder_b200 import metrics
    rng = np.random.default_rng(2)
    for h, w in ((176, 192), (201, 187), (320, 163)):
        x = O.synth_natural(1, 3, h, w, seed=w)[0].permute(1, 2, 0).contiguous().numpy()
        noisy = np.clip(x.astype(np.int32) + rng.integers(-25, 26, size=x.shape), 0, 255).astype(np.uint8)
        dark = (x * 0.7).astype(np.uint8)
        for y in (x, noisy, dark):
            got, want = metrics.ms_ssim(x, y), O.ms_ssim_u8(x, y)
            assert abs(got - want) < 2e-5, (h, w, got, want)
    with pytest.raises(ValueError):
        metrics.ms_ssim(np.zeros((100, 200, 3), np.uint8), np.zeros((100, 200, 3), np.uint8))
