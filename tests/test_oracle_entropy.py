"""Known-answer and round-trip tests pinning the restated CompressAI pieces
(parity unpinned upstream -- SURVEY.md 8c)."""
import os

import numpy as np
import pytest
import torch

from oracle import cae_oracle as O
from oracle import rans

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def test_init_analytic_kat():
    # 8c(iii): at init logits_cum is affine x/10 + c, quantiles (-10,0,10)
    torch.manual_seed(0)
    eb = O.EntropyBottleneck(5)
    eb.update(force=True)
    assert eb._offset.tolist() == [-10] * 5
    assert eb._cdf_length.tolist() == [23] * 5
    assert eb._quantized_cdf.shape == (5, 23)
    v = torch.linspace(-8, 8, 9).reshape(1, 1, -1).repeat(5, 1, 1)
    lc = eb._logits_cumulative(v, stop_gradient=True)
    slope = (lc[:, 0, 1:] - lc[:, 0, :-1]) / 2.0
    assert torch.allclose(slope, torch.full_like(slope, 0.1), atol=1e-5)
    eb.eval()
    y = torch.zeros(1, 5, 4, 4)
    _, p = eb(y)
    assert 0.012 < p.min().item() and p.max().item() < 0.0251
    bits = -torch.log2(p).mean().item()
    assert 5.2 < bits < 6.4


def test_cdf_invariants():
    torch.manual_seed(3)
    eb = O.EntropyBottleneck(7)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] -= torch.rand(7) * 20
        eb.quantiles[:, 0, 2] += torch.rand(7) * 20
    eb.update(force=True)
    cdf = eb._quantized_cdf.numpy()
    for c in range(7):
        n = int(eb._cdf_length[c])
        row = cdf[c, :n]
        assert row[0] == 0 and row[-1] == 1 << 16
        assert (np.diff(row) > 0).all()


def test_pmf_to_quantized_cdf_steals_for_zero_bins():
    pmf = np.array([0.5, 0.0, 1e-9, 0.5 - 1e-9, 0.0], dtype=np.float32)
    cdf = rans.pmf_to_quantized_cdf(pmf, 16)
    assert cdf[0] == 0 and cdf[-1] == 65536 and (np.diff(cdf.astype(np.int64)) > 0).all()
    with pytest.raises(ValueError):
        rans.pmf_to_quantized_cdf(np.array([0.5, -0.1], dtype=np.float32))


@pytest.mark.parametrize('n', [1, 2, 17, 4096])
def test_rans_roundtrip_with_escapes(n):
    rng = np.random.default_rng(n)
    C = 3
    torch.manual_seed(1)
    eb = O.EntropyBottleneck(C)
    eb.update(force=True)
    idx = rng.integers(0, C, size=n).astype(np.int32)
    sym = np.round(rng.normal(0, 9, size=n)).astype(np.int32)
    sym[rng.integers(0, n, size=max(1, n // 8))] = rng.integers(-100000, 100000, size=max(1, n // 8))
    s = rans.encode_with_indexes(sym, idx, eb._quantized_cdf.numpy(), eb._cdf_length.numpy(),
                                 eb._offset.numpy())
    assert len(s) % 4 == 0 and len(s) >= 8
    back = rans.decode_with_indexes(s, idx, eb._quantized_cdf.numpy(), eb._cdf_length.numpy(),
                                    eb._offset.numpy())
    assert np.array_equal(back, sym)


def test_entropy_golden_regression():
    g = torch.load(os.path.join(GOLDEN, 'entropy_kat.pt'), map_location='cpu', weights_only=False)
    eb = O.EntropyBottleneck(6)
    O.eb_load_state_dict(eb, g['state'])
    eb.eval()
    assert torch.equal(eb._quantized_cdf, g['quantized_cdf'])
    assert torch.equal(eb._cdf_length, g['cdf_length'])
    assert torch.equal(eb._offset, g['offset'])
    with torch.no_grad():
        y_q, p_y = eb(g['y'])
    assert torch.equal(y_q, g['y_q'])
    assert torch.allclose(p_y, g['p_y'], rtol=1e-6, atol=1e-12)
    strings = eb.compress(g['y'])
    for s, ref in zip(strings, g['strings']):
        assert np.array_equal(np.frombuffer(s, dtype=np.uint8), ref)
    assert torch.equal(eb.decompress(strings, (5, 7)), g['y_q'])
    assert torch.equal(eb.symbols(g['y']), g['symbols'])


def test_hist_rate_equals_elementwise_rate():
    # 8c(iv): in eval the likelihood is a function of (channel, integer symbol)
    torch.manual_seed(4)
    eb = O.EntropyBottleneck(4)
    eb.update(force=True)
    eb.eval()
    y = torch.randn(2, 4, 16, 16) * 5
    with torch.no_grad():
        _, p = eb(y)
    rate = -torch.log2(p).sum().item()
    sym = eb.symbols(y)
    total = 0.0
    for c in range(4):
        vals, counts = torch.unique(sym[:, c], return_counts=True)
        with torch.no_grad():
            _, pc = eb(vals.float().reshape(1, 1, -1, 1).expand(1, 4, -1, 1).contiguous())
        total += (-torch.log2(pc[0, c, :, 0]) * counts).sum().item()
    assert abs(total - rate) < 1e-3 * abs(rate)


def test_codec_roundtrip_and_metrics():
    chk = O.make_checkpoint(dict(channels_org=3, channels_net=8, channels_bn=8,
                                 compression_level=2, act_layer_type='LeakyReLU'), seed=5)
    om = O.OracleModel(chk)
    tile = O.synth_tissue_tile(0, 1, ps=32, seed=2)
    buf = om.codec_encode(tile)
    assert buf[:16] == (32).to_bytes(8, 'big') * 2
    rec = om.codec_decode(buf)
    assert rec.shape == tile.shape and rec.dtype == np.uint8
    assert np.isfinite(O.psnr_u8(tile, rec))
    assert O.bpp(len(buf), 32, 32) > 0
