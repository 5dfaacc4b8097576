"""Product entropy coder (C ABI, host) against the oracle's C restatement: the
streams must be byte-identical and both must decode each other's output."""
import numpy as np
import pytest
import torch

from cnn_autoencoder_b200 import _cabi as C
from cnn_autoencoder_b200 import _entropy as E
from oracle import cae_oracle as O
from oracle import rans as oracle_rans


def _tables(channels, seed, spread=0.0):
    torch.manual_seed(seed)
    eb = O.EntropyBottleneck(channels)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] -= torch.rand(channels) * spread
        eb.quantiles[:, 0, 2] += torch.rand(channels) * spread
        eb.quantiles[:, 0, 1] += torch.rand(channels) - 0.5
    eb.update(force=True)
    return (np.ascontiguousarray(eb._quantized_cdf.numpy(), dtype=np.int32),
            np.ascontiguousarray(eb._cdf_length.numpy(), dtype=np.int32),
            np.ascontiguousarray(eb._offset.numpy(), dtype=np.int32))


@pytest.mark.parametrize('c,hw,seed', [(1, 1, 0), (3, 7, 1), (16, 24, 2), (48, 64 * 64, 3)])
def test_streams_identical_to_oracle(c, hw, seed):
    cdf, sizes, offs = _tables(c, seed, spread=15.0)
    rng = np.random.default_rng(seed)
    sym = np.round(rng.normal(0, 8, size=(c, hw))).astype(np.int32)
    k = max(1, (c * hw) // 50)
    sym.reshape(-1)[rng.integers(0, c * hw, size=k)] = rng.integers(-70000, 70000, size=k)
    idx = np.repeat(np.arange(c, dtype=np.int32), hw)
    want = oracle_rans.encode_with_indexes(sym.reshape(-1), idx, cdf, sizes, offs)
    got = E.encode_symbols(sym, cdf, sizes, offs)
    assert got == want
    assert np.array_equal(E.decode_symbols(want, c, hw, cdf, sizes, offs), sym)
    assert np.array_equal(oracle_rans.decode_with_indexes(got, idx, cdf, sizes, offs).reshape(c, hw), sym)


def test_escape_lengths():
    # raw values needing 1..8 nibbles, incl. the unary-of-15 count path never reached (<=8 groups)
    cdf, sizes, offs = _tables(1, 9)
    vals = [0, 11, 12, -11, -12, 100, -100, 5000, -5000, 2 ** 20, -2 ** 20, 2 ** 26, -2 ** 26]
    sym = np.array([vals], dtype=np.int32)
    idx = np.zeros(len(vals), dtype=np.int32)
    want = oracle_rans.encode_with_indexes(sym.reshape(-1), idx, cdf, sizes, offs)
    assert E.encode_symbols(sym, cdf, sizes, offs) == want
    assert np.array_equal(E.decode_symbols(want, 1, len(vals), cdf, sizes, offs), sym)


def test_pmf_to_quantized_cdf_matches_oracle():
    rng = np.random.default_rng(0)
    L = C.lib()
    for n in (1, 2, 5, 64, 300):
        for trial in range(20):
            pmf = rng.random(n).astype(np.float32) ** 6
            pmf[rng.integers(0, n, size=n // 3)] = 0.0
            if pmf.sum() == 0:
                pmf[0] = 1.0
            pmf /= max(pmf.sum(), 1e-9) * rng.uniform(0.8, 1.3)
            if (np.round(pmf * 65536) > 0).sum() == 0:
                continue
            want = oracle_rans.pmf_to_quantized_cdf(pmf, 16)
            got = np.zeros(n + 1, dtype=np.uint32)
            C.check(L.cae_pmf_to_quantized_cdf(pmf.ctypes.data, n, 16, got.ctypes.data))
            assert np.array_equal(got, want)


def test_decode_rejects_garbage_sizes():
    cdf, sizes, offs = _tables(2, 1)
    with pytest.raises(C.CaeError):
        E.decode_symbols(b'\x00' * 6, 2, 4, cdf, sizes, offs)
    with pytest.raises(C.CaeError):
        bad = np.array([0.5, float('nan')], dtype=np.float32)
        out = np.zeros(3, dtype=np.uint32)
        C.check(C.lib().cae_pmf_to_quantized_cdf(bad.ctypes.data, 2, 16, out.ctypes.data))


def test_update_tables_match_oracle():
    torch.manual_seed(11)
    ref = O.EntropyBottleneck(5)
    with torch.no_grad():
        ref.quantiles[:, 0, 0] -= torch.rand(5) * 7
        ref.quantiles[:, 0, 2] += torch.rand(5) * 3
        ref._factor0.add_(torch.randn_like(ref._factor0) * 0.2)
    ref.update(force=True)
    mine = E.EntropyBottleneck(5)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    for k in ('_quantized_cdf', '_offset', '_cdf_length'):
        setattr(mine, k, sd[k])
    mine.load_state_dict(sd)
    mine.update(force=True)
    assert torch.equal(mine._quantized_cdf, ref._quantized_cdf)
    assert torch.equal(mine._cdf_length, ref._cdf_length)
    assert torch.equal(mine._offset, ref._offset)
    assert torch.allclose(mine.loss(), ref.loss())
    y = torch.randn(2, 5, 4, 4) * 6
    mine.eval(); ref.eval()
    with torch.no_grad():
        a, pa = mine._forward_torch(y)
        b, pb = ref(y)
    assert torch.equal(a, b) and torch.equal(pa, pb)
    with pytest.raises(C.CaeError):
        mine(y)   # CPU tensor in eval: no CPU fallback
