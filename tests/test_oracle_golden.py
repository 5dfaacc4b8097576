"""The oracle restatement against the committed reference outputs (CPU)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import cae_oracle as O
from oracle.make_golden import state_sha256

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
SMALL = sorted(glob.glob(os.path.join(GOLDEN, 'transforms_*.pt')))
NAMED = sorted(glob.glob(os.path.join(GOLDEN, 'named_*.pt')))


def _load(p):
    return torch.load(p, map_location='cpu', weights_only=False)


@pytest.mark.parametrize('path', SMALL, ids=[os.path.basename(p)[11:-3] for p in SMALL])
def test_small_archs_match_reference_outputs(path):
    torch.set_num_threads(1)
    g = _load(path)
    om = O.OracleModel(g['checkpoint'])
    out = om.forward(g['x_u8'].float() / 255.0)
    # same torch ops in the same order: bit-exact on the same build; allow 1 ulp-scale
    # slack for a different CPU's vectorisation of conv accumulation
    assert torch.allclose(out['y'], g['y'], rtol=0, atol=2e-6)
    assert torch.allclose(out['x_r'][0], g['x_r'], rtol=0, atol=2e-6)
    assert torch.equal(torch.round(out['y']), torch.round(g['y'])) or \
        (torch.round(out['y']) != torch.round(g['y'])).float().mean() < 1e-3


@pytest.mark.parametrize('path', NAMED, ids=[os.path.basename(p)[6:-3] for p in NAMED])
def test_named_archs_match_reference_outputs(path):
    torch.set_num_threads(1)
    g = _load(path)
    chk = O.make_checkpoint(g['arch'], seed=g['seed'])
    assert state_sha256(chk) == g['sha256'], 'seeded weight generator drifted'
    om = O.OracleModel(chk)
    out = om.forward(g['x_u8'].float() / 255.0)
    assert torch.allclose(out['y'], g['y'], rtol=0, atol=5e-6)
    assert torch.allclose(out['x_r'][0], g['x_r'], rtol=0, atol=5e-6)
    assert torch.allclose(out['p_y'], g['p_y'], rtol=1e-5, atol=1e-9)


def test_plan_param_counts():
    # parameter counts measured on the reference classes
    for name, (enc_n, dec_n) in dict(A=(353745, 374400), B=(1256994, 1920384),
                                     M=(2385, 4680)).items():
        enc, dec = O.init_transform_state(O.NAMED_ARCHS[name], seed=0)
        assert sum(v.numel() for v in enc.values()) == enc_n
        assert sum(v.numel() for v in dec.values()) == dec_n


def test_image_io_casts():
    # Appendix C: true division by 255 on input, truncating cast on output
    x = torch.tensor([[[0.0, 254.9 / 255.0, 1.2, -0.3]]]).reshape(1, 1, 4).expand(3, 1, 4)
    u8 = O.to_uint8_hwc(x)
    assert u8[0, :, 0].tolist() == [0, 254, 255, 0]
    buf = np.arange(12, dtype=np.uint8).reshape(2, 2, 3)
    f = O.to_float_chw(buf)
    assert f.shape == (1, 3, 2, 2) and f[0, 1, 0, 1].item() == np.float32(4) / np.float32(255)


def test_oracle_ssim_and_delta_e_known_answers():
    """Hand-checkable corners of the restated evaluation metrics (test_cae.py:21-54): identical
    images, a constant offset (closed form), the Lab values of black / white / the sRGB primaries
    (published values)."""
    import numpy as np
    from oracle import cae_oracle as O
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, size=(40, 50, 3), dtype=np.uint8)
    assert abs(O.ssim_u8(x, x) - 1.0) < 1e-12
    assert O.delta_cielab_u8(x, x) == 0.0
    # constant images a, b: variances vanish, SSIM = (2ab + C1) / (a^2 + b^2 + C1)
    a, b = np.full((20, 20, 1), 100, np.uint8), np.full((20, 20, 1), 120, np.uint8)
    C1 = (0.01 * 255) ** 2
    assert abs(O.ssim_u8(a, b) - (2 * 100 * 120 + C1) / (100 ** 2 + 120 ** 2 + C1)) < 1e-12
    lab = O.rgb2lab_u8(np.array([[[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255]]], np.uint8))[0]
    want = np.array([[0, 0, 0], [100, 0, 0], [53.24, 80.09, 67.20], [87.73, -86.18, 83.18], [32.30, 79.19, -107.86]])
    assert np.abs(lab - want).max() < 0.05, lab


def test_oracle_ms_ssim_known_answers():
    """Identical images give 1; constant images a, b give the closed form: every contrast term is
    1 (the variances vanish) and the last scale contributes ((2ab + C1) / (a^2 + b^2 + C1))^0.1333
    (the 2x2 pooling keeps a constant image constant only away from the zero padding, so sizes
    that stay even through four halvings are used)."""
    import numpy as np
    from oracle import cae_oracle as O
    rng = np.random.default_rng(1)
    x = rng.integers(0, 256, size=(192, 208, 3), dtype=np.uint8)
    assert abs(O.ms_ssim_u8(x, x) - 1.0) < 1e-12
    a, b = np.full((192, 208, 1), 90, np.uint8), np.full((192, 208, 1), 140, np.uint8)
    C1 = (0.01 * 255) ** 2
    want = ((2 * 90 * 140 + C1) / (90 ** 2 + 140 ** 2 + C1)) ** 0.1333
    assert abs(O.ms_ssim_u8(a, b) - want) < 1e-9
