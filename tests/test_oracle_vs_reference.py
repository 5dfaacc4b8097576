"""Oracle restatement vs. the reference's own classes.  Only runs where the
reference tree exists (the build container); the GPU box skips it."""
import pytest
import torch

from oracle import cae_oracle as O
from oracle import ref_loader as R
from oracle.make_golden import SMALL_ARCHS

pytestmark = pytest.mark.skipif(not R.available(), reason='reference tree not present')

ARCHS = dict(SMALL_ARCHS)
ARCHS['rgb_multiscale'] = dict(channels_org=3, channels_net=8, channels_bn=8, compression_level=3,
                               act_layer_type='LeakyReLU', multiscale_analysis=True)
ARCHS['rgb_gdn_res'] = dict(channels_org=3, channels_net=16, channels_bn=8, compression_level=2,
                            act_layer_type='GDN', use_residual=True)


@pytest.mark.parametrize('name', sorted(ARCHS))
def test_restatement_is_bit_exact(name):
    arch = ARCHS[name]
    chk = R.reference_checkpoint(arch, seed=77)
    ref = R.reference_model(chk)
    om = O.OracleModel(chk)
    x = O.synth_natural(2, arch['channels_org'], 48, 32, seed=3).float() / 255.0
    with torch.no_grad():
        y = ref['encoder'](x)
        y_q, _ = ref['fact_ent'](y)
        x_r, brg = ref['decoder'](y_q)
    out = om.forward(x)
    assert torch.equal(out['y'], y)
    for a, b in zip(out['x_r'], x_r):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.equal(a, b)
    for a, b in zip(out['fx_brg'], brg):
        assert torch.equal(a, b)


def test_reference_codec_runs_on_restated_entropy_model():
    """The reference's 'cae' codec class (R:530-584) executes unmodified on top of
    the restated EntropyBottleneck and agrees with the oracle's codec functions."""
    arch = dict(channels_org=3, channels_net=8, channels_bn=8, compression_level=2,
                act_layer_type='LeakyReLU')
    chk = R.reference_checkpoint(arch, seed=5)
    ref = R.load()
    codec = ref.ConvolutionalAutoencoder(checkpoint=chk, gpu=False)
    om = O.OracleModel(chk)
    tile = O.synth_tissue_tile(2, 3, ps=32, seed=2)
    b_ref = codec.encode(tile)
    assert b_ref == om.codec_encode(tile)
    assert (codec.decode(b_ref) == om.codec_decode(b_ref)).all()
