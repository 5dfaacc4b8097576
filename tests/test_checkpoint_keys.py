"""Checkpoints produced with the real CompressAI load key for key.

The reference saves ``model['fact_ent'].module.state_dict()`` of a CompressAI
``EntropyBottleneck`` (``src/utils/_loggers.py:105-127``) and loads it back with a strict
``load_state_dict`` (``src/models/tasks/_autoencoders.py:490-502``).  The CompressAI 1.2.x key
set is the one ``scripts/transfer_weights.py:5-20`` of the reference enumerates:
``_matrix0..4, _bias0..4, _factor0..3, quantiles, target, _offset, _quantized_cdf,
_cdf_length, likelihood_lower_bound.bound``.
"""
import torch

import cnn_autoencoder_b200 as M
from cnn_autoencoder_b200._entropy import EntropyBottleneck
from oracle import cae_oracle as O

COMPRESSAI_1_2_KEYS = (['_matrix%d' % i for i in range(5)] + ['_bias%d' % i for i in range(5)] +
                       ['_factor%d' % i for i in range(4)] +
                       ['quantiles', 'target', '_offset', '_quantized_cdf', '_cdf_length',
                        'likelihood_lower_bound.bound'])


def test_state_dict_has_exactly_the_compressai_keys():
    eb = EntropyBottleneck(16, filters=[3, 3, 3, 3])
    assert sorted(eb.state_dict().keys()) == sorted(COMPRESSAI_1_2_KEYS)
    assert sorted(O.EntropyBottleneck(16).state_dict().keys()) == sorted(COMPRESSAI_1_2_KEYS)


def test_checkpoint_with_bound_buffer_loads_strictly():
    chk = O.make_checkpoint(O.NAMED_ARCHS['M'], seed=3)
    assert 'likelihood_lower_bound.bound' in chk['fact_ent']
    chk['fact_ent']['likelihood_lower_bound.bound'] = torch.tensor([2e-9])
    model = M.autoencoder_from_state_dict(chk, gpu=False, train=False)
    eb = model['fact_ent'].module
    assert abs(eb.likelihood_bound - 2e-9) < 1e-15          # the loaded buffer is the bound in force
    assert eb._quantized_cdf.numel() > 0                    # update(force=True) ran (R:502)


def test_checkpoint_without_bound_buffer_still_loads():
    chk = O.make_checkpoint(O.NAMED_ARCHS['M'], seed=3)
    chk['fact_ent'].pop('likelihood_lower_bound.bound')     # fixtures written in round 1
    model = M.autoencoder_from_state_dict(chk, gpu=False, train=False)
    assert abs(model['fact_ent'].module.likelihood_bound - 1e-9) < 1e-15


def test_parameter_list_keys_of_newer_compressai_are_mapped():
    src = EntropyBottleneck(8, filters=[3, 3, 3, 3])
    sd = {}
    for k, v in src.state_dict().items():
        for old, new in (('_matrix', 'matrices.'), ('_bias', 'biases.'), ('_factor', 'factors.')):
            if k.startswith(old):
                k = new + k[len(old):]
        sd[k] = v.clone()
    dst = EntropyBottleneck(8, filters=[3, 3, 3, 3])
    dst.load_state_dict(sd)
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
