"""Generate the committed fixtures under ``tests/golden/``.
TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden

What is pinned by what:

* ``transforms_*.pt`` -- outputs of the REFERENCE's own ``Analyzer`` /
  ``Synthesizer`` classes (``_autoencoders.py:307-455`` imported through
  ``oracle/ref_loader.py``) on seeded inputs, with the reference-constructed
  random-init checkpoint stored beside them.  The oracle restatement and the
  CUDA path are both compared with these.
* ``named_*.pt`` -- the named architectures A / A_res / B are too large to
  commit their weights (3-13 MB), so the checkpoint comes from
  ``cae_oracle.make_checkpoint(arch, seed)`` (regenerated from the seed on any
  box; a sha256 of the weights is stored to catch generator drift) and the
  stored outputs are those of the REFERENCE classes loaded with that
  checkpoint.
* ``entropy_kat.pt`` -- EntropyBottleneck tables / likelihoods / rANS byte
  strings produced by the oracle restatement itself (CompressAI is absent:
  PARITY UNPINNED; these are regression vectors plus the analytic KAT of
  SURVEY.md 8c(iii)).
"""
import hashlib
import os
import sys

import numpy as np
import torch

from . import cae_oracle as O
from . import ref_loader as R

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          'tests', 'golden')

SMALL_ARCHS = {
    'M': O.NAMED_ARCHS['M'],
    'rgb_lrelu': dict(channels_org=3, channels_net=16, channels_bn=16, compression_level=3,
                      act_layer_type='LeakyReLU'),
    'rgb_lrelu_res': dict(channels_org=3, channels_net=16, channels_bn=16, compression_level=2,
                          act_layer_type='LeakyReLU', use_residual=True),
    'rgb_relu_bias': dict(channels_org=3, channels_net=16, channels_bn=8, compression_level=2,
                          act_layer_type='ReLU', bias=True),
    'rgb_noact': dict(channels_org=3, channels_net=16, channels_bn=8, compression_level=2),
    'rgb_gdn': dict(channels_org=3, channels_net=16, channels_bn=8, compression_level=3,
                    act_layer_type='GDN'),
    'rgb_bn_res': dict(channels_org=3, channels_net=16, channels_bn=8, compression_level=2,
                       act_layer_type='LeakyReLU', batch_norm=True, use_residual=True),
    'rgb_exp2': dict(channels_org=3, channels_net=8, channels_bn=8, compression_level=3,
                     channels_expansion=2, act_layer_type='LeakyReLU'),
    # groups=True builds every layer with groups=channels_in (R:68, 83, 119, 135, 153): torch
    # only accepts it when every layer's output channels are a multiple of its input channels
    'c4_groups': dict(channels_org=4, channels_net=4, channels_bn=4, compression_level=2,
                      act_layer_type='LeakyReLU', groups=True, bias=True),
    'c8_groups_res': dict(channels_org=8, channels_net=8, channels_bn=8, compression_level=2,
                          act_layer_type='LeakyReLU', groups=True, use_residual=True),
}

NAMED_INPUT = dict(A=(2, 64, 64), A_res=(1, 64, 64), B=(1, 64, 64))


def state_sha256(chk):
    h = hashlib.sha256()
    for part in ('encoder', 'decoder', 'fact_ent'):
        for k in sorted(chk[part]):
            if k == 'likelihood_lower_bound.bound':
                continue      # constant buffer added to the state dict in round 2; the committed
                              # fixtures' hashes cover the weights, which did not change
            h.update(k.encode())
            h.update(chk[part][k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def run_reference(chk, x_u8):
    """x_u8: N x C x H x W uint8.  Returns reference y and x_r[0] (eval)."""
    m = R.reference_model(chk)
    x = x_u8.float() / 255.0
    with torch.no_grad():
        y = m['encoder'](x)
        y_q, p_y = m['fact_ent'](y)
        x_r, _ = m['decoder'](y_q)
    return dict(y=y, y_q=y_q, p_y=p_y, x_r=x_r[0])


def main(only=None):
    """``only``: iterable of SMALL_ARCHS names to (re)generate; default = every fixture."""
    if not R.available():
        sys.exit('reference tree not found; fixtures can only be generated in the build container')
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)          # deterministic accumulation order

    for name, arch in SMALL_ARCHS.items():
        if only is not None and name not in only:
            continue
        chk = R.reference_checkpoint(arch, seed=1234)
        if arch.get('batch_norm'):
            g = torch.Generator().manual_seed(7)
            for part in ('encoder', 'decoder'):
                for k, v in chk[part].items():
                    if k.endswith('running_mean'):
                        v.copy_(torch.rand(v.shape, generator=g) * 0.2 - 0.1)
                    elif k.endswith('running_var'):
                        v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        c = arch['channels_org']
        x = O.synth_natural(2, c, 32, 48, seed=11)
        out = run_reference(chk, x)
        torch.save(dict(arch=arch, checkpoint=chk, x_u8=x, y=out['y'], y_q=out['y_q'],
                        p_y=out['p_y'], x_r=out['x_r']),
                   os.path.join(GOLDEN_DIR, f'transforms_{name}.pt'))
        print(name, 'y', tuple(out['y'].shape), 'x_r', tuple(out['x_r'].shape))

    if only is not None:
        return
    for name, (n, h, w) in NAMED_INPUT.items():
        arch = O.NAMED_ARCHS[name]
        chk = O.make_checkpoint(arch, seed=1234)
        x = O.synth_natural(n, 3, h, w, seed=12)
        out = run_reference(chk, x)
        torch.save(dict(arch=arch, seed=1234, sha256=state_sha256(chk), x_u8=x, y=out['y'],
                        y_q=out['y_q'], p_y=out['p_y'], x_r=out['x_r']),
                   os.path.join(GOLDEN_DIR, f'named_{name}.pt'))
        print(name, 'y', tuple(out['y'].shape), 'sha', state_sha256(chk)[:12])

    # Entropy model regression / known-answer vectors (oracle-generated: unpinned)
    torch.manual_seed(99)
    eb = O.EntropyBottleneck(6)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] -= torch.rand(6) * 3
        eb.quantiles[:, 0, 1] += torch.rand(6) - 0.5
        eb.quantiles[:, 0, 2] += torch.rand(6) * 5
        eb._factor1.add_(torch.randn_like(eb._factor1) * 0.3)
        eb._matrix2.add_(torch.randn_like(eb._matrix2) * 0.3)
    eb.update(force=True)
    eb.eval()
    g = torch.Generator().manual_seed(5)
    y = torch.randn(2, 6, 5, 7, generator=g) * 7
    y[0, 0, 0, 0] = 300.4          # escapes on both sides
    y[1, 5, 4, 6] = -1234.5
    y[0, 3, 2, 2] = 70000.0
    with torch.no_grad():
        y_q, p_y = eb(y)
    strings = eb.compress(y)
    torch.save(dict(state={k: v.detach().clone() for k, v in eb.state_dict().items()},
                    y=y, y_q=y_q, p_y=p_y, symbols=eb.symbols(y),
                    strings=[np.frombuffer(s, dtype=np.uint8).copy() for s in strings],
                    quantized_cdf=eb._quantized_cdf.clone(), cdf_length=eb._cdf_length.clone(),
                    offset=eb._offset.clone(), loss=eb.loss().detach()),
               os.path.join(GOLDEN_DIR, 'entropy_kat.pt'))
    print('entropy_kat', [len(s) for s in strings])


if __name__ == '__main__':
    main(only=sys.argv[1:] or None)
