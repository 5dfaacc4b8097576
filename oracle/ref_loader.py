"""Import the reference's own ``Analyzer`` / ``Synthesizer`` / model factory.
TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

``/root/reference/src/models/tasks/_autoencoders.py`` is pure ``torch.nn``
except for three imports that are not installed here (``compressai``,
``numcodecs``; lines 10-16).  This loader registers stub modules for them in
``sys.modules`` -- the stubs carry the oracle's restated ``EntropyBottleneck``
/ ``GDN`` and a minimal ``Codec`` base -- and then executes the reference file
unmodified from where it lies.  Nothing is copied.  ``/root/reference`` only
exists in the build container, so everything here is used to *generate and
validate* fixtures (``oracle/make_golden.py``) and by tests that skip when the
tree is absent; nothing that runs on the GPU box may call it.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import cae_oracle as O

REFERENCE_ROOT = os.environ.get('CAE_REFERENCE_ROOT', '/root/reference')
_AE_PATH = os.path.join(REFERENCE_ROOT, 'src', 'models', 'tasks', '_autoencoders.py')

_cached = None


def available():
    return os.path.isfile(_AE_PATH)


class _GDN(nn.Module):
    """nn.Module shell over the oracle's restated GDN with
    CompressAI's parameter names ``beta`` / ``gamma``."""

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        beta, gamma = O.gdn_init_params(in_channels)
        self.beta = nn.Parameter(beta)
        self.gamma = nn.Parameter(gamma)

    def forward(self, x):
        return O.gdn_forward(x, self.beta, self.gamma, self.inverse)


class _Codec:
    codec_id = None


def _ndarray_copy(src, dst):
    if dst is None:
        return src
    np.copyto(np.asarray(dst).reshape(src.shape), src)
    return dst


def load():
    """Returns the reference module object (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(_AE_PATH)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    saved = {k: sys.modules.get(k) for k in (
        'compressai', 'compressai.ans', 'compressai.layers', 'compressai.entropy_models',
        'numcodecs', 'numcodecs.abc', 'numcodecs.compat')}
    try:
        ans = mod('compressai.ans')
        layers = mod('compressai.layers', GDN=_GDN)
        em = mod('compressai.entropy_models', EntropyBottleneck=O.EntropyBottleneck)
        mod('compressai', ans=ans, layers=layers, entropy_models=em)
        abc = mod('numcodecs.abc', Codec=_Codec)
        compat = mod('numcodecs.compat', ndarray_copy=_ndarray_copy,
                     ensure_contiguous_ndarray=np.ascontiguousarray)
        mod('numcodecs', abc=abc, compat=compat)
        spec = importlib.util.spec_from_file_location('_reference_autoencoders', _AE_PATH)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = m
    return m


def reference_checkpoint(arch, seed=1234):
    """Random-init checkpoint produced by the REFERENCE's constructors under
    ``torch.manual_seed(seed)`` (``setup_modules`` R:458-479; initialiser
    R:37-42), in the checkpoint-dict format ``autoencoder_from_state_dict``
    consumes (R:505-512)."""
    ref = load()
    torch.manual_seed(seed)
    model = ref.setup_modules(**arch)
    chk = dict(arch)
    for k in ('encoder', 'decoder', 'fact_ent'):
        chk[k] = {n: v.detach().clone() for n, v in model[k].state_dict().items()}
    return chk


def reference_model(checkpoint):
    """The reference's own model dict (R:505-527), eval mode, CPU."""
    ref = load()
    return ref.autoencoder_from_state_dict(checkpoint, gpu=False, train=False)
