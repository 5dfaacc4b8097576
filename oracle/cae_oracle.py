"""CPU restatement (torch fp32 / numpy) of the reference's compress/decompress
hot path.  TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Every function cites the reference lines it follows.  ``R:`` below abbreviates
``/root/reference/src/models/tasks/_autoencoders.py``.

The transforms are written functionally on top of a *state dict* with the
reference's key names (``analysis_track.<i>.model.<j>.weight`` ...), so the
same checkpoint drives the reference classes, this oracle and the CUDA path.

Pieces that live in CompressAI (EntropyBottleneck, GDN, LowerBound) follow
SURVEY.md Appendix A.  **Parity unpinned** for those (no upstream source or
package in this container); the transforms are pinned against the reference's
own classes (``oracle/ref_loader.py``, ``tests/golden``).
"""
import math
import struct

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import rans as _rans

# --------------------------------------------------------------------------
# Architecture description
# --------------------------------------------------------------------------

ARCH_DEFAULTS = dict(          # Analyzer/Synthesizer ctor defaults, R:308-318, R:365-376
    channels_org=3, channels_net=8, channels_bn=16, compression_level=3,
    channels_expansion=1, kernel_size=3, groups=False, batch_norm=False,
    dropout=0.0, bias=False, use_residual=False, act_layer_type=None,
    multiscale_analysis=False, K=4, r=3)

# Named architecture family, SURVEY.md section 8.
NAMED_ARCHS = {
    'A': dict(channels_org=3, channels_net=128, channels_bn=48, compression_level=3,
              act_layer_type='LeakyReLU'),
    'A_res': dict(channels_org=3, channels_net=128, channels_bn=48, compression_level=3,
                  act_layer_type='LeakyReLU', use_residual=True),
    'B': dict(channels_org=3, channels_net=128, channels_bn=192, compression_level=4,
              act_layer_type='LeakyReLU', use_residual=True),
    'M': dict(channels_org=1, channels_net=8, channels_bn=16, compression_level=3,
              act_layer_type='LeakyReLU'),
}


def full_arch(**kw):
    a = dict(ARCH_DEFAULTS)
    a.update(kw)
    if a['kernel_size'] != 3:
        raise ValueError('oracle restates kernel_size=3 only (the reference never sets another)')
    return a


def _act_has_pre_conv(act):
    # R:62, R:129, R:144, R:187, R:256, R:274
    return act is not None and act != 'GDN'


def analysis_plan(arch):
    """List of units; each unit is a dict(res=[ops], main=[ops], residual=bool).

    An op is ('conv', key, cin, cout, stride), ('bn', key, ch), ('act', kind, key, ch).
    Keys are state-dict prefixes relative to the Analyzer.  R:53-101 (plain),
    R:104-174 (residual), R:307-357 (unit chaining, last unit act None).
    """
    a = full_arch(**arch)
    L = a['compression_level']
    units = []
    cin, cout = a['channels_org'], a['channels_net']
    for i in range(L):
        last = i == L - 1
        act = None if last else a['act_layer_type']
        co = a['channels_bn'] if last else cout
        pre = f'analysis_track.{i}.'
        unit = dict(residual=a['use_residual'], res=[], main=[])
        if a['use_residual']:
            j = 0
            unit['res'].append(('conv', f'{pre}res_model.{j}', cin, cin, 1)); j += 1
            if a['batch_norm']:
                unit['res'].append(('bn', f'{pre}res_model.{j}', cin)); j += 1
            unit['res'].append(('act', act, f'{pre}res_model.{j}', cin)); j += 1
            if _act_has_pre_conv(act):
                unit['res'].append(('conv', f'{pre}res_model.{j}', cin, cin, 1)); j += 1
                if a['batch_norm']:
                    unit['res'].append(('bn', f'{pre}res_model.{j}', cin)); j += 1
            j = 0
            if _act_has_pre_conv(act):
                unit['main'].append(('act', act, f'{pre}model.{j}', cin)); j += 1
        else:
            j = 0
            if _act_has_pre_conv(act):
                unit['main'].append(('conv', f'{pre}model.{j}', cin, cin, 1)); j += 1
                if a['batch_norm']:
                    unit['main'].append(('bn', f'{pre}model.{j}', cin)); j += 1
                unit['main'].append(('act', act, f'{pre}model.{j}', cin)); j += 1
        unit['main'].append(('conv', f'{pre}model.{j}', cin, co, 2)); j += 1
        if a['batch_norm']:
            unit['main'].append(('bn', f'{pre}model.{j}', co)); j += 1
        if act is not None:
            unit['main'].append(('act', act, f'{pre}model.{j}', co)); j += 1
        units.append(unit)
        cin = co
        cout = cin * a['channels_expansion']
    return units


def synthesis_plan(arch):
    """Mirror of :func:`analysis_plan` for the Synthesizer.  R:177-227 (plain),
    R:230-304 (residual; note the activation *after* the 2nd residual conv,
    R:270-271), R:364-415 (chaining)."""
    a = full_arch(**arch)
    L = a['compression_level']
    units = []
    cin = a['channels_bn']
    cout = a['channels_net'] * a['channels_expansion'] ** L
    for i in range(L):
        last = i == L - 1
        act = None if last else a['act_layer_type']
        co = a['channels_org'] if last else cout
        pre = f'synthesis_track.{i}.'
        unit = dict(residual=a['use_residual'], res=[], main=[])
        if a['use_residual']:
            j = 0
            unit['res'].append(('convT', f'{pre}res_model.{j}', cin, cin, 1)); j += 1
            if a['batch_norm']:
                unit['res'].append(('bn', f'{pre}res_model.{j}', cin)); j += 1
            unit['res'].append(('act', act, f'{pre}res_model.{j}', cin)); j += 1
            if _act_has_pre_conv(act):
                unit['res'].append(('convT', f'{pre}res_model.{j}', cin, cin, 1)); j += 1
                if a['batch_norm']:
                    unit['res'].append(('bn', f'{pre}res_model.{j}', cin)); j += 1
                unit['res'].append(('act', act, f'{pre}res_model.{j}', cin)); j += 1
            j = 0
            if _act_has_pre_conv(act):
                unit['main'].append(('act', act, f'{pre}model.{j}', cin)); j += 1
        else:
            j = 0
            if _act_has_pre_conv(act):
                unit['main'].append(('convT', f'{pre}model.{j}', cin, cin, 1)); j += 1
                if a['batch_norm']:
                    unit['main'].append(('bn', f'{pre}model.{j}', cin)); j += 1
                unit['main'].append(('act', act, f'{pre}model.{j}', cin)); j += 1
        unit['main'].append(('convT', f'{pre}model.{j}', cin, co, 2)); j += 1
        if a['batch_norm']:
            unit['main'].append(('bn', f'{pre}model.{j}', co)); j += 1
        if act is not None:
            unit['main'].append(('act', act, f'{pre}model.{j}', co)); j += 1
        units.append(unit)
        cin = co
        cout = cin // a['channels_expansion']
    return units


# --------------------------------------------------------------------------
# GDN (CompressAI) -- parity unpinned
# --------------------------------------------------------------------------

GDN_PEDESTAL = (2.0 ** -18) ** 2
GDN_BETA_MIN = 1e-6


def gdn_init_params(ch):
    ped = torch.tensor([GDN_PEDESTAL])
    beta = torch.sqrt(torch.max(torch.ones(ch) + ped, ped))
    gamma = torch.sqrt(torch.max(0.1 * torch.eye(ch) + ped, ped))
    return beta, gamma


def gdn_forward(x, beta_p, gamma_p, inverse):
    ped = torch.tensor([GDN_PEDESTAL], dtype=x.dtype)
    beta_bound = (GDN_BETA_MIN + GDN_PEDESTAL) ** 0.5
    gamma_bound = (0.0 + GDN_PEDESTAL) ** 0.5
    beta = torch.max(beta_p, torch.tensor([beta_bound])) ** 2 - ped
    gamma = torch.max(gamma_p, torch.tensor([gamma_bound])) ** 2 - ped
    C = x.size(1)
    norm = F.conv2d(x ** 2, gamma.reshape(C, C, 1, 1), beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


# --------------------------------------------------------------------------
# Transforms
# --------------------------------------------------------------------------

def _apply_act(x, kind, sd, key, track):
    # R:19-34
    if kind is None or kind == 'Identity':
        return x
    if kind == 'LeakyReLU':
        return F.leaky_relu(x, 0.01)
    if kind == 'ReLU':
        return F.relu(x)
    if kind == 'GDN':
        return gdn_forward(x, sd[key + '.beta'], sd[key + '.gamma'], inverse=(track == 'synthesis'))
    raise ValueError(f'Activation layer {kind} not supported')


def _apply_op(x, op, sd, groups, track):
    kind = op[0]
    if kind == 'conv':
        _, key, cin, cout, stride = op
        w = sd[key + '.weight']
        b = sd.get(key + '.bias')
        # padding_mode='reflect', pad 1 (R:63-70, R:78-85); SURVEY Appendix C
        xp = F.pad(x, (1, 1, 1, 1), mode='reflect')
        return F.conv2d(xp, w, b, stride=stride, groups=cin if groups else 1)
    if kind == 'convT':
        _, key, cin, cout, stride = op
        w = sd[key + '.weight']
        b = sd.get(key + '.bias')
        # R:189-196 (s1, output_padding 0), R:204-211 (s2, output_padding 1)
        return F.conv_transpose2d(x, w, b, stride=stride, padding=1,
                                  output_padding=1 if stride == 2 else 0,
                                  groups=cin if groups else 1)
    if kind == 'bn':
        _, key, ch = op
        # eval-mode BatchNorm2d(affine=True), R:72-73
        return F.batch_norm(x, sd[key + '.running_mean'], sd[key + '.running_var'],
                            sd.get(key + '.weight'), sd.get(key + '.bias'), False, 0.0, 1e-5)
    if kind == 'act':
        _, act, key, ch = op
        return _apply_act(x, act, sd, key, track)
    raise ValueError(kind)


def _run_units(x, units, sd, groups, track):
    outs = []
    for unit in units:
        if unit['residual']:
            fx = x
            for op in unit['res']:
                fx = _apply_op(fx, op, sd, groups, track)
            x = fx + x                     # R:172, R:302
        for op in unit['main']:
            x = _apply_op(x, op, sd, groups, track)
        outs.append(x)
    return x, outs


def analysis_forward(sd, x, arch):
    """Analyzer.forward, R:359-361.  x fp32 N x C x H x W in [0,1] -> y."""
    a = full_arch(**arch)
    y, _ = _run_units(x, analysis_plan(a), sd, a['groups'], 'analysis')
    return y


def synthesis_forward(sd, y_q, arch):
    """Synthesizer.forward, R:442-455.  Returns (x_r list, fx_brg list); x_r[0]
    is the full-resolution reconstruction, lower scales are None unless
    multiscale_analysis (R:417-436)."""
    a = full_arch(**arch)
    units = synthesis_plan(a)
    _, fx_brg = _run_units(y_q, units, sd, a['groups'], 'synthesis')
    L = a['compression_level']
    x_r = []
    for i, fx in enumerate(fx_brg):
        if i == L - 1:
            x_r_i = fx                                            # nn.Identity, R:434
        elif a['multiscale_analysis']:
            key = f'color_layers.{i}.0'
            xp = F.pad(fx, (1, 1, 1, 1), mode='reflect')
            x_r_i = F.conv2d(xp, sd[key + '.weight'], sd.get(key + '.bias'),
                             groups=a['channels_org'] if a['groups'] else 1)
        else:
            x_r_i = None                                          # NoneColorLayer, R:45-50
        x_r.insert(0, x_r_i)
    return x_r, fx_brg


def init_transform_state(arch, seed):
    """Random-init weights with the reference's initialiser (R:37-42: Xavier
    uniform gain sqrt(2/1.01) on every Conv/ConvT weight, bias 0.01), keyed by
    a seed.  Does NOT reproduce the reference's RNG consumption order; both the
    oracle and the CUDA path load the returned dict, which is all parity needs.
    Returns (encoder_sd, decoder_sd)."""
    a = full_arch(**arch)
    g = torch.Generator().manual_seed(seed)
    gain = math.sqrt(2 / 1.01)

    def fill(units, transposed):
        sd = {}
        for unit in units:
            for op in unit['res'] + unit['main']:
                if op[0] in ('conv', 'convT'):
                    _, key, cin, cout, stride = op
                    grp = cin if a['groups'] else 1
                    shape = (cin, cout // grp, 3, 3) if transposed else (cout, cin // grp, 3, 3)
                    fan_in = shape[1] * 9
                    fan_out = shape[0] * 9
                    bound = gain * math.sqrt(6.0 / (fan_in + fan_out))
                    sd[key + '.weight'] = (torch.rand(shape, generator=g) * 2 - 1) * bound
                    if a['bias']:
                        sd[key + '.bias'] = torch.full((cout,), 0.01)
                elif op[0] == 'bn':
                    _, key, ch = op
                    sd[key + '.weight'] = torch.rand(ch, generator=g) + 0.5
                    sd[key + '.bias'] = torch.rand(ch, generator=g) - 0.5
                    sd[key + '.running_mean'] = torch.rand(ch, generator=g) * 0.2 - 0.1
                    sd[key + '.running_var'] = torch.rand(ch, generator=g) + 0.5
                elif op[0] == 'act' and op[1] == 'GDN':
                    _, _, key, ch = op
                    beta, gamma = gdn_init_params(ch)
                    sd[key + '.beta'] = beta
                    sd[key + '.gamma'] = gamma
        return sd

    enc_sd, dec_sd = fill(analysis_plan(a), False), fill(synthesis_plan(a), True)
    if a['multiscale_analysis']:
        # colour heads, R:417-429: Conv2d(channels_net * e^i -> channels_org), i reversed
        L, e = a['compression_level'], a['channels_expansion']
        for u, i in enumerate(reversed(range(L - 1))):
            cin, cout = a['channels_net'] * e ** i, a['channels_org']
            grp = cout if a['groups'] else 1
            shape = (cout, cin // grp, 3, 3)
            bound = gain * math.sqrt(6.0 / (shape[1] * 9 + shape[0] * 9))
            dec_sd[f'color_layers.{u}.0.weight'] = (torch.rand(shape, generator=g) * 2 - 1) * bound
            if a['bias']:
                dec_sd[f'color_layers.{u}.0.bias'] = torch.full((cout,), 0.01)
    return enc_sd, dec_sd


# --------------------------------------------------------------------------
# EntropyBottleneck (CompressAI 1.2.x) -- parity unpinned
# --------------------------------------------------------------------------

class _LowerBoundFn(torch.autograd.Function):
    """max(x, bound) with the CompressAI gradient rule (A.1): pass the gradient
    where x >= bound or where it pushes x up (grad < 0)."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        pass_through = (x >= bound) | (g < 0)
        return pass_through.type(g.dtype) * g, None


class _LowerBoundModule(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer('bound', torch.tensor([float(bound)]))


class EntropyBottleneck(nn.Module):
    """Restated factorized-prior entropy model.  Call sites in the reference:
    ctor R:476-477, R:607-608; forward ``_taskutils.py:97``; compress R:549-551,
    R:645-647; decompress R:568-572, R:662-665; update R:502, R:615; loss
    ``_lossutils.py:70``."""

    def __init__(self, channels, filters=(3, 3, 3, 3), init_scale=10.0, tail_mass=1e-9,
                 likelihood_bound=1e-9, entropy_coder_precision=16):
        super().__init__()
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        self.entropy_coder_precision = int(entropy_coder_precision)
        # CompressAI keeps the bound in a LowerBound submodule whose buffer is part of every
        # state dict (key ``likelihood_lower_bound.bound``; scripts/transfer_weights.py lists it)
        self.likelihood_lower_bound = _LowerBoundModule(likelihood_bound)

        F_ = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / F_[i + 1]))
            m = torch.empty(self.channels, F_[i + 1], F_[i]).fill_(float(init))
            self.register_parameter(f'_matrix{i:d}', nn.Parameter(m))
            b = torch.empty(self.channels, F_[i + 1], 1)
            nn.init.uniform_(b, -0.5, 0.5)
            self.register_parameter(f'_bias{i:d}', nn.Parameter(b))
            if i < len(self.filters):
                f = torch.zeros(self.channels, F_[i + 1], 1)
                self.register_parameter(f'_factor{i:d}', nn.Parameter(f))
        q = torch.tensor([-self.init_scale, 0.0, self.init_scale]).repeat(self.channels, 1, 1)
        self.quantiles = nn.Parameter(q)
        t = np.log(2 / self.tail_mass - 1)
        self.register_buffer('target', torch.tensor([-t, 0.0, t], dtype=torch.float32))
        self.register_buffer('_offset', torch.IntTensor())
        self.register_buffer('_quantized_cdf', torch.IntTensor())
        self.register_buffer('_cdf_length', torch.IntTensor())

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # fixtures written before the bound buffer was restated lack its key
        state_dict.setdefault(prefix + 'likelihood_lower_bound.bound',
                              self.likelihood_lower_bound.bound.clone())
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    # -- density model ------------------------------------------------------
    def _logits_cumulative(self, v, stop_gradient):
        for i in range(len(self.filters) + 1):
            m = getattr(self, f'_matrix{i:d}')
            b = getattr(self, f'_bias{i:d}')
            if stop_gradient:
                m, b = m.detach(), b.detach()
            v = torch.matmul(F.softplus(m), v)
            v = v + b
            if i < len(self.filters):
                f = getattr(self, f'_factor{i:d}')
                if stop_gradient:
                    f = f.detach()
                v = v + torch.tanh(f) * torch.tanh(v)
        return v

    def _likelihood(self, v):
        lower = self._logits_cumulative(v - 0.5, stop_gradient=False)
        upper = self._logits_cumulative(v + 0.5, stop_gradient=False)
        sign = -torch.sign(lower + upper).detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def _medians(self):
        return self.quantiles[:, :, 1:2]

    def forward(self, x, training=None):
        if training is None:
            training = self.training
        nd = x.dim()
        perm = list(range(nd))
        perm[0], perm[1] = 1, 0
        xp = x.permute(*perm).contiguous()
        shape = xp.size()
        v = xp.reshape(xp.size(0), 1, -1)
        if training:
            v = v + torch.empty_like(v).uniform_(-0.5, 0.5)
        else:
            med = self._medians()
            v = torch.round(v - med) + med
        lik = self._likelihood(v)
        lik = _LowerBoundFn.apply(lik, self.likelihood_lower_bound.bound)
        out = v.reshape(shape).permute(*perm).contiguous()
        lik = lik.reshape(shape).permute(*perm).contiguous()
        return out, lik

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    # -- tables -------------------------------------------------------------
    def update(self, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        med = self.quantiles[:, 0, 1]
        minima = torch.clamp(torch.ceil(med - self.quantiles[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - med).int(), min=0)
        self._offset = -minima
        pmf_start = med - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max().item())
        samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative(samples - 0.5, stop_gradient=True)
        upper = self._logits_cumulative(samples + 0.5, stop_gradient=True)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        cdf = torch.zeros((self.channels, max_length + 2), dtype=torch.int32)
        for c in range(self.channels):
            prob = torch.cat((pmf[c, :int(pmf_length[c])], tail[c]), dim=0).detach()
            q = _rans.pmf_to_quantized_cdf(prob.numpy().astype(np.float32),
                                           self.entropy_coder_precision)
            cdf[c, :q.shape[0]] = torch.from_numpy(q.astype(np.int32))
        self._quantized_cdf = cdf
        self._cdf_length = (pmf_length + 2).int()
        return True

    # -- entropy coding -----------------------------------------------------
    def symbols(self, x):
        """round(x - median_c) as int32, shape of x (A.1 ``compress``)."""
        med = self._medians().detach().reshape(1, -1, *([1] * (x.dim() - 2)))
        return torch.round(x - med).int()

    def compress(self, x):
        sym = self.symbols(x)
        N, C = sym.shape[:2]
        hw = int(np.prod(sym.shape[2:]))
        idx = np.repeat(np.arange(C, dtype=np.int32), hw)
        cdf = self._quantized_cdf.numpy()
        out = []
        for n in range(N):
            out.append(_rans.encode_with_indexes(
                sym[n].reshape(-1).numpy(), idx, cdf,
                self._cdf_length.numpy(), self._offset.numpy()))
        return out

    def decompress(self, strings, size):
        C = self._quantized_cdf.size(0)
        hw = int(np.prod(size))
        idx = np.repeat(np.arange(C, dtype=np.int32), hw)
        cdf = self._quantized_cdf.numpy()
        med = self._medians().detach().reshape(1, C, *([1] * len(size)))
        outs = []
        for s in strings:
            v = _rans.decode_with_indexes(s, idx, cdf, self._cdf_length.numpy(),
                                          self._offset.numpy())
            outs.append(torch.from_numpy(v).reshape(C, *size))
        y = torch.stack(outs).type_as(med) + med
        return y


def eb_load_state_dict(eb, fact_ent):
    """``load_state_dict`` for the fact_ent entry, R:490-502: resize the three
    table buffers by hand, load, then ``update(force=True)``."""
    for k in ('_quantized_cdf', '_offset', '_cdf_length'):
        if k in fact_ent:
            setattr(eb, k, fact_ent[k])
    eb.load_state_dict(fact_ent)
    eb.update(force=True)


# --------------------------------------------------------------------------
# Model dict, codecs, step, criterion, metrics
# --------------------------------------------------------------------------

def arch_from_checkpoint(chk):
    return {k: chk[k] for k in ARCH_DEFAULTS if k in chk}


def make_checkpoint(arch, seed=1234):
    """A checkpoint dict in the reference's format (``utils/_loggers.py:105-127``:
    architecture kwargs + 'encoder'/'decoder'/'fact_ent' state dicts in one flat
    dict) with random-init weights."""
    a = full_arch(**arch)
    enc, dec = init_transform_state(a, seed)
    torch.manual_seed(seed + 1)
    eb = EntropyBottleneck(a['channels_bn'], filters=[a['r']] * a['K'])
    chk = dict(arch)
    chk.update(encoder=enc, decoder=dec,
               fact_ent={k: v.detach().clone() for k, v in eb.state_dict().items()})
    return chk


class OracleModel:
    """enc -> fact_ent -> dec on CPU fp32, driven by a checkpoint dict
    (``autoencoder_from_state_dict`` R:505-527, eval mode)."""

    def __init__(self, checkpoint):
        self.arch = full_arch(**arch_from_checkpoint(checkpoint))
        self.enc_sd = {k: v.float() for k, v in checkpoint['encoder'].items()}
        self.dec_sd = {k: v.float() for k, v in checkpoint['decoder'].items()}
        self.fact_ent = EntropyBottleneck(self.arch['channels_bn'],
                                          filters=[self.arch['r']] * self.arch['K'])
        if checkpoint.get('fact_ent') is not None:
            eb_load_state_dict(self.fact_ent, checkpoint['fact_ent'])
        else:
            self.fact_ent.update(force=True)
        self.fact_ent.eval()

    @torch.no_grad()
    def encoder(self, x):
        return analysis_forward(self.enc_sd, x, self.arch)

    @torch.no_grad()
    def decoder(self, y_q):
        return synthesis_forward(self.dec_sd, y_q, self.arch)

    @torch.no_grad()
    def forward(self, x):
        """forward_func, ``_taskutils.py:95-108`` (eval)."""
        y = self.encoder(x)
        y_q, p_y = self.fact_ent(y)
        x_r, fx_brg = self.decoder(y_q)
        return dict(x_r=x_r, fx_brg=fx_brg, y=y, y_q=y_q, p_y=p_y)

    # 'cae' codec, R:539-584
    @torch.no_grad()
    def codec_encode(self, buf):
        h, w, c = buf.shape
        x = torch.from_numpy(np.ascontiguousarray(buf)).permute(2, 0, 1).reshape(1, c, h, w)
        x = x.float() / 255.0
        y = self.encoder(x)
        return struct.pack('>QQ', h, w) + self.fact_ent.compress(y)[0]

    @torch.no_grad()
    def codec_decode(self, buf):
        L = self.arch['compression_level']
        h, w = struct.unpack('>QQ', buf[:16])
        y_q = self.fact_ent.decompress([buf[16:]], size=(h // 2 ** L, w // 2 ** L))
        x_r, _ = self.decoder(y_q)
        return to_uint8_hwc(x_r[0][0])


def _codec_trace(self, buf):
    """Everything the parity gates of a throughput run compare for one uint8 HWC chunk: the
    oracle's latent relative to the medians, its integer symbols (C x h x w) and its uint8
    reconstruction (the 'cae' codec flow R:539-584 without the byte stream)."""
    with torch.no_grad():
        x = to_float_chw(buf)
        y = self.encoder(x)
        med = self.fact_ent._medians().detach().reshape(1, -1, 1, 1)
        sym = self.fact_ent.symbols(y)
        x_r, _ = self.decoder(sym.type_as(med) + med)
        return dict(y_minus_median=(y - med)[0], symbols=sym[0], x_r_u8=to_uint8_hwc(x_r[0][0]))


def _symbol_bits(self, sym):
    """-sum log2 p of the given integer symbols (C x h x w) under the oracle's density: the
    numerator of the estimated rate (``_ratedist.py:51-52``) for ANY symbol array, so that two
    implementations' symbols are priced by the same model."""
    with torch.no_grad():
        med = self.fact_ent._medians().detach().reshape(1, -1, 1, 1)
        y_q = sym.reshape(1, *sym.shape).type_as(med) + med
        _, p = self.fact_ent(y_q)            # eval: round(y_q - med) + med == y_q
        return float(-torch.log2(p).sum())


OracleModel.codec_trace = _codec_trace
OracleModel.symbol_bits = _symbol_bits


def to_float_chw(buf_u8_hwc):
    """u8 HWC -> fp32 1xCxHxW, true division by 255 (R:542-545, compress.py:53-55)."""
    h, w, c = buf_u8_hwc.shape
    x = torch.from_numpy(np.ascontiguousarray(buf_u8_hwc)).permute(2, 0, 1).reshape(1, c, h, w)
    return x.float() / 255.0


def to_uint8_hwc(x_chw):
    """fp32 CxHxW -> u8 HWC: x*255, clip(0,255), truncating cast (R:576-581,
    decompress.py:33-35)."""
    x = (x_chw * 255.0).clip(0, 255).to(torch.uint8)
    return np.ascontiguousarray(x.permute(1, 2, 0).numpy())


def rate_loss(x, p_y):
    """RateLoss, ``_ratedist.py:49-54``: -sum(log2 p_y) / (N*H*W) of the IMAGE."""
    return -torch.sum(torch.log2(p_y)) / (x.size(0) * x.size(2) * x.size(3))


def dist_mse(x, x_r0):
    """DistMSELoss * 255^2, ``_ratedist.py:57-63``, ``_lossutils.py:19,58``."""
    return F.mse_loss(x_r0, x) * 255 ** 2


def general_loss(x, out, fact_ent, distortion_lambda=0.01):
    """GeneralLoss.forward for criterion 'RateMSE', ``_lossutils.py:54-72,100-109``."""
    dist = dist_mse(x, out['x_r'][0])
    rate = rate_loss(x, out['p_y'])
    return dict(dist=[dist], dist_loss=distortion_lambda * dist, rate_loss=rate,
                entropy_loss=fact_ent.loss(), loss=distortion_lambda * dist + rate)


def psnr_u8(x, x_r, max_val=255):
    """compute_psnr, ``test_cae.py:60-63`` -- restated with a widening cast
    (the reference subtracts uint8 arrays, which wraps)."""
    d = x.astype(np.float64) - x_r.astype(np.float64)
    mse = float((d ** 2).mean())
    if mse == 0:
        return float('inf')
    return 20 * math.log10(max_val) - 10 * math.log10(mse)


def bpp(nbytes_stored, h, w):
    """compute_rate, ``test_cae.py:71-73``."""
    return 8.0 * float(nbytes_stored) / (h * w)


# --------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------

def synth_natural(n, c, h, w, seed=1):
    """uint8 natural-image-like batch: three octaves of bicubic-upsampled
    uniform noise (config 2)."""
    g = torch.Generator().manual_seed(seed)
    acc = torch.zeros(n, c, h, w)
    for octave, amp in ((8, 0.5), (32, 0.3), (128, 0.2)):
        hh, ww = max(2, h // octave), max(2, w // octave)
        z = torch.rand(n, c, hh, ww, generator=g)
        acc += amp * F.interpolate(z, size=(h, w), mode='bicubic', align_corners=False)
    return (acc.clamp(0, 1) * 255).round().to(torch.uint8)


def synth_tissue_tile(ty, tx, ps=512, seed=2):
    """uint8 HWC tissue-like tile keyed on (tile_y, tile_x, seed): pink/purple
    base, low-frequency blobs, mid-frequency texture, 15 % white noise
    (configs 3-4)."""
    g = torch.Generator().manual_seed((seed * 1000003 + ty) * 1000003 + tx)
    base = torch.tensor([0.85, 0.55, 0.75]).view(1, 3, 1, 1)
    lo = F.interpolate(torch.rand(1, 1, max(2, ps // 64), max(2, ps // 64), generator=g),
                       size=(ps, ps), mode='bicubic', align_corners=False)
    mid = F.interpolate(torch.rand(1, 3, max(2, ps // 8), max(2, ps // 8), generator=g),
                        size=(ps, ps), mode='bilinear', align_corners=False)
    noise = torch.rand(1, 3, ps, ps, generator=g)
    img = base * (0.6 + 0.4 * lo) + 0.25 * (mid - 0.5) + 0.15 * (noise - 0.5)
    img = (img.clamp(0, 1) * 255).round().to(torch.uint8)
    return np.ascontiguousarray(img[0].permute(1, 2, 0).numpy())


def ssim_u8(x, x_r):
    """compute_ssim, ``test_cae.py:52-54``: ``skimage.metrics.structural_similarity(x, x_r,
    channel_axis=2)`` with its defaults, restated from the published algorithm (scikit-image is
    not installed here -- parity unpinned by the real package): per channel, float64, 7x7
    ``scipy.ndimage.uniform_filter`` moments, sample covariance (NP / (NP - 1)), K1 = 0.01,
    K2 = 0.03, data range 255 for uint8, the map cropped by (win - 1) // 2 = 3 pixels, mean; then
    the mean over the channels."""
    from scipy.ndimage import uniform_filter
    win, K1, K2, R = 7, 0.01, 0.03, 255.0
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    pad = (win - 1) // 2
    vals = []
    for c in range(x.shape[2]):
        a, b = x[..., c].astype(np.float64), x_r[..., c].astype(np.float64)
        ux, uy = uniform_filter(a, size=win), uniform_filter(b, size=win)
        uxx, uyy, uxy = uniform_filter(a * a, size=win), uniform_filter(b * b, size=win), uniform_filter(a * b, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        vals.append(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
    return float(np.mean(vals))


def rgb2lab_u8(x):
    """``skimage.color.rgb2lab`` of an 8-bit sRGB image (D65, 2 degree observer), restated from the
    published algorithm (parity unpinned by the real package)."""
    v = x.astype(np.float64) / 255.0
    lin = np.where(v > 0.04045, ((v + 0.055) / 1.055) ** 2.4, v / 12.92)
    M = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169],
                  [0.019334, 0.119193, 0.950227]])
    xyz = lin @ M.T
    xyz = xyz / np.array([0.95047, 1.0, 1.08883])
    f = np.where(xyz > 0.008856, np.cbrt(xyz), 7.787 * xyz + 16.0 / 116.0)
    L = 116.0 * f[..., 1] - 16.0
    a = 500.0 * (f[..., 0] - f[..., 1])
    b = 200.0 * (f[..., 1] - f[..., 2])
    return np.stack([L, a, b], axis=-1)


def delta_cielab_u8(x, x_r):
    """compute_deltaCIELAB, ``test_cae.py:21-44``: mean ``deltaE_cie76`` of the two images."""
    d = rgb2lab_u8(x) - rgb2lab_u8(x_r)
    return float(np.sqrt((d ** 2).sum(axis=-1)).mean())


def ms_ssim_u8(x, x_r):
    """compute_ms_ssim, ``test_cae.py:46-50``: ``pytorch_msssim.ms_ssim(x_r, x, data_range=255)``
    with its defaults, restated from the published algorithm (the package is not installed here
    -- parity unpinned by it): 11-tap Gaussian window (sigma 1.5, normalised), valid separable
    filtering of X, Y, X^2, Y^2, XY, per-channel means of the contrast-structure map on the first
    four scales and of the SSIM map on the fifth (all through relu), ``avg_pool2d(2, padding =
    size % 2)`` between scales, weighted product with (0.0448, 0.2856, 0.3001, 0.2363, 0.1333),
    mean over channels.  float64 throughout."""
    X = torch.from_numpy(np.moveaxis(x_r, -1, 0)[None].copy()).double()
    Y = torch.from_numpy(np.moveaxis(x, -1, 0)[None].copy()).double()
    coords = torch.arange(11, dtype=torch.float64) - 11 // 2
    g = torch.exp(-(coords ** 2) / (2 * 1.5 ** 2))
    g = g / g.sum()
    C = X.shape[1]

    def gauss(t):
        t = F.conv2d(t, g.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)
        return F.conv2d(t, g.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)

    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=torch.float64)
    mcs = []
    for i in range(5):
        mu1, mu2 = gauss(X), gauss(Y)
        s1, s2, s12 = gauss(X * X) - mu1 * mu1, gauss(Y * Y) - mu2 * mu2, gauss(X * Y) - mu1 * mu2
        cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
        ssim_map = ((2 * mu1 * mu2 + C1) / (mu1 ** 2 + mu2 ** 2 + C1)) * cs_map
        ssim_c, cs_c = ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)
        if i < 4:
            mcs.append(torch.relu(cs_c))
            pad = [s % 2 for s in X.shape[2:]]
            X, Y = F.avg_pool2d(X, 2, padding=pad), F.avg_pool2d(Y, 2, padding=pad)
    vals = torch.stack(mcs + [torch.relu(ssim_c)], dim=0)
    return float(torch.prod(vals ** weights.view(-1, 1, 1), dim=0).mean())
