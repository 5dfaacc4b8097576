"""Host-side mirror of the reference's ``models.tasks._autoencoders`` for the
compress/decompress hot path: same class names, constructor arguments,
state-dict keys and call conventions, with the arithmetic executed by the
sm_100a kernels behind ``include/cae_b200.h``.

Reference (R = ``/root/reference/src/models/tasks/_autoencoders.py``):
``DownsamplingUnit`` R:53-101, ``ResidualDownsamplingUnit`` R:104-174,
``UpsamplingUnit`` R:177-227, ``ResidualUpsamplingUnit`` R:230-304,
``Analyzer`` R:307-361, ``Synthesizer`` R:364-455, ``setup_modules`` R:458-479,
``load_state_dict`` R:482-502, ``autoencoder_from_state_dict`` R:505-527,
codecs ``'cae'`` R:530-584 and ``'cae_bn'`` R:587-673.

The unit classes only *hold parameters* (``nn.Conv2d`` / ``nn.ConvTranspose2d``
instances at the reference's Sequential indices, so checkpoints load unchanged)
and describe their dataflow as a ``layout`` list; ``_engine`` turns a track of
layouts into fused kernel launches.  In ``eval()`` mode ``forward`` runs only
CUDA kernels of this repo and requires CUDA tensors (there is no CPU fallback).
In ``train()`` mode ``forward`` is the same dataflow written with torch autograd
ops on the device (interim: the backward kernels are scheduled).
"""
import base64
import io
import math
import os
import struct
import threading

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi as C
from . import _engine as E
from . import _ops as O
from . import _train_conv as T
from ._entropy import EntropyBottleneck

try:                                    # zarr / numcodecs are optional at import time
    from numcodecs.abc import Codec
    from numcodecs.compat import ensure_contiguous_ndarray, ndarray_copy
except ImportError:                     # same minimal surface, so the codecs still work stand-alone
    class Codec:                        # noqa: D401
        codec_id = None

        def get_config(self):
            cfg = {'id': self.codec_id}
            cfg.update({k: v for k, v in self.__dict__.items() if not k.startswith('_')})
            return cfg

        @classmethod
        def from_config(cls, config):
            config = dict(config)
            config.pop('id', None)
            return cls(**config)

    def ensure_contiguous_ndarray(buf):
        return np.ascontiguousarray(buf)

    def ndarray_copy(src, dst):
        if dst is None:
            return src
        dst = np.asarray(dst)
        np.copyto(dst.reshape(src.shape).view(src.dtype) if dst.dtype != src.dtype else
                  dst.reshape(src.shape), src)
        return dst


_SUPPORTED_ACTS = (None, 'Identity', 'LeakyReLU', 'ReLU', 'GDN')


def _make_act(kind, channels, track):
    """R:19-34.  The modules are placeholders for state-dict index parity; the
    activation itself is an epilogue flag of the preceding convolution."""
    if kind is None or kind == 'Identity':
        return nn.Identity()
    if kind == 'LeakyReLU':
        return nn.LeakyReLU(inplace=False)
    if kind == 'ReLU':
        return nn.ReLU(inplace=False)
    if kind == 'GDN':
        return GDN(in_channels=channels, inverse=track == 'synthesis')
    raise ValueError(f'Activation layer {kind} not supported')


class _LowerBoundBuffer(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer('bound', torch.tensor([float(bound)]))


class _NonNegativeParam(nn.Module):
    """State-dict shell of CompressAI's NonNegativeParametrizer (buffers ``pedestal`` and
    ``lower_bound.bound``) so GDN checkpoints load key for key."""

    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        pedestal = float(reparam_offset) ** 2
        self.register_buffer('pedestal', torch.tensor([pedestal]))
        self.lower_bound = _LowerBoundBuffer((float(minimum) + pedestal) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return torch.max(x, self.lower_bound.bound) ** 2 - self.pedestal


class GDN(nn.Module):
    """Generalised divisive normalisation with CompressAI's parameters (``beta``, ``gamma`` and
    the re-parametrisation buffers), the ``act_layer_type='GDN'`` choice of
    R:29-30.  Eval-mode tracks run it as the ``cae_gdn`` kernel; this ``forward`` is the torch
    autograd form used in ``train()`` mode."""

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = _NonNegativeParam(minimum=float(beta_min))
        self.gamma_reparam = _NonNegativeParam()
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(int(in_channels))))
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(int(in_channels))))

    def effective(self):
        """(beta, gamma) after re-parametrisation, fp32, as the kernel consumes them."""
        return (self.beta_reparam(self.beta).detach().float().contiguous(),
                self.gamma_reparam(self.gamma).detach().float().contiguous())

    def forward(self, x):
        C = x.size(1)
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma).reshape(C, C, 1, 1)
        norm = F.conv2d(x ** 2, gamma, beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm


def initialize_weights(m):
    """R:37-42."""
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
        nn.init.xavier_uniform_(m.weight.data, gain=math.sqrt(2 / 1.01))
        if m.bias is not None:
            nn.init.constant_(m.bias.data, 0.01)


class NoneColorLayer(nn.Module):
    def forward(self, *args, **kwargs):
        return None


class _Unit(nn.Module):
    """Parameter holder + dataflow description shared by the four unit classes."""

    transposed = False
    residual = False

    def __init__(self, channels_in, channels_out, kernel_size=3, groups=False, batch_norm=False,
                 dropout=0.0, bias=False, act_layer_type=None):
        super().__init__()
        if kernel_size != 3:
            raise NotImplementedError('the CUDA kernels implement kernel_size=3 '
                                      '(the only value the reference uses)')
        if act_layer_type not in _SUPPORTED_ACTS:
            raise ValueError(f'Activation layer {act_layer_type} not supported')
        track = 'synthesis' if self.transposed else 'analysis'
        act = act_layer_type
        pre_conv = act is not None and act != 'GDN'
        seqs = {'model': [], 'res_model': []}
        layout = []

        def conv(seq, cin, cout, stride):
            if self.transposed:
                # groups=True: every layer is built with groups=channels_in (R:68, 83, 119, 135,
                # 153, 195, 210, 247); torch refuses channel counts the groups do not divide,
                # exactly as in the reference
                m = nn.ConvTranspose2d(cin, cout, 3, stride=stride, padding=1,
                                       output_padding=stride - 1, bias=bias,
                                       groups=cin if groups else 1)
            else:
                m = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=bias,
                              padding_mode='reflect', groups=cin if groups else 1)
            layout.append(('conv', seq, len(seqs[seq])))
            seqs[seq].append(m)
            if batch_norm:
                layout.append(('bn', seq, len(seqs[seq])))
                seqs[seq].append(nn.BatchNorm2d(cout, affine=True))

        def activation(seq, kind, ch):
            layout.append(('act', seq, len(seqs[seq]), kind))
            seqs[seq].append(_make_act(kind, ch, track))

        if self.residual:
            conv('res_model', channels_in, channels_in, 1)
            activation('res_model', act, channels_in)
            if pre_conv:
                conv('res_model', channels_in, channels_in, 1)
                if self.transposed:                      # R:270-271: decoder only
                    activation('res_model', act, channels_in)
            layout.append(('add',))
            if pre_conv:
                activation('model', act, channels_in)
        elif pre_conv:
            conv('model', channels_in, channels_in, 1)
            activation('model', act, channels_in)
        conv('model', channels_in, channels_out, 2)
        if act is not None:
            activation('model', act, channels_out)
        if dropout > 0.0:
            seqs['model'].append(nn.Dropout2d(dropout))
        if self.residual:
            self.res_model = nn.Sequential(*seqs['res_model'])
        self.model = nn.Sequential(*seqs['model'])
        self.layout = layout

    def forward(self, x):
        """torch-op dataflow (training / autograd path)."""
        if self.residual:
            x = self.res_model(x) + x
        return self.model(x)


class DownsamplingUnit(_Unit):
    pass


class ResidualDownsamplingUnit(_Unit):
    residual = True


class UpsamplingUnit(_Unit):
    transposed = True


class ResidualUpsamplingUnit(_Unit):
    transposed = True
    residual = True


class _Track(nn.Module):
    """Common execution plumbing of Analyzer / Synthesizer."""

    def __init__(self):
        super().__init__()
        self._exec = None
        self._lock = threading.Lock()
        # train() mode: run the wide layers on the CUDA kernels (CAE_TRAIN_TORCH=1: torch ops)
        self.train_kernels = not os.environ.get('CAE_TRAIN_TORCH')

    def _units(self):
        raise NotImplementedError

    def _executor(self):
        if self._exec is None:
            self._exec = E.TrackExecutor(E.steps_from_units(self._units()))
        return self._exec

    def _check_input(self, x):
        if not x.is_cuda:
            raise C.CaeError(f'{type(self).__name__}: eval-mode forward runs CUDA kernels only; '
                             'got a CPU tensor (no CPU fallback). Move the model and data to cuda.')
        p = next(self.parameters())
        if p.device != x.device:
            raise C.CaeError(f'{type(self).__name__}: parameters on {p.device}, input on {x.device}')


class Analyzer(_Track):
    def __init__(self, channels_org=3, channels_net=8, channels_bn=16, compression_level=3,
                 channels_expansion=1, kernel_size=3, groups=False, batch_norm=False, dropout=0.0,
                 bias=False, use_residual=False, act_layer_type=None, **kwargs):
        super().__init__()
        unit = ResidualDownsamplingUnit if use_residual else DownsamplingUnit
        common = dict(kernel_size=kernel_size, groups=groups, batch_norm=batch_norm,
                      dropout=dropout, bias=bias)
        track = []
        cin, cout = channels_org, channels_net
        for _ in range(compression_level - 1):
            track.append(unit(cin, cout, act_layer_type=act_layer_type, **common))
            cin, cout = cout, cout * channels_expansion
        if compression_level > 0:
            track.append(unit(cin, channels_bn, act_layer_type=None, **common))
        else:
            track.append(nn.Identity())
        self.analysis_track = nn.Sequential(*track)
        self.apply(initialize_weights)

    def _units(self):
        return [u for u in self.analysis_track if isinstance(u, _Unit)]

    def forward(self, x, quant=None):
        """x: fp32 N x C x H x W in [0,1] (or uint8 N x H x W x C, divided by 255 in the
        first kernel) -> latent y, fp32 N x C_bn x H/2^L x W/2^L.  ``quant`` (eval only): an
        ``EntropyBottleneck.quant_request()`` filled by the last layer's epilogue."""
        if self.training:
            if x.dtype == torch.uint8:
                x = x.permute(0, 3, 1, 2).float() / 255.0
            if self.train_kernels and x.is_cuda:
                # wide layers forward / backward on this repo's kernels (_train_conv.py); tracks
                # they do not cover run the torch formulation below
                r = T.run_track(self, x)
                if r is not None:
                    return r[0]
            return self.analysis_track(x)
        self._check_input(x)
        if not self._units():
            return x
        with self._lock, torch.no_grad():
            a = O.wrap_u8_hwc(x) if x.dtype == torch.uint8 else O.wrap_nchw(x)
            if a.c > 4 and a.fmt == C.FMT_F32_NCHW and self._executor().steps[0].groups == 1:
                # wide dense first layer: tensor-core path, input rounded to fp16 here (a grouped
                # first layer reads the fp32 image itself on the direct kernel)
                first = self._executor().steps[0]
                fmt = C.FMT_F16_SPLIT if first.kind == C.CONV_S2 else C.FMT_F16_PLANAR
                a = O.nchw_to_planar(a.t, fmt, C.HALO_REFLECT)
            y, _, _ = self._executor().run(a, C.FMT_F32_NCHW, quant=quant)
        return y.t


class Synthesizer(_Track):
    def __init__(self, channels_org=3, channels_net=8, channels_bn=16, compression_level=3,
                 channels_expansion=1, kernel_size=3, groups=False, batch_norm=False, dropout=0.0,
                 bias=False, use_residual=False, act_layer_type=None, multiscale_analysis=False,
                 **kwargs):
        super().__init__()
        unit = ResidualUpsamplingUnit if use_residual else UpsamplingUnit
        common = dict(kernel_size=kernel_size, groups=groups, batch_norm=batch_norm,
                      dropout=dropout, bias=bias)
        track = []
        cin = channels_bn
        cout = channels_net * channels_expansion ** compression_level
        for _ in range(compression_level - 1):
            track.append(unit(cin, cout, act_layer_type=act_layer_type, **common))
            cin, cout = cout, cout // channels_expansion
        if compression_level > 0:
            track.append(unit(cin, channels_org, act_layer_type=None, **common))
        else:
            track.append(nn.Identity())
        self.synthesis_track = nn.Sequential(*track)
        if multiscale_analysis:
            # R:417-429: a reflect-padded 3x3 colour head on every intermediate scale
            layers = [nn.Sequential(nn.Conv2d(channels_net * channels_expansion ** i, channels_org,
                                              kernel_size=kernel_size, stride=1,
                                              padding=kernel_size // 2, bias=bias,
                                              groups=channels_org if groups else 1,
                                              padding_mode='reflect'))
                      for i in reversed(range(compression_level - 1))]
        else:
            layers = [nn.Sequential(NoneColorLayer()) for _ in range(compression_level - 1)]
        layers.append(nn.Identity())
        self.multiscale = bool(multiscale_analysis)
        self.color_layers = nn.ModuleList(layers)
        self.rec_level = compression_level
        self.bridges = False       # set True to get fx_brg as fp32 tensors in eval mode
        self.apply(initialize_weights)

    def _units(self):
        return [u for u in self.synthesis_track if isinstance(u, _Unit)]

    def _unit_output_indices(self):
        idx, k = [], 0
        for u in self._units():
            k += sum(1 for op in u.layout if op[0] == 'conv')
            idx.append(k)
        return idx

    def forward(self, x, as_uint8=False, planar=None):
        """y_q fp32 N x C_bn x h x w -> (x_r, fx_brg) exactly as R:442-455 (``planar``: the same
        y_q already in the track's input layout, from the fused quantizer): x_r[0] is the
        full-resolution reconstruction, lower scales are None.  ``as_uint8`` (extension
        used by the codecs) additionally returns the N x H x W x C uint8 image produced
        by the last kernel's epilogue: ``(x_r, fx_brg, u8)``; ``as_uint8='only'`` skips the
        fp32 copy (``x_r`` / ``fx_brg`` entries are then None)."""
        if self.training:
            if self.train_kernels and x.is_cuda and not self.multiscale and not self.bridges:
                r = T.run_track(self, x)
                if r is not None:
                    final, tensors = r
                    idx = self._unit_output_indices()
                    # bridges inside the fused run are not materialised (as in eval mode)
                    fx_brg = [tensors.get(i) for i in idx]
                    x_r = [None] * len(idx)
                    x_r[0] = final
                    return x_r, fx_brg
            fx, fx_brg, x_r = x, [], []
            for up, col in zip(self.synthesis_track, self.color_layers):
                fx = up(fx)
                x_r.insert(0, col(fx))
                fx_brg.append(fx)
            return x_r, fx_brg
        self._check_input(x)
        if len(self._units()) == 0:
            return [x], [x]
        with self._lock, torch.no_grad():
            if planar is not None and x.shape[1] > 4:
                a = planar
            elif x.shape[1] > 4:
                # the planar copy of the latent lives in a buffer kept per shape (the zero halo
                # is written once at allocation)
                key = (tuple(x.shape), str(x.device))
                buf = self._in_planar if getattr(self, '_in_planar_key', None) == key else None
                a = O.nchw_to_planar(x, C.FMT_F16_PLANAR, C.HALO_KEEP, out=buf)
                self._in_planar, self._in_planar_key = a, key
            else:
                a = O.wrap_nchw(x)
            return self._run(a, as_uint8)

    def forward_planar(self, a, as_uint8=False):
        """``forward`` on a latent that is already in the track's input layout (an ``_ops.Act``
        in planar fp16 with a zero halo, e.g. written by ``cae_eb_dequantize_planar``)."""
        with self._lock, torch.no_grad():
            return self._run(a, as_uint8)

    def _run(self, a, as_uint8):
        n_units = len(self._units())
        if True:
            outs = self._unit_output_indices()
            keep = outs[:-1] if (self.bridges or self.multiscale) else ()
            final = C.FMT_U8_HWC if as_uint8 else C.FMT_F32_NCHW
            only_u8 = as_uint8 == 'only'
            last, kept, aux = self._executor().run(a, final, keep=keep, aux_last=not only_u8)
            if only_u8 and last.fmt == C.FMT_U8_HWC:
                return [None] * n_units, [None] * n_units, last.t
            x_full = aux if aux is not None else last.t
            u8 = last.t if (as_uint8 and last.fmt == C.FMT_U8_HWC) else None
            fx_brg = [O.planar_to_nchw(kept[i]) if (self.bridges and i in kept) else None
                      for i in outs[:-1]]
            fx_brg.append(x_full)
            x_r = [None] * (n_units - 1)
            if self.multiscale:
                # colour heads read the kept intermediate tensors (zero halo: their consumer is
                # a transposed conv), so they run on the direct kernel, which resolves the
                # reflect padding by index arithmetic on the interior
                for u, i in enumerate(outs[:-1]):
                    conv = self.color_layers[u][0]
                    a_in = kept[i]
                    out = O.alloc_act(C.FMT_F32_NCHW, a_in.n, conv.out_channels, a_in.h, a_in.w,
                                      device=a_in.t.device)
                    O.conv(C.CONV_S1, a_in, conv.weight.detach().float().contiguous(),
                           conv.out_channels, out, igemm=False,
                           bias=conv.bias.detach().float().contiguous() if conv.bias is not None else None,
                           pad_mode=C.PAD_REFLECT, groups=conv.groups)
                    x_r[n_units - 2 - u] = out.t
        x_r.insert(0, x_full)
        if as_uint8:
            if u8 is None:
                u8 = (x_full * 255.0).clip(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
            return x_r, fx_brg, u8
        return x_r, fx_brg


# --------------------------------------------------------------------------
# Model dict factory (R:458-527)
# --------------------------------------------------------------------------

def setup_modules(channels_bn=192, compression_level=4, K=4, r=3, enabled_modules=None, **kwargs):
    if enabled_modules is None:
        enabled_modules = ['encoder', 'decoder', 'fact_ent']
    model = {}
    if 'encoder' in enabled_modules:
        model['encoder'] = Analyzer(channels_bn=channels_bn, compression_level=compression_level,
                                    **kwargs)
    if 'decoder' in enabled_modules:
        model['decoder'] = Synthesizer(channels_bn=channels_bn,
                                       compression_level=compression_level, **kwargs)
    if 'fact_ent' in enabled_modules:
        model['fact_ent'] = EntropyBottleneck(channels=channels_bn, filters=[r] * K)
    return model


def load_state_dict(model, encoder=None, decoder=None, fact_ent=None, **kwargs):
    if 'encoder' in model and encoder is not None:
        model['encoder'].load_state_dict(encoder, strict=False)
    if 'decoder' in model and decoder is not None:
        model['decoder'].load_state_dict(decoder, strict=False)
    if 'fact_ent' in model and fact_ent is not None:
        for k in ('_quantized_cdf', '_offset', '_cdf_length'):
            if k in fact_ent:
                setattr(model['fact_ent'], k, fact_ent[k])
        model['fact_ent'].load_state_dict(fact_ent)
        model['fact_ent'].update(force=True)


class ModuleHandle(nn.Module):
    """The ``nn.DataParallel`` slot of the reference's model dict (R:514-520):
    callable, ``.module`` reach-through, ``train/eval/cuda/state_dict``.  One
    process drives one GPU here (tiles are sharded across processes by chunk
    range), so the handle simply forwards to the module."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def autoencoder_from_state_dict(checkpoint, gpu=False, train=False):
    """R:505-527.  ``gpu=False`` builds the modules on the CPU exactly like the
    reference; their eval-mode ``forward`` then refuses to run (no CPU fallback)."""
    if isinstance(checkpoint, str):
        state = torch.load(checkpoint, map_location='cpu', weights_only=False)
    else:
        state = checkpoint
    model = setup_modules(**state)
    load_state_dict(model, **state)
    for k in model:
        model[k] = ModuleHandle(model[k])
        if gpu and torch.cuda.is_available():
            model[k].cuda()
        model[k].train() if train else model[k].eval()
    return model


# --------------------------------------------------------------------------
# numcodecs codecs (R:530-673)
# --------------------------------------------------------------------------

def _device_of(handle):
    return next(handle.parameters()).device


class ConvolutionalAutoencoder(Codec):
    """zarr chunk codec ``'cae'``: uint8 HWC tile <-> 16-byte '>QQ' (h, w) header + rANS
    stream (R:530-584).  ``gpu`` is kept in the config for compatibility; the codec
    always executes on the current CUDA device."""
    codec_id = 'cae'

    def __init__(self, checkpoint, gpu=False):
        self.checkpoint = checkpoint
        self.gpu = gpu
        self._model = autoencoder_from_state_dict(checkpoint, gpu=True, train=False)
        if not torch.cuda.is_available():
            raise C.CaeError("codec 'cae' needs a CUDA device (no CPU fallback)")

    def encode(self, buf):
        buf = np.ascontiguousarray(buf)
        h, w, c = buf.shape
        dev = _device_of(self._model['encoder'])
        x = torch.from_numpy(buf).to(dev, non_blocking=True).reshape(1, h, w, c)
        y = self._model['encoder'](x)
        strings = self._model['fact_ent'].module.compress(y)
        return struct.pack('>QQ', h, w) + strings[0]

    def decode(self, buf, out=None):
        if out is not None:
            out = ensure_contiguous_ndarray(out)
        buf = bytes(buf)
        level = len(self._model['decoder'].module.synthesis_track)
        h, w = struct.unpack('>QQ', buf[:16])
        y_q = self._model['fact_ent'].module.decompress([buf[16:]],
                                                        size=(h // 2 ** level, w // 2 ** level))
        _, _, u8 = self._model['decoder'](y_q, as_uint8='only')
        img = np.ascontiguousarray(u8[0].cpu().numpy())
        return ndarray_copy(img, out)


class ConvolutionalAutoencoderBottleneck(Codec):
    """zarr chunk codec ``'cae_bn'``: fp32 latent HWC <-> header + rANS stream (R:587-673)."""
    codec_id = 'cae_bn'

    def __init__(self, channels_bn, fact_ent=None, filters=None, fact_ent_checkpoint=None,
                 gpu=False):
        if fact_ent is not None:
            filters = list(fact_ent.filters)
            fact_ent_checkpoint = {n: self._tensor2bytes(p)
                                   for n, p in fact_ent.named_parameters()}
        self.filters = filters
        self.channels_bn = channels_bn
        self.fact_ent_checkpoint = fact_ent_checkpoint
        self.gpu = gpu
        self._setup_encoder()

    def _setup_encoder(self):
        if not torch.cuda.is_available():
            raise C.CaeError("codec 'cae_bn' needs a CUDA device (no CPU fallback)")
        self._fact_ent = EntropyBottleneck(channels=self.channels_bn, filters=self.filters)
        state = {n: self._bytes2tensor(b) for n, b in self.fact_ent_checkpoint.items()}
        self._fact_ent.load_state_dict(state, strict=False)
        self._fact_ent.cuda().eval()
        self._fact_ent.update(force=True)

    @staticmethod
    def _tensor2bytes(tensor):
        buf = io.BytesIO()
        torch.save(tensor.cpu().detach(), buf)
        return base64.b64encode(buf.getvalue()).decode('ascii')

    @staticmethod
    def _bytes2tensor(buf):
        return torch.load(io.BytesIO(base64.b64decode(buf)), weights_only=False)

    def encode(self, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float32)
        h, w, c = buf.shape
        y = torch.from_numpy(buf).cuda().permute(2, 0, 1).reshape(1, c, h, w).contiguous()
        strings = self._fact_ent.compress(y)
        return struct.pack('>QQ', h, w) + strings[0]

    def decode(self, buf, out=None):
        if out is not None:
            out = ensure_contiguous_ndarray(out)
        buf = bytes(buf)
        h, w = struct.unpack('>QQ', buf[:16])
        y_q = self._fact_ent.decompress([buf[16:]], size=(h, w))
        arr = np.ascontiguousarray(y_q[0].permute(1, 2, 0).float().cpu().numpy())
        return ndarray_copy(arr, out)
