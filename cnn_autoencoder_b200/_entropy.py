"""Factorized-prior entropy model with the interface the reference uses from
CompressAI (``compressai.entropy_models.EntropyBottleneck``; imported at
``src/models/tasks/_autoencoders.py:12``; algorithm: SURVEY.md Appendix A.1).

Same constructor, state-dict keys (``_matrix{i}``, ``_bias{i}``, ``_factor{i}``,
``quantiles``, ``target``, ``_offset``, ``_quantized_cdf``, ``_cdf_length``) and
methods the reference calls: ``forward`` (``_taskutils.py:97``), ``compress`` /
``decompress`` (``_autoencoders.py:549-551, 568-572, 645-647, 662-665``),
``update(force=True)`` (``:502, :615``), ``loss()`` (``_lossutils.py:70``),
attributes ``channels`` / ``filters`` / ``quantiles``.

Eval-mode ``forward`` and the symbol extraction of ``compress`` run in one CUDA
kernel (``cae_eb_quantize``: round, per-(channel, symbol) likelihood table,
histogram, rate); the tables come from this module's own fp32 torch ops, so the
table lookup equals evaluating the density.  Entropy coding is the C++ range
coder behind the C ABI.  Training mode (additive-noise proxy, likelihood with
gradients) stays on torch autograd ops on the device.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi as C

_LUT_MARGIN = 32
_LUT_EDGE = 8          # symbols per side that must sit on the likelihood bound
_LUT_MAX = 2048        # widest half range of the likelihood table


class _LowerBound(torch.autograd.Function):
    """max(x, bound); gradient passes where x >= bound or where it pushes x up."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad):
        x, bound = ctx.saved_tensors
        keep = (x >= bound) | (grad < 0)
        return keep.type(grad.dtype) * grad, None


class _LikelihoodLowerBound(nn.Module):
    """State-dict shell of CompressAI's ``LowerBound`` submodule: every genuine
    EntropyBottleneck checkpoint carries ``likelihood_lower_bound.bound``
    (``scripts/transfer_weights.py`` of the reference lists it)."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer('bound', torch.tensor([float(bound)]))

    def forward(self, x):
        return _LowerBound.apply(x, self.bound.to(x.dtype))


class _EbTrainFn(torch.autograd.Function):
    """Training-mode forward / backward of the bottleneck as the two fused kernels
    ``cae_eb_train_fwd`` / ``cae_eb_train_bwd``.  ``blob`` (C x 58) holds the effective
    parameters softplus(_matrix_i), _bias_i, tanh(_factor_i); it is built with torch ops from the
    module's parameters, so autograd carries the kernel's ``d/d blob`` through those (tiny)
    Jacobians into ``_matrix*`` / ``_bias*`` / ``_factor*``."""

    @staticmethod
    def forward(ctx, y, noise, blob, bound):
        y = y.contiguous().float()
        n, c = y.shape[0], y.shape[1]
        hw = y[0, 0].numel()
        y_hat, lik = torch.empty_like(y), torch.empty_like(y)
        blob = blob.contiguous().float()
        stream = ctypes.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)
        C.check(C.lib().cae_eb_train_fwd(y.data_ptr(), noise.data_ptr() if noise is not None else None,
                                         blob.data_ptr(), n, c, hw, bound, y_hat.data_ptr(),
                                         lik.data_ptr(), stream))
        ctx.save_for_backward(y_hat, blob)
        ctx.bound = bound
        return y_hat, lik

    @staticmethod
    def backward(ctx, g_yhat, g_lik):
        y_hat, blob = ctx.saved_tensors
        n, c = y_hat.shape[0], y_hat.shape[1]
        hw = y_hat[0, 0].numel()
        g_y = torch.empty_like(y_hat)
        g_blob = torch.zeros_like(blob)
        g_yhat = g_yhat.contiguous().float() if g_yhat is not None else None
        g_lik = g_lik.contiguous().float() if g_lik is not None else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(y_hat.device).cuda_stream)
        C.check(C.lib().cae_eb_train_bwd(y_hat.data_ptr(), blob.data_ptr(), n, c, hw, ctx.bound,
                                         g_yhat.data_ptr() if g_yhat is not None else None,
                                         g_lik.data_ptr() if g_lik is not None else None,
                                         g_y.data_ptr(), g_blob.data_ptr(), stream))
        return g_y, None, g_blob, None


class QuantRequest:
    """Outputs of the quantizer when it runs inside the epilogue of the last analysis
    convolution (``cae_conv_desc.quant``): ``y_q`` (fp32 NCHW), ``hist`` (C x bins),
    ``rate`` (total bits, float64[1]), optionally ``sym`` (int32 NCHW) and ``planar`` (y_q in
    the synthesis track's input layout).  ``done`` stays False when the track could not fuse
    (tiny nets on the direct kernel, GDN / residual last layers): the caller then runs
    ``EntropyBottleneck.quantize_rate`` on the latent as before."""

    def __init__(self, eb, want_sym=False, want_planar=True, want_yq=True, want_stats=True):
        self.eb, self.want_sym, self.want_planar = eb, want_sym, want_planar
        self.want_yq, self.want_stats = want_yq, want_stats
        self.done = False
        self.y_q = self.sym = self.hist = self.rate = self.planar = self.status = None
        self._struct = None

    def prepare(self, n, c, h, w, device):
        from . import _ops as O
        eb = self.eb
        if c != eb.channels:
            raise ValueError(f'expected {eb.channels} channels, got {c}')
        tb = eb._device_tables()
        q = C.QuantFuse()
        q.tables = eb._abi_tables(tb)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        q.status = self.status.data_ptr()
        if self.want_yq:
            self.y_q = torch.empty((n, c, h, w), dtype=torch.float32, device=device)
            q.y_q = self.y_q.data_ptr()
        if self.want_stats:
            self.hist = torch.zeros((c, tb['lut_len']), dtype=torch.int32, device=device)
            self.rate = torch.zeros(1, dtype=torch.float64, device=device)
            q.hist, q.rate_bits = self.hist.data_ptr(), self.rate.data_ptr()
        if self.want_sym:
            self.sym = torch.empty((n, c, h, w), dtype=torch.int32, device=device)
            q.symbols = self.sym.data_ptr()
        if self.want_planar:
            self.planar = O.alloc_act(C.FMT_F16_PLANAR, n, c, h, w, C.HALO_KEEP, device=device)
            q.y_q_planar = self.planar.desc()
        else:
            q.y_q_planar = C.Tensor(None, C.FMT_NONE, 0, 0, 0)
        self._struct = q                      # keeps the table pointers alive until the launch
        return q


class EntropyBottleneck(nn.Module):
    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3),
                 likelihood_bound=1e-9, entropy_coder_precision=16, **kwargs):
        super().__init__()
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = float(likelihood_bound) > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = _LikelihoodLowerBound(likelihood_bound)

        dims = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = float(np.log(np.expm1(1 / scale / dims[i + 1])))
            self.register_parameter(f'_matrix{i:d}', nn.Parameter(
                torch.full((self.channels, dims[i + 1], dims[i]), init)))
            self.register_parameter(f'_bias{i:d}', nn.Parameter(
                torch.empty(self.channels, dims[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f'_factor{i:d}', nn.Parameter(
                    torch.zeros(self.channels, dims[i + 1], 1)))
        self.quantiles = nn.Parameter(
            torch.tensor([-self.init_scale, 0.0, self.init_scale]).repeat(self.channels, 1, 1))
        t = float(np.log(2 / self.tail_mass - 1))
        self.register_buffer('target', torch.tensor([-t, 0.0, t]))
        self.register_buffer('_offset', torch.IntTensor())
        self.register_buffer('_quantized_cdf', torch.IntTensor())
        self.register_buffer('_cdf_length', torch.IntTensor())
        self._tables = None
        self._tables_key = None
        self.noise_fn = None

    # ------------------------------------------------------------- density
    def _logits_cumulative(self, v, stop_gradient):
        for i in range(len(self.filters) + 1):
            m = getattr(self, f'_matrix{i:d}')
            b = getattr(self, f'_bias{i:d}')
            if stop_gradient:
                m, b = m.detach(), b.detach()
            v = torch.matmul(F.softplus(m), v) + b
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
        return self.quantiles[:, 0, 1].detach()

    @property
    def likelihood_bound(self):
        """The bound in force: the loaded ``likelihood_lower_bound.bound`` buffer."""
        return float(self.likelihood_lower_bound.bound.item()) if self.use_likelihood_bound else 0.0

    def _bound(self, ref):
        return self.likelihood_lower_bound.bound.to(device=ref.device, dtype=ref.dtype)

    # CompressAI >= 1.2.5 stores the density parameters as ParameterLists
    _NEW_KEYS = (('matrices.', '_matrix'), ('biases.', '_bias'), ('factors.', '_factor'))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for new, old in self._NEW_KEYS:
            for k in [k for k in state_dict if k.startswith(prefix + new)]:
                idx = k[len(prefix + new):]
                if idx.isdigit():
                    state_dict[prefix + old + idx] = state_dict.pop(k)
        if self.use_likelihood_bound:
            # fixtures written by this repo before the buffer existed lack the key
            state_dict.setdefault(prefix + 'likelihood_lower_bound.bound',
                                  self.likelihood_lower_bound.bound.clone())
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _effective_blob(self):
        """C x 58 effective parameters in the layout ``cae_eb_train_fwd`` documents, as a
        differentiable function of the module's parameters."""
        parts = []
        K = len(self.filters)
        for i in range(K + 1):
            parts.append(F.softplus(getattr(self, f'_matrix{i:d}')).reshape(self.channels, -1))
            parts.append(getattr(self, f'_bias{i:d}').reshape(self.channels, -1))
            if i < K:
                parts.append(torch.tanh(getattr(self, f'_factor{i:d}')).reshape(self.channels, -1))
        return torch.cat(parts, dim=1)

    def _forward_train_cuda(self, x):
        """Training-mode forward on the fused kernels (filters (3, 3, 3, 3) on a CUDA device)."""
        if self.noise_fn is not None:
            # reproducible runs supply the noise in CompressAI's C x 1 x M layout
            nd = x.dim()
            perm = [1, 0] + list(range(2, nd))
            shape = x.permute(*perm).shape
            noise = self.noise_fn(x.permute(*perm).reshape(shape[0], 1, -1))
            noise = noise.reshape(shape).permute(*perm).contiguous().to(x.device, torch.float32)
        else:
            noise = torch.empty_like(x, dtype=torch.float32).uniform_(-0.5, 0.5)
        return _EbTrainFn.apply(x, noise, self._effective_blob(), self.likelihood_bound_const)

    @property
    def likelihood_bound_const(self):
        # the bound as a Python float without a device synchronisation per step
        if getattr(self, '_bound_cache', None) is None or self._bound_cache[0] != self.likelihood_lower_bound.bound._version:
            self._bound_cache = (self.likelihood_lower_bound.bound._version, self.likelihood_bound)
        return self._bound_cache[1]

    def _forward_torch(self, x, training=False):
        """The model written with torch ops (training path; also what the tables
        are built from).  C x 1 x M layout as in CompressAI."""
        nd = x.dim()
        perm = list(range(nd))
        perm[0], perm[1] = 1, 0
        xp = x.permute(*perm).contiguous()
        shape = xp.size()
        v = xp.reshape(xp.size(0), 1, -1)
        if training:
            # additive uniform noise proxy; `noise_fn` (tests / reproducible data parallel runs)
            # may supply the noise for this rank's shard, e.g. keyed on the global sample index
            noise = self.noise_fn(v) if self.noise_fn is not None else \
                torch.empty_like(v).uniform_(-0.5, 0.5)
            v = v + noise
        else:
            med = self.quantiles[:, :, 1:2]
            v = torch.round(v - med) + med
        lik = self._likelihood(v)
        if self.use_likelihood_bound:
            lik = self.likelihood_lower_bound(lik)
        out = v.reshape(shape).permute(*perm).contiguous()
        lik = lik.reshape(shape).permute(*perm).contiguous()
        return out, lik

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    # -------------------------------------------------------------- tables
    def _host_clone(self):
        """Parameter-only copy on the CPU in fp32: every table (integer CDFs, likelihood
        LUT) is computed with CPU fp32 ops whatever device the module lives on, so that
        whoever decodes the stream or checks the rate can reproduce it."""
        host = EntropyBottleneck.__new__(EntropyBottleneck)
        nn.Module.__init__(host)
        host.filters = self.filters
        for name, prm in self.named_parameters():
            host.register_parameter(name, nn.Parameter(prm.detach().float().cpu(),
                                                       requires_grad=False))
        return host

    def update(self, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        # Tables are always computed with CPU fp32 ops, whatever device the module lives on:
        # the integer CDF must be reproducible by whoever decodes the stream.
        dev = self.quantiles.device
        with torch.no_grad():
            host = self._host_clone()
            q = host.quantiles
            med = q[:, 0, 1]
            minima = torch.clamp(torch.ceil(med - q[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(q[:, 0, 2] - med).int(), min=0)
            pmf_start = med - minima
            pmf_length = maxima + minima + 1
            max_length = int(pmf_length.max().item())
            samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
            lower = host._logits_cumulative(samples - 0.5, stop_gradient=True)
            upper = host._logits_cumulative(samples + 0.5, stop_gradient=True)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            pmf_h = pmf.numpy()
            tail_h = tail.numpy()
            len_h = pmf_length.numpy()
            cdf = np.zeros((self.channels, max_length + 2), dtype=np.int32)
            L = C.lib()
            for c in range(self.channels):
                prob = np.ascontiguousarray(
                    np.concatenate([pmf_h[c, :len_h[c]], tail_h[c]]), dtype=np.float32)
                row = np.zeros(prob.shape[0] + 1, dtype=np.uint32)
                C.check(L.cae_pmf_to_quantized_cdf(prob.ctypes.data, prob.shape[0],
                                                   self.entropy_coder_precision, row.ctypes.data))
                cdf[c, :row.shape[0]] = row.astype(np.int32)
            self._offset = (-minima).int().to(dev)
            self._quantized_cdf = torch.from_numpy(cdf).to(dev)
            self._cdf_length = (pmf_length + 2).int().to(dev)
        self._tables = None
        return True

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _device_tables(self):
        """Likelihood table / MLP blob / medians on the device, rebuilt whenever a
        parameter changed."""
        key = self._param_key()
        if self._tables is not None and self._tables_key == key:
            return self._tables
        dev = self.quantiles.device
        bound = self.likelihood_bound
        with torch.no_grad():
            host = self._host_clone()
            q = host.quantiles
            med = q[:, 0, 1]
            lo = int(torch.floor((q[:, 0, 0] - med).min()).item()) - _LUT_MARGIN
            hi = int(torch.ceil((q[:, 0, 2] - med).max()).item()) + _LUT_MARGIN
            # Widen the table until both ends sit in the region where the likelihood is the
            # lower bound (checked over a run of _LUT_EDGE symbols per side and channel): every
            # symbol outside the table then has exactly the bound and needs no density
            # evaluation on the device.  Give up (MLP fallback) beyond +-_LUT_MAX.
            tail_lik = 0.0
            while True:
                lo, hi = max(lo, -_LUT_MAX), min(hi, _LUT_MAX)
                syms = torch.arange(lo, hi + 1, dtype=torch.float32)
                v = syms[None, None, :] + q[:, :, 1:2]
                raw = host._likelihood(v)[:, 0, :]
                lik = torch.clamp(raw, min=bound) if bound > 0 else raw
                if bound <= 0:
                    break
                left_ok = bool((raw[:, :_LUT_EDGE] <= bound).all())
                right_ok = bool((raw[:, -_LUT_EDGE:] <= bound).all())
                if left_ok and right_ok:
                    tail_lik = bound
                    break
                if (lo <= -_LUT_MAX or left_ok) and (hi >= _LUT_MAX or right_ok):
                    break
                if not left_ok:
                    lo -= max(32, (hi - lo) // 2)
                if not right_ok:
                    hi += max(32, (hi - lo) // 2)
            blob = []
            K = len(self.filters)
            for i in range(K + 1):
                blob.append(F.softplus(getattr(host, f'_matrix{i:d}')).reshape(self.channels, -1))
                blob.append(getattr(host, f'_bias{i:d}').reshape(self.channels, -1))
                if i < K:
                    blob.append(torch.tanh(getattr(host, f'_factor{i:d}')).reshape(self.channels, -1))
            mlp = torch.cat(blob, dim=1).contiguous().float()
            tb = dict(medians=med.contiguous().float().clone().to(dev),
                      lut=lik.contiguous().float().to(dev), lut_min=lo, lut_len=hi - lo + 1,
                      mlp=mlp.to(dev), tail_lik=tail_lik, lik_bound=bound)
        if max((1,) + self.filters) > 8 or K + 1 > 9:
            tb['mlp'] = None
        self._tables, self._tables_key = tb, key
        return tb

    def _abi_tables(self, tb):
        t = C.EbTables()
        t.medians = tb['medians'].data_ptr()
        t.lut = tb['lut'].data_ptr()
        t.lut_min, t.lut_len = tb['lut_min'], tb['lut_len']
        dims = (1,) + self.filters + (1,)
        if tb['mlp'] is not None:
            t.mlp = tb['mlp'].data_ptr()
            t.n_layers = len(self.filters) + 1
            t.mlp_stride = tb['mlp'].shape[1]
            for i, d in enumerate(dims):
                t.dims[i] = d
        t.hist_min, t.hist_bins = tb['lut_min'], tb['lut_len']
        t.tail_lik = tb.get('tail_lik', 0.0)
        t.lik_bound = tb.get('lik_bound', 0.0)
        return t

    def _quantize_cuda(self, x, want_yq=True, want_p=True, want_sym=False, want_hist=False,
                       want_rate=False):
        if not x.is_cuda:
            raise C.CaeError('EntropyBottleneck: the eval path runs on the GPU only '
                             '(no CPU fallback); move the module and the latent to cuda')
        if self.quantiles.device != x.device:
            raise C.CaeError('EntropyBottleneck parameters and input are on different devices')
        x = x.contiguous().float()
        n, c = x.shape[0], x.shape[1]
        if c != self.channels:
            raise ValueError(f'expected {self.channels} channels, got {c}')
        hw = x[0, 0].numel()
        tb = self._device_tables()
        t = self._abi_tables(tb)
        y_q = torch.empty_like(x) if want_yq else None
        p_y = torch.empty_like(x) if want_p else None
        sym = torch.empty(x.shape, dtype=torch.int32, device=x.device) if want_sym else None
        hist = torch.zeros((c, tb['lut_len']), dtype=torch.int32, device=x.device) if want_hist else None
        rate = torch.zeros(1, dtype=torch.float64, device=x.device) if want_rate else None
        status = torch.zeros(1, dtype=torch.int32, device=x.device)
        ptr = lambda a: a.data_ptr() if a is not None else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        C.check(C.lib().cae_eb_quantize(x.data_ptr(), n, c, hw, ctypes.byref(t), ptr(y_q), ptr(p_y),
                                        ptr(sym), ptr(hist), ptr(rate), status.data_ptr(), stream))
        return y_q, p_y, sym, hist, rate

    # -------------------------------------------------------------- forward
    def forward(self, x, training=None):
        if training is None:
            training = self.training
        if training and x.is_cuda and self.filters == (3, 3, 3, 3) and self.use_likelihood_bound \
                and not os.environ.get('CAE_EB_TRAIN_TORCH'):
            return self._forward_train_cuda(x)
        if training or torch.is_grad_enabled() and x.requires_grad:
            return self._forward_torch(x, training=training)
        y_q, p_y, _, _, _ = self._quantize_cuda(x)
        return y_q, p_y

    def quant_request(self, want_sym=False, want_planar=True, want_yq=True, want_stats=True):
        """Request object for the quantizer fused into the latent layer's epilogue
        (``Analyzer.forward(x, quant=req)``); see ``QuantRequest``."""
        return QuantRequest(self, want_sym, want_planar, want_yq, want_stats)

    def quantize_rate(self, x):
        """(y_q, C x bins histogram, total bits) in one pass: what the encode+rate+decode
        pipeline needs between the two transforms."""
        y_q, _, _, hist, rate = self._quantize_cuda(x, want_yq=True, want_p=False,
                                                    want_hist=True, want_rate=True)
        return y_q, hist, rate

    def symbols_hist_rate(self, x):
        """(int32 symbols, C x bins histogram, total bits) in one pass (K11/K14)."""
        _, _, sym, hist, rate = self._quantize_cuda(x, want_yq=False, want_p=False, want_sym=True,
                                                    want_hist=True, want_rate=True)
        return sym, hist, rate

    # ------------------------------------------------------ entropy coding
    def _host_tables(self):
        if self._offset.numel() == 0:
            raise C.CaeError('EntropyBottleneck.update() must be called before compress/decompress')
        cdf = np.ascontiguousarray(self._quantized_cdf.cpu().numpy(), dtype=np.int32)
        sizes = np.ascontiguousarray(self._cdf_length.cpu().numpy().reshape(-1), dtype=np.int32)
        offs = np.ascontiguousarray(self._offset.cpu().numpy().reshape(-1), dtype=np.int32)
        return cdf, sizes, offs

    # Streams per call from which the batched device coder is used instead of the host thread
    # pool.  One thread per stream is a dependent chain (~70 ms for the 196 608 symbols of a
    # 512^2 tile whatever the stream count, tools/micro/coderbench.py), so it wins with a
    # hundred or more tiles in flight: 128 streams 0.43 GP/s, 1024 streams 1.4 GP/s including
    # the copy back (15 GP/s device side at 4096 streams); host pool of 16 cores 0.4 GP/s.
    GPU_CODER_MIN_STREAMS = 128

    def compress(self, x):
        _, _, sym, _, _ = self._quantize_cuda(x, want_yq=False, want_p=False, want_sym=True)
        if sym.shape[0] >= self.GPU_CODER_MIN_STREAMS:
            return self.encode_symbols_gpu(sym)
        sym_h = sym.reshape(sym.shape[0], sym.shape[1], -1).cpu().numpy()
        return [encode_symbols(sym_h[i], *self._host_tables()) for i in range(sym_h.shape[0])]

    def _dev_tables(self, device):
        """(cdf, sizes, offsets, encode table) on ``device``, rebuilt when ``update()`` changed
        the CDFs.  The encode table is the division-free form of the state update
        (``cae_rans_build_enc_table``)."""
        key = (str(device), self._quantized_cdf.data_ptr(), self._quantized_cdf._version,
               self._cdf_length._version, self._offset._version)
        cache = getattr(self, '_dev_tables_cache', None)
        if cache is not None and cache[0] == key:
            return cache[1]
        cdf = self._quantized_cdf.to(device=device, dtype=torch.int32).contiguous()
        sizes = self._cdf_length.to(device=device, dtype=torch.int32).reshape(-1).contiguous()
        offs = self._offset.to(device=device, dtype=torch.int32).reshape(-1).contiguous()
        L = C.lib()
        nbytes = L.cae_rans_enc_table_bytes(cdf.shape[0], cdf.shape[1])
        table = torch.empty(nbytes, dtype=torch.uint8, device=device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        C.check(L.cae_rans_build_enc_table(cdf.data_ptr(), cdf.shape[0], cdf.shape[1],
                                           sizes.data_ptr(), table.data_ptr(), stream))
        self._dev_tables_cache = (key, (cdf, sizes, offs, table))
        return cdf, sizes, offs, table

    def encode_symbols_device(self, sym):
        """int32 symbols N x C x ... (device) -> (packed uint8 device tensor holding the N
        streams back to back, int64 host array of N + 1 byte offsets).  All streams are coded
        concurrently (one thread per stream, ``cae_rans_encode_batch``); the only host
        synchronisation is reading the N stream lengths."""
        if self._offset.numel() == 0:
            raise C.CaeError('EntropyBottleneck.update() must be called before compress/decompress')
        sym = sym.contiguous()
        n, c = sym.shape[0], sym.shape[1]
        hw = sym[0, 0].numel()
        dev = sym.device
        cdf, sizes, offs, table = self._dev_tables(dev)
        cap = c * hw + 64
        words = torch.empty((n, cap), dtype=torch.int32, device=dev)
        nwords = torch.empty(n, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L = C.lib()
        use_table = True
        C.check(L.cae_rans_encode_batch(sym.data_ptr(), n, c, hw, cdf.data_ptr(), cdf.shape[1],
                                        sizes.data_ptr(), offs.data_ptr(),
                                        table.data_ptr() if use_table else None,
                                        words.data_ptr(), cap, nwords.data_ptr(),
                                        status.data_ptr(), stream))
        off_d = torch.empty(n + 1, dtype=torch.int64, device=dev)
        C.check(L.cae_rans_scan(nwords.data_ptr(), n, off_d.data_ptr(), stream))
        e = off_d.cpu().numpy()                            # sync: stream lengths are data dependent
        if int(status.item()) & 1:
            raise C.CaeError('device entropy coder: a stream overflowed its staging buffer')
        packed = torch.empty(int(e[-1]), dtype=torch.int32, device=dev)
        C.check(L.cae_rans_compact(words.data_ptr(), n, cap, nwords.data_ptr(), off_d.data_ptr(),
                                   packed.data_ptr(), stream))
        return packed.view(torch.uint8), e * 4

    def encode_symbols_gpu(self, sym):
        """int32 symbols N x C x ... (device) -> list of N byte strings."""
        packed, off = self.encode_symbols_device(sym)
        host = packed.cpu().numpy()
        return [host[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]

    def decode_streams_device(self, words, word_off, hw, defer_status=False):
        """``words``: int32 device tensor holding N streams back to back, ``word_off``: N + 1
        word offsets (host array) -> int32 symbols N x C x hw on the device, all streams decoded
        concurrently (``cae_rans_decode_batch``).  ``defer_status``: do not wait for the kernel;
        returns (symbols, status tensor) and the caller hands the tensor to
        ``check_decode_status`` once it has synchronised anyway."""
        if self._offset.numel() == 0:
            raise C.CaeError('EntropyBottleneck.update() must be called before compress/decompress')
        dev = words.device
        n, c = len(word_off) - 1, self.channels
        off_d = torch.from_numpy(np.ascontiguousarray(word_off, dtype=np.int64)).to(dev)
        cdf, sizes, offs, _ = self._dev_tables(dev)
        sym = torch.empty((n, c, hw), dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        C.check(C.lib().cae_rans_decode_batch(words.data_ptr(), off_d.data_ptr(), n, c, hw,
                                              cdf.data_ptr(), cdf.shape[1], sizes.data_ptr(),
                                              offs.data_ptr(), sym.data_ptr(), status.data_ptr(), stream))
        if defer_status:
            return sym, status
        self.check_decode_status(status)
        return sym

    @staticmethod
    def check_decode_status(status):
        if int(status.item()) & 2:
            raise C.CaeError('device entropy decoder: a stream is truncated')

    def decode_streams_gpu(self, strings, hw):
        """list of N byte strings -> int32 symbols N x C x hw on the device."""
        dev = self.quantiles.device
        lens = np.array([len(s) for s in strings], dtype=np.int64)
        if (lens % 4).any() or (lens < 8).any():
            raise C.CaeError('entropy-coded stream is not a whole number of words >= 2')
        off = np.zeros(len(strings) + 1, dtype=np.int64)
        np.cumsum(lens // 4, out=off[1:])
        blob = torch.from_numpy(np.frombuffer(b''.join(bytes(s) for s in strings), dtype=np.int32).copy())
        return self.decode_streams_device(blob.to(dev, non_blocking=True), off, hw)

    def decompress(self, strings, size):
        hw = int(np.prod(size))
        if len(strings) >= self.GPU_CODER_MIN_STREAMS and self.quantiles.is_cuda:
            sym = self.decode_streams_gpu(strings, hw)
            med = self._medians().reshape(1, -1, *([1] * len(size)))
            return sym.reshape(len(strings), self.channels, *size).type_as(med) + med
        cdf, sizes, offs = self._host_tables()
        out = np.empty((len(strings), self.channels, hw), dtype=np.int32)
        for i, s in enumerate(strings):
            out[i] = decode_symbols(s, self.channels, hw, cdf, sizes, offs)
        dev = self.quantiles.device
        y = torch.from_numpy(out).to(dev).reshape(len(strings), self.channels, *size)
        med = self._medians().reshape(1, -1, *([1] * len(size)))
        return y.type_as(med) + med


def encode_symbols(sym_chw, cdf, sizes, offs):
    """int32 symbols (C x hw, host) -> bytes, through ``cae_rans_encode``."""
    sym = np.ascontiguousarray(sym_chw, dtype=np.int32)
    c, hw = sym.shape
    cap = 4 * (2 * c * hw + 64)
    buf = np.empty(cap, dtype=np.uint8)
    nbytes = ctypes.c_size_t(0)
    C.check(C.lib().cae_rans_encode(sym.ctypes.data, c, hw, cdf.ctypes.data, cdf.shape[1],
                                    sizes.ctypes.data, offs.ctypes.data, buf.ctypes.data, cap,
                                    ctypes.byref(nbytes)))
    return buf[:nbytes.value].tobytes()


def decode_symbols(data, c, hw, cdf, sizes, offs):
    enc = np.frombuffer(bytes(data), dtype=np.uint8)
    out = np.empty((c, hw), dtype=np.int32)
    C.check(C.lib().cae_rans_decode(enc.ctypes.data, enc.shape[0], c, hw, cdf.ctypes.data,
                                    cdf.shape[1], sizes.ctypes.data, offs.ctypes.data,
                                    out.ctypes.data))
    return out
