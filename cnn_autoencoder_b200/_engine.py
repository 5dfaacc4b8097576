"""Layer-plan executor: turns a track of units (the parameter-holding modules
of ``_autoencoders.py``) into a list of fused convolution steps and runs them
through the C ABI.

One step = one ABI call ``out = post_act(pre_act(conv(in) + bias) + skip)``;
the unit structure of the reference (``src/models/tasks/_autoencoders.py``
53-304) maps onto steps as:

  plain unit     [conv s1 -> act] -> conv s2 -> [act]
                 = step(conv s1, pre=act) ; step(conv s2, pre=act)
  residual unit  res = conv_a -> act_a -> [conv_b (-> act_b in the decoder)];
                 fx = res + x ; [act_m] -> conv_c -> [act_c]
                 = step(conv_a, pre=act_a) ;
                   step(conv_b, pre=act_b, skip=x, post=act_m) ; step(conv_c, pre=act_c)
                 (without conv_b: step(conv_a, pre=act_a, skip=x, post=act_m))

Eval-mode BatchNorm (R:72-73) is folded into the weights/bias when they are
packed.  Every tensor between steps lives in the internal fp16 layout its
consumer wants (planar for stride-1 / transposed, parity-split for stride-2;
reflect halo for Conv2d consumers, zero halo for ConvTranspose2d consumers).
"""
import os

import torch
import torch.nn as nn

from . import _cabi as C
from . import _ops as O

_ACT_CODE = {None: C.ACT_NONE, 'Identity': C.ACT_NONE, 'LeakyReLU': C.ACT_LEAKY_RELU,
             'ReLU': C.ACT_RELU}


def act_code(kind):
    if kind not in _ACT_CODE:
        raise NotImplementedError(f'activation {kind!r} has no CUDA epilogue')
    return _ACT_CODE[kind]


class Step:
    """One fused convolution."""

    def __init__(self, conv, bn=None, pre_act=None, skip=None, post_act=None):
        self.conv = conv            # nn.Conv2d / nn.ConvTranspose2d used as the parameter holder
        self.bn = bn
        self.pre_act = pre_act
        self.skip = skip            # index of the tensor added before post_act (0 = track input)
        self.post_act = post_act
        self.gdn = None             # GDN module applied to the conv output before the skip add
        self.transposed = isinstance(conv, nn.ConvTranspose2d)
        stride = conv.stride[0]
        if conv.kernel_size != (3, 3) or conv.dilation != (1, 1):
            raise NotImplementedError('only 3x3 convolutions have CUDA kernels')
        self.groups = conv.groups     # > 1: depthwise-style layers of groups=True nets (direct kernel)
        self.kind = ({1: C.CONVT_S1, 2: C.CONVT_S2} if self.transposed
                     else {1: C.CONV_S1, 2: C.CONV_S2})[stride]
        self.c_in = conv.in_channels
        self.c_out = conv.out_channels
        self.pad_mode = C.PAD_ZERO if self.transposed else (
            C.PAD_REFLECT if conv.padding_mode == 'reflect' else C.PAD_ZERO)
        self._cache_key = None
        self._cache = None
        self._proj_key = None
        self._proj_cache = None

    # weights in the form the chosen kernel wants, rebuilt when a parameter changes
    def materialise(self, igemm):
        params = [self.conv.weight, self.conv.bias]
        if self.bn is not None:
            params += [self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        key = (igemm,) + tuple((p.data_ptr(), p._version) if p is not None else None
                               for p in params)
        if key == self._cache_key:
            return self._cache
        with torch.no_grad():
            w = self.conv.weight.detach().float()
            b = self.conv.bias.detach().float() if self.conv.bias is not None else None
            scale = None
            if self.bn is not None:
                inv = torch.rsqrt(self.bn.running_var.float() + self.bn.eps)
                gamma = self.bn.weight.float() if self.bn.weight is not None else torch.ones_like(inv)
                beta = self.bn.bias.float() if self.bn.bias is not None else torch.zeros_like(inv)
                scale = gamma * inv
                b0 = b if b is not None else torch.zeros_like(inv)
                b = (b0 - self.bn.running_mean.float()) * scale + beta
            if igemm:
                wdev = O.pack_weights(self.kind, w, scale=scale)
            else:
                if scale is not None:
                    if self.transposed and self.groups > 1:      # (c_in, c_out / G, 3, 3)
                        g = self.groups
                        w = (w.view(g, -1, w.shape[1], 3, 3) * scale.view(g, 1, -1, 1, 1)).view_as(w)
                    else:
                        w = w * (scale.view(1, -1, 1, 1) if self.transposed else scale.view(-1, 1, 1, 1))
                wdev = w.contiguous()
            self._cache = (wdev, b.contiguous() if b is not None else None)
        self._cache_key = key
        return self._cache


    def _folded(self):
        """(weight, bias, per-output-channel scale) with eval-mode BatchNorm folded."""
        w = self.conv.weight.detach().float()
        b = self.conv.bias.detach().float() if self.conv.bias is not None else None
        scale = None
        if self.bn is not None:
            inv = torch.rsqrt(self.bn.running_var.float() + self.bn.eps)
            gamma = self.bn.weight.float() if self.bn.weight is not None else torch.ones_like(inv)
            beta = self.bn.bias.float() if self.bn.bias is not None else torch.zeros_like(inv)
            scale = gamma * inv
            b0 = b if b is not None else torch.zeros_like(inv)
            b = (b0 - self.bn.running_mean.float()) * scale + beta
        return w, b, scale

    def materialise_proj(self):
        """The image layer as the projection operand of the layer before it (cae_conv_desc.proj)."""
        params = [self.conv.weight, self.conv.bias]
        if self.bn is not None:
            params += [self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        key = ('proj',) + tuple((p.data_ptr(), p._version) if p is not None else None for p in params)
        if key == self._proj_key:
            return self._proj_cache
        with torch.no_grad():
            w, b, scale = self._folded()
            self._proj_cache = (O.pack_proj_weights(w, scale=scale),
                                b.contiguous() if b is not None else None)
        self._proj_key = key
        return self._proj_cache


def steps_from_units(units):
    """units: iterable of unit modules exposing ``layout`` (list of op tuples built
    by ``_autoencoders._unit_layout``), ``model`` and (residual) ``res_model``."""
    steps = []
    tensor_idx = 0            # index of the tensor currently flowing (0 = input)
    for unit in units:
        unit_in = tensor_idx
        pending = None        # last Step, still accepting a trailing bn / activation
        skip_armed = False
        for op in unit.layout:
            tag = op[0]
            if tag == 'conv':
                seq = getattr(unit, op[1])
                pending = Step(seq[op[2]])
                steps.append(pending)
                tensor_idx = len(steps)
            elif tag == 'bn':
                seq = getattr(unit, op[1])
                if pending is None or pending.pre_act is not None or pending.skip is not None:
                    raise NotImplementedError('BatchNorm2d not directly after a convolution')
                pending.bn = seq[op[2]]
            elif tag == 'act':
                kind = op[3]
                if kind in (None, 'Identity'):
                    continue
                if pending is None:
                    raise NotImplementedError('activation on the raw track input')
                if kind == 'GDN':
                    if pending.gdn is not None or pending.pre_act is not None or skip_armed:
                        raise NotImplementedError('GDN must directly follow a convolution')
                    pending.gdn = getattr(unit, op[1])[op[2]]
                    continue
                act_code(kind)
                if pending.skip is not None or skip_armed:
                    if pending.post_act is not None:
                        raise NotImplementedError('two activations after a residual add')
                    pending.post_act = kind
                else:
                    if pending.pre_act is not None:
                        raise NotImplementedError('two activations after one convolution')
                    pending.pre_act = kind
            elif tag == 'add':
                if pending is None:
                    raise NotImplementedError('residual add without a convolution')
                pending.skip = unit_in
                skip_armed = True
            else:
                raise ValueError(tag)
            if tag == 'conv':
                skip_armed = False
    return steps


class TrackExecutor:
    """Runs a list of Steps on one input through the C ABI."""

    def __init__(self, steps):
        self.steps = steps
        self._buffers = {}
        # step index -> (args, kwargs) of its last _ops.conv call, or ('head', args, kwargs)
        # for a fused stem + stride-2 pair (recorded under the stem's index); _ops.replay
        self.last_calls = {}
        self.fuse_head = not os.environ.get('CAE_NO_HEAD_FUSION')
        # Experiment, OFF by default (CAE_TAIL_L2_MB = bytes of the exchange buffer per sub-batch):
        # the last two synthesis layers (ConvTranspose 128->128 s2, then the image layer) exchange
        # the largest tensor of the path; run over sub-batches that reuse ONE small buffer, the
        # exchange stays in the 126 MB L2.  Measured on B200 (net A, 128 x 256^2 per step, round 2):
        # 1.392 ms per step without, 1.531 / 1.456 / 1.450 ms with 40 / 72 / 100 MB sub-batches --
        # the extra launches and the wave quantisation of the smaller grids cost more than the HBM
        # round trip they save, so the whole-batch schedule stays.
        self.tail_l2_bytes = int(os.environ.get('CAE_TAIL_L2_MB', '0')) << 20
        # projection fusion of the last two synthesis layers (cae_conv_desc.proj)
        self.fuse_proj = not os.environ.get('CAE_NO_PROJ_FUSION')

    @staticmethod
    def _use_igemm(step, x):
        if step.groups != 1:
            return False                      # grouped layers: HBM-bound, CUDA-core direct kernel
        if x.fmt in (C.FMT_F32_NCHW, C.FMT_U8_HWC):
            return False                      # raw image / latent from the caller: direct kernel
        return not (step.c_in <= 4 and step.c_out <= 4)

    def _consumer_layout(self, k, final_fmt):
        """(fmt, halo) of the tensor produced by step k-1 (k = consumer index)."""
        if k >= len(self.steps):
            return final_fmt, C.HALO_KEEP
        nxt = self.steps[k]
        if nxt.kind == C.CONV_S1 and nxt.c_in <= 4 and nxt.c_out <= 4:
            return C.FMT_F32_NCHW, C.HALO_KEEP      # stem -> stem stays fp32 (tiny tensors)
        fmt = C.FMT_F16_SPLIT if nxt.kind == C.CONV_S2 else C.FMT_F16_PLANAR
        halo = C.HALO_REFLECT if (not nxt.transposed and nxt.pad_mode == C.PAD_REFLECT) else C.HALO_KEEP
        return fmt, halo

    def _head_match(self, k, cur, keep, final_fmt):
        """Number of steps (2: plain unit, 3: residual unit, 0: none) starting at k that form the
        first downsampling unit reading the raw image and that the fused kernel covers."""
        if not self.fuse_head or cur.fmt not in (C.FMT_U8_HWC, C.FMT_F32_NCHW):
            return 0
        for span in (2, 3):
            if k + span > len(self.steps) - 1 or any((k + j) in keep for j in range(1, span)):
                continue
            steps = self.steps[k:k + span]
            if any(s.gdn is not None or s.transposed or s.pad_mode != steps[0].pad_mode
                   or s.groups != 1 for s in steps):
                continue
            a, down = steps[0], steps[-1]
            stems_ok = all(s.kind == C.CONV_S1 and s.c_in <= 4 and s.c_out == s.c_in
                           for s in steps[:-1])
            if not (stems_ok and down.kind == C.CONV_S2 and down.c_in == a.c_in and
                    down.c_out <= 128 and down.skip is None and down.post_act is None):
                continue
            if a.skip is not None or a.post_act is not None:
                continue
            if span == 3:
                b = steps[1]
                if not (b.skip == k and b.pre_act is None):     # res + x, then the activation
                    continue
            fmt, _ = self._consumer_layout(k + span, final_fmt)
            if fmt == C.FMT_F16_PLANAR:
                return span
        return 0

    def _proj_match(self, k, cur, keep, final_fmt):
        """True when steps k, k + 1 are the last two layers of a synthesis track in the form the
        projection fusion covers: ConvTranspose2d(c, 128, s2) -> [act], then the <= 3 channel
        image layer (``cae_conv_desc.proj``: the 128-channel tensor between them never reaches
        HBM)."""
        if not self.fuse_proj or k != len(self.steps) - 2 or (k + 1) in keep:
            return False
        a, b = self.steps[k], self.steps[k + 1]
        if not (a.kind == C.CONVT_S2 and b.kind == C.CONVT_S2 and a.c_out == 128 and b.c_in == 128
                and b.c_out <= 3 and a.c_in % 16 == 0):
            return False
        if any(s.gdn is not None or s.skip is not None or s.groups != 1 for s in (a, b)):
            return False
        if a.post_act is not None:
            return False
        return cur.fmt == C.FMT_F16_PLANAR and final_fmt in (C.FMT_U8_HWC, C.FMT_F32_NCHW)

    def _run_proj(self, k, cur, final_fmt, aux_last):
        a, b = self.steps[k], self.steps[k + 1]
        wa, ba = a.materialise(True)
        vb, bb = b.materialise_proj()
        hu, wu = O.KIND_OUT[a.kind](cur.h, cur.w)
        ho, wo = O.KIND_OUT[b.kind](hu, wu)
        dev = cur.t.device
        key = ((k, 'proj'), cur.n, hu, wu, str(dev))
        buf = self._buffers.get((k, 'proj'))
        if buf is None or buf[0] != key:
            buf = (key, O.alloc_proj(cur.n, hu, wu, dev))
            self._buffers[(k, 'proj')] = buf
        rec = buf[1]
        out = O.alloc_act(C.FMT_U8_HWC, cur.n, b.c_out, ho, wo, device=dev) if final_fmt == C.FMT_U8_HWC else None
        aux = torch.empty((cur.n, b.c_out, ho, wo), dtype=torch.float32, device=dev) \
            if (aux_last or final_fmt != C.FMT_U8_HWC) else None
        call_a = ((a.kind, cur, wa, a.c_out, None),
                  dict(igemm=True, bias=ba, skip=None, pre_act=act_code(a.pre_act),
                       post_act=C.ACT_NONE, pad_mode=a.pad_mode, aux=None, proj=(vb, rec)))
        O.conv(*call_a[0], **call_a[1])
        call_b = ('image_from_proj', (rec, cur.n, hu, wu, b.c_out),
                  dict(bias=bb, pre_act=act_code(b.pre_act), post_act=act_code(b.post_act),
                       out=out, aux=aux))
        O.image_from_proj(*call_b[1], **call_b[2])
        self.last_calls[k], self.last_calls[k + 1] = call_a, call_b
        last = out if out is not None else O.Act(aux, C.FMT_F32_NCHW, cur.n, b.c_out, ho, wo)
        return last, aux

    def _buffer(self, key, fmt, n, c, h, w, halo, device):
        full = (key, fmt, n, c, h, w, halo, str(device))
        buf = self._buffers.get(key)
        if buf is None or buf[0] != full:
            if len(self._buffers) > 64:
                self._buffers.clear()
            buf = (full, O.alloc_act(fmt, n, c, h, w, halo, device=device))
            self._buffers[key] = buf
        return buf[1]

    def _tail_match(self, k, cur, keep, final_fmt):
        """Sub-batch size (images) for the L2-resident schedule of steps k, k + 1, or 0."""
        n_steps = len(self.steps)
        if k != n_steps - 2 or not self.tail_l2_bytes or (k + 1) in keep:
            return 0
        a, b = self.steps[k], self.steps[k + 1]
        if not (a.kind == C.CONVT_S2 and b.kind == C.CONVT_S2 and 4 * b.c_out <= 16 and 4 * a.c_out > 16):
            return 0
        if any(s.gdn is not None or s.skip is not None or s.groups != 1 for s in (a, b)):
            return 0
        if cur.fmt != C.FMT_F16_PLANAR or final_fmt not in (C.FMT_U8_HWC, C.FMT_F32_NCHW):
            return 0
        per_image = O.planes_for(a.c_out) * (2 * cur.h + 2) * (2 * cur.w + 2 + 2 * O.COL_PAD) * 16
        sub = self.tail_l2_bytes // per_image
        return int(sub) if 1 <= sub < cur.n else 0

    def _run_tail(self, k, cur, sub, final_fmt, aux_last):
        a, b = self.steps[k], self.steps[k + 1]
        wa, ba = a.materialise(True)
        wb, bb = b.materialise(True)
        hu, wu = O.KIND_OUT[a.kind](cur.h, cur.w)
        ho, wo = O.KIND_OUT[b.kind](hu, wu)
        dev = cur.t.device
        u = self._buffer((k, 'tail'), C.FMT_F16_PLANAR, sub, a.c_out, hu, wu, C.HALO_KEEP, dev)
        out = O.alloc_act(C.FMT_U8_HWC, cur.n, b.c_out, ho, wo, device=dev) if final_fmt == C.FMT_U8_HWC else None
        aux = torch.empty((cur.n, b.c_out, ho, wo), dtype=torch.float32, device=dev) \
            if (aux_last or final_fmt != C.FMT_U8_HWC) else None
        for n0 in range(0, cur.n, sub):
            n1 = min(cur.n, n0 + sub)
            m = n1 - n0
            x_s = O.Act(cur.t[n0:n1], cur.fmt, m, cur.c, cur.h, cur.w, cur.halo)
            u_s = O.Act(u.t[:m], u.fmt, m, a.c_out, hu, wu, u.halo)
            call_a = ((a.kind, x_s, wa, a.c_out, u_s),
                      dict(igemm=True, bias=ba, skip=None, pre_act=act_code(a.pre_act),
                           post_act=act_code(a.post_act), pad_mode=a.pad_mode, aux=None))
            O.conv(*call_a[0], **call_a[1])
            o_s = O.Act(out.t[n0:n1], out.fmt, m, b.c_out, ho, wo) if out is not None else None
            call_b = ((b.kind, u_s, wb, b.c_out, o_s),
                      dict(igemm=True, bias=bb, skip=None, pre_act=act_code(b.pre_act),
                           post_act=act_code(b.post_act), pad_mode=b.pad_mode,
                           aux=aux[n0:n1] if aux is not None else None))
            O.conv(*call_b[0], **call_b[1])
            if n0 == 0:
                self.last_calls[k], self.last_calls[k + 1] = call_a, call_b
        last = out if out is not None else O.Act(aux, C.FMT_F32_NCHW, cur.n, b.c_out, ho, wo)
        return last, aux

    def run(self, x, final_fmt, keep=(), aux_last=False, quant=None):
        """x: Act.  final_fmt: format of the last step's output (F32_NCHW, U8_HWC or
        planar).  keep: indices of intermediate tensors to return as well.
        aux_last: also return the last output as fp32 NCHW (alongside U8_HWC).
        quant: a ``_entropy.QuantRequest``; when the last step is a tensor-core layer writing
        the fp32 latent, the quantizer runs in its epilogue and ``quant.done`` is set.
        Returns (last Act or None, {index: Act}, aux tensor or None)."""
        tensors = {0: x}
        cur = x
        aux = None
        n_steps = len(self.steps)
        self.last_calls = {}
        skip_steps = 0
        for k, st in enumerate(self.steps):
            if skip_steps:                # consumed by the fused head launched before
                skip_steps -= 1
                continue
            if self._proj_match(k, cur, keep, final_fmt):
                cur, aux = self._run_proj(k, cur, final_fmt, aux_last)
                tensors[k + 2] = cur
                break
            sub_n = self._tail_match(k, cur, keep, final_fmt)
            if sub_n:
                cur, aux = self._run_tail(k, cur, sub_n, final_fmt, aux_last)
                tensors[k + 2] = cur
                break
            span = self._head_match(k, cur, keep, final_fmt)
            if span:
                down = self.steps[k + span - 1]
                w1, b1 = st.materialise(False)
                w2, b2 = down.materialise(False)
                ho, wo = O.KIND_OUT[down.kind](cur.h, cur.w)
                fmt, halo = self._consumer_layout(k + span, final_fmt)
                out = self._buffer(k + span - 1, fmt, cur.n, down.c_out, ho, wo, halo, cur.t.device)
                kw = dict(act_stem=act_code(st.pre_act), act_down=act_code(down.pre_act),
                          pad_mode=st.pad_mode)
                if span == 3:
                    mid = self.steps[k + 1]
                    kw['w_stem2'], kw['b_stem2'] = mid.materialise(False)
                    kw['act_mid'] = act_code(mid.post_act)
                call = ('head', (cur, w1, b1, w2, b2, down.c_out, out), kw)
                O.conv_head(*call[1], **call[2])
                self.last_calls[k] = call
                tensors[k + span] = out
                cur = out
                skip_steps = span - 1
                continue
            igemm = self._use_igemm(st, cur)
            if igemm and cur.fmt not in (C.FMT_F16_PLANAR, C.FMT_F16_SPLIT):
                raise C.CaeError('internal: igemm step fed a non-planar tensor')
            wdev, bias = st.materialise(igemm)
            ho, wo = O.KIND_OUT[st.kind](cur.h, cur.w)
            last = k == n_steps - 1
            fmt, halo = self._consumer_layout(k + 1, final_fmt)
            merged = st.kind == C.CONVT_S2 and 4 * st.c_out <= 16
            out = None
            aux_t = None
            convert_after = None
            if last and igemm and merged:
                # the final image layer writes uint8 HWC and / or fp32 NCHW
                if fmt == C.FMT_U8_HWC:
                    out = O.alloc_act(C.FMT_U8_HWC, cur.n, st.c_out, ho, wo, device=cur.t.device)
                if aux_last or fmt != C.FMT_U8_HWC:
                    aux_t = torch.empty((cur.n, st.c_out, ho, wo), dtype=torch.float32,
                                        device=cur.t.device)
            else:
                if igemm and merged:
                    raise NotImplementedError('a <=4-channel transposed layer in the middle of a track')
                if igemm and st.kind == C.CONVT_S2 and fmt in (C.FMT_F32_NCHW, C.FMT_U8_HWC):
                    # wide (>4 channel) image layer: pixel-shuffle epilogue writes planar fp16
                    convert_after = fmt
                    fmt = C.FMT_F16_PLANAR
                if fmt in (C.FMT_F32_NCHW, C.FMT_U8_HWC):
                    out = O.alloc_act(fmt, cur.n, st.c_out, ho, wo, device=cur.t.device)
                    if last and aux_last and fmt == C.FMT_U8_HWC:
                        aux_t = torch.empty((cur.n, st.c_out, ho, wo), dtype=torch.float32,
                                            device=cur.t.device)
                else:
                    out = self._buffer(k, fmt, cur.n, st.c_out, ho, wo, halo, cur.t.device)
            skip = tensors[st.skip] if st.skip is not None else None
            if st.gdn is not None:
                # conv -> (fp16 planar scratch) -> GDN (+ skip) -> the consumer's layout
                if out is None or out.fmt not in (C.FMT_F16_PLANAR, C.FMT_F16_SPLIT):
                    raise NotImplementedError('GDN on the last layer of a track')
                tmp = self._buffer((k, 'gdn'), C.FMT_F16_PLANAR, cur.n, st.c_out, ho, wo,
                                   C.HALO_KEEP, cur.t.device)
                call = ((st.kind, cur, wdev, st.c_out, tmp),
                        dict(igemm=igemm, bias=bias, skip=None, pre_act=C.ACT_NONE,
                             post_act=C.ACT_NONE, pad_mode=st.pad_mode, aux=None, groups=st.groups))
                O.conv(*call[0], **call[1])
                beta, gamma = st.gdn.effective()
                O.gdn(tmp, out, beta, gamma, st.gdn.inverse, skip=skip)
                if st.post_act is not None:
                    raise NotImplementedError('activation after a GDN residual add')
            else:
                q = None
                if (quant is not None and last and igemm and skip is None and out is not None
                        and out.fmt == C.FMT_F32_NCHW and convert_after is None):
                    q = quant.prepare(cur.n, st.c_out, ho, wo, cur.t.device)
                call = ((st.kind, cur, wdev, st.c_out, out),
                        dict(igemm=igemm, bias=bias, skip=skip, pre_act=act_code(st.pre_act),
                             post_act=act_code(st.post_act), pad_mode=st.pad_mode, aux=aux_t,
                             quant=q, groups=st.groups))
                O.conv(*call[0], **call[1])
                if q is not None:
                    quant.done = True
                    # the recorded call must stay replayable after the request's tensors are gone
                    call = (call[0], dict(call[1], quant=None))
            self.last_calls[k] = call
            if out is None:
                out = O.Act(aux_t, C.FMT_F32_NCHW, cur.n, st.c_out, ho, wo)
            if convert_after is not None:
                aux_t = O.planar_to_nchw(out)
                out = O.Act(aux_t, C.FMT_F32_NCHW, cur.n, st.c_out, ho, wo)
            tensors[k + 1] = out
            cur = out
            if last:
                aux = aux_t
        kept = {i: tensors[i] for i in keep if i in tensors}
        return cur, kept, aux
