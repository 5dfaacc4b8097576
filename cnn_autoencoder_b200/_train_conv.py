"""Training-mode forward / backward of the transforms' wide layers on this repo's kernels.

The reference trains ``Analyzer`` / ``Synthesizer`` through torch autograd
(``/root/reference/src/train_cae_ms.py:209-219`` -> ``loss.backward()`` through the
``nn.Conv2d`` / ``nn.ConvTranspose2d`` units of ``src/models/tasks/_autoencoders.py:53-304``).
Here a maximal run of consecutive wide layers (16..128 channels on both sides, dense, no
BatchNorm / GDN / residual add) of a track is ONE ``torch.autograd.Function``:

  forward   the inference kernels (``cae_conv_igemm``: bias + activation in the epilogue,
            fp16 activations in the planar / split layouts), every intermediate tensor kept;
  backward  per layer, last to first:
              dz  = fold(g) * act'(out)                       ``cae_act_grad`` (+ bias gradient)
              dW += dz (x) window(x)                          ``cae_conv_wgrad`` (tcgen05, fp32)
              g   = adjoint convolution of dz                 ``cae_conv_igemm``, transposed kind

The adjoint of a Conv2d is the ConvTranspose2d with the same weight tensor and vice versa (torch
stores them as (out, in, 3, 3) resp. (in, out, 3, 3)), so the data gradient needs no new kernel.
``padding_mode='reflect'`` is handled by embedding dz in a zero ring: the transposed convolution
then produces the gradient of the PADDED input and ``cae_act_grad`` of the layer below folds the
mirrored ring back onto the interior.  Gradients travel in fp16 with a power-of-two loss scale
taken from max|g| (device side, no host synchronisation) and leave in fp32.

The thin layers either side (3-channel stem, image layer: < 2 % of the FLOPs) stay torch ops, so
autograd stitches the pieces together; tracks with features the kernels do not cover fall back to
the torch formulation entirely (``eligible_chain``).
"""
import ctypes
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi as C
from . import _engine as E
from . import _ops as O

_ADJOINT = {C.CONV_S1: C.CONVT_S1, C.CONV_S2: C.CONVT_S2, C.CONVT_S1: C.CONV_S1,
            C.CONVT_S2: C.CONV_S2}
_NONE = C.Tensor(None, C.FMT_NONE, 0, 0, 0)


def _wide(st):
    """Layers the kernels cover: up to 256 channels on either side, plain or with the residual
    add of the residual units (LeakyReLU or nothing after the add: ReLU there is not invertible,
    and the branch's sign is recovered from the saved sum).  Thin sides (the 3-channel image) ride
    the same tensor-core kernels with their channels padded to 16 -- those layers are bound by
    the wide tensor's HBM traffic either way."""
    if not (1 <= st.c_in <= 256 and 1 <= st.c_out <= 256 and st.groups == 1 and st.bn is None
            and st.gdn is None):
        return False
    if st.skip is None:
        return st.post_act is None
    return st.post_act in (None, 'LeakyReLU') and st.kind in (C.CONV_S1, C.CONVT_S1) and st.c_in == st.c_out


def _merged(kind, c_out):
    """cae_conv_igemm runs a transposed stride-2 layer with <= 4 output channels as the merged
    image-layer kernel, which writes fp32 NCHW (aux) instead of the planar layout."""
    return kind == C.CONVT_S2 and 4 * c_out <= 16


def _plain(st):
    return st.groups == 1 and st.bn is None and st.gdn is None and (st.skip is not None or st.post_act is None)


def eligible_chain(track):
    """(steps, k0, k1) when the track's layers are plain convolutions + LeakyReLU / ReLU and
    steps[k0:k1] is its run of wide layers (at least one), else None."""
    units = track._units()
    if not units or any(isinstance(m, (nn.Dropout, nn.Dropout2d, nn.BatchNorm2d))
                        for u in units for m in u.modules()):
        return None
    try:
        steps = track._executor().steps
    except NotImplementedError:
        return None
    if not all(_plain(s) for s in steps):
        return None
    wide = [k for k, s in enumerate(steps) if _wide(s)]
    if not wide or wide != list(range(wide[0], wide[-1] + 1)):
        return None
    if any(s.skip is not None for k, s in enumerate(steps) if k < wide[0] or k > wide[-1]):
        return None
    for k in range(wide[0], wide[-1] + 1):
        if steps[k].pad_mode != (C.PAD_ZERO if steps[k].transposed else C.PAD_REFLECT):
            return None
        if _merged(steps[k].kind, steps[k].c_out) and k != wide[-1]:
            return None                       # an image-type layer in the middle of the run
        if steps[k].skip is not None and steps[k].skip < wide[0]:
            return None                       # residual source outside the run
    return steps, wide[0], wide[-1] + 1


def torch_step(st, x):
    """One Step as torch ops (the thin layers around a chain)."""
    conv = st.conv
    if st.transposed:
        y = F.conv_transpose2d(x, conv.weight, conv.bias, stride=conv.stride, padding=1,
                               output_padding=conv.stride[0] - 1)
    else:
        mode = 'reflect' if conv.padding_mode == 'reflect' else 'constant'
        y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode=mode), conv.weight, conv.bias, stride=conv.stride)
    if st.pre_act == 'LeakyReLU':
        y = F.leaky_relu(y, 0.01)
    elif st.pre_act == 'ReLU':
        y = F.relu(y)
    return y


def _consumer_layout(st):
    fmt = C.FMT_F16_SPLIT if st.kind == C.CONV_S2 else C.FMT_F16_PLANAR
    halo = C.HALO_REFLECT if not st.transposed else C.HALO_KEEP
    return fmt, halo


def _sptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Plan:
    """Everything one input shape needs, allocated once: the activations of the forward pass
    (kept for the backward pass), packed weights (refilled every step), gradient buffers, and
    the two CUDA graphs that replay the launches."""

    def __init__(self, chain, shape, device, weights, biases):
        steps = chain.steps
        n, c, h, w = shape
        self.shape, self.device = tuple(shape), device
        self.busy = False
        fmt, halo = _consumer_layout(steps[0])
        self.x0 = O.alloc_act(fmt, n, c, h, w, halo, device=device)
        self.xin, self.outs, self.wp, self.wt = [], [], [], []
        self.dz, self.embed, self.dz_off, self.gbuf = [], [], [], []
        self.gsum = []                 # residual layers: the gradient handed on to the skip source
        cur = self.x0
        for k, st in enumerate(steps):
            ho, wo = O.KIND_OUT[st.kind](cur.h, cur.w)
            last = k == len(steps) - 1
            if last and _merged(st.kind, st.c_out):
                out = O.Act(torch.empty((n, st.c_out, ho, wo), dtype=torch.float32, device=device),
                            C.FMT_F32_NCHW, n, st.c_out, ho, wo)
            elif last and st.kind != C.CONVT_S2:
                out = O.alloc_act(C.FMT_F32_NCHW, n, st.c_out, ho, wo, device=device)
            else:
                f2, h2 = _consumer_layout(steps[k + 1]) if not last else (C.FMT_F16_PLANAR, C.HALO_KEEP)
                out = O.alloc_act(f2, n, st.c_out, ho, wo, h2, device=device)
            self.xin.append(cur)
            self.outs.append(out)
            self.wp.append(O.pack_weights(st.kind, weights[k]))
            self.wt.append(O.pack_weights(_ADJOINT[st.kind], weights[k]))
            # gradient of the pre-activation output, in the layout its two consumers want
            if st.kind == C.CONV_S1:
                dz, off, emb = O.alloc_act(C.FMT_F16_PLANAR, n, st.c_out, ho + 2, wo + 2, device=device), (1, 1), 1
            elif st.kind == C.CONV_S2:
                dz, off, emb = O.alloc_act(C.FMT_F16_PLANAR, n, st.c_out, ho + 1, wo + 1, device=device), (1, 1), 1
            elif st.kind == C.CONVT_S1:
                dz, off, emb = O.alloc_act(C.FMT_F16_PLANAR, n, st.c_out, ho, wo, device=device), (0, 0), 0
            else:
                dz, off, emb = O.alloc_act(C.FMT_F16_SPLIT, n, st.c_out, ho, wo, device=device), (0, 0), 0
            self.dz.append(dz)
            self.dz_off.append(off)
            self.embed.append(emb)
            kt = _ADJOINT[st.kind]
            gh, gw = O.KIND_OUT[kt](dz.h, dz.w)
            if _merged(kt, st.c_in):      # adjoint of a thin stride-2 Conv2d: fp32 NCHW from the merged kernel
                g = O.Act(torch.empty((n, st.c_in, gh, gw), dtype=torch.float32, device=device),
                          C.FMT_F32_NCHW, n, st.c_in, gh, gw)
            else:
                g = O.alloc_act(C.FMT_F16_PLANAR, n, st.c_in, gh, gw, device=device)
            self.gbuf.append(g)
            self.gsum.append(O.alloc_act(C.FMT_F16_PLANAR, n, st.c_out, ho, wo, device=device)
                             if st.skip is not None else None)
            cur = out
        last_out = self.outs[-1]
        self.y = last_out.t if last_out.fmt == C.FMT_F32_NCHW else \
            torch.empty((n, last_out.c, last_out.h, last_out.w), dtype=torch.float32, device=device)
        # gradients: one flat fp32 buffer, [dW_0 .. dW_L-1 | db_0 ..]
        sizes = [wt.numel() for wt in weights] + [b.numel() if b is not None else 0 for b in biases]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        offs = [0]
        for sz in sizes:
            offs.append(offs[-1] + sz)
        L = len(steps)
        self.dW = [self.flat[offs[k]:offs[k + 1]].view_as(weights[k]) for k in range(L)]
        self.db = [self.flat[offs[L + k]:offs[L + k + 1]] if biases[k] is not None else None for k in range(L)]
        self.offs = offs
        self.g_in = torch.empty_like(self.y)
        self.dx = torch.empty((n, c, h, w), dtype=torch.float32, device=device)
        self.scale = torch.ones(1, dtype=torch.float32, device=device)
        self.inv_scale = torch.ones(1, dtype=torch.float32, device=device)
        self.ws = torch.empty(C.lib().cae_conv_wgrad_workspace_bytes(), dtype=torch.uint8, device=device)
        self.graph_fwd = self.graph_bwd = None
        self.fwd_ptrs = self.bwd_key = None


class WideChain:
    """Forward / backward of a run of steps of a track on the kernels."""

    def __init__(self, steps, first=0):
        self.steps = steps
        self.first = first             # track index of steps[0] (Step.skip counts track tensors)
        self._plans = {}
        self.use_graphs = not os.environ.get('CAE_TRAIN_NO_GRAPH')

    def _plan(self, x, weights, biases):
        key = (tuple(x.shape), str(x.device))
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) > 4:
                self._plans.clear()
            plan = self._plans[key] = _Plan(self, x.shape, x.device, weights, biases)
        elif plan.busy:
            # a second forward pass before the first one's backward: its own set of buffers
            plan = _Plan(self, x.shape, x.device, weights, biases)
        return plan

    @staticmethod
    def _tensor(plan, i):
        """Tensor i of the chain: 0 = its input, i = the output of step i - 1."""
        return plan.x0 if i == 0 else plan.outs[i - 1]

    # ------------------------------------------------------------------ forward
    def _launch_forward(self, plan, weights, biases):
        for k, st in enumerate(self.steps):
            O.pack_weights(st.kind, weights[k], out=plan.wp[k])
            out = plan.outs[k]
            aux = out.t if _merged(st.kind, st.c_out) else None
            skip = self._tensor(plan, st.skip - self.first) if st.skip is not None else None
            O.conv(st.kind, plan.xin[k], plan.wp[k], st.c_out, None if aux is not None else out,
                   igemm=True, bias=biases[k], skip=skip, pre_act=E.act_code(st.pre_act),
                   post_act=E.act_code(st.post_act), pad_mode=st.pad_mode, aux=aux)
        last = plan.outs[-1]
        if last.fmt != C.FMT_F32_NCHW:
            C.check(C.lib().cae_planar_to_nchw(last.desc(), last.n, last.c, last.h, last.w,
                                               plan.y.data_ptr(), _sptr()))

    def forward(self, x, weights, biases):
        """x: fp32 NCHW (device).  Returns (y fp32 NCHW, plan)."""
        weights = [w.detach() for w in weights]
        biases = [b.detach() if b is not None else None for b in biases]
        for t in weights + [b for b in biases if b is not None]:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise C.CaeError('the training kernels expect contiguous fp32 parameters')
        plan = self._plan(x, weights, biases)
        plan.busy = True
        O.nchw_to_planar(x.detach(), plan.x0.fmt, plan.x0.halo, out=plan.x0)
        # (under an outer capture -- train_step.GraphedTrainStep -- the launches go straight into
        # that graph)
        if not self.use_graphs or torch.cuda.is_current_stream_capturing():
            self._launch_forward(plan, weights, biases)
        else:
            ptrs = tuple(t.data_ptr() for t in weights) + tuple(b.data_ptr() if b is not None else 0 for b in biases)
            if plan.graph_fwd is None or plan.fwd_ptrs != ptrs:
                # warm-up launch outside the capture (kernel attributes, lazy module loading)
                self._launch_forward(plan, weights, biases)
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = C.lib().cae_launch_count()
                with torch.cuda.graph(g):
                    self._launch_forward(plan, weights, biases)
                plan.fwd_kernels = int(C.lib().cae_launch_count() - n0)
                plan.graph_fwd, plan.fwd_ptrs = g, ptrs
            plan.graph_fwd.replay()
            C.note_graph_replay(plan.fwd_kernels)
        return plan.y.clone(), plan

    # ----------------------------------------------------------------- backward
    def _act_grad(self, g, g_dims, g_off, fold, fold_shift, out, act, dz, dz_dims, dz_off, n, h, w,
                  c, scale, db, post_act=C.ACT_NONE, skip=None, gsum=None, g2=None):
        d = C.ActGradDesc()
        d.n, d.h, d.w, d.c = n, h, w, c
        d.g = g
        d.g_h, d.g_w, d.g_oy, d.g_ox = g_dims[0], g_dims[1], g_off[0], g_off[1]
        d.fold, d.fold_shift = fold, fold_shift
        d.out = out.desc() if out is not None else _NONE
        d.out_h, d.out_w = (out.h, out.w) if out is not None else (0, 0)
        d.act, d.post_act = act, post_act
        d.dz = dz
        d.dz_h, d.dz_w, d.dz_oy, d.dz_ox = dz_dims[0], dz_dims[1], dz_off[0], dz_off[1]
        d.skip = skip.desc() if skip is not None else _NONE
        d.gsum = gsum.desc() if gsum is not None else _NONE
        d.g2 = g2.desc() if g2 is not None else _NONE
        d.scale = scale.data_ptr() if scale is not None else None
        d.db = db.data_ptr() if db is not None else None
        C.check(C.lib().cae_act_grad(ctypes.byref(d), _sptr()))

    def _launch_backward(self, plan, weights, need_dx):
        steps = self.steps
        n = plan.shape[0]
        L = len(steps)
        plan.flat.zero_()
        g_desc = C.Tensor(plan.g_in.data_ptr(), C.FMT_F32_NCHW, 0, 0, 0)
        g_dims, g_off, fold, fold_shift = (plan.g_in.shape[2], plan.g_in.shape[3]), (0, 0), 0, 0
        pending = {}           # chain tensor index -> gradient arriving through a residual add
        for k in range(L - 1, -1, -1):
            st = steps[k]
            x, o, dz = plan.xin[k], plan.outs[k], plan.dz[k]
            last = k == L - 1
            skip = gsum = None
            if st.skip is not None:
                src = st.skip - self.first
                if src in pending:
                    raise C.CaeError('two residual connections from one tensor are not supported')
                skip, gsum = self._tensor(plan, src), plan.gsum[k]
                pending[src] = gsum
            self._act_grad(g_desc, g_dims, g_off, fold, fold_shift, o, E.act_code(st.pre_act),
                           dz.desc(), (dz.h, dz.w), plan.dz_off[k], n, o.h, o.w, st.c_out,
                           plan.scale if last else None, plan.db[k],
                           post_act=E.act_code(st.post_act), skip=skip, gsum=gsum,
                           g2=pending.pop(k + 1, None))
            C.check(C.lib().cae_conv_wgrad(st.kind, n, x.h, x.w, st.c_in, st.c_out, x.desc(),
                                           dz.desc(), plan.embed[k], plan.dW[k].data_ptr(),
                                           plan.inv_scale.data_ptr(), plan.ws.data_ptr(),
                                           plan.ws.numel(), _sptr()))
            if k == 0 and not need_dx:
                g_desc = None
                break
            kt = _ADJOINT[st.kind]
            O.pack_weights(kt, weights[k], out=plan.wt[k])
            gb = plan.gbuf[k]
            if gb.fmt == C.FMT_F32_NCHW:
                O.conv(kt, dz, plan.wt[k], st.c_in, None, igemm=True, bias=None, skip=None,
                       pre_act=C.ACT_NONE, post_act=C.ACT_NONE, pad_mode=C.PAD_ZERO, aux=gb.t)
            else:
                O.conv(kt, dz, plan.wt[k], st.c_in, gb, igemm=True, bias=None, skip=None,
                       pre_act=C.ACT_NONE, post_act=C.ACT_NONE, pad_mode=C.PAD_ZERO)
            g_desc = C.Tensor(gb.t.data_ptr(), gb.fmt, gb.planes, 0, 0)
            g_dims, g_off = (gb.h, gb.w), (0, 0)
            fold, fold_shift = (1, 0 if st.kind == C.CONV_S1 else 1) if not st.transposed else (0, 0)
        if g_desc is not None:
            x0 = plan.x0
            self._act_grad(g_desc, g_dims, g_off, fold, fold_shift, None, C.ACT_NONE,
                           C.Tensor(plan.dx.data_ptr(), C.FMT_F32_NCHW, 0, 0, 0), (x0.h, x0.w), (0, 0),
                           n, x0.h, x0.w, x0.c, plan.inv_scale, None, g2=pending.pop(0, None))
        # below the top layer the incoming gradient already carries the loss scale
        if L > 1 and plan.offs[2 * L] > plan.offs[L]:
            lo = plan.offs[L]
            hi = plan.offs[2 * L - 1]        # the biases of layers 0 .. L-2
            if hi > lo:
                plan.flat[lo:hi].mul_(plan.inv_scale)

    def backward(self, plan, g, weights, biases, need_dx=True):
        """g: fp32 NCHW gradient of the chain's output.  Returns (dx fp32 NCHW or None,
        [dW...], [db...]) with dW / db in torch layout, fp32."""
        weights = [w.detach() for w in weights]
        plan.g_in.copy_(g)
        # power-of-two loss scale: max |g| * s ~ 64 (fp16 head-room for the growth through the
        # layers, sub-normals 9 decades below); computed on the device, no host round trip
        amax = plan.g_in.abs().amax().clamp_min(1e-30)
        plan.scale.copy_(torch.exp2(torch.floor(torch.log2(64.0 / amax))).reshape(1))
        plan.inv_scale.copy_(1.0 / plan.scale)
        if not self.use_graphs or torch.cuda.is_current_stream_capturing():
            self._launch_backward(plan, weights, need_dx)
        else:
            ptrs = tuple(t.data_ptr() for t in weights)
            if plan.graph_bwd is None or plan.bwd_key != (need_dx, ptrs):
                self._launch_backward(plan, weights, need_dx)
                torch.cuda.current_stream().synchronize()
                gr = torch.cuda.CUDAGraph()
                n0 = C.lib().cae_launch_count()
                with torch.cuda.graph(gr):
                    self._launch_backward(plan, weights, need_dx)
                plan.bwd_kernels = int(C.lib().cae_launch_count() - n0)
                plan.graph_bwd, plan.bwd_key = gr, (need_dx, ptrs)
            plan.graph_bwd.replay()
            C.note_graph_replay(plan.bwd_kernels)
        grads = plan.flat.clone()              # (autograd may keep what it is handed)
        L = len(self.steps)
        offs = plan.offs
        dWs = [grads[offs[k]:offs[k + 1]].view_as(weights[k]) for k in range(L)]
        dbs = [grads[offs[L + k]:offs[L + k + 1]] if biases[k] is not None else None for k in range(L)]
        dx = plan.dx.clone() if need_dx else None
        plan.busy = False
        return dx, dWs, dbs


class _ChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, chain, n_layers, *params):
        weights, biases = params[:n_layers], params[n_layers:]
        y, plan = chain.forward(x, weights, biases)
        ctx.chain, ctx.plan = chain, plan
        ctx.n_layers = n_layers
        ctx.params = params
        ctx.need_dx = x.requires_grad
        if not any(ctx.needs_input_grad):
            plan.busy = False             # nobody will come back for the saved activations
        return y

    @staticmethod
    def backward(ctx, g):
        n = ctx.n_layers
        weights, biases = ctx.params[:n], ctx.params[n:]
        dx, dWs, dbs = ctx.chain.backward(ctx.plan, g, weights, biases, need_dx=ctx.need_dx)
        return (dx, None, None, *dWs, *dbs)


def run_track(track, x):
    """Training-mode forward of a whole track: torch ops for the thin layers, the kernels for
    the wide run.  Returns the list of step outputs of interest: (final tensor, {step index:
    tensor}) -- only the tensors that exist as torch tensors (chain boundaries, thin layers)."""
    info = getattr(track, '_train_chain', None)
    if info is None:
        found = eligible_chain(track)
        if found is None:
            track._train_chain = info = False
        else:
            steps, k0, k1 = found
            track._train_chain = info = (steps, k0, k1, WideChain(steps[k0:k1], k0))
    if info is False:
        return None
    steps, k0, k1, chain = info
    if any(st.conv.weight.dtype != torch.float32 or not st.conv.weight.is_cuda for st in steps[k0:k1]):
        return None                    # (half / CPU parameters: the torch formulation)
    tensors = {0: x}
    cur = x
    for k in range(k0):
        cur = torch_step(steps[k], cur)
        tensors[k + 1] = cur
    ws = [steps[k].conv.weight for k in range(k0, k1)]
    bs = [steps[k].conv.bias for k in range(k0, k1)]
    cur = _ChainFn.apply(cur, chain, k1 - k0, *ws, *bs)
    tensors[k1] = cur
    for k in range(k1, len(steps)):
        cur = torch_step(steps[k], cur)
        tensors[k + 1] = cur
    return cur, tensors
