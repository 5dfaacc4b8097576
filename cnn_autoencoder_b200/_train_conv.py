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
    """Layers the kernels cover: up to 128 channels on either side.  Thin sides (the 3-channel
    image) ride the same tensor-core kernels with their channels padded to 16 -- those layers
    are bound by the wide tensor's HBM traffic either way."""
    return (1 <= st.c_in <= 128 and 1 <= st.c_out <= 128 and st.groups == 1 and st.bn is None
            and st.gdn is None and st.skip is None and st.post_act is None)


def _merged(kind, c_out):
    """cae_conv_igemm runs a transposed stride-2 layer with <= 4 output channels as the merged
    image-layer kernel, which writes fp32 NCHW (aux) instead of the planar layout."""
    return kind == C.CONVT_S2 and 4 * c_out <= 16


def _plain(st):
    return (st.groups == 1 and st.bn is None and st.gdn is None and st.skip is None
            and st.post_act is None)


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
    for k in range(wide[0], wide[-1] + 1):
        if steps[k].pad_mode != (C.PAD_ZERO if steps[k].transposed else C.PAD_REFLECT):
            return None
        if _merged(steps[k].kind, steps[k].c_out) and k != wide[-1]:
            return None                       # an image-type layer in the middle of the run
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


class WideChain:
    """Forward / backward of steps[k0:k1] of a track on the kernels."""

    def __init__(self, steps):
        self.steps = steps
        self._bufs = {}

    def _buf(self, key, fmt, n, c, h, w, halo, device):
        full = (fmt, n, c, h, w, halo, str(device))
        b = self._bufs.get(key)
        if b is None or b[0] != full:
            b = (full, O.alloc_act(fmt, n, c, h, w, halo, device=device))
            self._bufs[key] = b
        return b[1]

    def _workspace(self, device):
        ws = self._bufs.get('ws')
        if ws is None or ws.device != device:
            ws = torch.empty(C.lib().cae_conv_wgrad_workspace_bytes(), dtype=torch.uint8, device=device)
            self._bufs['ws'] = ws
        return ws

    # ------------------------------------------------------------------ forward
    def forward(self, x, weights, biases):
        """x: fp32 NCHW (device).  Returns (y fp32 NCHW, saved)."""
        steps = self.steps
        n = x.shape[0]
        fmt, halo = _consumer_layout(steps[0])
        cur = O.nchw_to_planar(x.detach(), fmt, halo)      # (fresh: saved for the backward pass)
        xin, outs = [], []
        y = None
        for k, st in enumerate(steps):
            ho, wo = O.KIND_OUT[st.kind](cur.h, cur.w)
            last = k == len(steps) - 1
            wp = O.pack_weights(st.kind, weights[k].detach())
            b = biases[k].detach().float().contiguous() if biases[k] is not None else None
            aux = None
            if last and _merged(st.kind, st.c_out):
                aux = torch.empty((n, st.c_out, ho, wo), dtype=torch.float32, device=x.device)
                out = None
            elif last and st.kind != C.CONVT_S2:
                out = O.alloc_act(C.FMT_F32_NCHW, n, st.c_out, ho, wo, device=x.device)
            else:
                f2, h2 = _consumer_layout(steps[k + 1]) if not last else (C.FMT_F16_PLANAR, C.HALO_KEEP)
                # fresh buffers: they are saved for the backward pass of THIS call
                out = O.alloc_act(f2, n, st.c_out, ho, wo, h2, device=x.device)
            O.conv(st.kind, cur, wp, st.c_out, out, igemm=True, bias=b, skip=None,
                   pre_act=E.act_code(st.pre_act), post_act=C.ACT_NONE, pad_mode=st.pad_mode, aux=aux)
            if aux is not None:
                out = O.Act(aux, C.FMT_F32_NCHW, n, st.c_out, ho, wo)
            xin.append(cur)
            outs.append(out)
            cur = out
        y = cur.t if cur.fmt == C.FMT_F32_NCHW else O.planar_to_nchw(cur)
        return y, (xin, outs)

    # ----------------------------------------------------------------- backward
    def _act_grad(self, g, g_dims, g_off, fold, fold_shift, out, act, dz, dz_dims, dz_off, n, h, w,
                  c, scale, db):
        C.check(C.lib().cae_act_grad(
            g, g_dims[0], g_dims[1], g_off[0], g_off[1], fold, fold_shift,
            out.desc() if out is not None else _NONE,
            out.h if out is not None else 0, out.w if out is not None else 0, act,
            dz, dz_dims[0], dz_dims[1], dz_off[0], dz_off[1], n, h, w, c,
            scale.data_ptr() if scale is not None else None,
            db.data_ptr() if db is not None else None, _sptr()))

    def backward(self, saved, g, weights, biases, need_dx=True):
        """g: fp32 NCHW gradient of the chain's output.  Returns (dx fp32 NCHW or None,
        [dW...], [db...]) with dW / db in torch layout, fp32."""
        steps = self.steps
        xin, outs = saved
        dev = g.device
        n = g.shape[0]
        g = g.contiguous().float()
        # power-of-two loss scale: max |g| * s ~ 64 (fp16 head-room for the growth through the
        # layers, sub-normals 9 decades below)
        amax = g.abs().amax().clamp_min(1e-30)
        scale = torch.exp2(torch.floor(torch.log2(64.0 / amax))).reshape(1).float()
        inv_scale = (1.0 / scale).float()
        dWs, dbs = [None] * len(steps), [None] * len(steps)
        g_desc = C.Tensor(g.data_ptr(), C.FMT_F32_NCHW, 0, 0, 0)
        g_dims, g_off, fold, fold_shift = (g.shape[2], g.shape[3]), (0, 0), 0, 0
        keep = [g]
        for k in range(len(steps) - 1, -1, -1):
            st = steps[k]
            x, o = xin[k], outs[k]
            ho, wo = o.h, o.w
            conv_like = not st.transposed
            if st.kind == C.CONV_S1:
                dz = self._buf(('dz', k), C.FMT_F16_PLANAR, n, st.c_out, ho + 2, wo + 2, C.HALO_KEEP, dev)
                dz_off, embed = (1, 1), 1
            elif st.kind == C.CONV_S2:
                dz = self._buf(('dz', k), C.FMT_F16_PLANAR, n, st.c_out, ho + 1, wo + 1, C.HALO_KEEP, dev)
                dz_off, embed = (1, 1), 1
            elif st.kind == C.CONVT_S1:
                dz = self._buf(('dz', k), C.FMT_F16_PLANAR, n, st.c_out, ho, wo, C.HALO_KEEP, dev)
                dz_off, embed = (0, 0), 0
            else:
                dz = self._buf(('dz', k), C.FMT_F16_SPLIT, n, st.c_out, ho, wo, C.HALO_KEEP, dev)
                dz_off, embed = (0, 0), 0
            last = k == len(steps) - 1
            db = torch.zeros(st.c_out, dtype=torch.float32, device=dev) if biases[k] is not None else None
            self._act_grad(g_desc, g_dims, g_off, fold, fold_shift, o, E.act_code(st.pre_act),
                           dz.desc(), (dz.h, dz.w), dz_off, n, ho, wo, st.c_out,
                           scale if last else None, db)
            if db is not None and not last:
                db.mul_(inv_scale)            # below the top layer g already carries the scale
            dbs[k] = db
            dW = torch.zeros_like(weights[k], dtype=torch.float32)
            ws = self._workspace(dev)
            C.check(C.lib().cae_conv_wgrad(st.kind, n, x.h, x.w, st.c_in, st.c_out, x.desc(),
                                           dz.desc(), embed, dW.data_ptr(), inv_scale.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _sptr()))
            dWs[k] = dW
            if k == 0 and not need_dx:
                g_desc = None
                break
            kt = _ADJOINT[st.kind]
            wt = O.pack_weights(kt, weights[k].detach())
            gh, gw = O.KIND_OUT[kt](dz.h, dz.w)
            if _merged(kt, st.c_in):
                # adjoint of a stride-2 Conv2d with a thin input: the merged kernel, fp32 NCHW out
                gt = torch.empty((n, st.c_in, gh, gw), dtype=torch.float32, device=dev)
                O.conv(kt, dz, wt, st.c_in, None, igemm=True, bias=None, skip=None,
                       pre_act=C.ACT_NONE, post_act=C.ACT_NONE, pad_mode=C.PAD_ZERO, aux=gt)
                g_desc = C.Tensor(gt.data_ptr(), C.FMT_F32_NCHW, 0, 0, 0)
                keep.append(gt)
            else:
                gbuf = self._buf(('g', k), C.FMT_F16_PLANAR, n, st.c_in, gh, gw, C.HALO_KEEP, dev)
                O.conv(kt, dz, wt, st.c_in, gbuf, igemm=True, bias=None, skip=None,
                       pre_act=C.ACT_NONE, post_act=C.ACT_NONE, pad_mode=C.PAD_ZERO)
                g_desc = gbuf.desc()
            g_dims, g_off = (gh, gw), (0, 0)
            fold, fold_shift = (1, 0 if st.kind == C.CONV_S1 else 1) if conv_like else (0, 0)
            keep.append(wt)
        dx = None
        if g_desc is not None:
            x0 = xin[0]
            dx = torch.empty((n, x0.c, x0.h, x0.w), dtype=torch.float32, device=dev)
            self._act_grad(g_desc, g_dims, g_off, fold, fold_shift, None, C.ACT_NONE,
                           C.Tensor(dx.data_ptr(), C.FMT_F32_NCHW, 0, 0, 0), (x0.h, x0.w), (0, 0),
                           n, x0.h, x0.w, x0.c, inv_scale, None)
        return dx, dWs, dbs


class _ChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, chain, n_layers, *params):
        weights, biases = params[:n_layers], params[n_layers:]
        y, saved = chain.forward(x, weights, biases)
        ctx.chain, ctx.saved = chain, saved
        ctx.n_layers = n_layers
        ctx.params = params
        ctx.need_dx = x.requires_grad
        return y

    @staticmethod
    def backward(ctx, g):
        n = ctx.n_layers
        weights, biases = ctx.params[:n], ctx.params[n:]
        dx, dWs, dbs = ctx.chain.backward(ctx.saved, g, weights, biases, need_dx=ctx.need_dx)
        grads = [dw.to(w.dtype) for dw, w in zip(dWs, weights)]
        grads += [db.to(b.dtype) if b is not None else None for db, b in zip(dbs, biases)]
        return (dx, None, None, *grads)


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
            track._train_chain = info = (steps, k0, k1, WideChain(steps[k0:k1]))
    if info is False:
        return None
    steps, k0, k1, chain = info
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
