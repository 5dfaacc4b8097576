"""Thin torch-tensor front end over the C ABI: device buffers in the internal
layouts and one Python call per ABI entry point.  PyTorch is used for device
memory and streams only; every computation below is a kernel of
``libcae_b200.so``.
"""
import ctypes
from dataclasses import dataclass

import torch

from . import _cabi as C


def _stream_ptr(t=None):
    """Current stream of the current device.  The library launches on the CURRENT device, so a
    tensor living on another one is refused here instead of faulting inside a kernel."""
    if t is not None and t.is_cuda and t.device.index != torch.cuda.current_device():
        raise C.CaeError(f'tensor on {t.device} but the current CUDA device is '
                         f'cuda:{torch.cuda.current_device()}: wrap the call in torch.cuda.device(...)')
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise C.CaeError(f'{name} must be a CUDA tensor: the hot path has no CPU fallback')


COL_PAD = 3     # CAE_COL_PAD of include/cae_b200.h: unused units left / right of the halo


def planes_for(c):
    """8-channel planes holding ``c`` channels padded to a multiple of 16 (MMA K step)."""
    return ((c + 15) // 16) * 2


@dataclass
class Act:
    """An activation tensor in one of the ABI formats."""
    t: torch.Tensor
    fmt: int
    n: int
    c: int
    h: int
    w: int
    halo: int = C.HALO_KEEP

    @property
    def planes(self):
        return self.t.shape[-4] if self.fmt in (C.FMT_F16_PLANAR, C.FMT_F16_SPLIT) else 0

    def desc(self, halo=None):
        return C.Tensor(self.t.data_ptr(), self.fmt, self.planes,
                        self.halo if halo is None else halo, 0)


def alloc_act(fmt, n, c, h, w, halo=C.HALO_KEEP, device='cuda', planes=None):
    """Zero-initialised buffer (the zero halo IS the zero padding of the transposed
    convolutions, so it is cleared once here and never written with halo KEEP)."""
    if fmt == C.FMT_F16_PLANAR:
        p = planes or planes_for(c)
        t = torch.zeros((n, p, h + 2, w + 2 + 2 * COL_PAD, 8), dtype=torch.float16, device=device)
    elif fmt == C.FMT_F16_SPLIT:
        if h % 2 or w % 2:
            raise C.CaeError(f'split layout needs even spatial size, got {h}x{w}')
        p = planes or planes_for(c)
        t = torch.zeros((n, 4, p, (h + 2) // 2, (w + 2 + 2 * COL_PAD) // 2, 8), dtype=torch.float16,
                        device=device)
    elif fmt == C.FMT_F32_NCHW:
        t = torch.empty((n, c, h, w), dtype=torch.float32, device=device)
    elif fmt == C.FMT_U8_HWC:
        t = torch.empty((n, h, w, c), dtype=torch.uint8, device=device)
    else:
        raise C.CaeError(f'bad format {fmt}')
    return Act(t, fmt, n, c, h, w, halo)


def wrap_nchw(x):
    _require_cuda(x, 'input')
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    n, c, h, w = x.shape
    return Act(x, C.FMT_F32_NCHW, n, c, h, w)


def wrap_u8_hwc(x):
    _require_cuda(x, 'input')
    x = x.contiguous()
    if x.dtype != torch.uint8 or x.dim() != 4:
        raise C.CaeError('expected an N x H x W x C uint8 tensor')
    n, h, w, c = x.shape
    return Act(x, C.FMT_U8_HWC, n, c, h, w)


KIND_OUT = {C.CONV_S1: lambda h, w: (h, w), C.CONV_S2: lambda h, w: ((h - 1) // 2 + 1, (w - 1) // 2 + 1),
            C.CONVT_S1: lambda h, w: (h, w), C.CONVT_S2: lambda h, w: (2 * h, 2 * w)}


def pack_weights(kind, weight, scale=None, ck=0, out=None):
    """fp32 torch-layout weight (device) -> packed fp16 image for the igemm kernel (``out``: a
    buffer from an earlier call with the same shape, refilled in place)."""
    _require_cuda(weight, 'weight')
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.contiguous().float()
    transposed = kind in (C.CONVT_S1, C.CONVT_S2)
    c_in, c_out = (w.shape[0], w.shape[1]) if transposed else (w.shape[1], w.shape[0])
    L = C.lib()
    nbytes = L.cae_packed_weight_bytes(kind, c_in, c_out, ck)
    packed = out if out is not None else torch.empty(nbytes // 2, dtype=torch.float16, device=w.device)
    if packed.numel() * 2 != nbytes:
        raise C.CaeError('pack_weights: the buffer given does not match this layer')
    sc = None
    if scale is not None:
        sc = scale.detach().contiguous().float()
    C.check(L.cae_pack_weights(kind, c_in, c_out, ck, w.data_ptr(),
                               sc.data_ptr() if sc is not None else None,
                               packed.data_ptr(), _stream_ptr(w)))
    return packed


def pack_proj_weights(weight, scale=None):
    """fp32 ConvTranspose2d weight (128, c, 3, 3) of the image layer (device) -> the packed
    projection operand of ``cae_conv_desc.proj`` (``cae_pack_proj_weights``)."""
    _require_cuda(weight, 'weight')
    w = weight.detach().contiguous().float()
    L = C.lib()
    packed = torch.empty(L.cae_proj_weight_bytes() // 2, dtype=torch.float16, device=w.device)
    sc = scale.detach().contiguous().float() if scale is not None else None
    C.check(L.cae_pack_proj_weights(w.shape[0], w.shape[1], w.data_ptr(),
                                    sc.data_ptr() if sc is not None else None,
                                    packed.data_ptr(), _stream_ptr(w)))
    return packed


def alloc_proj(n, h, w, device):
    """Record buffer of the projection fusion for an n x h x w intermediate tensor."""
    return torch.empty(C.lib().cae_proj_bytes(n, h, w) // 2, dtype=torch.float16, device=device)


def image_from_proj(proj, n, h, w, c_out, *, bias=None, pre_act=C.ACT_NONE, post_act=C.ACT_NONE,
                    out=None, aux=None):
    """Second half of the projection fusion (``cae_image_from_proj``): records of the h x w
    intermediate tensor -> uint8 HWC ``out`` (an Act) and / or fp32 NCHW ``aux``, both 2h x 2w."""
    C.check(C.lib().cae_image_from_proj(
        proj.data_ptr(), n, h, w, c_out, bias.data_ptr() if bias is not None else None,
        pre_act, post_act, out.t.data_ptr() if out is not None else None,
        aux.data_ptr() if aux is not None else None, _stream_ptr(proj)))
    return out


def conv(kind, x, weights, c_out, out, *, igemm, bias=None, skip=None, pre_act=C.ACT_NONE,
         post_act=C.ACT_NONE, pad_mode=C.PAD_REFLECT, aux=None, ck=0, mt=0, grid=0, quant=None,
         groups=1, proj=None):
    """out = post_act(pre_act(conv(x) + bias) + skip) through the C ABI.  ``quant``: a
    ``_cabi.QuantFuse`` for the latent layer (igemm, fp32 NCHW output).  ``proj``: (packed
    projection weights, record buffer) -- the layer writes the records of the projection fusion
    instead of ``out`` (which is then None)."""
    d = C.ConvDesc()
    d.kind = kind
    d.n, d.h_in, d.w_in, d.c_in, d.c_out = x.n, x.h, x.w, x.c, c_out
    d.inp = x.desc()
    d.out = out.desc() if out is not None else C.Tensor(None, C.FMT_NONE, 0, 0, 0)
    d.skip = skip.desc() if skip is not None else C.Tensor(None, C.FMT_NONE, 0, 0, 0)
    d.weights = weights.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.pre_act, d.post_act, d.pad_mode = pre_act, post_act, pad_mode
    d.ck, d.mt, d.grid = ck, mt, grid
    d.aux_out = aux.data_ptr() if aux is not None else None
    d.quant = ctypes.addressof(quant) if quant is not None else None
    d.groups = groups
    pf = None
    if proj is not None:
        pf = C.ProjFuse(proj[0].data_ptr(), proj[1].data_ptr())
        d.proj = ctypes.addressof(pf)
    L = C.lib()
    fn = L.cae_conv_igemm if igemm else L.cae_conv_direct
    C.check(fn(ctypes.byref(d), _stream_ptr(x.t)))
    return out


def conv_head(x, w_stem, b_stem, w_down, b_down, c_out, out, *, act_stem=C.ACT_NONE,
              act_down=C.ACT_NONE, pad_mode=C.PAD_REFLECT, w_stem2=None, b_stem2=None,
              act_mid=C.ACT_NONE):
    """out = act_down(conv_s2(act_stem(conv_s1(x) + b_stem)) + b_down): the fused first
    downsampling unit (``cae_conv_head``).  x: U8_HWC / F32_NCHW Act, out: planar Act.
    With ``w_stem2`` the residual form: the stride-2 convolution reads
    ``act_mid(conv_s1_b(act_stem(conv_s1_a(x) + b_stem)) + b_stem2 + x)``."""
    d = C.HeadDesc()
    d.n, d.h_in, d.w_in, d.c_in, d.c_out = x.n, x.h, x.w, x.c, c_out
    d.inp = x.desc()
    d.out = out.desc()
    d.w_stem, d.b_stem = w_stem.data_ptr(), (b_stem.data_ptr() if b_stem is not None else None)
    d.w_down, d.b_down = w_down.data_ptr(), (b_down.data_ptr() if b_down is not None else None)
    d.act_stem, d.act_down, d.pad_mode = act_stem, act_down, pad_mode
    if w_stem2 is not None:
        d.residual = 1
        d.w_stem2 = w_stem2.data_ptr()
        d.b_stem2 = b_stem2.data_ptr() if b_stem2 is not None else None
        d.act_mid = act_mid
    C.check(C.lib().cae_conv_head(ctypes.byref(d), _stream_ptr(x.t)))
    return out


def replay(call):
    """Re-issue a call recorded by ``TrackExecutor.last_calls`` (benchmarks / profiling)."""
    if call[0] == 'head':
        return conv_head(*call[1], **call[2])
    if call[0] == 'image_from_proj':
        return image_from_proj(*call[1], **call[2])
    return conv(*call[0], **call[1])


def nchw_to_planar(x, fmt=C.FMT_F16_PLANAR, halo=C.HALO_KEEP, out=None):
    """fp32 NCHW -> internal layout.  ``out``: a buffer of the right shape to reuse (its zero
    halo persists: only the interior, and the reflect halo if asked for, are rewritten)."""
    a = wrap_nchw(x)
    if out is None:
        out = alloc_act(fmt, a.n, a.c, a.h, a.w, halo, device=x.device)
    C.check(C.lib().cae_nchw_to_planar(a.t.data_ptr(), a.n, a.c, a.h, a.w, out.desc(),
                                       _stream_ptr(a.t)))
    return out


def planar_to_nchw(a):
    out = torch.empty((a.n, a.c, a.h, a.w), dtype=torch.float32, device=a.t.device)
    C.check(C.lib().cae_planar_to_nchw(a.desc(), a.n, a.c, a.h, a.w, out.data_ptr(),
                                       _stream_ptr(a.t)))
    return out


def gdn(x, out, beta, gamma, inverse, skip=None):
    """out = GDN(x) (+ skip) through ``cae_gdn``; x / out / skip are planar fp16 Acts."""
    none = C.Tensor(None, C.FMT_NONE, 0, 0, 0)
    C.check(C.lib().cae_gdn(x.desc(), out.desc(), skip.desc() if skip is not None else none,
                            x.n, x.h, x.w, x.c, beta.data_ptr(), gamma.data_ptr(),
                            1 if inverse else 0, _stream_ptr(x.t)))
    return out
