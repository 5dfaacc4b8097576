"""Evaluation metrics of the reference's ``src/test_cae.py`` for the codec path, with the sums
computed on the device:

* ``rmse`` / ``psnr``: ``compute_rmse`` :57-58, ``compute_psnr`` :60-63 —
  ``20 log10(255) - 10 log10(mean((x - x_r)^2))``.  The reference subtracts two uint8 arrays
  without widening (the differences wrap); here the squared differences are
  summed exactly in 64-bit integers by ``cae_sse_u8``.
* ``bpp``: ``compute_rate`` :71-73 — ``8 * bytes_stored / (H * W)``.
* ``ssim``: ``compute_ssim`` :52-54 — ``skimage.metrics.structural_similarity(x, x_r,
  channel_axis=2)`` with its defaults (7x7 uniform window, sample covariance, data range 255, map
  cropped by 3 pixels), window sums on the device (``cae_ssim_u8``).
* ``delta_cielab``: ``compute_deltaCIELAB`` :21-44 — mean CIE76 distance after
  ``skimage.color.rgb2lab`` of both images (``cae_delta_e_u8``).
* ``ms_ssim``: ``compute_ms_ssim`` :46-50 — ``pytorch_msssim.ms_ssim(x_r, x, data_range=255)``
  with its defaults (11-tap Gaussian window, sigma 1.5, five scales): window moments, average
  pooling and the u8 -> plane conversion are kernels (``cae_ssim_gauss_planes_f32``,
  ``cae_avgpool2_planes_f32``, ``cae_u8_to_planes_f32``); the weighted product of the five scale
  means is a handful of scalar torch ops.

The reference computes these on the host after downloading the whole reconstruction; here the
reconstruction can stay where the synthesis transform left it.  Inputs are uint8 CUDA tensors
(or ndarrays, which are uploaded); there is no CPU implementation.
"""
import ctypes
import math

import numpy as np
import torch

from . import _cabi as C


def _as_cuda_u8(a, device):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    if a.dtype != torch.uint8:
        raise TypeError('expected uint8 images')
    if not a.is_cuda:
        if device is None:
            if not torch.cuda.is_available():
                raise C.CaeError('metrics run on the GPU only (no CPU fallback)')
            device = torch.device('cuda', torch.cuda.current_device())
        a = a.to(device, non_blocking=True)
    return a.contiguous()


def sse_u8(x, x_r, per_image=False):
    """Sum of squared differences of two uint8 arrays of the same shape (exact, int64).
    ``per_image``: one sum per leading index (N x ... inputs) instead of the total."""
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else (
        x_r.device if isinstance(x_r, torch.Tensor) and x_r.is_cuda else None)
    a, b = _as_cuda_u8(x, dev), _as_cuda_u8(x_r, dev)
    if a.shape != b.shape:
        raise ValueError('shape mismatch %r vs %r' % (tuple(a.shape), tuple(b.shape)))
    n = a.shape[0] if per_image else 1
    per = a.numel() // max(n, 1)
    out = torch.zeros(n, dtype=torch.int64, device=a.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
    with torch.cuda.device(a.device):
        for n0 in range(0, n, 65535):
            m = min(65535, n - n0)
            C.check(C.lib().cae_sse_u8(a.data_ptr() + n0 * per, b.data_ptr() + n0 * per, m, per,
                                       out[n0:].data_ptr(), stream))
    return out if per_image else out[0]


def mse(x, x_r):
    n = x.numel() if isinstance(x, torch.Tensor) else x.size
    return float(sse_u8(x, x_r).item()) / n


def rmse(x, x_r):
    """``compute_rmse`` (test_cae.py:57-58)."""
    return math.sqrt(mse(x, x_r))


def psnr(x, x_r, max_val=255.0):
    """``compute_psnr`` (test_cae.py:60-63), with the widening the reference lacks."""
    m = mse(x, x_r)
    return float('inf') if m == 0 else 20.0 * math.log10(max_val) - 10.0 * math.log10(m)


def bpp(nbytes_stored, height, width):
    """``compute_rate`` (test_cae.py:71-73)."""
    return 8.0 * float(nbytes_stored) / (height * width)


def _pair(x, x_r):
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else (
        x_r.device if isinstance(x_r, torch.Tensor) and x_r.is_cuda else None)
    a, b = _as_cuda_u8(x, dev), _as_cuda_u8(x_r, dev)
    if a.shape != b.shape:
        raise ValueError('shape mismatch %r vs %r' % (tuple(a.shape), tuple(b.shape)))
    if a.dim() == 3:
        a, b = a[None], b[None]
    if a.dim() != 4:
        raise ValueError('expected H x W x C or N x H x W x C uint8 images')
    return a, b


def ssim(x, x_r, per_image=False):
    """``compute_ssim`` (test_cae.py:52-54) of uint8 H x W x C images (or a batch N x H x W x C:
    the mean over the batch unless ``per_image``)."""
    a, b = _pair(x, x_r)
    n, h, w, c = a.shape
    out = torch.zeros(n, dtype=torch.float64, device=a.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
    with torch.cuda.device(a.device):
        C.check(C.lib().cae_ssim_u8(a.data_ptr(), b.data_ptr(), n, h, w, c, out.data_ptr(), stream))
    out = out / float((h - 6) * (w - 6) * c)
    return out if per_image else float(out.mean().item())


def delta_cielab(x, x_r, per_image=False):
    """``compute_deltaCIELAB`` (test_cae.py:21-44): mean CIE76 distance of 8-bit sRGB images."""
    a, b = _pair(x, x_r)
    n, h, w, c = a.shape
    if c != 3:
        raise ValueError('delta_cielab needs RGB images')
    out = torch.zeros(n, dtype=torch.float64, device=a.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
    with torch.cuda.device(a.device):
        C.check(C.lib().cae_delta_e_u8(a.data_ptr(), b.data_ptr(), n, h * w, out.data_ptr(), stream))
    out = out / float(h * w)
    return out if per_image else float(out.mean().item())


MS_SSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def ms_ssim(x, x_r, per_image=False):
    """``compute_ms_ssim`` (test_cae.py:46-50) of uint8 H x W x C images (or a batch)."""
    a, b = _pair(x, x_r)
    n, h, w, c = a.shape
    levels = len(MS_SSIM_WEIGHTS)
    if min(h, w) <= (11 - 1) * 2 ** (levels - 1):
        raise ValueError('ms_ssim needs images larger than %d pixels on the short side' % ((11 - 1) * 2 ** (levels - 1)))
    L = C.lib()
    dev = a.device
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    planes = n * c
    with torch.cuda.device(dev):
        pa = torch.empty((planes, h, w), dtype=torch.float32, device=dev)
        pb = torch.empty_like(pa)
        C.check(L.cae_u8_to_planes_f32(a.data_ptr(), n, h, w, c, pa.data_ptr(), stream))
        C.check(L.cae_u8_to_planes_f32(b.data_ptr(), n, h, w, c, pb.data_ptr(), stream))
        vals = []
        for lvl in range(levels):
            hh, ww = pa.shape[1], pa.shape[2]
            s_ssim = torch.zeros(planes, dtype=torch.float64, device=dev)
            s_cs = torch.zeros(planes, dtype=torch.float64, device=dev)
            C.check(L.cae_ssim_gauss_planes_f32(pa.data_ptr(), pb.data_ptr(), planes, hh, ww, 255.0,
                                                s_ssim.data_ptr(), s_cs.data_ptr(), stream))
            count = float((hh - 10) * (ww - 10))
            if lvl < levels - 1:
                vals.append(torch.relu(s_cs / count))
                oh, ow = (hh + 2 * (hh & 1) - 2) // 2 + 1, (ww + 2 * (ww & 1) - 2) // 2 + 1
                na = torch.empty((planes, oh, ow), dtype=torch.float32, device=dev)
                nb = torch.empty_like(na)
                C.check(L.cae_avgpool2_planes_f32(pa.data_ptr(), planes, hh, ww, na.data_ptr(), stream))
                C.check(L.cae_avgpool2_planes_f32(pb.data_ptr(), planes, hh, ww, nb.data_ptr(), stream))
                pa, pb = na, nb
            else:
                vals.append(torch.relu(s_ssim / count))
        wts = torch.tensor(MS_SSIM_WEIGHTS, dtype=torch.float64, device=dev).view(-1, 1)
        ms = torch.prod(torch.stack(vals) ** wts, dim=0).view(n, c).mean(dim=1)
    return ms if per_image else float(ms.mean().item())
