"""Evaluation of a checkpoint on an image, mirroring ``src/test_cae.py`` of the reference:
``test_image`` compresses the image, decompresses it and scores the reconstruction with the
metrics of ``metric_fun`` (``test_cae.py:21-89``).  The reference writes the reconstruction to a
zarr group, reads both images back on the host and runs scikit-image / pytorch_msssim there; here
the reconstruction is decoded straight into an array, uploaded once together with the source, and
every sum is taken on the device (``metrics.py``).  Same metric names, same returned keys."""
import os
import shutil
import tempfile
from time import perf_counter

import numpy as np
import torch

from . import metrics
from .compress import as_yxc, compress_image, open_source
from .decompress import decompress_image


def compute_deltaCIELAB(x=None, x_r=None, **kwargs):
    return metrics.delta_cielab(x, x_r), None


def compute_ms_ssim(x=None, x_r=None, **kwargs):
    return metrics.ms_ssim(x, x_r), None


def compute_ssim(x=None, x_r=None, **kwargs):
    return metrics.ssim(x, x_r), None


def compute_psnr(x=None, x_r=None, max_val=255, **kwargs):
    return metrics.psnr(x, x_r, max_val=max_val), None


def compute_rmse(x=None, x_r=None, **kwargs):
    return metrics.rmse(x, x_r), None


def compute_rate(x=None, x_r=None, nbytes_stored=None, **kwargs):
    # the reference asks zarr for ``nbytes_stored`` of the compressed array (:71-73)
    return metrics.bpp(nbytes_stored, x.shape[0], x.shape[1]), None


metric_fun = {'dist': compute_rmse,
              'rate': compute_rate,
              'ms-ssim': compute_ms_ssim,
              'ssim': compute_ssim,
              'psnr': compute_psnr,
              'delta_cielab': compute_deltaCIELAB}


def test_image(checkpoint, input_filename, patch_size=512, source_format='zarr', data_group='0/0',
               data_axes='TCZYX', gpu=True, progress_bar=False, temp_output_filename=None, **kwargs):
    """``test_cae.py:91-163``: returns ``{metric: score, metric + '_time': seconds, ...,
    'execution_time', 'evaluation_time'}``.  ``kwargs`` go to the tile loops (``batch_tiles``,
    ``coder_tiles``, ``workers``)."""
    if not torch.cuda.is_available():
        raise RuntimeError('test_image needs a CUDA device (no CPU fallback)')
    own_tmp = temp_output_filename is None
    if own_tmp:
        base = '/dev/shm' if os.path.isdir('/dev/shm') else None
        temp_output_filename = os.path.join(tempfile.mkdtemp(prefix='cae_eval_', dir=base), 'temp.zarr')
    try:
        src = as_yxc(open_source(input_filename, data_group), data_axes)
        H, W, c = src.shape
        x_r = np.zeros((H, W, c), dtype=np.uint8)
        e_time = perf_counter()
        cs = compress_image('CAE', checkpoint, input_filename, temp_output_filename,
                            patch_size=patch_size, source_format=source_format, data_group=data_group,
                            data_axes=data_axes, progress_bar=progress_bar, gpu=gpu, **kwargs)
        decompress_image(temp_output_filename, x_r, data_group=data_group, checkpoint=checkpoint,
                         progress_bar=progress_bar, gpu=gpu, **kwargs)
        e_time = perf_counter() - e_time
        x = np.ascontiguousarray(src[0:H, 0:W])
        # one upload of each image; every metric then reads HBM
        xd = torch.from_numpy(x).cuda(non_blocking=True)
        xrd = torch.from_numpy(x_r).cuda(non_blocking=True)
        all_metrics = {}
        eval_time = perf_counter()
        for m_k, fn in metric_fun.items():
            t0 = perf_counter()
            try:
                score, _ = fn(x=xd, x_r=xrd, nbytes_stored=cs['bytes'])
            except ValueError:                     # e.g. ms-ssim on an image that is too small
                score = float('nan')
            torch.cuda.synchronize()
            all_metrics[m_k + '_time'] = perf_counter() - t0
            all_metrics[m_k] = score if score >= 0.0 else np.nan
        all_metrics['evaluation_time'] = perf_counter() - eval_time
        all_metrics['execution_time'] = e_time
        return all_metrics
    finally:
        if own_tmp:
            shutil.rmtree(os.path.dirname(temp_output_filename), ignore_errors=True)
