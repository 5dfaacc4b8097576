"""The measured hot path as one call: uint8 tiles -> analysis transform ->
quantize + likelihood/rate + histogram -> synthesis transform -> uint8 tiles.

This is the body of the reference's per-tile work (``ConvolutionalAutoencoder.
encode/decode``, ``/root/reference/src/models/tasks/_autoencoders.py:539-584``,
and ``forward_func``, ``_taskutils.py:95-108``) for a BATCH of tiles, with the
entropy coder left out (it is host code; see ``compress.py``).
"""
import os

import torch

from . import _cabi


class GraphedPipeline:
    """``CodecPipeline.__call__`` on one fixed input buffer, captured once into a CUDA graph
    and replayed: a step is one graph launch instead of ~11 kernel launches issued from
    Python, which is what keeps a host-fed tile loop GPU-bound.  ``x`` (the static input)
    must be refilled in stream order before each ``replay()``; the returned dict holds the
    static outputs of the capture (valid until the next replay)."""

    def __init__(self, pipe, x_static, warmup=3):
        self.x = x_static
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # executor buffers / packed weights: normal pool
            for _ in range(warmup):
                pipe(x_static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _cabi.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = pipe(x_static)
        self.launches = _cabi.launch_count() - n0   # kernels of this library per replay
        # the graph holds raw pointers into the executors' intermediate buffers: keep those
        # tensors alive for as long as the graph, whatever the executors cache later
        self._keep = [dict(pipe.model[k].module._executor()._buffers)
                      for k in ('encoder', 'decoder') if k in pipe.model]
        self._keep.append(getattr(pipe.model['decoder'].module, '_in_planar', None))

    def replay(self):
        self.graph.replay()
        return self.out


class CodecPipeline:
    def __init__(self, model):
        self.model = model
        self.level = len(model['decoder'].module.synthesis_track)
        # The quantizer can run inside the last analysis layer's epilogue (cae_conv_desc.quant).
        # Measured on B200 (ncu, net A, 128 x 256^2): that layer is epilogue bound with only 3.5
        # waves of tiles, so the fused form takes 142 us against 45 + 33 + 15 us for the layer,
        # the stand-alone quantizer and the layout conversion -- the default stays unfused.
        self.fuse_quantizer = bool(os.environ.get('CAE_FUSED_QUANT'))

    @torch.no_grad()
    def __call__(self, x_u8):
        """x_u8: N x H x W x C uint8 on the device.  Returns dict(x_r_u8, y, y_q, hist,
        bits, bpp): bpp is the estimated rate ``-sum(log2 p_y) / (N*H*W)``
        (``_ratedist.py:49-54``)."""
        n, h, w, _ = x_u8.shape
        fact_ent = self.model['fact_ent'].module
        req = fact_ent.quant_request() if self.fuse_quantizer else None
        y = self.model['encoder'](x_u8, quant=req)
        if req is not None and req.done:
            # quantize + likelihood + histogram + rate ran in the last encoder layer's epilogue
            y_q, hist, bits, planar = req.y_q, req.hist, req.rate, req.planar
        else:
            y_q, hist, bits = fact_ent.quantize_rate(y)
            planar = None
        _, _, x_r_u8 = self.model['decoder'](y_q, as_uint8='only', planar=planar)
        return dict(x_r_u8=x_r_u8, y=y, y_q=y_q, hist=hist, bits=bits,
                    bpp=bits / float(n * h * w))

    def graphed(self, x_static, warmup=3):
        """Capture this pipeline on the device buffer ``x_static`` (see GraphedPipeline)."""
        return GraphedPipeline(self, x_static, warmup)
