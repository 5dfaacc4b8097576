"""Two slides in flight: ``compress_image`` calls on one thread and CUDA stream,
``decompress_image`` calls on another.

The reference works through a list of slides one call after the other (``compress.py:197-209``,
``decompress.py:172-184``: a loop over the files of ``args.data_dir``).  One call alone leaves parts of the
machine idle that the next call could use: the upload direction of the host link is idle during
``decompress_image`` and the download direction during ``compress_image``; the GPU idles under
the tail of a compress call (last coder call, stream download, chunk-file writes) and under the
prologue of a decompress call (file reads, upload, first decode).  ``SlideJobs`` runs the two kinds
of call on two worker threads, each with a CUDA stream of its own as the call's main stream, so
that a compress call of one slide overlaps the decompress call of another.  The calls themselves
are the public tile loops, unchanged; the batched engine keeps separate buffers, graphs and coder
streams for the two directions (``_slide.TileCodec``).  CUDA graphs are captured on a model's first
call of each kind, and a capture does not tolerate launches from another thread: run one compress
and one decompress call alone before putting two in flight.
"""
import threading
from concurrent.futures import ThreadPoolExecutor

import torch

__all__ = ['SlideJobs']


class SlideJobs:
    STAGES = ('compress', 'decompress')

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('SlideJobs needs a CUDA device (no CPU fallback)')
        self.device = torch.cuda.current_device() if device is None else device
        self._tls = threading.local()
        self._pools = {s: ThreadPoolExecutor(max_workers=1, thread_name_prefix='cae-' + s,
                                             initializer=self._init_thread) for s in self.STAGES}

    def _init_thread(self):
        torch.cuda.set_device(self.device)
        self._tls.stream = torch.cuda.Stream(self.device)

    def _run(self, fn, args, kwargs):
        with torch.cuda.stream(self._tls.stream):
            return fn(*args, **kwargs)

    def submit(self, stage, fn, *args, **kwargs):
        """Queue ``fn(*args, **kwargs)`` on the worker of ``stage`` ('compress' or 'decompress');
        calls of one stage run in submission order.  Returns a ``concurrent.futures.Future``."""
        return self._pools[stage].submit(self._run, fn, args, kwargs)

    def close(self, wait=True):
        for p in self._pools.values():
            p.shutdown(wait=wait)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
