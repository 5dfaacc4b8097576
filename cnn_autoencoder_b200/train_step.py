"""Rate-distortion training step with a data-parallel gradient all-reduce, mirroring the
step body of the reference trainer (``/root/reference/src/train_cae_ms.py:209-230``):
forward closure -> criterion -> ``loss.backward()`` and ``entropy_loss.backward()`` -> per
optimizer, when ``step % mod_grad_accumulate[k] == 0``: ``clip_grad_norm_(1.0)``, ``step``,
``zero_grad``.

The reference replicates modules with ``nn.DataParallel`` inside one process
(``_autoencoders.py:514-517``): parameters are broadcast and gradients reduced to GPU 0 on every
step.  Here one process drives one GPU and the replicas' gradients live in ONE persistent flat
fp32 bucket (every ``p.grad`` is a view into it: no per-step ``cat`` / scatter), ordered by
optimizer: ``[decoder | encoder | fact_ent | fact_ent_aux]``.  The synthesis transform's
gradients (half of the bucket) are complete as soon as the gradient of the latent ``y`` reaches
the analysis transform, so their all-reduce is launched from a tensor hook on ``y`` and overlaps
the analysis transform's backward pass; the rest follows the auxiliary backward.  Gradients are divided by the world size, so the update equals the single-process
update on the concatenated batch (``torch.mean`` over the batch, :214).

Kernels: the training-mode bottleneck (noise proxy + likelihood, forward and backward) runs on
this repo's fused kernels (``cae_eb_train_fwd`` / ``cae_eb_train_bwd``); the transforms'
``train()``-mode forward / backward are torch autograd ops (cuDNN) on the device -- dgrad / wgrad
kernels are not written.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

ORDER = ('decoder', 'encoder', 'fact_ent', 'fact_ent_aux')


def _split_params(model):
    """Parameter lists keyed like the reference's optimizers (``train_cae_ms.py:592-596``):
    encoder, decoder, fact_ent (everything but the quantiles), fact_ent_aux (the quantiles)."""
    groups = {}
    for k in ('encoder', 'decoder'):
        if k in model:
            groups[k] = [p for p in model[k].parameters() if p.requires_grad]
    if 'fact_ent' in model:
        named = [(n, p) for n, p in model['fact_ent'].named_parameters() if p.requires_grad]
        groups['fact_ent'] = [p for n, p in named if 'quantiles' not in n]
        groups['fact_ent_aux'] = [p for n, p in named if 'quantiles' in n]
    return groups


def setup_optimizers(model, lr=1e-4, aux_lr=1e-3, capturable=False):
    """Per-module Adam optimizers split like ``train_cae_ms.py:592-596``.  ``capturable``: keep
    Adam's step counters on the device so that the step can be part of a CUDA graph
    (``GraphedTrainStep``)."""
    groups = _split_params(model)
    return {k: torch.optim.Adam(v, lr=aux_lr if k == 'fact_ent_aux' else lr, capturable=capturable)
            for k, v in groups.items() if v}


class GradBucket:
    """One flat fp32 gradient buffer for the whole model dict; ``p.grad`` of every parameter is
    a view into it (autograd accumulates in place).  ``ranges[k]`` is the slice of optimizer
    ``k``'s parameters."""

    def __init__(self, model):
        groups = _split_params(model)
        self.params = {k: groups[k] for k in ORDER if groups.get(k)}
        total = sum(p.numel() for v in self.params.values() for p in v)
        first = next(p for v in self.params.values() for p in v)
        self.flat = torch.zeros(total, dtype=torch.float32, device=first.device)
        self.ranges = {}
        off = 0
        for k, plist in self.params.items():
            start = off
            for p in plist:
                if p.dtype != torch.float32:
                    raise TypeError('GradBucket expects fp32 parameters')
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.ranges[k] = (start, off)
        self.pending = []

    def numel(self):
        return self.flat.numel()

    def span(self, keys):
        keys = [k for k in keys if k in self.ranges]
        if not keys:
            return None
        return min(self.ranges[k][0] for k in keys), max(self.ranges[k][1] for k in keys)

    def attached(self):
        """Views survive ``zero_grad(set_to_none=False)``; re-attach if somebody replaced them."""
        for plist in self.params.values():
            for p in plist:
                if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                        p.grad.data_ptr() >= self.flat.data_ptr() + 4 * self.flat.numel():
                    return False
        return True

    def allreduce(self, keys, group=None, async_op=False):
        """Sum this span over the ranks and divide by the world size."""
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(group)
        span = self.span(keys)
        if world == 1 or span is None:
            return None
        view = self.flat[span[0]:span[1]]
        work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            self.pending.append((work, view, world))
            return work
        view.div_(world)
        return None

    def wait(self):
        for work, view, world in self.pending:
            work.wait()
            view.div_(world)
        self.pending = []

    def zero(self, key):
        a, b = self.ranges[key]
        self.flat[a:b].zero_()


def allreduce_gradients(model, group=None):
    """Sum the gradients of every parameter of the model dict across ranks (one flat bucket) and
    divide by the world size.  Stand-alone form for callers without a ``GradBucket``.  Returns
    the bucket size in elements."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    params = [p for k in sorted(model) for p in model[k].parameters() if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1).float() for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()


def train_step(x, model, criterion, optimizers, forward_func, targets=None, max_norm=1.0,
               group=None, bucket=None, step=0, mod_grad_accumulate=None):
    """One step on this rank's shard ``x`` of the global batch.  Returns the loss dict.

    ``bucket``: a ``GradBucket`` (built once per model); without it the gradients are gathered
    into a temporary flat tensor each step.  ``mod_grad_accumulate``: ``{optimizer: period}``
    as in ``train_cae_ms.py:221-222`` (an optimizer clips / steps / zeroes only on steps where
    ``step % period == 0``; its gradients are all-reduced on those steps only, so accumulated
    local contributions are summed exactly once)."""
    acc = mod_grad_accumulate or {}
    stepping = [k for k in optimizers if step % int(acc.get(k, 1)) == 0]
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    output = forward_func(x, model)
    early = ('decoder',)
    overlap = (bucket is not None and distributed and all(k in stepping for k in optimizers)
               and isinstance(output.get('y'), torch.Tensor) and output['y'].requires_grad)
    if overlap:
        # the synthesis transform's gradients are final once d loss / d y is handed to the
        # analysis transform: start their all-reduce there, under that transform's backward
        output['y'].register_hook(lambda g: (bucket.allreduce(early, group, async_op=True), g)[1])
    loss_dict = criterion(inputs=x, outputs=output, targets=targets, net=model)
    loss = torch.mean(loss_dict['loss'])
    loss.backward()
    if 'entropy_loss' in loss_dict:
        torch.mean(loss_dict['entropy_loss']).backward()
    if bucket is None:
        if stepping:
            allreduce_gradients(model, group)
    elif overlap:
        bucket.allreduce(('encoder', 'fact_ent', 'fact_ent_aux'), group, async_op=True)
        bucket.wait()
    elif distributed:
        for k in stepping:
            bucket.allreduce((k,), group)
    for k in stepping:
        opt = optimizers[k]
        nn.utils.clip_grad_norm_(opt.param_groups[0]['params'], max_norm=max_norm)
        opt.step()
        if bucket is not None:
            bucket.zero(k)
        else:
            opt.zero_grad()
    return loss_dict


class GraphedTrainStep:
    """``train_step`` captured once as ONE CUDA graph and replayed: forward closure, criterion,
    both backward passes, the gradient all-reduce (NCCL kernels are graph nodes like any other),
    clipping and the Adam updates.  The eager step spends more host time on launching its ~500
    small kernels than the GPU spends running them; replayed, a step costs one launch.

    Restrictions of a static graph: fixed batch shape (``x`` is copied into a static buffer),
    every optimizer steps on every call (``mod_grad_accumulate`` periods of 1), optimizers built
    with ``capturable=True``.  The warm-up iterations before the capture ARE training steps (the
    capture itself records a step without running it).
    Returns the loss dict of the step as static tensors (overwritten by the next call)."""

    def __init__(self, x_example, model, criterion, optimizers, forward_func, bucket,
                 max_norm=1.0, group=None, warmup=3):
        for k, opt in optimizers.items():
            if not all(g.get('capturable', False) for g in opt.param_groups):
                raise ValueError(f'optimizer {k!r} must be built with capturable=True')
        self.x = x_example.detach().clone()
        args = (model, criterion, optimizers, forward_func)
        kw = dict(max_norm=max_norm, group=group, bucket=bucket, step=0)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                train_step(self.x, *args, **kw)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _cabi
        n0 = _cabi.lib().cae_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = train_step(self.x, *args, **kw)
            self.out = {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
        self.kernels = int(_cabi.lib().cae_launch_count() - n0)
        self._note = _cabi.note_graph_replay

    def __call__(self, x):
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        self._note(self.kernels)
        return self.out
