"""Rate-distortion training step with a data-parallel gradient all-reduce, mirroring the
step body of the reference trainer (``/root/reference/src/train_cae_ms.py:209-230``):
forward closure -> criterion -> ``loss.backward()`` and ``entropy_loss.backward()`` -> per
optimizer ``clip_grad_norm_(1.0)``, ``step``, ``zero_grad``.

The reference replicates modules with ``nn.DataParallel`` inside one process
(``_autoencoders.py:514-517``).  Here one process drives one GPU and the replicas' gradients
are summed with ONE flat-bucket all-reduce (NCCL on GPUs, gloo in the CPU tests) between the
backward passes and the clipping, then divided by the world size so the update equals the
single-process update on the concatenated batch (``torch.mean`` over the batch, :214).

Status: the ``train()``-mode forward / backward of the transforms runs on torch autograd ops
on the device: the tensor-core kernels of this repo are the inference
path; dgrad / wgrad kernels are not written yet.
"""
import torch
import torch.distributed as dist
import torch.nn as nn


def setup_optimizers(model, lr=1e-4, aux_lr=1e-3):
    """Per-module Adam optimizers split like ``train_cae_ms.py:592-596``: encoder, decoder,
    fact_ent (everything but the quantiles) and fact_ent_aux (the quantiles)."""
    opts = {}
    for k in ('encoder', 'decoder'):
        if k in model:
            opts[k] = torch.optim.Adam(model[k].parameters(), lr=lr)
    if 'fact_ent' in model:
        main = [p for n, p in model['fact_ent'].named_parameters() if 'quantiles' not in n]
        aux = [p for n, p in model['fact_ent'].named_parameters() if 'quantiles' in n]
        opts['fact_ent'] = torch.optim.Adam(main, lr=lr)
        opts['fact_ent_aux'] = torch.optim.Adam(aux, lr=aux_lr)
    return opts


def allreduce_gradients(model, group=None):
    """Sum the gradients of every parameter of the model dict across ranks with one flat fp32
    bucket and divide by the world size.  Returns the bucket size in elements."""
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
               group=None):
    """One step on this rank's shard ``x`` of the global batch.  Returns the loss dict."""
    output = forward_func(x, model)
    loss_dict = criterion(inputs=x, outputs=output, targets=targets, net=model)
    loss = torch.mean(loss_dict['loss'])
    loss.backward()
    if 'entropy_loss' in loss_dict:
        torch.mean(loss_dict['entropy_loss']).backward()
    allreduce_gradients(model, group)
    for opt in optimizers.values():
        nn.utils.clip_grad_norm_(opt.param_groups[0]['params'], max_norm=max_norm)
        opt.step()
        opt.zero_grad()
    return loss_dict
