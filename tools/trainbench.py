"""Config 5 of BASELINE.json on real GPUs: the train_cae_ms rate-distortion step, batch 16 per
GPU of 3 x 256 x 256 patches, data parallel with ONE flat-bucket NCCL all-reduce per step.

    python tools/trainbench.py --steps 10
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/trainbench.py --steps 10

Reports per-step time (CUDA events, max over ranks), samples/s of the whole job, the size of the
gradient bucket, the loss trajectory, and checks that every rank holds bit-identical parameters
after the last step (the replicas must not drift).  The transforms' train()-mode forward and
backward are torch autograd ops on the device for now."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cnn_autoencoder_b200 as M  # noqa: E402
from oracle import cae_oracle as O  # noqa: E402  (random-init checkpoint + synthetic patches only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--arch', default='A')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    chk = O.make_checkpoint(O.NAMED_ARCHS[args.arch], seed=1234)
    model = M.autoencoder_from_state_dict(chk, gpu=True, train=True)
    fwd = M.decorate_trainable_modules(trainable_modules=['encoder', 'decoder', 'fact_ent'],
                                       enabled_modules=['encoder', 'decoder', 'fact_ent'])
    crit = M.setup_loss('RateMSE', distortion_lambda=0.01)
    opts = M.setup_optimizers(model, lr=1e-4, aux_lr=1e-3)
    x = (O.synth_natural(args.batch, 3, 256, 256, seed=100 + rank).float() / 255.0).cuda()
    losses = []
    for _ in range(args.warmup):
        out = M.train_step(x, model, crit, opts, fwd)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for s, e in ev:
        s.record()
        out = M.train_step(x, model, crit, opts, fwd)
        e.record()
        losses.append(out['loss'] if isinstance(out, dict) and 'loss' in out else None)
    torch.cuda.synchronize()
    ms = torch.tensor([sum(s.elapsed_time(e) for s, e in ev)], dtype=torch.float64, device='cuda')
    flat = torch.cat([p.detach().reshape(-1).double() for k in sorted(model)
                      for p in model[k].parameters()])
    lo, hi = flat.clone(), flat.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        step_ms = ms.item() / args.steps
        loss_vals = [None if v is None else float(torch.as_tensor(v).float().mean().item())
                     for v in losses]
        print(json.dumps(dict(config='train_cae_ms step, net %s, batch %d x 3 x 256 x 256 per GPU' %
                              (args.arch, args.batch), n_gpus=world, steps=args.steps,
                              ms_per_step=round(step_ms, 3),
                              samples_per_s=round(world * args.batch / (step_ms / 1e3), 1),
                              grad_bucket_elems=int(flat.numel()),
                              replicas_identical=bool(torch.equal(lo, hi)),
                              loss_first=loss_vals[0], loss_last=loss_vals[-1],
                              finite=all(v is None or v == v for v in loss_vals))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
