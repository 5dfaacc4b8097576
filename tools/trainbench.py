"""Config 5 of BASELINE.json on real GPUs: the train_cae_ms rate-distortion step
(/root/reference/src/train_cae_ms.py:209-230), batch 16 per GPU of 3 x 256 x 256 patches, data
parallel over one process per GPU with the persistent flat gradient bucket of
``cnn_autoencoder_b200.train_step`` (NCCL all-reduce, the synthesis half overlapped with the
analysis transform's backward).

    python bench.py --workload train [--steps K]            (1 GPU)
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 bench.py --workload train --gpus 8

Prints the bench.py JSON line: ``value`` = samples/s of the whole job with the batch resident in
HBM (CUDA events, max over ranks), ``e2e`` = the same with each step's batch copied from pinned
host memory and the loss read back, plus the gradient-bucket size, the all-reduce's share, the
loss trajectory and whether every rank holds bit-identical parameters after the last step.
Kernels: the training-mode bottleneck (cae_eb_train_fwd / bwd) and the transforms' forward /
backward (cae_conv_igemm, cae_act_grad, cae_conv_wgrad) run on this repo's kernels; CAE_TRAIN_TORCH=1
switches the transforms back to torch autograd (cuDNN) for comparison."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(args):
    import torch
    import torch.distributed as dist
    import cnn_autoencoder_b200 as M
    from cnn_autoencoder_b200 import _cabi
    from oracle import cae_oracle as O     # random-init checkpoint + synthetic patches only

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    batch = 16
    warmup = max(args.warmup, 3)
    chk = O.make_checkpoint(O.NAMED_ARCHS[args.arch], seed=1234)
    model = M.autoencoder_from_state_dict(chk, gpu=True, train=True)
    fwd = M.decorate_trainable_modules(trainable_modules=['encoder', 'decoder', 'fact_ent'],
                                       enabled_modules=['encoder', 'decoder', 'fact_ent'])
    crit = M.setup_loss('RateMSE', distortion_lambda=0.01)
    use_graph = not os.environ.get('CAE_TRAIN_EAGER_STEP')
    opts = M.setup_optimizers(model, lr=1e-4, aux_lr=1e-3, capturable=use_graph)
    bucket = M.GradBucket(model)
    x_pin = (O.synth_natural(batch, 3, 256, 256, seed=100 + rank).float() / 255.0).pin_memory()
    x = x_pin.cuda(non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = [0]
    graphed = None
    step_form = 'eager launches'
    if use_graph:
        # the whole step as one CUDA graph (train_step.GraphedTrainStep); eager if capture fails
        try:
            graphed = M.GraphedTrainStep(x, model, crit, opts, fwd, bucket, warmup=3)
            step_form = 'one CUDA graph per step (%d kernels of this library inside)' % graphed.kernels
        except Exception as exc:        # noqa: BLE001 -- report and fall back
            print('trainbench: graph capture failed (%s: %s), eager step' % (type(exc).__name__, exc),
                  file=sys.stderr)
            graphed = None
            torch.cuda.synchronize()

    def one(xb):
        if graphed is not None:
            out = graphed(xb)
        else:
            out = M.train_step(xb, model, crit, opts, fwd, bucket=bucket, step=step[0])
        step[0] += 1
        return out

    for _ in range(warmup):
        out = one(x)
    barrier()
    launches0 = _cabi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    losses = []
    for s, e in ev:
        s.record()
        out = one(x)
        e.record()
        losses.append(out['loss'].detach().clone())
    barrier()
    launches = _cabi.launch_count() - launches0
    ms = torch.tensor([sum(s.elapsed_time(e) for s, e in ev)], dtype=torch.float64, device='cuda')
    # e2e: the step's batch from pinned host memory, the loss back to the host
    x_stage = torch.empty_like(x)
    loss_host = torch.empty(1).pin_memory()
    barrier()
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        x_stage.copy_(x_pin, non_blocking=True)
        out = one(x_stage)
        loss_host.copy_(torch.mean(out['loss']).detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e0.record()
    barrier()
    e2e_ms = torch.tensor([s0.elapsed_time(e0)], dtype=torch.float64, device='cuda')
    # all-reduce alone (the same spans, no compute to hide under)
    ar_ms = 0.0
    if world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            dist.all_reduce(bucket.flat)
        barrier()
        a0.record()
        for _ in range(10):
            dist.all_reduce(bucket.flat)
        a1.record()
        barrier()
        ar_ms = a0.elapsed_time(a1) / 10
        bucket.flat.zero_()
    flat = torch.cat([p.detach().reshape(-1).double() for k in sorted(model)
                      for p in model[k].parameters()])
    lo, hi = flat.clone(), flat.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    def on_kernels(key):
        t = model[key].module
        return bool(getattr(t, 'train_kernels', False) and getattr(t, '_train_chain', None))
    if on_kernels('encoder') and on_kernels('decoder'):
        transforms = ('this repo: cae_conv_igemm forward and data gradient, cae_act_grad, cae_conv_wgrad '
                      '(tcgen05), fp16 operands / fp32 accumulate and master weights')
    else:
        transforms = 'torch autograd (cuDNN)'
    if rank == 0:
        step_ms = ms.item() / args.steps
        lv = [float(torch.mean(v).item()) for v in losses]
        print(json.dumps({
            'metric': 'train_step_samples_per_sec', 'value': round(world * batch / (step_ms / 1e3), 1),
            'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
            'ms_per_step': round(step_ms, 3), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'train_cae_ms rate-distortion step, net %s, batch %d x 3 x 256 x 256 per GPU, '
                                   'RateMSE lambda 0.01, Adam 1e-4 / aux 1e-3, clip 1.0' % (args.arch, batch),
                       'kernels': 'bottleneck fwd/bwd: cae_eb_train_fwd / cae_eb_train_bwd (this repo); '
                                  'transforms fwd/bwd: ' + transforms,
                       'step': step_form,
                       'parallelism': 'dp%d, one persistent flat fp32 gradient bucket, NCCL all-reduce, '
                                      'synthesis half launched under the analysis backward' % world},
            'e2e': {'value': round(world * batch * args.steps / (e2e_ms.item() / 1e3), 1), 'unit': 'samples/s',
                    'h2d_bytes_per_step': int(x_pin.numel() * 4), 'd2h_bytes_per_step': 4},
            'gpu_launches': int(launches),
            'grad_bucket_elems': int(bucket.numel()), 'allreduce_alone_ms': round(ar_ms, 4),
            'replicas_identical': bool(torch.equal(lo, hi)),
            'loss_first': lv[0], 'loss_last': lv[-1], 'finite': all(v == v for v in lv)}))
    if world > 1:
        # a captured graph holds NCCL kernels: release it before the communicator goes away, and do
        # not let a stuck teardown keep the GPUs (observed: destroy_process_group hanging with the
        # graph alive)
        import gc
        import threading
        graphed = None
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--arch', default='A')
    run(ap.parse_args())
