#!/usr/bin/env python
"""Benchmark of the compress/decompress hot path (BASELINE.json metric).

Workload (configs[1]): ImageNet-style 256x256 RGB patches, batch 128 per GPU,
net A (3->128->128->48, level 3, LeakyReLU), random-init weights, synthetic
natural-image-like uint8 tiles.  One step = encode (analysis transform) +
quantize + factorized-prior rate/histogram + decode (synthesis transform, uint8
out) of one batch.  Tiles are independent, so N GPUs = N processes each running
its own batches with no data-path collective ("weak" scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = megapixels/s with inputs resident in
HBM; `e2e` = the same through the public API from pinned host buffers with the
H2D / D2H copies inside the timed region; `roofline` = the dominant kernel (the
128->128 stride-1 implicit-GEMM layer) timed alone with CUDA events against the
measured bf16 tensor peak; `cpu_baseline` = the CPU oracle on this box's cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH_NAME = 'A'
BATCH, SIZE = 128, 256
FLOPS_PER_PX = {'A': 136746, 'A_res': 321876, 'B': 367236, 'M': 738}
METRIC = 'wsi_encode_decode_megapixels_per_sec'


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d['bf16_tflops'], bf16_sustained=d['bf16_tflops_sustained'],
                    hbm=d['hbm_gbs'], source='measured')
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-i', str(self.index), '-lms', '25'], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        import datetime
        for r in self.rows:
            try:
                ts = datetime.datetime.strptime(r[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                if self.t0 is not None and not (self.t0 - 0.05 <= ts <= self.t1 + 0.05):
                    continue
                r = r[1:]
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                    samples=len(sm))


def run_reference(args, rank, world):
    """The reference's CPU path for the same metric: its Analyzer/Synthesizer arithmetic
    (oracle restatement, bit-exact against the reference classes) + restated
    EntropyBottleneck, fp32, all host threads, on a bounded sample of the workload."""
    if rank != 0:
        return
    import torch
    from oracle import cae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    chk = O.make_checkpoint(O.NAMED_ARCHS[ARCH_NAME], seed=1234)
    model = O.OracleModel(chk)
    n = 8
    x_u8 = O.synth_natural(n, 3, SIZE, SIZE, seed=1)

    def step():
        x = x_u8.float() / 255.0
        out = model.forward(x)
        O.rate_loss(x, out['p_y'])
        (out['x_r'][0] * 255.0).clip(0, 255).to(torch.uint8)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    mp = n * SIZE * SIZE * args.steps / 1e6
    val = mp / dt
    sample = f'{n} of {BATCH} tiles of {SIZE}x{SIZE} per step (CPU fp32, torch {torch.__version__})'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': round(val, 4), 'unit': 'MP/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(dt / args.steps * 1e3, 3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'net {ARCH_NAME} (3->128->128->48 L3 LeakyReLU), '
                               f'{BATCH}x3x{SIZE}x{SIZE} uint8 patches per GPU, '
                               'encode+quantize+rate+decode', 'sample': sample},
        'cpu_baseline': {'value': round(val, 4), 'unit': 'MP/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': round(val, 4), 'unit': 'MP/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--arch', default=ARCH_NAME)
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true',
                    help='issue every kernel from Python instead of replaying a CUDA graph')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from oracle import cae_oracle as O       # weights/tiles generator + cpu_baseline leg only
    import cnn_autoencoder_b200 as M
    from cnn_autoencoder_b200 import _cabi, _ops
    from cnn_autoencoder_b200.pipeline import CodecPipeline

    if not torch.cuda.is_available():
        sys.exit('bench.py: no CUDA device (the hot path has no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    warmup = max(args.warmup, 3)
    arch = O.NAMED_ARCHS[args.arch]
    chk = O.make_checkpoint(arch, seed=1234)
    model = M.autoencoder_from_state_dict(chk, gpu=True, train=False)
    pipe = CodecPipeline(model)
    B = args.batch
    # distinct tiles per rank: the slide is sharded by chunk range, no exchange between ranks
    x_host = O.synth_natural(B, 3, SIZE, SIZE, seed=1 + rank).permute(0, 2, 3, 1).contiguous()
    x_pin = x_host.pin_memory()
    x_dev = x_pin.cuda(non_blocking=True)
    out_pin = torch.empty_like(x_pin).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2

    def flush_l2():
        # write 256 MiB (evicts everything), then read it back so the L2 is left holding
        # clean lines: otherwise the first timed kernel pays for writing back the fill
        flush.fill_(1)
        flush.view(torch.int64).sum()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    # One step = one replay of the pipeline captured in a CUDA graph (the same kernels, in the
    # same order, as the eager call; --no-graph issues them one by one from Python).
    if not args.no_graph and any(k in os.environ for k in (
            'CUDA_INJECTION64_PATH', 'NV_COMPUTE_PROFILER_PERFWORKS_DIR', 'NV_NSIGHT_INJECTION_PORT_BASE')):
        # a profiler is attached: ncu --set full cannot replay the tensor-map kernels as graph
        # nodes (LaunchFailed on its second pass, profiles/README.md), so launch them one by one
        print('bench.py: profiler detected, launching kernels eagerly (no CUDA graph)', file=sys.stderr)
        args.no_graph = True
    if args.no_graph:
        run_dev = lambda: pipe(x_dev)
        launches_per_step = None
    else:
        try:
            g_dev = pipe.graphed(x_dev, warmup=warmup)
            run_dev = g_dev.replay
            launches_per_step = g_dev.launches
        except RuntimeError as exc:            # capture refused: time the eager launches instead
            print(f'bench.py: CUDA graph capture failed ({exc}); timing eager launches',
                  file=sys.stderr)
            args.no_graph = True
            torch.cuda.synchronize()
            run_dev = lambda: pipe(x_dev)
            launches_per_step = None
    for _ in range(warmup):
        out = run_dev()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        time.sleep(0.5)                     # let nvidia-smi start sampling (untimed)
        for _ in range(3):
            out = run_dev()
        barrier()
        launches0 = _cabi.launch_count()
        t_begin = time.time()
        for s, e in ev:
            flush_l2()                      # evict L2 between timed iterations (untimed)
            s.record()
            out = run_dev()
            e.record()
        barrier()
        clk.window(t_begin, time.time())
    launches = _cabi.launch_count() - launches0
    if launches_per_step is not None:
        launches = launches_per_step * args.steps     # graph replays bypass the ABI's counter
    step_ms = [s.elapsed_time(e) for s, e in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = total_ms.item()
    px_per_step = B * SIZE * SIZE
    value = world * px_per_step * args.steps / (total_ms / 1e3) / 1e6
    bpp = out['bpp'].item()

    # ---------------- end to end through the public API, host buffers ----------------
    # Every step: H2D of that step's uint8 tiles from pinned memory, the pipeline call, D2H of
    # the reconstructed tiles and of the rate (the step's metric).  Copies run on their own
    # streams so step i+1's upload and step i-1's download overlap step i's kernels, the way
    # the tile loop (compress.py / decompress.py) drives the device.
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    x_stage = [torch.empty_like(x_dev) for _ in range(2)]
    out_stage = [torch.empty_like(x_pin).pin_memory() for _ in range(2)]
    bpp_stage = [torch.empty(1, dtype=torch.float64).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    g_e2e = None if args.no_graph else [pipe.graphed(x_stage[k]) for k in range(2)]

    def e2e_run(n_steps):
        for i in range(n_steps):
            k = i % 2
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_free[k])          # step i-2 has consumed this buffer
                x_stage[k].copy_(x_pin, non_blocking=True)
                ev_in[k].record(s_in)
            main.wait_event(ev_in[k])
            if g_e2e is None:
                o = pipe(x_stage[k])
            else:
                if i >= 2:
                    main.wait_event(ev_out[k])           # step i-2's results have left the device
                o = g_e2e[k].replay()
            ev_free[k].record(main)
            ev_done[k].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[k])
                if g_e2e is None:
                    o['x_r_u8'].record_stream(s_out)
                    o['bpp'].record_stream(s_out)
                out_stage[k].copy_(o['x_r_u8'], non_blocking=True)
                bpp_stage[k].copy_(o['bpp'], non_blocking=True)
                ev_out[k].record(s_out)
        s_out.synchronize()
        s_in.synchronize()

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    e2e_run(args.steps)
    main.wait_stream(s_out)
    e0.record()
    barrier()
    e2e_ms = torch.tensor([max(s0.elapsed_time(e0), 0.0)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * px_per_step * args.steps / (e2e_ms.item() / 1e3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (rank 0, kernel alone) ----------------
    peaks = measured_peaks()
    ex = model['encoder'].module._executor()
    dom = None
    for k, st in enumerate(ex.steps):
        if st.kind == _cabi.CONV_S1 and st.c_in >= 64:
            dom = k
            break
    roof = None
    if dom is not None and ex.last_calls:
        call = ex.last_calls[dom]
        # MEASURED_PEAKS.json's burst figure (the roofline denominator) is a best-of-10 on a
        # device that was not under sustained load; give this kernel the same footing after
        # the long timed loops above by letting the device idle for a moment, then warm up
        torch.cuda.synchronize()
        time.sleep(2.0)
        for _ in range(3):
            _ops.replay(call)
        torch.cuda.synchronize()
        reps = 20
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(reps)]
        # the kernel is timed ALONE (one launch between synchronisations, as MEASURED_PEAKS.json's
        # burst figure was taken), with the L2 flush queued right before it so that the device
        # does not drop out of its boost clocks while idle
        for s, e in evs:
            flush_l2()
            s.record()
            _ops.replay(call)
            e.record()
            torch.cuda.synchronize()
        times = sorted(s.elapsed_time(e) for s, e in evs)
        t_ms = statistics.mean(times)            # the average launch, as the contract asks
        t_med = statistics.median(times)
        x_in = call[0][1]
        st = ex.steps[dom]
        flops = 2.0 * 9 * st.c_in * st.c_out * x_in.n * x_in.h * x_in.w
        ach = flops / (t_ms / 1e3) / 1e12
        roof = {'bound': 'tensor', 'achieved': round(ach, 2), 'peak': peaks['bf16'],
                'unit': 'TFLOP/s', 'frac': round(ach / peaks['bf16'], 4),
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed
                # ncu --set full capture (profiles/r01_ncu_full_enc2_128to128_s1_batch128.json);
                # only meaningful for the default batch
                'traffic': 1113868800 if (B == BATCH and args.arch == ARCH_NAME) else None,
                'kernel': 'igemm_conv_kernel<EPI_ACT> conv3x3 s1 %d->%d @%dx%d x%d' % (
                    st.c_in, st.c_out, x_in.h, x_in.w, x_in.n),
                'ms': round(t_ms, 4), 'ms_median': round(t_med, 4), 'ms_min': round(times[0], 4),
                'ms_max': round(times[-1], 4), 'peak_source': peaks['source'] + ' bf16 burst'}

    # ---------------- CPU baseline (bounded sample, rank 0, N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        oracle = O.OracleModel(chk)
        n = 8
        xs = O.synth_natural(n, 3, SIZE, SIZE, seed=1)

        def cpu_step():
            x = xs.float() / 255.0
            o = oracle.forward(x)
            O.rate_loss(x, o['p_y'])
            (o['x_r'][0] * 255.0).clip(0, 255).to(torch.uint8)

        cpu_step()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            cpu_step()
        dt = (time.perf_counter() - t0) / reps
        cpu = {'value': round(n * SIZE * SIZE / 1e6 / dt, 4), 'unit': 'MP/s', 'cores': cores,
               'kind': 'port',
               'sample': f'{n} of {B} tiles of {SIZE}x{SIZE}, {reps} repeats after 1 warm-up '
                         '(oracle: reference transforms restated + restated EntropyBottleneck)'}

    tf = value * 1e6 * FLOPS_PER_PX[args.arch] / 1e12
    line = {
        'metric': METRIC, 'value': round(value, 2), 'unit': 'MP/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': warmup, 'ms_per_step': round(total_ms / args.steps, 4),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16',
        'data': 'synthetic',
        'config': {'workload': f'net {args.arch} (3->128->128->48 L3 LeakyReLU), '
                               f'{B}x3x{SIZE}x{SIZE} uint8 patches per GPU, '
                               'encode+quantize+rate+decode',
                   'l2': 'flushed between timed iterations (256 MiB write, then read back)',
                   'accumulate': 'f32', 'est_bpp': round(bpp, 4),
                   'launch': 'eager, one ABI call per kernel' if args.no_graph
                   else 'one CUDA graph replay per step'},
        'clocks': clk.summary(),
        'e2e': {'value': round(e2e_value, 2), 'unit': 'MP/s',
                'h2d_bytes_per_step': int(x_pin.numel()),
                'd2h_bytes_per_step': int(out_pin.numel()) + 8},
        'gpu_launches': int(launches),
        'pipeline_tflops': round(tf / world, 2),
        'pipeline_tensor_frac': round(tf / world / peaks['bf16_sustained'], 4),
        'roofline': roof, 'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
