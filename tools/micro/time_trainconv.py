"""Transforms of net A in train() mode, forward + backward: this repo's kernels against the torch
formulation (cuDNN), CUDA-event times and (when the profiler is available) the kernel table.

    python tools/micro/time_trainconv.py [--batch 16] [--size 256] [--profile]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import cae_oracle as O  # noqa: E402
import cnn_autoencoder_b200 as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--profile', action='store_true')
    args = ap.parse_args()
    chk = O.make_checkpoint(dict(O.NAMED_ARCHS['A'], bias=True), seed=4)
    x = (O.synth_natural(args.batch, 3, args.size, args.size, seed=3).float() / 255.0).cuda()
    for mode in ('kernels', 'torch'):
        model = M.autoencoder_from_state_dict(chk, gpu=True, train=True)
        enc, dec = model['encoder'], model['decoder']
        enc.module.train_kernels = dec.module.train_kernels = mode == 'kernels'

        def step():
            y = enc(x)
            x_r, _ = dec(y)
            loss = ((x_r[0] - x) ** 2).mean()
            loss.backward()
            return loss

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        print(f'{mode:8s} fwd+bwd median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f} ms')
        if args.profile:
            try:
                from torch.profiler import profile, ProfilerActivity
                with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                    step()
                    torch.cuda.synchronize()
                print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25,
                                                max_name_column_width=70))
            except Exception as exc:  # profiler not available on this box
                print('profiler unavailable:', exc)


if __name__ == '__main__':
    main()
