"""Per-layer timing of the bench pipeline (CUDA events, L2 flushed between repeats), with
optional sweeps over the igemm bring-up knobs (environment variables read per launch).

    python tools/layerbench.py [--arch A] [--batch 128] [--size 256] [--sweep]
"""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import cae_oracle as O  # noqa: E402
import cnn_autoencoder_b200 as M  # noqa: E402
from cnn_autoencoder_b200 import _cabi as C, _ops  # noqa: E402
from cnn_autoencoder_b200.pipeline import CodecPipeline  # noqa: E402

KIND = {0: 'conv_s1', 1: 'conv_s2', 2: 'convT_s1', 3: 'convT_s2'}


def time_call(call, flush, reps=5):
    for _ in range(2):
        _ops.replay(call)
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        flush.view(torch.int64).sum()   # leave the L2 clean (see bench.py)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _ops.replay(call)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--arch', default='A')
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--sweep', action='store_true')
    ap.add_argument('--debug-sweep', action='store_true')
    ap.add_argument('--skip-sweep', action='store_true')
    ap.add_argument('--head-sweep', action='store_true')
    ap.add_argument('--env', action='append', default=[],
                    help='KEY=VAL[,KEY=VAL] : one extra settings row per option')
    args = ap.parse_args()
    chk = O.make_checkpoint(O.NAMED_ARCHS[args.arch], seed=1234)
    model = M.autoencoder_from_state_dict(chk, gpu=True, train=False)
    pipe = CodecPipeline(model)
    x = O.synth_natural(args.batch, 3, args.size, args.size, seed=1).permute(0, 2, 3, 1).contiguous().cuda()
    pipe(x)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    settings = [dict()]
    if args.sweep:
        settings = [dict(), dict(CAE_IGEMM_DEBUG='16'), dict(CAE_IGEMM_EPI_WARPS='8'),
                    dict(CAE_IGEMM_MT='1')]
    if args.debug_sweep:
        settings = [dict(), dict(CAE_IGEMM_DEBUG='8'), dict(CAE_IGEMM_DEBUG='1'),
                    dict(CAE_IGEMM_DEBUG='2'), dict(CAE_IGEMM_DEBUG='3'), dict(CAE_IGEMM_DEBUG='4'),
                    dict(CAE_IGEMM_DEBUG='11'), dict(CAE_IGEMM_DEBUG='15')]
    if args.skip_sweep:
        settings = [dict(), dict(CAE_IGEMM_DEBUG='32'), dict(CAE_IGEMM_DEBUG='64'),
                    dict(CAE_IGEMM_DEBUG='96'), dict(CAE_IGEMM_EPI_WARPS='8'),
                    dict(CAE_IGEMM_NO_FAST_EPILOGUE='1')]
    if args.head_sweep:
        settings = [dict(), dict(CAE_HEAD_DEBUG='4'), dict(CAE_HEAD_DEBUG='6'), dict(CAE_HEAD_DEBUG='2')]
    for spec in args.env:
        settings.append(dict(kv.split('=', 1) for kv in spec.split(',')))
    known = set(k for e in settings for k in e)
    for env in settings:
        for k in set(('CAE_IGEMM_EPI_WARPS', 'CAE_IGEMM_MT', 'CAE_IGEMM_NO_FAST_EPILOGUE', 'CAE_IGEMM_DEBUG',
                      'CAE_IGEMM_TPB', 'CAE_HEAD_DEBUG')) | known:
            os.environ.pop(k, None)
        os.environ.update(env)
        print('== settings', env or 'default')
        total = 0.0
        for name in ('encoder', 'decoder'):
            ex = model[name].module._executor()
            for k, st in enumerate(ex.steps):
                call = ex.last_calls.get(k)
                if call is None:          # second half of a fused head pair
                    continue
                ms = time_call(call, flush)
                total += ms
                if call[0] == 'head':
                    xin = call[1][0]
                    print(f'  {name[:3]}{k}+{k + 1} fused head {st.c_in}->{st.c_in}->{call[1][5]} '
                          f'@{xin.h}x{xin.w} {ms * 1e3:9.1f} us')
                    continue
                if call[0] == 'image_from_proj':
                    print(f'  {name[:3]}{k} image_from_proj {call[1][4]} ch @{call[1][2]}x{call[1][3]} '
                          f'{ms * 1e3:9.1f} us')
                    continue
                xin = call[0][1]
                flops = 2.0 * 9 * st.c_in * st.c_out * xin.n * xin.h * xin.w
                if st.kind == C.CONV_S2:
                    flops /= 4
                print(f'  {name[:3]}{k} {KIND[st.kind]:9s} {st.c_in:4d}->{st.c_out:4d} @{xin.h}x{xin.w} '
                      f'{"igemm" if call[1]["igemm"] else "direct":6s} {ms * 1e3:9.1f} us '
                      f'{flops / ms / 1e9:8.1f} TFLOP/s{" +skip" if call[1].get("skip") is not None else ""}'
                      f'{" +proj" if call[1].get("proj") is not None else ""}')
        print(f'  sum of conv layers {total * 1e3:.1f} us')


if __name__ == '__main__':
    main()
