"""Kernel bring-up harness: every case runs in its own process (a trapped kernel
poisons its CUDA context) under a timeout, compares one ABI call against plain
torch fp32 ops on the same device, and the summary lands in gpurun_out/.

    python tools/bringup.py            # all cases
    python tools/bringup.py --case igemm_s1_small
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _mods():
    import torch
    import torch.nn.functional as F
    from cnn_autoencoder_b200 import _cabi as C, _ops as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch, F, C, O


def ref_conv(kind, x, w, pad_mode):
    torch, F, C, O = _mods()
    if kind in (C.CONV_S1, C.CONV_S2):
        xp = F.pad(x, (1, 1, 1, 1), mode='reflect' if pad_mode == C.PAD_REFLECT else 'constant')
        return F.conv2d(xp, w, stride=2 if kind == C.CONV_S2 else 1)
    s = 2 if kind == C.CONVT_S2 else 1
    return F.conv_transpose2d(x, w, stride=s, padding=1, output_padding=1 if s == 2 else 0)


def padded_view(act):
    """Act (planar/split) -> fp32 N x C x (H+2) x (W+2) including the halo."""
    torch, F, C, O = _mods()
    t = act.t.float()
    if act.fmt == C.FMT_F16_PLANAR:
        n, p, hp, wp, _ = t.shape
        return t.permute(0, 1, 4, 2, 3).reshape(n, p * 8, hp, wp)[..., O.COL_PAD:wp - O.COL_PAD]
    n, _, p, hh, wh, _ = t.shape
    full = torch.zeros(n, p * 8, hh * 2, wh * 2, device=t.device)
    for par in range(4):
        py, px = par >> 1, par & 1
        full[:, :, py::2, px::2] = t[:, par].permute(0, 1, 4, 2, 3).reshape(n, p * 8, hh, wh)
    return full[..., O.COL_PAD:wh * 2 - O.COL_PAD]


def fp16_vals(shape, gen, scale=1.0):
    torch, *_ = _mods()
    return (torch.randn(shape, generator=gen, device='cuda') * scale).half().float()


def check(name, got, want, tol):
    torch, *_ = _mods()
    if got.numel() == 0:
        print(f'  {name}: empty ok')
        return True
    err = (got - want).abs().max().item()
    ref = want.abs().max().item()
    ok = err <= tol * max(1.0, ref)
    print(f'  {name}: max_err={err:.3e} ref_max={ref:.3e} tol={tol:g} {"ok" if ok else "MISMATCH"}')
    return ok


# ------------------------------------------------------------------ cases
def case_direct(kind_name, pad_name='reflect'):
    torch, F, C, O = _mods()
    kind = getattr(C, kind_name)
    pad = C.PAD_REFLECT if pad_name == 'reflect' else C.PAD_ZERO
    g = torch.Generator(device='cuda').manual_seed(1)
    n, ci, co, h, w = 2, 3, 5, 12, 10
    x = torch.randn(n, ci, h, w, generator=g, device='cuda')
    transposed = kind in (C.CONVT_S1, C.CONVT_S2)
    wt = torch.randn((ci, co, 3, 3) if transposed else (co, ci, 3, 3), generator=g, device='cuda')
    bias = torch.randn(co, generator=g, device='cuda')
    ho, wo = O.KIND_OUT[kind](h, w)
    skip = torch.randn(n, co, ho, wo, generator=g, device='cuda')
    out = O.alloc_act(C.FMT_F32_NCHW, n, co, ho, wo)
    O.conv(kind, O.wrap_nchw(x), wt, co, out, igemm=False, bias=bias, skip=O.wrap_nchw(skip),
           pre_act=C.ACT_LEAKY_RELU, post_act=C.ACT_RELU, pad_mode=pad)
    want = torch.relu(F.leaky_relu(ref_conv(kind, x, wt, pad) + bias.view(1, -1, 1, 1), 0.01) + skip)
    ok = check('f32->f32', out.t, want, 1e-5)
    # u8 in, planar out with reflect halo, then back
    xu = torch.randint(0, 256, (n, h, w, ci), generator=g, device='cuda', dtype=torch.uint8)
    if not transposed:
        for fmt in (C.FMT_F16_PLANAR, C.FMT_F16_SPLIT):
            if fmt == C.FMT_F16_SPLIT and (ho % 2 or wo % 2):
                continue
            o2 = O.alloc_act(fmt, n, co, ho, wo, halo=C.HALO_REFLECT)
            O.conv(kind, O.wrap_u8_hwc(xu), wt, co, o2, igemm=False, pad_mode=pad)
            xf = xu.permute(0, 3, 1, 2).float() / 255.0
            want2 = ref_conv(kind, xf, wt, pad)
            pv = padded_view(o2)[:, :co]
            ok &= check(f'u8->fmt{fmt} interior', pv[:, :, 1:-1, 1:-1], want2, 2e-3)
            ok &= check(f'u8->fmt{fmt} halo', pv, F.pad(pv[:, :, 1:-1, 1:-1], (1, 1, 1, 1), mode='reflect'), 0)
            ok &= check(f'fmt{fmt}->nchw', O.planar_to_nchw(o2), pv[:, :, 1:-1, 1:-1], 0)
    else:
        o3 = O.alloc_act(C.FMT_U8_HWC, n, co, ho, wo)
        aux = torch.empty(n, co, ho, wo, device='cuda')
        O.conv(kind, O.wrap_nchw(x * 0.2), wt, co, o3, igemm=False, aux=aux)
        want3 = ref_conv(kind, x * 0.2, wt, pad)
        ok &= check('aux', aux, want3, 1e-5)
        wu8 = (aux * 255.0).clip(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
        ok &= check('u8 out', o3.t.float(), wu8.float(), 0)
    return ok


def case_layout():
    torch, F, C, O = _mods()
    g = torch.Generator(device='cuda').manual_seed(2)
    ok = True
    x = fp16_vals((2, 19, 6, 8), g)
    for fmt in (C.FMT_F16_PLANAR, C.FMT_F16_SPLIT):
        for halo in (C.HALO_KEEP, C.HALO_REFLECT):
            a = O.nchw_to_planar(x, fmt, halo)
            pv = padded_view(a)
            want = F.pad(x, (1, 1, 1, 1), mode='reflect' if halo else 'constant')
            ok &= check(f'fmt{fmt} halo{halo}', pv[:, :19], want, 0)
            ok &= check(f'fmt{fmt} pad-channels', pv[:, 19:], torch.zeros_like(pv[:, 19:]), 0)
            ok &= check(f'fmt{fmt} back', O.planar_to_nchw(a), x, 0)
    return ok


def _igemm_case(kind_name, n, ci, co, h, w, *, out_fmt=None, halo_out=None, mt=0, ck=0,
                bias=False, skip=False, pre=0, post=0, grid=0, tol=2e-3):
    torch, F, C, O = _mods()
    kind = getattr(C, kind_name)
    g = torch.Generator(device='cuda').manual_seed(3)
    transposed = kind in (C.CONVT_S1, C.CONVT_S2)
    x = fp16_vals((n, ci, h, w), g)
    wt = fp16_vals((ci, co, 3, 3) if transposed else (co, ci, 3, 3), g, scale=0.2)
    pad = C.PAD_ZERO if transposed else C.PAD_REFLECT
    in_fmt = C.FMT_F16_SPLIT if kind == C.CONV_S2 else C.FMT_F16_PLANAR
    xin = O.nchw_to_planar(x, in_fmt, C.HALO_KEEP if transposed else C.HALO_REFLECT)
    packed = O.pack_weights(kind, wt, ck=ck)
    ho, wo = O.KIND_OUT[kind](h, w)
    b = fp16_vals((co,), g) if bias else None
    sk = fp16_vals((n, co, ho, wo), g) if skip else None
    want = ref_conv(kind, x, wt, pad)
    if b is not None:
        want = want + b.view(1, -1, 1, 1)
    act = {0: lambda t: t, 1: lambda t: F.leaky_relu(t, 0.01), 2: torch.relu}
    want = act[pre](want)
    if sk is not None:
        want = want + sk
    want = act[post](want)
    merged = kind == C.CONVT_S2 and co * 4 <= 16
    ok = True
    if merged:
        out = O.alloc_act(C.FMT_U8_HWC, n, co, ho, wo)
        aux = torch.empty(n, co, ho, wo, device='cuda')
        O.conv(kind, xin, packed, co, out, igemm=True, bias=b, aux=aux, ck=ck, mt=mt, grid=grid,
               pre_act=pre, post_act=post)
        ok &= check('aux fp32', aux, want, tol)
        wu8 = (aux * 255.0).clip(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
        ok &= check('u8', out.t.float(), wu8.float(), 0)
        return ok
    fmts = [out_fmt] if out_fmt is not None else [C.FMT_F16_PLANAR, C.FMT_F32_NCHW]
    for fmt in fmts:
        if fmt == C.FMT_F32_NCHW and kind == C.CONVT_S2:
            continue
        halo = halo_out if halo_out is not None else (C.HALO_KEEP if transposed else C.HALO_REFLECT)
        out = O.alloc_act(fmt, n, co, ho, wo, halo=halo)
        skip_act = O.nchw_to_planar(sk, C.FMT_F16_PLANAR) if sk is not None else None
        O.conv(kind, xin, packed, co, out, igemm=True, bias=b, skip=skip_act, pre_act=pre,
               post_act=post, ck=ck, mt=mt, grid=grid)
        if fmt == C.FMT_F32_NCHW:
            ok &= check('f32 nchw', out.t, want, tol)
        else:
            pv = padded_view(out)
            got = pv[:, :co, 1:-1, 1:-1]
            ok &= check(f'fmt{fmt} interior', got, want, tol * 2)
            mode = 'reflect' if halo == C.HALO_REFLECT else 'constant'
            ok &= check(f'fmt{fmt} halo', pv, F.pad(pv[:, :, 1:-1, 1:-1], (1, 1, 1, 1), mode=mode), 0)
            ok &= check(f'fmt{fmt} pad-ch', pv[:, co:], torch.zeros_like(pv[:, co:]), 0)
    return ok


def case_head(n, ci, co, h, w, *, u8=True, pad_name='reflect', halo_out=None, act1=1, act2=1,
              tol=3e-3, residual=False):
    """cae_conv_head against conv_s2(fp16(act(conv_s1(x)))) in torch fp32."""
    torch, F, C, O = _mods()
    g = torch.Generator(device='cuda').manual_seed(11)
    pad = C.PAD_REFLECT if pad_name == 'reflect' else C.PAD_ZERO
    if u8:
        img = torch.randint(0, 256, (n, h, w, ci), generator=g, device='cuda', dtype=torch.uint8)
        x = img.permute(0, 3, 1, 2).float() / 255.0
        xin = O.wrap_u8_hwc(img)
    else:
        x = torch.rand((n, ci, h, w), generator=g, device='cuda')
        xin = O.wrap_nchw(x)
    w1 = torch.randn((ci, ci, 3, 3), generator=g, device='cuda') * 0.4
    b1 = torch.randn((ci,), generator=g, device='cuda') * 0.1
    w2 = fp16_vals((co, ci, 3, 3), g, scale=0.3)
    b2 = torch.randn((co,), generator=g, device='cuda') * 0.1
    act = {0: lambda t: t, 1: lambda t: F.leaky_relu(t, 0.01), 2: torch.relu}
    s = act[act1](ref_conv(C.CONV_S1, x, w1, pad) + b1.view(1, -1, 1, 1))
    extra = {}
    if residual:
        w1b = torch.randn((ci, ci, 3, 3), generator=g, device='cuda') * 0.4
        b1b = torch.randn((ci,), generator=g, device='cuda') * 0.1
        s = act[act1](ref_conv(C.CONV_S1, s, w1b, pad) + b1b.view(1, -1, 1, 1) + x)
        extra = dict(w_stem2=w1b, b_stem2=b1b, act_mid=act1)
    s = s.half().float()
    want = act[act2](ref_conv(C.CONV_S2, s, w2, pad) + b2.view(1, -1, 1, 1))
    ho, wo = O.KIND_OUT[C.CONV_S2](h, w)
    halo = halo_out if halo_out is not None else C.HALO_REFLECT
    out = O.alloc_act(C.FMT_F16_PLANAR, n, co, ho, wo, halo=halo)
    O.conv_head(xin, w1, b1, w2, b2, co, out, act_stem=act1, act_down=act2, pad_mode=pad, **extra)
    pv = padded_view(out)
    ok = check('interior', pv[:, :co, 1:-1, 1:-1], want, tol)
    mode = 'reflect' if halo == C.HALO_REFLECT else 'constant'
    ok &= check('halo', pv, F.pad(pv[:, :, 1:-1, 1:-1], (1, 1, 1, 1), mode=mode), 0)
    ok &= check('pad-ch', pv[:, co:], torch.zeros_like(pv[:, co:]), 0)
    return ok


def case_eb():
    torch, F, C, O = _mods()
    import ctypes
    from cnn_autoencoder_b200 import _entropy
    torch.manual_seed(5)
    eb = _entropy.EntropyBottleneck(6).cuda()
    with torch.no_grad():
        eb.quantiles[:, 0, 1] += torch.rand(6, device='cuda') - 0.5
        eb._factor1.add_(torch.randn_like(eb._factor1) * 0.3)
    eb.update(force=True)
    eb.eval()
    y = torch.randn(3, 6, 9, 11, device='cuda') * 8
    y[0, 0, 0, 0] = 900.25
    y[2, 5, 8, 10] = -3000.75
    with torch.no_grad():
        y_q, p = eb(y)
        y_q_t, p_t = eb._forward_torch(y)
    ok = check('y_q', y_q, y_q_t, 0)
    ok &= check('p_y (rel)', p / p_t, torch.ones_like(p), 2e-4)
    sym, hist, bits = eb.symbols_hist_rate(y)
    ok &= check('symbols', sym.float(), torch.round(y - eb._medians().view(1, -1, 1, 1)), 0)
    ok &= check('rate', bits.float(), -torch.log2(p_t).sum().double().float().reshape(1), 1e-4)
    ok &= bool(hist.sum().item() == y.numel())
    return ok


CASES = {
    'direct_s1_reflect': lambda: case_direct('CONV_S1'),
    'direct_s1_zero': lambda: case_direct('CONV_S1', 'zero'),
    'direct_s2': lambda: case_direct('CONV_S2'),
    'direct_t1': lambda: case_direct('CONVT_S1'),
    'direct_t2': lambda: case_direct('CONVT_S2'),
    'layout': case_layout,
    # smallest tensor-core case: one tile, one chunk, mt=1
    'igemm_s1_tiny': lambda: _igemm_case('CONV_S1', 1, 16, 16, 16, 8, mt=1),
    'igemm_s1_small': lambda: _igemm_case('CONV_S1', 2, 32, 32, 20, 24),
    'igemm_s1_128': lambda: _igemm_case('CONV_S1', 2, 128, 128, 40, 48, bias=True, pre=1, skip=True, post=1),
    'igemm_s1_ck16_mt1': lambda: _igemm_case('CONV_S1', 1, 64, 48, 33 - 1, 26, ck=16, mt=1),
    'igemm_s1_grid3': lambda: _igemm_case('CONV_S1', 3, 64, 64, 36, 36, grid=3),
    'igemm_s2': lambda: _igemm_case('CONV_S2', 2, 32, 48, 24, 40),
    'igemm_s2_128_48': lambda: _igemm_case('CONV_S2', 2, 128, 48, 64, 64, out_fmt=2),
    'igemm_s2_split_out': lambda: _igemm_case('CONV_S2', 1, 64, 64, 40, 72, out_fmt=4, pre=1),
    'igemm_s2_16': lambda: _igemm_case('CONV_S2', 2, 3, 128, 32, 48, out_fmt=3, pre=1),
    'igemm_t1': lambda: _igemm_case('CONVT_S1', 2, 48, 48, 20, 24, bias=True, pre=1),
    'igemm_t1_192': lambda: _igemm_case('CONVT_S1', 1, 192, 192, 16, 16),
    'igemm_t2': lambda: _igemm_case('CONVT_S2', 2, 48, 128, 20, 24, pre=1, bias=True),
    'igemm_t2_128': lambda: _igemm_case('CONVT_S2', 1, 128, 128, 32, 32),
    'igemm_t2_merged': lambda: _igemm_case('CONVT_S2', 2, 128, 3, 24, 40, bias=True),
    'igemm_t2_merged_c1': lambda: _igemm_case('CONVT_S2', 1, 16, 1, 16, 16),
    'head_u8_3_128': lambda: case_head(2, 3, 128, 64, 96),
    'head_u8_multi_tile': lambda: case_head(3, 3, 128, 256, 256),
    'head_odd_c1': lambda: case_head(2, 1, 24, 37, 51, u8=False, act2=0),
    'head_c4_zero_pad': lambda: case_head(1, 4, 64, 40, 72, pad_name='zero', halo_out=0, act1=2),
    'head_small': lambda: case_head(1, 3, 48, 6, 10, u8=False),
    'head_res_u8': lambda: case_head(2, 3, 128, 64, 96, residual=True),
    'head_res_multi_tile': lambda: case_head(2, 3, 128, 256, 256, residual=True),
    'head_res_odd_c1': lambda: case_head(2, 1, 24, 37, 51, u8=False, act2=0, residual=True),
    'head_res_zero_pad': lambda: case_head(1, 4, 64, 40, 72, pad_name='zero', halo_out=0, act1=2, residual=True),
    'head_res_small': lambda: case_head(1, 3, 48, 6, 10, u8=False, residual=True),
    'eb': case_eb,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case')
    ap.add_argument('--only', default='')
    ap.add_argument('--timeout', type=int, default=180)
    args = ap.parse_args()
    if args.case:
        torch, *_ = _mods()
        ok = CASES[args.case]()
        torch.cuda.synchronize()
        print('RESULT', 'PASS' if ok else 'FAIL')
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    summary = {}
    for name in CASES:
        if args.only and args.only not in name:
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', name],
                               capture_output=True, text=True, timeout=args.timeout)
            status = 'PASS' if r.returncode == 0 else f'FAIL({r.returncode})'
            out = r.stdout + r.stderr
        except subprocess.TimeoutExpired as e:
            status, out = 'TIMEOUT', (e.stdout or b'').decode() + (e.stderr or b'').decode()
        summary[name] = dict(status=status, secs=round(time.time() - t0, 1), log=out[-3000:])
        print(f'== {name}: {status} ({summary[name]["secs"]}s)')
        print(out[-1500:])
        sys.stdout.flush()
    with open(os.path.join(ROOT, 'gpurun_out', 'bringup.json'), 'w') as f:
        json.dump(summary, f, indent=1)
    n_pass = sum(v['status'] == 'PASS' for v in summary.values())
    print(f'SUMMARY {n_pass}/{len(summary)} passed:',
          {k: v['status'] for k, v in summary.items() if v['status'] != 'PASS'})


if __name__ == '__main__':
    main()
