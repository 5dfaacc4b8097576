"""One small invocation of the hot path on cuda:0, checked against the CPU oracle."""
import os
import sys

import numpy as np
import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import cae_oracle as OR        # the checker (allowed in smoke())
    import cnn_autoencoder_b200 as M
    from cnn_autoencoder_b200.pipeline import CodecPipeline

    if not torch.cuda.is_available():
        raise RuntimeError('smoke() needs a CUDA device')
    torch.cuda.set_device(0)
    arch = dict(channels_org=3, channels_net=32, channels_bn=16, compression_level=3,
                act_layer_type='LeakyReLU', use_residual=True)
    chk = OR.make_checkpoint(arch, seed=4321)
    model = M.autoencoder_from_state_dict(chk, gpu=True, train=False)
    oracle = OR.OracleModel(chk)
    x_u8 = OR.synth_natural(2, 3, 64, 96, seed=5)                       # N x C x H x W uint8
    x_hwc = x_u8.permute(0, 2, 3, 1).contiguous().cuda()
    out = CodecPipeline(model)(x_hwc)
    torch.cuda.synchronize()
    ref = oracle.forward(x_u8.float() / 255.0)
    sym = torch.round(out['y'].cpu() - oracle.fact_ent._medians().reshape(1, -1, 1, 1))
    sym_ref = torch.round(ref['y'] - oracle.fact_ent._medians().reshape(1, -1, 1, 1))
    agree = (sym == sym_ref).float().mean().item()
    ref_u8 = (ref['x_r'][0] * 255.0).clip(0, 255).to(torch.uint8).permute(0, 2, 3, 1).numpy()
    got_u8 = out['x_r_u8'].cpu().numpy()
    img = x_u8.permute(0, 2, 3, 1).numpy()
    d_psnr = abs(OR.psnr_u8(img, got_u8) - OR.psnr_u8(img, ref_u8))
    bpp_ref = OR.rate_loss(x_u8.float(), ref['p_y']).item()
    d_bpp = abs(out['bpp'].item() - bpp_ref) / bpp_ref
    print(f'smoke: symbol agreement {agree * 100:.3f}%  |dPSNR| {d_psnr:.4f} dB  '
          f'|dbpp| {d_bpp * 100:.3f}%  max|u8 diff| {np.abs(got_u8.astype(int) - ref_u8.astype(int)).max()}')
    assert agree >= 0.995, agree
    assert d_psnr <= 0.05, d_psnr
    assert d_bpp <= 0.005, d_bpp
    from cnn_autoencoder_b200 import _cabi
    assert _cabi.launch_count() > 0
