"""Training-mode transforms on the repo's kernels (cae_act_grad, cae_conv_wgrad and the adjoint
cae_conv_igemm calls) against torch autograd on the same parameters (the reference's
formulation: nn.Conv2d / nn.ConvTranspose2d + LeakyReLU, src/models/tasks/_autoencoders.py:53-304
under src/train_cae_ms.py:209-219)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize('kind,c_in,c_out,h,w', [
    ('conv_s1', 128, 128, 32, 48), ('conv_s2', 128, 48, 32, 32), ('convt_s1', 48, 48, 20, 24),
    ('convt_s2', 48, 128, 16, 24), ('conv_s1', 32, 64, 16, 8), ('convt_s2', 128, 128, 24, 16),
    ('conv_s1', 3, 3, 32, 40), ('conv_s2', 3, 128, 32, 48), ('convt_s2', 128, 3, 16, 24)])
def test_single_layer_gradients_match_autograd(kind, c_in, c_out, h, w):
    from cnn_autoencoder_b200 import _engine as E, _train_conv as T
    import torch.nn as nn
    torch.manual_seed(3)
    if kind.startswith('convt'):
        s = 2 if kind.endswith('s2') else 1
        conv = nn.ConvTranspose2d(c_in, c_out, 3, stride=s, padding=1, output_padding=s - 1, bias=True)
    else:
        s = 2 if kind.endswith('s2') else 1
        conv = nn.Conv2d(c_in, c_out, 3, stride=s, padding=1, bias=True, padding_mode='reflect')
    conv = conv.cuda()
    st = E.Step(conv, pre_act='LeakyReLU')
    chain = T.WideChain([st])
    x = torch.randn(3, c_in, h, w, device='cuda', requires_grad=True)
    y = T._ChainFn.apply(x, chain, 1, conv.weight, conv.bias)
    # reference: the same fp16-rounded operands in fp32 arithmetic (TF32 off), and the SAME
    # activation mask (an output within rounding noise of zero may take the other LeakyReLU
    # branch, and a rare flipped mask entry moves a gradient sum by its whole term: that is
    # fp16-forward training, not a kernel error, and is bounded in the whole-model test below)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x16 = x.detach().half().float().requires_grad_(True)
    w16 = conv.weight.detach().half().float().requires_grad_(True)
    b32 = conv.bias.detach().clone().requires_grad_(True)
    if st.transposed:
        z = F.conv_transpose2d(x16, w16, b32, stride=conv.stride, padding=1, output_padding=s - 1)
    else:
        z = F.conv2d(F.pad(x16, (1, 1, 1, 1), mode='reflect'), w16, b32, stride=conv.stride)
    mask = torch.where(y.detach() > 0, 1.0, 0.01)
    y_ref = z * mask
    assert torch.allclose(y, y_ref, atol=2e-2, rtol=2e-2)
    g = torch.randn_like(y_ref) * 1e-3
    gx, gw, gb = torch.autograd.grad(y, [x, conv.weight, conv.bias], g)
    rx, rw, rb = torch.autograd.grad(y_ref, [x16, w16, b32], g)
    assert _rel(gw, rw) < 3e-3, ('dW', _rel(gw, rw))
    assert _rel(gb, rb) < 3e-3, ('db', _rel(gb, rb))
    assert _rel(gx, rx) < 3e-3, ('dx', _rel(gx, rx))


@pytest.mark.parametrize('arch', ['A', 'A_res', 'B'])
def test_whole_model_gradients_match_autograd(arch):
    """Encoder and decoder of a named net in train() mode: loss and every parameter gradient,
    kernels against the torch formulation on identical parameters and inputs."""
    from oracle import cae_oracle as O
    import cnn_autoencoder_b200 as M
    chk = O.make_checkpoint(dict(O.NAMED_ARCHS[arch], bias=True), seed=4)
    grads, losses = {}, {}
    x = (O.synth_natural(2, 3, 64, 96, seed=3).float() / 255.0).cuda()
    for mode in ('kernels', 'torch'):
        model = M.autoencoder_from_state_dict(chk, gpu=True, train=True)
        enc, dec = model['encoder'], model['decoder']
        enc.module.train_kernels = dec.module.train_kernels = mode == 'kernels'
        y = enc(x)
        x_r, _ = dec(y)
        loss = ((x_r[0] - x) ** 2).mean() + 1e-3 * y.abs().mean()
        loss.backward()
        losses[mode] = loss.item()
        grads[mode] = {n: p.grad.clone() for m in (enc, dec) for n, p in m.named_parameters()}
        if mode == 'kernels':
            assert enc.module._train_chain and dec.module._train_chain, 'the named nets are covered by the kernels'
    assert abs(losses['kernels'] - losses['torch']) <= 2e-3 * abs(losses['torch'])
    for n, gt in grads['torch'].items():
        assert _rel(grads['kernels'][n], gt) < 2e-2, (n, _rel(grads['kernels'][n], gt))


def test_train_step_matches_the_torch_formulation():
    """One whole rate-distortion step (train_cae_ms.py:209-230: forward closure, criterion, two
    backward passes, clip, Adam) with the transforms on the kernels against the same step with
    the transforms as torch autograd ops (the reference's formulation) from identical parameters,
    batch and noise: loss terms, the clipped gradients Adam saw (exp_avg after the first step =
    0.1 x gradient) and the post-step weights."""
    from oracle import cae_oracle as O
    import cnn_autoencoder_b200 as M
    chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=7)
    x = (O.synth_natural(4, 3, 128, 128, seed=5).float() / 255.0).cuda()
    res = {}
    for mode in ('kernels', 'torch'):
        model = M.autoencoder_from_state_dict(chk, gpu=True, train=True)
        for k in ('encoder', 'decoder'):
            model[k].module.train_kernels = mode == 'kernels'
        fwd = M.decorate_trainable_modules(trainable_modules=['encoder', 'decoder', 'fact_ent'],
                                           enabled_modules=['encoder', 'decoder', 'fact_ent'])
        crit = M.setup_loss('RateMSE', distortion_lambda=0.01)
        opts = M.setup_optimizers(model, lr=1e-4, aux_lr=1e-3)
        bucket = M.GradBucket(model)
        torch.manual_seed(11)                    # the bottleneck's additive noise
        out = M.train_step(x, model, crit, opts, fwd, bucket=bucket, step=0)
        torch.cuda.synchronize()
        res[mode] = dict(
            loss=float(torch.mean(out['loss'])), dist=float(torch.mean(out['dist_loss'])),
            rate=float(torch.mean(out['rate_loss'])),
            m={k: torch.cat([opts[k].state[p]['exp_avg'].reshape(-1) for p in opts[k].param_groups[0]['params']])
               for k in ('encoder', 'decoder')},
            w={k: torch.cat([p.detach().reshape(-1) for p in model[k].parameters()]).clone()
               for k in ('encoder', 'decoder')})
        if mode == 'kernels':
            assert model['encoder'].module._train_chain and model['decoder'].module._train_chain
    a, b = res['kernels'], res['torch']
    for key in ('loss', 'dist', 'rate'):
        assert abs(a[key] - b[key]) <= 2e-3 * abs(b[key]), (key, a[key], b[key])
    for k in ('encoder', 'decoder'):
        assert _rel(a['m'][k], b['m'][k]) < 3e-2, (k, _rel(a['m'][k], b['m'][k]))
        # Adam's first step moves every weight by +-lr: a gradient entry near zero may take the
        # other sign, never more than 2 lr away, and only a few of them
        d = (a['w'][k] - b['w'][k]).abs()
        assert float(d.max()) <= 2.05e-4, (k, float(d.max()))
        assert float((d > 2e-5).float().mean()) < 0.03, (k, float((d > 2e-5).float().mean()))


def test_graphed_train_step_tracks_the_eager_step():
    """GraphedTrainStep (the whole rate-distortion step replayed as one CUDA graph) against the
    eager train_step from the same start: the additive noise of the bottleneck comes from
    different generator offsets, so the two runs agree statistically, not bit for bit -- after
    the same number of Adam steps every weight sits within (steps x lr) of the eager run, the
    losses are finite and close, and a replay really advances the model."""
    from oracle import cae_oracle as O
    import cnn_autoencoder_b200 as M
    chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=9)
    x = (O.synth_natural(4, 3, 128, 128, seed=6).float() / 255.0).cuda()
    steps, warm = 3, 2

    def build():
        model = M.autoencoder_from_state_dict(chk, gpu=True, train=True)
        fwd = M.decorate_trainable_modules(trainable_modules=['encoder', 'decoder', 'fact_ent'],
                                           enabled_modules=['encoder', 'decoder', 'fact_ent'])
        crit = M.setup_loss('RateMSE', distortion_lambda=0.01)
        opts = M.setup_optimizers(model, lr=1e-4, aux_lr=1e-3, capturable=True)
        return model, fwd, crit, opts, M.GradBucket(model)

    torch.manual_seed(0)
    model, fwd, crit, opts, bucket = build()
    for k in range(warm + steps):            # (capturing records a step, it does not run one)
        out = M.train_step(x, model, crit, opts, fwd, bucket=bucket, step=0)
    loss_e = float(torch.mean(out['loss']).detach())
    w_e = torch.cat([p.detach().reshape(-1) for k in ('encoder', 'decoder') for p in model[k].parameters()])

    torch.manual_seed(0)
    model, fwd, crit, opts, bucket = build()
    w0 = torch.cat([p.detach().reshape(-1) for k in ('encoder', 'decoder') for p in model[k].parameters()]).clone()
    g = M.GraphedTrainStep(x, model, crit, opts, fwd, bucket, warmup=warm)     # warm steps so far
    before = torch.cat([p.detach().reshape(-1) for k in ('encoder', 'decoder') for p in model[k].parameters()]).clone()
    for k in range(steps):
        out = g(x)
    torch.cuda.synchronize()
    loss_g = float(torch.mean(out['loss']))
    w_g = torch.cat([p.detach().reshape(-1) for k in ('encoder', 'decoder') for p in model[k].parameters()])
    n_steps = warm + steps
    assert loss_g == loss_g and abs(loss_g - loss_e) <= 0.05 * abs(loss_e), (loss_g, loss_e)
    assert float((w_g - before).abs().max()) > 1e-5, 'a replay must move the weights'
    assert float((w_g - w0).abs().max()) <= 1.05e-4 * n_steps
    assert float((w_g - w_e).abs().max()) <= 2.1e-4 * n_steps
    assert float((w_g - w_e).abs().mean()) <= 0.5e-4 * n_steps
