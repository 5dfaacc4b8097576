"""The rate-distortion training step (config 5) on CPU: two gloo ranks with one flat-bucket
gradient all-reduce reproduce the single-process step on the concatenated batch, and the
single-process step agrees with the oracle's loss terms."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cnn_autoencoder_b200 as M
from oracle import cae_oracle as O

ARCH = dict(channels_org=3, channels_net=8, channels_bn=8, compression_level=2,
            act_layer_type='LeakyReLU', use_residual=True)


def _build(seed=3):
    chk = O.make_checkpoint(ARCH, seed=seed)
    model = M.autoencoder_from_state_dict(chk, gpu=False, train=True)
    return chk, model


def _noise_for(shard, total):
    """Deterministic noise keyed on the GLOBAL sample index."""
    def fn(v):                       # v: C x 1 x (n_local * h * w)
        g = torch.Generator().manual_seed(1234)
        c = v.shape[0]
        per = v.shape[2] // (shard.stop - shard.start)
        full = torch.rand((c, 1, total, per), generator=g) - 0.5
        return full[:, :, shard].reshape(c, 1, -1)
    return fn


def _step(model, x, shard, total, group=None, bucket=False, steps=1, acc=None):
    model['fact_ent'].module.noise_fn = _noise_for(shard, total)
    fwd = M.decorate_trainable_modules(trainable_modules=['encoder', 'decoder', 'fact_ent'],
                                       enabled_modules=['encoder', 'decoder', 'fact_ent'])
    crit = M.setup_loss('RateMSE', distortion_lambda=0.01)
    opts = M.setup_optimizers(model, lr=1e-3, aux_lr=1e-2)
    b = M.GradBucket(model) if bucket else None
    for s in range(steps):
        out = M.train_step(x, model, crit, opts, fwd, group=group, bucket=b, step=s,
                           mod_grad_accumulate=acc)
    return out


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    _, model = _build()
    x = O.synth_natural(4, 3, 32, 32, seed=9).float() / 255.0
    shard = slice(rank * 2, rank * 2 + 2)
    _step(model, x[shard], shard, 4, bucket=True)      # persistent bucket, overlapped all-reduce
    if rank == 0:
        torch.save({k: v.state_dict() for k, v in model.items()}, out)
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_step(tmp_path):
    torch.set_num_threads(1)
    _, ref = _build()
    x = O.synth_natural(4, 3, 32, 32, seed=9).float() / 255.0
    before = {k: {n: p.detach().clone() for n, p in v.named_parameters()} for k, v in ref.items()}
    _step(ref, x, slice(0, 4), 4)
    out = str(tmp_path / 'rank0.pt')
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    got = torch.load(out)
    moved = 0
    for k in ref:
        for n, p in ref[k].state_dict().items():
            assert torch.allclose(got[k][n].float(), p.float(), rtol=2e-4, atol=2e-6), (k, n)
        for n, p in ref[k].named_parameters():
            moved += int(not torch.equal(p, before[k][n]))
    assert moved > 10                      # the step really updated the parameters


def test_train_mode_loss_terms_match_oracle():
    chk, model = _build(seed=5)
    for m in model.values():
        m.eval()                            # deterministic quantizer for the comparison
    x = O.synth_natural(2, 3, 32, 32, seed=4).float() / 255.0
    with torch.no_grad():
        y = model['encoder'].module.analysis_track(x)
        y_q, p_y = model['fact_ent'].module._forward_torch(y)
        x_r = model['decoder'].module.synthesis_track(y_q)
    oracle = O.OracleModel(chk)
    ref = oracle.forward(x)
    assert torch.allclose(y, ref['y'], atol=1e-6)
    assert torch.allclose(x_r, ref['x_r'][0], atol=1e-6)
    crit = M.setup_loss('RateMSE', distortion_lambda=0.01)
    loss = crit(inputs=x, outputs=dict(x_r=[x_r], p_y=p_y), net=model)
    want = O.general_loss(x, ref, oracle.fact_ent, distortion_lambda=0.01)
    for k in ('loss', 'rate_loss', 'dist_loss', 'entropy_loss'):
        assert torch.allclose(loss[k], want[k], rtol=1e-5), k


def test_allreduce_is_a_noop_without_a_process_group():
    _, model = _build()
    assert M.allreduce_gradients(model) == 0


def test_persistent_bucket_equals_the_gathered_bucket():
    torch.set_num_threads(1)
    x = O.synth_natural(4, 3, 32, 32, seed=9).float() / 255.0
    _, a = _build()
    _, b = _build()
    _step(a, x, slice(0, 4), 4, bucket=False, steps=2)
    _step(b, x, slice(0, 4), 4, bucket=True, steps=2)
    for k in a:
        for (n, p), (_, q) in zip(a[k].named_parameters(), b[k].named_parameters()):
            assert torch.equal(p, q), (k, n)
    bucket = M.GradBucket(b)
    assert bucket.attached() and bucket.numel() == sum(p.numel() for k in b for p in b[k].parameters())
    assert list(bucket.ranges) == ['decoder', 'encoder', 'fact_ent', 'fact_ent_aux']


def test_mod_grad_accumulate_gates_the_optimizers():
    """train_cae_ms.py:221-222: an optimizer clips / steps / zeroes only when step % period == 0,
    its gradients accumulating in between."""
    torch.set_num_threads(1)
    x = O.synth_natural(2, 3, 32, 32, seed=2).float() / 255.0
    _, m = _build()
    before = {k: [p.detach().clone() for p in v.parameters()] for k, v in m.items()}
    fwd = M.decorate_trainable_modules(trainable_modules=['encoder', 'decoder', 'fact_ent'],
                                       enabled_modules=['encoder', 'decoder', 'fact_ent'])
    crit = M.setup_loss('RateMSE', distortion_lambda=0.01)
    opts = M.setup_optimizers(m, lr=1e-3, aux_lr=1e-2)
    bucket = M.GradBucket(m)
    m['fact_ent'].module.noise_fn = _noise_for(slice(0, 2), 2)
    acc = {'encoder': 2, 'decoder': 1, 'fact_ent': 1, 'fact_ent_aux': 1}
    M.train_step(x, m, crit, opts, fwd, bucket=bucket, step=1, mod_grad_accumulate=acc)
    enc_same = all(torch.equal(p, q) for p, q in zip(m['encoder'].parameters(), before['encoder']))
    dec_moved = any(not torch.equal(p, q) for p, q in zip(m['decoder'].parameters(), before['decoder']))
    assert enc_same and dec_moved
    a, b = bucket.ranges['encoder']
    g1 = bucket.flat[a:b].clone()
    assert g1.abs().sum() > 0                                   # kept for the next step
    a2, b2 = bucket.ranges['decoder']
    assert bucket.flat[a2:b2].abs().sum() == 0                  # zeroed after its step
    M.train_step(x, m, crit, opts, fwd, bucket=bucket, step=2, mod_grad_accumulate=acc)
    assert any(not torch.equal(p, q) for p, q in zip(m['encoder'].parameters(), before['encoder']))
    assert bucket.flat[a:b].abs().sum() == 0


def test_training_chain_eligibility_is_decided_from_the_track():
    """Which tracks train on the kernels (``_train_conv.eligible_chain``, host logic only): the
    named nets A, A+residual and B in full; BatchNorm, GDN and groups=True fall back to the torch
    formulation; a ReLU directly after a residual add does too (not invertible)."""
    from cnn_autoencoder_b200 import _train_conv as T
    import cnn_autoencoder_b200 as M

    def tracks(**arch):
        chk = O.make_checkpoint(dict(O.NAMED_ARCHS['A'], **arch), seed=1)
        model = M.autoencoder_from_state_dict(chk, gpu=False, train=True)
        return model['encoder'].module, model['decoder'].module

    for arch in (dict(), dict(use_residual=True),
                 dict(use_residual=True, channels_bn=192, compression_level=4)):
        for tr in tracks(**arch):
            found = T.eligible_chain(tr)
            assert found is not None, arch
            steps, k0, k1 = found
            assert (k0, k1) == (0, len(steps)), (arch, k0, k1)      # the whole track, thin layers included
    for arch in (dict(batch_norm=True), dict(act_layer_type='GDN'), dict(groups=True, channels_net=126, channels_bn=48),
                 dict(use_residual=True, act_layer_type='ReLU')):
        try:
            trs = tracks(**arch)
        except Exception:          # an architecture the constructors refuse is not a chain either
            continue
        enc, dec = trs
        assert T.eligible_chain(enc) is None or T.eligible_chain(dec) is None, arch
