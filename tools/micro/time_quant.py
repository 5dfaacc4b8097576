"""Times the analysis track with and without the fused quantizer (bring-up aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cae_oracle as O
import cnn_autoencoder_b200 as M

chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
model = M.autoencoder_from_state_dict(chk, gpu=True, train=False)
enc, eb = model['encoder'], model['fact_ent'].module
x = O.synth_natural(128, 3, 256, 256, seed=1).permute(0, 2, 3, 1).contiguous().cuda()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


y = enc(x)
print('encoder            %8.1f us' % timeit(lambda: enc(x)))
print('encoder + fused q  %8.1f us' % timeit(lambda: enc(x, quant=eb.quant_request())))
print('  no hist/rate     %8.1f us' % timeit(lambda: enc(x, quant=eb.quant_request())) if False else '', end='')
print('standalone quant   %8.1f us' % timeit(lambda: eb.quantize_rate(y)))
for knob in ('CAE_QUANT_NO_HIST', 'CAE_QUANT_NO_RATE', 'CAE_QUANT_NO_YQ'):
    os.environ[knob] = '1'
    print('fused, %-18s %8.1f us' % (knob, timeit(lambda: enc(x, quant=eb.quant_request()))))
    del os.environ[knob]
