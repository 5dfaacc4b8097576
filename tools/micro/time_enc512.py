import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cae_oracle as O
import cnn_autoencoder_b200 as M
chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
model = M.autoencoder_from_state_dict(chk, gpu=True, train=False)
enc, eb = model['encoder'], model['fact_ent'].module
x = torch.randint(0, 255, (32, 512, 512, 3), dtype=torch.uint8, device='cuda')
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
print('encoder                 %.3f ms' % timeit(lambda: enc(x)))
print('encoder + fused (sym)   %.3f ms' % timeit(lambda: enc(x, quant=eb.quant_request(want_sym=True, want_planar=False))))
print('encoder + fused (plan)  %.3f ms' % timeit(lambda: enc(x, quant=eb.quant_request())))
y = enc(x)
print('standalone sym          %.3f ms' % timeit(lambda: eb._quantize_cuda(y, want_yq=False, want_p=False, want_sym=True)))
