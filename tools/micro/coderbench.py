"""Device entropy coder throughput against the number of concurrent tile streams."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cae_oracle as O
import cnn_autoencoder_b200 as M

chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
fe = M.autoencoder_from_state_dict(chk, gpu=True, train=False)['fact_ent'].module
c, h, w = 48, 64, 64          # the latent of one 512 x 512 tile of net A
mp_per_tile = 512 * 512 / 1e6
for n in (128, 1024, 4096):
    g = torch.Generator(device='cuda').manual_seed(n)
    sym = torch.round(torch.randn(n, c, h * w, generator=g, device='cuda') * 3).int()
    for env in ({}, {'CAE_RANS_NO_TABLE': '1'}):
        os.environ.pop('CAE_RANS_NO_TABLE', None)
        os.environ.update(env)
        if n > 1024 and env:
            continue
        fe.encode_symbols_gpu(sym[:8])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        packed, off = fe.encode_symbols_device(sym)
        torch.cuda.synchronize()
        print(f'  device part only n={n}: {time.perf_counter() - t0:.3f} s')
        t0 = time.perf_counter()
        streams = fe.encode_symbols_gpu(sym)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print(f'encode n={n:5d} {"division" if env else "table   "} {t1 - t0:7.3f} s '
              f'{n * mp_per_tile / (t1 - t0) / 1e3:7.2f} GP/s  {sum(map(len, streams)) * 8 / (n * c * h * w):.2f} bits/sym')
    os.environ.pop('CAE_RANS_NO_TABLE', None)
    t0 = time.perf_counter()
    back = fe.decode_streams_gpu(streams, h * w)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f'decode n={n:5d}          {t1 - t0:7.3f} s {n * mp_per_tile / (t1 - t0) / 1e3:7.2f} GP/s  ok={bool(torch.equal(back, sym))}')
