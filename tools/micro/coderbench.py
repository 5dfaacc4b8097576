"""Device entropy coder: kernel time (CUDA events) against the number of concurrent tile
streams, warp-staged kernels (default) vs. the round-1 one-thread-per-stream kernels
(``CAE_DEBUG=1 CAE_RANS_V1=1`` in a second process), and a bit-exactness check against the host
coder.  Usage: python tools/micro/coderbench.py [n ...]"""
import ctypes, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cae_oracle as O
import cnn_autoencoder_b200 as M
from cnn_autoencoder_b200 import _cabi as C
from cnn_autoencoder_b200._entropy import encode_symbols

chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
fe = M.autoencoder_from_state_dict(chk, gpu=True, train=False)['fact_ent'].module
c, h, w = 48, 64, 64          # the latent of one 512 x 512 tile of net A
mp_per_tile = 512 * 512 / 1e6
tag = 'v1' if os.environ.get('CAE_RANS_V1') else 'v2'
for n in [int(a) for a in sys.argv[1:]] or (128, 1024, 2048, 4096):
    g = torch.Generator(device='cuda').manual_seed(n)
    sym = torch.round(torch.randn(n, c, h * w, generator=g, device='cuda') * 3).int()
    sym[0, 0, :8] = torch.tensor([-400, 300, 70000, -70000, 11, -11, 12, -12], device='cuda')  # escapes
    dev = sym.device
    cdf, sizes, offs, table = fe._dev_tables(dev)
    cap = c * h * w + 64
    words = torch.empty((n, cap), dtype=torch.int32, device=dev)
    nwords = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    L = C.lib()
    def enc():
        C.check(L.cae_rans_encode_batch(sym.data_ptr(), n, c, h * w, cdf.data_ptr(), cdf.shape[1],
                                        sizes.data_ptr(), offs.data_ptr(), table.data_ptr(),
                                        words.data_ptr(), cap, nwords.data_ptr(), status.data_ptr(), st))
    enc(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); enc(); e1.record(); torch.cuda.synchronize()
    t_enc = e0.elapsed_time(e1)
    packed, off = fe.encode_symbols_device(sym)
    host = packed.cpu().numpy()
    ref0 = encode_symbols(sym[0].cpu().numpy(), *fe._host_tables())
    same = host[int(off[0]):int(off[1])].tobytes() == ref0
    wordsd = packed.view(torch.int32)
    def dec():
        return fe.decode_streams_device(wordsd, off // 4, h * w)
    dec(); torch.cuda.synchronize()
    e0.record(); back = dec(); e1.record(); torch.cuda.synchronize()
    t_dec = e0.elapsed_time(e1)
    print(f'{tag} n={n:5d} encode {t_enc:8.2f} ms {n * mp_per_tile / t_enc:7.2f} GP/s | decode {t_dec:8.2f} ms '
          f'{n * mp_per_tile / t_dec:7.2f} GP/s | {int(off[-1]) * 8 / (n * c * h * w):.2f} bits/sym '
          f'host-identical={same} roundtrip={bool(torch.equal(back.reshape(sym.shape), sym))}', flush=True)
