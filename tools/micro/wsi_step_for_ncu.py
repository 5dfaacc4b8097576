"""One small WSI device round trip (256 chunks of 512 x 512, eager launches) for the ncu launch
list / full captures: every kernel of the path runs a few times, nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from oracle import cae_oracle as O
from cnn_autoencoder_b200 import compress as CMP, _slide

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
model = CMP.load_model(chk)
tc = _slide.TileCodec(model, 512, 3, 32, graphs=False)
x = torch.from_numpy(np.stack([bench.slide_tile(O, i // 64, i % 64) for i in range(T)])).cuda()
out = torch.empty_like(x)
for _ in range(2):
    _slide.device_roundtrip(tc, x, out, T)
torch.cuda.synchronize()
print('ok', T)
