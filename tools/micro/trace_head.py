import os, sys, torch
sys.path.insert(0, '/root/repo')
from oracle import cae_oracle as O
import cnn_autoencoder_b200 as M
from cnn_autoencoder_b200.pipeline import CodecPipeline
chk = O.make_checkpoint(O.NAMED_ARCHS['A'], seed=1234)
model = M.autoencoder_from_state_dict(chk, gpu=True, train=False)
pipe = CodecPipeline(model)
x = O.synth_natural(128, 3, 256, 256, seed=1).permute(0, 2, 3, 1).contiguous().cuda()
for _ in range(3): pipe(x)
torch.cuda.synchronize()
os.environ['CAE_HEAD_TRACE'] = '1'
pipe(x)
torch.cuda.synchronize()
