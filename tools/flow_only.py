"""Flow stage only (encoder + CFM) at a mid-size shape, for ncu captures of gemm_tc / attn_tc."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as W
from chatterbox_b200 import Engine, S3Gen
B = int(os.environ.get("FB", 16))
eng = Engine(0)
s3 = S3Gen(eng, W.make_flow_weights(0), None)
_, cg = W.make_conds(1234)
g = torch.Generator().manual_seed(3)
toks = [torch.randint(0, 6561, (int(n),), generator=g) for n in torch.randint(300, 1000, (B,), generator=g)]
mels = eng.flow_mel(toks, cg, n_timesteps=int(os.environ.get("NT", 2)))
torch.cuda.synchronize()
print("frames", sum(2 * (250 + t.numel()) for t in toks), "launches", eng.h.launch_count())
