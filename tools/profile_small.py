"""Small end-to-end pass (B=8, short budgets) for the ncu launch list: every kernel of the hot path appears with
bench-like shapes per row, but the run stays short enough to profile launch by launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as W  # noqa: E402
from chatterbox_b200 import ChatterboxTTS, Conditionals, T3, T3Cond, S3Gen, Engine  # noqa: E402

B = int(os.environ.get("PB", 8))
STEPS = int(os.environ.get("PSTEPS", 24))
eng = Engine(0)
t3 = T3(eng, W.make_t3_weights(0))
s3 = S3Gen(eng, W.make_flow_weights(0), W.make_hift_weights(0))
c3, cg = W.make_conds(1234)
tts = ChatterboxTTS(t3, s3, None, "cuda", Conditionals(T3Cond(**c3), cg))
g = torch.Generator().manual_seed(1)
texts = [torch.randint(1, 255, (int(n),), generator=g) for n in torch.randint(16, 160, (B,), generator=g)]
tm = {}
wavs = tts.generate_batch(texts, max_new_tokens=[STEPS + 2 * i for i in range(B)], timings=tm)
print({k: round(v, 2) if isinstance(v, float) else v for k, v in tm.items()}, "launches", eng.h.launch_count())
