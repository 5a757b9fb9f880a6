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
import time
for cls in os.environ.get("FCLS", "none").split(","):
    eng.h.set_option("time_kernel", cls)
    torch.cuda.synchronize(); t0 = time.time()
    mels = eng.flow_mel(toks, cg, n_timesteps=int(os.environ.get("NT", 2)))
    torch.cuda.synchronize(); dt = time.time() - t0
    ms, n, work = eng.h.timer_read()
    if cls == "all":
        print("  per class: " + ", ".join(f"{c} {r['ms']:.1f} ms / {r['n']}" for c, r in ((c, eng.h.timer_read_class(c)) for c in
              ("gemm_tc", "wres", "stream", "attn_tc", "flash")) if r["n"]), flush=True)
    print(f"class={cls} wall={dt*1e3:.1f}ms kernel_ms={ms:.1f} launches={n} tflops={work/1e9/max(ms,1e-9):.1f} "
          f"frames={sum(2 * (250 + t.numel()) for t in toks)} total_launches={eng.h.launch_count()}", flush=True)
