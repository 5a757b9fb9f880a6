"""HiFT stage only at a bench-like chunk (24 000 mel frames), per-kernel-class device time via the library's event timer."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as W
from chatterbox_b200 import Engine, S3Gen
eng = Engine(0)
s3 = S3Gen(eng, W.make_flow_weights(0), W.make_hift_weights(0))
g = torch.Generator().manual_seed(3)
T = [int(n) * 2 for n in torch.randint(75, 1000, (int(os.environ.get("HB", 24)),), generator=g)]
mels = [torch.randn(80, t, generator=g).cuda() * 0.5 - 2 for t in T]
for cls in ["none", "none", "gemm_tc"]:
    eng.h.set_option("time_kernel", cls)
    l0 = eng.h.launch_count()
    torch.cuda.synchronize(); t0 = time.time()
    eng.hift(mels, seed=1)
    torch.cuda.synchronize(); dt = time.time() - t0
    ms, n, work = eng.h.timer_read()
    print(f"class={cls} wall={dt*1e3:.1f}ms kernel_ms={ms:.1f} launches={n} total_launches={eng.h.launch_count()-l0} "
          f"tflops={work/1e9/max(ms,1e-9):.1f} frames={sum(T)} -> {612.3e6*sum(T)/1e12/dt:.1f} TFLOP/s algorithmic", flush=True)
