"""Per-kernel-class device time of the flow stage (CFM + encoder) at a bench-like shape, via the library's event timer."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as W
from chatterbox_b200 import Engine, S3Gen
B = int(os.environ.get("FB", 32))
eng = Engine(0)
s3 = S3Gen(eng, W.make_flow_weights(0), W.make_hift_weights(0))
if os.environ.get("ATTN_PREC"):          # fp16 | bf16x3 (opt-in operand formats)
    eng.set_attention_precision(os.environ["ATTN_PREC"])
if os.environ.get("CFM_ACT"):            # fp16 | bf16x2
    eng.set_cfm_activation_precision(os.environ["CFM_ACT"])
_, cg = W.make_conds(1234)
g = torch.Generator().manual_seed(3)
toks = [torch.randint(0, 6561, (int(n),), generator=g) for n in torch.randint(75, 1000, (B,), generator=g)]
for cls in ["none", "flash", "gemm_tc"]:
    eng.h.set_option("time_kernel", cls)
    torch.cuda.synchronize(); t0 = time.time()
    l0 = eng.h.launch_count()
    mels = eng.flow_mel(toks, cg)
    torch.cuda.synchronize(); dt = time.time() - t0
    ms, n, work = eng.h.timer_read()
    print(f"class={cls} wall={dt:.2f}s kernel_ms={ms:.1f} launches={n} total_launches={eng.h.launch_count()-l0} tflops={work/1e9/max(ms,1e-9):.1f} frames={sum(2*(250+t.numel()) for t in toks)}", flush=True)
