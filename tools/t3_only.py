"""T3 stage only at a bench-like shape, with the library's event timer on one kernel class (paged / gemm_tc / gemv)."""
import os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as W
from chatterbox_b200 import Engine, T3
B = int(os.environ.get("TB", 256))
STEPS = int(os.environ.get("TSTEPS", 400))
eng = Engine(0)
t3 = T3(eng, W.make_t3_weights(0))
c3, _ = W.make_conds(1234)
cond = eng.t3_cond(c3["speaker_emb"].reshape(1, 256), c3["cond_prompt_speech_tokens"].reshape(1, -1),
                   torch.as_tensor(c3["emotion_adv"]).reshape(-1)[:1].float())
g = torch.Generator().manual_seed(1)
texts = [F.pad(F.pad(torch.randint(1, 255, (int(n),), generator=g), (1, 0), value=255), (0, 1), value=0)
         for n in torch.randint(16, 160, (B,), generator=g)]
for cls in os.environ.get("TCLS", "paged,gemm_tc").split(","):
    eng.stats.update(paged_bytes=0.0, paged_launches=0, decode_steps=0, decode_row_steps=0)
    eng.h.set_option("time_kernel", cls)
    if cls == "none":        # graph path: one untimed pass first (graph capture / instantiation)
        eng.t3_generate(texts, cond, max_new_tokens=STEPS, seed=3, kv_dtype="bf16")
        eng.stats.update(paged_bytes=0.0, paged_launches=0, decode_steps=0, decode_row_steps=0)
    torch.cuda.synchronize(); t0 = time.time()
    toks = eng.t3_generate(texts, cond, max_new_tokens=STEPS, seed=3, kv_dtype="bf16")
    torch.cuda.synchronize(); dt = time.time() - t0
    ms, n, work = eng.h.timer_read()
    gbs = eng.stats["paged_bytes"] / 1e9 / (ms / 1e3) if cls == "paged" and ms > 0 else 0.0
    print(f"class={cls} wall={dt:.2f}s kernel_ms={ms:.1f} launches={n} paged_GBps={gbs:.0f} tflops={work/1e9/max(ms,1e-9):.1f} "
          f"steps={eng.stats['decode_steps']} mean_len={sum(int(t.numel()) for t in toks)/len(toks):.0f}", flush=True)
