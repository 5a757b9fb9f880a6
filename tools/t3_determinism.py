"""Is the fp32-KV decode deterministic run to run?  Repeats the batch of tests/test_gpu_long.py::test_batch64_... and a set of
B=1 runs and reports every utterance whose ids differ between repetitions (a race shows up as run-to-run variation)."""
import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_long import _t3
g, sd, c3, t3, cond = _t3(os.path.join(ROOT, "tests", "golden"))
eng = t3.engine
gen = torch.Generator().manual_seed(2026)
B = 64
n_text = torch.randint(8, 90, (B,), generator=gen)
budgets = torch.randint(3, 70, (B,), generator=gen).tolist()
texts = [F.pad(F.pad(torch.randint(1, 255, (int(n),), generator=gen), (1, 0), value=255), (0, 1), value=0) for n in n_text]
cnd = t3.prepare_conditioning(cond)
kw = dict(cfg_weight=0.5, temperature=0.8, top_p=1.0, min_p=1.0, repetition_penalty=1.2, kv_dtype=os.environ.get("KV", "fp32"))
reps = int(os.environ.get("REPS", 8))
eng.decode_steps_per_call = 5
ref = None
for r in range(reps):
    out = eng.t3_generate(texts, cnd, max_new_tokens=budgets, **kw)
    if ref is None: ref = out; continue
    bad = [b for b in range(B) if not torch.equal(out[b], ref[b])]
    print(f"batch rep {r}: {len(bad)} utterances differ from rep 0 {bad[:8]}", flush=True)
eng.decode_steps_per_call = 16
nbad = 0
for b in range(0, B, 3):
    outs = [eng.t3_generate([texts[b]], cnd, max_new_tokens=budgets[b], **kw)[0] for _ in range(max(2, reps // 2))]
    var = any(not torch.equal(o, outs[0]) for o in outs[1:])
    eq = torch.equal(outs[0], ref[b])
    if var or not eq:
        nbad += 1
        first = next((i for i in range(min(len(outs[0]), len(ref[b]))) if outs[0][i] != ref[b][i]), -1)
        print(f"utt {b}: single runs vary={var} single==batch {eq} (first differing step {first} of {budgets[b]})", flush=True)
print("done: singles with an issue:", nbad)
