"""GPU parity at the BENCHMARKED configuration: contexts beyond 1000 tokens, bf16 and fp32 KV,
CFM sequences beyond 2000 frames, a 64-utterance mixed-length batch through device-side retirement.

Fixtures: tests/golden/t3_long_golden.pt and flow_long_golden.pt are outputs of the unmodified reference
(oracle/make_golden.py long_t3 / long_flow: `T3.inference` for 900 greedy steps + a teacher-forced pass of the reference
backbone; `CausalMaskedDiffWithXvec.inference` at T = 2040 frames)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

_cache = {}


def _t3(golden_dir):
    if "t3" not in _cache:
        from gpu_util import engine
        from oracle import weights as W
        from chatterbox_b200.t3 import T3, T3Cond
        g = torch.load(os.path.join(golden_dir, "t3_long_golden.pt"))
        sd = W.make_t3_weights(g["weights_seed"])
        c3, _ = W.make_conds(g["conds_seed"])
        t3 = T3(engine(), sd)
        cond = T3Cond(speaker_emb=c3["speaker_emb"], cond_prompt_speech_tokens=c3["cond_prompt_speech_tokens"],
                      emotion_adv=c3["emotion_adv"])
        _cache["t3"] = (g, sd, c3, t3, cond)
    return _cache["t3"]


def _forced(t3, cond, g, n, kv_dtype, act_dtype=None):
    """Teacher-forced run over the reference's first n ids; returns (logits of both CFG rows after n tokens, the ids the
    engine's own greedy pick would have been).  The utterance is run as a batch of 6 identical copies (12 CFG rows): above
    8 rows the decode step takes the tensor-core path (split-K projections, plane / fp16 operands) that the bench uses;
    a single utterance would exercise the GEMV kernels instead (covered by tests/test_gpu_t3.py)."""
    eng = t3.engine
    cnd = t3.prepare_conditioning(cond)
    ids = g["tokens"][0]
    nb = 6
    out, st = eng.t3_generate([g["text_tokens"][0]] * nb, cnd, max_new_tokens=n, cfg_weight=0.5, temperature=0.8, top_p=1.0,
                              min_p=1.0, repetition_penalty=1.2, kv_dtype=kv_dtype, return_state=True,
                              force_tokens=[ids[:n]] * nb, act_dtype=act_dtype)
    torch.cuda.synchronize()
    assert out[0].tolist() == ids[:n].tolist()                  # the forced ids were recorded as the utterance's tokens
    return st["logits"][:2, :8194].cpu(), st["sampled"][0, :n].cpu()


@pytest.mark.parametrize("kv_dtype,act_dtype,tol", [("fp32", "bf16x2", 2e-3), ("bf16", "bf16x2", 2e-2), ("bf16", "fp16", 6e-2),
                                                    ("fp8", "fp16", 1.0)])
def test_t3_teacher_forced_logits_at_long_context(golden_dir, kv_dtype, act_dtype, tol):
    """Logits of both CFG rows after 1 / 64 / 256 / 512 / 768 / 900 generated tokens (context 189 .. 1088) against the
    reference backbone fed the same ids.  fp32 KV + hi/lo activation planes: fp32-faithful; bf16 KV: the error of rounding
    K/V to 8 mantissa bits; bf16 KV + one fp16 activation plane = the BENCH configuration; fp8 KV: opt-in."""
    g, sd, c3, t3, cond = _t3(golden_dir)
    worst = 0.0
    for n, ref in sorted(g["taps"].items()):
        logits, _ = _forced(t3, cond, g, n, kv_dtype, act_dtype)
        err = (logits - ref).abs().max().item()
        worst = max(worst, err)
        print(f"[long t3 kv {kv_dtype} act {act_dtype}] after {n} tokens (context {g['s0'] + n}): max|dlogit| = {err:.3e} (|logit| max {ref.abs().max():.2f})")
        assert err < tol, f"{kv_dtype} KV / {act_dtype} activations, {n} tokens: max|dlogit|={err}"


@pytest.mark.parametrize("kv_dtype,act_dtype,min_rate", [("fp32", "bf16x2", 0.998), ("bf16", "bf16x2", 0.99), ("bf16", "fp16", 0.98),
                                                         ("fp8", "fp16", 0.80)])
def test_t3_teacher_forced_argmax_agreement(golden_dir, kv_dtype, act_dtype, min_rate):
    """Over 900 teacher-forced steps the engine's own greedy pick equals the reference's id at (almost) every step;
    a disagreement can only be a near-tie of the top-2 logits."""
    g, sd, c3, t3, cond = _t3(golden_dir)
    n = g["tokens"].shape[1]
    _, sampled = _forced(t3, cond, g, n, kv_dtype, act_dtype)
    agree = (sampled == g["tokens"][0].to(torch.int32)).float().mean().item()
    print(f"[long t3 kv {kv_dtype} act {act_dtype}] teacher-forced argmax agreement over {n} steps: {agree:.4f}")
    assert agree >= min_rate, f"{kv_dtype} KV / {act_dtype}: agreement {agree}"


def test_t3_free_running_900_steps_fp32_kv(golden_dir):
    """Free-running greedy decode for 900 steps (fp32 KV): the ids equal the reference's own `T3.inference` output for
    at least the first 256 steps (an fp32 near-tie later on may fork the sequence; the teacher-forced tests above cover
    every step independently of such a fork)."""
    g, sd, c3, t3, cond = _t3(golden_dir)
    toks = t3.inference(t3_cond=cond, text_tokens=g["text_tokens"], max_new_tokens=900, temperature=0.8, top_p=1.0,
                        min_p=1.0, repetition_penalty=1.2, cfg_weight=0.5, kv_dtype="fp32").cpu()
    ref = g["tokens"]
    n = min(toks.shape[1], ref.shape[1])
    same = (toks[0, :n] == ref[0, :n]).int()
    prefix = int(same.cumprod(0).sum())
    print(f"[long t3] free-running fp32-KV greedy: {prefix} of {ref.shape[1]} leading ids equal the reference's")
    assert prefix >= 256, f"only {prefix} leading ids agree"


def test_batch64_mixed_lengths_equals_single_runs(golden_dir):
    """64 utterances of mixed text length and budget in one batch (paged KV, device-side retirement while others keep
    decoding, capacity shrinking through the graph buckets) produce exactly the ids of 64 independent B=1 runs."""
    import torch.nn.functional as F
    g, sd, c3, t3, cond = _t3(golden_dir)
    eng = t3.engine
    gen = torch.Generator().manual_seed(2026)
    B = 64
    n_text = torch.randint(8, 90, (B,), generator=gen)
    budgets = torch.randint(3, 70, (B,), generator=gen).tolist()
    texts = [F.pad(F.pad(torch.randint(1, 255, (int(n),), generator=gen), (1, 0), value=255), (0, 1), value=0) for n in n_text]
    cnd = t3.prepare_conditioning(cond)
    kw = dict(cfg_weight=0.5, temperature=0.8, top_p=1.0, min_p=1.0, repetition_penalty=1.2, kv_dtype="fp32")
    eng.decode_steps_per_call = 5
    try:
        batch = eng.t3_generate(texts, cnd, max_new_tokens=budgets, **kw)
    finally:
        eng.decode_steps_per_call = 16
    assert [len(b) for b in batch] == budgets or all(len(b) <= m for b, m in zip(batch, budgets))
    for b in range(0, B, 3):                      # every third utterance on its own (22 single runs)
        single = eng.t3_generate([texts[b]], cnd, max_new_tokens=budgets[b], **kw)[0]
        assert torch.equal(batch[b], single), (b, batch[b].tolist(), single.tolist())


def test_eos_retirement_does_not_disturb_neighbours(golden_dir):
    """Device-side EOS retirement: a 7-utterance batch is teacher-forced with its own greedy ids, except that three
    utterances are fed EOS (6562) at different steps.  Those stop right there (ids end with the EOS); every other
    utterance keeps producing, step for step, exactly the ids of the undisturbed run while its neighbours leave the
    active list and the slots are re-packed on the device."""
    import torch.nn.functional as F
    g, sd, c3, t3, cond = _t3(golden_dir)
    eng = t3.engine
    gen = torch.Generator().manual_seed(7)
    B, steps = 7, 30
    texts = [F.pad(F.pad(torch.randint(1, 255, (12 + 5 * i,), generator=gen), (1, 0), value=255), (0, 1), value=0) for i in range(B)]
    cnd = t3.prepare_conditioning(cond)
    kw = dict(cfg_weight=0.5, temperature=0.8, top_p=1.0, min_p=1.0, repetition_penalty=1.2, kv_dtype="fp32")
    free = eng.t3_generate(texts, cnd, max_new_tokens=steps, **kw)
    assert all(len(t) == steps for t in free)
    forced = [t.clone() for t in free]
    eos_at = {1: 4, 3: 17, 6: 9}
    for b, k in eos_at.items():
        forced[b][k] = 6562
    eng.decode_steps_per_call = 3
    try:
        out, st = eng.t3_generate(texts, cnd, max_new_tokens=steps, return_state=True, force_tokens=forced, **kw)
    finally:
        eng.decode_steps_per_call = 16
    sampled = st["sampled"].cpu()
    for b in range(B):
        if b in eos_at:
            k = eos_at[b]
            assert out[b].tolist() == forced[b][:k + 1].tolist(), (b, out[b].tolist())
            assert sampled[b, :k + 1].tolist() == free[b][:k + 1].tolist()
        else:
            assert out[b].tolist() == free[b].tolist()
            assert sampled[b, :steps].tolist() == free[b].tolist(), (b, sampled[b, :steps].tolist(), free[b].tolist())
    # the three EOS utterances left the active list on the device; the four that ran to their budget are flagged done in
    # the last step and would be dropped by the next step's compaction
    assert int(st["n_act"].item()) == B - len(eos_at) and int(st["done"].sum().item()) == B


# ---------------------------------------------------------------------------------------------------------- flow
def test_cfm_mel_at_2040_frames(golden_dir):
    """10-step CFM at T = 2040 frames (32 query tiles x 32 key blocks per head in the tcgen05 attention; the bench's long
    utterances): mel RMS <= 1e-3 against the reference's own output, encoder mu on a 1-in-16 frame sample."""
    from gpu_util import engine
    from oracle import weights as W
    from chatterbox_b200.s3gen import S3Gen
    g = torch.load(os.path.join(golden_dir, "flow_long_golden.pt"))
    fsd = W.make_flow_weights(g["weights_seed"])
    hsd = W.make_hift_weights(g["weights_seed"])
    s3 = _cache.setdefault("s3", S3Gen(engine(), fsd, hsd))
    _, cg = W.make_conds(seed=1234, n_gen_prompt=g["n_prompt"])
    T = 2 * (g["n_prompt"] + g["n"])
    torch.manual_seed(g["rng_seed"])
    z = torch.randn(1, 80, T)
    assert torch.equal(z[..., :8], g["z_head"]) and abs(float(z.double().sum()) - g["z_sum"]) < 1e-6
    mus, _ = s3.engine.flow_mel([g["tokens"][0]], cg, return_mu=True)
    mu_err = (mus[0].cpu()[::16] - g["mu_sample"][0]).abs().max().item()
    print(f"[long flow] T={T}: encoder max|dmu| (sampled) = {mu_err:.3e}")
    assert mu_err < 4e-3, mu_err
    for fmt in ("bf16x3", "fp16", "fp16+act16"):
        s3.engine.set_attention_precision("fp16" if fmt != "bf16x3" else "bf16x3")
        s3.engine.set_cfm_activation_precision("fp16" if fmt == "fp16+act16" else "bf16x2")
        try:
            mel = s3.flow_inference(g["tokens"][0], ref_dict=cg, z=z[0]).cpu()
        finally:
            s3.engine.set_cfm_activation_precision("fp16")
            s3.engine.set_attention_precision("fp16")
        rms = ((mel - g["mel"]) ** 2).mean().sqrt().item()
        mx = (mel - g["mel"]).abs().max().item()
        print(f"[long flow] operands {fmt}: mel RMS {rms:.3e}, max {mx:.3e} (mel std {g['mel'].std():.3f})")
        assert mel.shape == g["mel"].shape and rms < 1e-3, f"{fmt}: mel RMS {rms}"
