"""GPU parity of the T3 path (C ABI -> CUDA) against the oracle restatement and the reference's own outputs
(tests/golden/t3_golden.pt, produced by the real reference T3.inference)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

_cache = {}


def _setup(golden_dir):
    if "t3" not in _cache:
        from gpu_util import engine
        from oracle import weights as W
        from chatterbox_b200.t3 import T3, T3Cond
        g = torch.load(os.path.join(golden_dir, "t3_golden.pt"))
        sd = W.make_t3_weights(g["weights_seed"])
        c3, _ = W.make_conds(g["conds_seed"])
        t3 = T3(engine(), sd)
        cond = T3Cond(speaker_emb=c3["speaker_emb"], cond_prompt_speech_tokens=c3["cond_prompt_speech_tokens"],
                      emotion_adv=c3["emotion_adv"])
        _cache["t3"] = (g, sd, c3, t3, cond)
    return _cache["t3"]


def test_cond_encode_matches_reference(golden_dir):
    g, sd, c3, t3, cond = _setup(golden_dir)
    out = t3.prepare_conditioning(cond).cpu()
    ref = g["cases"][0]["cond_emb"]
    err = (out - ref).abs().max().item()
    assert out.shape == ref.shape and err < 2e-4, err       # fp32 reference vs split-bf16 tensor-core GEMMs


@pytest.mark.parametrize("kv_dtype,tol", [("fp32", 2e-3), ("bf16", 6e-2)])
def test_prefill_logits_match_reference(golden_dir, kv_dtype, tol):
    """logits of the last prefill position for both CFG rows vs the reference backbone."""
    g, sd, c3, t3, cond = _setup(golden_dir)
    for case in g["cases"]:
        tt = case["text_tokens"]
        cnd = t3.prepare_conditioning(cond)
        st = t3.engine.t3_generate([tt[0]], cnd, max_new_tokens=4, cfg_weight=0.5, kv_dtype=kv_dtype,
                                   return_state="prefill")
        torch.cuda.synchronize()
        logits = st["logits"][:2, :8194].cpu()
        err = (logits - case["prefill_logits"]).abs().max().item()
        assert err < tol, f"n_text={case['n_text']} max|dlogit|={err}"


def test_greedy_tokens_bit_exact(golden_dir):
    """min_p=1.0 greedy emulation: ids must equal the reference's, fp32 KV cache."""
    g, sd, c3, t3, cond = _setup(golden_dir)
    for case in g["cases"]:
        if case["min_p"] != 1.0:
            continue
        toks = t3.inference(t3_cond=cond, text_tokens=case["text_tokens"], max_new_tokens=case["steps"],
                            temperature=0.8, top_p=1.0, min_p=1.0, repetition_penalty=1.2, cfg_weight=0.5,
                            kv_dtype="fp32")
        assert torch.equal(toks.cpu(), case["tokens"]), (toks.cpu(), case["tokens"])


def test_sampled_tokens_bit_exact_with_injected_noise(golden_dir):
    """multinomial(p,1) == argmax(p/q): feeding the reference's Exp(1) draws reproduces its sampled ids."""
    g, sd, c3, t3, cond = _setup(golden_dir)
    case = [c for c in g["cases"] if c["min_p"] != 1.0][0]
    torch.manual_seed(case["rng_seed"])
    q = torch.stack([torch.empty(8194).exponential_(1) for _ in range(case["steps"])])
    toks = t3.inference(t3_cond=cond, text_tokens=case["text_tokens"], max_new_tokens=case["steps"], temperature=0.8,
                        top_p=1.0, min_p=case["min_p"], repetition_penalty=1.2, cfg_weight=0.5, q_noise=q, kv_dtype="fp32")
    assert torch.equal(toks.cpu(), case["tokens"]), (toks.cpu(), case["tokens"])


def test_bf16_kv_greedy_agreement(golden_dir):
    """bench configuration (bf16 KV cache): teacher-forcing is not available through the API, so require the
    free-running greedy ids to agree with the reference on a prefix (bf16 rounding of K/V may flip a near-tie)."""
    g, sd, c3, t3, cond = _setup(golden_dir)
    case = [c for c in g["cases"] if c["min_p"] == 1.0][0]
    toks = t3.inference(t3_cond=cond, text_tokens=case["text_tokens"], max_new_tokens=case["steps"], temperature=0.8,
                        top_p=1.0, min_p=1.0, repetition_penalty=1.2, cfg_weight=0.5, kv_dtype="bf16").cpu()
    same = (toks[0] == case["tokens"][0]).int()
    prefix = int(same.cumprod(0).sum())
    assert prefix >= 8, f"only {prefix} leading ids agree: {toks} vs {case['tokens']}"


def test_batch_equals_independent_runs(golden_dir):
    """B=3 mixed-length batch (paged KV, early retirement) == three B=1 runs == oracle."""
    from oracle.t3_ref import T3Oracle
    import torch.nn.functional as F
    g, sd, c3, t3, cond = _setup(golden_dir)
    eng = t3.engine
    gen = torch.Generator().manual_seed(99)
    texts, budgets = [], [5, 9, 7]
    for n in (11, 40, 23):
        t = torch.randint(1, 255, (n,), generator=gen)
        texts.append(F.pad(F.pad(t, (1, 0), value=255), (0, 1), value=0))
    cnd = t3.prepare_conditioning(cond)
    batch = eng.t3_generate(texts, cnd, max_new_tokens=budgets, cfg_weight=0.5, temperature=0.8, top_p=1.0, min_p=1.0,
                            repetition_penalty=1.2, kv_dtype="fp32", max_sync_steps=3)
    orc = T3Oracle(sd)
    for b in range(3):
        single = eng.t3_generate([texts[b]], cnd, max_new_tokens=budgets[b], cfg_weight=0.5, temperature=0.8,
                                 top_p=1.0, min_p=1.0, repetition_penalty=1.2, kv_dtype="fp32")[0]
        assert torch.equal(batch[b], single), (b, batch[b], single)
        ref = orc.inference(c3, torch.stack([texts[b], texts[b]]), budgets[b], temperature=0.8, top_p=1.0, min_p=1.0,
                            repetition_penalty=1.2, cfg_weight=0.5)[0]
        assert torch.equal(batch[b], ref), (b, batch[b], ref)


def test_top_p_path_matches_oracle(golden_dir):
    """top_p < 1 (T3.inference default 0.95): sort-based filter vs the restated TopPLogitsWarper."""
    from oracle.t3_ref import T3Oracle
    g, sd, c3, t3, cond = _setup(golden_dir)
    case = g["cases"][0]
    steps = 6
    torch.manual_seed(42)
    q = torch.stack([torch.empty(8194).exponential_(1) for _ in range(steps)])
    toks = t3.inference(t3_cond=cond, text_tokens=case["text_tokens"], max_new_tokens=steps, temperature=0.8, top_p=0.8,
                        min_p=0.0, repetition_penalty=1.2, cfg_weight=0.5, q_noise=q, kv_dtype="fp32").cpu()
    ref = T3Oracle(sd).inference(c3, case["text_tokens"], steps, temperature=0.8, top_p=0.8, min_p=0.0,
                                 repetition_penalty=1.2, cfg_weight=0.5, q_noise=q)
    assert torch.equal(toks, ref), (toks, ref)


def test_decode_graph_replay_matches_direct_launches(golden_dir):
    """CUDA-graph replay of the decode steps (the default) gives the same ids as direct launches."""
    import time
    g, sd, c3, t3, cond = _setup(golden_dir)
    case = [c for c in g["cases"] if c["min_p"] == 1.0][0]
    kw = dict(t3_cond=cond, text_tokens=case["text_tokens"], max_new_tokens=case["steps"], temperature=0.8, top_p=1.0,
              min_p=1.0, repetition_penalty=1.2, cfg_weight=0.5, kv_dtype="fp32")
    eng = t3.engine
    eng.set_decode_graph(False)
    try:
        t3.inference(**kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        direct = t3.inference(**kw).cpu()
        t_direct = time.perf_counter() - t0
    finally:
        eng.set_decode_graph(True)
    t3.inference(**kw)                                   # warm (graph instantiation)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    replay = t3.inference(**kw).cpu()
    t_graph = time.perf_counter() - t0
    print(f"decode {case['steps']} steps: direct {t_direct * 1e3:.1f} ms, graph {t_graph * 1e3:.1f} ms")
    assert torch.equal(direct, case["tokens"]) and torch.equal(replay, direct), (replay, direct)


def test_fp32_checkpoint_weight_rounding_is_bounded(golden_dir):
    """every other parity test uses bf16-representable weights, so the one deliberate approximation of the
    engine -- matmul weights rounded once to bf16 (relative 2^-9) -- never shows.  Here the checkpoint is NOT
    bf16-representable (fp32 draws, like a real checkpoint) and the fp32 oracle keeps the exact values: the engine's
    prefill logits stay within a small bound of the oracle's and the greedy ids agree on a prefix (a near-tie may fork
    later, exactly as two fp32 BLAS libraries would fork the reference against itself at a coarser scale)."""
    from gpu_util import engine
    from oracle import weights as W
    from oracle.t3_ref import T3Oracle
    from chatterbox_b200 import Engine, T3, T3Cond
    L = 30
    sd = W.make_t3_weights(3, n_layers=L, bf16=False)
    assert not torch.equal(sd["tfmr.layers.0.mlp.up_proj.weight"], sd["tfmr.layers.0.mlp.up_proj.weight"].bfloat16().float())
    c3, _ = W.make_conds(1234)
    eng = Engine(0)
    t3 = T3(eng, sd)
    cond = T3Cond(**c3)
    tt = torch.randint(1, 255, (40,), generator=torch.Generator().manual_seed(4))
    tt = torch.nn.functional.pad(torch.nn.functional.pad(tt, (1, 0), value=255), (0, 1), value=0)
    tt2 = torch.stack([tt, tt])
    steps = 24
    ref = T3Oracle(sd, n_layers=L).inference(c3, tt2, steps, temperature=0.8, top_p=1.0, min_p=1.0, repetition_penalty=1.2,
                                             cfg_weight=0.5)
    out, st = eng.t3_generate([tt], t3.prepare_conditioning(cond), max_new_tokens=steps, cfg_weight=0.5, temperature=0.8,
                              top_p=1.0, min_p=1.0, repetition_penalty=1.2, kv_dtype="fp32", return_state=True,
                              force_tokens=[ref[0]])
    sampled = st["sampled"][0, :steps].cpu()
    agree = (sampled == ref[0].to(torch.int32)).float().mean().item()
    # logits after the forced sequence vs an oracle built from the SAME weights rounded to bf16: isolates the weight rounding
    sd_r = {k: (v.bfloat16().float() if v.dim() == 2 and "emb" not in k and "norm" not in k and min(v.shape) >= 8 else v) for k, v in sd.items()}
    ref_r = T3Oracle(sd_r, n_layers=L).inference(c3, tt2, steps, temperature=0.8, top_p=1.0, min_p=1.0, repetition_penalty=1.2,
                                               cfg_weight=0.5)
    agree_r = (ref_r[0] == ref[0]).float().mean().item()
    print(f"[weights] fp32 (non-bf16-representable) checkpoint: engine greedy pick == fp32-oracle id on {agree:.3f} of {steps} "
          f"teacher-forced steps; a bf16-weight ORACLE free-runs to the same ids on {agree_r:.3f} of them")
    assert agree >= 0.9, agree
