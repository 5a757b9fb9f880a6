"""GPU parity of the Turbo path: GPT-2 backbone T3 + `inference_turbo` sampler through the C ABI,
against the reference's own outputs (tests/golden/turbo_golden.pt, written by the real `T3(hp).inference_turbo`) and the
oracle restatement (oracle/t3_turbo_ref.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

_cache = {}


def _setup(golden_dir):
    if "t" not in _cache:
        from oracle import weights as W
        from chatterbox_b200 import Engine, T3, T3Cond
        g = torch.load(os.path.join(golden_dir, "turbo_golden.pt"))
        sd = W.make_t3_turbo_weights(g["weights_seed"], text_vocab=g["text_vocab"])
        c3, cg = W.make_conds(g["conds_seed"], n_t3_prompt=g["n_prompt"], n_gen_prompt=40)
        eng = Engine(0)                   # own engine: the handle holds one T3 backbone
        t3 = T3(eng, sd)
        assert eng.t3_turbo and eng.t3_layers == 24
        cond = T3Cond(speaker_emb=c3["speaker_emb"], cond_prompt_speech_tokens=c3["cond_prompt_speech_tokens"],
                      emotion_adv=c3["emotion_adv"])
        _cache["t"] = (g, sd, c3, cg, eng, t3, cond)
    return _cache["t"]


def test_turbo_cond_encode_matches_reference(golden_dir):
    g, sd, c3, cg, eng, t3, cond = _setup(golden_dir)
    out = t3.prepare_conditioning(cond).cpu()
    case = g["cases"][0]
    assert out.shape == (1, case["len_cond"], 1024)
    assert (out[:, :4] - case["cond_emb_head"]).abs().max().item() < 2e-4
    emb = sd["speech_emb.weight"][c3["cond_prompt_speech_tokens"][0]]
    assert torch.equal(out[0, 1:], emb)                      # prompt rows are plain gathers


@pytest.mark.parametrize("kv_dtype,tol", [("fp32", 2e-3), ("bf16", 6e-2)])
def test_turbo_prefill_logits_match_reference(golden_dir, kv_dtype, tol):
    g, sd, c3, cg, eng, t3, cond = _setup(golden_dir)
    cnd = t3.prepare_conditioning(cond)
    for case in g["cases"][1:3]:
        st = eng.t3_generate([case["text_tokens"][0]], cnd, max_new_tokens=4, cfg_weight=0.0, kv_dtype=kv_dtype,
                             return_state="prefill", top_k=1)
        torch.cuda.synchronize()
        logits = st["logits"][:1, :6563].cpu()
        err = (logits - case["prefill_logits"]).abs().max().item()
        assert err < tol, f"n_text={case['n_text']} max|dlogit|={err}"


def test_turbo_greedy_tokens_bit_exact(golden_dir):
    """top_k=1 greedy emulation: ids equal the reference's, fp32 KV cache."""
    g, sd, c3, cg, eng, t3, cond = _setup(golden_dir)
    for case in g["cases"]:
        if case["top_k"] != 1:
            continue
        toks = t3.inference_turbo(cond, case["text_tokens"], temperature=0.8, top_k=1, top_p=case["top_p"],
                                  repetition_penalty=case["rep"], max_gen_len=case["steps"], kv_dtype="fp32").cpu()
        assert torch.equal(toks, case["tokens"]), (toks, case["tokens"])


def test_turbo_sampled_tokens_bit_exact_with_injected_noise(golden_dir):
    """temperature -> top-k -> top-p -> repetition penalty -> multinomial: feeding the reference's Exp(1) draws
    (multinomial(p,1) == argmax(p/q)) reproduces its sampled ids, for top_k 1000 / top_p 0.95 and top_k 50 / top_p 0.8."""
    g, sd, c3, cg, eng, t3, cond = _setup(golden_dir)
    for case in g["cases"]:
        if case["top_k"] == 1:
            continue
        torch.manual_seed(case["rng_seed"])
        q = torch.stack([torch.empty(6563).exponential_(1) for _ in range(case["steps"] + 1)])
        toks = t3.inference_turbo(cond, case["text_tokens"], temperature=0.8, top_k=case["top_k"], top_p=case["top_p"],
                                  repetition_penalty=case["rep"], max_gen_len=case["steps"], q_noise=q,
                                  kv_dtype="fp32").cpu()
        assert torch.equal(toks, case["tokens"]), (toks, case["tokens"])


def test_turbo_batch_equals_single(golden_dir):
    """A batch of Turbo utterances (one row each, no CFG) equals running them one by one (greedy, fp32 KV)."""
    g, sd, c3, cg, eng, t3, cond = _setup(golden_dir)
    cnd = t3.prepare_conditioning(cond)
    cases = [c for c in g["cases"] if c["top_k"] == 1 and c["rep"] == 1.2] * 2 + [g["cases"][2]]
    texts = [c["text_tokens"][0] for c in cases]
    budgets = [5, 9, 7]
    out = eng.t3_generate(texts, cnd, max_new_tokens=budgets, cfg_weight=0.0, temperature=0.8, top_p=0.95, min_p=0.0,
                          repetition_penalty=1.2, kv_dtype="fp32", top_k=1)
    for t, b, o in zip(texts, budgets, out):
        single = eng.t3_generate([t], cnd, max_new_tokens=b, cfg_weight=0.0, temperature=0.8, top_p=0.95, min_p=0.0,
                                 repetition_penalty=1.2, kv_dtype="fp32", top_k=1)[0]
        assert torch.equal(o, single), (o, single)


def test_turbo_generate_matches_oracle_pipeline(golden_dir):
    """ChatterboxTurboTTS.generate_tokens vs the oracle pipeline with the same torch seed: Turbo T3 -> +3 silence
    tokens -> 2-step meanflow CFM -> HiFT (reference tts_turbo.py:272-321)."""
    from oracle import weights as W
    from oracle.t3_turbo_ref import TurboOracle
    from oracle.flow_ref import FlowOracle
    from oracle.hift_ref import HiFTOracle
    from chatterbox_b200 import ChatterboxTurboTTS, Conditionals, S3Gen, T3Cond
    g, sd, c3, cg, eng, t3, cond = _setup(golden_dir)
    fsd, hsd = W.make_flow_weights(0, meanflow=True), W.make_hift_weights(0)
    tts = ChatterboxTurboTTS(t3, S3Gen(eng, fsd, hsd, meanflow=True), None, "cuda", Conditionals(T3Cond(**c3), dict(cg)))
    text = g["cases"][0]["text_tokens"]
    steps = 14
    torch.manual_seed(777)
    wav, mid = tts.generate_tokens(text, max_gen_len=steps, rng="torch_cpu", kv_dtype="fp32", return_intermediates=True)
    torch.manual_seed(777)
    toks = TurboOracle(sd).inference_turbo(c3, text, temperature=0.8, top_k=1000, top_p=0.95, repetition_penalty=1.2,
                                           max_gen_len=steps)
    assert torch.equal(mid["tokens"].cpu(), toks), (mid["tokens"], toks)
    st = toks[0]
    st = torch.cat([st[st < 6561], torch.tensor([4299, 4299, 4299])])
    assert torch.equal(mid["speech_tokens"].cpu(), st)
    noised = torch.randn(1, 80, 2 * st.numel())                                  # s3gen.py:316
    mel = FlowOracle(fsd, meanflow=True).inference(st, cg, 2, noised_mels=noised)
    rms = ((mid["mel"].cpu() - mel) ** 2).mean().sqrt().item()
    assert rms < 1e-3, f"mel RMS {rms}"
    ho = HiFTOracle(hsd)
    ref_wav2, _ = ho.inference(mid["mel"].cpu(), s=mid["source"].cpu())
    err2 = (wav - ref_wav2).abs().max().item()
    assert wav.shape == ref_wav2.shape and err2 < 2e-4, f"vocoder max|dwav|={err2}"
    # batched front-end: one T3 row per utterance, 2-step meanflow, HiFT; 960 samples per speech token (+3 silence tokens)
    tm = {}
    wavs = tts.generate_batch([text[0], text[0][:10]], max_gen_len=[6, 9], top_k=1, timings=tm)
    assert len(wavs) == 2 and all(torch.isfinite(w).all() for w in wavs)
    assert sum(int(w.numel()) for w in wavs) == int(round(tm["audio_s"] * 25)) * 960
    assert all(int(w.numel()) >= 4 * 960 for w in wavs)
