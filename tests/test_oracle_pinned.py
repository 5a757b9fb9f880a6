"""Pin the CPU restatement (oracle/) against outputs of the real reference modules.

tests/golden/*.pt were written by oracle/make_golden.py, which imports the unmodified reference from
/root/reference/src and runs T3.inference / flow.inference / HiFTGenerator.inference on the seeded
synthetic checkpoints.  These tests run anywhere (no /root/reference, no GPU).
"""
import os

import torch

from oracle import weights as W
from oracle.t3_ref import T3Oracle
from oracle.flow_ref import FlowOracle
from oracle.hift_ref import HiFTOracle
from oracle.t3_turbo_ref import TurboOracle



def test_t3_oracle_matches_reference_tokens_and_logits(golden_dir):
    g = torch.load(os.path.join(golden_dir, "t3_golden.pt"))
    sd = W.make_t3_weights(g["weights_seed"])
    c3, _ = W.make_conds(g["conds_seed"])
    orc = T3Oracle(sd)
    for case in g["cases"]:
        cond = orc.prepare_conditioning(c3["speaker_emb"], c3["cond_prompt_speech_tokens"], c3["emotion_adv"])
        assert torch.allclose(cond, case["cond_emb"], atol=2e-5, rtol=1e-5)
        torch.manual_seed(case["rng_seed"])
        toks, logits = orc.inference(c3, case["text_tokens"], case["steps"], temperature=0.8, top_p=1.0,
                                     min_p=case["min_p"], repetition_penalty=1.2, cfg_weight=0.5,
                                     return_logits=True)
        assert torch.equal(toks, case["tokens"]), (toks, case["tokens"])       # bit-exact ids
        err = (logits[0] - case["prefill_logits"]).abs().max().item()
        assert err < 2e-4, err


def test_turbo_oracle_matches_reference_tokens_and_logits(golden_dir):
    """Turbo T3 (GPT-2 backbone, reference t3.py:392-468): the restatement reproduces the reference's sampled and
    greedy ids bit for bit (same torch CPU RNG stream) and its prefill logits."""
    g = torch.load(os.path.join(golden_dir, "turbo_golden.pt"))
    sd = W.make_t3_turbo_weights(g["weights_seed"], text_vocab=g["text_vocab"])
    c3, _ = W.make_conds(g["conds_seed"], n_t3_prompt=g["n_prompt"])
    orc = TurboOracle(sd)
    for case in g["cases"]:
        cond = orc.prepare_conditioning(c3["speaker_emb"], c3["cond_prompt_speech_tokens"])
        assert cond.shape[1] == case["len_cond"] == 1 + g["n_prompt"]
        assert torch.allclose(cond[:, :4], case["cond_emb_head"], atol=2e-5, rtol=1e-5)
        torch.manual_seed(case["rng_seed"])
        toks, logits = orc.inference_turbo(c3, case["text_tokens"], temperature=0.8, top_k=case["top_k"],
                                           top_p=case["top_p"], repetition_penalty=case["rep"],
                                           max_gen_len=case["steps"], return_logits=True)
        assert torch.equal(toks, case["tokens"]), (toks, case["tokens"])       # bit-exact ids
        err = (logits[0] - case["prefill_logits"]).abs().max().item()
        assert err < 2e-4, err


def test_flow_and_hift_oracle_match_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "s3gen_golden.pt"))
    fsd = W.make_flow_weights(g["weights_seed"])
    hsd = W.make_hift_weights(g["weights_seed"])
    fo, ho = FlowOracle(fsd), HiFTOracle(hsd)
    for case in g["cases"]:
        _, cg = W.make_conds(seed=1234, n_gen_prompt=case["n_prompt"])
        mu, spks, cond, mask, l1 = fo.encode(case["tokens"], cg["prompt_token"], cg["prompt_feat"], cg["embedding"])
        assert (mu.transpose(1, 2) - case["mu"]).abs().max().item() < 1e-4
        with torch.inference_mode():
            v = fo.estimator(case["z"], mask, mu, torch.tensor([case["nfe_t"]]), spks, cond)
        assert (v - case["nfe_v"]).abs().max().item() < 1e-4
        mel = fo.inference(case["tokens"], cg, 10, z=case["z"])
        rms = ((mel - case["mel"]) ** 2).mean().sqrt().item()
        assert rms < 1e-4, rms                                                  # north_star bar: 1e-3
        torch.manual_seed(case["rng_seed"] + 100)
        wav, s = ho.inference(case["mel"], trim_fade=False)
        assert (s - case["source"]).abs().max().item() < 1e-6
        assert (wav - case["wav"]).abs().max().item() < 1e-5                    # north_star bar: 1e-4


def test_weight_norm_fold_is_bf16_representable_to_1ulp_fp32():
    """g = ||v|| * 2^k makes the folded weight equal to a bf16 value up to the fp32 rounding of
    torch._weight_norm's own norm (<= 2e-7 relative) - the engine's bf16 packing is then lossless."""
    hsd = W.make_hift_weights(0)
    folded = W.fold_weight_norm(hsd)
    for k, v in folded.items():
        if k.endswith(".weight") and v.dim() == 3 and "source_downs" not in k:
            assert ((v - W.bf16_round(v)).abs() <= 2e-7 * v.abs()).all(), k


def test_variants_meanflow_and_multilingual(golden_dir):
    """meanflow 2-step estimator and the multilingual T3 vocabulary, against the reference's outputs."""
    g = torch.load(os.path.join(golden_dir, "variants_golden.pt"))
    mf = g["meanflow"]
    fo = FlowOracle(W.make_flow_weights(g["weights_seed"], meanflow=True), meanflow=True)
    _, cg = W.make_conds(seed=1234, n_gen_prompt=mf["n_prompt"])
    mel = fo.inference(mf["tokens"], cg, 2, z=mf["z"])
    assert ((mel - mf["mel"]) ** 2).mean().sqrt().item() < 1e-4
    mt = g["mtl"]
    sd = W.make_t3_weights(mt["weights_seed"], text_vocab=2454)
    c3, _ = W.make_conds()
    toks = T3Oracle(sd).inference(c3, mt["text_tokens"], 10, temperature=0.8, top_p=1.0, min_p=1.0,
                                  repetition_penalty=2.0, cfg_weight=0.5)
    assert torch.equal(toks, mt["tokens"])
