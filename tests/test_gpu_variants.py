"""GPU parity of the variants on the path: meanflow 2-step CFM (Turbo's decoder) and the
multilingual T3 (text vocabulary 2454), against the reference's own outputs in tests/golden/variants_golden.pt."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_meanflow_two_step_mel(golden_dir):
    from oracle import weights as W
    from chatterbox_b200 import Engine, S3Gen
    g = torch.load(os.path.join(golden_dir, "variants_golden.pt"))["meanflow"]
    eng = Engine(0)                       # own engine: the estimator weights differ (time_embed_mixer)
    s3 = S3Gen(eng, W.make_flow_weights(0, meanflow=True), None, meanflow=True)
    assert eng.meanflow
    _, cg = W.make_conds(seed=1234, n_gen_prompt=g["n_prompt"])
    mel = s3.flow_inference(g["tokens"][0], ref_dict=cg, z=g["z"][0]).cpu()
    rms = ((mel - g["mel"]) ** 2).mean().sqrt().item()
    assert mel.shape == g["mel"].shape and rms < 1e-3, f"meanflow mel RMS {rms}"


def test_multilingual_t3_greedy_ids(golden_dir):
    from oracle import weights as W
    from chatterbox_b200 import Engine, T3, T3Cond
    g = torch.load(os.path.join(golden_dir, "variants_golden.pt"))["mtl"]
    eng = Engine(0)
    t3 = T3(eng, W.make_t3_weights(g["weights_seed"], text_vocab=2454))
    c3, _ = W.make_conds()
    toks = t3.inference(t3_cond=T3Cond(**c3), text_tokens=g["text_tokens"], max_new_tokens=10, temperature=0.8,
                        top_p=1.0, min_p=1.0, repetition_penalty=2.0, cfg_weight=0.5, kv_dtype="fp32").cpu()
    assert torch.equal(toks, g["tokens"]), (toks, g["tokens"])
