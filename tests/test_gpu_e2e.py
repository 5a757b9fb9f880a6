"""End-to-end: ChatterboxTTS.generate_tokens on the GPU vs the oracle pipeline with the same torch seed."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_generate_matches_oracle_pipeline():
    from gpu_util import engine
    from oracle import weights as W
    from oracle.t3_ref import T3Oracle
    from oracle.flow_ref import FlowOracle
    from oracle.hift_ref import HiFTOracle
    from chatterbox_b200 import ChatterboxTTS, Conditionals, T3, T3Cond, S3Gen
    t3_sd, fsd, hsd = W.make_t3_weights(0), W.make_flow_weights(0), W.make_hift_weights(0)
    c3, cg = W.make_conds(1234, n_gen_prompt=60)
    eng = engine()
    tts = ChatterboxTTS(T3(eng, t3_sd), S3Gen(eng, fsd, hsd), None, "cuda",
                        Conditionals(T3Cond(**c3), dict(cg)))
    text = torch.randint(1, 255, (1, 30), generator=torch.Generator().manual_seed(17))
    steps = 20
    torch.manual_seed(2024)
    wav, mid = tts.generate_tokens(text, max_new_tokens=steps, rng="torch_cpu", kv_dtype="fp32", return_intermediates=True)
    # oracle pipeline with the same global seed (reference order of RNG draws)
    torch.manual_seed(2024)
    tt = F.pad(F.pad(torch.cat([text, text]), (1, 0), value=255), (0, 1), value=0)
    toks = T3Oracle(t3_sd).inference(c3, tt, steps, temperature=0.8, top_p=1.0, min_p=0.05, repetition_penalty=1.2,
                                     cfg_weight=0.5)
    assert torch.equal(mid["tokens"].cpu(), toks), (mid["tokens"], toks)
    st = toks[0]
    st = st[st < 6561]
    mel = FlowOracle(fsd).inference(st, cg, 10)
    rms = ((mid["mel"].cpu() - mel) ** 2).mean().sqrt().item()
    assert rms < 1e-3, f"mel RMS {rms}"
    ho = HiFTOracle(hsd)
    ref_wav, _ = ho.inference(mel)
    # vocoder in isolation: the oracle decodes the ENGINE's mel with the ENGINE's source (hifigan.py cache_source hook)
    ref_wav2, _ = ho.inference(mid["mel"].cpu(), s=mid["source"].cpu())
    err2 = (wav - ref_wav2).abs().max().item()
    assert wav.shape == ref_wav2.shape and err2 < 2e-4, f"vocoder max|dwav|={err2}"
    # whole pipeline: the mel difference (<= 1e-3 RMS) perturbs f0, whose running sum is the source phase, so the two
    # waveforms drift apart with utterance length (reference-vs-reference on another BLAS would too); loose bound
    err = (wav - ref_wav).abs().max().item()
    assert err < 2e-1, f"pipeline max|dwav|={err}"
