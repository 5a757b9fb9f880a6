"""The vocoder's waveform bar and the f0 -> phase integration.

HiFT's harmonic source integrates f0 over EVERY sample: theta_h(n) = 2*pi*frac(cumsum_m<=n f0(m)*h/24000)
(hifigan.py:204-211).  Two f0 tracks that differ by df (last fp32 bits of a 5-conv stack: another BLAS, another thread
count, another GPU) therefore drift apart by 2*pi*h*sum(df)/24000 rad -- in the reference's own arithmetic as much as in
this engine.  These tests turn that statement into assertions:

  * CPU: in the reference arithmetic (oracle), |s(f0 + d) - s(f0)| <= the per-sample bound computed from d, and the
    reference differs from ITSELF between 1 and N threads by a measurable f0 / waveform amount;
  * GPU: the engine's source differs from the reference's by no more than the bound computed from the engine's own
    |f0 - f0_ref|, and decoding the engine's source with the oracle reproduces the engine's waveform to 1e-4: the phase
    drift of the source explains the whole end-to-end deviation.
"""
import os

import numpy as np
import pytest
import torch


def _ulp32(x):
    x = x.abs().clamp_min(1e-30).double()
    return torch.pow(2.0, torch.floor(torch.log2(x)) - 23)


def source_drift_bound(f0_a, f0_b, lin_w):
    """Per-sample bound on |s_a - s_b| for two f0 tracks [T] with the same phases / noise and the same voiced flags.
    C_h(n) = cumsum(f0*h/24000) is accumulated in fp64 and rounded to fp32 once per output (torch CPU cumsum), `% 1` is
    exact, sin and tanh are 1-Lipschitz, the merge is a 9-tap linear layer:
        |ds(n)| <= sum_h |w_h| * 0.1 * 2*pi * (|C_h^a(n) - C_h^b(n)| + ulp32(C_h(n))) + 3e-6 (fp32 sin / tanh evaluation)."""
    up = lambda f: torch.repeat_interleave(f.double(), 480)
    bound = torch.zeros(f0_a.numel() * 480, dtype=torch.float64)
    for h in range(1, 10):
        ca = torch.cumsum((up(f0_a).float() * h / 24000).double(), 0)
        cb = torch.cumsum((up(f0_b).float() * h / 24000).double(), 0)
        bound += abs(float(lin_w[h - 1])) * 0.1 * 2 * np.pi * ((ca - cb).abs() + _ulp32(torch.maximum(ca, cb)))
    return bound + 3e-6


def _case(golden_dir):
    from oracle import weights as W
    from oracle.hift_ref import HiFTOracle
    g = torch.load(os.path.join(golden_dir, "s3gen_golden.pt"))
    hsd = W.make_hift_weights(g["weights_seed"])
    return g["cases"][0], hsd, HiFTOracle(hsd)


def _draws(case):
    from torch.distributions.uniform import Uniform
    T = case["mel"].shape[-1]
    torch.manual_seed(case["rng_seed"] + 100)
    phase = Uniform(low=-np.pi, high=np.pi).sample(sample_shape=(1, 9, 1))
    phase[:, 0, :] = 0
    noise = torch.randn(1, 9, 480 * T)
    return phase, noise


def test_phase_drift_bound_in_the_reference_arithmetic(golden_dir):
    case, hsd, ho = _case(golden_dir)
    phase, noise = _draws(case)
    with torch.inference_mode():
        f0 = ho.f0_predictor(case["mel"])
        s_ref = ho.source(f0, phase, noise)
        assert (s_ref - case["source"]).abs().max() < 1e-6            # the oracle IS the reference here (pinned)
        gen = torch.Generator().manual_seed(0)
        for rel in (1e-7, 1e-6, 1e-5):                               # 1 ulp .. 100 ulp of f0
            d = f0 * rel * (torch.rand(f0.shape, generator=gen) * 2 - 1)
            f0b = torch.where(f0 > 10, (f0 + d).clamp_min(10.001), f0)   # voiced flags unchanged
            s_b = ho.source(f0b, phase, noise)
            err = (s_b - s_ref)[0, 0].abs().double()
            bound = source_drift_bound(f0[0], f0b[0], ho.sd["m_source.l_linear.weight"][0])
            assert (err <= bound).all(), f"rel {rel}: max excess {(err - bound).max():.3e}"
            w_b = ho.decode(case["mel"], s_b)
            print(f"[drift] reference arithmetic, f0 perturbed by {rel:.0e} relative: max|ds| = {err.max():.3e} "
                  f"(bound max {bound.max():.3e}), max|dwav| = {(w_b - case['wav']).abs().max():.3e}")


def test_reference_differs_from_itself_across_thread_counts(golden_dir):
    """The reference's own f0 (5 convs + Linear, fp32) depends on the thread count of the host BLAS / conv kernels; the
    waveform it produces inherits that through the phase integral.  Printed for scale next to the engine's numbers."""
    case, hsd, ho = _case(golden_dir)
    phase, noise = _draws(case)
    n0 = torch.get_num_threads()
    try:
        with torch.inference_mode():
            torch.set_num_threads(1)
            f1 = ho.f0_predictor(case["mel"]); w1, _ = ho.inference(case["mel"], None, phase, noise, trim_fade=False)
            torch.set_num_threads(max(2, n0))
            fn = ho.f0_predictor(case["mel"]); wn, _ = ho.inference(case["mel"], None, phase, noise, trim_fade=False)
    finally:
        torch.set_num_threads(n0)
    df, dw = (f1 - fn).abs().max().item(), (w1 - wn).abs().max().item()
    print(f"[drift] reference vs itself (1 vs {max(2, n0)} threads): max|df0| = {df:.3e} Hz, max|dwav| = {dw:.3e}")
    assert df < 1e-2 and dw <= 2.0          # informational: both may be exactly 0 on a deterministic backend


@pytest.mark.gpu
def test_engine_source_within_the_drift_bound_and_decode_explained(golden_dir):
    from gpu_util import engine
    from oracle import weights as W
    from chatterbox_b200.s3gen import S3Gen
    case, hsd, ho = _case(golden_dir)
    phase, noise = _draws(case)
    s3 = S3Gen(engine(), W.make_flow_weights(0), hsd)
    with torch.inference_mode():
        f0_ref = ho.f0_predictor(case["mel"])[0]
        f0_eng = s3.engine.hift_f0([case["mel"][0]])[0].cpu()
        wav, src = s3.hift_inference(case["mel"], None, phase_vec=phase.reshape(9), noise=noise[0], trim_fade=False)
        src, wav = src.cpu().reshape(-1), wav.cpu()
        err = (src - case["source"].reshape(-1)).abs().double()
        bound = source_drift_bound(f0_ref, f0_eng, ho.sd["m_source.l_linear.weight"][0]) + 2e-5     # + engine sin/tanh (2e-5 with equal f0)
        df = (f0_eng - f0_ref).abs().max().item()
        print(f"[drift] engine: max|df0| = {df:.3e} Hz, max|ds| = {err.max():.3e} (bound max {bound.max():.3e})")
        assert ((f0_eng > 10) == (f0_ref > 10)).all()
        assert (err <= bound).all(), f"source exceeds the phase-drift bound by {(err - bound).max():.3e}"
        # the whole end-to-end deviation is the source's: the reference decoder fed the ENGINE's source gives the engine's waveform
        w_or = ho.decode(case["mel"], src.reshape(1, 1, -1))
        e2 = (wav.reshape(-1) - w_or.reshape(-1)).abs().max().item()
        print(f"[drift] engine waveform vs reference decoder on the engine's source: max|dwav| = {e2:.3e}; "
              f"vs the reference's own waveform: {(wav.reshape(-1) - case['wav'].reshape(-1)).abs().max():.3e}")
        assert e2 < 1e-4, e2
