"""CPU study (oracle arithmetic): how many split-precision terms does the CFM attention need for the 1e-3 mel-RMS bar?
Emulates the tensor-core operand formats inside the oracle's transformer blocks and reports the mel RMS against fp32.
Modes: operands rounded to bf16 / fp16, `terms` = 1 (hi.hi), 2 ((hi+lo).hi: A operand exact, B operand rounded once),
3 (hi.hi + hi.lo + lo.hi: what attn_tc_kernel issues today with bf16)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights as W
from oracle.flow_ref import FlowOracle


def split(x, dt):
    hi = x.to(dt).float()
    lo = (x - hi).to(dt).float()
    return hi, lo


def mm_terms(a, b, dt, terms):
    """a @ b with both operands in split format dt (fp32 accumulate)."""
    if dt is None:
        return a @ b
    ah, al = split(a, dt)
    bh, bl = split(b, dt)
    if terms == 1:
        return ah @ bh
    if terms == 2:
        return ah @ bh + al @ bh
    return ah @ bh + ah @ bl + al @ bh


class StudyOracle(FlowOracle):
    def __init__(self, sd, dt, terms):
        super().__init__(sd)
        self.dt, self.terms = dt, terms

    def _tfmr(self, p, x, bias):
        sd = self.sd
        B, T, _ = x.shape
        h = F.layer_norm(x, (256,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        sp = lambda w: F.linear(h, sd[p + w]).view(B, T, 8, 64).transpose(1, 2)
        q, k, v = sp("attn1.to_q.weight"), sp("attn1.to_k.weight"), sp("attn1.to_v.weight")
        s = mm_terms(q, k.transpose(-1, -2), self.dt, self.terms) * 0.125 + bias
        pr = torch.softmax(s, dim=-1)
        o = mm_terms(pr, v, self.dt, self.terms)
        o = o.transpose(1, 2).reshape(B, T, 512)
        x = x + F.linear(o, sd[p + "attn1.to_out.0.weight"], sd[p + "attn1.to_out.0.bias"])
        h = F.layer_norm(x, (256,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], 1e-5)
        h = F.gelu(F.linear(h, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"]))
        return x + F.linear(h, sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"])


if __name__ == "__main__":
    torch.set_num_threads(min(16, len(os.sched_getaffinity(0))))
    fsd = W.make_flow_weights(0)
    n_p, n = int(os.environ.get("NP", 60)), int(os.environ.get("N", 90))
    _, cg = W.make_conds(1234, n_gen_prompt=n_p)
    tok = torch.randint(0, 6561, (n,), generator=torch.Generator().manual_seed(5))
    z = torch.randn(1, 80, 2 * (n_p + n), generator=torch.Generator().manual_seed(6))
    ref = StudyOracle(fsd, None, 0).inference(tok, cg, 10, z=z)
    print(f"frames={2 * (n_p + n)} mel std={float(ref.std()):.3f}")
    for name, dt, terms in [("bf16 x3 (today)", torch.bfloat16, 3), ("bf16 x2", torch.bfloat16, 2), ("bf16 x1", torch.bfloat16, 1),
                            ("fp16 x2", torch.float16, 2), ("fp16 x1", torch.float16, 1)]:
        mel = StudyOracle(fsd, dt, terms).inference(tok, cg, 10, z=z)
        rms = ((mel - ref) ** 2).mean().sqrt().item()
        print(f"{name:18s} mel RMS vs fp32 = {rms:.3e}   max = {(mel - ref).abs().max().item():.3e}")


def study_gemm_inputs():
    """Second question: could the CFM estimator's GEMM/conv *inputs* travel as one fp16 (or bf16) value instead of the
    bf16 hi+lo pair?  (halves the activation bytes of the HBM-bound K=256 GEMMs and the MMA count.)  Emulated by rounding
    the input of every F.linear / F.conv1d inside the estimator; weights are bf16-exact already."""
    import oracle.flow_ref as FR
    fsd = W.make_flow_weights(0)
    n_p, n = int(os.environ.get("NP", 60)), int(os.environ.get("N", 90))
    _, cg = W.make_conds(1234, n_gen_prompt=n_p)
    tok = torch.randint(0, 6561, (n,), generator=torch.Generator().manual_seed(5))
    z = torch.randn(1, 80, 2 * (n_p + n), generator=torch.Generator().manual_seed(6))
    ref = FlowOracle(fsd).inference(tok, cg, 10, z=z)
    lin, conv = F.linear, F.conv1d
    for name, dt, attn in [("gemm in fp16, attn fp32", torch.float16, (None, 0)), ("gemm in fp16, attn fp16 x1", torch.float16, (torch.float16, 1)),
                           ("gemm in bf16, attn bf16 x1", torch.bfloat16, (torch.bfloat16, 1))]:
        orc = StudyOracle(fsd, *attn)
        est = orc.estimator

        def wrapped(*a, _est=est, _dt=dt, **k):
            F.linear = lambda x, w, b=None: lin(x.to(_dt).float(), w, b)
            F.conv1d = lambda x, w, b=None, *aa, **kk: conv(x.to(_dt).float(), w, b, *aa, **kk)
            try:
                return _est(*a, **k)
            finally:
                F.linear, F.conv1d = lin, conv
        orc.estimator = wrapped
        mel = orc.inference(tok, cg, 10, z=z)
        rms = ((mel - ref) ** 2).mean().sqrt().item()
        print(f"{name:28s} mel RMS vs fp32 = {rms:.3e}   max = {(mel - ref).abs().max().item():.3e}")


if __name__ == "__main__" and os.environ.get("GEMM_STUDY", "1") == "1":
    study_gemm_inputs()


def study_hift():
    """Third question: the vocoder's conv inputs as one fp16 value?  No: max |dwav| 7e-4 against the 1e-4 bar."""
    from oracle.hift_ref import HiFTOracle
    hsd = W.make_hift_weights(0)
    ho = HiFTOracle(hsd)
    g = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "s3gen_golden.pt"))
    mel = g["cases"][0]["mel"]
    torch.manual_seed(1)
    wav_ref, s = ho.inference(mel, trim_fade=False)
    conv, convt = F.conv1d, F.conv_transpose1d
    for name, dt in [("fp16", torch.float16), ("bf16", torch.bfloat16)]:
        F.conv1d = lambda x, w, b=None, *a, _dt=dt, **k: conv(x.to(_dt).float(), w, b, *a, **k)
        F.conv_transpose1d = lambda x, w, b=None, *a, _dt=dt, **k: convt(x.to(_dt).float(), w, b, *a, **k)
        try:
            wav, _ = ho.inference(mel, s=s, trim_fade=False)
        finally:
            F.conv1d, F.conv_transpose1d = conv, convt
        print(f"HiFT conv inputs as one {name} value: max|dwav| = {(wav - wav_ref).abs().max().item():.3e} (bar 1e-4)")


if __name__ == "__main__" and os.environ.get("HIFT_STUDY", "1") == "1":
    study_hift()
