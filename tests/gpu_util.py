"""Helpers shared by the GPU parity tests (call libcbx through the C ABI; the oracle is only the checker)."""
import ctypes as C

import numpy as np
import torch

from chatterbox_b200.engine import Engine, PackedLayout, _ptr
from chatterbox_b200._lib import Layout

ACT = dict(none=0, silu=1, gelu=2, mish=3, elu=4, lrelu=5, snake=6, tanh=7)
_engine = None


def engine():
    global _engine
    if _engine is None:
        _engine = Engine(0)
    return _engine


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def run_gemm(eng, A, w, bias=None, mode=0, dil=1, pad=0, stride=1, out_layout=None, in_layout=None, act="none",
             act_p=0.0, res=None, swiglu=False, M=None, impl="tc"):
    """A: [M_in, lda] fp32 cuda; w: [N, cin, taps] fp32 cpu (bf16-representable).  Returns C [M, N_out] cuda."""
    eng.h.set_option("gemm", impl)
    N, cin, taps = w.shape
    M_in = A.shape[0]
    M = M if M is not None else (out_layout.rows if out_layout is not None else M_in)
    n_out = N // 2 if swiglu else N
    Cout = torch.full((M, n_out), float("nan"), device=A.device, dtype=torch.float32)
    wc = w.contiguous().float()
    bc = bias.contiguous().float() if bias is not None else None
    null_l = C.POINTER(Layout)()
    eng.h.call("cbx_test_gemm", _ptr(A), A.shape[1], M_in, M, C.c_void_p(wc.data_ptr()),
               C.c_void_p(bc.data_ptr()) if bc is not None else C.c_void_p(0), N, cin, taps, mode, dil, pad, stride,
               C.byref(out_layout.c) if out_layout is not None else null_l,
               C.byref(in_layout.c) if in_layout is not None else null_l,
               ACT[act], float(act_p), _ptr(res) if res is not None else C.c_void_p(0),
               res.shape[1] if res is not None else 0, 1 if swiglu else 0, _ptr(Cout), n_out,
               C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    eng.h.set_option("gemm", "tc")
    return Cout


def run_attention(eng, q, k, v, lens, n_heads, scale, causal=False, bias=None, bias_rel=False, bias_center=0, impl="tc"):
    """q,k,v: [rows, n_heads*64] packed by PackedLayout(lens).  Returns O [rows, n_heads*64]."""
    eng.h.set_option("attn", impl)
    L = PackedLayout(lens, q.device)
    O = torch.zeros_like(q)
    bh, bld = 0, 0
    if bias is not None:
        bh, bld = bias.shape[1] * bias.shape[2], bias.shape[2]
    eng.h.call("cbx_test_attention", _ptr(q), _ptr(k), _ptr(v), q.shape[1], _ptr(O), O.shape[1], n_heads, C.byref(L.c),
               float(scale), 1 if causal else 0, _ptr(bias) if bias is not None else C.c_void_p(0), bh, bld,
               1 if bias_rel else 0, int(bias_center), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    eng.h.set_option("attn", "tc")
    return O, L


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))
