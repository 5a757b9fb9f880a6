"""Does a row-shifted SWIZZLE_128B UMMA descriptor read the rows one expects?  (hardware probe, see csrc/probe.cu)"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chatterbox_b200 import Engine
from chatterbox_b200.engine import _ptr
eng = Engine(0)
g = torch.Generator().manual_seed(0)
A = torch.randn(160, 64, generator=g).bfloat16().cuda()
W = torch.randn(64, 64, generator=g).bfloat16().cuda()
for mode in (0, 1):
    res = []
    for shift in range(0, 32):
        Cc = torch.zeros(128, 64, device="cuda")
        eng.h.call("cbx_test_umma_rowshift", _ptr(A), _ptr(W), shift, mode, _ptr(Cc), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ref = A[shift:shift + 128].float() @ W.float().t()
        res.append(float((Cc - ref).abs().max()))
    print(f"mode {mode}: max|err| per shift:", " ".join(f"{e:.1e}" for e in res), flush=True)
