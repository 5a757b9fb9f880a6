"""Micro-benchmark of the four decode-step projections at a given row count (default 512 = B 256 with CFG): the persistent
streaming kernel against the one-tile-per-CTA kernel, tile / split variants, weights streaming from HBM (16 copies round-robin).
    python tools/decode_gemm_bench.py            # env: DM (rows), DREPS"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chatterbox_b200 import Engine
from chatterbox_b200.engine import _ptr
eng = Engine(0)
M = int(os.environ.get("DM", 512))
reps = int(os.environ.get("DREPS", 4))
ws = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
shapes = [("qkv", 3072, 1024, 0), ("o", 1024, 1024, 0), ("gate_up", 8192, 1024, 1), ("down", 1024, 4096, 0), ("head", 8256, 1024, 0)]
variants = {"qkv": [(1, 64, 1), (1, 128, 1), (1, 256, 0)], "o": [(2, 64, 0), (4, 64, 0), (8, 64, 0), (2, 128, 0), (4, 128, 0)],
            "gate_up": [(1, 128, 1), (1, 256, 0), (1, 64, 1)], "down": [(4, 64, 0), (8, 64, 0), (2, 64, 0), (4, 128, 0), (8, 128, 0)],
            "head": [(1, 0, 0), (1, 64, 0)]}
for name, N, K, swiglu in shapes:
    nw = max(2, min(48, int(300e6 / (N * K * 2))))
    floor_us = max(N * K * 2 / 6.57e12, 2.0 * M * N * K / 1457e12) * 1e6
    for splitk, bn, dual in variants[name]:
        us = C.c_float(0.0)
        eng.h.call("cbx_bench_gemm_f16", M, N, K, splitk, bn, dual, swiglu, nw, reps, C.byref(us), _ptr(ws), ws.numel(),
                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
        print(f"{name:8s} M={M} N={N} K={K} splitk={splitk} bn={bn} dual={dual}: {us.value:7.2f} us  (floor {floor_us:.2f} us, "
              f"{N*K*2/us.value/1e3:.0f} GB/s weights, {2.0*M*N*K/us.value/1e6:.0f} TFLOP/s)", flush=True)
