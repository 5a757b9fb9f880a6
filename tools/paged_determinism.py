"""Bitwise run-to-run determinism of the paged decode attention at a decode-like shape (many rows, short ragged contexts)."""
import os, sys, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_kernels import _paged_case, _eng
from gpu_util import _ptr
eng = _eng()
g = torch.Generator().manual_seed(5)
R = int(os.environ.get("PR", 128))
S_list = torch.randint(20, 200, (R,), generator=g).tolist()
for kv_dtype in os.environ.get("PD_DT", "fp32,bf16").split(","):
    for nsplit in (1, 4):
        for fuse in (0, 1):
            pt, pool, qkv, cos, sin = _paged_case(S_list, kv_dtype, 7, fuse)
            d = lambda t: t.cuda().contiguous()
            qkv_d, pt_d, cos_d, sin_d = d(qkv), d(pt), d(cos), d(sin)
            slot_row = torch.arange(R, dtype=torch.int32).cuda()
            pos_d = torch.tensor([s - 1 for s in S_list], dtype=torch.int32).cuda()
            ws = torch.empty(R * 16 * nsplit * 66 * 4 + 4096, dtype=torch.uint8, device="cuda")
            ref, nbad, worst = None, 0, 0.0
            for rep in range(int(os.environ.get("PD_REPS", 40))):
                pool_d = d(pool)
                out = torch.zeros(R, 1024, device="cuda")
                eng.h.call("cbx_test_paged_decode", _ptr(qkv_d), _ptr(pool_d), {"bf16": 0, "fp32": 1, "fp8": 2}[kv_dtype], pool.shape[0],
                           _ptr(pt_d), pt.shape[1], _ptr(slot_row), _ptr(pos_d), R, nsplit, 0, fuse, _ptr(cos_d), _ptr(sin_d), _ptr(out),
                           _ptr(ws), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                o = out.cpu()
                if ref is None: ref = o; continue
                if not torch.equal(o, ref):
                    nbad += 1
                    worst = max(worst, (o - ref).abs().max().item())
            print(f"{kv_dtype} rows={R} nsplit={nsplit} fuse={fuse}: {nbad} of the repetitions differ bitwise from the first (max |d| {worst:.3e})", flush=True)
