"""Repeat the long ragged paged-attention case (tests/test_gpu_kernels.py) and print where the error sits: per row, per head."""
import os, sys, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_kernels import _paged_case, _paged_reference, _eng
from gpu_util import _ptr
eng = _eng()
S_list = [1, 33, 1000, 1190, 32, 64, 65, 517]
R = len(S_list)
for nsplit in (1, 4, 16):
    for fuse in (0, 1):
        for kv_dtype in os.environ.get("PD_DT", "bf16,fp32").split(","):
            pt, pool, qkv, cos, sin = _paged_case(S_list, kv_dtype, 11 + nsplit, fuse)
            ref, _ = _paged_reference(pt, pool, qkv, cos, sin, S_list, fuse)
            d = lambda t: t.cuda().contiguous()
            qkv_d, pt_d, cos_d, sin_d = d(qkv), d(pt), d(cos), d(sin)
            slot_row = torch.arange(R, dtype=torch.int32).cuda()
            pos_d = torch.tensor([s - 1 for s in S_list], dtype=torch.int32).cuda()
            ws = torch.empty(R * 16 * nsplit * 66 * 4 + 4096, dtype=torch.uint8, device="cuda")
            worst, bad = 0.0, 0
            where = {}
            for rep in range(int(os.environ.get("PD_REPS", 20))):
                pool_d = d(pool)
                out = torch.zeros(R, 1024, device="cuda")
                ws.fill_(int(os.environ.get("PD_FILL", "255")))       # NaN pattern: an unwritten scratch entry shows up in the output
                eng.h.call("cbx_test_paged_decode", _ptr(qkv_d), _ptr(pool_d), {"bf16": 0, "fp32": 1, "fp8": 2}[kv_dtype], pool.shape[0],
                           _ptr(pt_d), pt.shape[1], _ptr(slot_row), _ptr(pos_d), R, nsplit, 0, fuse, _ptr(cos_d), _ptr(sin_d), _ptr(out),
                           _ptr(ws), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                e = torch.nan_to_num((out.cpu().double() - ref).abs(), nan=9e9).view(R, 16, 64).amax(-1)
                worst = max(worst, e.max().item())
                if e.max().item() > 2e-5:
                    bad += 1
                    for r, h in (e > 2e-5).nonzero().tolist():
                        where[(r, h)] = where.get((r, h), 0) + 1
            print(f"{kv_dtype} nsplit={nsplit} fuse={fuse}: worst={worst:.3e} bad_reps={bad} where(row,head)->count={dict(sorted(where.items())[:12])}", flush=True)
