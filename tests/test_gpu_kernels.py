"""GPU unit tests of the libcbx kernels against plain torch fp32 expressions (tolerances written per test)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _eng():
    from gpu_util import engine
    return engine()


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("M,K,N", [(1, 1024, 3072), (2, 4096, 1024), (5, 256, 1024), (8, 1024, 8192),
                                   (37, 1024, 1024), (128, 64, 64), (300, 320, 256), (1000, 1024, 3072),
                                   (513, 512, 80), (260, 2048, 512)])
def test_linear_gemm(impl, M, K, N):
    from gpu_util import run_gemm, bf16r, relerr
    g = torch.Generator().manual_seed(M * 7 + K + N)
    A = torch.randn(M, K, generator=g).cuda()
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g) * 0.1
    C_ = run_gemm(_eng(), A, w[:, :, None].contiguous(), b, impl=impl)
    ref = A.double() @ w.double().t().cuda() + b.double().cuda()
    err = relerr(C_, ref)
    # activations keep 16 significand bits (bf16 hi+lo), weights are exact, fp32 accumulation
    assert torch.isfinite(C_).all() and err < 2e-5, f"{impl} M{M} K{K} N{N}: rel err {err}"


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("act", ["silu", "gelu", "mish", "elu", "lrelu", "snake"])
def test_gemm_epilogues(impl, act):
    from gpu_util import run_gemm, bf16r, relerr
    g = torch.Generator().manual_seed(5)
    M, K, N = 200, 256, 128
    A = torch.randn(M, K, generator=g).cuda()
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g) * 0.1
    res = torch.randn(M, N, generator=g).cuda()
    C_ = run_gemm(_eng(), A, w[:, :, None].contiguous(), b, act=act, act_p=0.3, res=res, impl=impl)
    y = (A.double() @ w.double().t().cuda() + b.double().cuda()).float()
    f = dict(silu=F.silu, gelu=F.gelu, mish=F.mish, elu=F.elu, lrelu=lambda t: F.leaky_relu(t, 0.3),
             snake=lambda t: t + (1.0 / (0.3 + 1e-9)) * torch.sin(t * 0.3) ** 2)[act]
    ref = f(y) + res
    err = relerr(C_, ref)
    assert err < 3e-5, f"{impl} {act}: rel err {err}"


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_gemm_swiglu(impl):
    from gpu_util import run_gemm, bf16r, relerr
    g = torch.Generator().manual_seed(6)
    for M in (2, 200):
        K, H = 1024, 512
        A = torch.randn(M, K, generator=g).cuda()
        wg = bf16r(torch.randn(H, K, generator=g) / math.sqrt(K))
        wu = bf16r(torch.randn(H, K, generator=g) / math.sqrt(K))
        w = torch.stack([wg, wu], dim=1).reshape(2 * H, K)          # interleaved rows (gate_j, up_j)
        C_ = run_gemm(_eng(), A, w[:, :, None].contiguous(), None, swiglu=True, impl=impl)
        ref = F.silu(A.double() @ wg.double().t().cuda()) * (A.double() @ wu.double().t().cuda())
        err = relerr(C_, ref)
        assert err < 3e-5, f"{impl} M{M}: rel err {err}"


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("cin,cout,k,dil,pad,causal", [(320, 256, 3, 1, 2, True), (256, 256, 7, 3, 9, False),
                                                      (64, 64, 11, 5, 25, False), (512, 512, 4, 1, 0, False),
                                                      (64, 18, 7, 1, 3, False)])
def test_conv_taps_packed(impl, cin, cout, k, dil, pad, causal):
    """implicit-GEMM conv on a packed variable-length batch == F.conv1d per sequence (zero padding)."""
    from gpu_util import run_gemm, bf16r, relerr
    from chatterbox_b200.engine import PackedLayout
    g = torch.Generator().manual_seed(cin + k)
    lens = [70, 128, 301]
    L = PackedLayout(lens, torch.device("cuda"))
    x = torch.randn(L.rows, cin, generator=g).cuda()
    w = bf16r(torch.randn(cout, cin, k, generator=g) / math.sqrt(cin * k))
    b = torch.randn(cout, generator=g) * 0.1
    C_ = run_gemm(_eng(), x, w, b, mode=0, dil=dil, pad=pad, stride=1, out_layout=L, in_layout=L, impl=impl)
    for s, n in enumerate(lens):
        xs = x[L.starts[s]:L.starts[s] + n].t()[None].double()
        right = dil * (k - 1) - pad
        ref = F.conv1d(F.pad(xs, (pad, right)), w.double().cuda(), b.double().cuda(), dilation=dil)[0].t()
        got = C_[L.starts[s]:L.starts[s] + n]
        assert relerr(got, ref) < 3e-5, f"{impl} seq{s}: {relerr(got, ref)}"
        tail = C_[L.starts[s] + n:(L.starts[s + 1] if s + 1 < len(lens) else L.rows)]
        assert (tail == 0).all(), "layout padding rows must be zeroed"


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("cin,cout,k,stride,pad", [(18, 256, 30, 15, 7), (18, 128, 6, 3, 1), (18, 64, 1, 1, 0),
                                                   (80, 512, 7, 1, 3), (80, 512, 3, 1, 1)])
def test_conv_window_strided(impl, cin, cout, k, stride, pad):
    from gpu_util import run_gemm, bf16r, relerr
    from chatterbox_b200.engine import PackedLayout
    g = torch.Generator().manual_seed(cin + k + stride)
    tin = [15 * 9 + 1, 15 * 20 + 1]
    tout = [(t + 2 * pad - k) // stride + 1 for t in tin]
    Lin = PackedLayout(tin, torch.device("cuda"))
    Lout = PackedLayout(tout, torch.device("cuda"))
    x = torch.randn(Lin.rows, cin, generator=g).cuda()
    w = bf16r(torch.randn(cout, cin, k, generator=g) / math.sqrt(cin * k))
    b = torch.randn(cout, generator=g) * 0.1
    C_ = run_gemm(_eng(), x, w, b, mode=1, dil=0, pad=pad, stride=stride, out_layout=Lout, in_layout=Lin, impl=impl)
    for s in range(2):
        xs = x[Lin.starts[s]:Lin.starts[s] + tin[s]].t()[None].double()
        ref = F.conv1d(xs, w.double().cuda(), b.double().cuda(), stride=stride, padding=pad)[0].t()
        got = C_[Lout.starts[s]:Lout.starts[s] + tout[s]]
        assert relerr(got, ref) < 3e-5, f"{impl} seq{s}: {relerr(got, ref)}"


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("causal", [False, True])
def test_flash_attention_varlen(impl, causal):
    from gpu_util import run_attention, relerr
    g = torch.Generator().manual_seed(3)
    lens, H = [5, 64, 130, 257], 8
    from chatterbox_b200.engine import PackedLayout
    L0 = PackedLayout(lens, torch.device("cuda"))
    q, k, v = (torch.randn(L0.rows, H * 64, generator=g).cuda() for _ in range(3))
    O, L = run_attention(_eng(), q, k, v, lens, H, 0.125, causal=causal, impl=impl)
    for s, n in enumerate(lens):
        sl = slice(L.starts[s], L.starts[s] + n)
        sp = lambda t: t[sl].view(n, H, 64).transpose(0, 1)[None].double()
        ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v), is_causal=causal)[0].transpose(0, 1).reshape(n, H * 64)
        err = relerr(O[sl], ref)
        assert err < 3e-5, f"{impl} causal={causal} seq{s} len{n}: {err}"


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_flash_attention_relpos_bias(impl):
    """bias[h][i][center - i + j] addressing == espnet rel_shift (transformer/attention.py:225-247)."""
    from gpu_util import run_attention, relerr
    from oracle.flow_ref import rel_shift
    g = torch.Generator().manual_seed(4)
    T, H = 150, 8
    lens = [T]
    from chatterbox_b200.engine import PackedLayout
    L0 = PackedLayout(lens, torch.device("cuda"))
    q, k, v = (torch.randn(L0.rows, H * 64, generator=g).cuda() for _ in range(3))
    Tmax = 170                                    # table built for a longer sequence
    raw = torch.randn(H, L0.rows, 2 * Tmax - 1, generator=g).cuda()
    O, L = run_attention(_eng(), q, k, v, lens, H, 0.125, bias=raw, bias_rel=True, bias_center=Tmax - 1, impl=impl)
    sp = lambda t: t[:T].view(T, H, 64).transpose(0, 1)[None].double()
    raw_T = raw[:, :T, Tmax - T:Tmax + T - 1].cpu()[None]         # the (2T-1)-wide slice the reference would see
    bd = rel_shift(raw_T).cuda().double()
    scores = (sp(q) @ sp(k).transpose(-1, -2) + bd) * 0.125
    ref = (scores.softmax(-1) @ sp(v))[0].transpose(0, 1).reshape(T, H * 64)
    err = relerr(O[:T], ref)
    assert err < 3e-5, f"{impl}: {err}"


def test_linear_gemm_wide_tile_bn256():
    """enough tiles for the 128x256 tile variant (3-stage pipeline)."""
    from gpu_util import run_gemm, bf16r, relerr
    g = torch.Generator().manual_seed(12)
    M, K, N = 4096, 320, 1024
    A = torch.randn(M, K, generator=g).cuda()
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    C_ = run_gemm(_eng(), A, w[:, :, None].contiguous(), None, impl="tc")
    ref = A.double() @ w.double().t().cuda()
    err = relerr(C_, ref)
    assert err < 2e-5, f"rel err {err}"


def test_tcgen05_attention_matches_sdpa():
    """CFM-path attention on the tcgen05 tensor cores (TMA operands, S/PV in TMEM) vs fp64 SDPA, packed varlen."""
    import ctypes as C
    from gpu_util import relerr
    from chatterbox_b200.engine import PackedLayout, _ptr
    eng = _eng()
    g = torch.Generator().manual_seed(21)
    lens, H = [5, 64, 130, 300, 777], 8
    L = PackedLayout(lens, torch.device("cuda"))
    qkv = torch.randn(L.rows, 3 * H * 64, generator=g).cuda()
    O = torch.zeros(L.rows, H * 64, device="cuda")
    ws = torch.empty(L.rows * 3 * H * 64 * 4 + (1 << 20), dtype=torch.uint8, device="cuda")
    eng.h.call("cbx_test_attention_tc", _ptr(qkv), _ptr(O), H, C.byref(L.c), 0.125, _ptr(ws), ws.numel(),
               C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    for s, n in enumerate(lens):
        sl = slice(L.starts[s], L.starts[s] + n)
        sp = lambda t: t[sl].view(n, H, 64).transpose(0, 1)[None].double()
        q, k, v = qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:]
        ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))[0].transpose(0, 1).reshape(n, H * 64)
        err = relerr(O[sl], ref)
        assert err < 3e-5, f"seq{s} len{n}: {err}"


def test_linear_gemm_dual_cta_tiles():
    """enough 128x128 tiles for the two-CTAs-per-SM configuration (2-stage pipeline, GELU epilogue + residual)."""
    from gpu_util import run_gemm, bf16r, relerr
    g = torch.Generator().manual_seed(13)
    M, K, N = 8192, 256, 1024
    A = torch.randn(M, K, generator=g).cuda()
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K))
    b = torch.randn(N, generator=g) * 0.1
    res = torch.randn(M, N, generator=g).cuda()
    C_ = run_gemm(_eng(), A, w[:, :, None].contiguous(), b, act="gelu", res=res, impl="tc")
    ref = F.gelu((A.double() @ w.double().t().cuda() + b.double().cuda()).float()) + res
    err = relerr(C_, ref)
    assert err < 3e-5, f"rel err {err}"


# ---------------------------------------------------------------------------------------------------------------------
# round 2: decode-path kernels
# ---------------------------------------------------------------------------------------------------------------------
def _paged_case(S_list, kv_dtype, seed, fuse):
    """Random paged cache for rows with context lengths S_list (token S-1 is the step's own token)."""
    from chatterbox_b200.engine import llama3_rope_tables
    g = torch.Generator().manual_seed(seed)
    R = len(S_list)
    pages_per_row = [(s + 31) // 32 + 1 for s in S_list]
    n_pages = sum(pages_per_row)
    perm = torch.randperm(n_pages, generator=g)                 # pages of a row are scattered over the pool
    max_pages = max(pages_per_row)
    pt = torch.zeros(R, max_pages, dtype=torch.int32)
    o = 0
    for r in range(R):
        pt[r, :pages_per_row[r]] = perm[o:o + pages_per_row[r]].to(torch.int32)
        o += pages_per_row[r]
    dt = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp8": torch.float8_e4m3fn}[kv_dtype]
    pool = (torch.randn(n_pages, 2, 16, 32, 64, generator=g) * 0.7).to(dt)
    qkv = torch.randn(R, 3072, generator=g)
    cos, sin = llama3_rope_tables(max(S_list) + 8)
    return pt, pool, qkv, cos, sin


def _paged_reference(pt, pool, qkv, cos, sin, S_list, fuse):
    """fp64 softmax(q k^T / 8) v per row and head from the gathered cache (+ RoPE / append of the new token when fused)."""
    R = len(S_list)
    out = torch.zeros(R, 1024, dtype=torch.float64)
    new_k = torch.zeros(R, 16, 64)
    for r, S in enumerate(S_list):
        pos = S - 1
        pages = pt[r, :(S + 31) // 32].long()
        K = pool[pages, 0].float().double().permute(1, 0, 2, 3).reshape(16, -1, 64)[:, :S].clone()      # [16, S, 64]
        V = pool[pages, 1].float().double().permute(1, 0, 2, 3).reshape(16, -1, 64)[:, :S].clone()
        q = qkv[r, :1024].view(16, 64)
        if fuse:
            c = torch.cat([cos[pos], cos[pos]]); s = torch.cat([sin[pos], sin[pos]])
            rot = lambda x: x * c + torch.cat([-x[:, 32:], x[:, :32]], -1) * s
            q = rot(q)
            k = rot(qkv[r, 1024:2048].view(16, 64)).to(pool.dtype)
            v = qkv[r, 2048:].view(16, 64).to(pool.dtype)
            new_k[r] = k.float()
            K[:, pos] = k.float().double(); V[:, pos] = v.float().double()
        sc = torch.einsum("hd,hsd->hs", q.double(), K) * 0.125
        p = torch.softmax(sc, -1)
        out[r] = torch.einsum("hs,hsd->hd", p, V).reshape(-1)
    return out, new_k


@pytest.mark.parametrize("kv_dtype", ["bf16", "fp32", "fp8"])
@pytest.mark.parametrize("impl,fuse", [(0, 0), (0, 1), (1, 0)])
@pytest.mark.parametrize("nsplit", [1, 4, 16])
def test_paged_decode_attention_long_ragged(kv_dtype, impl, fuse, nsplit):
    """paged decode attention (bulk-copy staged kernel, with and without the fused RoPE + KV append, and the round-1 __ldg
    kernel) against fp64 SDPA at the bench's context lengths: ragged rows S in {1, 33, 1000, 1190, ...}, every split count
    the engine uses."""
    import ctypes as C
    from gpu_util import _ptr
    if kv_dtype == "fp8" and impl == 1:
        pytest.skip("the fp8 cache is served by the bulk-copy kernel only")
    eng = _eng()
    S_list = [1, 33, 1000, 1190, 32, 64, 65, 517]
    pt, pool, qkv, cos, sin = _paged_case(S_list, kv_dtype, 11 + nsplit, fuse)
    ref, new_k = _paged_reference(pt, pool, qkv, cos, sin, S_list, fuse)
    R = len(S_list)
    d = lambda t: t.cuda().contiguous()
    pool_d, qkv_d, pt_d, cos_d, sin_d = d(pool), d(qkv), d(pt), d(cos), d(sin)
    slot_row = torch.arange(R, dtype=torch.int32).cuda()
    pos_d = torch.tensor([s - 1 for s in S_list], dtype=torch.int32).cuda()
    out = torch.zeros(R, 1024, device="cuda")
    ws = torch.empty(R * 16 * nsplit * 66 * 4 + 4096, dtype=torch.uint8, device="cuda")
    eng.h.call("cbx_test_paged_decode", _ptr(qkv_d), _ptr(pool_d), {"bf16": 0, "fp32": 1, "fp8": 2}[kv_dtype], pool.shape[0], _ptr(pt_d),
               pt.shape[1], _ptr(slot_row), _ptr(pos_d), R, nsplit, impl, fuse, _ptr(cos_d), _ptr(sin_d), _ptr(out), _ptr(ws),
               ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 2e-5, f"{kv_dtype} impl{impl} fuse{fuse} nsplit{nsplit}: max|dO| = {err}"
    if fuse:            # the step's k / v landed in the cache at `pos`, rotated
        pool_after = pool_d.cpu()
        for r, S in enumerate(S_list):
            pos = S - 1
            page = int(pt[r, pos // 32])
            kk = pool_after[page, 0, :, pos % 32].float()
            vv = pool_after[page, 1, :, pos % 32].float()
            assert (kk - new_k[r]).abs().max().item() < (1e-6 if kv_dtype == "fp32" else 1e-2 if kv_dtype == "bf16" else 0.26)
            assert torch.equal(vv, qkv[r, 2048:].view(16, 64).to(pool.dtype).float())


@pytest.mark.parametrize("M,K,N,splitk,bn", [(512, 1024, 1024, 2, 64), (512, 4096, 1024, 4, 64), (300, 4096, 1024, 8, 64),
                                              (128, 1024, 1024, 4, 64), (512, 1024, 1024, 1, 128), (40, 4096, 768, 8, 64)])
def test_gemm_splitk_partials_reduce_deterministically(M, K, N, splitk, bn):
    """decode-path projection: bf16 hi/lo planes by TMA, split-K partial sums, fixed-order reduction (resid_norm)."""
    import ctypes as C
    from gpu_util import _ptr, bf16r, relerr
    eng = _eng()
    g = torch.Generator().manual_seed(M + K + splitk)
    A = torch.randn(M, K, generator=g).cuda()
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K)).contiguous()
    outs = []
    for _ in range(2):
        Cc = torch.full((M, N), float("nan"), device="cuda")
        ws = torch.empty(M * K * 4 + splitk * M * N * 4 + 8192, dtype=torch.uint8, device="cuda")
        eng.h.call("cbx_test_gemm_splitk", _ptr(A), C.c_void_p(w.data_ptr()), M, N, K, splitk, bn, _ptr(Cc), _ptr(ws), ws.numel(),
                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        outs.append(Cc)
    ref = A.double() @ w.double().t().cuda()
    err = relerr(outs[0], ref)
    assert err < 2e-5, f"rel err {err}"
    assert torch.equal(outs[0], outs[1])           # bit-identical run to run


@pytest.mark.parametrize("M,K,N,act,out_half,use_res", [
    (4096, 256, 1536, 0, 1, 0),      # qkv projection  -> fp16 plane          (weight-resident kernel, BN 256)
    (3000, 256, 1024, 2, 1, 0),      # ff1 + GELU      -> fp16 plane, ragged last row tile
    (4096, 512, 256, 0, 0, 1),       # out projection + residual -> fp32      (weight-resident kernel, BN 128)
    (20000, 256, 1536, 0, 1, 0),     # many row tiles per CTA (accumulator ring wraps several times)
    (640, 512, 256, 0, 0, 1),        # fewer row tiles than CTA groups
    (4096, 1024, 256, 0, 0, 1),      # ff2 (K = 1024): the one-tile-per-CTA kernel with the same operand format
    (512, 1024, 3072, 0, 0, 0),      # decode qkv at 512 rows: persistent streaming kernel (M <= 1024, K >= 512)
    (300, 4096, 1024, 0, 0, 1),      # decode down projection, ragged rows, residual
    (64, 1024, 1024, 2, 1, 0),       # half a row tile, GELU, fp16 plane out (streaming kernel)
])
def test_fp16_plane_gemm_weight_resident(M, K, N, act, out_half, use_res):
    """fp16-plane operand format of the CFM block projections (A fp16 x fp16 copy of W, fp32 accumulate) against fp64 on
    the same fp16-rounded operands."""
    import ctypes as C
    from gpu_util import _ptr, bf16r, relerr
    eng = _eng()
    g = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=g).cuda()
    w = bf16r(torch.randn(N, K, generator=g) / math.sqrt(K)).contiguous()
    b = (torch.randn(N, generator=g) * 0.1).contiguous()
    res = torch.randn(M, N, generator=g).cuda() if use_res else None
    Cc = torch.full((M, N), float("nan"), device="cuda")
    ws = torch.empty(M * K * 2 + M * N * 2 + 8192, dtype=torch.uint8, device="cuda")
    eng.h.call("cbx_test_gemm_f16", _ptr(A), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), _ptr(res) if use_res else C.c_void_p(0),
               M, N, K, act, out_half, _ptr(Cc), _ptr(ws), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = A.half().double() @ w.half().double().t().cuda() + b.double().cuda()
    if act == 2:
        ref = F.gelu(ref)
    if use_res:
        ref = ref + res.double()
    if out_half:
        ref = ref.half().double()
    err = relerr(Cc, ref)
    tol = 1e-3 if out_half else (2e-6 if K <= 1024 else 6e-6)      # fp32 accumulation error grows with the reduction length
    assert torch.isfinite(Cc).all() and err < tol, f"rel err {err}"
