// GEMM family of libcbx:  C[M,N] = epilogue( A_gathered[M,K] (fp32) x W[N,K]^T (bf16) )
//
//  * gemm_tc_kernel<BN>   tcgen05 tensor-core tiles (M=128 x N=BN x K=64 per stage), accumulator in TMEM,
//                         W tiles by TMA (SWIZZLE_128B), A tiles gathered from fp32 activations by 8 producer
//                         warps (implicit im2col), split on the fly into bf16 hi+lo so that the product keeps
//                         ~16 significand bits of the activations (2 MMAs per K step share one W tile).
//  * gemv_kernel<R,NB>    weight-streaming GEMV for the AR decode step with few rows (HBM-bound).
//  * gemm_simt_kernel     plain fp32 tiles -- debug reference for the two kernels above (CBX_GEMM=simt).
//
// Algorithmic HBM bytes per launch (roofline): N*K*2 (weights) + M*K_in*4 (activations) + M*N*4 (output).
#include "ops.h"
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace cbx {

// ================================================================================================
// shared device helpers: row geometry + epilogue
// ================================================================================================
struct RowGeom {
  long in_row0;  // input row feeding tap 0 (may be outside [lo,hi))
  int lo, hi;    // valid input rows
  int valid;     // output row is a real row (not layout padding / beyond M)
};

__device__ __forceinline__ RowGeom row_geom(const GemmDev& g, int row) {
  RowGeom r;
  if (row >= g.M) { r.in_row0 = 0; r.lo = 0; r.hi = 0; r.valid = 0; return r; }
  if (g.has_seq) {
    int tile = row / kTileM;
    int s = g.seq.tile_seq[tile];
    if (s < 0) { r.in_row0 = 0; r.lo = 0; r.hi = 0; r.valid = 0; return r; }
    int t = row - g.seq.out_start[s];
    int ist = g.seq.in_start[s];
    r.in_row0 = (long)ist + (long)t * g.stride - g.pad;
    r.lo = ist;
    r.hi = ist + g.seq.in_len[s];
    r.valid = (t < g.seq.out_len[s]);
    if (!r.valid) { r.lo = 0; r.hi = 0; }
  } else {
    r.in_row0 = (long)row * g.stride - g.pad;
    r.lo = 0; r.hi = g.M_in; r.valid = 1;
  }
  return r;
}

// one A element (used by the SIMT reference and by the scalar tails of the producers)
static __device__ __noinline__ float load_a_elem(const GemmDev& g, const RowGeom& rg, int k) {
  if (!rg.valid) return 0.0f;
  if (g.a_mode == A_TAPS) {
    int tap = k / g.ctap, c = k - tap * g.ctap;
    long ir = rg.in_row0 + (long)tap * g.dil;
    if (tap >= g.ntaps || c >= g.c_in || ir < rg.lo || ir >= rg.hi) return 0.0f;
    return g.A[ir * g.lda + c];
  } else {
    if (k >= g.k_total) return 0.0f;
    long flat = rg.in_row0 * g.c_in + k;
    if (flat < (long)rg.lo * g.c_in || flat >= (long)rg.hi * g.c_in) return 0.0f;
    return g.A[flat];
  }
}

// epilogue for one output element (non-swiglu)
static __device__ __noinline__ void epilogue_store(const GemmDev& g, int row, int n, float acc, bool row_valid) {
  if (n >= g.n_out) return;
  float v = acc * g.alpha + (g.bias ? g.bias[n] : 0.0f);
  v = act_apply(g.act, v, g.act_vec ? g.act_vec[n] : g.act_p);
  if (g.res) v += g.res[(long)row * g.ldr + n];
  v *= g.out_scale;
  if (g.Chi) {                       // bf16 hi/lo planes (operands of the tcgen05 attention)
    if (!row_valid) v = 0.0f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    g.Chi[(long)row * g.ldcb + n] = h;
    g.Clo[(long)row * g.ldcb + n] = l;
    if (!g.C) return;
  }
  float* cp = g.C + (long)row * g.ldc + n;
  if (g.accumulate) v += *cp;
  if (!row_valid) v = 0.0f;
  *cp = v;
  if (g.C2) g.C2[(long)row * g.ldc2 + n] = row_valid ? act_apply(g.act2, v, g.act2_vec ? g.act2_vec[n] : g.act2_p) : 0.0f;
}
static __device__ __noinline__ void epilogue_store_swiglu(const GemmDev& g, int row, int n_even, float a0, float a1,
                                                      bool row_valid) {
  if (n_even >= g.n_out) return;
  float v0 = a0 * g.alpha + (g.bias ? g.bias[n_even] : 0.0f);
  float v1 = a1 * g.alpha + (g.bias ? g.bias[n_even + 1] : 0.0f);
  float v = (v0 / (1.0f + expf(-v0))) * v1;
  if (!row_valid) v = 0.0f;
  g.C[(long)row * g.ldc + (n_even >> 1)] = v;
}

// ================================================================================================
// tcgen05 kernel
//   warp 8  : TMA producer  -- W tile (bf16, SWIZZLE_128B) and, when the A operand is a plain strided fp32 matrix
//             (Linear / conv taps), the raw fp32 A tile [128 rows][64 floats] of the current tap
//   warps 0-7: converters   -- fp32 tile -> bf16 hi/lo planes IN PLACE in the swizzled UMMA layout (zeroing rows whose
//             tap falls outside their sequence); or, for window/strided convs, a register gather straight from global
//   warp 9  : MMA issuer    -- one thread, tcgen05.mma M128 x N{64,128,256} x K16, 2 MMAs (hi, lo) per K step, TMEM
//   warps 0-7 again: epilogue -- tcgen05.ld -> smem transpose -> coalesced bias/activation/residual/stores
// ================================================================================================
constexpr int TC_BM = 128, TC_BK = 64;
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 2) * 32;   // + TMA warp + MMA warp

// DUAL = 1: shallow 2-stage pipeline but two CTAs per SM, so one CTA's epilogue (LSU / ALU) runs under the other's
// main loop (TMA / tensor pipe) without a persistent-kernel restructure.
template <int BN, int DUAL = 0> struct TcCfg {
  static constexpr int STAGES = DUAL ? 2 : ((BN == 256) ? 3 : 4);
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;        // fp32 tile == two bf16 planes of 16 KB (converted in place)
  static constexpr int W_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM <= 232448, "shared memory budget");
};

// fp32 pair -> bf16x2 hi plane word and lo plane word (one F2FP each; bf16 -> fp32 is a shift)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float fa = __uint_as_float(hi << 16), fb = __uint_as_float(hi & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - fa, b - fb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}


// ---- register-direct epilogue ---------------------------------------------------------------------------------------
// One thread = one output row; a tcgen05.ld brings 32 consecutive accumulator columns of that row into registers and the
// whole epilogue (bias, activation, residual, SwiGLU, plane split) runs on them with 32 independent chains, then leaves
// as 16 / 32-byte vector stores (whole sectors per thread: a 32-column fp32 segment is one 128-byte line).  No shared-memory
// transpose, no rolled per-row loops: the transposing epilogue below needed ~5 k dependent instructions per warp and, on
// decode-sized GEMMs (<= 4 row tiles, one wave), WAS the kernel (ncu r2_ncu_gemm_tc_decode_before.txt: 14 cycles per issued
// instruction, 70 k of 77 k cycles).
__device__ __forceinline__ void stg256(float* p, const float* v) {      // 32-byte aligned
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void ldg256(const float* p, float* v) {      // 32-byte aligned
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]),
               "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2x(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// NV values of one row -> fp32 / fp16 plane / bf16 hi+lo planes at element offset `off` of the row (vector stores when the
// whole group is in range and aligned, guarded scalar stores for the ragged last group of a matrix)
template <int NV>
__device__ __forceinline__ void store_row_f32(float* dst, const float* v, int nvalid, bool wide) {
  if (nvalid >= NV) {
    if (wide) {
#pragma unroll
      for (int j = 0; j < NV; j += 8) stg256(dst + j, v + j);
    } else {
#pragma unroll
      for (int j = 0; j < NV; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) if (j < nvalid) dst[j] = v[j];
  }
}
template <int NV>
__device__ __forceinline__ void store_row_f16(__half* dst, const float* v, int nvalid) {
  if (nvalid >= NV) {
#pragma unroll
    for (int j = 0; j < NV; j += 8)
      *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_h2x(v[j], v[j + 1]), pack_h2x(v[j + 2], v[j + 3]), pack_h2x(v[j + 4], v[j + 5]), pack_h2x(v[j + 6], v[j + 7]));
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) if (j < nvalid) dst[j] = __float2half_rn(v[j]);
  }
}
template <int NV>
__device__ __forceinline__ void store_row_hilo(__nv_bfloat16* dh, __nv_bfloat16* dl, const float* v, int nvalid) {
  if (nvalid >= NV) {
#pragma unroll
    for (int j = 0; j < NV; j += 8) {
      uint4 h, l;
      split_pair(v[j], v[j + 1], h.x, l.x); split_pair(v[j + 2], v[j + 3], h.y, l.y);
      split_pair(v[j + 4], v[j + 5], h.z, l.z); split_pair(v[j + 6], v[j + 7], h.w, l.w);
      *reinterpret_cast<uint4*>(dh + j) = h; *reinterpret_cast<uint4*>(dl + j) = l;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) if (j < nvalid) { __nv_bfloat16 hh, ll; split_bf16(v[j], hh, ll); dh[j] = hh; dl[j] = ll; }
  }
}

// epilogue of one warp's quarter (32 rows) x column half of the tile; Cz = C + split offset
// stg / tmC (optional, g.epi_direct bit 3): fp32 rows leave through a per-warp 32-row x 128-byte staging tile (SWIZZLE_128B,
// 1024-byte aligned) and ONE bulk tensor store per 32-column group -- whole 128-byte lines per request; the tensor map
// clips rows >= M and columns >= n_out.
template <int BN>
__device__ __forceinline__ void epilogue_direct(const GemmDev& g, float* Cz, uint32_t tmem_base, int m0, int n0, int q, int half,
                                                int lane, bool row_valid, uint8_t* stg = nullptr, const CUtensorMap* tmC = nullptr) {
  const int row = m0 + q * 32 + lane;
  const bool tma_out = stg != nullptr && (g.epi_direct & 8) != 0;
  const bool row_ok = row < g.M;
  const bool wide_c = (g.epi_direct & 2) != 0, wide_r = (g.epi_direct & 4) != 0;
#pragma unroll 1
  for (int cc = 0; cc < BN / 2; cc += 32) {
    const int col = half * (BN / 2) + cc;
    const int n = n0 + col;
    if (n >= g.n_out) break;                     // warp-uniform
    const int nvalid = g.n_out - n;              // >= 32 except in the ragged last chunk
    float rres[32];
    const bool has_res = g.res != nullptr;
    if (has_res && row_ok && nvalid >= 32) {
      const float* rp = g.res + (long)row * g.ldr + n;
      if (wide_r) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) ldg256(rp + j, rres + j);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4) { const float4 t = *reinterpret_cast<const float4*>(rp + j); rres[j] = t.x; rres[j + 1] = t.y; rres[j + 2] = t.z; rres[j + 3] = t.w; }
      }
    } else if (has_res && row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j) rres[j] = j < nvalid ? g.res[(long)row * g.ldr + n + j] : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) rres[j] = 0.f;
    }
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
    tmem_ld_wait();
    float v[32];
    if (g.bias) {                                // Npad-padded: the whole 32-column group is readable
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n + j));
        v[j] = fmaf(__uint_as_float(r[j]), g.alpha, b.x); v[j + 1] = fmaf(__uint_as_float(r[j + 1]), g.alpha, b.y);
        v[j + 2] = fmaf(__uint_as_float(r[j + 2]), g.alpha, b.z); v[j + 3] = fmaf(__uint_as_float(r[j + 3]), g.alpha, b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * g.alpha;
    }
    if (g.swiglu) {
      // columns (2j, 2j+1) = (gate_j, up_j) -> out[j] = silu(gate) * up: 16 outputs at column n/2
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { const float v0 = v[2 * j]; o[j] = row_valid ? (v0 / (1.0f + expf(-v0))) * v[2 * j + 1] : 0.f; }
      if (row_ok) {
        const int no = n >> 1, nv = nvalid >> 1;
        if (g.Chi && g.c_half) store_row_f16<16>(reinterpret_cast<__half*>(g.Chi) + (long)row * g.ldcb + no, o, nv);
        else if (g.Chi) store_row_hilo<16>(g.Chi + (long)row * g.ldcb + no, g.Clo + (long)row * g.ldcb + no, o, nv);
        else store_row_f32<16>(g.C + (long)row * g.ldc + no, o, nv, wide_c);
      }
      continue;
    }
    switch (g.act) {       // warp-uniform; one 32-wide copy of each hot activation
      case ACT_NONE: break;
      case ACT_LRELU:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * g.act_p;
        break;
      case ACT_SILU:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] / (1.0f + expf(-v[j]));
        break;
      case ACT_GELU:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752440f));
        break;
      default:             // ACT_GELU_TANH (transformers NewGELUActivation)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + tanhf(0.79788456080286535588f * (v[j] + 0.044715f * (v[j] * v[j] * v[j]))));
        break;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) { v[j] = (v[j] + rres[j]) * g.out_scale; if (!row_valid) v[j] = 0.f; }
    if (tma_out) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous group has left the staging tile
      __syncwarp();
      const uint32_t srow = smem_u32(stg) + (uint32_t)lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((uint32_t)(j ^ (lane & 7)) << 4)), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                     "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(tmC), "r"(smem_u32(stg)), "r"(n), "r"(m0 + q * 32) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else if (row_ok) {
      if (g.Chi && g.c_half) store_row_f16<32>(reinterpret_cast<__half*>(g.Chi) + (long)row * g.ldcb + n, v, nvalid);
      else if (g.Chi) store_row_hilo<32>(g.Chi + (long)row * g.ldcb + n, g.Clo + (long)row * g.ldcb + n, v, nvalid);
      else store_row_f32<32>(Cz + (long)row * g.ldc + n, v, nvalid, wide_c);
    }
  }
  if (tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory may be reused / released
}

template <int BN, int DUAL>
__global__ void __launch_bounds__(TC_THREADS, DUAL ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmapW, const __grid_constant__ CUtensorMap tmapA,
               const __grid_constant__ CUtensorMap tmapA2, const __grid_constant__ CUtensorMap tmapC, const GemmDev g) {
  using Cfg = TcCfg<BN, DUAL>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // keep the __shared__ address space (LDS/STS instead of generic LD/ST): offset the shared pointer, do not round-trip through an integer
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* tma_full = bars;                  // [STAGES] TMA bytes landed (W, and the fp32 A tile in TMA mode)
  uint64_t* conv_full = bars + STAGES;        // [STAGES] 8 converter-warp arrivals: bf16 planes ready
  uint64_t* empty_bar = bars + 2 * STAGES;    // [STAGES] 1 arrival (tcgen05.commit): stage consumed
  uint64_t* accum_bar = bars + 3 * STAGES;    // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;             // N tiles fastest: CTAs sharing an A tile run together (L2 reuse)
  const int m0 = blockIdx.y * TC_BM;
  pdl_wait();                                 // (decode step) the producer of A / m_live has completed
  pdl_launch_dependents();
  if (g.m_live && m0 >= *g.m_live) return;    // device-side retirement: the row tile holds no live decode row
  // split-K: blockIdx.z owns K blocks [kb0, kb0 + KB) and writes its partial sums to C + z * split_stride (plain
  // stores in a fixed order: the reduction happens in the consumer kernel, deterministically)
  const int KBt = g.Kpad / TC_BK;
  const int kb0 = (g.splitk > 1) ? (int)((long)blockIdx.z * KBt / g.splitk) : 0;
  const int KB = (g.splitk > 1) ? (int)((long)(blockIdx.z + 1) * KBt / g.splitk) - kb0 : KBt;
  float* const Cz = g.C ? g.C + (long)blockIdx.z * g.split_stride : nullptr;
  const bool dbg = g.dbg && blockIdx.y == gridDim.y / 2 && blockIdx.x == 0;
  if (dbg && threadIdx.x == 0) g.dbg[5] = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&tma_full[s], 1); mbar_init(&conv_full[s], TC_PRODUCER_WARPS); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == TC_PRODUCER_WARPS + 1) tmem_alloc<BN>(tmem_slot);   // MMA warp owns TMEM
  if (warp == TC_PRODUCER_WARPS && lane == 0) {
    tma_prefetch_desc(&tmapW);
    if (g.a_tma) tma_prefetch_desc(&tmapA);
    if (g.a_tma == 2) tma_prefetch_desc(&tmapA2);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (dbg && threadIdx.x == 0) g.dbg[0] = clock64();

  if (warp < TC_PRODUCER_WARPS) {
    const int t = threadIdx.x;            // 0..255
    const int c4 = t & 15;                // float4 chunk inside the 64-wide K block
    const int rsub = t >> 4;              // 0..15
    if (g.a_tma >= 2) {
      // A arrives as bf16 planes (2) or one fp16 plane (3) by TMA: nothing to convert, go wait for the accumulator
    } else if (g.a_tma) {
      // ===================== converters: smem fp32 tile -> bf16 hi/lo planes, in place ===================
      // software pipelined: the LDS of K block kb+1 are in flight while block kb is converted and stored.
      // (One group of 8 warps on every block: two groups on alternating blocks would skip mbarrier phases, and a
      // parity wait cannot tell phase k from phase k+2.)
      const bool need_mask = g.has_seq || g.ntaps > 1;
      int lo_rel[8], hi_rel[8];       // per row: tap offsets (in input rows) that stay inside the row's own sequence
      if (need_mask) {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const RowGeom rr = row_geom(g, m0 + p * 16 + rsub);
          lo_rel[p] = rr.valid ? (int)(rr.lo - rr.in_row0) : 1;
          hi_rel[p] = rr.valid ? (int)(rr.hi - rr.in_row0) : 0;
        }
      }
      auto fetch = [&](int kb, float4 (&v)[8]) {
        const int s = kb % STAGES;
        mbar_wait(&tma_full[s], (kb / STAGES) & 1);
        const uint8_t* a_st = smem + s * Cfg::STAGE_BYTES;
#pragma unroll
        for (int p = 0; p < 8; ++p) v[p] = *reinterpret_cast<const float4*>(a_st + (p * 16 + rsub) * 256 + c4 * 16);
        if (need_mask) {
          const int toff = (((kb0 + kb) * TC_BK) / g.ctap) * g.dil;
#pragma unroll
          for (int p = 0; p < 8; ++p)
            if (toff < lo_rel[p] || toff >= hi_rel[p]) v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float4 v[8], vn[8];
      fetch(0, v);
      asm volatile("bar.sync 1, 256;" ::: "memory");        // every converter has read its part of stage 0
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        uint8_t* a_st = smem + s * Cfg::STAGE_BYTES;
        if (dbg && threadIdx.x == 0 && kb < 16) g.dbg[8 + kb] = clock64();
        if (kb + 1 < KB) fetch(kb + 1, vn);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int row = p * 16 + rsub;
          const uint32_t off = row * 128 + ((((uint32_t)c4 >> 1) ^ ((uint32_t)row & 7)) << 4) + (c4 & 1) * 8;
          uint32_t h0, l0, h1, l1;
          split_pair(v[p].x, v[p].y, h0, l0);
          split_pair(v[p].z, v[p].w, h1, l1);
          *reinterpret_cast<uint2*>(a_st + off) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(a_st + 16384 + off) = make_uint2(l0, l1);
        }
        fence_proxy_async_smem();   // make generic-proxy stores visible to the tensor-core (async) proxy
        asm volatile("bar.sync 1, 256;" ::: "memory");      // reads of stage kb+1 done before anyone overwrites it
        if (lane == 0) mbar_arrive(&conv_full[s]);
#pragma unroll
        for (int p = 0; p < 8; ++p) v[p] = vn[p];
      }
    } else {
      // ===================== register gather (window / strided convs, unaligned operands) ==============
      RowGeom rg[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) rg[p] = row_geom(g, m0 + p * 16 + rsub);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        uint8_t* a_st = smem + s * Cfg::STAGE_BYTES;
        const int k0 = (kb0 + kb) * TC_BK + c4 * 4;
        float4 v[8];
#pragma unroll 1
        for (int p = 0; p < 8; ++p) {
          v[p].x = load_a_elem(g, rg[p], k0 + 0); v[p].y = load_a_elem(g, rg[p], k0 + 1);
          v[p].z = load_a_elem(g, rg[p], k0 + 2); v[p].w = load_a_elem(g, rg[p], k0 + 3);
        }
        mbar_wait(&empty_bar[s], ph ^ 1);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int row = p * 16 + rsub;
          const uint32_t off = row * 128 + ((((uint32_t)c4 >> 1) ^ ((uint32_t)row & 7)) << 4) + (c4 & 1) * 8;
          uint32_t h0, l0, h1, l1;
          split_pair(v[p].x, v[p].y, h0, l0);
          split_pair(v[p].z, v[p].w, h1, l1);
          *reinterpret_cast<uint2*>(a_st + off) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(a_st + 16384 + off) = make_uint2(l0, l1);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_full[s]);
      }
    }
    // ===================== epilogue: TMEM -> smem transpose -> coalesced global ==========================
    if (dbg && threadIdx.x == 0) g.dbg[1] = clock64();
    mbar_wait(accum_bar, 0);
    tcgen05_fence_after();
    if (dbg && threadIdx.x == 0) g.dbg[2] = clock64();
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = warp >> 2;             // column half
    const int row_own = m0 + q * 32 + lane; // the row whose accumulator this thread loads
    const RowGeom er = row_geom(g, row_own);
    const int myvalid = (er.valid && row_own < g.M) ? 1 : 0;
    float* st = reinterpret_cast<float*>(smem + warp * (32 * 33 * 4));   // pipeline buffers are idle now
    const int rows_here = min(32, g.M - (m0 + q * 32));
    if (g.epi_direct) epilogue_direct<BN>(g, Cz, tmem_base, m0, n0, q, half, lane, myvalid != 0, smem + warp * 4096, &tmapC);   // pipeline buffers are idle now
    else
#pragma unroll 1
    for (int cc = 0; cc < BN / 2; cc += 32) {
      const int col = half * (BN / 2) + cc;
      if (n0 + col >= g.n_out) break;
      uint32_t r[32];
      if (dbg && threadIdx.x == 0 && cc == 0) g.dbg[40] = clock64();
      // residual rows of this chunk: all 32 loads in flight before the TMEM read and the transpose (C may alias res, so
      // the compiler cannot hoist them over the stores itself; each element is read and written by the same thread)
      float rres[32];
      const bool hot_res = g.res && !g.swiglu && !g.C2 && !g.Chi && !g.accumulate;
      if (hot_res) {
        const int nh = n0 + col + lane;
        const float* rb = g.res + (long)(m0 + q * 32) * g.ldr + nh;
#pragma unroll
        for (int u = 0; u < 32; ++u) rres[u] = (nh < g.n_out && u < rows_here) ? rb[(long)u * g.ldr] : 0.f;
      } else {
#pragma unroll
        for (int u = 0; u < 32; ++u) rres[u] = 0.f;
      }
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
      tmem_ld_wait();
      if (dbg && threadIdx.x == 0 && cc == 0) g.dbg[41] = clock64();
#pragma unroll
      for (int j = 0; j < 32; ++j) st[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      if (dbg && threadIdx.x == 0 && cc == 0) g.dbg[42] = clock64();
      if (g.swiglu) {
        // columns (2j, 2j+1) = (gate_j, up_j) -> out[j] = silu(gate) * up ; lanes 0..15 own one output column each
        const int nin = n0 + col + 2 * lane;
        const int nout = (n0 + col) / 2 + lane;
        const bool colok = lane < 16 && nin < g.n_out;
        const float b0 = (colok && g.bias) ? g.bias[nin] : 0.f, b1 = (colok && g.bias) ? g.bias[nin + 1] : 0.f;
        const unsigned vmask = __ballot_sync(0xffffffffu, myvalid != 0);
#pragma unroll 1
        for (int rr = 0; rr < rows_here; ++rr) {
          const bool rv = (vmask >> rr) & 1u;
          if (colok) {
            const float v0 = st[rr * 33 + 2 * lane] * g.alpha + b0, v1 = st[rr * 33 + 2 * lane + 1] * g.alpha + b1;
            float v = act_apply_slow(ACT_SILU, v0, 0.f) * v1;
            if (!rv) v = 0.f;
            if (g.Chi && g.c_half) {        // one fp16 plane
              reinterpret_cast<__half*>(g.Chi)[(long)(m0 + q * 32 + rr) * g.ldcb + nout] = __float2half_rn(v);
            } else if (g.Chi) {
              __nv_bfloat16 hh, ll;
              split_bf16(v, hh, ll);
              g.Chi[(long)(m0 + q * 32 + rr) * g.ldcb + nout] = hh;
              g.Clo[(long)(m0 + q * 32 + rr) * g.ldcb + nout] = ll;
            } else {
              g.C[(long)(m0 + q * 32 + rr) * g.ldc + nout] = v;
            }
          }
        }
      } else {
        const int n = n0 + col + lane;
        const bool colok = n < g.n_out;
        const float bias = (colok && g.bias) ? g.bias[n] : 0.f;
        const float ap = (colok && g.act_vec) ? g.act_vec[n] : g.act_p;
        const float ap2 = (colok && g.act2_vec) ? g.act2_vec[n] : g.act2_p;
        const unsigned vmask = __ballot_sync(0xffffffffu, myvalid != 0);     // bit rr = row rr of this warp is a real row
        const long grow0 = m0 + q * 32;
        if (!g.C2 && !g.Chi && !g.accumulate) {
          // ---- hot case: C = act(acc*alpha + bias) (+ res), coalesced 128-byte rows
          const bool pre = g.act != ACT_NONE && colok;
          if (pre) {
            // activation pass in place on the staged tile: ONE inline copy of each activation in a rolled loop
            // (calls from the store loop spill its live registers; inlining into the unrolled loop thrashes the I-cache)
#pragma unroll 1
            for (int rr = 0; rr < rows_here; ++rr) {
              float v = st[rr * 33 + lane] * g.alpha + bias;
              switch (g.act) {
                case ACT_GELU: v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); break;
                case ACT_SILU: v = v / (1.0f + expf(-v)); break;
                case ACT_LRELU: v = v > 0.f ? v : v * ap; break;
                case ACT_SNAKE: { const float sn = sinf(v * ap); v = v + (1.0f / (ap + 1e-9f)) * (sn * sn); } break;
                case ACT_ELU: v = v > 0.f ? v : expm1f(v); break;
                default: v = act_apply_slow(g.act, v, ap); break;
              }
              st[rr * 33 + lane] = v;
            }
          }
          const float a_mul = pre ? 1.0f : g.alpha, a_add = pre ? 0.0f : bias;
          float* cbase = Cz + grow0 * g.ldc + n;
#pragma unroll
          for (int r0 = 0; r0 < 32; r0 += 8) {
            if (r0 >= rows_here) break;
            float xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = (colok && (r0 + u) < rows_here) ? st[(r0 + u) * 33 + lane] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float v = (xv[u] * a_mul + a_add + rres[r0 + u]) * g.out_scale;
              if (!((vmask >> (r0 + u)) & 1u)) v = 0.f;
              if (colok && (r0 + u) < rows_here) cbase[(long)(r0 + u) * g.ldc] = v;
            }
          }
        } else if (g.Chi && !g.C && !g.C2 && !g.res) {
          // ---- bf16 hi/lo planes only (QKV projection feeding the tcgen05 attention)
          const bool pre = g.act != ACT_NONE && colok;
          if (pre) {       // same in-place activation pass as the fp32 path (GELU of ff1 feeding ff2's plane operand)
#pragma unroll 1
            for (int rr = 0; rr < rows_here; ++rr) {
              float v = st[rr * 33 + lane] * g.alpha + bias;
              switch (g.act) {
                case ACT_GELU: v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); break;
                case ACT_SILU: v = v / (1.0f + expf(-v)); break;
                default: v = act_apply_slow(g.act, v, ap); break;
              }
              st[rr * 33 + lane] = v;
            }
          }
          const float a_mul = pre ? 1.0f : g.alpha, a_add = pre ? 0.0f : bias;
          __nv_bfloat16* hb = g.Chi + grow0 * g.ldcb + n;
          __nv_bfloat16* lb = g.Clo + grow0 * g.ldcb + n;
          for (int r0 = 0; r0 < rows_here; r0 += 8) {
            float xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = (colok && (r0 + u) < rows_here) ? st[(r0 + u) * 33 + lane] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float v = (xv[u] * a_mul + a_add) * g.out_scale;
              if (!((vmask >> (r0 + u)) & 1u)) v = 0.f;
              if (g.c_half) {          // one fp16 plane
                if (colok && (r0 + u) < rows_here) reinterpret_cast<__half*>(hb)[(long)(r0 + u) * g.ldcb] = __float2half_rn(v);
                continue;
              }
              __nv_bfloat16 h, l;
              split_bf16(v, h, l);
              if (colok && (r0 + u) < rows_here) { hb[(long)(r0 + u) * g.ldcb] = h; lb[(long)(r0 + u) * g.ldcb] = l; }
            }
          }
        } else {
          // ---- general case (second output, accumulate, planes + fp32)
#pragma unroll 1
          for (int rr = 0; rr < rows_here; ++rr) {
            if (!colok) continue;
            const long grow = grow0 + rr;
            const bool rv = (vmask >> rr) & 1u;
            float v = st[rr * 33 + lane] * g.alpha + bias;
            v = act_apply(g.act, v, ap);
            if (g.res) v += g.res[grow * g.ldr + n];
            v *= g.out_scale;
            if (!rv) v = 0.f;
            if (g.Chi) {
              __nv_bfloat16 h, l;
              split_bf16(v, h, l);
              g.Chi[grow * g.ldcb + n] = h;
              g.Clo[grow * g.ldcb + n] = l;
            }
            if (g.C) {
              if (g.accumulate && rv) v += g.C[grow * g.ldc + n];
              g.C[grow * g.ldc + n] = v;
            }
            if (g.C2) g.C2[grow * g.ldc2 + n] = rv ? act_apply(g.act2, v, ap2) : 0.f;
          }
        }
      }
      __syncwarp();
      if (dbg && threadIdx.x == 0 && cc == 0) g.dbg[43] = clock64();
    }
    if (dbg && threadIdx.x == 0) g.dbg[3] = clock64();
  } else if (warp == TC_PRODUCER_WARPS) {
    // ===================== TMA producer =================================================================
    if (lane == 0) {
      long in_row0 = 0;
      if (g.a_tma) {                       // input row feeding tap 0 of the tile's first output row
        if (g.has_seq) {
          const int sq = g.seq.tile_seq[m0 / kTileM];
          if (sq >= 0) in_row0 = (long)g.seq.in_start[sq] + (long)(m0 - g.seq.out_start[sq]) - g.pad;
        } else {
          in_row0 = (long)m0 - g.pad;
        }
      }
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_st = smem + s * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&tma_full[s], g.a_tma == 3 ? (16384 + Cfg::W_BYTES) : g.a_tma ? (Cfg::A_BYTES + Cfg::W_BYTES) : Cfg::W_BYTES);
        tma_load_2d(a_st + Cfg::A_BYTES, &tmapW, &tma_full[s], (kb0 + kb) * TC_BK, n0);
        if (g.a_tma >= 2) {
          // plane operands: (channel block, input row of this tap) -- for a plain Linear that is (k, m0); for a stride-1 conv
          // the same tile shifted by tap * dil rows.  Rows of other sequences never enter: the producers of conv planes write
          // zeros on layout padding rows and every sequence is followed by >= pad of them (engine.py layouts).
          const int k0 = (kb0 + kb) * TC_BK;
          const int tap = k0 / g.ctap, c = k0 - tap * g.ctap;
          const int arow = (int)(in_row0 + (long)tap * g.dil);
          tma_load_2d(a_st, &tmapA, &tma_full[s], c, arow);                              // hi plane (or the fp16 plane), SWIZZLE_128B
          if (g.a_tma == 2) tma_load_2d(a_st + 16384, &tmapA2, &tma_full[s], c, arow);   // lo plane
        } else if (g.a_tma) {
          const int k0 = (kb0 + kb) * TC_BK;
          const int tap = k0 / g.ctap, c = k0 - tap * g.ctap;
          tma_load_2d(a_st, &tmapA, &tma_full[s], c, (int)(in_row0 + (long)tap * g.dil));
        }
      }
    }
  } else {
    // ===================== MMA issuer (one elected thread) ============================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, BN);
      constexpr uint32_t idesc_a16 = idesc & ~(7u << 7) & ~(7u << 10);   // a_format = b_format = F16 (0): A fp16 x W fp16 copy
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&tma_full[s], ph);      // W tile landed (and A planes in plane mode)
        if (g.a_tma < 2) mbar_wait(&conv_full[s], ph);      // bf16 planes of A written by the converters
        tcgen05_fence_after();
        if (dbg && kb < 8) g.dbg[32 + kb] = clock64();
        const uint32_t a_hi = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t a_lo = a_hi + 16384;
        const uint32_t wb = a_hi + Cfg::A_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < TC_BK / 16; ++k4) {   // UMMA_K = 16 bf16 = 32 bytes inside the swizzle row
          const uint64_t db = umma_desc_sw128(wb + k4 * 32);
          if (g.a_tma == 3) {     // one fp16 plane: a single term
            umma_bf16(tmem_base, umma_desc_sw128(a_hi + k4 * 32), db, idesc_a16, (kb | k4) != 0 ? 1u : 0u);
            continue;
          }
          // small plane first so the fp32 accumulator adds the correction before the leading term
          umma_bf16(tmem_base, umma_desc_sw128(a_lo + k4 * 32), db, idesc, (kb | k4) != 0 ? 1u : 0u);
          umma_bf16(tmem_base, umma_desc_sw128(a_hi + k4 * 32), db, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);     // frees the smem stage once these MMAs have read it
      }
      umma_commit(accum_bar);           // accumulator complete
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == TC_PRODUCER_WARPS + 1) tmem_dealloc<BN>(tmem_base);
  if (dbg && threadIdx.x == 0) g.dbg[4] = clock64();
}

// ================================================================================================
// gemm_wres_kernel: persistent, WEIGHT-RESIDENT tcgen05 GEMM for the CFM transformer-block projections (K <= 512):
//     C[M, N] = epilogue( A16[M, K] (one fp16 plane, TMA) x W16[N, K]^T (fp16 copy of the weight) )
// Why: with K = 256 a 128 x 128 output tile moves 128 KB of operands through L2 -> shared memory for 1024 cycles of MMA;
// at ~42 B/clk per SM of L2 read bandwidth that alone caps the one-tile-per-CTA kernel at a third of the tensor peak, and
// the TMEM allocation, barrier set-up and epilogue of every tile sit on top.  Here a CTA loads its BN x K weight panel ONCE
// (128 KB), then walks the row tiles: A streams through a 4-stage ring (16 KB per K block; 231.7 KB of shared memory in all), the accumulator is double
// buffered in TMEM (2 x BN columns) so the epilogue of tile i runs under the MMAs of tile i+1, and the epilogue stores each
// thread's 32-column row segment straight from registers (whole 32-byte sectors, no shared-memory transpose).
//   warp 0: TMA producer   warp 1: MMA issuer (+ TMEM owner)   warps 2-9: epilogue (lane quarter = warp % 4, column half = (warp-2)/4)
// Algorithmic HBM bytes per launch: M*K*2 (A) + N*K*2 (W, once) + outputs (+ residual).
// ================================================================================================
constexpr int WR_THREADS = 320, WR_STAGES = 4, WR_W_BYTES = 131072, WR_A_STAGE = 16384;
constexpr int WR_STG = 8 * 4096;        // per epilogue warp: one 32-row x 128-byte staging tile of the TMA-store epilogue
constexpr int WR_SMEM = WR_W_BYTES + WR_STAGES * WR_A_STAGE + WR_STG + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*align*/;

__host__ __device__ constexpr uint32_t umma_idesc_f16f16(int M, int N) {       // fp16 x fp16 -> f32, K-major A and B
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int BN>
__global__ void __launch_bounds__(WR_THREADS, 1)
gemm_wres_kernel(const __grid_constant__ CUtensorMap tmapW, const __grid_constant__ CUtensorMap tmapA,
                 const __grid_constant__ CUtensorMap tmapC, const GemmDev g, const int n_mtiles, const int n_panels) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                   // [KB][BN rows x 128 B]  (SWIZZLE_128B)
  uint8_t* sA = smem + WR_W_BYTES;                      // [WR_STAGES][128 rows x 128 B]
  uint8_t* sStg = sA + WR_STAGES * WR_A_STAGE;          // [8 epilogue warps][32 rows x 128 B], SWIZZLE_128B (1024-byte aligned)
  float* sBias = reinterpret_cast<float*>(sStg + WR_STG);   // [BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + 1024);
  uint64_t* w_full = bars;                  // 1
  uint64_t* a_full = bars + 1;              // [WR_STAGES]
  uint64_t* a_empty = a_full + WR_STAGES;   // [WR_STAGES]
  uint64_t* acc_full = a_empty + WR_STAGES; // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = g.Kpad / TC_BK;
  const int panel = blockIdx.x % n_panels;
  const int n0 = panel * BN;
  const int mt0 = blockIdx.x / n_panels, mt_step = gridDim.x / n_panels;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < WR_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmapW); tma_prefetch_desc(&tmapA); tma_prefetch_desc(&tmapC); }
  for (int i = threadIdx.x; i < BN; i += WR_THREADS) sBias[i] = (g.bias && n0 + i < g.Npad) ? g.bias[n0 + i] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: the weight panel once, then the A tiles of every row tile =============
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, (uint32_t)(KB * BN * 128));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * BN * 128, &tmapW, w_full, kb * TC_BK, n0);
      uint32_t gi = 0;
      for (int mt = mt0; mt < n_mtiles; mt += mt_step) {
        for (int kb = 0; kb < KB; ++kb, ++gi) {
          const int s = gi % WR_STAGES;
          mbar_wait(&a_empty[s], ((gi / WR_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&a_full[s], WR_A_STAGE);
          tma_load_2d(sA + s * WR_A_STAGE, &tmapA, &a_full[s], kb * TC_BK, mt * TC_BM);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ===================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16f16(TC_BM, BN);
      mbar_wait(w_full, 0);
      uint32_t gi = 0;
      int it = 0;
      for (int mt = mt0; mt < n_mtiles; mt += mt_step, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < KB; ++kb, ++gi) {
          const int s = gi % WR_STAGES;
          mbar_wait(&a_full[s], (gi / WR_STAGES) & 1);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * WR_A_STAGE);
          const uint32_t w_addr = smem_u32(sW + kb * BN * 128);
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4)
            umma_bf16(d, umma_desc_sw128(a_addr + k4 * 32), umma_desc_sw128(w_addr + k4 * 32), idesc, (kb | k4) != 0 ? 1u : 0u);
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> swizzled staging tile -> TMA store ======================
    // One thread = one output row.  Its 128-byte groups (64 fp16 or 32 fp32 columns) go through a per-warp 32 x 128 B
    // staging tile and leave as ONE bulk tensor store: whole 128-byte lines reach L2, instead of 32 scattered 16-byte
    // pieces per store instruction (session 12: the K = 256 projections ran at ~1.9 TB/s, bound by the L2 request rate of
    // those partial-sector stores, not by HBM or the tensor pipe).  Rows beyond M are clipped by the tensor map.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* stg = sStg + (warp - 2) * 4096;
    const uint32_t stg_row = smem_u32(stg) + (uint32_t)lane * 128;
    const bool out16 = g.Chi != nullptr;                  // one fp16 plane (c_half), else fp32 (+ residual)
    int it = 0;
    for (int mt = mt0; mt < n_mtiles; mt += mt_step, ++it) {
      const int buf = it & 1;
      const int row = mt * TC_BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < BN / 2; cc += 32) {
        const int col = half * (BN / 2) + cc;
        const int n = n0 + col;
        if (n >= g.n_out) break;
        float rres[32];
        if (g.res && row_ok) {
          const float* rp = g.res + (long)row * g.ldr + n;
          if (g.epi_direct & 4) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) ldg256(rp + j, rres + j);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 t = reinterpret_cast<const float4*>(rp)[j]; rres[4 * j] = t.x; rres[4 * j + 1] = t.y; rres[4 * j + 2] = t.z; rres[4 * j + 3] = t.w; }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) rres[j] = 0.f;
        }
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + col), r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * g.alpha + sBias[col + j];
        if (g.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752440f));
        }
        // staging: 16-byte unit u of the row lands at unit (u ^ (row & 7)) of its 128-byte line (SWIZZLE_128B)
        const bool first = out16 ? ((cc & 32) == 0) : true;      // fp16: two 32-column chunks share one 128-byte group
        if (first) {                                             // the previous store has finished reading the staging tile
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        if (out16) {
          const int u0 = (cc & 32) ? 4 : 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t a = stg_row + ((uint32_t)((u0 + j) ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_h2(v[8 * j], v[8 * j + 1])), "r"(pack_h2(v[8 * j + 2], v[8 * j + 3])),
                         "r"(pack_h2(v[8 * j + 4], v[8 * j + 5])), "r"(pack_h2(v[8 * j + 6], v[8 * j + 7])) : "memory");
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = stg_row + ((uint32_t)(j ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"((v[4 * j] + rres[4 * j]) * g.out_scale), "f"((v[4 * j + 1] + rres[4 * j + 1]) * g.out_scale),
                         "f"((v[4 * j + 2] + rres[4 * j + 2]) * g.out_scale), "f"((v[4 * j + 3] + rres[4 * j + 3]) * g.out_scale) : "memory");
          }
        }
        const bool last = out16 ? ((cc & 32) != 0) : true;
        if (last) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int c0 = out16 ? (n - 32) : n;                 // first column of the 128-byte group
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&tmapC), "r"(smem_u32(stg)), "r"(c0), "r"(mt * TC_BM + q * 32) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // the stores have left shared memory and completed
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * BN>(tmem_base);
}


// ================================================================================================
// gemm_stream_kernel: persistent tcgen05 GEMM for the decode-step projections (M <= a few row tiles, K = 1024 / 4096):
//     C[M, N] = epilogue( A16[M, K] (one fp16 plane, TMA) x W16[N, K]^T (fp16 copy of the weight, TMA) )
// Why a second kernel: at 512 rows a projection is 2-9 GFLOP against 2-17 MB of weights -- a few microseconds of tensor or
// HBM time -- and the one-tile-per-CTA kernel spends several times that on everything else: two shallow stages (its shared
// memory is laid out for the fp32 A tile of the converter path), a cold pipeline per tile, TMEM allocation and barrier set-up
// per tile, an epilogue that starts only after the last MMA.  Here each CTA owns a list of (row tile, column tile, K split)
// items and keeps ONE pipeline running across them: 6 stages of (A 16 KB + W BN x 128 B) always in flight -- the loads of
// the next item start while the current one is still in its MMAs -- the accumulator is double buffered in TMEM so the
// register-direct epilogue (epilogue_direct) of item i runs under the main loop of item i+1, and all set-up happens once.
// Items are ordered row-tile-fastest: the CTAs that share a weight tile run at the same time and it crosses HBM once.
//   warp 0: TMA producer   warp 1: MMA issuer (+ TMEM owner)   warps 2-9: epilogue
// Algorithmic HBM bytes per launch: M*K*2 (A) + N*K*2 (W, once) + outputs.
// ================================================================================================
template <int BN> struct StreamCfg {
  static constexpr int STAGES = BN == 64 ? 8 : (BN == 128 ? 6 : 4);
  static constexpr int STAGE_BYTES = 16384 + BN * 128;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*align*/;
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <int BN>
__global__ void __launch_bounds__(WR_THREADS, 1)
gemm_stream_kernel(const __grid_constant__ CUtensorMap tmapW, const __grid_constant__ CUtensorMap tmapA, const GemmDev g,
                   const int n_ntiles) {
  using Cfg = StreamCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + STAGES;          // [STAGES]
  uint64_t* acc_full = empty + STAGES;      // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmapW); tma_prefetch_desc(&tmapA); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  // Only the TMA producer waits for the preceding kernel (griddepcontrol.wait, below): everything the other warps touch is
  // downstream of its A loads.  m_live is read before that wait on purpose -- it is written by the step's first kernel
  // (t3_compact) and every projection sits at least two launches later, so the writer completed before our predecessor started.
  // work list: live row tiles x column tiles x K splits, row tile fastest
  const int m_rows = g.m_live ? min(*g.m_live, g.M) : g.M;
  const int n_mt = (m_rows + TC_BM - 1) / TC_BM;
  const int splitk = g.splitk > 1 ? g.splitk : 1;
  const int n_items = n_mt * n_ntiles * splitk;
  const int KBt = g.Kpad / TC_BK;
  auto item = [&](int it, int& m0, int& n0, int& kb0, int& KB, int& ks) {
    const int mt = it % n_mt;
    const int rest = it / n_mt;
    const int nt = rest % n_ntiles;
    ks = rest / n_ntiles;
    m0 = mt * TC_BM; n0 = nt * BN;
    kb0 = (int)((long)ks * KBt / splitk);
    KB = (int)((long)(ks + 1) * KBt / splitk) - kb0;
  };

  if (warp == 0) {
    // ===================== TMA producer =================================================================
    if (lane == 0) {
      uint32_t gi = 0;
      // The weights do not depend on the previous kernel: the W tiles of the first item's first stages are already in flight
      // (HBM latency, ~1.5 us) while that kernel drains; the A tiles of the same stages follow the wait.
      int pre = 0;
      if ((int)blockIdx.x < n_items) {
        int m0, n0, kb0, KB, ks;
        item(blockIdx.x, m0, n0, kb0, KB, ks);
        pre = KB < STAGES ? KB : STAGES;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_arrive_expect_tx(&full[kb], Cfg::STAGE_BYTES);
          tma_load_2d(smem + kb * Cfg::STAGE_BYTES + 16384, &tmapW, &full[kb], (kb0 + kb) * TC_BK, n0);
        }
      }
      pdl_wait();                               // the producer of A has completed
      bool first = true;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        int m0, n0, kb0, KB, ks;
        item(it, m0, n0, kb0, KB, ks);
        for (int kb = 0; kb < KB; ++kb, ++gi) {
          const int s = gi % STAGES;
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          if (first && kb < pre) {               // W already requested, barrier already armed
            tma_load_2d(st, &tmapA, &full[s], (kb0 + kb) * TC_BK, m0);
            continue;
          }
          mbar_wait(&empty[s], ((gi / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
          tma_load_2d(st, &tmapA, &full[s], (kb0 + kb) * TC_BK, m0);
          tma_load_2d(st + 16384, &tmapW, &full[s], (kb0 + kb) * TC_BK, n0);
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ===================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16f16(TC_BM, BN);
      uint32_t gi = 0;
      int i = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++i) {
        int m0, n0, kb0, KB, ks;
        item(it, m0, n0, kb0, KB, ks);
        const int buf = i & 1;
        mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);          // the epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < KB; ++kb, ++gi) {
          const int s = gi % STAGES;
          mbar_wait(&full[s], (gi / STAGES) & 1);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t w_addr = a_addr + 16384;
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4)
            umma_bf16(d, umma_desc_sw128(a_addr + k4 * 32), umma_desc_sw128(w_addr + k4 * 32), idesc, (kb | k4) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =========================================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int i = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++i) {
      int m0, n0, kb0, KB, ks;
      item(it, m0, n0, kb0, KB, ks);
      const int buf = i & 1;
      mbar_wait(&acc_full[buf], (i >> 1) & 1);
      tcgen05_fence_after();
      float* Cz = g.C ? g.C + (long)ks * g.split_stride : nullptr;
      epilogue_direct<BN>(g, Cz, tmem_base + (uint32_t)(buf * BN), m0, n0, q, half, lane, m0 + q * 32 + lane < g.M);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * BN>(tmem_base);
}

// ================================================================================================
// SIMT reference tiles (debug only)
// ================================================================================================
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmDev g) {
  __shared__ float As[16][64 + 1];
  __shared__ float Ws[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < g.Kpad; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int r = i >> 4, k = i & 15;
      RowGeom rg = row_geom(g, m0 + r);
      As[k][r] = load_a_elem(g, rg, k0 + k);
      int n = n0 + r;
      Ws[k][r] = (n < g.Npad) ? __bfloat162float(g.Wp[(long)n * g.Kpad + k0 + k]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    int row = m0 + ty * 4 + i;
    if (row >= g.M) continue;
    RowGeom rg = row_geom(g, row);
    if (g.swiglu) {
      for (int j = 0; j < 4; j += 2) {
        int n = n0 + tx * 4 + j;
        if (n < g.n_out) epilogue_store_swiglu(g, row, n, acc[i][j], acc[i][j + 1], rg.valid);
      }
    } else {
      for (int j = 0; j < 4; ++j) epilogue_store(g, row, n0 + tx * 4 + j, acc[i][j], rg.valid);
    }
  }
}

// ================================================================================================
// GEMV (few rows, Linear only): CTA = 4 warps; each CTA owns NB output columns; warp w streams the K
// slices [w*256 + i*1024, +256) with one 16-byte (8 x bf16) load per lane per column.
// ================================================================================================
template <int R, int NB>
__global__ void __launch_bounds__(128) gemv_kernel(const GemmDev g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb0 = blockIdx.x * NB;
  const int K = g.Kpad;
  // The weights do not depend on the previous kernel: the first K slice of every column is already in flight while that
  // kernel drains (PDL) -- for the 1024-wide projections that is the CTA's whole weight traffic.
  const int kf = warp * 256 + lane * 8;
  uint4 wf[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int n = nb0 + j;
    wf[j] = (n < g.Npad && kf < K) ? __ldg(reinterpret_cast<const uint4*>(g.Wp + (long)n * K + kf)) : make_uint4(0, 0, 0, 0);
  }
  pdl_wait();
  pdl_launch_dependents();
  // optional fused row norm of the raw residual stream (dim = k_total <= 1024): statistics once per CTA
  __shared__ float s_stat[R][2];
  if (g.norm_w) {
    for (int r = warp; r < R; r += 4) {
      float s1 = 0.f, s2 = 0.f;
      if (r < g.M) {
        for (int k = lane * 4; k < g.k_total; k += 128) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(g.A + (long)r * g.lda + k));
          s1 += (a.x + a.y) + (a.z + a.w);
          s2 += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
        }
      }
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if (lane == 0) {
        if (g.norm_ln) {
          const float mean = s1 / g.k_total;
          float var = 0.f;            // second pass for the variance (matches the two-pass LayerNorm kernels)
          s_stat[r][0] = mean; s_stat[r][1] = var;
        } else {
          s_stat[r][0] = 0.f; s_stat[r][1] = rsqrtf(s2 / g.k_total + g.norm_eps);
        }
      }
    }
    __syncthreads();
    if (g.norm_ln) {
      for (int r = warp; r < R; r += 4) {
        float v = 0.f;
        const float mean = s_stat[r][0];
        if (r < g.M) {
          for (int k = lane * 4; k < g.k_total; k += 128) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(g.A + (long)r * g.lda + k));
            const float d0 = a.x - mean, d1 = a.y - mean, d2 = a.z - mean, d3 = a.w - mean;
            v += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
          }
        }
        v = warp_sum(v);
        if (lane == 0) s_stat[r][1] = rsqrtf(v / g.k_total + g.norm_eps);
      }
      __syncthreads();
    }
  }
  float acc[R][NB];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[r][j] = 0.0f;
  for (int k = kf; k < K; k += 1024) {
    uint4 w[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      int n = nb0 + j;
      w[j] = (k == kf) ? wf[j] : ((n < g.Npad) ? __ldg(reinterpret_cast<const uint4*>(g.Wp + (long)n * K + k)) : make_uint4(0, 0, 0, 0));
    }
    float x[R][8];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r < g.M && k < g.k_total) {   // k_total == logical K (multiple of 8 for every Linear on this path)
        const float4 a = __ldg(reinterpret_cast<const float4*>(g.A + (long)r * g.lda + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(g.A + (long)r * g.lda + k + 4));
        x[r][0] = a.x; x[r][1] = a.y; x[r][2] = a.z; x[r][3] = a.w;
        x[r][4] = b.x; x[r][5] = b.y; x[r][6] = b.z; x[r][7] = b.w;
        if (g.norm_w) {
          const float4 wa = __ldg(reinterpret_cast<const float4*>(g.norm_w + k));
          const float4 wb = __ldg(reinterpret_cast<const float4*>(g.norm_w + k + 4));
          const float nw[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
          const float mean = s_stat[r][0], inv = s_stat[r][1];
          if (g.norm_ln) {
            const float4 ba = __ldg(reinterpret_cast<const float4*>(g.norm_b + k));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(g.norm_b + k + 4));
            const float nb[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) x[r][i] = (x[r][i] - mean) * inv * nw[i] + nb[i];
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[r][i] = nw[i] * (x[r][i] * inv);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[r][i] = 0.0f;
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const uint32_t ww[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float w0 = __uint_as_float(ww[i] << 16);            // bf16 -> fp32 is a 16-bit shift
        const float w1 = __uint_as_float(ww[i] & 0xFFFF0000u);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[r][j] = fmaf(x[r][2 * i], w0, acc[r][j]);
          acc[r][j] = fmaf(x[r][2 * i + 1], w1, acc[r][j]);
        }
      }
    }
  }
  __shared__ float red[4][R][NB];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float v = warp_sum(acc[r][j]);
      if (lane == 0) red[warp][r][j] = v;
    }
  __syncthreads();
  if (threadIdx.x < R * NB) {
    const int r = threadIdx.x / NB, j = threadIdx.x % NB;
    if (r < g.M) {
      if (g.swiglu) {
        if ((j & 1) == 0) {
          float a0 = red[0][r][j] + red[1][r][j] + red[2][r][j] + red[3][r][j];
          float a1 = red[0][r][j + 1] + red[1][r][j + 1] + red[2][r][j + 1] + red[3][r][j + 1];
          if (nb0 + j < g.n_out) epilogue_store_swiglu(g, r, nb0 + j, a0, a1, true);
        }
      } else {
        float a = red[0][r][j] + red[1][r][j] + red[2][r][j] + red[3][r][j];
        epilogue_store(g, r, nb0 + j, a, true);
      }
    }
  }
}

// ================================================================================================
// host side
// ================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cbx: cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

void make_tmaps_for(Weight& W) {
  PFN_encodeTiled enc = get_encode_fn();
  const int boxes[3] = {64, 128, 256};
  for (int i = 0; i < 3; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)W.Kpad, (cuuint64_t)W.Npad};
    cuuint64_t strides[1] = {(cuuint64_t)W.Kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)boxes[i]};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&W.tmap[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W.w, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cbx: cuTensorMapEncodeTiled failed");
  }
}

// 2-D map over a [rows][cols] bf16 plane, box 64 x 64, SWIZZLE_128B (tcgen05 attention operands)
void make_plane_tmap(CUtensorMap* tm, const __nv_bfloat16* base, long rows, int cols, int box_rows, int ld) {
  PFN_encodeTiled enc = get_encode_fn();
  if (ld == 0) ld = cols;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cbx: cuTensorMapEncodeTiled (plane) failed");
}

static inline uint16_t f2bf16_host(float f) {   // round-to-nearest-even
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
  uint32_t lsb = (u >> 16) & 1;
  u += 0x7FFFu + lsb;
  return (uint16_t)(u >> 16);
}

static void upload_packed(Weight& W, const std::vector<uint16_t>& host, const float* host_bias) {
  CBX_CHECK(cudaMalloc(&W.w, host.size() * 2));
  CBX_CHECK(cudaMemcpy(W.w, host.data(), host.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> b(W.Npad, 0.0f);
  if (host_bias) for (int i = 0; i < W.N; ++i) b[i] = host_bias[i];
  CBX_CHECK(cudaMalloc(&W.bias, W.Npad * 4));
  CBX_CHECK(cudaMemcpy(W.bias, b.data(), W.Npad * 4, cudaMemcpyHostToDevice));
  make_tmaps_for(W);
}

void pack_linear(Weight& W, const float* w, const float* bias, int N, int K, bool half_copy) {
  W.N = N; W.K = K; W.Npad = (N + 63) / 64 * 64; W.Kpad = (K + 63) / 64 * 64;
  std::vector<uint16_t> h((size_t)W.Npad * W.Kpad, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) h[(size_t)n * W.Kpad + k] = f2bf16_host(w[(size_t)n * K + k]);
  upload_packed(W, h, bias);
  if (half_copy) {       // the bf16-rounded values again, as fp16
    std::vector<__half> hh((size_t)W.Npad * W.Kpad, __float2half_rn(0.f));
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        uint32_t u = (uint32_t)h[(size_t)n * W.Kpad + k] << 16;
        float f; memcpy(&f, &u, 4);
        hh[(size_t)n * W.Kpad + k] = __float2half_rn(f);
      }
    CBX_CHECK(cudaMalloc(&W.w16, hh.size() * 2));
    CBX_CHECK(cudaMemcpy(W.w16, hh.data(), hh.size() * 2, cudaMemcpyHostToDevice));
    Weight tmp = W; tmp.w = reinterpret_cast<__nv_bfloat16*>(W.w16);     // 2-byte elements: same map geometry
    make_tmaps_for(tmp);
    for (int i = 0; i < 3; ++i) W.tmap16[i] = tmp.tmap[i];
  }
}
// torch Conv1d weight [N][cin][taps] -> [N][tap][ctap], ctap = cin rounded up to 64
void pack_conv_taps(Weight& W, const float* w, const float* bias, int N, int cin, int taps) {
  const int ctap = (cin + 63) / 64 * 64;
  W.N = N; W.K = taps * ctap; W.Npad = (N + 63) / 64 * 64; W.Kpad = W.K;
  std::vector<uint16_t> h((size_t)W.Npad * W.Kpad, 0);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < taps; ++t)
        h[(size_t)n * W.Kpad + (size_t)t * ctap + c] = f2bf16_host(w[((size_t)n * cin + c) * taps + t]);
  upload_packed(W, h, bias);
}
// torch Conv1d weight [N][cin][taps] -> [N][tap*cin + c]  (contiguous window of a channel-last input)
void pack_conv_window(Weight& W, const float* w, const float* bias, int N, int cin, int taps) {
  W.N = N; W.K = taps * cin; W.Npad = (N + 63) / 64 * 64; W.Kpad = (W.K + 63) / 64 * 64;
  std::vector<uint16_t> h((size_t)W.Npad * W.Kpad, 0);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < taps; ++t)
        h[(size_t)n * W.Kpad + (size_t)t * cin + c] = f2bf16_host(w[((size_t)n * cin + c) * taps + t]);
  upload_packed(W, h, bias);
}
void free_weight(Weight& W) {
  if (W.w) cudaFree(W.w);
  if (W.bias) cudaFree(W.bias);
  if (W.w16) cudaFree(W.w16);
  W.w = nullptr; W.bias = nullptr; W.w16 = nullptr;
}

GemmDev gemm_args_linear(const float* A, int lda, int M, const Weight& W, float* C, int ldc) {
  GemmDev g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.M = M; g.M_in = M;
  g.a_mode = A_TAPS; g.ntaps = 1; g.ctap = W.Kpad; g.c_in = W.K; g.dil = 0; g.pad = 0; g.stride = 1;
  g.k_total = W.K;
  g.has_seq = 0;
  g.Wp = W.w; g.Kpad = W.Kpad; g.Npad = W.Npad;
  g.C = C; g.ldc = ldc; g.n_out = W.N; g.bias = W.bias; g.alpha = 1.0f;
  g.act = ACT_NONE; g.out_scale = 1.0f;
  return g;
}

static void make_a_tmap(CUtensorMap* tm, const GemmDev& g) {
  PFN_encodeTiled enc = get_encode_fn();
  // [M_in rows][c_in cols] fp32 view of the activation matrix; columns >= c_in and rows outside [0, M_in) read as zero
  cuuint64_t dims[2] = {(cuuint64_t)g.c_in, (cuuint64_t)g.M_in};
  cuuint64_t strides[1] = {(cuuint64_t)g.lda * 4};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g.A), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cbx: cuTensorMapEncodeTiled (A operand) failed");
}

// opt every tcgen05 instantiation into its dynamic shared memory size on the CURRENT device (called once per handle,
// before any launch or stream capture)
template <int BN, int DUAL> static void set_tc_attr() {
  CBX_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN, DUAL>::SMEM));
}
void gemm_init() {
  set_tc_attr<64, 0>(); set_tc_attr<64, 1>(); set_tc_attr<128, 0>(); set_tc_attr<128, 1>(); set_tc_attr<256, 0>();
  CBX_CHECK(cudaFuncSetAttribute(gemm_wres_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_SMEM));
  CBX_CHECK(cudaFuncSetAttribute(gemm_wres_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_SMEM));
  CBX_CHECK(cudaFuncSetAttribute(gemm_stream_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, StreamCfg<64>::SMEM));
  CBX_CHECK(cudaFuncSetAttribute(gemm_stream_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, StreamCfg<128>::SMEM));
  CBX_CHECK(cudaFuncSetAttribute(gemm_stream_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, StreamCfg<256>::SMEM));
}

// register-direct epilogue (see epilogue_direct): every destination vector-aligned, one of the inline activations
static void set_epi_direct(GemmDev& g) {
  auto al = [](const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) & (uintptr_t)(a - 1)) == 0; };
  static const bool direct_on = !(getenv("CBX_EPI_DIRECT") && atoi(getenv("CBX_EPI_DIRECT")) == 0);
  const bool act_ok = !g.act_vec && (g.act == ACT_NONE || g.act == ACT_LRELU || g.act == ACT_SILU || g.act == ACT_GELU || g.act == ACT_GELU_TANH);
  bool ok = direct_on && act_ok && !g.C2 && !g.accumulate && !(g.Chi && g.C) && (g.Chi || g.C) && !g.dbg;
  if (g.swiglu) ok = ok && !g.res && g.act == ACT_NONE && (g.n_out % 2) == 0;
  if (g.C && !g.Chi) ok = ok && (g.ldc % 4) == 0 && al(g.C, 16) && (g.split_stride % 4) == 0;
  if (g.Chi) ok = ok && (g.ldcb % 8) == 0 && al(g.Chi, 16) && (g.c_half || al(g.Clo, 16));
  if (g.res) ok = ok && (g.ldr % 4) == 0 && al(g.res, 16);
  g.epi_direct = 0;
  if (ok) {
    g.epi_direct = 1;
    if (g.C && (g.ldc % 8) == 0 && al(g.C, 32) && (g.split_stride % 8) == 0) g.epi_direct |= 2;       // 32-byte stores
    if (g.res && (g.ldr % 8) == 0 && al(g.res, 32)) g.epi_direct |= 4;                                 // 32-byte residual loads
  }
}

template <int BN, int DUAL> static void launch_tc(Ctx& ctx, GemmDev g, const Weight& W, int tmap_idx) {
  // TMA-fed A operand: plain strided fp32 rows (Linear, stride-1 conv taps)
  g.a_tma = (g.a_mode == A_TAPS && g.stride == 1 && (g.lda % 4) == 0 && (g.c_in % 4) == 0 &&
             (reinterpret_cast<uintptr_t>(g.A) & 15) == 0 && g.M_in > 0) ? 1 : 0;
  CUtensorMap tmA, tmA2;
  if (g.A16) {
    CBX_REQUIRE(g.ntaps == 1 && !g.has_seq && (g.lda16 % 8) == 0, "fp16 plane operand needs a plain Linear");
    CBX_REQUIRE(W.w16 != nullptr, "fp16 activations need the fp16 copy of the weight (pack_linear half_copy)");
    g.a_tma = 3;
    make_plane_tmap(&tmA, reinterpret_cast<const __nv_bfloat16*>(g.A16), g.M, g.k_total, 128, g.lda16);   // 2-byte elements
    tmA2 = tmA;
  } else if (g.Ahi) {
    CBX_REQUIRE(g.a_mode == A_TAPS && g.stride == 1 && g.Alo && (g.ldab % 8) == 0, "plane operands need a Linear or a stride-1 conv");
    g.a_tma = 2;
    const long a_rows = (g.ntaps > 1 || g.has_seq) ? g.M_in : g.M;
    const int a_cols = (g.ntaps > 1 || g.has_seq) ? g.c_in : g.k_total;
    make_plane_tmap(&tmA, g.Ahi, a_rows, a_cols, 128, g.ldab);
    make_plane_tmap(&tmA2, g.Alo, a_rows, a_cols, 128, g.ldab);
  } else {
    if (g.a_tma) make_a_tmap(&tmA, g); else tmA = W.tmap[tmap_idx];
    tmA2 = tmA;
  }
  if (g.splitk < 1) g.splitk = 1;
  if (g.splitk > 1) {
    CBX_REQUIRE(g.C && !g.bias && !g.res && !g.C2 && !g.Chi && !g.accumulate && !g.swiglu && g.act == ACT_NONE && g.alpha == 1.0f &&
                g.out_scale == 1.0f && g.splitk <= g.Kpad / TC_BK, "split-K writes raw partial sums");
  }
  set_epi_direct(g);
  dim3 grid((g.Npad + BN - 1) / BN, (g.M + TC_BM - 1) / TC_BM, g.splitk);
  if (ctx.timer && ctx.timer->on(K_GEMM_TC)) {
    const double flops = 2.0 * (double)g.M * (double)g.n_out * (double)g.k_total;
    // algorithmic HBM bytes: weights once (bf16), the activation matrix once (fp32 or hi+lo planes = 4 B per element),
    // every output element once per destination, the residual once
    const double a_rows = (g.a_mode == A_TAPS && g.ntaps > 1) ? (double)g.M_in : (double)g.M * g.stride;
    double b = (double)g.Npad * g.Kpad * 2.0 + a_rows * (double)g.c_in * (g.A16 ? 2.0 : 4.0);
    const double out_elems = (double)g.M * (double)(g.swiglu ? g.n_out / 2 : g.n_out);
    if (g.C) b += out_elems * 4.0;
    if (g.Chi) b += out_elems * (g.c_half ? 2.0 : 4.0);
    if (g.C2) b += out_elems * 4.0;
    if (g.res) b += out_elems * 4.0;
    if (g.accumulate) b += out_elems * 4.0;
    ctx.timer->add(K_GEMM_TC, flops, b);
  }
  // fp32 result of a full-K launch: TMA-store epilogue (32-row x 128-byte boxes)
  CUtensorMap tmC = tmA;
  static const bool tma_store_on = !(getenv("CBX_TMA_STORE") && atoi(getenv("CBX_TMA_STORE")) == 0);
  if (tma_store_on && (g.epi_direct & 1) && g.C && !g.Chi && !g.swiglu && g.splitk == 1 && (g.ldc % 4) == 0 && g.M >= 4 * TC_BM) {
    PFN_encodeTiled enc = get_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)g.n_out, (cuuint64_t)g.M};
    cuuint64_t strides[1] = {(cuuint64_t)g.ldc * 4};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)g.C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) g.epi_direct |= 8;
  }
  if (ctx.timer) ctx.timer->begin(K_GEMM_TC, ctx.stream);
  launch_kernel(ctx, gemm_tc_kernel<BN, DUAL>, grid, dim3(TC_THREADS), (size_t)TcCfg<BN, DUAL>::SMEM,
                g.A16 ? W.tmap16[tmap_idx] : W.tmap[tmap_idx], tmA, tmA2, tmC, g);
  if (ctx.timer) ctx.timer->end(K_GEMM_TC, ctx.stream);
}

// weight-resident persistent kernel: eligible launches (see gemm())
template <int BN> static void launch_wres(Ctx& ctx, GemmDev g, const Weight& W) {
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; CBX_CHECK(cudaGetDevice(&dev)); CBX_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)); }
  CUtensorMap tmA;
  make_plane_tmap(&tmA, reinterpret_cast<const __nv_bfloat16*>(g.A16), g.M, g.k_total, 128, g.lda16);   // 2-byte elements
  const int n_mtiles = (g.M + TC_BM - 1) / TC_BM;
  const int n_panels = g.Npad / BN;
  int groups = n_sm / n_panels;                         // row-tile groups: every CTA of a group owns one weight panel
  if (groups > n_mtiles) groups = n_mtiles;
  if (groups < 1) groups = 1;
  if (ctx.timer && ctx.timer->on(K_WRES)) {
    double b = (double)g.Npad * g.Kpad * 2.0 + (double)g.M * g.k_total * 2.0 + (double)g.M * g.n_out * (g.Chi ? 2.0 : 4.0);
    if (g.res) b += (double)g.M * g.n_out * 4.0;
    ctx.timer->add(K_WRES, 2.0 * (double)g.M * (double)g.n_out * (double)g.k_total, b);
  }
  // output map of the TMA-store epilogue: 32-row x 128-byte boxes (64 fp16 or 32 fp32 columns), SWIZZLE_128B
  CUtensorMap tmC;
  {
    PFN_encodeTiled enc = get_encode_fn();
    const bool h16 = g.Chi != nullptr;
    cuuint64_t dims[2] = {(cuuint64_t)g.n_out, (cuuint64_t)g.M};
    cuuint64_t strides[1] = {h16 ? (cuuint64_t)g.ldcb * 2 : (cuuint64_t)g.ldc * 4};
    cuuint32_t box[2] = {h16 ? 64u : 32u, 32u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmC, h16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h16 ? (void*)g.Chi : (void*)g.C, dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cbx: cuTensorMapEncodeTiled (C store) failed");
  }
  set_epi_direct(g);       // (only the residual-load width bit is used by this kernel)
  if (ctx.timer) ctx.timer->begin(K_WRES, ctx.stream);
  gemm_wres_kernel<BN><<<groups * n_panels, WR_THREADS, WR_SMEM, ctx.stream>>>(W.tmap16[BN == 256 ? 2 : 1], tmA, tmC, g, n_mtiles, n_panels);
  if (ctx.timer) ctx.timer->end(K_WRES, ctx.stream);
}

// persistent streaming kernel for the decode-step projections: eligible launches (see gemm())
template <int BN> static void launch_stream(Ctx& ctx, GemmDev g, const Weight& W) {
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; CBX_CHECK(cudaGetDevice(&dev)); CBX_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)); }
  CUtensorMap tmA;
  make_plane_tmap(&tmA, reinterpret_cast<const __nv_bfloat16*>(g.A16), g.M, g.k_total, 128, g.lda16);   // 2-byte elements
  if (g.splitk < 1) g.splitk = 1;
  const int n_ntiles = (g.Npad + BN - 1) / BN;
  const int n_items = ((g.M + TC_BM - 1) / TC_BM) * n_ntiles * g.splitk;
  if (ctx.timer && ctx.timer->on(K_STREAM)) {
    const double out_elems = (double)g.M * (double)(g.swiglu ? g.n_out / 2 : g.n_out);
    double b = (double)g.Npad * g.Kpad * 2.0 + (double)g.M * g.k_total * 2.0;
    if (g.C) b += out_elems * 4.0 * g.splitk;
    if (g.Chi) b += out_elems * (g.c_half ? 2.0 : 4.0);
    if (g.res) b += out_elems * 4.0;
    ctx.timer->add(K_STREAM, 2.0 * (double)g.M * (double)g.n_out * (double)g.k_total, b);
  }
  if (ctx.timer) ctx.timer->begin(K_STREAM, ctx.stream);
  launch_kernel(ctx, gemm_stream_kernel<BN>, dim3(n_items < n_sm ? n_items : n_sm), dim3(WR_THREADS), (size_t)StreamCfg<BN>::SMEM,
                W.tmap16[BN == 64 ? 0 : (BN == 128 ? 1 : 2)], tmA, g, n_ntiles);
  if (ctx.timer) ctx.timer->end(K_STREAM, ctx.stream);
}

template <int R> static void launch_gemv(Ctx& ctx, const GemmDev& g) {
  if (ctx.timer) ctx.timer->begin(K_GEMV, ctx.stream);
  struct End { Ctx& c; ~End() { if (c.timer) c.timer->end(K_GEMV, c.stream); } } _end{ctx};
  if (g.Npad <= 2048) {
    launch_kernel(ctx, gemv_kernel<R, 2>, dim3((g.Npad + 1) / 2), dim3(128), 0, g);
  } else {
    launch_kernel(ctx, gemv_kernel<R, 4>, dim3((g.Npad + 3) / 4), dim3(128), 0, g);
  }
}

void gemm(Ctx& ctx, GemmDev g, const Weight& W) {
  CBX_REQUIRE(g.Kpad % 64 == 0 && g.Npad % 64 == 0, "padded weight dims");
  if (g.a_mode == A_TAPS) CBX_REQUIRE(g.ctap % 64 == 0, "TAPS mode needs 64-channel chunks");
  if (ctx.dry) return;
  CBX_REQUIRE(!g.c_half || (ctx.gemm_impl == 0 && g.Chi && !g.C && !g.C2 && !g.res && g.M > 8), "fp16 plane output needs the tcgen05 planes-only epilogue");
  CBX_REQUIRE(!g.A16 || W.w16 != nullptr, "fp16 activations need the fp16 copy of the weight");
  ctx.launches++;
  if (ctx.gemm_impl == 1) {
    dim3 grid((g.M + 63) / 64, (g.Npad + 63) / 64);
    gemm_simt_kernel<<<grid, 256, 0, ctx.stream>>>(g);
  } else {
    const bool plain = !g.has_seq && g.a_mode == A_TAPS && g.ntaps == 1 && g.stride == 1 && g.pad == 0 &&
                       (g.lda % 4 == 0) && (g.k_total % 8 == 0) && !g.C2 && !g.Chi && !g.Ahi && !g.A16 && (!g.norm_w || g.k_total <= 1024) &&
                       ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    // fp16-plane Linear with a short reduction (the CFM block projections qkv / ff1 / out): weight-resident persistent kernel
    static const bool wres_on = !(getenv("CBX_WRES") && atoi(getenv("CBX_WRES")) == 0);
    const bool wres_epi = !g.swiglu && !g.C2 && !g.accumulate && (g.act == ACT_NONE || g.act == ACT_GELU) && !g.act_vec &&
                          ((g.Chi && g.c_half && !g.C && !g.res) || (g.C && !g.Chi)) &&
                          (!g.C || ((g.ldc % 4) == 0 && (g.n_out % 32) == 0 && (!g.res || (g.ldr % 4) == 0))) &&
                          (!g.Chi || ((g.ldcb % 8) == 0 && (g.n_out % 64) == 0));      // TMA-store groups: 64 fp16 / 32 fp32 columns
    if (wres_on && g.A16 && W.w16 && !g.has_seq && g.ntaps == 1 && g.splitk <= 1 && wres_epi && g.M >= 4 * TC_BM) {
      if (g.Kpad == 256 && g.Npad % 256 == 0) { launch_wres<256>(ctx, g, W); CBX_CHECK(cudaGetLastError()); return; }
      if (g.Kpad == 512 && g.Npad % 128 == 0) { launch_wres<128>(ctx, g, W); CBX_CHECK(cudaGetLastError()); return; }
    }
    // fp16-plane Linear at decode size (a few row tiles, long reduction): persistent streaming kernel
    static const int stream_max_m = getenv("CBX_GEMM_STREAM") ? atoi(getenv("CBX_GEMM_STREAM")) : 1024;
    if (g.A16 && W.w16 && !g.has_seq && g.ntaps == 1 && g.M > 8 && g.M <= stream_max_m && g.Kpad >= 512) {
      GemmDev gs = g;
      if (gs.splitk < 1) gs.splitk = 1;
      set_epi_direct(gs);
      const bool split_ok = gs.splitk == 1 || (gs.C && !gs.bias && !gs.res && !gs.Chi && !gs.swiglu && gs.act == ACT_NONE && gs.alpha == 1.0f &&
                                               gs.out_scale == 1.0f && gs.splitk <= gs.Kpad / TC_BK);
      if (gs.epi_direct && split_ok) {
        const int bn = (g.tile_bn == 64 || g.tile_bn == 128 || g.tile_bn == 256) && g.Npad % g.tile_bn == 0 ? g.tile_bn : (g.Npad % 128 == 0 ? 128 : 64);
        if (bn == 64) launch_stream<64>(ctx, gs, W); else if (bn == 128) launch_stream<128>(ctx, gs, W); else launch_stream<256>(ctx, gs, W);
        CBX_CHECK(cudaGetLastError());
        return;
      }
    }
    CBX_REQUIRE(!g.norm_w || (plain && g.M <= 8), "the fused row norm exists in the GEMV kernel only");
    if (plain && g.M <= 8) {
      if (g.M <= 2) launch_gemv<2>(ctx, g);
      else if (g.M <= 4) launch_gemv<4>(ctx, g);
      else launch_gemv<8>(ctx, g);
    } else {
      const int mt = (g.M + TC_BM - 1) / TC_BM;
      // widest N tile that still gives every SM a tile (L2->SM traffic per flop falls with BN)
      static const int tile_mode = getenv("CBX_TILE") ? atoi(getenv("CBX_TILE")) : 0;   // 0 auto, 1 no dual (experiments)
      const long t64 = (long)mt * (g.Npad / 64);
      const long t128 = (g.Npad % 128 == 0) ? (long)mt * (g.Npad / 128) : 0;
      const long t256 = (g.Npad % 256 == 0) ? (long)mt * (g.Npad / 256) : 0;
      // fewest waves first: two CTAs per SM when there are >= 2 tiles per SM, else the widest tile that still spreads
      // over at least half of the SMs (per-tile time is latency-bound, so fewer, fatter tiles beat a second wave)
      if (g.tile_bn == 64) { if (g.tile_dual) launch_tc<64, 1>(ctx, g, W, 0); else launch_tc<64, 0>(ctx, g, W, 0); }
      else if (g.tile_bn == 128 && t128 > 0) { if (g.tile_dual) launch_tc<128, 1>(ctx, g, W, 1); else launch_tc<128, 0>(ctx, g, W, 1); }
      else if (g.tile_bn == 256 && t256 > 0) launch_tc<256, 0>(ctx, g, W, 2);
      else if (tile_mode == 0 && t128 >= 2 * 148) launch_tc<128, 1>(ctx, g, W, 1);
      else if (t256 >= 120) launch_tc<256, 0>(ctx, g, W, 2);
      else if (t128 >= 74) launch_tc<128, 0>(ctx, g, W, 1);
      else if (tile_mode == 0 && t64 >= 2 * 148) launch_tc<64, 1>(ctx, g, W, 0);
      else launch_tc<64, 0>(ctx, g, W, 0);
    }
  }
  CBX_CHECK(cudaGetLastError());
}

}  // namespace cbx
