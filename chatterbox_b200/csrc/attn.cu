// Attention kernels of libcbx.
//
//  * flash_attn_kernel      head_dim 64, fp32 in/out, online softmax in fp32; Q.K^T and P.V on the legacy
//                           tensor path (mma.sync m16n8k16 bf16) with every operand split into bf16 hi+lo
//                           (3 MMAs per product: hi.hi + hi.lo + lo.hi) so results track the fp32 reference.
//                           Options: causal, per-sequence kv length, additive bias incl. espnet rel-pos shift.
//                           Used by: T3 prefill, conformer encoder, CFM estimator blocks.
//  * attn_simt_kernel       same contract on CUDA cores (debug reference, CBX_ATTN=simt).
//  * attn_generic_kernel    tiny generic-head-dim attention (perceiver resampler, 4 heads x 256).
//  * paged_decode_kernel    one query token per row against the paged KV cache (HBM-bound: every K/V byte is
//                           read exactly once, 16-byte coalesced loads, flash-decoding split over CTAs).
//  * rope_store_kernel      llama3 RoPE on q,k (in place) + append k,v to the paged cache.
#include "ops.h"
#include <cuda_fp8.h>
#include <cstdlib>
#include <string>

namespace cbx {

// ================================================================================================
// flash attention (mma.sync bf16, 3-term split)
// ================================================================================================
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat16 h0, l0, h1, l1;
  split_bf16(x, h0, l0);
  split_bf16(y, h1, l1);
  hi = pack_bf16(h0, h1);
  lo = pack_bf16(l0, l1);
}

constexpr int FA_BM = 64, FA_BN = 64, FA_LD = 72;   // smem row stride (bf16) - conflict-free fragment loads

struct FlashDev {
  const float* Q; const float* K; const float* V; float* O;
  int ldq, ldk, ldv, ldo;
  const int* q_start; const int* q_len; const int* kv_start; const int* kv_len;
  float scale; int causal;
  const float* bias; long bias_head_stride; int bias_ld; long bias_row0; int bias_center; int bias_rel;
};

__global__ void __launch_bounds__(128) flash_attn_kernel(const FlashDev p) {
  const int seq = blockIdx.z, head = blockIdx.y;
  const int qlen = p.q_len[seq], kvlen = p.kv_len[seq];
  const int q0 = blockIdx.x * FA_BM;
  if (q0 >= qlen) return;
  const long qrow0 = p.q_start[seq], krow0 = p.kv_start[seq];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  __shared__ __align__(16) __nv_bfloat16 Kh[FA_BN * FA_LD], Kl[FA_BN * FA_LD];   // [key][d]
  __shared__ __align__(16) __nv_bfloat16 Vh[64 * FA_LD], Vl[64 * FA_LD];         // [d][key] (transposed)

  // ---- Q fragments (A operand, 16 rows x 64 d per warp), hi and lo planes
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;     // local query indices of this thread's two rows
  uint32_t qh[4][4], ql[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int d0 = ks * 16 + 2 * t;
    float2 a = make_float2(0.f, 0.f), b = a, c = a, d = a;
    if (r0 < qlen) {
      const float* q = p.Q + (qrow0 + r0) * p.ldq + head * 64;
      a = *reinterpret_cast<const float2*>(q + d0);
      c = *reinterpret_cast<const float2*>(q + d0 + 8);
    }
    if (r1 < qlen) {
      const float* q = p.Q + (qrow0 + r1) * p.ldq + head * 64;
      b = *reinterpret_cast<const float2*>(q + d0);
      d = *reinterpret_cast<const float2*>(q + d0 + 8);
    }
    split2(a.x, a.y, qh[ks][0], ql[ks][0]);
    split2(b.x, b.y, qh[ks][1], ql[ks][1]);
    split2(c.x, c.y, qh[ks][2], ql[ks][2]);
    split2(d.x, d.y, qh[ks][3], ql[ks][3]);
  }

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int coff = kvlen - qlen;   // causal offset
  int kv_end = kvlen;
  if (p.causal) { int last = q0 + FA_BM - 1 + coff; if (last + 1 < kv_end) kv_end = last + 1; }

  for (int kb = 0; kb < kv_end; kb += FA_BN) {
    __syncthreads();   // previous tile fully consumed
    // ---- stage K tile [64 keys][64 d] and V tile transposed [64 d][64 keys], split to hi/lo
    for (int i = threadIdx.x; i < 64 * 16; i += 128) {          // K: key = i/16, d4 = i%16
      const int key = i >> 4, d4 = (i & 15) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kb + key < kvlen) x = *reinterpret_cast<const float4*>(p.K + (krow0 + kb + key) * p.ldk + head * 64 + d4);
      uint32_t h0, l0_, h1, l1_;
      split2(x.x, x.y, h0, l0_); split2(x.z, x.w, h1, l1_);
      *reinterpret_cast<uint2*>(&Kh[key * FA_LD + d4]) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(&Kl[key * FA_LD + d4]) = make_uint2(l0_, l1_);
    }
    for (int i = threadIdx.x; i < 32 * 16; i += 128) {          // V: key pair = i/16, d4 = i%16
      const int kp = (i >> 4) * 2, d4 = (i & 15) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
      if (kb + kp < kvlen) x = *reinterpret_cast<const float4*>(p.V + (krow0 + kb + kp) * p.ldv + head * 64 + d4);
      if (kb + kp + 1 < kvlen) y = *reinterpret_cast<const float4*>(p.V + (krow0 + kb + kp + 1) * p.ldv + head * 64 + d4);
      const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t h, l;
        split2(xs[e], ys[e], h, l);
        *reinterpret_cast<uint32_t*>(&Vh[(d4 + e) * FA_LD + kp]) = h;
        *reinterpret_cast<uint32_t*>(&Vl[(d4 + e) * FA_LD + kp]) = l;
      }
    }
    __syncthreads();

    // ---- S = Q K^T (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int off = (j * 8 + g) * FA_LD + ks * 16 + 2 * t;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&Kh[off]);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&Kh[off + 8]);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&Kl[off]);
        const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&Kl[off + 8]);
        mma_bf16_16816(s[j], ql[ks], bh0, bh1);
        mma_bf16_16816(s[j], qh[ks], bl0, bl1);
        mma_bf16_16816(s[j], qh[ks], bh0, bh1);
      }
    }
    // ---- bias, scale, mask, online softmax
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = (e < 2) ? r0 : r1;
        const int kj = kb + j * 8 + 2 * t + (e & 1);
        float v = s[j][e];
        bool ok = (kj < kvlen) && (qi < qlen) && (!p.causal || kj <= qi + coff);
        if (ok && p.bias) {
          const long brow = (qrow0 + qi) - p.bias_row0;
          const long bcol = p.bias_rel ? (long)(p.bias_center - qi + kj) : (long)kj;
          v += p.bias[(long)head * p.bias_head_stride + brow * p.bias_ld + bcol];
        }
        v = ok ? v * p.scale : -INFINITY;
        s[j][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = (mn0 == -INFINITY) ? 1.f : expf(m0 - mn0);
    const float c1 = (mn1 == -INFINITY) ? 1.f : expf(m1 - mn1);
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = (s[j][0] == -INFINITY) ? 0.f : expf(s[j][0] - mn0);
      s[j][1] = (s[j][1] == -INFINITY) ? 0.f : expf(s[j][1] - mn0);
      s[j][2] = (s[j][2] == -INFINITY) ? 0.f : expf(s[j][2] - mn1);
      s[j][3] = (s[j][3] == -INFINITY) ? 0.f : expf(s[j][3] - mn1);
      rs0 += s[j][0] + s[j][1];
      rs1 += s[j][2] + s[j][3];
      o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
    }
    l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
    m0 = mn0; m1 = mn1;

    // ---- O += P V   (P from the S accumulators, V^T tile as the col-major B operand)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {         // 16 keys per step
      uint32_t ph[4], pl[4];
      split2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
      split2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
      split2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
      split2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int jd = 0; jd < 8; ++jd) {
        const int off = (jd * 8 + g) * FA_LD + kk * 16 + 2 * t;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(&Vh[off]);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(&Vh[off + 8]);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(&Vl[off]);
        const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(&Vl[off + 8]);
        mma_bf16_16816(o[jd], pl, bh0, bh1);
        mma_bf16_16816(o[jd], ph, bl0, bl1);
        mma_bf16_16816(o[jd], ph, bh0, bh1);
      }
    }
  }
  // ---- finalize
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
  for (int jd = 0; jd < 8; ++jd) {
    const int d = jd * 8 + 2 * t;
    if (r0 < qlen)
      *reinterpret_cast<float2*>(p.O + (qrow0 + r0) * p.ldo + head * 64 + d) = make_float2(o[jd][0] * i0, o[jd][1] * i0);
    if (r1 < qlen)
      *reinterpret_cast<float2*>(p.O + (qrow0 + r1) * p.ldo + head * 64 + d) = make_float2(o[jd][2] * i1, o[jd][3] * i1);
  }
}

// SIMT reference: one warp per (seq, head, query); lanes own 2 of the 64 dims.
__global__ void __launch_bounds__(128) attn_simt_kernel(const FlashDev p) {
  const int seq = blockIdx.z, head = blockIdx.y;
  const int qlen = p.q_len[seq], kvlen = p.kv_len[seq];
  const int qi = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (qi >= qlen) return;
  const int lane = threadIdx.x & 31;
  const long qrow0 = p.q_start[seq], krow0 = p.kv_start[seq];
  const float2 q = *reinterpret_cast<const float2*>(p.Q + (qrow0 + qi) * p.ldq + head * 64 + 2 * lane);
  const int coff = kvlen - qlen;
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int kj = 0; kj < kvlen; ++kj) {
    if (p.causal && kj > qi + coff) break;
    const float2 k = *reinterpret_cast<const float2*>(p.K + (krow0 + kj) * p.ldk + head * 64 + 2 * lane);
    float sc = warp_sum(q.x * k.x + q.y * k.y);
    if (p.bias) {
      const long brow = (qrow0 + qi) - p.bias_row0;
      const long bcol = p.bias_rel ? (long)(p.bias_center - qi + kj) : (long)kj;
      sc += p.bias[(long)head * p.bias_head_stride + brow * p.bias_ld + bcol];
    }
    sc *= p.scale;
    const float mn = fmaxf(m, sc);
    const float c = (m == -INFINITY) ? 0.f : expf(m - mn);
    const float e = expf(sc - mn);
    const float2 v = *reinterpret_cast<const float2*>(p.V + (krow0 + kj) * p.ldv + head * 64 + 2 * lane);
    o0 = o0 * c + e * v.x; o1 = o1 * c + e * v.y;
    l = l * c + e; m = mn;
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  *reinterpret_cast<float2*>(p.O + (qrow0 + qi) * p.ldo + head * 64 + 2 * lane) = make_float2(o0 * inv, o1 * inv);
}

void attention(Ctx& ctx, const AttnArgs& a) {
  if (ctx.dry) return;
  FlashDev p;
  p.Q = a.Q; p.K = a.K; p.V = a.V; p.O = a.O; p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv; p.ldo = a.ldo;
  p.q_start = a.q_start; p.q_len = a.q_len; p.kv_start = a.kv_start; p.kv_len = a.kv_len;
  p.scale = a.scale; p.causal = a.causal;
  p.bias = a.bias; p.bias_head_stride = a.bias_head_stride; p.bias_ld = a.bias_ld; p.bias_row0 = a.bias_row0;
  p.bias_center = a.bias_center; p.bias_rel = a.bias_rel;
  ctx.launches++;
  if (ctx.attn_impl == 1) {
    dim3 grid((a.max_q_len + 3) / 4, a.n_heads, a.n_seq);
    attn_simt_kernel<<<grid, 128, 0, ctx.stream>>>(p);
  } else {
    dim3 grid((a.max_q_len + FA_BM - 1) / FA_BM, a.n_heads, a.n_seq);
    if (ctx.timer) ctx.timer->begin(K_FLASH, ctx.stream);
    flash_attn_kernel<<<grid, 128, 0, ctx.stream>>>(p);
    if (ctx.timer) ctx.timer->end(K_FLASH, ctx.stream);
  }
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// tiny generic attention (any head_dim <= 256): one CTA per (query, head)
// ================================================================================================
__global__ void attn_generic_kernel(const float* Q, const float* K, const float* V, float* O, int n_q, int n_kv,
                                    int head_dim, int ldq, int ldk, int ldv, int ldo, float scale) {
  extern __shared__ float sm[];         // scores [n_kv]
  const int qi = blockIdx.x, head = blockIdx.y;
  const float* q = Q + (long)qi * ldq + head * head_dim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < n_kv; j += nw) {
    const float* k = K + (long)j * ldk + head * head_dim;
    float acc = 0.f;
    for (int d = lane; d < head_dim; d += 32) acc += q[d] * k[d];
    acc = warp_sum(acc);
    if (lane == 0) sm[j] = acc * scale;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = 0; j < n_kv; ++j) mx = fmaxf(mx, sm[j]);
  float sum = 0.f;
  for (int j = 0; j < n_kv; ++j) sum += expf(sm[j] - mx);
  for (int d = threadIdx.x; d < head_dim; d += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < n_kv; ++j) acc += expf(sm[j] - mx) * V[(long)j * ldv + head * head_dim + d];
    O[(long)qi * ldo + head * head_dim + d] = acc / sum;
  }
}
void attention_generic(Ctx& ctx, const float* Q, const float* K, const float* V, float* O, int n_q, int n_kv,
                       int n_heads, int head_dim, int ldq, int ldk, int ldv, int ldo, float scale) {
  if (ctx.dry) return;
  ctx.launches++;
  attn_generic_kernel<<<dim3(n_q, n_heads), 256, n_kv * sizeof(float), ctx.stream>>>(Q, K, V, O, n_q, n_kv, head_dim,
                                                                                    ldq, ldk, ldv, ldo, scale);
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// paged KV cache: layout [layer][page][k|v][head][token][64]
// ================================================================================================
template <typename T> struct KvTraits;
// lanes per token, dims per lane.  PAGED_WIDE: 32 bytes (two 16-byte loads) of K and of V per lane and trip, half the
// per-byte instruction count of the 16-byte version (softmax bookkeeping, shuffles and address arithmetic amortised).
#ifndef PAGED_WIDE
#define PAGED_WIDE 1
#endif
#if PAGED_WIDE
template <> struct KvTraits<__nv_bfloat16> { static constexpr int LPT = 4, DPL = 16; };
template <> struct KvTraits<float> { static constexpr int LPT = 8, DPL = 8; };
#else
template <> struct KvTraits<__nv_bfloat16> { static constexpr int LPT = 8, DPL = 8; };
template <> struct KvTraits<float> { static constexpr int LPT = 16, DPL = 4; };
#endif

#ifndef PAGED_PF
#define PAGED_PF 1
#endif
// one 16-byte piece of a K/V row -> fp32
template <typename T> struct KvPiece;
template <> struct KvPiece<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void decode(const uint4& u, float* x) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[2 * i] = __uint_as_float(w[i] << 16); x[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
};
template <> struct KvPiece<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void decode(const uint4& u, float* x) {
    x[0] = __uint_as_float(u.x); x[1] = __uint_as_float(u.y); x[2] = __uint_as_float(u.z); x[3] = __uint_as_float(u.w);
  }
};

template <> struct KvPiece<__nv_fp8_e4m3> {     // opt-in fp8 KV cache: 16 values per 16-byte piece
  static constexpr int N = 16;
  static __device__ __forceinline__ void decode(const uint4& u, float* x) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] & 0xFFFFu), __NV_E4M3);
      const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(w[i] >> 16), __NV_E4M3);
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
      x[4 * i] = a.x; x[4 * i + 1] = a.y; x[4 * i + 2] = b.x; x[4 * i + 3] = b.y;
    }
  }
};
// value as it will read back from a cache of element type T
template <typename T> __device__ __forceinline__ float kv_round(float v) { return (float)(T)v; }
template <> __device__ __forceinline__ float kv_round<float>(float v) { return v; }

struct PagedDev {
  const float* qkv; int ldqkv;        // [slots][3*H*64] (q rotated in place by rope_store)
  const void* pages; int n_pages, n_layers, n_heads, page_tokens, page_shift, layer;   // page_tokens = 1 << page_shift
  const int* page_table; int max_pages;
  const int* slot_row;                // compact slot -> physical row
  const int* positions;               // [rows] index of the current token (attend to 0..pos)
  float* out; int ldo;                // [slots][H*64]
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;   // when set: the output as bf16 hi/lo planes (operand of the o GEMM)
  float* scratch;                     // [slots][H][nsplit][66] partial (m, l, o[64])
  int nsplit; float scale;
  // bulk-copy kernel only: fuse_rope = 1 -> qkv holds the raw projections of the step's token; the kernel rotates q and k
  // (llama3 RoPE tables cos_t / sin_t [pos][32]), appends k / v to the cache and attends to them from registers
  int fuse_rope; const float* cos_t; const float* sin_t;
  const int* n_live;                  // optional device scalar: slots >= *n_live are retired (CTA exits)
  __half* out16;                      // when set: the output as one fp16 plane
  int dbg_early_release;              // A/B switch (CBX_PB_EARLY=1): release a stage before its loads are known to have returned
};

template <typename T>
__global__ void __launch_bounds__(128) paged_decode_kernel(const PagedDev p) {
  constexpr int LPT = KvTraits<T>::LPT, DPL = KvTraits<T>::DPL, TPI = 32 / LPT;   // tokens per warp-iteration
  constexpr int NP = DPL / KvPiece<T>::N;                                      // 16-byte pieces per lane and row
  const int slot = blockIdx.x, head = blockIdx.y, split = blockIdx.z;
  if (p.n_live && slot >= *p.n_live) return;
  const int row = p.slot_row[slot];
  const int S = p.positions[row] + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPT, grp = lane / LPT;
  const int H = p.n_heads;
  float q[DPL];
  {
    const float* qp = p.qkv + (long)slot * p.ldqkv + head * 64 + sub * DPL;
#pragma unroll
    for (int i = 0; i < DPL; ++i) q[i] = qp[i] * p.scale;
  }
  float m = -INFINITY, l = 0.f, o[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) o[i] = 0.f;
  const long slab = (long)p.page_tokens * 64;                         // elements per (page,layer,kv,head)
  const long page_stride = 2L * H * slab;
  const T* base = reinterpret_cast<const T*>(p.pages) + (long)p.layer * p.n_pages * page_stride + (long)head * slab;
  const int* pt = p.page_table + (long)row * p.max_pages;
  const int n_iter = (S + TPI - 1) / TPI;
  const int stride = 4 * p.nsplit;
  // Software-pipelined stream: the raw 16-byte K and V chunks of the next PF warp-iterations are already in flight
  // while the current one is folded into the online softmax (packed bf16 stays packed until use, 8 registers per
  // stage).  Round-1 history at the bench shape: no prefetch 3.25 TB/s (49% of the copy peak, 9 CTAs/SM x 1 KB per warp
  // in flight); 4 converted token groups per trip needed 107 registers and fell to 34%.
  constexpr int PF = PAGED_PF;
  struct Raw { uint4 k[NP], v[NP]; };
  const int pmask = p.page_tokens - 1;
  auto issue = [&](int it, Raw& r) {
    const int tok = it * TPI + grp;
    if (it < n_iter && tok < S) {
      const int page = pt[tok >> p.page_shift];
      const uint4* kp = reinterpret_cast<const uint4*>(base + (long)page * page_stride + (long)(tok & pmask) * 64 + sub * DPL);
      const uint4* vp = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(kp) + (long)H * slab);
#pragma unroll
      for (int j = 0; j < NP; ++j) { r.k[j] = __ldg(kp + j); r.v[j] = __ldg(vp + j); }
    } else {
#pragma unroll
      for (int j = 0; j < NP; ++j) { r.k[j] = make_uint4(0, 0, 0, 0); r.v[j] = r.k[j]; }
    }
  };
  Raw rq[PF + 1];
  const int it0 = split * 4 + warp;
#pragma unroll
  for (int f = 0; f < PF; ++f) issue(it0 + f * stride, rq[f]);
  for (int it = it0; it < n_iter; it += stride) {
    issue(it + PF * stride, rq[PF]);
    const bool ok = it * TPI + grp < S;
    float kx[DPL], vx[DPL];
#pragma unroll
    for (int j = 0; j < NP; ++j) KvPiece<T>::decode(rq[0].k[j], kx + j * KvPiece<T>::N);
    float sc = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) sc = fmaf(q[i], kx[i], sc);
#pragma unroll
    for (int ofs = 1; ofs < LPT; ofs <<= 1) sc += __shfl_xor_sync(0xffffffffu, sc, ofs);
    if (ok) {
      const float mn = fmaxf(m, sc);
      const float c = (m == -INFINITY) ? 0.f : expf(m - mn);
      const float e = expf(sc - mn);
#pragma unroll
      for (int j = 0; j < NP; ++j) KvPiece<T>::decode(rq[0].v[j], vx + j * KvPiece<T>::N);
#pragma unroll
      for (int i = 0; i < DPL; ++i) o[i] = o[i] * c + e * vx[i];
      l = l * c + e; m = mn;
    }
#pragma unroll
    for (int f = 0; f < PF; ++f) rq[f] = rq[f + 1];
  }
  // ---- merge the 4*TPI independent streams of this CTA
  __shared__ float sm_m[4 * 32], sm_l[4 * 32], sm_o[4 * 32 * DPL];
  sm_m[threadIdx.x] = m; sm_l[threadIdx.x] = l;
#pragma unroll
  for (int i = 0; i < DPL; ++i) sm_o[threadIdx.x * DPL + i] = o[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int d = threadIdx.x;
    const int osub = d / DPL, oi = d % DPL;
    float mt = -INFINITY;
    for (int w = 0; w < 4; ++w)
      for (int gq = 0; gq < TPI; ++gq) mt = fmaxf(mt, sm_m[w * 32 + gq * LPT + osub]);
    float lt = 0.f, ot = 0.f;
    for (int w = 0; w < 4; ++w)
      for (int gq = 0; gq < TPI; ++gq) {
        const int th = w * 32 + gq * LPT + osub;
        const float ms = sm_m[th];
        if (ms == -INFINITY) continue;
        const float c = expf(ms - mt);
        lt += sm_l[th] * c;
        ot += sm_o[th * DPL + oi] * c;
      }
    if (p.nsplit == 1) {
      const float ov = lt > 0.f ? ot / lt : 0.f;
      const long oidx = (long)slot * p.ldo + head * 64 + d;
      if (p.out_hi) { __nv_bfloat16 hh, ll; split_bf16(ov, hh, ll); p.out_hi[oidx] = hh; p.out_lo[oidx] = ll; }
      else p.out[oidx] = ov;
    } else {
      float* sp = p.scratch + (((long)slot * H + head) * p.nsplit + split) * 66;
      sp[2 + d] = ot;
      if (d == 0) { sp[0] = mt; sp[1] = lt; }
    }
  }
}


// ================================================================================================
// paged_bulk_kernel: the decode-step attention of the fused T3 layer.
//   One CTA per (slot, head, split).  A producer warp streams the row's pages with the bulk-copy engine
//   (cp.async.bulk global -> shared, mbarrier complete_tx): with the layer-major pool one (page, head) slab of K is
//   32 tokens x 64 dims = 4 KB (bf16) of contiguous HBM, V likewise, so a stage is two bulk copies and no thread ever
//   touches a global K/V address.  Four consumer warps fold the staged tiles into an online softmax (4 lanes per
//   token, shuffle-reduced scores, exp in fp32) -- 8 stages x 8 KB in flight per CTA, 3 CTAs per SM.
//   fuse_rope: RoPE on q and on the step's k, the KV-cache append and the attention over the new token all happen
//   here (the token's k / v never make a round trip through the cache before they are used).
// Algorithmic HBM bytes: every K/V byte of rows [0, pos] once + 3 x 64 floats of q/k/v per (slot, head).
// ================================================================================================
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// stages per CTA: OCC = 3 -> 64 KB of K/V in flight per CTA (8 bf16 stages), OCC = 4 -> 48 KB (6 bf16 stages): the same ~200 KB per SM,
// spread over 12 or 16 consumer warps
template <typename T, int OCC> struct BulkCfg;
template <int OCC> struct BulkCfg<__nv_bfloat16, OCC> { static constexpr int NST = OCC == 4 ? 6 : 8; };
template <int OCC> struct BulkCfg<float, OCC> { static constexpr int NST = OCC == 4 ? 3 : 4; };
template <int OCC> struct BulkCfg<__nv_fp8_e4m3, OCC> { static constexpr int NST = OCC == 4 ? 12 : 16; };
constexpr int PB_TOK = 32;                    // tokens per page (required by this kernel)
constexpr int PB_THREADS = 160;               // 4 consumer warps + 1 producer warp
constexpr int PB_OCC = 3;                     // persistent CTAs per SM (default; CBX_PB_OCC=4 selects the 4-CTA instantiation)

// shared memory: [NST stages of (K slab | V slab)] [512 B barriers] [merge buffer 1152 B] [2 x (q | k_new | v_new) 768 B]
template <typename T, int OCC = PB_OCC> constexpr int pb_smem_bytes() { return BulkCfg<T, OCC>::NST * 2 * PB_TOK * 64 * (int)sizeof(T) + 512 + 4 * 4 * 18 * 4 + 2 * 192 * 4; }

// Persistent: gridDim.x CTAs walk the (slot, head, split) work items with a fixed stride.  The producer warp runs ahead
// of the consumers ACROSS items (the stage ring and its mbarrier phases never reset), so while the consumers merge one
// row's partial results, the next row's first pages are already landing in shared memory -- and so is its QUERY: the
// producer warp's 32 lanes fetch the next item's q / k / v (and the RoPE table row) while the current item's pages stream,
// rotate them, append the new k / v to the cache and leave q (scaled), k_new, v_new in a double-buffered shared-memory
// slot.  The consumers used to do this themselves at the top of every item behind three dependent global loads
// (slot -> row -> position -> q, ~1.5 us) -- a fifth of an average item.
template <typename T, int OCC = PB_OCC>
__global__ void __launch_bounds__(PB_THREADS, OCC) paged_bulk_kernel(const PagedDev p, const int n_items) {
  constexpr int NST = BulkCfg<T, OCC>::NST;
  constexpr int SLAB = PB_TOK * 64 * (int)sizeof(T);      // one (page, head) slab of K or of V
  constexpr int STAGE = 2 * SLAB;
  constexpr int DPL = 16, NPC = DPL / KvPiece<T>::N;      // dims per lane, 16-byte pieces per lane
  extern __shared__ __align__(128) uint8_t pb_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(pb_smem + NST * STAGE);
  uint64_t* empty = full + NST;
  uint64_t* q_full = empty + NST;                         // [2] query slot staged (1 arrival: the producer warp)
  uint64_t* q_empty = q_full + 2;                         // [2] query slot consumed (4 arrivals: the consumer warps)
  float* mrg = reinterpret_cast<float*>(pb_smem + NST * STAGE + 512);     // [4 warps][4 subs][18]: m, l, o[16]
  float* qs = mrg + 4 * 4 * 18;                           // [2][q 64 | k_new 64 | v_new 64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.n_heads;
  const int per_slot = H * p.nsplit;
  pdl_wait();                                             // (decode step) the QKV projection has completed
  pdl_launch_dependents();
  const int n_live = p.n_live ? *p.n_live : 0x7fffffff;
  const long slab_e = (long)PB_TOK * 64;                  // elements per slab
  const long page_stride = 2L * H * slab_e;
  const T* layer_base = reinterpret_cast<const T*>(p.pages) + (long)p.layer * p.n_pages * page_stride;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 4); }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == 4) {
    // ===================== producer warp ==========================================================================
    // All 32 lanes fetch metadata (the page ids of the next 32 pages in one coalesced load, the next item's row / position /
    // query while the current item streams); lane 0 alone talks to the copy engine.
    struct QRaw { float qa, qb, ka, kb, va, vb, c, s; };   // dims lane and lane + 32 of q, k, v; cos / sin of dim lane
    auto q_fetch = [&](int slot_, int head_, int pos_, QRaw& r) {
      const float* qp = p.qkv + (long)slot_ * p.ldqkv + head_ * 64;
      r.qa = qp[lane]; r.qb = qp[lane + 32];
      r.ka = r.kb = r.va = r.vb = 0.f; r.c = 1.f; r.s = 0.f;
      if (p.fuse_rope) {
        const float* kq = qp + H * 64; const float* vq = qp + 2 * H * 64;
        r.ka = kq[lane]; r.kb = kq[lane + 32]; r.va = vq[lane]; r.vb = vq[lane + 32];
        r.c = p.cos_t[(long)pos_ * 32 + lane]; r.s = p.sin_t[(long)pos_ * 32 + lane];
      }
    };
    auto q_stage = [&](int li, const QRaw& r, int row_, int head_, int split_, int pos_) {
      // rotate_half convention (modeling_llama.py:138-167): out = x*cos + rotate_half(x)*sin, products rounded separately like
      // the reference's elementwise ops (no fused multiply-add)
      float q0 = r.qa, q1 = r.qb, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
      if (p.fuse_rope) {
        q0 = __fadd_rn(__fmul_rn(r.qa, r.c), __fmul_rn(-r.qb, r.s));
        q1 = __fadd_rn(__fmul_rn(r.qb, r.c), __fmul_rn(r.qa, r.s));
        k0 = kv_round<T>(__fadd_rn(__fmul_rn(r.ka, r.c), __fmul_rn(-r.kb, r.s)));
        k1 = kv_round<T>(__fadd_rn(__fmul_rn(r.kb, r.c), __fmul_rn(r.ka, r.s)));
        v0 = kv_round<T>(r.va); v1 = kv_round<T>(r.vb);
      }
      const int b = li & 1;
      mbar_wait(&q_empty[b], ((li >> 1) & 1) ^ 1);        // the consumers are done with the item that used this slot
      float* d = qs + b * 192;
      d[lane] = q0 * p.scale; d[lane + 32] = q1 * p.scale;
      d[64 + lane] = k0; d[96 + lane] = k1; d[128 + lane] = v0; d[160 + lane] = v1;
      // the split that owns the new token's page appends it to the cache
      const int new_pg = pos_ >> 5, new_t = pos_ & 31;
      if (p.fuse_rope && (new_pg % p.nsplit) == split_) {
        const int* pt_ = p.page_table + (long)row_ * p.max_pages;
        T* kd = const_cast<T*>(layer_base) + (long)head_ * slab_e + (long)pt_[new_pg] * page_stride + (long)new_t * 64;
        T* vd = kd + (long)H * slab_e;
        kd[lane] = (T)k0; kd[lane + 32] = (T)k1; vd[lane] = (T)v0; vd[lane + 32] = (T)v1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_full[b]);
    };
    uint32_t gi = 0;
    int li = 0;
    int it = blockIdx.x;
    int slot = it < n_items ? it / per_slot : n_live;
    int row = 0, pos = 0;
    if (slot < n_live) {
      row = p.slot_row[slot]; pos = p.positions[row];
      const int rem0 = it - slot * per_slot;
      QRaw r0;
      q_fetch(slot, rem0 / p.nsplit, pos, r0);
      q_stage(0, r0, row, rem0 / p.nsplit, rem0 % p.nsplit, pos);
    }
    while (it < n_items && slot < n_live) {
      const int rem = it - slot * per_slot;
      const int head = rem / p.nsplit, split = rem - head * p.nsplit;
      const int npg = (pos + PB_TOK) / PB_TOK;
      const int* pt = p.page_table + (long)row * p.max_pages;
      const T* base = layer_base + (long)head * slab_e;
      // next item's metadata and query: issued now, needed after this item's pages
      const int it_n = it + gridDim.x;
      const int slot_n = it_n < n_items ? it_n / per_slot : n_live;
      int row_n = 0, pos_n = 0, head_n = 0, split_n = 0;
      QRaw rn;
      if (slot_n < n_live) {
        row_n = p.slot_row[slot_n]; pos_n = p.positions[row_n];
        const int rem_n = it_n - slot_n * per_slot;
        head_n = rem_n / p.nsplit; split_n = rem_n - head_n * p.nsplit;
        q_fetch(slot_n, head_n, pos_n, rn);
      }
      int i = 0;
      for (int pj0 = split; pj0 < npg; pj0 += 32 * p.nsplit) {
        const int pj_l = pj0 + lane * p.nsplit;
        const int pg_l = pj_l < npg ? pt[pj_l] : 0;
        const int cnt = min(32, (npg - pj0 + p.nsplit - 1) / p.nsplit);
        for (int u = 0; u < cnt; ++u, ++i, ++gi) {
          const int page = __shfl_sync(0xffffffffu, pg_l, u);
          if (lane == 0) {
            const int s = gi % NST;
            mbar_wait(&empty[s], ((gi / NST) & 1) ^ 1);
            const T* kp = base + (long)page * page_stride;
            uint8_t* st = pb_smem + s * STAGE;
            mbar_arrive_expect_tx(&full[s], STAGE);
            bulk_g2s(st, kp, SLAB, &full[s]);
            bulk_g2s(st + SLAB, kp + (long)H * slab_e, SLAB, &full[s]);
          }
        }
      }
      __syncwarp();
      ++li;
      if (slot_n < n_live) q_stage(li, rn, row_n, head_n, split_n, pos_n);
      it = it_n; slot = slot_n; row = row_n; pos = pos_n;
    }
    return;
  }
  // ===================== consumers ========================================================================
  const int sub = lane & 3, grp = lane >> 2;              // 4 lanes per token, 8 tokens per warp and page
  const int d0 = sub * DPL;
  uint32_t gi = 0;
  int li = 0;
  for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
    const int slot = it / per_slot;
    if (slot >= n_live) break;
    const int rem = it - slot * per_slot;
    const int head = rem / p.nsplit, split = rem - head * p.nsplit;
    const int row = p.slot_row[slot];
    const int pos = p.positions[row];
    const int S = pos + 1;
    const int npg = (S + PB_TOK - 1) / PB_TOK;
    const float* qsl = qs + (li & 1) * 192;
    mbar_wait(&q_full[li & 1], (li >> 1) & 1);            // q (rotated, scaled), k_new, v_new of this item are staged
    float q[DPL];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const float4 t = *reinterpret_cast<const float4*>(qsl + d0 + 4 * i4);
      q[4 * i4] = t.x; q[4 * i4 + 1] = t.y; q[4 * i4 + 2] = t.z; q[4 * i4 + 3] = t.w;
    }
    float m = -INFINITY, l = 0.f, o[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[i] = 0.f;
    for (int pj = split; pj < npg; pj += p.nsplit, ++gi) {
      const int s = gi % NST;
      mbar_wait(&full[s], (gi / NST) & 1);
      const int tl = warp * 8 + grp;                      // token inside the page
      const int tok = pj * PB_TOK + tl;
      const uint4* kp = reinterpret_cast<const uint4*>(pb_smem + s * STAGE + (tl * 64 + d0) * (int)sizeof(T));
      const uint4* vp = reinterpret_cast<const uint4*>(pb_smem + s * STAGE + SLAB + (tl * 64 + d0) * (int)sizeof(T));
      float kx[DPL], vx[DPL];
      uint4 kr[NPC], vr[NPC];
#pragma unroll
      for (int j = 0; j < NPC; ++j) { kr[j] = kp[j]; vr[j] = vp[j]; }
      // Release the stage only once the shared-memory loads have RETURNED: the arrive takes a value derived from every
      // load as an (unused) operand, so it cannot issue while one of them is still in flight; and order the generic-proxy
      // reads against the async-proxy write of the producer's next bulk copy into the same bytes (A/B: CBX_PB_NOFENCE=1).
      uint32_t dep = 0;
#pragma unroll
      for (int j = 0; j < NPC; ++j) dep ^= kr[j].x ^ vr[j].x;
      if (!(p.dbg_early_release & 2)) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (p.dbg_early_release & 1) mbar_arrive(&empty[s]);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])), "r"(dep) : "memory");
      }
#pragma unroll
      for (int j = 0; j < NPC; ++j) { KvPiece<T>::decode(kr[j], kx + j * KvPiece<T>::N); KvPiece<T>::decode(vr[j], vx + j * KvPiece<T>::N); }
      if (p.fuse_rope && tok == pos) {                    // the step's own token: taken from the staged slot, its cache entry is being written now
#pragma unroll
        for (int j = 0; j < DPL; ++j) { kx[j] = qsl[64 + d0 + j]; vx[j] = qsl[128 + d0 + j]; }
      }
      float sc = 0.f;
      if (tok < S) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};                 // four independent chains instead of one of 16 dependent FMAs
#pragma unroll
        for (int j = 0; j < DPL; ++j) s4[j & 3] = fmaf(q[j], kx[j], s4[j & 3]);
        sc = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      if (tok < S) {
        const float mn = fmaxf(m, sc);
        const float c = (m == -INFINITY) ? 0.f : expf(m - mn);
        const float e = expf(sc - mn);
#pragma unroll
        for (int j = 0; j < DPL; ++j) o[j] = o[j] * c + e * vx[j];
        l = l * c + e; m = mn;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&q_empty[li & 1]);         // this warp no longer reads the item's query slot
    // ---- merge the 32 token streams of this CTA: 8 lane groups per warp by shuffles, then the 4 warps through smem
#pragma unroll
    for (int ofs = 4; ofs < 32; ofs <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, ofs);
      const float l2 = __shfl_xor_sync(0xffffffffu, l, ofs);
      const float mn = fmaxf(m, m2);
      const float c1 = (m == -INFINITY) ? 0.f : expf(m - mn);
      const float c2 = (m2 == -INFINITY) ? 0.f : expf(m2 - mn);
      l = l * c1 + l2 * c2;
#pragma unroll
      for (int j = 0; j < DPL; ++j) o[j] = o[j] * c1 + __shfl_xor_sync(0xffffffffu, o[j], ofs) * c2;
      m = mn;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");       // the previous item's merge buffer has been read
    if (grp == 0) {
      float* w = mrg + (warp * 4 + sub) * 18;
      w[0] = m; w[1] = l;
#pragma unroll
      for (int j = 0; j < DPL; ++j) w[2 + j] = o[j];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x < 64) {
      const int d = threadIdx.x;
      const int osub = d / DPL, oi = d % DPL;
      float mt = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) mt = fmaxf(mt, mrg[(w * 4 + osub) * 18]);
      float lt = 0.f, ot = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float* r = mrg + (w * 4 + osub) * 18;
        if (r[0] == -INFINITY) continue;
        const float c = expf(r[0] - mt);
        lt += r[1] * c;
        ot += r[2 + oi] * c;
      }
      if (p.nsplit == 1) {
        const float ov = lt > 0.f ? ot / lt : 0.f;
        const long oidx = (long)slot * p.ldo + head * 64 + d;
        if (p.out16) p.out16[oidx] = __float2half_rn(ov);
        else if (p.out_hi) { __nv_bfloat16 hh, ll; split_bf16(ov, hh, ll); p.out_hi[oidx] = hh; p.out_lo[oidx] = ll; }
        else p.out[oidx] = ov;
      } else {
        float* sp = p.scratch + (((long)slot * H + head) * p.nsplit + split) * 66;
        sp[2 + d] = ot;
        if (d == 0) { sp[0] = mt; sp[1] = lt; }
      }
    }
  }
}

__global__ void paged_combine_kernel(const float* scratch, float* out, int ldo, int H, int nsplit,
                                     __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, const int* n_live, __half* out16) {
  const int slot = blockIdx.x, head = blockIdx.y, d = threadIdx.x;
  pdl_wait();
  pdl_launch_dependents();
  if (n_live && slot >= *n_live) return;
  const float* sp = scratch + ((long)slot * H + head) * nsplit * 66;
  float mt = -INFINITY;
  for (int s = 0; s < nsplit; ++s) mt = fmaxf(mt, sp[s * 66]);
  float lt = 0.f, ot = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = sp[s * 66];
    if (ms == -INFINITY) continue;
    const float c = expf(ms - mt);
    lt += sp[s * 66 + 1] * c;
    ot += sp[s * 66 + 2 + d] * c;
  }
  const float ov = lt > 0.f ? ot / lt : 0.f;
  const long oi = (long)slot * ldo + head * 64 + d;
  if (out16) out16[oi] = __float2half_rn(ov);
  else if (out_hi) { __nv_bfloat16 hh, ll; split_bf16(ov, hh, ll); out_hi[oi] = hh; out_lo[oi] = ll; }
  else out[oi] = ov;
}

void paged_attention_init() {     // per device, before any stream capture
  CBX_CHECK(cudaFuncSetAttribute(paged_bulk_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb_smem_bytes<__nv_bfloat16>()));
  CBX_CHECK(cudaFuncSetAttribute(paged_bulk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb_smem_bytes<float>()));
  CBX_CHECK(cudaFuncSetAttribute(paged_bulk_kernel<__nv_fp8_e4m3>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb_smem_bytes<__nv_fp8_e4m3>()));
  CBX_CHECK(cudaFuncSetAttribute(paged_bulk_kernel<__nv_bfloat16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb_smem_bytes<__nv_bfloat16, 4>()));
}

void paged_decode_attention(Ctx& ctx, const float* qkv, int ldqkv, const PagedKV& kv, int layer, const int* slot_row,
                            int n_slots, const int* positions, float* out, int ldo, float* scratch, int nsplit,
                            __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, const PagedOpts* opts) {
  if (ctx.dry) return;
  PagedDev p;
  p.out_hi = out_hi; p.out_lo = out_lo;
  p.qkv = qkv; p.ldqkv = ldqkv; p.pages = kv.pages; p.n_pages = kv.n_pages; p.n_layers = kv.n_layers; p.n_heads = kv.n_heads;
  p.page_tokens = kv.page_tokens; p.layer = layer;
  p.page_shift = 0;
  while ((1 << p.page_shift) < kv.page_tokens) ++p.page_shift;
  CBX_REQUIRE((1 << p.page_shift) == kv.page_tokens, "page_tokens must be a power of two"); p.page_table = kv.page_table; p.max_pages = kv.max_pages_per_row;
  p.slot_row = slot_row; p.positions = positions; p.out = out; p.ldo = ldo; p.scratch = scratch; p.nsplit = nsplit;
  p.scale = 0.125f;
  p.fuse_rope = opts ? opts->fuse_rope : 0; p.cos_t = opts ? opts->cos_t : nullptr; p.sin_t = opts ? opts->sin_t : nullptr;
  p.n_live = opts ? opts->n_live : nullptr;
  p.out16 = opts ? opts->out16 : nullptr;
  static const bool early = getenv("CBX_PB_EARLY") != nullptr;
  static const bool nofence = getenv("CBX_PB_NOFENCE") != nullptr;
  p.dbg_early_release = (early ? 1 : 0) | (nofence ? 2 : 0);
  CBX_REQUIRE(!p.out16 || (kv.page_tokens == PB_TOK && !(opts && opts->impl == 1)), "fp16 output needs the bulk-copy kernel");
  // default: the bulk-copy (TMA engine) kernel; CBX_PAGED=ldg keeps the round-1 __ldg kernel for A/B runs
  static const bool force_ldg = getenv("CBX_PAGED") && std::string(getenv("CBX_PAGED")) == "ldg";
  const bool bulk = kv.page_tokens == PB_TOK && !(force_ldg && !p.fuse_rope) && !(opts && opts->impl == 1);
  CBX_REQUIRE(bulk || !p.fuse_rope, "fused RoPE + KV append needs the bulk-copy kernel (32-token pages)");
  CBX_REQUIRE(bulk || kv.kv_fp32 != 2, "the fp8 KV cache needs the bulk-copy kernel");
  dim3 grid(n_slots, kv.n_heads, nsplit);
  ctx.launches++;
  if (ctx.timer) ctx.timer->begin(K_PAGED, ctx.stream);
  if (bulk) {
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; CBX_CHECK(cudaGetDevice(&dev)); CBX_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)); }
    const int n_items = n_slots * kv.n_heads * nsplit;
    // bf16 cache: 4 CTAs per SM x 6 stages (4.77 TB/s against 4.49 with 3 x 8 at B = 256, 400 steps; session 20); CBX_PB_OCC=3 for A/B
    static const int occ = getenv("CBX_PB_OCC") ? atoi(getenv("CBX_PB_OCC")) : 4;
    const int g = n_items < PB_OCC * n_sm ? n_items : PB_OCC * n_sm;
    if (kv.kv_fp32 == 0 && occ == 4) {
      const int g4 = n_items < 4 * n_sm ? n_items : 4 * n_sm;
      launch_kernel(ctx, paged_bulk_kernel<__nv_bfloat16, 4>, dim3(g4), dim3(PB_THREADS), (size_t)pb_smem_bytes<__nv_bfloat16, 4>(), p, n_items);
    } else if (kv.kv_fp32 == 1) launch_kernel(ctx, paged_bulk_kernel<float>, dim3(g), dim3(PB_THREADS), (size_t)pb_smem_bytes<float>(), p, n_items);
    else if (kv.kv_fp32 == 2) launch_kernel(ctx, paged_bulk_kernel<__nv_fp8_e4m3>, dim3(g), dim3(PB_THREADS), (size_t)pb_smem_bytes<__nv_fp8_e4m3>(), p, n_items);
    else launch_kernel(ctx, paged_bulk_kernel<__nv_bfloat16>, dim3(g), dim3(PB_THREADS), (size_t)pb_smem_bytes<__nv_bfloat16>(), p, n_items);
  } else {
    if (kv.kv_fp32 == 1) paged_decode_kernel<float><<<grid, 128, 0, ctx.stream>>>(p);
    else paged_decode_kernel<__nv_bfloat16><<<grid, 128, 0, ctx.stream>>>(p);
  }
  if (ctx.timer) ctx.timer->end(K_PAGED, ctx.stream);
  if (nsplit > 1) {
    ctx.launches++;
    launch_kernel(ctx, paged_combine_kernel, dim3(n_slots, kv.n_heads), dim3(64), 0, (const float*)scratch, out, ldo, kv.n_heads, nsplit,
                  out_hi, out_lo, p.n_live, p.out16);
  }
  CBX_CHECK(cudaGetLastError());
}

// ---- RoPE (rotate_half convention) on q,k in place + append k,v to the paged cache -------------------
// token i of the launch: qkv row i, physical cache row tok_row[i], position tok_pos[i]
template <typename T>
__global__ void __launch_bounds__(256) rope_store_kernel(float* qkv, int ldqkv, void* pages, int n_pages, int H,
                                                         int page_tokens, int layer, const int* page_table,
                                                         int max_pages, const int* tok_row, const int* tok_pos,
                                                         int pos_is_per_row, const float* cos_t, const float* sin_t) {
  const int i = blockIdx.x;
  const int row = tok_row[i];
  const int pos = pos_is_per_row ? tok_pos[row] : tok_pos[i];
  float* base = qkv + (long)i * ldqkv;
  const long slab = (long)page_tokens * 64;
  const long page_stride = 2L * H * slab;
  const int page = page_table[(long)row * max_pages + pos / page_tokens];
  T* kbase = reinterpret_cast<T*>(pages) + ((long)layer * n_pages + page) * page_stride + (long)(pos % page_tokens) * 64;
  for (int idx = threadIdx.x; idx < H * 32; idx += blockDim.x) {
    const int head = idx >> 5, j = idx & 31;
    const float c = cos_t[(long)pos * 32 + j], s = sin_t[(long)pos * 32 + j];
    float* q = base + head * 64;
    float* k = base + H * 64 + head * 64;
    const float* v = base + 2 * H * 64 + head * 64;
    const float q1 = q[j], q2 = q[j + 32];
    q[j] = q1 * c - q2 * s; q[j + 32] = q2 * c + q1 * s;
    const float k1 = k[j], k2 = k[j + 32];
    const float kr1 = k1 * c - k2 * s, kr2 = k2 * c + k1 * s;
    k[j] = kr1; k[j + 32] = kr2;
    T* kd = kbase + (long)head * slab;
    T* vd = kd + (long)H * slab;
    kd[j] = (T)kr1; kd[j + 32] = (T)kr2;
    vd[j] = (T)v[j]; vd[j + 32] = (T)v[j + 32];
  }
}

void rope_and_store_kv(Ctx& ctx, float* qkv, int ldqkv, const PagedKV& kv, int layer, const int* tok_row,
                       const int* tok_pos, int pos_is_per_row, int n_tok, const float* cos_t, const float* sin_t) {
  if (ctx.dry || n_tok == 0) return;
  ctx.launches++;
  if (kv.kv_fp32 == 1)
    rope_store_kernel<float><<<n_tok, 256, 0, ctx.stream>>>(qkv, ldqkv, kv.pages, kv.n_pages, kv.n_heads, kv.page_tokens,
                                                           layer, kv.page_table, kv.max_pages_per_row, tok_row, tok_pos,
                                                           pos_is_per_row, cos_t, sin_t);
  else if (kv.kv_fp32 == 2)
    rope_store_kernel<__nv_fp8_e4m3><<<n_tok, 256, 0, ctx.stream>>>(qkv, ldqkv, kv.pages, kv.n_pages, kv.n_heads,
                                                                   kv.page_tokens, layer, kv.page_table,
                                                                   kv.max_pages_per_row, tok_row, tok_pos,
                                                                   pos_is_per_row, cos_t, sin_t);
  else
    rope_store_kernel<__nv_bfloat16><<<n_tok, 256, 0, ctx.stream>>>(qkv, ldqkv, kv.pages, kv.n_pages, kv.n_heads,
                                                                   kv.page_tokens, layer, kv.page_table,
                                                                   kv.max_pages_per_row, tok_row, tok_pos,
                                                                   pos_is_per_row, cos_t, sin_t);
  CBX_CHECK(cudaGetLastError());
}

}  // namespace cbx
