// tcgen05 flash attention for the CFM estimator blocks (non-causal, key-length mask, head_dim 64).
//
// One CTA = 128 queries x 1 head of one sequence; key blocks of 64.  Operands are the bf16 hi/lo planes the QKV
// GEMM epilogue writes ([rows][1536] bf16 each), moved by TMA (SWIZZLE_128B):
//   S = Q K^T      tcgen05.mma M128 N64 K64, 3 terms (hi.hi + hi.lo + lo.hi)  -> TMEM (double buffered)
//   softmax        4 warps, one query row per thread (tcgen05.ld 64 columns), online max / sum in fp32,
//                  P split into bf16 hi/lo and written to swizzled smem as the next A operand
//   PV = P V       tcgen05.mma M128 N64 K64, V consumed in its natural [key][d] layout as an MN-major B operand
//   O accumulate   in registers (O = O * exp(m_old - m_new) + PV), normalised and stored as fp32 or as bf16 planes
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 (2..9 with SPLIT 2) = softmax / correction / epilogue.
#include "ops.h"
#include <cstdlib>

namespace cbx {

constexpr int AT_BM = 128, AT_BN = 64, AT_STAGES = 3;
constexpr int AT_TILE = AT_BN * 128;                 // 64 rows x 128 B = 8 KB (one plane of a 64-row tile)
constexpr int AT_Q_BYTES = 2 * 2 * AT_TILE;          // 128 rows x 2 planes = 32 KB
constexpr int AT_KV_STAGE = 4 * AT_TILE;             // K hi, K lo, V hi, V lo = 32 KB
constexpr int AT_P_BYTES = 2 * 2 * AT_TILE;          // P hi, P lo (128 rows each) = 32 KB
constexpr int AT_SMEM = AT_Q_BYTES + AT_STAGES * AT_KV_STAGE + AT_P_BYTES + 1024 + 256 + 2048;   // + barriers + row exchange
// F16 variant (one fp16 plane per operand, one MMA term): half the tile bytes, twice the K/V stages
// OCC = CTAs per SM of the F16 variant: 1 -> 6 K/V stages (128 KB), 2 -> 4 stages (99 KB per CTA; two CTAs share the SM so one
// CTA's softmax (MUFU / ALU) runs under the other's MMAs, and every scheduler holds two softmax warps)
__host__ __device__ constexpr int at_stages_f16(int occ) { return occ == 2 ? 4 : 6; }
__host__ __device__ constexpr int at_smem_f16(int occ) { return AT_Q_BYTES / 2 + at_stages_f16(occ) * (AT_KV_STAGE / 2) + AT_P_BYTES / 2 + 1024 + 256 + 2048; }

struct AttnTcDev {
  float* O; int ldo;
  __nv_bfloat16* Ohi; __nv_bfloat16* Olo;   // when set, O is written as bf16 hi/lo planes [rows][ldo] (operand of the out GEMM)
  const int* q_start; const int* q_len; const int* kv_start; const int* kv_len;
  float scale_log2e;      // softmax scale * log2(e)
  int q_col, k_col, v_col;   // column offsets of head 0 inside the packed planes
  __half* O16;               // F16 variant only: when set, O is written as one fp16 plane [rows][ldo]
};

__device__ __forceinline__ float fast_exp2(float x) {      // ex2.approx: 2 ulp, -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// fp32 pair -> bf16x2 hi word and lo word (one F2FP each)
__device__ __forceinline__ void split_pair_at(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float fa = __uint_as_float(hi << 16), fb = __uint_as_float(hi & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - fa, b - fb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// fp32 pair -> packed fp16x2 (round to nearest)
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// instruction descriptor for fp16 x fp16 -> f32 (a_format = b_format = 0), K-major A and B
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// MN-major SWIZZLE_128B descriptor (B operand stored [k][n], n contiguous, 64 n = one 128-byte row):
// 8 k-rows per 1024-byte atom, atoms along k are SBO = 1024 B apart (cute/arch/mma_sm100_desc.hpp).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 16;     // leading byte offset: next 64-wide n atom (unused, N = 64)
  d |= (uint64_t)(1024 >> 4) << 32;     // stride byte offset: next group of 8 k rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// SPLIT = softmax threads per query row.  SPLIT 1 (default): 4 softmax warps, one row per thread.  SPLIT 2: 8 softmax warps, the two warps that share a TMEM lane
// quarter each own 32 of the 64 key columns of a block (and 32 of the 64 output columns); the row maximum and the
// final row sum are exchanged through shared memory behind a 64-thread named barrier.  Two warps per scheduler hide
// the MUFU / tcgen05.ld latencies that a single in-order warp exposes (round-1 profile: IPC 0.2 with SPLIT 1).
// F16 = 1: the operands are ONE fp16 plane each (Q, K, V from the QKV GEMM epilogue, P from the softmax) and every
// product is a single MMA term.  The CPU study tools/attn_precision_study.py (CPU fp32 arithmetic, 10-step CFM) puts the mel
// RMS of that format at 2e-5 against fp32 -- 50x inside the 1e-3 bar -- while it cuts the tensor work and the
// shared-memory operand traffic of this kernel by 3x.  F16 = 0 (default until measured on the GPU): bf16 hi/lo planes,
// three terms.
template <int SPLIT, int F16, int OCC>
__global__ void __launch_bounds__(64 + 128 * SPLIT, OCC)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const AttnTcDev p) {
  constexpr int NSW = 4 * SPLIT;            // softmax warps
  constexpr int NC = 64 / SPLIT;            // key columns (and output columns) per softmax thread
  constexpr int NPL = F16 ? 1 : 2;          // planes per operand
  constexpr int STAGES = F16 ? at_stages_f16(OCC) : AT_STAGES;
  constexpr int Q_BYTES = NPL * 2 * AT_TILE, KV_STAGE = NPL * 2 * AT_TILE, P_BYTES = NPL * 2 * AT_TILE;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int qlen = p.q_len[seq], kvlen = p.kv_len[seq];
  const int q0 = blockIdx.x * AT_BM;
  if (q0 >= qlen) return;
  const int qrow0 = p.q_start[seq] + q0, krow0 = p.kv_start[seq];
  const int nblk = (kvlen + AT_BN - 1) / AT_BN;

  extern __shared__ uint8_t smem_raw[];
  // keep the __shared__ address space (LDS/STS instead of generic LD/ST): offset the shared pointer, do not round-trip through an integer
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // [hi: 128 rows][lo: 128 rows]   (F16: one plane)
  uint8_t* sKV = sQ + Q_BYTES;                          // stages of [Khi][Klo][Vhi][Vlo] (F16: [K][V])
  uint8_t* sP = sKV + STAGES * KV_STAGE;                // [hi: 128 rows][lo: 128 rows]   (F16: one plane)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES;    // [STAGES]
  uint64_t* s_full = kv_empty + STAGES;     // [2]
  uint64_t* s_empty = s_full + 2;           // [2]
  uint64_t* p_full = s_empty + 2;           // 1
  uint64_t* pv_full = p_full + 1;           // 1
  uint64_t* pv_empty = pv_full + 1;         // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_empty + 1);
  float* xch = reinterpret_cast<float*>(bars + 32);     // [2 buffers][2 halves][128 rows] row-max / row-sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], NSW); }
    mbar_init(p_full, NSW); mbar_init(pv_full, 1); mbar_init(pv_empty, NSW);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_hi); tma_prefetch_desc(&tm_lo); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tS1 = tmem_base + 64, tPV = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer ===============================================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      const int qc = p.q_col + head * 64;
      tma_load_2d(sQ, &tm_hi, q_full, qc, qrow0);
      tma_load_2d(sQ + AT_TILE, &tm_hi, q_full, qc, qrow0 + 64);
      if (!F16) {
        tma_load_2d(sQ + 2 * AT_TILE, &tm_lo, q_full, qc, qrow0);
        tma_load_2d(sQ + 3 * AT_TILE, &tm_lo, q_full, qc, qrow0 + 64);
      }
      const int kc = p.k_col + head * 64, vc = p.v_col + head * 64;
      for (int j = 0; j < nblk; ++j) {
        const int s = j % STAGES;
        mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
        uint8_t* st = sKV + s * KV_STAGE;
        mbar_arrive_expect_tx(&kv_full[s], KV_STAGE);
        const int r = krow0 + j * AT_BN;
        if (F16) {
          tma_load_2d(st, &tm_hi, &kv_full[s], kc, r);
          tma_load_2d(st + AT_TILE, &tm_hi, &kv_full[s], vc, r);
        } else {
          tma_load_2d(st, &tm_hi, &kv_full[s], kc, r);
          tma_load_2d(st + AT_TILE, &tm_lo, &kv_full[s], kc, r);
          tma_load_2d(st + 2 * AT_TILE, &tm_hi, &kv_full[s], vc, r);
          tma_load_2d(st + 3 * AT_TILE, &tm_lo, &kv_full[s], vc, r);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =================================================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = F16 ? umma_idesc_f16(AT_BM, AT_BN) : umma_idesc_bf16(AT_BM, AT_BN);   // A, B K-major
      constexpr uint32_t idesc_pv = (F16 ? umma_idesc_f16(AT_BM, 64) : umma_idesc_bf16(AT_BM, 64)) | (1u << 16);   // B (= V) MN-major
      const uint32_t q_hi = smem_u32(sQ), q_lo = q_hi + 2 * AT_TILE;
      const uint32_t p_hi = smem_u32(sP), p_lo = p_hi + 2 * AT_TILE;
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(&kv_full[s], (j / STAGES) & 1);
        mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t k_hi = smem_u32(sKV + s * KV_STAGE), k_lo = k_hi + AT_TILE;
        const uint32_t d = (j & 1) ? tS1 : tS0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint64_t dkh = umma_desc_sw128(k_hi + k4 * 32), dkl = umma_desc_sw128(k_lo + k4 * 32);
          const uint64_t dqh = umma_desc_sw128(q_hi + k4 * 32), dql = umma_desc_sw128(q_lo + k4 * 32);
          if (F16) {
            umma_bf16(d, dqh, dkh, idesc_s, k4 != 0 ? 1u : 0u);     // kind::f16 covers fp16 and bf16 operands alike
          } else {
            umma_bf16(d, dql, dkh, idesc_s, k4 != 0 ? 1u : 0u);
            umma_bf16(d, dqh, dkl, idesc_s, 1u);
            umma_bf16(d, dqh, dkh, idesc_s, 1u);
          }
        }
        umma_commit(&s_full[j & 1]);
      };
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);            // S of the next block overlaps the softmax of this one
        const int s = j % STAGES;
        mbar_wait(p_full, j & 1);                    // P_j is in smem
        mbar_wait(pv_empty, (j & 1) ^ 1);            // PV accumulator of block j-1 has been read
        tcgen05_fence_after();
        const uint32_t v_hi = smem_u32(sKV + s * KV_STAGE + NPL * AT_TILE), v_lo = v_hi + AT_TILE;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {             // 16 keys per step: 32 B along P rows, 2048 B along V rows
          const uint64_t dvh = umma_desc_sw128_mn(v_hi + k4 * 2048), dvl = umma_desc_sw128_mn(v_lo + k4 * 2048);
          const uint64_t dph = umma_desc_sw128(p_hi + k4 * 32), dpl = umma_desc_sw128(p_lo + k4 * 32);
          if (F16) {
            umma_bf16(tPV, dph, dvh, idesc_pv, k4 != 0 ? 1u : 0u);
          } else {
            umma_bf16(tPV, dpl, dvh, idesc_pv, k4 != 0 ? 1u : 0u);
            umma_bf16(tPV, dph, dvl, idesc_pv, 1u);
            umma_bf16(tPV, dph, dvh, idesc_pv, 1u);
          }
        }
        umma_commit(pv_full);                        // PV_j ready, P smem free
        umma_commit(&kv_empty[s]);                   // K/V stage free
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (SPLIT threads per query row) =================
    const int quarter = warp & 3;                     // TMEM lane quarter of this warp (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;                 // which NC-column slice of the row (always 0 when SPLIT == 1)
    const int row = quarter * 32 + lane;              // query row inside the tile
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int col0 = half * NC;
    float o[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f, c_prev = 1.f;
    uint8_t* prow_hi = sP + row * 128;
    uint8_t* prow_lo = sP + 2 * AT_TILE + row * 128;
    auto fold_pv = [&](bool release) {                // O = O * c_prev + PV (this thread's NC output columns)
      uint32_t v0[32];
#pragma unroll
      for (int cc = 0; cc < NC; cc += 32) {
        tmem_ld_32x32(tPV + lane_off + col0 + cc, v0);
        tmem_ld_wait();
        if (release && cc + 32 >= NC) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(pv_empty);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) o[cc + i] = o[cc + i] * c_prev + __uint_as_float(v0[i]);
      }
    };
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tcgen05_fence_after();
      uint32_t r[NC];
      const uint32_t ts = ((j & 1) ? tS1 : tS0) + lane_off + col0;
#pragma unroll
      for (int cc = 0; cc < NC; cc += 32) tmem_ld_32x32(ts + cc, *reinterpret_cast<uint32_t(*)[32]>(r + cc));
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[j & 1]);
      // scores in the log2 domain; keys beyond the sequence are masked
      const int kbase = j * AT_BN + col0;
      if (j * AT_BN + AT_BN > kvlen) {             // only the last key block needs the length mask
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (kbase + i >= kvlen) r[i] = 0xff800000u;        // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < NC; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if (SPLIT == 2) {                            // row maximum over both halves (double-buffered slot, one barrier)
        float* slot = xch + (j & 1) * 256;
        slot[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        mx = fmaxf(mx, slot[(half ^ 1) * 128 + row]);
      }
      mx = fmaxf(mx, m);
      // raw scores are unscaled; the (positive) scale commutes with max, so scale once here
      const float mxs = mx * p.scale_log2e;
      const float c = (m == -INFINITY) ? 1.f : fast_exp2(m * p.scale_log2e - mxs);
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const float a = fast_exp2(fmaf(__uint_as_float(r[i]), p.scale_log2e, -mxs));
        sum4[i & 3] += a;
        r[i] = __float_as_uint(a);
      }
      l = l * c + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      m = mx;
      // fold the previous block's PV into O (it was computed relative to the previous max)
      if (j > 0) {
        mbar_wait(pv_full, (j - 1) & 1);
        tcgen05_fence_after();
        fold_pv(true);
      }
      c_prev = c;
      // P_j -> bf16 hi/lo planes in swizzled smem (row-major 128 B rows, 16-byte chunk XOR row%8)
#pragma unroll
      for (int cq = 0; cq < NC / 8; ++cq) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (F16) hi[e] = pack_half2(__uint_as_float(r[cq * 8 + e * 2]), __uint_as_float(r[cq * 8 + e * 2 + 1]));
          else split_pair_at(__uint_as_float(r[cq * 8 + e * 2]), __uint_as_float(r[cq * 8 + e * 2 + 1]), hi[e], lo[e]);
        }
        const int ch = half * (NC / 8) + cq;
        const uint32_t off = ((uint32_t)(ch ^ (row & 7))) << 4;
        *reinterpret_cast<uint4*>(prow_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (!F16) *reinterpret_cast<uint4*>(prow_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // last block's PV
    mbar_wait(pv_full, (nblk - 1) & 1);
    tcgen05_fence_after();
    fold_pv(false);
    if (SPLIT == 2) {                              // row sum over both halves
      float* slot = xch + (nblk & 1) * 256;
      slot[half * 128 + row] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      l += slot[(half ^ 1) * 128 + row];
    }
    if (q0 + row < qlen) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      const long off = (long)(qrow0 + row) * p.ldo + head * 64 + col0;
      if (F16 && p.O16) {              // (the fp16 output plane only exists together with the fp16 operand format)
        uint4* d16 = reinterpret_cast<uint4*>(p.O16 + off);
#pragma unroll
        for (int i = 0; i < NC; i += 8)
          d16[i / 8] = make_uint4(pack_half2(o[i] * inv, o[i + 1] * inv), pack_half2(o[i + 2] * inv, o[i + 3] * inv),
                                  pack_half2(o[i + 4] * inv, o[i + 5] * inv), pack_half2(o[i + 6] * inv, o[i + 7] * inv));
      } else if (p.Ohi) {
        uint4* dh = reinterpret_cast<uint4*>(p.Ohi + off);
        uint4* dl = reinterpret_cast<uint4*>(p.Olo + off);
#pragma unroll
        for (int i = 0; i < NC; i += 8) {
          uint4 h, lw;
          split_pair_at(o[i] * inv, o[i + 1] * inv, h.x, lw.x);
          split_pair_at(o[i + 2] * inv, o[i + 3] * inv, h.y, lw.y);
          split_pair_at(o[i + 4] * inv, o[i + 5] * inv, h.z, lw.z);
          split_pair_at(o[i + 6] * inv, o[i + 7] * inv, h.w, lw.w);
          dh[i / 8] = h; dl[i / 8] = lw;
        }
      } else {
        float* dst = p.O + off;
#pragma unroll
        for (int i = 0; i < NC; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      }
    } else if (F16 && p.O16) {
      uint4* d16 = reinterpret_cast<uint4*>(p.O16 + (long)(qrow0 + row) * p.ldo + head * 64 + col0);
#pragma unroll
      for (int i = 0; i < NC / 8; ++i) d16[i] = make_uint4(0, 0, 0, 0);       // finite padding rows, see below
    } else if (p.Ohi) {
      // Padding rows of the sequence's last 128-row tile (packed layouts start every sequence on a tile boundary): keep
      // them finite.  They flow through the out GEMM into x, come back as K/V padding rows of the next block, and a NaN
      // there would poison real rows through 0 * NaN in P.V.
      const long off = (long)(qrow0 + row) * p.ldo + head * 64 + col0;
      uint4* dh = reinterpret_cast<uint4*>(p.Ohi + off);
      uint4* dl = reinterpret_cast<uint4*>(p.Olo + off);
#pragma unroll
      for (int i = 0; i < NC / 8; ++i) { dh[i] = make_uint4(0, 0, 0, 0); dl[i] = make_uint4(0, 0, 0, 0); }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}


// ================================================================================================
// attn_f16_kernel: the fp16-operand variant, rebuilt around the softmax warps' ISSUE budget.
// ncu of attn_tc_kernel<1,1,2> (profiles/r2_ncu_attn_tc.txt): tensor pipe idle most of the time, 9.7 issued instructions per
// (row, key) element, schedulers 47 % busy with 2.6 warps each -- the softmax is the kernel.  Head dim 64 makes that
// structural: a 128 x 64 key block costs 256 tensor cycles (QK^T + PV, one fp16 term) but 8192 ex2 = 512 MUFU cycles, so the
// exp pipe is the roof and every other softmax instruction has to fit under it:
//   * row max with 3-input FMNMX3 (32 instead of 64 instructions per 64 keys)
//   * scale / subtract, row sum and the O rescale on PACKED fp32 pairs (fma.rn.f32x2 / add.f32x2: FFMA2 / FADD2)
//   * PTM = 1: P never touches shared memory.  The softmax writes its fp16 row straight back into the TMEM columns of the S
//     tile it came from (tcgen05.st) and the PV product takes A from TMEM (.ts form of tcgen05.mma): no STS, no proxy
//     fence, and the PV MMAs read only V (2 KB per K step) from shared memory instead of P + V (6 KB, over the 128 B/clk
//     the SS form can pull for an N = 64 tile).
// Same roles and barriers as attn_tc_kernel; S tiles double buffered, S_{j+2} cannot overtake PV_j because one thread
// issues both in that order and the tensor pipe executes in issue order.
// ================================================================================================
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// accumulator forms: the destination IS the loop-carried register (a plain "=l" output lands in the dying operand's register
// and is then moved back: 64 extra MOVs per key block in the first build)
__device__ __forceinline__ void fma2_acc(uint64_t& acc, uint64_t scale, uint64_t add) { asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc) : "l"(scale), "l"(add)); }
__device__ __forceinline__ void add2_acc(uint64_t& acc, uint64_t b) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(b)); }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
// D[tmem] (+)= A[tmem] * B[smem desc]   (.ts form: the A operand -- 128 rows x 16 fp16, two per 32-bit column -- comes from TMEM)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ constexpr int af_stages(int occ, int ptm) { return occ == 2 ? (ptm ? 5 : 4) : (ptm ? 8 : 6); }
__host__ __device__ constexpr int af_smem(int occ, int ptm) {
  return AT_Q_BYTES / 2 + af_stages(occ, ptm) * (AT_KV_STAGE / 2) + (ptm ? 0 : AT_P_BYTES / 2) + 1024 + 256 + 2048;    // + align, barriers, row exchange
}

template <int OCC, int PTM>
__global__ void __launch_bounds__(192, OCC)
attn_f16_kernel(const __grid_constant__ CUtensorMap tm, const AttnTcDev p) {
  constexpr int STAGES = af_stages(OCC, PTM);
  constexpr int Q_BYTES = 2 * AT_TILE, KV_STAGE = 2 * AT_TILE, P_BYTES = PTM ? 0 : 2 * AT_TILE;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int qlen = p.q_len[seq], kvlen = p.kv_len[seq];
  const int q0 = blockIdx.x * AT_BM;
  if (q0 >= qlen) return;
  const int qrow0 = p.q_start[seq] + q0, krow0 = p.kv_start[seq];
  const int nblk = (kvlen + AT_BN - 1) / AT_BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // 128 rows x 128 B
  uint8_t* sKV = sQ + Q_BYTES;                          // stages of [K 64 rows][V 64 rows]
  uint8_t* sP = sKV + STAGES * KV_STAGE;                // PTM = 0 only: 128 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES;    // [STAGES]
  uint64_t* s_full = kv_empty + STAGES;     // [2]
  uint64_t* s_empty = s_full + 2;           // [2]
  uint64_t* p_full = s_empty + 2;           // 1
  uint64_t* pv_full = p_full + 1;           // 1
  uint64_t* pv_empty = pv_full + 1;         // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); }
    mbar_init(p_full, 4); mbar_init(pv_full, 1); mbar_init(pv_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tS1 = tmem_base + 64, tPV = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer ===============================================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      const int qc = p.q_col + head * 64;
      tma_load_2d(sQ, &tm, q_full, qc, qrow0);
      tma_load_2d(sQ + AT_TILE, &tm, q_full, qc, qrow0 + 64);
      const int kc = p.k_col + head * 64, vc = p.v_col + head * 64;
      for (int j = 0; j < nblk; ++j) {
        const int s = j % STAGES;
        mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
        uint8_t* st = sKV + s * KV_STAGE;
        mbar_arrive_expect_tx(&kv_full[s], KV_STAGE);
        const int r = krow0 + j * AT_BN;
        tma_load_2d(st, &tm, &kv_full[s], kc, r);
        tma_load_2d(st + AT_TILE, &tm, &kv_full[s], vc, r);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =================================================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AT_BM, AT_BN);                   // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BM, 64) | (1u << 16);        // B (= V) MN-major
      const uint32_t q_a = smem_u32(sQ), p_a = smem_u32(sP);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(&kv_full[s], (j / STAGES) & 1);
        mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t k_a = smem_u32(sKV + s * KV_STAGE);
        const uint32_t d = (j & 1) ? tS1 : tS0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_bf16(d, umma_desc_sw128(q_a + k4 * 32), umma_desc_sw128(k_a + k4 * 32), idesc_s, k4 != 0 ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
      };
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);            // S of the next block overlaps the softmax of this one
        const int s = j % STAGES;
        mbar_wait(p_full, j & 1);                    // P_j is in TMEM (or smem)
        mbar_wait(pv_empty, (j & 1) ^ 1);            // PV accumulator of block j-1 has been read
        tcgen05_fence_after();
        const uint32_t v_a = smem_u32(sKV + s * KV_STAGE + AT_TILE);
        const uint32_t tP = (j & 1) ? tS1 : tS0;     // PTM: P_j sits in the first 32 columns of its S tile
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {             // 16 keys per step: 8 TMEM columns (32 B of a P row), 2048 B along V rows
          const uint64_t dv = umma_desc_sw128_mn(v_a + k4 * 2048);
          if (PTM) umma_f16_ts(tPV, tP + k4 * 8, dv, idesc_pv, k4 != 0 ? 1u : 0u);
          else umma_bf16(tPV, umma_desc_sw128(p_a + k4 * 32), dv, idesc_pv, k4 != 0 ? 1u : 0u);
        }
        umma_commit(pv_full);                        // PV_j ready, P free
        umma_commit(&kv_empty[s]);                   // K/V stage free
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: one query row per thread ===================
    const int quarter = warp & 3;                     // TMEM lane quarter of this warp (hardware: warp id % 4)
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint64_t o2[32];                                  // 64 output columns as packed pairs
#pragma unroll
    for (int i = 0; i < 32; ++i) o2[i] = 0ull;
    float m = -INFINITY, l = 0.f, c_prev = 1.f;
    uint8_t* prow = sP + row * 128;
    const uint64_t sc2 = pk2(p.scale_log2e, p.scale_log2e);
    auto fold_pv = [&](bool release) {                // O = O * c_prev + PV
      const uint64_t c2 = pk2(c_prev, c_prev);
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        uint32_t v0[32];
        tmem_ld_32x32(tPV + lane_off + cc, v0);
        tmem_ld_wait();
        if (release && cc == 32) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(pv_empty);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
          fma2_acc(o2[cc / 2 + i], c2, pk2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1])));
      }
    };
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tcgen05_fence_after();
      uint32_t r[64];
      const uint32_t ts = ((j & 1) ? tS1 : tS0) + lane_off;
      tmem_ld_32x32(ts, *reinterpret_cast<uint32_t(*)[32]>(r));
      tmem_ld_32x32(ts + 32, *reinterpret_cast<uint32_t(*)[32]>(r + 32));
      tmem_ld_wait();
      if (!PTM) {                                     // PTM: the S tile doubles as the P buffer, released by the PV commit order
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[j & 1]);
      }
      if (j * AT_BN + AT_BN > kvlen) {                // only the last key block needs the length mask
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (j * AT_BN + i >= kvlen) r[i] = 0xff800000u;        // -inf
      }
      float mxa = m, mxb = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mxa = max3(mxa, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        mxb = max3(mxb, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
      const float mx = fmaxf(mxa, mxb);
      // raw scores are unscaled; the (positive) scale commutes with max, so scale once here
      const float mxs = mx * p.scale_log2e;
      const float c = (m == -INFINITY) ? 1.f : fast_exp2(m * p.scale_log2e - mxs);
      const uint64_t nm2 = pk2(-mxs, -mxs);
      uint64_t sa = 0ull, sb = 0ull;
      uint32_t ph[32];                                // P_j as packed fp16 pairs
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        float x0, x1, x2, x3;
        upk2(fma2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, nm2), x0, x1);
        upk2(fma2(pk2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), sc2, nm2), x2, x3);
        x0 = fast_exp2(x0); x1 = fast_exp2(x1); x2 = fast_exp2(x2); x3 = fast_exp2(x3);
        add2_acc(sa, pk2(x0, x1));
        add2_acc(sb, pk2(x2, x3));
        ph[i / 2] = pack_half2(x0, x1);
        ph[i / 2 + 1] = pack_half2(x2, x3);
      }
      {
        float s0, s1, s2, s3;
        upk2(sa, s0, s1); upk2(sb, s2, s3);
        l = l * c + ((s0 + s1) + (s2 + s3));
      }
      m = mx;
      // hand P_j to the MMA warp first (its registers die here), then fold the previous block's PV into O
      if (PTM) {
        tmem_st_32x32_x32(ts, ph);                    // row `row`, columns 0..31 of the S tile = 64 fp16 keys
        tmem_st_wait();
        tcgen05_fence_before();
      } else {
        if (j > 0) mbar_wait(pv_full, (j - 1) & 1);   // PV_{j-1} has finished reading the P buffer
#pragma unroll
        for (int cq = 0; cq < 8; ++cq)
          *reinterpret_cast<uint4*>(prow + (((uint32_t)(cq ^ (row & 7))) << 4)) = make_uint4(ph[4 * cq], ph[4 * cq + 1], ph[4 * cq + 2], ph[4 * cq + 3]);
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (PTM) { if (lane == 0) mbar_arrive(&s_empty[j & 1]); }   // bookkeeping only: S_{j+2} is issued after PV_j by the same thread
      if (j > 0) {                                    // PV_{j-1} was computed relative to the previous max: O = O * c_prev + PV
        mbar_wait(pv_full, (j - 1) & 1);
        tcgen05_fence_after();
        fold_pv(true);
      }
      c_prev = c;
    }
    // last block's PV
    mbar_wait(pv_full, (nblk - 1) & 1);
    tcgen05_fence_after();
    fold_pv(false);
    float o[64];
#pragma unroll
    for (int i = 0; i < 32; ++i) upk2(o2[i], o[2 * i], o[2 * i + 1]);
    const long off = (long)(qrow0 + row) * p.ldo + head * 64;
    if (q0 + row < qlen) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      if (p.O16) {
        uint4* d16 = reinterpret_cast<uint4*>(p.O16 + off);
#pragma unroll
        for (int i = 0; i < 64; i += 8)
          d16[i / 8] = make_uint4(pack_half2(o[i] * inv, o[i + 1] * inv), pack_half2(o[i + 2] * inv, o[i + 3] * inv),
                                  pack_half2(o[i + 4] * inv, o[i + 5] * inv), pack_half2(o[i + 6] * inv, o[i + 7] * inv));
      } else if (p.Ohi) {
        uint4* dh = reinterpret_cast<uint4*>(p.Ohi + off);
        uint4* dl = reinterpret_cast<uint4*>(p.Olo + off);
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          uint4 h, lw;
          split_pair_at(o[i] * inv, o[i + 1] * inv, h.x, lw.x);
          split_pair_at(o[i + 2] * inv, o[i + 3] * inv, h.y, lw.y);
          split_pair_at(o[i + 4] * inv, o[i + 5] * inv, h.z, lw.z);
          split_pair_at(o[i + 6] * inv, o[i + 7] * inv, h.w, lw.w);
          dh[i / 8] = h; dl[i / 8] = lw;
        }
      } else {
        float* dst = p.O + off;
#pragma unroll
        for (int i = 0; i < 64; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      }
    } else if (p.O16) {          // padding rows of the sequence's last tile stay finite (see attn_tc_kernel)
      uint4* d16 = reinterpret_cast<uint4*>(p.O16 + off);
#pragma unroll
      for (int i = 0; i < 8; ++i) d16[i] = make_uint4(0, 0, 0, 0);
    } else if (p.Ohi) {
      uint4* dh = reinterpret_cast<uint4*>(p.Ohi + off);
      uint4* dl = reinterpret_cast<uint4*>(p.Olo + off);
#pragma unroll
      for (int i = 0; i < 8; ++i) { dh[i] = make_uint4(0, 0, 0, 0); dl[i] = make_uint4(0, 0, 0, 0); }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}


// ================================================================================================
// attn_otm_kernel: O stays in TMEM.  Measured on the B200 (session 11): attn_f16_kernel cut the softmax instruction count by a
// third and gained 11 % -- the kernel is not issue-bound but TMEM-READ bound.  tcgen05.ld moves 64 B per clock per SM
// (B300_MICROARCH.md, TMEM table); a 128 x 64 key block made every CTA read the fp32 S tile (32 KB) AND the fp32 PV tile
// (32 KB) back into registers: 1024 clocks per block against 256 clocks of MMA and 512 of MUFU.  Here the PV product
// ACCUMULATES in TMEM across key blocks (enable-input-d after the first block) and the softmax never reads it back, except
//   * once at the end (normalise + store), and
//   * when a row's running maximum grows by more than 2^8 over the reference maximum its probabilities are expressed in
//     ("lazy rescale", as in FlashAttention-4): then the warp multiplies its 32 accumulator rows by 2^(m_ref_old - m_ref_new)
//     in place (tcgen05.ld / tcgen05.st).  P = 2^(s - m_ref) <= 2^8 fits fp16; the row sum is kept in fp32.
// That halves the TMEM reads (512 clocks per block, level with the MUFU floor) and drops 64 accumulator registers.
// ================================================================================================
constexpr float AT_RESCALE_LOG2 = 8.0f;

template <int OCC>
__global__ void __launch_bounds__(192, OCC)
attn_otm_kernel(const __grid_constant__ CUtensorMap tm, const AttnTcDev p) {
  constexpr int STAGES = af_stages(OCC, 1);
  constexpr int Q_BYTES = 2 * AT_TILE, KV_STAGE = 2 * AT_TILE;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int qlen = p.q_len[seq], kvlen = p.kv_len[seq];
  const int q0 = blockIdx.x * AT_BM;
  if (q0 >= qlen) return;
  const int qrow0 = p.q_start[seq] + q0, krow0 = p.kv_start[seq];
  const int nblk = (kvlen + AT_BN - 1) / AT_BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // 128 rows x 128 B
  uint8_t* sKV = sQ + Q_BYTES;                          // stages of [K 64 rows][V 64 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * KV_STAGE);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES;    // [STAGES]
  uint64_t* s_full = kv_empty + STAGES;     // [3]
  uint64_t* p_full = s_full + 3;            // 1   (4 softmax warps: P_j written, O rescaled if it had to be)
  uint64_t* pv_full = p_full + 1;           // 1   (PV_j accumulated)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 3; ++s) mbar_init(&s_full[s], 1);
    mbar_init(p_full, 4); mbar_init(pv_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // three S tiles: S_{j+2} is issued before the softmax of block j has finished (it used to be issued only then, and the
  // softmax warps spent a fifth of their time waiting for the next tile: profiles/r2_ncu_attn_otm_prefetch.txt)
  const uint32_t tO = tmem_base + 192;
  auto tS = [&](int j) { return tmem_base + (uint32_t)((j % 3) * 64); };

  if (warp == 0) {
    // ===================== TMA producer ===============================================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      const int qc = p.q_col + head * 64;
      tma_load_2d(sQ, &tm, q_full, qc, qrow0);
      tma_load_2d(sQ + AT_TILE, &tm, q_full, qc, qrow0 + 64);
      const int kc = p.k_col + head * 64, vc = p.v_col + head * 64;
      for (int j = 0; j < nblk; ++j) {
        const int s = j % STAGES;
        mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
        uint8_t* st = sKV + s * KV_STAGE;
        mbar_arrive_expect_tx(&kv_full[s], KV_STAGE);
        const int r = krow0 + j * AT_BN;
        tma_load_2d(st, &tm, &kv_full[s], kc, r);
        tma_load_2d(st + AT_TILE, &tm, &kv_full[s], vc, r);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =================================================================
    // Issue order S_0, S_1, S_2, PV_0, S_3, PV_1, ...: S_{j+3} overwrites the tile that held S_j / P_j, and it is issued after
    // PV_j by this same thread (the tensor pipe executes in issue order), PV_j in turn after the softmax has read S_j.
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AT_BM, AT_BN);                   // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BM, 64) | (1u << 16);        // B (= V) MN-major
      const uint32_t q_a = smem_u32(sQ);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(&kv_full[s], (j / STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t k_a = smem_u32(sKV + s * KV_STAGE);
        const uint32_t d = tS(j);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_bf16(d, umma_desc_sw128(q_a + k4 * 32), umma_desc_sw128(k_a + k4 * 32), idesc_s, k4 != 0 ? 1u : 0u);
        umma_commit(&s_full[j % 3]);
      };
      issue_s(0);
      if (nblk > 1) issue_s(1);
      for (int j = 0; j < nblk; ++j) {
        if (j + 2 < nblk) issue_s(j + 2);            // two S tiles ahead of the softmax
        const int s = j % STAGES;
        mbar_wait(p_full, j & 1);                    // P_j is in TMEM, O carries the right scale
        tcgen05_fence_after();
        const uint32_t v_a = smem_u32(sKV + s * KV_STAGE + AT_TILE);
        const uint32_t tP = tS(j);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)               // 16 keys per step: 8 TMEM columns of P, 2048 B along V rows
          umma_f16_ts(tO, tP + k4 * 8, umma_desc_sw128_mn(v_a + k4 * 2048), idesc_pv, (j | k4) != 0 ? 1u : 0u);
        umma_commit(pv_full);                        // PV_j accumulated
        umma_commit(&kv_empty[s]);                   // K/V stage free
      }
    }
  } else {
    // ===================== softmax: one query row per thread ============================================
    const int quarter = warp & 3;                     // TMEM lane quarter of this warp (hardware: warp id % 4)
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float m_ref = -INFINITY, l = 0.f;                 // reference maximum of the row's probabilities, row sum relative to it
    const uint64_t sc2 = pk2(p.scale_log2e, p.scale_log2e);
    const float thr = AT_RESCALE_LOG2 / p.scale_log2e;           // raw-score units
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[j % 3], (j / 3) & 1);
      tcgen05_fence_after();
      uint32_t r[64];
      const uint32_t ts = tS(j) + lane_off;
      tmem_ld_32x32(ts, *reinterpret_cast<uint32_t(*)[32]>(r));
      tmem_ld_32x32(ts + 32, *reinterpret_cast<uint32_t(*)[32]>(r + 32));
      tmem_ld_wait();
      if (j * AT_BN + AT_BN > kvlen) {                // only the last key block needs the length mask
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (j * AT_BN + i >= kvlen) r[i] = 0xff800000u;        // -inf
      }
      float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mxa = max3(mxa, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        mxb = max3(mxb, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
      const float mx = fmaxf(mxa, mxb);
      // lazy rescale: move the reference only when this block's maximum exceeds it by more than 2^8
      const bool move = mx > m_ref + thr;             // also true for the first block (m_ref = -inf) and never for mx = -inf
      float c = 1.f;
      if (move) {
        c = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - mx) * p.scale_log2e);
        m_ref = mx;
        l *= c;
      }
      bool synced = false;                            // waited for PV_{j-1} in this iteration
      if (j > 0 && __any_sync(0xffffffffu, move)) {   // this warp's 32 accumulator rows, in place
        mbar_wait(pv_full, (j - 1) & 1);              // PV_{j-1} has landed; PV_j waits for our p_full arrival
        synced = true;
        tcgen05_fence_after();
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          uint32_t ov[32];
          tmem_ld_32x32(tO + lane_off + cc, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * c);
          tmem_st_32x32_x32(tO + lane_off + cc, ov);
        }
      }
      const float mrs = m_ref * p.scale_log2e;
      const uint64_t nm2 = pk2(-mrs, -mrs);
      uint64_t sa = 0ull, sb = 0ull;
      uint32_t ph[32];                                // P_j as packed fp16 pairs
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        float x0, x1, x2, x3;
        upk2(fma2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, nm2), x0, x1);
        upk2(fma2(pk2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), sc2, nm2), x2, x3);
        x0 = fast_exp2(x0); x1 = fast_exp2(x1); x2 = fast_exp2(x2); x3 = fast_exp2(x3);
        add2_acc(sa, pk2(x0, x1));
        add2_acc(sb, pk2(x2, x3));
        ph[i / 2] = pack_half2(x0, x1);
        ph[i / 2 + 1] = pack_half2(x2, x3);
      }
      {
        float s0, s1, s2, s3;
        upk2(sa, s0, s1); upk2(sb, s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      tmem_st_32x32_x32(ts, ph);                      // row `row`, columns 0..31 of the S tile = 64 fp16 keys
      tmem_st_wait();                                 // (covers the accumulator rescale stores as well)
      tcgen05_fence_before();
      // Every warp observes EVERY phase of pv_full (a parity wait cannot tell phase k from phase k+2): PV_{j-1} was issued
      // when the slowest warp finished block j-1 and has long completed by now, so this does not stall -- it only keeps the
      // warp within one phase of the barrier for the rescale wait above and for the final wait below.
      if (j > 0 && !synced) mbar_wait(pv_full, (j - 1) & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: the accumulated O, normalised
    mbar_wait(pv_full, (nblk - 1) & 1);
    tcgen05_fence_after();
    const long off = (long)(qrow0 + row) * p.ldo + head * 64;
    const bool live = q0 + row < qlen;
    const float inv = (live && l > 0.f) ? 1.f / l : 0.f;        // padding rows of the sequence's last tile are written as zeros
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
      uint32_t ov[32];
      tmem_ld_32x32(tO + lane_off + cc, ov);
      tmem_ld_wait();
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = live ? __uint_as_float(ov[i]) * inv : 0.f;
      if (p.O16) {
        uint4* d16 = reinterpret_cast<uint4*>(p.O16 + off + cc);
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          d16[i / 8] = make_uint4(pack_half2(o[i], o[i + 1]), pack_half2(o[i + 2], o[i + 3]), pack_half2(o[i + 4], o[i + 5]), pack_half2(o[i + 6], o[i + 7]));
      } else if (p.Ohi) {
        uint4* dh = reinterpret_cast<uint4*>(p.Ohi + off + cc);
        uint4* dl = reinterpret_cast<uint4*>(p.Olo + off + cc);
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 h, lw;
          split_pair_at(o[i], o[i + 1], h.x, lw.x); split_pair_at(o[i + 2], o[i + 3], h.y, lw.y);
          split_pair_at(o[i + 4], o[i + 5], h.z, lw.z); split_pair_at(o[i + 6], o[i + 7], h.w, lw.w);
          dh[i / 8] = h; dl[i / 8] = lw;
        }
      } else if (live) {
        float* dst = p.O + off + cc;
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// attn_otm2_kernel: the same kernel with TWO softmax threads per query row (8 softmax warps; the two warps that share a TMEM
// lane quarter own 32 of the 64 key columns of a block each, and 32 of the 64 output columns).  ncu of attn_otm_kernel
// (profiles/r2_ncu_attn_otm.txt): the MUFU pipe is 43 % busy and the schedulers idle 68 % of the time on fixed-latency and
// tcgen05.ld dependencies -- two in-order softmax warps per scheduler cannot cover their own latencies.  Four per scheduler
// can, at ~100 registers per thread.  The pair agrees on the row's reference maximum through a double-buffered shared-memory
// slot and one 64-thread named barrier per key block; with the lazy rescale that decision is the only per-block exchange.
template <int OCC>
__global__ void __launch_bounds__(320, OCC)
attn_otm2_kernel(const __grid_constant__ CUtensorMap tm, const AttnTcDev p) {
  constexpr int STAGES = af_stages(OCC, 1);
  constexpr int Q_BYTES = 2 * AT_TILE, KV_STAGE = 2 * AT_TILE;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int qlen = p.q_len[seq], kvlen = p.kv_len[seq];
  const int q0 = blockIdx.x * AT_BM;
  if (q0 >= qlen) return;
  const int qrow0 = p.q_start[seq] + q0, krow0 = p.kv_start[seq];
  const int nblk = (kvlen + AT_BN - 1) / AT_BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // 128 rows x 128 B
  uint8_t* sKV = sQ + Q_BYTES;                          // stages of [K 64 rows][V 64 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * KV_STAGE);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES;    // [STAGES]
  uint64_t* s_full = kv_empty + STAGES;     // [2]
  uint64_t* p_full = s_full + 2;            // 1   (4 softmax warps: P_j written, O rescaled if it had to be)
  uint64_t* pv_full = p_full + 1;           // 1   (PV_j accumulated)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + 1);
  float* xch = reinterpret_cast<float*>(bars + 32);     // [2 buffers][2 halves][128 rows] row-maximum / row-sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) mbar_init(&s_full[s], 1);
    mbar_init(p_full, 8); mbar_init(pv_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tS1 = tmem_base + 64, tO = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer ===============================================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
      const int qc = p.q_col + head * 64;
      tma_load_2d(sQ, &tm, q_full, qc, qrow0);
      tma_load_2d(sQ + AT_TILE, &tm, q_full, qc, qrow0 + 64);
      const int kc = p.k_col + head * 64, vc = p.v_col + head * 64;
      for (int j = 0; j < nblk; ++j) {
        const int s = j % STAGES;
        mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
        uint8_t* st = sKV + s * KV_STAGE;
        mbar_arrive_expect_tx(&kv_full[s], KV_STAGE);
        const int r = krow0 + j * AT_BN;
        tma_load_2d(st, &tm, &kv_full[s], kc, r);
        tma_load_2d(st + AT_TILE, &tm, &kv_full[s], vc, r);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =================================================================
    // Issue order S_0, S_1, PV_0, S_2, PV_1, ...: S_{j+2} overwrites the tile that held S_j / P_j, and it is issued after
    // PV_j by this same thread (the tensor pipe executes in issue order), PV_j in turn after the softmax has read S_j.
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AT_BM, AT_BN);                   // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BM, 64) | (1u << 16);        // B (= V) MN-major
      const uint32_t q_a = smem_u32(sQ);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(&kv_full[s], (j / STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t k_a = smem_u32(sKV + s * KV_STAGE);
        const uint32_t d = (j & 1) ? tS1 : tS0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma_bf16(d, umma_desc_sw128(q_a + k4 * 32), umma_desc_sw128(k_a + k4 * 32), idesc_s, k4 != 0 ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
      };
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);            // S of the next block overlaps the softmax of this one
        const int s = j % STAGES;
        mbar_wait(p_full, j & 1);                    // P_j is in TMEM, O carries the right scale
        tcgen05_fence_after();
        const uint32_t v_a = smem_u32(sKV + s * KV_STAGE + AT_TILE);
        const uint32_t tP = (j & 1) ? tS1 : tS0;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)               // 16 keys per step: 8 TMEM columns of P, 2048 B along V rows
          umma_f16_ts(tO, tP + k4 * 8, umma_desc_sw128_mn(v_a + k4 * 2048), idesc_pv, (j | k4) != 0 ? 1u : 0u);
        umma_commit(pv_full);                        // PV_j accumulated
        umma_commit(&kv_empty[s]);                   // K/V stage free
      }
    }
  } else {
    // ===================== softmax: two threads per query row ===========================================
    const int quarter = warp & 3;                     // TMEM lane quarter of this warp (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;                 // which 32 key columns (and 32 output columns) of the row
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float m_ref = -INFINITY, l = 0.f;                 // reference maximum of the row's probabilities, this thread's part of the row sum
    const uint64_t sc2 = pk2(p.scale_log2e, p.scale_log2e);
    const float thr = AT_RESCALE_LOG2 / p.scale_log2e;           // raw-score units
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tcgen05_fence_after();
      uint32_t r[32];
      const uint32_t ts = ((j & 1) ? tS1 : tS0) + lane_off;
      tmem_ld_32x32(ts + half * 32, r);
      tmem_ld_wait();
      const int kbase = j * AT_BN + half * 32;
      if (j * AT_BN + AT_BN > kvlen) {                // only the last key block needs the length mask
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (kbase + i >= kvlen) r[i] = 0xff800000u;            // -inf
      }
      float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mxa = max3(mxa, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        mxb = max3(mxb, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
      }
      float mx = fmaxf(mxa, mxb);
      {                                               // the row's maximum over both halves
        float* slot = xch + (j & 1) * 256;
        slot[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        mx = fmaxf(mx, slot[(half ^ 1) * 128 + row]);
      }
      // lazy rescale: move the reference only when this block's maximum exceeds it by more than 2^8 (both threads of the row decide alike)
      const bool move = mx > m_ref + thr;
      float c = 1.f;
      if (move) {
        c = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - mx) * p.scale_log2e);
        m_ref = mx;
        l *= c;
      }
      bool synced = false;
      if (j > 0 && __any_sync(0xffffffffu, move)) {   // this warp's 32 rows x 32 accumulator columns, in place
        mbar_wait(pv_full, (j - 1) & 1);
        synced = true;
        tcgen05_fence_after();
        uint32_t ov[32];
        tmem_ld_32x32(tO + lane_off + half * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * c);
        tmem_st_32x32_x32(tO + lane_off + half * 32, ov);
      }
      const float mrs = m_ref * p.scale_log2e;
      const uint64_t nm2 = pk2(-mrs, -mrs);
      uint64_t sa = 0ull, sb = 0ull;
      uint32_t ph[16];                                // this thread's 32 probabilities as packed fp16 pairs
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float x0, x1, x2, x3;
        upk2(fma2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, nm2), x0, x1);
        upk2(fma2(pk2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), sc2, nm2), x2, x3);
        x0 = fast_exp2(x0); x1 = fast_exp2(x1); x2 = fast_exp2(x2); x3 = fast_exp2(x3);
        add2_acc(sa, pk2(x0, x1));
        add2_acc(sb, pk2(x2, x3));
        ph[i / 2] = pack_half2(x0, x1);
        ph[i / 2 + 1] = pack_half2(x2, x3);
      }
      {
        float s0, s1, s2, s3;
        upk2(sa, s0, s1); upk2(sb, s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      tmem_st_32x32_x16(ts + half * 16, ph);          // columns [16 half, 16 half + 16) of the S tile = this thread's 32 fp16 keys
      tmem_st_wait();
      tcgen05_fence_before();
      if (j > 0 && !synced) mbar_wait(pv_full, (j - 1) & 1);     // every warp observes every phase of pv_full (see attn_otm_kernel)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: the accumulated O, normalised by the row sum of both halves
    {
      float* slot = xch + (nblk & 1) * 256;
      slot[half * 128 + row] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      l += slot[(half ^ 1) * 128 + row];
    }
    mbar_wait(pv_full, (nblk - 1) & 1);
    tcgen05_fence_after();
    const long off = (long)(qrow0 + row) * p.ldo + head * 64 + half * 32;
    const bool live = q0 + row < qlen;
    const float inv = (live && l > 0.f) ? 1.f / l : 0.f;        // padding rows of the sequence's last tile are written as zeros
    uint32_t ov[32];
    tmem_ld_32x32(tO + lane_off + half * 32, ov);
    tmem_ld_wait();
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = live ? __uint_as_float(ov[i]) * inv : 0.f;
    if (p.O16) {
      uint4* d16 = reinterpret_cast<uint4*>(p.O16 + off);
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        d16[i / 8] = make_uint4(pack_half2(o[i], o[i + 1]), pack_half2(o[i + 2], o[i + 3]), pack_half2(o[i + 4], o[i + 5]), pack_half2(o[i + 6], o[i + 7]));
    } else if (p.Ohi) {
      uint4* dh = reinterpret_cast<uint4*>(p.Ohi + off);
      uint4* dl = reinterpret_cast<uint4*>(p.Olo + off);
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 h, lw;
        split_pair_at(o[i], o[i + 1], h.x, lw.x); split_pair_at(o[i + 2], o[i + 3], h.y, lw.y);
        split_pair_at(o[i + 4], o[i + 5], h.z, lw.z); split_pair_at(o[i + 6], o[i + 7], h.w, lw.w);
        dh[i / 8] = h; dl[i / 8] = lw;
      }
    } else if (live) {
      float* dst = p.O + off;
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}



// ================================================================================================
// attn_pp_kernel: TWO query tiles per CTA with their softmax warp groups in enforced anti-phase ("ping-pong", the structure
// FlashAttention-4 uses).  Why: with one tile per CTA and two CTAs per SM (attn_otm_kernel) the two CTAs run in phase -- both
// read their S tile through the 64 B/clk tcgen05.ld port, then both queue on the MUFU pipe -- and a key block costs the SUM
// of the two floors (measured 1040 clocks) instead of their maximum (512).  Here one CTA owns 256 query rows of a (sequence,
// head): tile A and tile B share every K/V stage (half the TMA traffic), each has its own S ring and O accumulator in TMEM
// (2 x (128 + 64) = 384 columns), and a pair of mbarriers hands the TMEM read port back and forth: group B may load S_B(j) only
// after group A has S_A(j) in registers, group A may load S_A(j+1) only after group B has S_B(j) -- so one group's
// exponentials always run under the other group's tcgen05.ld.
//   warp 0: TMA producer   warp 1: MMA issuer (+ TMEM owner)   warps 2-5: softmax of tile A   warps 6-9: softmax of tile B
// ================================================================================================
constexpr int APP_STAGES = 8;
constexpr int APP_SMEM = 2 * (2 * AT_TILE) + APP_STAGES * (2 * AT_TILE) + 512 + 1024;

template <int HANDOFF>
__global__ void __launch_bounds__(320, 1)
attn_pp_kernel(const __grid_constant__ CUtensorMap tm, const AttnTcDev p) {
  constexpr int STAGES = APP_STAGES;
  constexpr int Q_TILE = 2 * AT_TILE, KV_STAGE = 2 * AT_TILE;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int qlen = p.q_len[seq], kvlen = p.kv_len[seq];
  const int q0 = blockIdx.x * (2 * AT_BM);
  if (q0 >= qlen) return;
  const bool b_on = q0 + AT_BM < qlen;                  // the second tile holds real rows
  const int qrow0 = p.q_start[seq] + q0, krow0 = p.kv_start[seq];
  const int nblk = (kvlen + AT_BN - 1) / AT_BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // [tile][128 rows x 128 B]
  uint8_t* sKV = sQ + 2 * Q_TILE;                       // stages of [K 64 rows][V 64 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * KV_STAGE);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // [STAGES]
  uint64_t* kv_empty = kv_full + STAGES;    // [STAGES]
  uint64_t* s_full = kv_empty + STAGES;     // [tile][2]
  uint64_t* p_full = s_full + 4;            // [tile]
  uint64_t* pv_full = p_full + 2;           // [tile]
  uint64_t* ld_done = pv_full + 2;          // [tile]  the group has its S tile of the current block in registers (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 4; ++s) mbar_init(&s_full[s], 1);
    for (int t = 0; t < 2; ++t) { mbar_init(&p_full[t], 4); mbar_init(&pv_full[t], 1); mbar_init(&ld_done[t], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto tS = [&](int t, int j) { return tmem_base + (uint32_t)(t * 192 + (j & 1) * 64); };
  auto tO = [&](int t) { return tmem_base + (uint32_t)(t * 192 + 128); };

  if (warp == 0) {
    // ===================== TMA producer ===============================================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * Q_TILE);
      const int qc = p.q_col + head * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) tma_load_2d(sQ + i * AT_TILE, &tm, q_full, qc, qrow0 + 64 * i);      // rows past the buffer read as zeros
      const int kc = p.k_col + head * 64, vc = p.v_col + head * 64;
      for (int j = 0; j < nblk; ++j) {
        const int s = j % STAGES;
        mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
        uint8_t* st = sKV + s * KV_STAGE;
        mbar_arrive_expect_tx(&kv_full[s], KV_STAGE);
        const int r = krow0 + j * AT_BN;
        tma_load_2d(st, &tm, &kv_full[s], kc, r);
        tma_load_2d(st + AT_TILE, &tm, &kv_full[s], vc, r);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =================================================================
    // Issue order per key block: S_A(j+1), S_B(j+1), PV_A(j), PV_B(j).  S_t(j+1) overwrites the tile that held S_t(j-1) / P_t(j-1),
    // after PV_t(j-1) in issue order (the tensor pipe executes in issue order).
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(AT_BM, AT_BN);                   // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_f16(AT_BM, 64) | (1u << 16);        // B (= V) MN-major
      const int nt = b_on ? 2 : 1;
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(&kv_full[s], (j / STAGES) & 1);
        tcgen05_fence_after();
        const uint32_t k_a = smem_u32(sKV + s * KV_STAGE);
        for (int t = 0; t < nt; ++t) {
          const uint32_t q_a = smem_u32(sQ + t * Q_TILE);
          const uint32_t d = tS(t, j);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16(d, umma_desc_sw128(q_a + k4 * 32), umma_desc_sw128(k_a + k4 * 32), idesc_s, k4 != 0 ? 1u : 0u);
          umma_commit(&s_full[t * 2 + (j & 1)]);
        }
      };
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        const int s = j % STAGES;
        const uint32_t v_a = smem_u32(sKV + s * KV_STAGE + AT_TILE);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(&p_full[t], j & 1);              // P_t(j) is in TMEM, O_t carries the right scale
          tcgen05_fence_after();
          const uint32_t tP = tS(t, j);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_f16_ts(tO(t), tP + k4 * 8, umma_desc_sw128_mn(v_a + k4 * 2048), idesc_pv, (j | k4) != 0 ? 1u : 0u);
          umma_commit(&pv_full[t]);
        }
        umma_commit(&kv_empty[s]);                   // K/V stage free (both tiles' S and PV products of block j have read it)
      }
    }
  } else {
    // ===================== softmax: group A (warps 2-5) / group B (warps 6-9), one query row per thread =================
    const int t = (warp - 2) >> 2;                    // tile of this group
    if (t == 1 && !b_on) { /* nothing to do: fall through to the teardown */ }
    else {
      const int quarter = warp & 3;                   // TMEM lane quarter of this warp (hardware: warp id % 4)
      const int row = quarter * 32 + lane;
      const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
      const bool handoff = HANDOFF && b_on;
      float m_ref = -INFINITY, l = 0.f;
      const uint64_t sc2 = pk2(p.scale_log2e, p.scale_log2e);
      const float thr = AT_RESCALE_LOG2 / p.scale_log2e;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&s_full[t * 2 + (j & 1)], (j >> 1) & 1);
        // the TMEM read port goes A(0) B(0) A(1) B(1) ...: wait for the other group's load of the previous slot in that order
        if (handoff) {
          if (t == 0) { if (j > 0) mbar_wait(&ld_done[1], (j - 1) & 1); }
          else mbar_wait(&ld_done[0], j & 1);
        }
        tcgen05_fence_after();
        uint32_t r[64];
        const uint32_t ts = tS(t, j) + lane_off;
        tmem_ld_32x32(ts, *reinterpret_cast<uint32_t(*)[32]>(r));
        tmem_ld_32x32(ts + 32, *reinterpret_cast<uint32_t(*)[32]>(r + 32));
        tmem_ld_wait();
        if (handoff) { __syncwarp(); if (lane == 0) mbar_arrive(&ld_done[t]); }
        if (j * AT_BN + AT_BN > kvlen) {              // only the last key block needs the length mask
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (j * AT_BN + i >= kvlen) r[i] = 0xff800000u;
        }
        float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          mxa = max3(mxa, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
          mxb = max3(mxb, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        }
        const float mx = fmaxf(mxa, mxb);
        const bool move = mx > m_ref + thr;           // lazy rescale (see attn_otm_kernel)
        float c = 1.f;
        if (move) {
          c = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - mx) * p.scale_log2e);
          m_ref = mx;
          l *= c;
        }
        bool synced = false;
        if (j > 0 && __any_sync(0xffffffffu, move)) {
          mbar_wait(&pv_full[t], (j - 1) & 1);
          synced = true;
          tcgen05_fence_after();
#pragma unroll
          for (int cc = 0; cc < 64; cc += 32) {
            uint32_t ov[32];
            tmem_ld_32x32(tO(t) + lane_off + cc, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * c);
            tmem_st_32x32_x32(tO(t) + lane_off + cc, ov);
          }
        }
        const float mrs = m_ref * p.scale_log2e;
        const uint64_t nm2 = pk2(-mrs, -mrs);
        uint64_t sa = 0ull, sb = 0ull;
        uint32_t ph[32];
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          float x0, x1, x2, x3;
          upk2(fma2(pk2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, nm2), x0, x1);
          upk2(fma2(pk2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), sc2, nm2), x2, x3);
          x0 = fast_exp2(x0); x1 = fast_exp2(x1); x2 = fast_exp2(x2); x3 = fast_exp2(x3);
          add2_acc(sa, pk2(x0, x1));
          add2_acc(sb, pk2(x2, x3));
          ph[i / 2] = pack_half2(x0, x1);
          ph[i / 2 + 1] = pack_half2(x2, x3);
        }
        {
          float s0, s1, s2, s3;
          upk2(sa, s0, s1); upk2(sb, s2, s3);
          l += (s0 + s1) + (s2 + s3);
        }
        tmem_st_32x32_x32(ts, ph);
        tmem_st_wait();
        tcgen05_fence_before();
        if (j > 0 && !synced) mbar_wait(&pv_full[t], (j - 1) & 1);     // every warp observes every phase of pv_full
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // epilogue: the accumulated O, normalised
      mbar_wait(&pv_full[t], (nblk - 1) & 1);
      tcgen05_fence_after();
      const int qr = q0 + t * AT_BM + row;
      const long off = (long)(qrow0 + t * AT_BM + row) * p.ldo + head * 64;
      const bool live = qr < qlen;
      const float inv = (live && l > 0.f) ? 1.f / l : 0.f;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 32) {
        uint32_t ov[32];
        tmem_ld_32x32(tO(t) + lane_off + cc, ov);
        tmem_ld_wait();
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = live ? __uint_as_float(ov[i]) * inv : 0.f;
        if (p.O16) {
          uint4* d16 = reinterpret_cast<uint4*>(p.O16 + off + cc);
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            d16[i / 8] = make_uint4(pack_half2(o[i], o[i + 1]), pack_half2(o[i + 2], o[i + 3]), pack_half2(o[i + 4], o[i + 5]), pack_half2(o[i + 6], o[i + 7]));
        } else if (p.Ohi) {
          uint4* dh = reinterpret_cast<uint4*>(p.Ohi + off + cc);
          uint4* dl = reinterpret_cast<uint4*>(p.Olo + off + cc);
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 h, lw;
            split_pair_at(o[i], o[i + 1], h.x, lw.x); split_pair_at(o[i + 2], o[i + 3], h.y, lw.y);
            split_pair_at(o[i + 4], o[i + 5], h.z, lw.z); split_pair_at(o[i + 6], o[i + 7], h.w, lw.w);
            dh[i / 8] = h; dl[i / 8] = lw;
          }
        } else if (live) {
          float* dst = p.O + off + cc;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---- host --------------------------------------------------------------------------------------------

template <int SPLIT, int F16, int OCC> static void at_attr() {
  CBX_CHECK(cudaFuncSetAttribute(attn_tc_kernel<SPLIT, F16, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, F16 ? at_smem_f16(OCC) : AT_SMEM));
}
void attention_tc_init() {      // per device
  at_attr<1, 0, 1>(); at_attr<2, 0, 1>(); at_attr<1, 1, 1>(); at_attr<2, 1, 1>(); at_attr<1, 1, 2>(); at_attr<2, 1, 2>();
  CBX_CHECK(cudaFuncSetAttribute(attn_pp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, APP_SMEM));
  CBX_CHECK(cudaFuncSetAttribute(attn_pp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, APP_SMEM));
  CBX_CHECK(cudaFuncSetAttribute(attn_otm2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, af_smem(1, 1)));
  CBX_CHECK(cudaFuncSetAttribute(attn_otm2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, af_smem(2, 1)));
  CBX_CHECK(cudaFuncSetAttribute(attn_otm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, af_smem(1, 1)));
  CBX_CHECK(cudaFuncSetAttribute(attn_otm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, af_smem(2, 1)));
  CBX_CHECK(cudaFuncSetAttribute(attn_f16_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, af_smem(1, 0)));
  CBX_CHECK(cudaFuncSetAttribute(attn_f16_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, af_smem(1, 1)));
  CBX_CHECK(cudaFuncSetAttribute(attn_f16_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, af_smem(2, 0)));
  CBX_CHECK(cudaFuncSetAttribute(attn_f16_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, af_smem(2, 1)));
}

void attention_tc(Ctx& ctx, const AttnTcArgs& a) {
  if (ctx.dry) return;
  AttnTcDev p;
  p.O = a.O; p.ldo = a.ldo; p.Ohi = a.Ohi; p.Olo = a.Olo; p.O16 = a.O16; p.q_start = a.q_start; p.q_len = a.q_len; p.kv_start = a.kv_start; p.kv_len = a.kv_len;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  p.q_col = a.q_col; p.k_col = a.k_col; p.v_col = a.v_col;
  // CBX_ATTN_TC = softmax threads per query row (1 | 2), CBX_ATTN_OCC = CTAs per SM of the fp16 variant (1 | 2).
  // bf16x3 operands: the kernel is bound by shared-memory operand traffic of the N=64 SS-mode MMAs, one thread per row is
  // best (round 1: 1123 ms vs 1170 ms at B=32).  fp16 operands cut the MMA and operand traffic 3x and leave the softmax
  // (64 ex2 per row and key block on one warp per scheduler) exposed: see DESIGN.md for the measured variants.
  static const int variant = getenv("CBX_ATTN_TC") ? atoi(getenv("CBX_ATTN_TC")) : 1;
  static const int occ = getenv("CBX_ATTN_OCC") ? atoi(getenv("CBX_ATTN_OCC")) : 2;
  ctx.launches++;
  if (ctx.timer) ctx.timer->add(K_ATTN_TC, a.work, 0.0);
  if (ctx.timer) ctx.timer->begin(K_ATTN_TC, ctx.stream);
  dim3 grid((a.max_q_len + AT_BM - 1) / AT_BM, a.n_heads, a.n_seq);
  // CBX_ATTN_F16 = 2: attn_f16_kernel with P in TMEM (.ts PV product); 1: same kernel, P through shared memory;
  // 0: the round-2 first version (attn_tc_kernel<*, 1, *>)
  // 3 (default): attn_otm_kernel -- P and O both stay in TMEM (lazy rescale)
  // 4: attn_otm2_kernel -- the same with two softmax threads per query row (measured 5 % slower than 3, session 15)
  static const int f16_kernel = getenv("CBX_ATTN_F16") ? atoi(getenv("CBX_ATTN_F16")) : 3;
  if (a.f16 && f16_kernel >= 5 && variant == 1) {        // 5: attn_pp_kernel (two query tiles per CTA, anti-phase hand-off); 6: without the hand-off
    dim3 grid2((a.max_q_len + 2 * AT_BM - 1) / (2 * AT_BM), a.n_heads, a.n_seq);
    if (f16_kernel == 5) attn_pp_kernel<1><<<grid2, 320, APP_SMEM, ctx.stream>>>(*a.tm_hi, p);
    else attn_pp_kernel<0><<<grid2, 320, APP_SMEM, ctx.stream>>>(*a.tm_hi, p);
  } else if (a.f16 && f16_kernel >= 4 && variant == 1) {
    if (occ == 2) attn_otm2_kernel<2><<<grid, 320, af_smem(2, 1), ctx.stream>>>(*a.tm_hi, p);
    else attn_otm2_kernel<1><<<grid, 320, af_smem(1, 1), ctx.stream>>>(*a.tm_hi, p);
  } else if (a.f16 && f16_kernel >= 3 && variant == 1) {
    if (occ == 2) attn_otm_kernel<2><<<grid, 192, af_smem(2, 1), ctx.stream>>>(*a.tm_hi, p);
    else attn_otm_kernel<1><<<grid, 192, af_smem(1, 1), ctx.stream>>>(*a.tm_hi, p);
  } else if (a.f16 && f16_kernel >= 1 && variant == 1) {
    const int ptm = f16_kernel >= 2 ? 1 : 0;
    if (occ == 2) {
      if (ptm) attn_f16_kernel<2, 1><<<grid, 192, af_smem(2, 1), ctx.stream>>>(*a.tm_hi, p);
      else attn_f16_kernel<2, 0><<<grid, 192, af_smem(2, 0), ctx.stream>>>(*a.tm_hi, p);
    } else {
      if (ptm) attn_f16_kernel<1, 1><<<grid, 192, af_smem(1, 1), ctx.stream>>>(*a.tm_hi, p);
      else attn_f16_kernel<1, 0><<<grid, 192, af_smem(1, 0), ctx.stream>>>(*a.tm_hi, p);
    }
  } else
  if (a.f16) {       // single fp16 plane per operand (a.tm_hi maps it; a.tm_lo is not read)
    if (occ == 2) {
      if (variant == 1) attn_tc_kernel<1, 1, 2><<<grid, 192, at_smem_f16(2), ctx.stream>>>(*a.tm_hi, *a.tm_hi, p);
      else attn_tc_kernel<2, 1, 2><<<grid, 320, at_smem_f16(2), ctx.stream>>>(*a.tm_hi, *a.tm_hi, p);
    } else {
      if (variant == 1) attn_tc_kernel<1, 1, 1><<<grid, 192, at_smem_f16(1), ctx.stream>>>(*a.tm_hi, *a.tm_hi, p);
      else attn_tc_kernel<2, 1, 1><<<grid, 320, at_smem_f16(1), ctx.stream>>>(*a.tm_hi, *a.tm_hi, p);
    }
  } else {
    if (variant == 1) attn_tc_kernel<1, 0, 1><<<grid, 192, AT_SMEM, ctx.stream>>>(*a.tm_hi, *a.tm_lo, p);
    else attn_tc_kernel<2, 0, 1><<<grid, 320, AT_SMEM, ctx.stream>>>(*a.tm_hi, *a.tm_lo, p);
  }
  if (ctx.timer) ctx.timer->end(K_ATTN_TC, ctx.stream);
  CBX_CHECK(cudaGetLastError());
}

}  // namespace cbx
