// HiFT ResBlock convolutions (reference src/chatterbox/models/s3gen/hifigan.py:106-166): dilated 1-D convolutions with
// C_in = C_out = C in {64, 128, 256}, k in {3, 7, 11} taps, dilation in {1, 3, 5}, on packed channel-last rows.
//
// hift_conv_kernel<C>: persistent tcgen05 kernel, one CTA per SM walking 128-row output tiles.
//   * The input tile is staged ONCE per 64-channel block as 128 + (k-1)*dil rows of bf16 hi/lo planes (TMA, SWIZZLE_128B;
//     rows before the first / after the last row of the buffer and the zero gap between two sequences supply the
//     padding) and every tap reads it through a ROW-SHIFTED UMMA descriptor (start address + tap*dil*128 B): the
//     dilated receptive field lives in shared memory, each input row crosses L2 -> SM once per tile instead of k times
//     (tools/probe_rowshift.py: shifted SWIZZLE_128B descriptors are exact for every shift, base-offset field 0).
//   * Weights: resident in shared memory when all k*(C/64) tiles fit (C = 64; C = 128 with k = 3), else a ring.
//   * Accumulator double-buffered in TMEM (2 x C columns): the epilogue of tile i runs under the MMAs of tile i+1.
//   * Epilogue straight from registers, one 32-channel row segment per thread and trip, three fused forms:
//       mode 0  t      = snake(acc + bias; alpha)                    -> bf16 hi/lo planes (input of the next conv)
//       mode 1  x_new  = acc + bias + res   (fp32, in place allowed) [+ planes of snake(x_new; alpha) for the next branch]
//       mode 2  dst    = (accumulate ? dst : 0) + (acc + bias + res) * scale   (ResBlock output into the stage sum)
//     so no activation ever makes a separate pass over HBM.
//   warp 0: TMA producer   warp 1: MMA issuer (+ TMEM)   warps 2-9: epilogue
// Operands keep the fp32-faithful two-term split (x = hi + lo): the vocoder's 1e-4 waveform bar needs it.
// Algorithmic HBM bytes per launch: rows*C*4 (planes in) + k*C*C*2 (weights) + outputs (4 B / element each) (+ residual).
#include "engine.h"

namespace cbx {

constexpr int HC_THREADS = 320;
constexpr int HC_A_STAGES = 2;
constexpr int HC_RIN_MAX = 184;                         // 128 + 10 * 5 = 178 rows, rounded up to a multiple of 8
constexpr int HC_A_STAGE = 2 * HC_RIN_MAX * 128;        // hi + lo planes of one 64-channel block
constexpr int HC_W_BYTES = 98304;                       // weight region: resident tiles or a ring
constexpr int HC_SMEM = HC_A_STAGES * HC_A_STAGE + HC_W_BYTES + 2048 /*bias, alpha*/ + 512 /*barriers*/ + 1024 /*align*/;

struct HcDev {
  int k, dil, pad, rin;                 // taps, dilation, left padding in rows ((k-1)/2*dil), staged rows per tile
  int M, n_tiles;
  const int* tile_seq; const int* start; const int* len;
  const float* bias; const float* alpha;
  int mode;
  const float* res; float* out; int accumulate; float scale;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int w_res;                            // all weight tiles resident
};

__device__ __forceinline__ float snake_f(float v, float a) {      // hifigan.py:79-82
  const float s = sinf(v * a);
  return v + (1.0f / (a + 1e-9f)) * (s * s);
}

template <int C>
__global__ void __launch_bounds__(HC_THREADS, 1)
hift_conv_kernel(const __grid_constant__ CUtensorMap tmHi, const __grid_constant__ CUtensorMap tmLo,
                 const __grid_constant__ CUtensorMap tmW, const HcDev p) {
  constexpr int KB = C / 64;                            // 64-channel blocks
  constexpr int WT = C * 128;                           // bytes of one weight tile [C out rows][64 in channels]
  constexpr int NW = HC_W_BYTES / WT;                   // ring stages (3 | 6 | 12)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sW = smem + HC_A_STAGES * HC_A_STAGE;
  float* sBias = reinterpret_cast<float*>(sW + HC_W_BYTES);
  float* sAlpha = sBias + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAlpha + 256);
  uint64_t* a_full = bars;                      // [2]
  uint64_t* a_empty = bars + 2;                 // [2]
  uint64_t* w_full = bars + 4;                  // [NW <= 12] (slot 0 doubles as "all resident tiles landed")
  uint64_t* w_empty = bars + 16;                // [NW]
  uint64_t* acc_full = bars + 28;               // [2]
  uint64_t* acc_empty = bars + 30;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int plane = p.rin * 128;                        // bytes of one staged plane

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    for (int s = 0; s < NW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * C>(tmem_slot);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmHi); tma_prefetch_desc(&tmLo); tma_prefetch_desc(&tmW); }
  for (int i = threadIdx.x; i < C; i += HC_THREADS) { sBias[i] = p.bias ? p.bias[i] : 0.f; sAlpha[i] = p.alpha ? p.alpha[i] : 1.f; }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer ====================================================================
    if (lane == 0) {
      if (p.w_res) {
        mbar_arrive_expect_tx(&w_full[0], (uint32_t)(p.k * KB * WT));
        for (int i = 0; i < p.k * KB; ++i) tma_load_2d(sW + i * WT, &tmW, &w_full[0], i * 64, 0);
      }
      uint32_t ga = 0, gw = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int row0 = tile * 128 - p.pad;            // may be negative / run past the end: TMA fills zeros
        for (int kb = 0; kb < KB; ++kb, ++ga) {
          const int s = ga % HC_A_STAGES;
          mbar_wait(&a_empty[s], ((ga / HC_A_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&a_full[s], (uint32_t)(2 * plane));
          tma_load_2d(sA + s * HC_A_STAGE, &tmHi, &a_full[s], kb * 64, row0);
          tma_load_2d(sA + s * HC_A_STAGE + plane, &tmLo, &a_full[s], kb * 64, row0);
          if (!p.w_res) {
            for (int t = 0; t < p.k; ++t, ++gw) {
              const int sw = gw % NW;
              mbar_wait(&w_empty[sw], ((gw / NW) & 1) ^ 1);
              mbar_arrive_expect_tx(&w_full[sw], WT);
              tma_load_2d(sW + sw * WT, &tmW, &w_full[sw], (t * KB + kb) * 64, 0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, C);
      if (p.w_res) mbar_wait(&w_full[0], 0);
      uint32_t ga = 0, gw = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(buf * C);
        uint32_t first = 0;
        for (int kb = 0; kb < KB; ++kb, ++ga) {
          const int s = ga % HC_A_STAGES;
          mbar_wait(&a_full[s], (ga / HC_A_STAGES) & 1);
          tcgen05_fence_after();
          const uint32_t a_stage = smem_u32(sA + s * HC_A_STAGE);
          for (int t = 0; t < p.k; ++t) {
            uint32_t w_addr;
            int sw = 0;
            if (p.w_res) {
              w_addr = smem_u32(sW + (t * KB + kb) * WT);
            } else {
              sw = gw % NW;
              mbar_wait(&w_full[sw], (gw / NW) & 1);
              tcgen05_fence_after();
              w_addr = smem_u32(sW + sw * WT);
            }
            const uint32_t a_hi = a_stage + (uint32_t)(t * p.dil) * 128u;      // tap t: the same tile, t*dil rows further down
            const uint32_t a_lo = a_hi + (uint32_t)plane;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t db = umma_desc_sw128(w_addr + k4 * 32);
              umma_bf16(d, umma_desc_sw128(a_lo + k4 * 32), db, idesc, first);      // small plane first
              umma_bf16(d, umma_desc_sw128(a_hi + k4 * 32), db, idesc, 1u);
              first = 1u;
            }
            if (!p.w_res) { umma_commit(&w_empty[sw]); ++gw; }
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue ========================================================================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const long row = (long)tile * 128 + q * 32 + lane;
      const int seq = p.tile_seq[tile];
      const bool valid = seq >= 0 && row < p.M && (int)(row - p.start[seq]) < p.len[seq];
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < C / 2; cc += 32) {
        const int n = half * (C / 2) + cc;
        float rres[32];
        if (p.mode != 0 && valid) {
          const float4* rp = reinterpret_cast<const float4*>(p.res + row * C + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 t = rp[j]; rres[4 * j] = t.x; rres[4 * j + 1] = t.y; rres[4 * j + 2] = t.z; rres[4 * j + 3] = t.w; }
          if (p.mode == 2 && p.accumulate) {
            const float4* dp = reinterpret_cast<const float4*>(p.out + row * C + n);
            // v = dst + (acc + bias + res) * scale : fold dst / scale into the residual term
            const float inv = 1.0f / p.scale;
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 t = dp[j]; rres[4 * j] += t.x * inv; rres[4 * j + 1] += t.y * inv; rres[4 * j + 2] += t.z * inv; rres[4 * j + 3] += t.w * inv; }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) rres[j] = 0.f;
        }
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * C + n), r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + sBias[n + j];
        if (row >= p.M) continue;
        if (p.mode == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? snake_f(v[j], sAlpha[n + j]) : 0.f;
        } else {
          const float sc = p.mode == 2 ? p.scale : 1.0f;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? (v[j] + rres[j]) * sc : 0.f;
          float4* dst = reinterpret_cast<float4*>(p.out + row * C + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (p.out_hi) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = valid ? snake_f(v[j], sAlpha[n + j]) : 0.f;
          }
        }
        if (p.out_hi) {
          uint4* dh = reinterpret_cast<uint4*>(p.out_hi + row * C + n);
          uint4* dl = reinterpret_cast<uint4*>(p.out_lo + row * C + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(v[8 * j + 2 * e], h0, l0);
              split_bf16(v[8 * j + 2 * e + 1], h1, l1);
              h[e] = pack_bf16(h0, h1); l[e] = pack_bf16(l0, l1);
            }
            dh[j] = make_uint4(h[0], h[1], h[2], h[3]);
            dl[j] = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * C>(tmem_base);
}

// planes of snake(x; alpha) (the first activation of a ResBlock, hifigan.py:156): one pass over x
__global__ void snake_planes_kernel(const float* x, int C, const float* alpha, __nv_bfloat16* hi, __nv_bfloat16* lo,
                                    const int* tile_seq, const int* start, const int* len, long rows) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= rows * C) return;
  const long r = i / C; const int c = (int)(i - r * C);
  const int seq = tile_seq[r / kTileM];
  const bool valid = seq >= 0 && (int)(r - start[seq]) < len[seq];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    v = *reinterpret_cast<const float4*>(x + i);
    v.x = snake_f(v.x, alpha[c]); v.y = snake_f(v.y, alpha[c + 1]); v.z = snake_f(v.z, alpha[c + 2]); v.w = snake_f(v.w, alpha[c + 3]);
  }
  __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
  split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
  *reinterpret_cast<uint2*>(hi + i) = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
  *reinterpret_cast<uint2*>(lo + i) = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
}

void hift_conv_init() {
  CBX_CHECK(cudaFuncSetAttribute(hift_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, HC_SMEM));
  CBX_CHECK(cudaFuncSetAttribute(hift_conv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, HC_SMEM));
  CBX_CHECK(cudaFuncSetAttribute(hift_conv_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, HC_SMEM));
}

void snake_planes(Ctx& ctx, const float* x, int C, const float* alpha, __nv_bfloat16* hi, __nv_bfloat16* lo,
                  const cbx_layout& L) {
  if (ctx.dry) return;
  ctx.launches++;
  const long n4 = (long)L.rows * C / 4;
  snake_planes_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, ctx.stream>>>(x, C, alpha, hi, lo, L.tile_seq, L.start, L.len, L.rows);
  CBX_CHECK(cudaGetLastError());
}

// one dilated conv of a ResBlock on planes (see the kernel header for the epilogue modes)
void hift_conv(Ctx& ctx, const Weight& W, int C, int k, int dil, const cbx_layout& L, const __nv_bfloat16* in_hi,
               const __nv_bfloat16* in_lo, int mode, const float* alpha, const float* res, float* out, int accumulate,
               float scale, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo) {
  if (ctx.dry) return;
  CBX_REQUIRE(C == 64 || C == 128 || C == 256, "ResBlock width");
  CBX_REQUIRE(W.N == C && W.Kpad == k * C && L.rows % 128 == 0, "packed conv weight / layout");
  HcDev p;
  memset(&p, 0, sizeof(p));
  p.k = k; p.dil = dil; p.pad = (k - 1) / 2 * dil;
  p.rin = (128 + (k - 1) * dil + 7) / 8 * 8;
  CBX_REQUIRE(p.rin <= HC_RIN_MAX, "receptive field larger than the staged tile");
  p.M = L.rows; p.n_tiles = L.rows / 128;
  p.tile_seq = L.tile_seq; p.start = L.start; p.len = L.len;
  p.bias = W.bias; p.alpha = alpha; p.mode = mode; p.res = res; p.out = out; p.accumulate = accumulate; p.scale = scale;
  p.out_hi = out_hi; p.out_lo = out_lo;
  const int tiles_w = k * (C / 64);
  p.w_res = (tiles_w * C * 128 <= HC_W_BYTES) ? 1 : 0;
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; CBX_CHECK(cudaGetDevice(&dev)); CBX_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)); }
  CUtensorMap tmHi, tmLo;
  make_plane_tmap(&tmHi, in_hi, L.rows, C, p.rin, C);
  make_plane_tmap(&tmLo, in_lo, L.rows, C, p.rin, C);
  const int grid = p.n_tiles < n_sm ? p.n_tiles : n_sm;
  ctx.launches++;
  if (ctx.timer) ctx.timer->add(K_HIFT_CONV, 2.0 * (double)L.rows * C * (double)(k * C),
                                (double)L.rows * C * 4.0 + (double)k * C * C * 2.0 + (double)L.rows * C * 4.0 * ((out ? 1 : 0) + (out_hi ? 1 : 0) + (res ? 1 : 0)));
  if (ctx.timer) ctx.timer->begin(K_HIFT_CONV, ctx.stream);
  if (C == 64) hift_conv_kernel<64><<<grid, HC_THREADS, HC_SMEM, ctx.stream>>>(tmHi, tmLo, W.tmap[0], p);
  else if (C == 128) hift_conv_kernel<128><<<grid, HC_THREADS, HC_SMEM, ctx.stream>>>(tmHi, tmLo, W.tmap[1], p);
  else hift_conv_kernel<256><<<grid, HC_THREADS, HC_SMEM, ctx.stream>>>(tmHi, tmLo, W.tmap[2], p);
  if (ctx.timer) ctx.timer->end(K_HIFT_CONV, ctx.stream);
  CBX_CHECK(cudaGetLastError());
}

}  // namespace cbx
