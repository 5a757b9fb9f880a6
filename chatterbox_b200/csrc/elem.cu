// Row-wise norms, elementwise kernels, gathers and the T3 sampler of libcbx (all fp32, HBM-bound).
#include "ops.h"
#include "kernels.h"

namespace cbx {

// ------------------------------------------------------------------------------------------------
// block reduce helpers (blockDim.x <= 1024)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (warp == 0) { r = warp_sum(r); if (lane == 0) sh[0] = r; }
  __syncthreads();
  r = sh[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : -INFINITY;
  if (warp == 0) { r = warp_max(r); if (lane == 0) sh[0] = r; }
  __syncthreads();
  r = sh[0];
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm (modeling_llama.py LlamaRMSNorm: x * rsqrt(mean(x^2)+eps) * w), one CTA per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* x, int ldx, const float* w, float* y, int ldy, int dim,
                                                      float eps, const int* row_idx, __nv_bfloat16* yhi, __nv_bfloat16* ylo) {
  __shared__ float sh[32];
  const int r = blockIdx.x;
  const float* xr = x + (long)(row_idx ? row_idx[r] : r) * ldx;
  float ss = 0.f;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) { float v = xr[i]; ss += v * v; }
  ss = block_sum(ss, sh);
  const float inv = rsqrtf(ss / dim + eps);
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    const float v = w[i] * (xr[i] * inv);
    if (yhi) { __nv_bfloat16 h, l; split_bf16(v, h, l); yhi[(long)r * ldy + i] = h; ylo[(long)r * ldy + i] = l; }
    else y[(long)r * ldy + i] = v;
  }
}
void rmsnorm(Ctx& ctx, const float* x, int ldx, const float* w, float* y, int ldy, int rows, int dim, float eps,
             const int* row_idx, __nv_bfloat16* yhi, __nv_bfloat16* ylo) {
  if (ctx.dry || rows == 0) return;
  ctx.launches++;
  rmsnorm_kernel<<<rows, 256, 0, ctx.stream>>>(x, ldx, w, y, ldy, dim, eps, row_idx, yhi, ylo);
  CBX_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over channels (+ activation, + per-sequence vector add, + scale), one warp per row.
// y = (act(LN(x)) * valid + seq_add[seq]) * out_scale ; padding rows of the packed layout are zeroed.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* x, int ldx, const float* w, const float* b, float* y,
                                                        int ldy, int rows, int dim, float eps, int act, float out_scale,
                                                        const float* seq_add, int seq_add_ld, int has_seq, SeqMap seq,
                                                        __nv_bfloat16* yhi, __nv_bfloat16* ylo, __half* y16) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  int s = 0; bool valid = true;
  if (has_seq) {
    s = seq.tile_seq[r / kTileM];
    valid = (s >= 0) && (r - seq.out_start[s] < seq.out_len[s]);
  }
  float* yr = y + (long)r * ldy;
  __nv_bfloat16* hr = yhi ? yhi + (long)r * ldy : nullptr;
  __nv_bfloat16* lr = yhi ? ylo + (long)r * ldy : nullptr;
  __half* r16 = y16 ? y16 + (long)r * ldy : nullptr;
  if (!valid) {
    for (int i = lane; i < dim; i += 32) {
      if (r16) r16[i] = __float2half_rn(0.f);
      else if (yhi) { hr[i] = __float2bfloat16(0.f); lr[i] = __float2bfloat16(0.f); } else yr[i] = 0.f;
    }
    return;
  }
  const float* xr = x + (long)r * ldx;
  float sum = 0.f;
  for (int i = lane; i < dim; i += 32) sum += xr[i];
  const float mean = warp_sum(sum) / dim;
  float var = 0.f;
  for (int i = lane; i < dim; i += 32) { float d = xr[i] - mean; var += d * d; }
  const float inv = rsqrtf(warp_sum(var) / dim + eps);
  for (int i = lane; i < dim; i += 32) {
    float v = (xr[i] - mean) * inv * w[i] + b[i];
    v = act_apply(act, v, 0.f);
    if (seq_add) v += seq_add[(long)s * seq_add_ld + i];
    v *= out_scale;
    if (r16) r16[i] = __float2half_rn(v);
    else if (yhi) { __nv_bfloat16 h, l; split_bf16(v, h, l); hr[i] = h; lr[i] = l; } else yr[i] = v;
  }
}
// Register-resident variant for dim = 128 * CH (256 / 512): the row is read ONCE (one 16-byte load per lane and 128-column
// chunk), mean and variance come from the registers with the same two-pass arithmetic, and the result leaves as 16-byte
// (fp32), 8-byte (fp16 plane) or 2 x 8-byte (bf16 hi/lo planes) stores.  Same operation order per element as layernorm_kernel.
template <int CH>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float* x, int ldx, const float* w, const float* b, float* y,
                                                            int ldy, int rows, float eps, int act, float out_scale,
                                                            const float* seq_add, int seq_add_ld, int has_seq, SeqMap seq,
                                                            __nv_bfloat16* yhi, __nv_bfloat16* ylo, __half* y16) {
  constexpr int dim = CH * 128;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  int s = 0; bool valid = true;
  if (has_seq) {
    s = seq.tile_seq[r / kTileM];
    valid = (s >= 0) && (r - seq.out_start[s] < seq.out_len[s]);
  }
  float v[CH][4];
  if (valid) {
    const float* xr = x + (long)r * ldx;
#pragma unroll
    for (int c = 0; c < CH; ++c) { const float4 t = *reinterpret_cast<const float4*>(xr + c * 128 + lane * 4); v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w; }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) sum += v[c][e];
    const float mean = warp_sum(sum) / dim;
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float d = v[c][e] - mean; var += d * d; }
    const float inv = rsqrtf(warp_sum(var) / dim + eps);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i0 = c * 128 + lane * 4;
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + i0)), bv = __ldg(reinterpret_cast<const float4*>(b + i0));
      const float ww[4] = {wv.x, wv.y, wv.z, wv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t = (v[c][e] - mean) * inv * ww[e] + bb[e];
        t = act_apply(act, t, 0.f);
        if (seq_add) t += seq_add[(long)s * seq_add_ld + i0 + e];
        v[c][e] = t * out_scale;
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) v[c][e] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const long o = (long)r * ldy + c * 128 + lane * 4;
    if (y16) {
      const __half2 h0 = __floats2half2_rn(v[c][0], v[c][1]), h1 = __floats2half2_rn(v[c][2], v[c][3]);
      *reinterpret_cast<uint2*>(y16 + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    } else if (yhi) {
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_bf16(v[c][e], h[e], l[e]);
      *reinterpret_cast<uint2*>(yhi + o) = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
      *reinterpret_cast<uint2*>(ylo + o) = make_uint2(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]));
    } else {
      *reinterpret_cast<float4*>(y + o) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
    }
  }
}

void layernorm(Ctx& ctx, const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int rows, int dim,
               float eps, int act, float out_scale, const float* seq_add, int seq_add_ld, const SeqMap* seq,
               __nv_bfloat16* yhi, __nv_bfloat16* ylo, __half* y16) {
  if (ctx.dry || rows == 0) return;
  ctx.launches++;
  SeqMap sm; if (seq) sm = *seq;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const void* outp = y16 ? (const void*)y16 : yhi ? (const void*)yhi : (const void*)y;
  const bool vec_ok = (dim == 256 || dim == 512) && (ldx % 4) == 0 && (ldy % 4) == 0 && al16(x) && al16(w) && al16(b) && al16(outp) &&
                      (!yhi || al16(ylo));
  if (vec_ok) {
    if (dim == 256) layernorm_vec_kernel<2><<<(rows + 7) / 8, 256, 0, ctx.stream>>>(x, ldx, w, b, y, ldy, rows, eps, act, out_scale, seq_add, seq_add_ld,
                                                                                   seq ? 1 : 0, sm, yhi, ylo, y16);
    else layernorm_vec_kernel<4><<<(rows + 7) / 8, 256, 0, ctx.stream>>>(x, ldx, w, b, y, ldy, rows, eps, act, out_scale, seq_add, seq_add_ld,
                                                                         seq ? 1 : 0, sm, yhi, ylo, y16);
    CBX_CHECK(cudaGetLastError());
    return;
  }
  layernorm_kernel<<<(rows + 7) / 8, 256, 0, ctx.stream>>>(x, ldx, w, b, y, ldy, rows, dim, eps, act, out_scale, seq_add,
                                                          seq_add_ld, seq ? 1 : 0, sm, yhi, ylo, y16);
  CBX_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------
__global__ void ew_act_kernel(const float* x, int ldx, float* y, int ldy, long rows, int cols, int act, float p,
                              const float* vec) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long r = i / cols; const int c = (int)(i - r * cols);
  y[r * ldy + c] = act_apply(act, x[r * ldx + c], vec ? vec[c] : p);
}
void ew_act(Ctx& ctx, const float* x, int ldx, float* y, int ldy, long rows, int cols, int act, float p, const float* vec) {
  if (ctx.dry || rows == 0) return;
  ctx.launches++;
  ew_act_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, ctx.stream>>>(x, ldx, y, ldy, rows, cols, act, p, vec);
  CBX_CHECK(cudaGetLastError());
}
__global__ void copy2d_kernel(const float* src, int lds, float* dst, int ldd, long rows, int cols) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long r = i / cols; const int c = (int)(i - r * cols);
  dst[r * ldd + c] = src[r * lds + c];
}
void copy2d(Ctx& ctx, const float* src, int lds, float* dst, int ldd, long rows, int cols) {
  if (ctx.dry || rows == 0) return;
  ctx.launches++;
  copy2d_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, ctx.stream>>>(src, lds, dst, ldd, rows, cols);
  CBX_CHECK(cudaGetLastError());
}
__global__ void fill_kernel(float* p, long n, float v) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
void fill(Ctx& ctx, float* p, long n, float v) {
  if (ctx.dry || n == 0) return;
  ctx.launches++;
  fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx.stream>>>(p, n, v);
  CBX_CHECK(cudaGetLastError());
}

// out[r] = table[ids[r]] (+ add_table[add_ids ? add_ids[r] : r]) ; one CTA per row
__global__ void gather_rows_kernel(const float* table, int ld, const int* ids, float* out, int ldo, int dim,
                                   const float* add_table, int add_ld, const int* add_ids, int id_limit) {
  const int r = blockIdx.x;
  int id = ids[r];
  if (id < 0 || id >= id_limit) id = 0;
  const float* t = table + (long)id * ld;
  const float* a = add_table ? add_table + (long)(add_ids ? add_ids[r] : r) * add_ld : nullptr;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) out[(long)r * ldo + i] = t[i] + (a ? a[i] : 0.f);
}
void gather_rows(Ctx& ctx, const float* table, int ld, const int* ids, float* out, int ldo, int rows, int dim,
                 const float* add_table, int add_ld, const int* add_ids, int id_limit) {
  if (ctx.dry || rows == 0) return;
  ctx.launches++;
  gather_rows_kernel<<<rows, 128, 0, ctx.stream>>>(table, ld, ids, out, ldo, dim, add_table, add_ld, add_ids, id_limit);
  CBX_CHECK(cudaGetLastError());
}

// fp32 [N][K] (ld) -> padded bf16 hi / lo planes [Npad][Kpad] (activation-derived "weights", e.g. rel-pos P)
__global__ void pack_hilo_kernel(const float* src, int ld, int N, int K, __nv_bfloat16* hi, __nv_bfloat16* lo, int Npad,
                                 int Kpad) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)Npad * Kpad) return;
  const int n = (int)(i / Kpad), k = (int)(i - (long)n * Kpad);
  float v = (n < N && k < K) ? src[(long)n * ld + k] : 0.f;
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  hi[i] = h; lo[i] = l;
}
// fp32 [N][K] -> one bf16 matrix [Npad][2K] = [hi | lo] (so that x.[hi|lo]^T with x repeated twice along K is x.W^T)
__global__ void pack_hilo_cat_kernel(const float* src, int ld, int N, int K, __nv_bfloat16* out, int Npad) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)Npad * K) return;
  const int n = (int)(i / K), k = (int)(i - (long)n * K);
  float v = (n < N) ? src[(long)n * ld + k] : 0.f;
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  out[(long)n * 2 * K + k] = h;
  out[(long)n * 2 * K + K + k] = l;
}
void pack_hilo_cat(Ctx& ctx, const float* src, int ld, int N, int K, __nv_bfloat16* out, int Npad) {
  if (ctx.dry) return;
  ctx.launches++;
  pack_hilo_cat_kernel<<<(unsigned)(((long)Npad * K + 255) / 256), 256, 0, ctx.stream>>>(src, ld, N, K, out, Npad);
  CBX_CHECK(cudaGetLastError());
}
// fp32 activations [rows][cols] (ld) of a packed batch -> bf16 hi/lo planes [rows][ldp], ZERO on layout padding rows: the
// operand of a plane-fed conv GEMM (its taps read the rows before a sequence start, which are the padding rows of the
// previous sequence).  4 elements per thread (16-byte load, two 8-byte stores); cols % 4 == 0.
__global__ void pack_planes_seq_kernel(const float* src, int ld, long rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, int ldp,
                                       SeqMap seq) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = cols >> 2;
  if (i >= rows * c4) return;
  const long r = i / c4; const int c = (int)(i - r * c4) << 2;
  const int s = seq.tile_seq[r / kTileM];
  const bool valid = (s >= 0) && (r - seq.out_start[s] < seq.out_len[s]);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) v = *reinterpret_cast<const float4*>(src + r * ld + c);
  __nv_bfloat16 h[4], l[4];
  split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
  *reinterpret_cast<uint2*>(hi + r * ldp + c) = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
  *reinterpret_cast<uint2*>(lo + r * ldp + c) = make_uint2(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]));
}
void pack_planes_seq(Ctx& ctx, const float* src, int ld, long rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, int ldp,
                     const SeqMap& seq) {
  if (ctx.dry || rows == 0) return;
  CBX_REQUIRE((cols % 4) == 0 && (ld % 4) == 0 && (ldp % 4) == 0, "pack_planes_seq: 4-element groups");
  ctx.launches++;
  pack_planes_seq_kernel<<<(unsigned)((rows * (cols >> 2) + 255) / 256), 256, 0, ctx.stream>>>(src, ld, rows, cols, hi, lo, ldp, seq);
  CBX_CHECK(cudaGetLastError());
}
void pack_hilo(Ctx& ctx, const float* src, int ld, int N, int K, __nv_bfloat16* hi, __nv_bfloat16* lo, int Npad, int Kpad) {
  if (ctx.dry) return;
  ctx.launches++;
  pack_hilo_kernel<<<(unsigned)(((long)Npad * Kpad + 255) / 256), 256, 0, ctx.stream>>>(src, ld, N, K, hi, lo, Npad, Kpad);
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// T3: embedding assembly for the prefill
// token i: row = tok_row[i], pos = tok_pos[i];  [cond(34) | text(n_text) | BOS | BOS]
// ================================================================================================
__global__ void t3_embed_kernel(float* out, int n_tok, const int* tok_row, const int* tok_pos, const float* cond,
                                const int* row_voice, int len_cond, const int* text_flat, const int* text_start,
                                const int* n_text, const int* row_uncond, const float* text_emb, int text_vocab,
                                const float* text_pos, const float* speech_emb, const float* speech_pos, int bos_id,
                                const float* wpe) {
  const int i = blockIdx.x;
  const int row = tok_row[i], pos = tok_pos[i];
  float* o = out + (long)i * 1024;
  const int nt = n_text[row];
  if (pos < len_cond) {
    const float* c = cond + ((long)row_voice[row] * len_cond + pos) * 1024;
    for (int d = threadIdx.x; d < 1024; d += blockDim.x) o[d] = c[d];
  } else if (pos < len_cond + nt) {
    const int j = pos - len_cond;
    int id = text_flat[text_start[row] + j];
    if (id < 0 || id >= text_vocab) id = 0;
    const float* e = text_emb + (long)id * 1024;
    const float* pe = text_pos ? text_pos + (long)j * 1024 : nullptr;       // Turbo has no learned input tables
    const bool unc = row_uncond[row] != 0;
    for (int d = threadIdx.x; d < 1024; d += blockDim.x) o[d] = (unc ? 0.f : e[d]) + (pe ? pe[d] : 0.f);
  } else {
    const float* e = speech_emb + (long)bos_id * 1024;
    for (int d = threadIdx.x; d < 1024; d += blockDim.x) o[d] = e[d] + (speech_pos ? speech_pos[d] : 0.f);
  }
  if (wpe) {   // GPT2Model adds wpe[position] to inputs_embeds (modeling_gpt2.py GPT2Model.forward)
    __syncthreads();
    const float* w = wpe + (long)pos * 1024;
    for (int d = threadIdx.x; d < 1024; d += blockDim.x) o[d] += w[d];
  }
}
void t3_embed(Ctx& ctx, float* out, int n_tok, const int* tok_row, const int* tok_pos, const float* cond,
              const int* row_voice, int len_cond, const int* text_flat, const int* text_start, const int* n_text,
              const int* row_uncond, const float* text_emb, int text_vocab, const float* text_pos,
              const float* speech_emb, const float* speech_pos, int bos_id, const float* wpe) {
  if (ctx.dry || n_tok == 0) return;
  ctx.launches++;
  t3_embed_kernel<<<n_tok, 256, 0, ctx.stream>>>(out, n_tok, tok_row, tok_pos, cond, row_voice, len_cond, text_flat,
                                                text_start, n_text, row_uncond, text_emb, text_vocab, text_pos,
                                                speech_emb, speech_pos, bos_id, wpe);
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// T3 sampler: CFG combine -> repetition penalty -> temperature -> min-p -> top-p -> softmax ->
// multinomial(1) == argmax(p / q), q ~ Exp(1)   (reference t3.py:339-368; transformers logits processors)
// one CTA (1024 threads) per active utterance; also emits the next input embedding for its slots.
// ================================================================================================
__device__ __forceinline__ uint32_t philox_mix(uint64_t seed, uint32_t a, uint32_t b, uint32_t c) {
  // counter-based hash (splitmix-style); statistical quality is ample for Exp(1) sampling noise
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)a + 1) + 0xBF58476D1CE4E5B9ull * ((uint64_t)b + 1) +
               0x94D049BB133111EBull * ((uint64_t)c + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}

constexpr int SV = 8194;        // speech vocab
constexpr int SV_PAD = 16384;   // bitonic sort size

__global__ void __launch_bounds__(1024) t3_sample_kernel(T3SampleDev p) {
  extern __shared__ float smf[];
  float* lg = smf;                              // [SV] logits -> probabilities
  float* sh = smf + SV;                         // [64] reduction scratch
  float* skey = sh + 64;                        // [SV_PAD] sort keys (top-k / top-p only)
  int* sidx = reinterpret_cast<int*>(skey + SV_PAD);   // [SV_PAD]
  const int j = blockIdx.x;
  pdl_wait();
  pdl_launch_dependents();
  if (p.n_act && j >= *p.n_act) return;         // retired slot (device-side compaction, t3_compact_kernel)
  const int utt = p.act_utt[j];
  if (p.done[utt]) return;
  const int V = p.vocab;                        // 8194, Turbo 6563 (tables keep the SV stride)
  const bool turbo = p.turbo != 0;
  const int rows_per = p.cfg ? 2 : 1;
  // logits were written at the slot the utterance had BEFORE this step's compaction
  const float* lc = p.logits + (long)((p.src_slot ? p.src_slot[j] : j) * rows_per) * p.ldl;
  const float* lu = p.cfg ? lc + p.ldl : nullptr;
  const int step = p.n_gen[utt];
  unsigned char* seen = p.seen + (long)utt * SV;
  const float w = p.cfg_weight, rp = p.rep_penalty;
  // T3.inference (t3.py:339-356): CFG combine, repetition penalty (history incl. BOS), temperature, min-p, top-p.
  // T3.inference_turbo (t3.py:396-404): temperature, top-k, top-p, repetition penalty (history: BOS for the first
  // token, then the generated ids only).
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float l = lc[v];
    if (!turbo) {
      if (lu) l = l + w * (l - lu[v]);
      if (seen[v]) l = (l < 0.f) ? l * rp : l / rp;
      if (p.temperature != 1.0f) l = l / p.temperature;
    } else if (p.temperature > 0.f && p.temperature != 1.0f) {
      l = l / p.temperature;
    }
    lg[v] = l;
    mx = fmaxf(mx, l);
  }
  mx = block_max(mx, sh);
  if (!turbo) {
    // min-p: drop softmax(l) < min_p * max prob (max prob = 1/Z); keep >= 1 token (the max itself)
    float z = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) z += expf(lg[v] - mx);
    z = block_sum(z, sh);
    const float pmax = 1.0f / z;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float pr = expf(lg[v] - mx) / z;
      if (pr < p.min_p * pmax && lg[v] < mx) lg[v] = -INFINITY;
    }
  }
  __syncthreads();
  const bool use_topk = turbo && p.top_k > 0 && p.top_k < V;
  if (p.top_p < 1.0f || use_topk) {
    // ascending bitonic sort of (logit, id); ids >= V are +inf padding and end up behind the vocabulary
    for (int v = threadIdx.x; v < SV_PAD; v += blockDim.x) { skey[v] = (v < V) ? lg[v] : INFINITY; sidx[v] = v; }
    __syncthreads();
    for (int k = 2; k <= SV_PAD; k <<= 1)
      for (int jj = k >> 1; jj > 0; jj >>= 1) {
        for (int i = threadIdx.x; i < SV_PAD; i += blockDim.x) {
          const int ixj = i ^ jj;
          if (ixj > i) {
            const bool up = ((i & k) == 0);
            const float a = skey[i], b = skey[ixj];
            const bool sw = up ? (a > b || (a == b && sidx[i] > sidx[ixj])) : (a < b || (a == b && sidx[i] < sidx[ixj]));
            if (sw) { skey[i] = b; skey[ixj] = a; const int t = sidx[i]; sidx[i] = sidx[ixj]; sidx[ixj] = t; }
          }
        }
        __syncthreads();
      }
    if (use_topk) {
      // TopKLogitsWarper: remove scores < k-th largest (ties with the threshold stay)
      const float kth = skey[V - p.top_k];
      __syncthreads();
      for (int v = threadIdx.x; v < V; v += blockDim.x) {
        if (lg[v] < kth) lg[v] = -INFINITY;
        if (skey[v] < kth) skey[v] = -INFINITY;
      }
      __syncthreads();
    }
    if (p.top_p < 1.0f) {
      // TopPLogitsWarper: ascending order; remove while cumulative prob <= 1 - top_p; never the last (largest) one
      float z2 = 0.f;
      for (int v = threadIdx.x; v < V; v += blockDim.x) z2 += (lg[v] == -INFINITY) ? 0.f : expf(lg[v] - mx);
      z2 = block_sum(z2, sh);
      // sequential fp64 cumulative sum over the ascending order (torch CPU cumsum accumulates float in fp64)
      if (threadIdx.x == 0) {
        double cum = 0.0;
        const float thr = 1.0f - p.top_p;
        for (int i = 0; i < V - 1; ++i) {
          const float pr = (skey[i] == -INFINITY) ? 0.f : expf(skey[i] - mx) / z2;
          cum += (double)pr;
          if ((float)cum <= thr) lg[sidx[i]] = -INFINITY; else break;
        }
      }
      __syncthreads();
    }
  }
  if (turbo) {
    // RepetitionPenaltyLogitsProcessor comes last in inference_turbo; the softmax below needs the new maximum
    float m2 = -INFINITY;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      float l = lg[v];
      if (seen[v] && rp != 1.0f) { l = (l < 0.f) ? l * rp : l / rp; lg[v] = l; }
      m2 = fmaxf(m2, l);
    }
    mx = block_max(m2, sh);
    __syncthreads();
  }
  // 6: softmax + sample
  float z3 = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) z3 += (lg[v] == -INFINITY) ? 0.f : expf(lg[v] - mx);
  z3 = block_sum(z3, sh);
  float best = -1.f; int besti = SV;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float pr = (lg[v] == -INFINITY) ? 0.f : expf(lg[v] - mx) / z3;
    float q;
    if (p.q_noise) q = p.q_noise[((long)step * p.n_utts + utt) * SV + v];
    else {
      const uint32_t u = philox_mix(p.seed, (uint32_t)utt, (uint32_t)step, (uint32_t)v);
      q = -logf(((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f));
    }
    const float sc = pr / q;
    if (sc > best || (sc == best && v < besti)) { best = sc; besti = v; }
  }
  // block argmax (first index wins ties, like torch.argmax)
  __shared__ float bv[32]; __shared__ int bi[32];
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = besti; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = bv[threadIdx.x]; besti = bi[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if (threadIdx.x == 0) bi[0] = besti;
  }
  __syncthreads();
  int tok = bi[0];
  if (p.force_tokens) {           // teacher forcing (parity tests): record the engine's own pick, feed the given id
    if (threadIdx.x == 0 && p.sampled_out) p.sampled_out[(long)utt * p.max_tokens + step] = tok;
    tok = p.force_tokens[(long)utt * p.max_tokens + step];
  }
  // 7: bookkeeping + next embedding (speech_emb[tok] + speech_pos[step+1], both CFG rows)
  // inference_turbo samples its first token from the prefill without an EOS check (t3.py:428-433)
  const bool finished = (tok == p.eos_id && !(turbo && step == 0)) || (step + 1 >= p.max_new[utt]);
  if (threadIdx.x == 0) {
    p.tokens[(long)utt * p.max_tokens + step] = tok;
    p.n_gen[utt] = step + 1;
    if (turbo && step == 0) seen[p.bos_id] = 0;     // from the second token on the history is the generated ids only
    seen[tok] = 1;
    if (finished) p.done[utt] = 1;
    for (int r = 0; r < rows_per; ++r) {
      const int row = utt * rows_per + r;
      p.positions[row] = p.base_pos[row] + step;
    }
  }
  const float* e = p.speech_emb + (long)tok * 1024;
  // next input: speech_emb[tok] + speech_pos_emb[step+1] (t3.py:371-372); Turbo: + wpe[absolute position] (GPT2Model)
  const float* pe = turbo ? p.wpe + (long)(p.base_pos[utt * rows_per] + step) * 1024 : p.speech_pos + (long)(step + 1) * 1024;
  for (int d = threadIdx.x; d < 1024; d += blockDim.x) {
    const float v = e[d] + pe[d];
    for (int r = 0; r < rows_per; ++r) p.x[((long)(j * rows_per + r)) * 1024 + d] = v;
  }
}
constexpr int T3_SAMPLE_SMEM = (SV + 64 + SV_PAD) * 4 + SV_PAD * 4;
void t3_sample_init() {     // per device, before any launch / stream capture
  CBX_CHECK(cudaFuncSetAttribute(t3_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T3_SAMPLE_SMEM));
}
void t3_sample(Ctx& ctx, const T3SampleDev& p, int n_act) {
  if (ctx.dry || n_act == 0) return;
  const int smem = T3_SAMPLE_SMEM;
  ctx.launches++;
  launch_kernel(ctx, t3_sample_kernel, dim3(n_act), dim3(1024), (size_t)smem, p);
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// T3 decode: device-side retirement of finished utterances (replaces the host sync of t3.py:366).
// Stable compaction of the active list, once per step, one CTA: act_utt[0, n_act) loses the utterances whose `done`
// flag the sampler set in the previous step; src_slot[j] = slot the survivor had before (its logits still live
// there), slot_row[j*rp + r] = physical KV row, m_live = live decode rows (GEMM row tiles above it exit).
// ================================================================================================
__global__ void __launch_bounds__(1024) t3_compact_kernel(int* act_utt, int* n_act, int* src_slot, int* slot_row, int* m_live,
                                                          const int* done, int rows_per) {
  __shared__ int wsum[32];
  __shared__ int s_base;
  pdl_wait();
  pdl_launch_dependents();
  const int n = *n_act;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 1024) {
    const int j = c0 + threadIdx.x;
    const int utt = j < n ? act_utt[j] : -1;
    const int keep = (utt >= 0 && !done[utt]) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int wpre = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();                       // every entry of this chunk is in a register before anyone writes
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    int total = 0;
    for (int w = 0; w < 32; ++w) total += wsum[w];
    const int base = s_base;
    if (keep) {
      const int d = base + woff + wpre;    // d <= j: only positions that have already been read are overwritten
      act_utt[d] = utt; src_slot[d] = j;
      for (int r = 0; r < rows_per; ++r) slot_row[d * rows_per + r] = utt * rows_per + r;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) { *n_act = s_base; *m_live = s_base * rows_per; }
}
void t3_compact(Ctx& ctx, int* act_utt, int* n_act, int* src_slot, int* slot_row, int* m_live, const int* done, int rows_per) {
  if (ctx.dry) return;
  ctx.launches++;
  launch_kernel(ctx, t3_compact_kernel, dim3(1), dim3(1024), 0, act_utt, n_act, src_slot, slot_row, m_live, done, rows_per);
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// resid_norm: the glue between two GEMMs of a decode layer in ONE pass over the row (dim <= 1024):
//   x[r] += bias + sum_z part[z][r]      (split-K partial sums of the o / down projection, fixed order -> deterministic)
//   y[r]  = norm(x[r]) * w (+ b)         RMSNorm (Llama) or LayerNorm (GPT-2), written as the bf16 hi/lo planes the
//                                        next GEMM loads by TMA (or fp32 for the GEMV path)
// Replaces: residual add in the GEMM epilogue + rmsnorm kernel.  Bytes: (nsplit + 2) x 4 KB in, 4 KB (+ 4 KB) out per row.
// ================================================================================================
__global__ void __launch_bounds__(256) resid_norm_kernel(const ResidNormDev p) {
  __shared__ float sh[32];
  const int r = blockIdx.x;
  pdl_wait();
  pdl_launch_dependents();
  if (p.m_live && r >= *p.m_live) return;
  const int i = threadIdx.x * 4;
  const bool on = i < p.dim;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  float* xr = p.x + (long)r * p.ldx;
  if (on) {
    v = *reinterpret_cast<const float4*>(xr + i);
    if (p.nsplit > 0) {
      if (p.bias) { const float4 b = *reinterpret_cast<const float4*>(p.bias + i); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
      for (int z = 0; z < p.nsplit; ++z) {
        const float4 a = *reinterpret_cast<const float4*>(p.part + (long)z * p.split_stride + (long)r * p.ldp + i);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      *reinterpret_cast<float4*>(xr + i) = v;
    }
  }
  if (!p.w) return;
  float o[4];
  if (p.layernorm) {
    const float mean = block_sum(on ? (v.x + v.y) + (v.z + v.w) : 0.f, sh) / p.dim;
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float var = block_sum(on ? (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) : 0.f, sh) / p.dim;
    const float inv = rsqrtf(var + p.eps);
    if (on) {
      const float4 w = *reinterpret_cast<const float4*>(p.w + i);
      const float4 b = *reinterpret_cast<const float4*>(p.b + i);
      o[0] = d0 * inv * w.x + b.x; o[1] = d1 * inv * w.y + b.y; o[2] = d2 * inv * w.z + b.z; o[3] = d3 * inv * w.w + b.w;
    }
  } else {
    const float ss = block_sum(on ? (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w) : 0.f, sh);
    const float inv = rsqrtf(ss / p.dim + p.eps);
    if (on) {
      const float4 w = *reinterpret_cast<const float4*>(p.w + i);
      o[0] = w.x * (v.x * inv); o[1] = w.y * (v.y * inv); o[2] = w.z * (v.z * inv); o[3] = w.w * (v.w * inv);
    }
  }
  if (!on) return;
  if (p.y16) {
    const __half2 a = __floats2half2_rn(o[0], o[1]), b = __floats2half2_rn(o[2], o[3]);
    *reinterpret_cast<uint2*>(p.y16 + (long)r * p.ldy + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  } else if (p.yhi) {
    uint32_t h0, l0, h1, l1;
    {
      __nv_bfloat16 a, b, c, d;
      split_bf16(o[0], a, b); split_bf16(o[1], c, d);
      h0 = pack_bf16(a, c); l0 = pack_bf16(b, d);
      split_bf16(o[2], a, b); split_bf16(o[3], c, d);
      h1 = pack_bf16(a, c); l1 = pack_bf16(b, d);
    }
    *reinterpret_cast<uint2*>(p.yhi + (long)r * p.ldy + i) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(p.ylo + (long)r * p.ldy + i) = make_uint2(l0, l1);
  } else {
    *reinterpret_cast<float4*>(p.y + (long)r * p.ldy + i) = make_float4(o[0], o[1], o[2], o[3]);
  }
}
void resid_norm(Ctx& ctx, const ResidNormDev& p, int rows) {
  if (ctx.dry || rows == 0) return;
  CBX_REQUIRE(p.dim <= 1024 && p.dim % 4 == 0, "resid_norm handles rows of <= 1024 floats");
  ctx.launches++;
  launch_kernel(ctx, resid_norm_kernel, dim3(rows), dim3(256), 0, p);
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// conformer encoder helpers
// ================================================================================================
// q_u = q + pos_bias_u, q_v = q + pos_bias_v  (transformer/attention.py:303-306); q lives in qkv[:, 0:512]
__global__ void add_pos_bias_kernel(const float* qkv, int ld, const float* u, const float* v, float* qu, float* qv,
                                    long rows) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 512) return;
  const long r = i >> 9; const int c = (int)(i & 511);
  const float q = qkv[r * ld + c];
  qu[r * 512 + c] = q + u[c];
  qv[r * 512 + c] = q + v[c];
}
void add_pos_bias(Ctx& ctx, const float* qkv, int ld, const float* u, const float* v, float* qu, float* qv, long rows) {
  if (ctx.dry || rows == 0) return;
  ctx.launches++;
  add_pos_bias_kernel<<<(unsigned)((rows * 512 + 255) / 256), 256, 0, ctx.stream>>>(qkv, ld, u, v, qu, qv, rows);
  CBX_CHECK(cudaGetLastError());
}
// espnet relative positional table for length T: row p (0..2T-2) <-> relative position (T-1-p)
// (transformer/embedding.py:229-258): even cols sin(pos*div), odd cols cos(pos*div)
__global__ void relpos_table_kernel(float* pe, int T, int d_model) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)(2 * T - 1) * d_model) return;
  const int prow = (int)(i / d_model), c = (int)(i - (long)prow * d_model);
  const float rel = (float)(T - 1 - prow);
  const float div = expf((float)(c & ~1) * -(logf(10000.0f) / (float)d_model));
  const float a = rel * div;
  pe[i] = (c & 1) ? cosf(a) : sinf(a);
}
void relpos_table(Ctx& ctx, float* pe, int T, int d_model) {
  if (ctx.dry) return;
  ctx.launches++;
  relpos_table_kernel<<<(unsigned)(((long)(2 * T - 1) * d_model + 255) / 256), 256, 0, ctx.stream>>>(pe, T, d_model);
  CBX_CHECK(cudaGetLastError());
}
// nearest x2 upsample between two packed layouts: out row (start2[s] + t) = in row (start1[s] + t/2)
__global__ void upsample2_kernel(const float* x, float* y, int C, const int* tile_seq2, const int* start2,
                                 const int* len2, const int* start1, long rows2) {
  const long r = blockIdx.x;
  const int s = tile_seq2[r / kTileM];
  float* yr = y + r * C;
  bool valid = s >= 0;
  int t = 0;
  if (valid) { t = (int)(r - start2[s]); valid = t < len2[s]; }
  if (!valid) { for (int c = threadIdx.x; c < C; c += blockDim.x) yr[c] = 0.f; return; }
  const float* xr = x + ((long)start1[s] + (t >> 1)) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) yr[c] = xr[c];
}
void upsample2(Ctx& ctx, const float* x, float* y, int C, const int* tile_seq2, const int* start2, const int* len2,
               const int* start1, long rows2) {
  if (ctx.dry || rows2 == 0) return;
  ctx.launches++;
  upsample2_kernel<<<(unsigned)rows2, 128, 0, ctx.stream>>>(x, y, C, tile_seq2, start2, len2, start1, rows2);
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// CFM solver helpers
// ================================================================================================
// sinusoidal embedding (matcha/decoder.py:20-29): emb[k] = scale*t*exp(-k*ln(1e4)/(half-1)); out=[sin | cos]
__global__ void time_sinusoid_kernel(const float* t, float* out, int n, int dim, float scale) {
  const int i = blockIdx.x, half = dim / 2;
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float f = expf((float)k * -(logf(10000.0f) / (float)(half - 1)));
    const float a = scale * t[i] * f;
    out[(long)i * dim + k] = sinf(a);
    out[(long)i * dim + half + k] = cosf(a);
  }
}
void time_sinusoid(Ctx& ctx, const float* t, float* out, int n, int dim, float scale) {
  if (ctx.dry || n == 0) return;
  ctx.launches++;
  time_sinusoid_kernel<<<n, 160, 0, ctx.stream>>>(t, out, n, dim, scale);
  CBX_CHECK(cudaGetLastError());
}
// xin[row] = [x(80) | mu(80) | spk(80) | cond(80)] for the conditional half, [x | 0 | 0 | 0] for the uncond half.
// layout3 has 2B sequences: s < B conditional, s >= B unconditional copy of sequence s-B (layout2 rows).
__global__ void cfm_assemble_kernel(float* xin, const float* x, const float* mu, const float* spk, const float* cond,
                                    const int* tile_seq3, const int* start3, const int* len3, const int* start2, int B,
                                    int write_static) {
  const long r = blockIdx.x;
  const int s = tile_seq3[r / kTileM];
  float* o = xin + r * 320;
  bool valid = s >= 0; int t = 0;
  if (valid) { t = (int)(r - start3[s]); valid = t < len3[s]; }
  if (!valid) { for (int c = threadIdx.x; c < 320; c += blockDim.x) o[c] = 0.f; return; }
  const int sb = s < B ? s : s - B;
  const long r2 = (long)start2[sb] + t;
  for (int c = threadIdx.x; c < 80; c += blockDim.x) o[c] = x[r2 * 80 + c];
  if (write_static) {
    for (int c = threadIdx.x; c < 80; c += blockDim.x) {
      o[80 + c] = s < B ? mu[r2 * 80 + c] : 0.f;
      o[160 + c] = s < B ? spk[(long)sb * 80 + c] : 0.f;
      o[240 + c] = s < B ? cond[r2 * 80 + c] : 0.f;
    }
  }
}
void cfm_assemble(Ctx& ctx, float* xin, const float* x, const float* mu, const float* spk, const float* cond,
                  const int* tile_seq3, const int* start3, const int* len3, const int* start2, int B, long rows3,
                  int write_static) {
  if (ctx.dry || rows3 == 0) return;
  ctx.launches++;
  cfm_assemble_kernel<<<(unsigned)rows3, 96, 0, ctx.stream>>>(xin, x, mu, spk, cond, tile_seq3, start3, len3, start2, B,
                                                            write_static);
  CBX_CHECK(cudaGetLastError());
}
// x += dt * ((1+w) v_c - w v_u)   (flow_matching.py:138-141);  meanflow/no-CFG: x += dt * v_c
__global__ void cfm_euler_kernel(float* x, const float* v, const int* tile_seq2, const int* start2, const int* len2,
                                 const int* start3, int B, float dt, float w, int cfg, long rows2) {
  const long r = blockIdx.x;
  const int s = tile_seq2[r / kTileM];
  if (s < 0) return;
  const int t = (int)(r - start2[s]);
  if (t >= len2[s]) return;
  const long rc = (long)start3[s] + t, ru = cfg ? (long)start3[s + B] + t : 0;
  for (int c = threadIdx.x; c < 80; c += blockDim.x) {
    const float vc = v[rc * 80 + c];
    float d = vc;
    if (cfg) d = (1.0f + w) * vc - w * v[ru * 80 + c];
    x[r * 80 + c] = x[r * 80 + c] + dt * d;
  }
}
void cfm_euler(Ctx& ctx, float* x, const float* v, const int* tile_seq2, const int* start2, const int* len2,
               const int* start3, int B, float dt, float w, int cfg, long rows2) {
  if (ctx.dry || rows2 == 0) return;
  ctx.launches++;
  cfm_euler_kernel<<<(unsigned)rows2, 96, 0, ctx.stream>>>(x, v, tile_seq2, start2, len2, start3, B, dt, w, cfg, rows2);
  CBX_CHECK(cudaGetLastError());
}

// ================================================================================================
// HiFT helpers
// ================================================================================================
// harmonic-plus-noise source (hifigan.py SineGen.forward :200-231 + SourceModuleHnNSF :267-283)
// Phase: torch.cumsum on CPU accumulates float32 sequentially in fp64 and rounds every output to fp32
// (ATen cpu_cum_base_kernel, acc_type<float> = double); the '% 1' comes after.  The engine replays that scan EXACTLY, but
// not sample by sample: inside one mel frame the increment d = (double)(f0 * h / 24000) is constant for 480 samples, so
//     c_k = c_0 + k * d   holds bit for bit whenever no addition of the frame rounds,
// i.e. when c_0 and d are both multiples of u = ulp(c_0 + 480 d): every partial sum is then a multiple of u below 2^53 u.
// hift_phase_frames_kernel walks the FRAMES of one (sequence, harmonic) sequentially (a few thousand steps instead of
// ~10^6 dependent fp64 additions), records c_0 and the exact / inexact flag per frame and falls back to the 480 sequential
// additions only for a frame that does round (binade crossings, increments with bits below the running sum's ulp);
// hift_source_kernel then evaluates c_0 + (k+1) d per sample (or replays the frame's additions when flagged).  The round-2
// sample-by-sample kernel was 33 % of the HiFT stage (profiles/r2_launches_hift.txt) and wrote / re-read 9 floats per sample.
struct PhaseTab { double* c0; unsigned char* inexact; };   // [frame row][9]
__device__ __forceinline__ bool frame_adds_exact(double c, double d) {
  if (d == 0.0) return true;
  const double cend = c + 480.0 * d;                                   // >= d > 0, normal
  const int e = ((__double2hiint(cend) >> 20) & 0x7ff) - 1023;         // ilogb(cend); a rounded-up cend only makes u coarser
  const double u_inv = __hiloint2double((1023 + 52 - e) << 20, 0);     // 1 / ulp(cend)
  const double a = c * u_inv, b = d * u_inv;
  return a == rint(a) && b == rint(b);
}
__global__ void hift_phase_frames_kernel(const float* f0, PhaseTab tab, const int* startT, const int* lenT, int n_seq) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_seq * 9) return;
  const int s = id / 9, h = id % 9;
  double c = 0.0;
  const long r0 = startT[s];
  for (int tau = 0; tau < lenT[s]; ++tau) {
    const float inc = f0[r0 + tau] * (float)(h + 1) / 24000.0f;        // F_mat value (fp32)
    const double d = (double)inc;
    const bool ex = frame_adds_exact(c, d);
    tab.c0[(r0 + tau) * 9 + h] = c;
    tab.inexact[(r0 + tau) * 9 + h] = ex ? 0 : 1;
    if (ex) c = c + 480.0 * d;
    else for (int k = 0; k < 480; ++k) c += d;
  }
}
__global__ void __launch_bounds__(480) hift_source_kernel(const float* f0, PhaseTab tab, const float* phase_vec,
                                                          const float* noise, const float* lin_w, float lin_b,
                                                          float* s_out, const int* startT, const int* lenT,
                                                          const long* startS, int n_seq, unsigned long long seed) {
  const int s = blockIdx.y, tau = blockIdx.x;
  if (tau >= lenT[s]) return;
  const int o = threadIdx.x;
  const float f = f0[(long)startT[s] + tau];
  const long L = (long)lenT[s] * 480;
  const long n = (long)tau * 480 + o;
  const float uv = f > 10.0f ? 1.0f : 0.0f;
  const float namp = uv * 0.003f + (1.0f - uv) * 0.1f / 3.0f;
  float acc = lin_b;
#pragma unroll
  for (int h = 0; h < 9; ++h) {
    const long ti = ((long)startT[s] + tau) * 9 + h;
    const double d = (double)(f * (float)(h + 1) / 24000.0f);
    double cacc = tab.c0[ti];
    if (!tab.inexact[ti]) cacc = cacc + (double)(o + 1) * d;              // exact: the product has <= 33 significant bits
    else for (int k = 0; k <= o; ++k) cacc += d;                          // this frame rounds: replay its additions
    const float cf = (float)cacc;
    const float frac = cf - floorf(cf);                                     // torch '% 1' on non-negative values
    const float theta = 6.283185307179586f * frac;
    const float ph = phase_vec ? phase_vec[s * 9 + h] : 0.f;
    float sw = 0.1f * sinf(theta + ph);
    float nz;
    if (noise) nz = noise[startS[s] * 9 + (long)h * L + n];
    else {
      const uint32_t u1 = philox_mix(seed, (uint32_t)s, (uint32_t)h, (uint32_t)(2 * n));
      const uint32_t u2 = philox_mix(seed, (uint32_t)s, (uint32_t)h, (uint32_t)(2 * n + 1));
      const float a = ((float)(u1 >> 8) + 0.5f) * (1.0f / 16777216.0f), b = ((float)(u2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
      nz = sqrtf(-2.0f * logf(a)) * cosf(6.283185307179586f * b);
    }
    sw = sw * uv + namp * nz;
    acc += sw * lin_w[h];
  }
  s_out[startS[s] + n] = tanhf(acc);
}
void hift_source(Ctx& ctx, const float* f0, float* cumf, const float* phase_vec, const float* noise, const float* lin_w,
                 float lin_b, float* s_out, const int* startT, const int* lenT, const long* startS, int n_seq, int maxT,
                 unsigned long long seed, long n_frame_rows) {
  if (ctx.dry || n_seq == 0) return;
  ctx.launches += 2;
  // the per-frame table lives in the caller's `cumf` scratch (sized for 9 floats per SAMPLE: far more than 9 x 9 bytes per frame)
  PhaseTab tab;
  tab.c0 = reinterpret_cast<double*>(cumf);
  tab.inexact = reinterpret_cast<unsigned char*>(tab.c0 + (size_t)n_frame_rows * 9);
  hift_phase_frames_kernel<<<(n_seq * 9 + 31) / 32, 32, 0, ctx.stream>>>(f0, tab, startT, lenT, n_seq);
  hift_source_kernel<<<dim3(maxT, n_seq), 480, 0, ctx.stream>>>(f0, tab, phase_vec, noise, lin_w, lin_b, s_out, startT,
                                                               lenT, startS, n_seq, seed);
  CBX_CHECK(cudaGetLastError());
}
// |Linear(512->1)| per frame (f0_predictor.py:52-55): one warp per row
__global__ void f0_head_kernel(const float* x, int ld, const float* w, float b, float* f0, long rows) {
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int i = lane; i < 512; i += 32) acc += x[r * ld + i] * w[i];
  acc = warp_sum(acc);
  if (lane == 0) f0[r] = fabsf(acc + b);
}
void f0_head(Ctx& ctx, const float* x, int ld, const float* w, float b, float* f0, long rows) {
  if (ctx.dry || rows == 0) return;
  ctx.launches++;
  f0_head_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx.stream>>>(x, ld, w, b, f0, rows);
  CBX_CHECK(cudaGetLastError());
}
// STFT(n_fft 16, hop 4, hann, center/reflect) of the source: out[frame][0..8]=Re, [9..17]=Im (hifigan.py:396-402)
__global__ void hift_stft_kernel(const float* s, float* out, const int* startF, const int* lenT, const long* startS,
                                 int n_seq) {
  const int sq = blockIdx.y;
  const long F = (long)lenT[sq] * 120 + 1, L = (long)lenT[sq] * 480;
  const long f = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (f >= F) return;
  const int k = threadIdx.x & 31;
  if (k >= 9) return;
  const float* sp = s + startS[sq];
  float re = 0.f, im = 0.f;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    long idx = f * 4 + n - 8;
    if (idx < 0) idx = -idx;                      // reflect
    if (idx >= L) idx = 2 * (L - 1) - idx;
    const float w = 0.5f - 0.5f * cospif((float)n / 8.0f);   // periodic hann(16)
    const float v = w * sp[idx];
    float sn, cs;
    sincospif((float)((k * n) & 15) / 8.0f, &sn, &cs);
    re += v * cs; im -= v * sn;
  }
  float* o = out + ((long)startF[sq] + f) * 18;
  o[k] = re; o[9 + k] = im;
}
void hift_stft(Ctx& ctx, const float* s, float* out, const int* startF, const int* lenT, const long* startS, int n_seq,
               int maxT) {
  if (ctx.dry || n_seq == 0) return;
  ctx.launches++;
  const long maxF = (long)maxT * 120 + 1;
  hift_stft_kernel<<<dim3((unsigned)((maxF + 7) / 8), n_seq), 256, 0, ctx.stream>>>(s, out, startF, lenT, startS, n_seq);
  CBX_CHECK(cudaGetLastError());
}
// reflection pad (1,0) of the last upsampling stage (hifigan.py:421-422): rows were written at +1; row0 = row2
__global__ void reflect_row0_kernel(float* x, int C, const int* start, int n_seq) {
  const int s = blockIdx.x;
  float* base = x + (long)start[s] * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) base[c] = base[2 * C + c];
}
void reflect_row0(Ctx& ctx, float* x, int C, const int* start, int n_seq) {
  if (ctx.dry || n_seq == 0) return;
  ctx.launches++;
  reflect_row0_kernel<<<n_seq, 64, 0, ctx.stream>>>(x, C, start, n_seq);
  CBX_CHECK(cudaGetLastError());
}
// conv_post output [F][18] -> magnitude/phase -> iSTFT(16,4,hann) -> clamp +-0.99 -> trim-fade (s3gen.py:254-258)
__global__ void hift_istft_kernel(const float* y, float* wav, const int* startF, const int* lenT, const long* startS,
                                  int trim_fade) {
  const int sq = blockIdx.y;
  const long L = (long)lenT[sq] * 480, F = (long)lenT[sq] * 120 + 1;
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;     // output sample
  if (n >= L) return;
  const long np = n + 8;                                          // position in the padded signal
  float num = 0.f, den = 0.f;
  for (long f = np / 4; f >= 0 && f * 4 + 15 >= np; --f) {
    if (f >= F) continue;
    const int m = (int)(np - f * 4);
    const float* yf = y + ((long)startF[sq] + f) * 18;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float mag = expf(yf[k]);
      mag = fminf(mag, 100.0f);
      const float ph = sinf(yf[9 + k]);
      float sp, cp; sincosf(ph, &sp, &cp);
      const float re = mag * cp, im = mag * sp;
      float sn, cs; sincospif((float)((k * m) & 15) / 8.0f, &sn, &cs);
      const float ck = (k == 0 || k == 8) ? 1.0f : 2.0f;
      acc += ck * (re * cs - ((k == 0 || k == 8) ? 0.f : im * sn));
    }
    const float w = 0.5f - 0.5f * cospif((float)m / 8.0f);
    num += w * (acc / 16.0f);
    den += w * w;
  }
  float v = den > 1e-11f ? num / den : num;
  v = fminf(fmaxf(v, -0.99f), 0.99f);
  if (trim_fade && n < 960) {
    float fd = 0.f;
    if (n >= 480) fd = (cosf(3.14159265358979323846f * (1.0f - (float)(n - 480) / 479.0f)) + 1.0f) * 0.5f;
    v *= fd;
  }
  wav[startS[sq] + n] = v;
}
void hift_istft(Ctx& ctx, const float* y, float* wav, const int* startF, const int* lenT, const long* startS, int n_seq,
                int maxT, int trim_fade) {
  if (ctx.dry || n_seq == 0) return;
  ctx.launches++;
  const long maxL = (long)maxT * 480;
  hift_istft_kernel<<<dim3((unsigned)((maxL + 255) / 256), n_seq), 256, 0, ctx.stream>>>(y, wav, startF, lenT, startS,
                                                                                      trim_fade);
  CBX_CHECK(cudaGetLastError());
}

}  // namespace cbx
