// Host-side op layer of libcbx: packed weights, workspace arena, kernel launch wrappers.
#pragma once
#include "common.cuh"
#include <map>
#include <cstring>
#include <vector>
#include <string>

namespace cbx {

// ----------------------------------------------------------------------------------------------
// Packed weight: W[Npad][Kpad] bf16, K-major (exactly torch's Linear [out, in] / conv [out, tap*Cin_pad]).
// ----------------------------------------------------------------------------------------------
struct Weight {
  __nv_bfloat16* w = nullptr;
  float* bias = nullptr;     // [Npad] fp32 (zeros when the layer has none)
  int N = 0, K = 0;          // logical sizes
  int Npad = 0, Kpad = 0;    // N padded to 64, K padded to 64
  CUtensorMap tmap[3];       // TMA maps with box {64 (K), 64|128|256 (N)}, SWIZZLE_128B
  __half* w16 = nullptr;     // optional fp16 copy of the same values (exact for |w| in [2^-14, 65504]): operand of the
  CUtensorMap tmap16[3];     // single-plane fp16-activation GEMMs -- tcgen05 kind::f16 wants A and B in ONE 16-bit format
};

// Geometry of a packed variable-length batch.  Every sequence starts at a multiple of kTileM rows, so
// one 128-row tile never straddles two sequences.  All arrays live on the device.
struct SeqMap {
  const int* tile_seq = nullptr;   // [n_tiles] sequence id of each output tile
  const int* out_start = nullptr;  // [n_seq] first output row of the sequence
  const int* out_len = nullptr;    // [n_seq] valid output rows
  const int* in_start = nullptr;   // [n_seq] first input row (input buffer may use another layout)
  const int* in_len = nullptr;     // [n_seq] valid input rows
};

enum AMode : int { A_TAPS = 0, A_WINDOW = 1 };

// Device-visible argument block of every GEMM flavour (tcgen05 tiles, SIMT tiles, GEMV).
struct GemmDev {
  // ---- A operand: fp32 activations gathered as an implicit im2col -----------------------------
  const float* A;
  int lda;        // floats per input row
  int M;          // output rows (grid covers ceil(M/128) tiles)
  int M_in;       // input rows (flat case bound)
  int a_mode;     // A_TAPS: k = tap*ctap + c, input row = in_row0 + tap*dil, needs ctap % 64 == 0
                  // A_WINDOW: element k of output row = flat input element in_row0*c_in + k (lda == c_in)
  int ntaps, ctap, c_in, dil, pad, stride;
  int k_total;    // logical K (ntaps*c_in for WINDOW, ntaps*ctap for TAPS)
  int has_seq;
  SeqMap seq;
  int a_tma;      // set by the launcher: A tile is fetched by TMA (plain strided fp32 rows)
  // ---- B operand -------------------------------------------------------------------------------
  const __nv_bfloat16* Wp;  // [Npad][Kpad]
  int Kpad, Npad;
  // ---- epilogue --------------------------------------------------------------------------------
  float* C; int ldc;
  int n_out;                // valid output columns
  const float* bias;        // [Npad] or null
  float alpha;              // v = acc*alpha + bias
  int act; float act_p; const float* act_vec;   // activation (+ per-channel parameter, e.g. snake alpha)
  const float* res; int ldr;                    // v += res[r][n]
  int accumulate;                               // v += C[r][n] (old value)
  float out_scale;                              // v *= out_scale (after act/res)
  int swiglu;                                   // columns (2j,2j+1) -> out[j] = silu(v0)*v1
  float* C2; int ldc2; int act2; float act2_p; const float* act2_vec;  // optional 2nd output act2(v)
  __nv_bfloat16* Chi; __nv_bfloat16* Clo; int ldcb;   // optional bf16 hi/lo planes of v (C may then be null)
  long long* dbg;   // optional [64] clock64 timestamps of CTA (0,0) (kernel anatomy debugging)
  // A operand already split into bf16 hi/lo planes [M][ldab] by the producing kernel (Linear only): both planes are
  // fetched by TMA straight into the UMMA layout and the converter warps have nothing to do in the main loop
  const __nv_bfloat16* Ahi; const __nv_bfloat16* Alo; int ldab;
  // Chi-only epilogue: store ONE fp16 value per element into Chi (reinterpreted as __half[M][ldcb]) instead of the bf16
  // hi/lo pair (operand format of the single-term tcgen05 attention)
  int c_half;
  // A operand as ONE fp16 plane [M][lda16] written by the producer (plain Linear): a single TMA box per K block and a
  // single MMA term (A fp16 x W bf16).  Only where the precision study allows it (CFM transformer blocks).
  const __half* A16; int lda16;
  // split-K (decode GEMMs with few row tiles): grid.z = splitk CTAs per output tile, raw fp32 partial sums at
  // C + z * split_stride (elements); the consumer (resid_norm) adds them in a fixed order
  int splitk; long split_stride;
  const int* m_live;        // optional device scalar: row tiles with m0 >= *m_live exit at once (device-side retirement)
  int tile_bn;              // 0 = heuristic; 64 | 128 | 256 forces the N tile (decode shapes are tuned by measurement)
  int tile_dual;            // with tile_bn: 1 = the two-CTAs-per-SM configuration
  // GEMV path only (<= 8 rows, B=1 latency): A is the RAW residual stream and the kernel applies the row norm itself
  // (RMSNorm: norm_w * (x * rsqrt(mean x^2 + eps)); LayerNorm when norm_ln) -- one launch less per projection
  const float* norm_w; const float* norm_b; int norm_ln; float norm_eps;
  int epi_direct;           // set by the launcher: bit 0 = register-direct epilogue, bit 1 = 32-byte C stores, bit 2 = 32-byte residual loads, bit 3 = TMA-store of fp32 rows
};

struct Arena {  // bump allocator over caller-owned workspace; dry=true only counts
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  void* alloc(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    if (off > peak) peak = off;
    if (dry) return reinterpret_cast<void*>(size_t(0x1000) + a);  // fake non-null pointer, never dereferenced
    if (off > cap) throw std::runtime_error("cbx: workspace too small");
    return base + a;
  }
  template <typename T> T* get(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
  size_t mark() const { return off; }
  void reset(size_t m) { off = m; }
};

// Optional CUDA-event timer around one class of kernel launches (bench.py roofline: events on the launching stream).
enum KClass : int { K_NONE = 0, K_GEMM_TC = 1, K_GEMV = 2, K_FLASH = 3, K_PAGED = 4, K_WRES = 5, K_STREAM = 6, K_ATTN_TC = 7,
                    K_HIFT_CONV = 8, K_NCLASS = 9, K_ALL = 99 };
// One class (cbx_set_option "time_kernel" = its name) or every class at once ("all"): each timed launch is bracketed by a
// pair of events tagged with its class; drain() adds the elapsed times up per class.
struct KTimer {
  int cls = K_NONE;
  std::vector<cudaEvent_t> ev;   // pairs (start, stop)
  std::vector<int> tag;          // class of each pair
  size_t used = 0;
  double ms_c[K_NCLASS] = {}; long long n_c[K_NCLASS] = {};
  double work_c[K_NCLASS] = {};  // algorithmic work of the timed launches (flops), added by the launcher
  double bytes_c[K_NCLASS] = {}; // algorithmic HBM bytes of the timed launches (operands once in, results once out)
  bool on(int c) const { return cls == c || cls == K_ALL; }
  void add(int c, double work, double bytes) { if (on(c)) { work_c[c] += work; bytes_c[c] += bytes; } }
  void begin(int c, cudaStream_t st) {
    if (!on(c)) return;
    if (used + 2 > ev.size()) {
      if (ev.size() >= 32768) { drain(); }
      else { size_t old = ev.size(); ev.resize(old + 2048); tag.resize(ev.size() / 2 + 1); for (size_t i = old; i < ev.size(); ++i) cudaEventCreate(&ev[i]); }
    }
    tag[used / 2] = c;
    cudaEventRecord(ev[used], st);
  }
  void end(int c, cudaStream_t st) { if (!on(c)) return; cudaEventRecord(ev[used + 1], st); used += 2; }
  void drain() {
    for (size_t i = 0; i + 1 < used; i += 2) {
      cudaEventSynchronize(ev[i + 1]);
      float t = 0.f; cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      ms_c[tag[i / 2]] += t; n_c[tag[i / 2]] += 1;
    }
    used = 0;
  }
  void reset() { drain(); for (int c = 0; c < K_NCLASS; ++c) { ms_c[c] = 0; n_c[c] = 0; work_c[c] = 0; bytes_c[c] = 0; } }
  double ms() const { double t = 0; for (int c = 0; c < K_NCLASS; ++c) t += ms_c[c]; return t; }
  long long n() const { long long t = 0; for (int c = 0; c < K_NCLASS; ++c) t += n_c[c]; return t; }
  double work() const { double t = 0; for (int c = 0; c < K_NCLASS; ++c) t += work_c[c]; return t; }
  double bytes() const { double t = 0; for (int c = 0; c < K_NCLASS; ++c) t += bytes_c[c]; return t; }
};

struct Ctx {
  KTimer* timer = nullptr;
  cudaStream_t stream = nullptr;
  Arena ws;
  bool dry = false;       // size-only pass: no launches
  int gemm_impl = 0;      // 0 = tcgen05 (default), 1 = SIMT reference tiles (debug, env CBX_GEMM=simt)
  int attn_impl = 0;      // 0 = tensor-core flash (default), 1 = SIMT reference (debug, env CBX_ATTN=simt)
  int attn_f16 = 0;       // CFM attention operands: 0 = bf16 hi/lo planes, 3 MMA terms (default), 1 = one fp16 plane, 1 term
  int cfm_act_f16 = 0;    // CFM transformer-block GEMM inputs: 0 = bf16 hi/lo planes, 2 terms (default), 1 = one fp16 plane
  long launches = 0;      // kernels launched through this context
  bool pdl = false;       // launch with programmatic stream serialization (decode step; kernels call pdl_wait())
};

// kernel launch through cudaLaunchKernelEx so that the PDL attribute can ride along (ctx.pdl)
template <typename... KArgs, typename... Args>
inline void launch_kernel(Ctx& ctx, void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx.stream;
  cudaLaunchAttribute at[1];
  if (ctx.pdl) {
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
  }
  CBX_CHECK(cudaLaunchKernelEx(&cfg, k, static_cast<KArgs>(args)...));
}

// ---- weights ------------------------------------------------------------------------------------
// host fp32 [N][K] (row-major) -> device packed bf16 + TMA maps.  taps/cin describe conv weights given
// as [N][cin][taps] (torch Conv1d layout); they are re-ordered to [N][tap][cin_pad].
void pack_linear(Weight& W, const float* host_w, const float* host_bias, int N, int K, bool half_copy = false);
void pack_conv_taps(Weight& W, const float* host_w, const float* host_bias, int N, int cin, int taps);    // k = tap*ctap + c
void pack_conv_window(Weight& W, const float* host_w, const float* host_bias, int N, int cin, int taps);  // k = tap*cin + c
void free_weight(Weight& W);
void make_tmaps_for(Weight& W);   // (re)build the TMA maps of a weight whose .w/.Npad/.Kpad are set

// ---- GEMM ---------------------------------------------------------------------------------------
void umma_rowshift_probe(Ctx& ctx, const __nv_bfloat16* A, const __nv_bfloat16* W, int shift, int mode, float* C);   // probe.cu
void gemm_init();   // per device: dynamic shared memory opt-in of every tcgen05 instantiation
GemmDev gemm_args_linear(const float* A, int lda, int M, const Weight& W, float* C, int ldc);
void gemm(Ctx& ctx, GemmDev g, const Weight& W);

// ---- attention ------------------------------------------------------------------------------------
struct AttnArgs {
  const float* Q; const float* K; const float* V;  // row-major [rows][ld*], head h at column h*64
  int ldq, ldk, ldv;
  float* O; int ldo;
  int n_seq, n_heads;
  const int* q_start; const int* q_len;     // per sequence (device)
  const int* kv_start; const int* kv_len;
  int max_q_len;                            // host-side bound for the grid
  float scale;                              // S = (Q K^T + bias) * scale
  int causal;                               // query i attends keys j <= i + (kv_len - q_len)
  const float* bias = nullptr;              // optional additive bias, indexed by packed q row
  long bias_head_stride = 0; int bias_ld = 0; long bias_row0 = 0;
  int bias_rel = 0; int bias_center = 0;    // bias_rel: column = bias_center - i + j (espnet rel_shift)
};
void attention(Ctx& ctx, const AttnArgs& a);
// tcgen05 variant (non-causal, no bias): operands are bf16 hi/lo planes [rows][ld] addressed through TMA maps
struct AttnTcArgs {
  const CUtensorMap* tm_hi; const CUtensorMap* tm_lo;
  int q_col, k_col, v_col;                    // column of head 0 of Q / K / V inside the planes
  float* O; int ldo;
  __nv_bfloat16* Ohi = nullptr; __nv_bfloat16* Olo = nullptr;   // optional: write the output as bf16 planes [rows][ldo]
  int f16 = 0;              // 1: Q/K/V are ONE fp16 plane (tm_hi maps it), single-term products (see attn_tc.cu)
  __half* O16 = nullptr;    // optional: write the output as one fp16 plane [rows][ldo]
  int n_seq, n_heads;
  const int* q_start; const int* q_len; const int* kv_start; const int* kv_len;
  int max_q_len; float scale;
  double work = 0.0;        // algorithmic flops of this launch (4 * 64 * heads * sum len^2), for the event timer only
};
void attention_tc(Ctx& ctx, const AttnTcArgs& a);
void attention_tc_init();
void make_plane_tmap(CUtensorMap* tm, const __nv_bfloat16* base, long rows, int cols, int box_rows = 64, int ld = 0);
void attention_generic(Ctx& ctx, const float* Q, const float* K, const float* V, float* O, int n_q, int n_kv,
                       int n_heads, int head_dim, int ldq, int ldk, int ldv, int ldo, float scale);

// ---- T3 paged KV cache ------------------------------------------------------------------------------
struct PagedKV {
  void* pages;            // [n_layers][n_pages][2][n_heads][page_tokens][64] of kv dtype (layer-major)
  int n_pages;
  int kv_fp32;            // element type of the cache: 0 = bf16, 1 = fp32, 2 = fp8 e4m3 (opt-in, bulk-copy kernel only)
  int n_layers, n_heads, page_tokens;
  const int* page_table;  // [rows][max_pages_per_row]
  int max_pages_per_row;
};
struct PagedOpts {
  int fuse_rope = 0;                  // qkv holds the raw projections: rotate q/k, append k/v, attend (bulk kernel only)
  const float* cos_t = nullptr; const float* sin_t = nullptr;   // [pos][32] RoPE tables
  const int* n_live = nullptr;        // device scalar: slots >= *n_live exit (device-side retirement)
  int impl = 0;                       // 1 = round-1 __ldg kernel (tests / A-B runs)
  __half* out16 = nullptr;            // write the output as one fp16 plane [slots][ldo] (fp16-activation decode mode)
};
void paged_attention_init();
void paged_decode_attention(Ctx& ctx, const float* qkv, int ldqkv, const PagedKV& kv, int layer, const int* slot_row,
                            int n_slots, const int* positions, float* out, int ldo, float* scratch, int nsplit,
                            __nv_bfloat16* out_hi = nullptr, __nv_bfloat16* out_lo = nullptr,   // planes instead of out
                            const PagedOpts* opts = nullptr);
void rope_and_store_kv(Ctx& ctx, float* qkv, int ldqkv, const PagedKV& kv, int layer, const int* tok_row,
                       const int* tok_pos, int pos_is_per_row, int n_tok, const float* cos_t, const float* sin_t);

// ---- norms / fill -------------------------------------------------------------------------------------
void rmsnorm(Ctx& ctx, const float* x, int ldx, const float* w, float* y, int ldy, int rows, int dim, float eps,
             const int* row_idx,
             __nv_bfloat16* yhi = nullptr, __nv_bfloat16* ylo = nullptr);   // yhi/ylo: bf16 planes [rows][ldy] instead of y
void layernorm(Ctx& ctx, const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int rows, int dim,
               float eps, int act, float out_scale, const float* seq_add, int seq_add_ld, const SeqMap* seq,
               __nv_bfloat16* yhi = nullptr, __nv_bfloat16* ylo = nullptr,    // yhi/ylo: bf16 planes instead of fp32 y
               __half* y16 = nullptr);                                        // y16: one fp16 plane instead
void fill(Ctx& ctx, float* p, long n, float v);

}  // namespace cbx
