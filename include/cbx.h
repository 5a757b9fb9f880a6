/* libcbx -- C ABI of the B200-native Chatterbox inference engine (sm_100a).
 *
 * The reference (resemble-ai/chatterbox) has no FFI / plugin interface: its boundary is the Python
 * module API.  This header is the C-ABI that the thin Python shim in chatterbox_b200/
 * binds with ctypes; each entry point cites the reference method whose arithmetic it replaces.
 *
 * Conventions
 *  - every function returns an int status (CBX_OK == 0); nothing throws or aborts across the boundary;
 *    cbx_last_error() returns the message of the last failure on that handle;
 *  - the caller owns every activation buffer (KV pages, workspace, inputs, outputs) and passes raw device
 *    pointers + explicit sizes + a cudaStream_t; the library owns only its packed weights;
 *  - no hidden allocation or host synchronisation on the hot path; calls are ordered by the stream;
 *  - there is no CPU fallback: without a CUDA device cbx_create() fails.
 *
 * Packed variable-length batches ("layout"): sequences are stored back to back in one [rows, C] fp32
 * channel-last buffer, each sequence starting at a multiple of 128 rows (tile_seq[row/128] names the owner).
 */
#ifndef CBX_H_
#define CBX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBX_OK 0
#define CBX_ERR_INVALID 1
#define CBX_ERR_CUDA 2
#define CBX_ERR_WORKSPACE 3

typedef struct cbx_handle cbx_handle;
typedef void* cbx_stream; /* cudaStream_t */

typedef struct {
  int n_seq;           /* sequences */
  int rows;            /* total rows incl. alignment padding (multiple of 128) */
  int max_len;         /* longest sequence (host-side bound for grids) */
  const int* tile_seq; /* device [rows/128] sequence id per tile (-1 = unused) */
  const int* start;    /* device [n_seq] first row */
  const int* len;      /* device [n_seq] valid rows */
  const int* h_start;  /* host copy of start (needed by cbx_flow_encode to chunk the rel-pos bias), may be NULL elsewhere */
  const int* h_len;    /* host copy of len */
} cbx_layout;

/* ---- lifetime / weights -------------------------------------------------------------------------
 * replaces: ChatterboxTTS.from_local() module construction + load_state_dict  (reference tts.py:133-163) */
int cbx_create(int device, cbx_handle** out);
void cbx_destroy(cbx_handle* h);
const char* cbx_last_error(cbx_handle* h);
int cbx_version(void);
/* options (all per handle):
 *   "gemm" = "tc"|"simt", "attn" = "tc"|"simt"   SIMT = plain reference kernels for bisecting (debug)
 *   "time_kernel" = class name                   see cbx_timer_read below
 *   "decode_graph" = "1"|"0"                     every decode step replays a CUDA graph captured once per (state, capacity)
 *                                                (default 1; the call must then be issued on a capturable, non-default stream)
 *   "attn_prec" = "fp16"|"bf16x3"                operand format of the CFM attention: one fp16 plane, 1 MMA term (default; mel RMS
 *                                                1.1e-5 vs the reference at T = 2040) or bf16 hi/lo planes, 3 terms (fp32-faithful)
 *   "cfm_act" = "fp16"|"bf16x2"                  operand format of the CFM transformer-block GEMM inputs: one fp16 plane against an fp16
 *                                                copy of the weights (default; mel RMS 1.5e-4, bar 1e-3) or bf16 hi/lo planes;
 *                                                "fp16" takes effect together with attn_prec = fp16 */
int cbx_set_option(cbx_handle* h, const char* key, const char* value);
/* number of kernels launched through this handle so far (bench.py's gpu_launches) */
long long cbx_launch_count(cbx_handle* h);
/* cbx_set_option(h, "time_kernel", "paged"|"gemm_tc"|"gemv"|"flash"|"none") brackets every launch of that kernel
 * class with CUDA events on the launching stream; cbx_timer_read waits for them and returns the summed device
 * time and the number of launches since the option was set, plus (GEMM classes) the algorithmic flops 2*M*N*K of
 * those launches (bench.py's roofline object) */
int cbx_timer_read(cbx_handle* h, double* ms, long long* launches, double* work);
/* algorithmic HBM bytes of the same launches (GEMM family: weights, activations and results once each) */
int cbx_timer_read_bytes(cbx_handle* h, double* bytes);
/* per-class readout after cbx_set_option(h, "time_kernel", "all" | <class>): classes gemm_tc | wres | stream | gemv | attn_tc |
 * flash | paged | hift_conv (one CUDA-event pair per launch on the launching stream) */
int cbx_timer_read_class(cbx_handle* h, const char* cls, double* ms, long long* launches, double* work, double* bytes);
/* host fp32 tensor with the reference's state-dict name ("t3." / "flow." / "hift." prefix added by the caller) */
int cbx_load_tensor(cbx_handle* h, const char* name, const float* host_data, int ndim, const int64_t* shape);
/* pack loaded tensors of one model ("t3" | "flow" | "hift"): bf16 K-major weights + TMA maps, QKV concat,
 * gate/up interleave, weight-norm folding (reference hifigan.py weight_norm parametrisations) */
int cbx_finalize_weights(cbx_handle* h, const char* model);

/* ---- T3 (reference src/chatterbox/models/t3/) ----------------------------------------------------- */
typedef struct {
  int n_utts, n_rows, cfg;    /* n_rows = n_utts * (cfg ? 2 : 1); rows (2b, 2b+1) = (cond, uncond) */
  /* paged KV cache, layer-major: pages[layer][page][k|v][head][token][64] (one layer's pages are contiguous so that
   * a decode step of that layer stays inside ~1/n_layers of the pool: TLB reach) */
  void* kv_pages; int kv_dtype; /* 0 = bf16, 1 = fp32 (parity mode), 2 = fp8 e4m3 (opt-in: halves the KV term of the decode roofline) */ int page_tokens;
  const int* page_table; int max_pages_per_row; /* device [n_rows][max_pages_per_row] */
  int n_pages;             /* pages in the pool (stride between layers) */
  int* positions;          /* device [n_rows] rope position of the token being fed */
  const int* base_pos;     /* device [n_rows] prefill length S0 */
  int* tokens; int max_tokens; /* device [n_utts][max_tokens] generated ids */
  int* n_gen;              /* device [n_utts] */
  const int* max_new;      /* device [n_utts] per-utterance budget (reference max_new_tokens) */
  int* done;               /* device [n_utts] */
  unsigned char* seen;     /* device [n_utts][8194] repetition-penalty history, BOS (6561) preset to 1 */
  float* x;                /* device [n_rows][1024] next input embedding per slot */
  float* logits; int ldl;  /* device [n_rows][ldl] logits per slot, ldl >= 8194 */
  float cfg_weight, rep_penalty, temperature, min_p, top_p;
  const float* q_noise;    /* optional device [steps][n_utts][8194] Exp(1) draws (torch.multinomial parity) */
  unsigned long long seed; /* counter-RNG seed when q_noise == NULL */
  int sampler;             /* 0 = T3.inference order (CFG, repetition penalty, temperature, min-p, top-p; t3.py:339-356),
                              1 = T3.inference_turbo order (temperature, top-k, top-p, repetition penalty; t3.py:396-404);
                              must match the loaded backbone (Llama / GPT-2) */
  int top_k;               /* sampler 1 only; <= 0 disables */
  /* device-side retirement (replaces the per-step host sync of t3.py:366): caller-owned device buffers that the decode
   * step keeps up to date.  Before the first cbx_t3_decode: act_utt = 0..n_utts-1, n_act = n_utts; the rest is scratch. */
  int* act_utt;            /* device [n_utts] active utterance ids, packed at the front (stable order) */
  int* n_act;              /* device scalar: live entries of act_utt; the host may poll it asynchronously */
  int* src_slot;           /* device [n_utts] scratch */
  int* slot_row;           /* device [n_rows] physical KV row of each live slot */
  int* m_live;             /* device scalar: live decode rows */
  /* teacher forcing (parity tests): when set, step i of utterance u feeds force_tokens[u][i] (stride max_tokens) and
   * the id the sampler picked itself goes to sampled_out[u][i] (optional) */
  const int* force_tokens;
  int* sampled_out;
  /* operand format of the decode-step projections: 0 = activations as bf16 hi/lo planes (fp32-faithful: bit-exact greedy
   * ids with an fp32 KV cache), 1 = one fp16 plane against the fp16 copy of the weights (throughput mode: half the tensor
   * work; error of the order of the bf16 KV rounding it is meant to be combined with) */
  int act_fp16;
} cbx_t3_state;

/* replaces T3.prepare_conditioning + T3CondEnc.forward + Perceiver.forward
 * (t3.py:92-100, modules/cond_enc.py:64-97, modules/perceiver.py:200-212).  cond_out [n_voices][34][1024].
 * Turbo checkpoint (GPT-2 backbone, tts_turbo.py:151-159: no perceiver, no emotion row, no position table):
 * cond_out [n_voices][1 + n_prompt][1024] = [spkr_enc(speaker_emb) | speech_emb(prompt_tokens)]; emotion_adv unused */
int cbx_t3_cond_encode(cbx_handle* h, const float* speaker_emb, const int* prompt_tokens, int n_prompt,
                       const float* emotion_adv, int n_voices, float* cond_out, void* ws, size_t ws_bytes,
                       cbx_stream stream);
/* replaces T3.prepare_input_embeds + the prefill forward of T3.inference (t3.py:102-130, 303-335 ->
 * transformers LlamaModel.forward) for a packed batch of n_tok tokens; fills the KV pages and st->logits.
 * Row layout [cond(len_cond) | text | BOS | BOS]; Turbo (T3.inference_turbo prefill, t3.py:407-424 -> transformers
 * GPT2Model.forward with inputs_embeds, wpe added inside): [cond | text | BOS], one row per utterance */
int cbx_t3_prefill(cbx_handle* h, const cbx_t3_state* st, int n_tok, const int* tok_row, const int* tok_pos,
                   const int* row_start, const int* row_len, int max_row_len, const float* cond,
                   const int* row_voice, int len_cond, const int* text_flat, const int* text_start,
                   const int* n_text, const int* row_uncond, void* ws, size_t ws_bytes, cbx_stream stream);
/* replaces n_steps iterations of the sampling loop of T3.inference (t3.py:338-386): sample (CFG, repetition
 * penalty, temperature, min-p, top-p, multinomial) then one cached forward, for the n_act active utterances.
 * With st->sampler == 1 it is the loop of T3.inference_turbo (t3.py:426-461): the first token comes from the prefill
 * logits without an EOS check, history for the repetition penalty is BOS for that token and the generated ids after */
/* `capacity` = slots launched per step: any upper bound of the live utterances (st->n_act after the step's own
 * compaction); finished utterances are retired on the device every step, the host only shrinks `capacity` when it
 * learns (asynchronously) that fewer are left.  Each step = one CUDA graph replay (option "decode_graph", default on:
 * the call must then be issued on a capturable, non-default stream). */
int cbx_t3_decode(cbx_handle* h, const cbx_t3_state* st, int capacity, int n_steps, void* ws, size_t ws_bytes,
                  cbx_stream stream);
size_t cbx_t3_workspace_bytes(cbx_handle* h, int n_tok_prefill, int n_rows);

/* ---- S3Gen flow: token -> mel (reference src/chatterbox/models/s3gen/) ---------------------------- */
/* replaces CausalMaskedDiffWithXvec.inference up to the decoder call (flow.py:149-185) incl.
 * UpsampleConformerEncoder.forward (transformer/upsample_encoder.py:237-304).
 * tokens: device int32 [L1.rows] (prompt+generated ids per sequence); xvec [n_seq][192];
 * outputs mu [L2.rows][80], spk [n_seq][80] */
int cbx_flow_encode(cbx_handle* h, const int* tokens, const cbx_layout* L1, const cbx_layout* L2,
                    const float* xvec, float* mu, float* spk, void* ws, size_t ws_bytes, cbx_stream stream);
/* replaces CausalConditionalCFM.forward / solve_euler / basic_euler (flow_matching.py:195-246, 78-145) with
 * ConditionalDecoder.forward (decoder.py:243-333) as the estimator.  x: in = noise z, out = mel, [L2.rows][80]
 * channel-last; cond [L2.rows][80] (prompt mel then zeros); L3 = CFG layout with 2*n_seq sequences
 * (or L2 again when meanflow / no CFG). */
int cbx_cfm_solve(cbx_handle* h, const float* mu, const float* spk, const float* cond, float* x,
                  const cbx_layout* L2, const cbx_layout* L3, int n_steps, float cfg_rate, int meanflow,
                  void* ws, size_t ws_bytes, cbx_stream stream);
size_t cbx_flow_workspace_bytes(cbx_handle* h, const cbx_layout* L1, const cbx_layout* L2, const cbx_layout* L3);

/* ---- HiFT vocoder: mel -> waveform (reference hifigan.py, f0_predictor.py) ------------------------ */
typedef struct {
  cbx_layout LT;          /* mel frames T per sequence */
  cbx_layout L8, L40;     /* 8T, 40T rows */
  cbx_layout L120;        /* 120T + 1 rows (also the STFT frame layout) */
  const long long* sample_start; /* device [n_seq] first sample of each sequence in s / wav (480T samples each) */
  long long total_samples;
} cbx_hift_geom;
/* replaces ConvRNNF0Predictor.forward + f0_upsamp + SourceModuleHnNSF.forward (f0_predictor.py:52-55,
 * hifigan.py:200-231,267-283, 462-466).  phase_vec [n_seq][9] and noise (per sequence [9][480T] at 9*sample_start)
 * may be NULL (phase 0 / counter RNG).  s_out [total_samples]; f0_out optional [LT.rows]; f0_in optional [LT.rows]
 * overrides the predictor (the source integrates f0 over every sample, so parity tests inject the reference f0) */
int cbx_hift_source(cbx_handle* h, const float* mel, const cbx_hift_geom* g, const float* phase_vec,
                    const float* noise, unsigned long long seed, float* s_out, const float* f0_in, float* f0_out,
                    void* ws, size_t ws_bytes, cbx_stream stream);
/* replaces HiFTGenerator.decode (hifigan.py:412-444) + the trim-fade of S3Token2Wav.inference (s3gen.py:359-360) */
int cbx_hift_decode(cbx_handle* h, const float* mel, const float* s, const cbx_hift_geom* g, float* wav,
                    int trim_fade, void* ws, size_t ws_bytes, cbx_stream stream);
size_t cbx_hift_workspace_bytes(cbx_handle* h, const cbx_hift_geom* g);

/* ---- diagnostic entry points (unit tests of single kernels) --------------------------------------- */
/* C[M][N] = act(A_gather x W^T + bias) with W given on the host [N][cin][taps] (torch conv layout; taps=1, cin=K
 * for Linear).  mode 0 = TAPS (k = tap*ceil64(cin)+c), 1 = WINDOW.  A, C device pointers.
 * act = activation code (0 none, 1 silu, 2 gelu, 3 mish, 4 elu, 5 lrelu, 6 snake, 7 tanh); +100 selects the 3-plane
 * (24-bit) activation split. */
int cbx_test_gemm(cbx_handle* h, const float* A, int lda, int M_in, int M, const float* w_host, const float* bias_host,
                  int N, int cin, int taps, int mode, int dil, int pad, int stride, const cbx_layout* out_layout,
                  const cbx_layout* in_layout, int act, float act_p, const float* res, int ldr, int swiglu,
                  float* C, int ldc, cbx_stream stream);
/* O = softmax((Q K^T + bias) * scale) V on packed rows, head_dim 64 */
int cbx_test_attention(cbx_handle* h, const float* Q, const float* K, const float* V, int ld, float* O, int ldo,
                       int n_heads, const cbx_layout* L, float scale, int causal, const float* bias, long long bias_head_stride,
                       int bias_ld, int bias_rel, int bias_center, cbx_stream stream);

/* tcgen05 attention (CFM path): qkv fp32 [rows][3*n_heads*64] (q | k | v), split to bf16 hi/lo planes in ws */
int cbx_test_attention_tc(cbx_handle* h, const float* qkv, float* O, int n_heads, const cbx_layout* L, float scale,
                          void* ws, size_t ws_bytes, cbx_stream stream);

/* paged decode attention of one layer over a caller-built cache pages[n_pages][2][16][32][64] (kv_dtype 0 = bf16, 1 = fp32):
 * qkv [n_slots][3072] fp32 (q | k | v of the step's token), out [n_slots][1024].  impl 0 = bulk-copy staged kernel,
 * 1 = __ldg kernel.  fuse_rope = 1: qkv is un-rotated; the kernel applies RoPE (cos_t / sin_t [pos][32]), appends
 * k / v at `positions` and attends to them. */
int cbx_test_paged_decode(cbx_handle* h, const float* qkv, void* pages, int kv_dtype, int n_pages, const int* page_table,
                          int max_pages, const int* slot_row, const int* positions, int n_slots, int nsplit, int impl,
                          int fuse_rope, const float* cos_t, const float* sin_t, float* out, void* ws, size_t ws_bytes,
                          cbx_stream stream);
/* the decode-path projection: A [M][K] as bf16 hi/lo planes x W [N][K]^T with split-K partial sums reduced in a fixed
 * order (N <= 1024).  tile_bn 0 = heuristic. */
int cbx_test_gemm_splitk(cbx_handle* h, const float* A, const float* w_host, int M, int N, int K, int splitk, int tile_bn,
                         float* C, void* ws, size_t ws_bytes, cbx_stream stream);

/* C[M][N] = act(A x W^T + bias) (+ res) through the fp16-plane operand format of the CFM transformer blocks (A rounded to one
 * fp16 plane, fp16 copy of W; K = 256 / 512 take the weight-resident persistent kernel).  act: 0 none, 2 gelu.  out_half = 1:
 * the result passes through an fp16 plane (as the qkv / ff1 projections write it) before it is widened into C. */
int cbx_test_gemm_f16(cbx_handle* h, const float* A, const float* w_host, const float* bias_host, const float* res, int M,
                      int N, int K, int act, int out_half, float* C, void* ws, size_t ws_bytes, cbx_stream stream);

/* micro-benchmark of one decode-step projection (A fp16 plane [M][K], fp16 weight copy, optional SwiGLU epilogue, split-K partial
 * sums): `n_weights` distinct weight copies are used round-robin so that the weights stream from HBM as they do in the real
 * step; *us_out = average device time per launch in microseconds (tools/decode_gemm_bench.py). */
int cbx_bench_gemm_f16(cbx_handle* h, int M, int N, int K, int splitk, int tile_bn, int tile_dual, int swiglu, int n_weights,
                       int reps, float* us_out, void* ws, size_t ws_bytes, cbx_stream stream);

/* hardware probe: D[128][64] = A[shift .. shift+127][0..63] . W[64][64]^T, A (bf16 [160][64]) staged once in shared memory and
 * read through a row-shifted SWIZZLE_128B UMMA descriptor (mode 1: with the descriptor's base-offset field set) */
int cbx_test_umma_rowshift(cbx_handle* h, const void* A_bf16, const void* W_bf16, int shift, int mode, float* C, cbx_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* CBX_H_ */
