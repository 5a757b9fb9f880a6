// C ABI of libcbx (include/cbx.h): handle lifetime, weight staging, exception -> status-code fence.
#include "engine.h"
#include <vector>
#include <cuda_fp16.h>
#include <cstdlib>

using namespace cbx;

namespace cbx {
const HostTensor& host_tensor(cbx_handle* h, const std::string& name) {
  auto it = h->host.find(name);
  if (it == h->host.end()) throw std::runtime_error("cbx: missing tensor '" + name + "'");
  return it->second;
}
bool has_tensor(cbx_handle* h, const std::string& name) { return h->host.count(name) != 0; }
DevVec upload_vec(cbx_handle* h, const float* p, size_t n) {
  DevVec v; v.n = n;
  CBX_CHECK(cudaMalloc(&v.p, n * sizeof(float)));
  CBX_CHECK(cudaMemcpy(v.p, p, n * sizeof(float), cudaMemcpyHostToDevice));
  h->owned.push_back(v.p);
  return v;
}
DevVec upload_tensor(cbx_handle* h, const std::string& name) {
  const HostTensor& t = host_tensor(h, name);
  return upload_vec(h, t.data.data(), t.data.size());
}
}  // namespace cbx

static Ctx make_ctx(cbx_handle* h, void* ws, size_t ws_bytes, cbx_stream stream, bool dry = false) {
  Ctx c;
  c.stream = reinterpret_cast<cudaStream_t>(stream);
  c.ws.base = static_cast<char*>(ws); c.ws.cap = ws_bytes; c.ws.dry = dry; c.dry = dry;
  c.gemm_impl = h->gemm_impl; c.attn_impl = h->attn_impl; c.attn_f16 = h->attn_f16; c.cfm_act_f16 = h->cfm_act_f16;
  c.timer = (h->timer.cls != K_NONE && !dry) ? &h->timer : nullptr;
  return c;
}

#define CBX_GUARD_BEGIN try {
#define CBX_GUARD_END(h)                                                                   \
  } catch (const std::exception& e) {                                                      \
    if (h) (h)->err = e.what();                                                            \
    const std::string _m = e.what();                                                       \
    if (_m.find("workspace") != std::string::npos) return CBX_ERR_WORKSPACE;               \
    if (_m.find("CUDA error") != std::string::npos) return CBX_ERR_CUDA;                   \
    return CBX_ERR_INVALID;                                                                \
  }                                                                                        \
  return CBX_OK;

extern "C" {

int cbx_version(void) { return 1; }

int cbx_create(int device, cbx_handle** out) {
  if (!out) return CBX_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device >= n) return CBX_ERR_CUDA;   // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return CBX_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CBX_ERR_CUDA;
  if (prop.major != 10) return CBX_ERR_CUDA;          // libcbx is built for sm_100a only; no handle is handed out
  cbx_handle* h = new cbx_handle();
  h->device = device;
  const char* e = getenv("CBX_GEMM");
  if (e && std::string(e) == "simt") h->gemm_impl = 1;
  e = getenv("CBX_ATTN");
  if (e && std::string(e) == "simt") h->attn_impl = 1;
  e = getenv("CBX_DECODE_GRAPH");
  if (e) h->decode_graph = atoi(e);
  e = getenv("CBX_DECODE_PDL");
  if (e) h->decode_pdl = atoi(e);
  try {        // per-device kernel attributes, before any launch or stream capture
    gemm_init(); attention_tc_init(); paged_attention_init(); t3_sample_init(); hift_conv_init();
  } catch (const std::exception& ex) {
    delete h;
    return CBX_ERR_CUDA;
  }
  *out = h;
  return CBX_OK;
}

void cbx_destroy(cbx_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (void* p : h->owned) cudaFree(p);
  for (auto& kv : h->decode_graphs) cudaGraphExecDestroy(kv.second.exec);
  for (cudaEvent_t e : h->timer.ev) cudaEventDestroy(e);
  auto fw = [](Weight& w) { free_weight(w); };      // no-op for weights that were never packed
  for (auto& l : h->t3.layers) { fw(l.qkv); fw(l.o); fw(l.gu); fw(l.down); }
  fw(h->t3.head); fw(h->t3.spkr); fw(h->t3.pq); fw(h->t3.pk); fw(h->t3.pv); fw(h->t3.pproj);
  FlowModel& f = h->flow;
  fw(f.spk_affine); fw(f.enc_proj); fw(f.embed.lin); fw(f.up_embed.lin); fw(f.pre_conv1); fw(f.pre_conv2); fw(f.up_conv);
  auto fenc = [&](EncLayer& e) { fw(e.qkv); fw(e.out); fw(e.pos); fw(e.w1); fw(e.w2); };
  for (auto& e : f.enc) fenc(e);
  for (auto& e : f.up_enc) fenc(e);
  fw(f.time1); fw(f.time2); fw(f.time_mixer);
  auto fstage = [&](CfmStage& st) {
    fw(st.res.conv1); fw(st.res.conv2); fw(st.res.res); fw(st.res.mlp);
    for (auto& t : st.t) { fw(t.qkv); fw(t.out); fw(t.ff1); fw(t.ff2); }
  };
  fstage(f.down); for (auto& st : f.mid) fstage(st); fstage(f.up);
  fw(f.down_conv); fw(f.up_conv2); fw(f.final_conv); fw(f.final_proj);
  HiftModel& v = h->hift;
  for (auto& w : v.f0conv) fw(w);
  fw(v.conv_pre); fw(v.conv_post);
  for (auto& w : v.ups) fw(w);
  for (auto& w : v.src_down) fw(w);
  auto frb = [&](HiftResBlock& rb) { for (auto& w : rb.c1) fw(w); for (auto& w : rb.c2) fw(w); };
  for (auto& rb : v.src_rb) frb(rb);
  for (auto& rb : v.rb) frb(rb);
  delete h;
}

const char* cbx_last_error(cbx_handle* h) { return h ? h->err.c_str() : "null handle"; }

int cbx_set_option(cbx_handle* h, const char* key, const char* value) {
  if (!h || !key || !value) return CBX_ERR_INVALID;
  const std::string k = key, v = value;
  if (k == "gemm") h->gemm_impl = (v == "simt") ? 1 : 0;
  else if (k == "attn") h->attn_impl = (v == "simt") ? 1 : 0;
  else if (k == "decode_graph") h->decode_graph = (v == "1" || v == "on") ? 1 : 0;
  else if (k == "decode_pdl") h->decode_pdl = (v == "1" || v == "on") ? 1 : 0;
  else if (k == "attn_prec") h->attn_f16 = (v == "fp16") ? 1 : 0;     // CFM attention operand format (default bf16x3)
  else if (k == "cfm_act") h->cfm_act_f16 = (v == "fp16") ? 1 : 0;   // CFM transformer-block GEMM inputs (default bf16x2)
  else if (k == "time_kernel") {
    h->timer.reset();
    h->timer.cls = v == "gemm_tc" ? K_GEMM_TC : v == "gemv" ? K_GEMV : v == "flash" ? K_FLASH : v == "paged" ? K_PAGED : v == "wres" ? K_WRES :
                   v == "stream" ? K_STREAM : v == "attn_tc" ? K_ATTN_TC : v == "hift_conv" ? K_HIFT_CONV : v == "all" ? K_ALL : K_NONE;
  }
  else { h->err = "unknown option " + k; return CBX_ERR_INVALID; }
  return CBX_OK;
}

long long cbx_launch_count(cbx_handle* h) { return h ? h->launches : 0; }

int cbx_timer_read(cbx_handle* h, double* ms, long long* launches, double* work) {
  if (!h || !ms || !launches) return CBX_ERR_INVALID;
  h->timer.drain();
  *ms = h->timer.ms(); *launches = h->timer.n();
  if (work) *work = h->timer.work();
  return CBX_OK;
}

int cbx_timer_read_bytes(cbx_handle* h, double* bytes) {
  if (!h || !bytes) return CBX_ERR_INVALID;
  h->timer.drain();
  *bytes = h->timer.bytes();
  return CBX_OK;
}

int cbx_timer_read_class(cbx_handle* h, const char* cls, double* ms, long long* launches, double* work, double* bytes) {
  if (!h || !cls || !ms || !launches) return CBX_ERR_INVALID;
  const std::string v = cls;
  const int c = v == "gemm_tc" ? K_GEMM_TC : v == "gemv" ? K_GEMV : v == "flash" ? K_FLASH : v == "paged" ? K_PAGED : v == "wres" ? K_WRES :
                v == "stream" ? K_STREAM : v == "attn_tc" ? K_ATTN_TC : v == "hift_conv" ? K_HIFT_CONV : K_NONE;
  if (c == K_NONE) return CBX_ERR_INVALID;
  h->timer.drain();
  *ms = h->timer.ms_c[c]; *launches = h->timer.n_c[c];
  if (work) *work = h->timer.work_c[c];
  if (bytes) *bytes = h->timer.bytes_c[c];
  return CBX_OK;
}

int cbx_load_tensor(cbx_handle* h, const char* name, const float* host_data, int ndim, const int64_t* shape) {
  if (!h) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  CBX_REQUIRE(name && host_data && ndim >= 0 && ndim <= 4, "bad tensor");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.data.assign(host_data, host_data + n);
  h->host[name] = std::move(t);
  CBX_GUARD_END(h)
}

int cbx_finalize_weights(cbx_handle* h, const char* model) {
  if (!h || !model) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  CBX_CHECK(cudaSetDevice(h->device));
  const std::string m = model;
  if (m == "t3") t3_finalize(h);
  else if (m == "flow") flow_finalize(h);
  else if (m == "hift") hift_finalize(h);
  else throw std::runtime_error("unknown model " + m);
  // staged host copies of this model are no longer needed
  for (auto it = h->host.begin(); it != h->host.end();) {
    if (it->first.compare(0, m.size() + 1, m + ".") == 0) it = h->host.erase(it); else ++it;
  }
  CBX_CHECK(cudaDeviceSynchronize());
  CBX_GUARD_END(h)
}

// ---- T3 -------------------------------------------------------------------------------------------
int cbx_t3_cond_encode(cbx_handle* h, const float* speaker_emb, const int* prompt_tokens, int n_prompt,
                       const float* emotion_adv, int n_voices, float* cond_out, void* ws, size_t ws_bytes,
                       cbx_stream stream) {
  if (!h) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  t3_cond_encode(h, c, speaker_emb, prompt_tokens, n_prompt, emotion_adv, n_voices, cond_out);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}
int cbx_t3_prefill(cbx_handle* h, const cbx_t3_state* st, int n_tok, const int* tok_row, const int* tok_pos,
                   const int* row_start, const int* row_len, int max_row_len, const float* cond, const int* row_voice,
                   int len_cond, const int* text_flat, const int* text_start, const int* n_text, const int* row_uncond,
                   void* ws, size_t ws_bytes, cbx_stream stream) {
  if (!h || !st) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  t3_prefill(h, c, *st, n_tok, tok_row, tok_pos, row_start, row_len, max_row_len, cond, row_voice, len_cond, text_flat,
             text_start, n_text, row_uncond);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}
int cbx_t3_decode(cbx_handle* h, const cbx_t3_state* st, int capacity, int n_steps, void* ws, size_t ws_bytes,
                  cbx_stream stream) {
  if (!h || !st) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  t3_decode(h, c, *st, capacity, n_steps);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}
size_t cbx_t3_workspace_bytes(cbx_handle* h, int n_tok_prefill, int n_rows) {
  if (!h) return 0;
  try {
    cbx_t3_state st; memset(&st, 0, sizeof(st));
    const bool gpt = h->t3.gpt;     // Turbo: one row per utterance, inference_turbo sampler
    st.n_rows = n_rows; st.cfg = gpt ? 0 : 1; st.n_utts = gpt ? n_rows : (n_rows + 1) / 2; st.ldl = 8256;
    st.sampler = gpt ? 1 : 0;
    Ctx c = make_ctx(h, nullptr, 0, nullptr, true);
    t3_prefill(h, c, st, n_tok_prefill, nullptr, nullptr, nullptr, nullptr, 1, nullptr, nullptr, 34, nullptr, nullptr,
               nullptr, nullptr);
    size_t a = c.ws.peak;
    Ctx d = make_ctx(h, nullptr, 0, nullptr, true);
    t3_decode(h, d, st, st.n_utts, 1);
    size_t b = d.ws.peak + 4096;
    Ctx e = make_ctx(h, nullptr, 0, nullptr, true);
    t3_cond_encode(h, e, nullptr, nullptr, 512, nullptr, 1, nullptr);
    size_t m = a > b ? a : b;
    if (e.ws.peak > m) m = e.ws.peak;
    return m + (1 << 20);
  } catch (const std::exception& e) { h->err = e.what(); return 0; }
}

// ---- flow -----------------------------------------------------------------------------------------
int cbx_flow_encode(cbx_handle* h, const int* tokens, const cbx_layout* L1, const cbx_layout* L2, const float* xvec,
                    float* mu, float* spk, void* ws, size_t ws_bytes, cbx_stream stream) {
  if (!h || !L1 || !L2) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  flow_encode(h, c, tokens, *L1, *L2, xvec, mu, spk);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}
int cbx_cfm_solve(cbx_handle* h, const float* mu, const float* spk, const float* cond, float* x, const cbx_layout* L2,
                  const cbx_layout* L3, int n_steps, float cfg_rate, int meanflow, void* ws, size_t ws_bytes,
                  cbx_stream stream) {
  if (!h || !L2 || !L3) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  cfm_solve(h, c, mu, spk, cond, x, *L2, *L3, n_steps, cfg_rate, meanflow);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}
size_t cbx_flow_workspace_bytes(cbx_handle* h, const cbx_layout* L1, const cbx_layout* L2, const cbx_layout* L3) {
  if (!h || !L1 || !L2 || !L3) return 0;
  try {
    Ctx c = make_ctx(h, nullptr, 0, nullptr, true);
    flow_encode(h, c, nullptr, *L1, *L2, nullptr, nullptr, nullptr);
    Ctx d = make_ctx(h, nullptr, 0, nullptr, true);
    cfm_solve(h, d, nullptr, nullptr, nullptr, nullptr, *L2, *L3, h->flow.meanflow ? 2 : 10, 0.7f, h->flow.meanflow ? 1 : 0);
    return (c.ws.peak > d.ws.peak ? c.ws.peak : d.ws.peak) + (1 << 20);
  } catch (const std::exception& e) { h->err = e.what(); return 0; }
}

// ---- hift -----------------------------------------------------------------------------------------
int cbx_hift_source(cbx_handle* h, const float* mel, const cbx_hift_geom* g, const float* phase_vec, const float* noise,
                    unsigned long long seed, float* s_out, const float* f0_in, float* f0_out, void* ws, size_t ws_bytes,
                    cbx_stream stream) {
  if (!h || !g) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  hift_source_run(h, c, mel, *g, phase_vec, noise, seed, s_out, f0_in, f0_out);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}
int cbx_hift_decode(cbx_handle* h, const float* mel, const float* s, const cbx_hift_geom* g, float* wav, int trim_fade,
                    void* ws, size_t ws_bytes, cbx_stream stream) {
  if (!h || !g) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  hift_decode_run(h, c, mel, s, *g, wav, trim_fade);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}
size_t cbx_hift_workspace_bytes(cbx_handle* h, const cbx_hift_geom* g) {
  if (!h || !g) return 0;
  try {
    Ctx c = make_ctx(h, nullptr, 0, nullptr, true);
    hift_source_run(h, c, nullptr, *g, nullptr, nullptr, 0, nullptr, nullptr, nullptr);
    Ctx d = make_ctx(h, nullptr, 0, nullptr, true);
    hift_decode_run(h, d, nullptr, nullptr, *g, nullptr, 1);
    return (c.ws.peak > d.ws.peak ? c.ws.peak : d.ws.peak) + (1 << 20);
  } catch (const std::exception& e) { h->err = e.what(); return 0; }
}

// ---- diagnostics ------------------------------------------------------------------------------------
int cbx_test_gemm(cbx_handle* h, const float* A, int lda, int M_in, int M, const float* w_host, const float* bias_host,
                  int N, int cin, int taps, int mode, int dil, int pad, int stride, const cbx_layout* out_layout,
                  const cbx_layout* in_layout, int act, float act_p, const float* res, int ldr, int swiglu, float* C,
                  int ldc, cbx_stream stream) {
  if (!h) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Weight W;
  if (mode == 1) pack_conv_window(W, w_host, bias_host, N, cin, taps);
  else pack_conv_taps(W, w_host, bias_host, N, cin, taps);
  Ctx c = make_ctx(h, nullptr, 0, stream);
  GemmDev g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.M = M; g.M_in = M_in;
  g.a_mode = mode; g.ntaps = taps; g.ctap = (cin + 63) / 64 * 64; g.c_in = cin; g.dil = dil; g.pad = pad; g.stride = stride;
  g.k_total = mode == 1 ? taps * cin : taps * g.ctap;
  if (out_layout && in_layout) { g.has_seq = 1; g.seq = seqmap(*out_layout, *in_layout); }
  g.Wp = W.w; g.Kpad = W.Kpad; g.Npad = W.Npad;
  g.C = C; g.ldc = ldc; g.n_out = N; g.bias = W.bias; g.alpha = 1.f; g.act = act; g.act_p = act_p; g.out_scale = 1.f;
  g.res = res; g.ldr = ldr; g.swiglu = swiglu;
  long long* dbgbuf = nullptr;
  if (act >= 1000) { g.act = act - 1000; CBX_CHECK(cudaMalloc(&dbgbuf, 64 * 8)); CBX_CHECK(cudaMemset(dbgbuf, 0, 64 * 8)); g.dbg = dbgbuf; }
  gemm(c, g, W);
  if (dbgbuf) {
    long long hbuf[64]; CBX_CHECK(cudaMemcpy(hbuf, dbgbuf, sizeof(hbuf), cudaMemcpyDeviceToHost)); cudaFree(dbgbuf);
    const long long t0 = hbuf[5];
    printf("[gemm dbg] M=%d N=%d K=%d: setup_done=%lld prod_done=%lld accum_ready=%lld epi_done=%lld exit=%lld\n", M, N, cin * taps,
           hbuf[0] - t0, hbuf[1] - t0, hbuf[2] - t0, hbuf[3] - t0, hbuf[4] - t0);
    printf("   epilogue chunk0: ldtm=%lld to_smem=%lld rows_loop=%lld\n", hbuf[41] - hbuf[40], hbuf[42] - hbuf[41], hbuf[43] - hbuf[42]);
    for (int kb = 0; kb < 8; ++kb) printf("   kb%d: conv_start=%lld mma_issue=%lld\n", kb, hbuf[8 + kb] - t0, hbuf[32 + kb] - t0);
    fflush(stdout);
  }
  CBX_CHECK(cudaStreamSynchronize(c.stream));
  free_weight(W);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}
int cbx_test_attention(cbx_handle* h, const float* Q, const float* K, const float* V, int ld, float* O, int ldo,
                       int n_heads, const cbx_layout* L, float scale, int causal, const float* bias,
                       long long bias_head_stride, int bias_ld, int bias_rel, int bias_center, cbx_stream stream) {
  if (!h || !L) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, nullptr, 0, stream);
  AttnArgs a;
  a.Q = Q; a.K = K; a.V = V; a.ldq = a.ldk = a.ldv = ld; a.O = O; a.ldo = ldo; a.n_seq = L->n_seq; a.n_heads = n_heads;
  a.q_start = L->start; a.q_len = L->len; a.kv_start = L->start; a.kv_len = L->len; a.max_q_len = L->max_len;
  a.scale = scale; a.causal = causal; a.bias = bias; a.bias_head_stride = bias_head_stride; a.bias_ld = bias_ld;
  a.bias_row0 = 0; a.bias_rel = bias_rel; a.bias_center = bias_center;
  attention(c, a);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}

int cbx_test_attention_tc(cbx_handle* h, const float* qkv, float* O, int n_heads, const cbx_layout* L, float scale,
                          void* ws, size_t ws_bytes, cbx_stream stream) {
  if (!h || !L) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  const int ld = 3 * n_heads * 64;
  __nv_bfloat16* hi = c.ws.get<__nv_bfloat16>((size_t)L->rows * ld);
  __nv_bfloat16* lo = c.ws.get<__nv_bfloat16>((size_t)L->rows * ld);
  pack_hilo(c, qkv, ld, L->rows, ld, hi, lo, L->rows, ld);
  CUtensorMap tmh, tml;
  make_plane_tmap(&tmh, hi, L->rows, ld);
  make_plane_tmap(&tml, lo, L->rows, ld);
  AttnTcArgs a;
  a.tm_hi = &tmh; a.tm_lo = &tml; a.q_col = 0; a.k_col = n_heads * 64; a.v_col = 2 * n_heads * 64; a.O = O;
  a.ldo = n_heads * 64; a.n_seq = L->n_seq; a.n_heads = n_heads; a.q_start = L->start; a.q_len = L->len;
  a.kv_start = L->start; a.kv_len = L->len; a.max_q_len = L->max_len; a.scale = scale;
  attention_tc(c, a);
  CBX_CHECK(cudaStreamSynchronize(c.stream));
  h->launches += c.launches;
  CBX_GUARD_END(h)
}

/* paged decode attention over a caller-built cache (layer 0 of a 1-layer pool) */
int cbx_test_paged_decode(cbx_handle* h, const float* qkv, void* pages, int kv_dtype, int n_pages, const int* page_table,
                          int max_pages, const int* slot_row, const int* positions, int n_slots, int nsplit, int impl,
                          int fuse_rope, const float* cos_t, const float* sin_t, float* out, void* ws, size_t ws_bytes,
                          cbx_stream stream) {
  if (!h) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  PagedKV kv;
  kv.pages = pages; kv.n_pages = n_pages; kv.kv_fp32 = kv_dtype; kv.n_layers = 1; kv.n_heads = 16; kv.page_tokens = 32;
  kv.page_table = page_table; kv.max_pages_per_row = max_pages;
  float* scratch = c.ws.get<float>((size_t)n_slots * 16 * nsplit * 66);
  PagedOpts po;
  po.fuse_rope = fuse_rope; po.cos_t = cos_t; po.sin_t = sin_t; po.impl = impl;
  paged_decode_attention(c, qkv, 3072, kv, 0, slot_row, n_slots, positions, out, 1024, scratch, nsplit, nullptr, nullptr, &po);
  CBX_CHECK(cudaStreamSynchronize(c.stream));
  h->launches += c.launches;
  CBX_GUARD_END(h)
}

/* C[M][N] (N <= 1024) = A[M][K] x W[N][K]^T through the decode path: A as bf16 hi/lo planes, split-K partial sums,
 * fixed-order reduction in resid_norm */
int cbx_test_gemm_splitk(cbx_handle* h, const float* A, const float* w_host, int M, int N, int K, int splitk, int tile_bn,
                         float* C, void* ws, size_t ws_bytes, cbx_stream stream) {
  if (!h) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  CBX_REQUIRE(N <= 1024 && N % 4 == 0 && K % 64 == 0, "test shape");
  Weight W;
  pack_linear(W, w_host, nullptr, N, K);
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  __nv_bfloat16* hi = c.ws.get<__nv_bfloat16>((size_t)M * K);
  __nv_bfloat16* lo = c.ws.get<__nv_bfloat16>((size_t)M * K);
  float* part = c.ws.get<float>((size_t)splitk * M * N);
  pack_hilo(c, A, K, M, K, hi, lo, M, K);
  GemmDev g = gemm_args_linear(nullptr, K, M, W, part, N);
  g.Ahi = hi; g.Alo = lo; g.ldab = K; g.bias = nullptr;
  g.splitk = splitk; g.split_stride = (long)M * N; g.tile_bn = tile_bn;
  gemm(c, g, W);
  CBX_CHECK(cudaMemsetAsync(C, 0, (size_t)M * N * 4, c.stream));
  ResidNormDev rn;
  memset(&rn, 0, sizeof(rn));
  rn.x = C; rn.ldx = N; rn.part = part; rn.nsplit = splitk; rn.split_stride = (long)M * N; rn.ldp = N; rn.dim = N;
  resid_norm(c, rn, M);
  CBX_CHECK(cudaStreamSynchronize(c.stream));
  free_weight(W);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}

__global__ void f32_to_f16_kernel(const float* x, __half* y, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2half_rn(x[i]);
}
__global__ void f16_to_f32_kernel(const __half* x, float* y, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __half2float(x[i]);
}
/* C[M][N] = act(A x W^T + bias) (+ res) through the fp16-plane operand format of the CFM transformer blocks (A rounded to
 * one fp16 plane, fp16 copy of W): picks the weight-resident persistent kernel for K = 256 / 512.  out_half = 1: the result
 * goes through an fp16 plane (as the qkv / ff1 projections write it) before it is widened into C. */
int cbx_test_gemm_f16(cbx_handle* h, const float* A, const float* w_host, const float* bias_host, const float* res, int M, int N,
                      int K, int act, int out_half, float* C, void* ws, size_t ws_bytes, cbx_stream stream) {
  if (!h) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Weight W;
  pack_linear(W, w_host, bias_host, N, K, true);
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  __half* a16 = c.ws.get<__half>((size_t)M * K);
  __half* c16 = c.ws.get<__half>((size_t)M * N);
  f32_to_f16_kernel<<<(unsigned)(((long)M * K + 255) / 256), 256, 0, c.stream>>>(A, a16, (long)M * K);
  GemmDev g = gemm_args_linear(nullptr, K, M, W, out_half ? nullptr : C, N);
  g.A16 = a16; g.lda16 = K; g.act = act;
  if (out_half) { g.Chi = reinterpret_cast<__nv_bfloat16*>(c16); g.Clo = g.Chi; g.ldcb = N; g.c_half = 1; }
  else if (res) { g.res = res; g.ldr = N; }
  gemm(c, g, W);
  if (out_half) f16_to_f32_kernel<<<(unsigned)(((long)M * N + 255) / 256), 256, 0, c.stream>>>(c16, C, (long)M * N);
  CBX_CHECK(cudaStreamSynchronize(c.stream));
  free_weight(W);
  h->launches += c.launches;
  CBX_GUARD_END(h)
}

/* micro-benchmark of one decode-step projection: C = A16[M][K] x W16[N][K]^T (+ SwiGLU) with `n_weights` distinct copies of
 * the weight used round-robin (together larger than L2, as in the real step where 1 GB of weights streams per step);
 * returns the average device time per launch in microseconds (CUDA events around reps * n_weights launches). */
int cbx_bench_gemm_f16(cbx_handle* h, int M, int N, int K, int splitk, int tile_bn, int tile_dual, int swiglu, int n_weights,
                       int reps, float* us_out, void* ws, size_t ws_bytes, cbx_stream stream) {
  if (!h || !us_out) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  CBX_REQUIRE(n_weights >= 1 && n_weights <= 64 && reps >= 1, "bench shape");
  std::vector<float> hw((size_t)N * K);
  uint32_t x = 12345u;
  for (auto& v : hw) { x = x * 1664525u + 1013904223u; v = ((float)(x >> 8) / 16777216.0f - 0.5f) * 0.06f; }
  std::vector<Weight> Ws(n_weights);
  pack_linear(Ws[0], hw.data(), nullptr, N, K, true);
  for (int i = 1; i < n_weights; ++i) {
    Ws[i] = Ws[0];
    Ws[i].w = nullptr; Ws[i].bias = nullptr;
    CBX_CHECK(cudaMalloc(&Ws[i].w16, (size_t)Ws[0].Npad * Ws[0].Kpad * 2));
    CBX_CHECK(cudaMemcpy(Ws[i].w16, Ws[0].w16, (size_t)Ws[0].Npad * Ws[0].Kpad * 2, cudaMemcpyDeviceToDevice));
    Weight tmp = Ws[i]; tmp.w = reinterpret_cast<__nv_bfloat16*>(Ws[i].w16);
    make_tmaps_for(tmp);
    for (int j = 0; j < 3; ++j) Ws[i].tmap16[j] = tmp.tmap[j];
    Ws[i].bias = Ws[0].bias;
  }
  Ctx c = make_ctx(h, ws, ws_bytes, stream);
  c.timer = nullptr;
  __half* a16 = c.ws.get<__half>((size_t)M * K);
  CBX_CHECK(cudaMemsetAsync(a16, 0x2c, (size_t)M * K * 2, c.stream));        // 0x2c2c = 0.0652 in fp16
  const int n_out = swiglu ? N / 2 : N;
  float* C = c.ws.get<float>((size_t)(splitk > 1 ? splitk : 1) * M * n_out);
  __half* c16 = c.ws.get<__half>((size_t)M * n_out);
  cudaEvent_t e0, e1;
  CBX_CHECK(cudaEventCreate(&e0)); CBX_CHECK(cudaEventCreate(&e1));
  auto one = [&](const Weight& W) {
    GemmDev g = gemm_args_linear(nullptr, K, M, W, C, n_out);
    g.A16 = a16; g.lda16 = K; g.bias = nullptr;
    g.tile_bn = tile_bn; g.tile_dual = tile_dual;
    if (swiglu) { g.swiglu = 1; g.C = nullptr; g.Chi = reinterpret_cast<__nv_bfloat16*>(c16); g.Clo = g.Chi; g.ldcb = n_out; g.c_half = 1; }
    if (splitk > 1) { g.splitk = splitk; g.split_stride = (long)M * n_out; }
    gemm(c, g, W);
  };
  for (int i = 0; i < n_weights; ++i) one(Ws[i]);                              // warm-up
  CBX_CHECK(cudaEventRecord(e0, c.stream));
  for (int r = 0; r < reps; ++r)
    for (int i = 0; i < n_weights; ++i) one(Ws[i]);
  CBX_CHECK(cudaEventRecord(e1, c.stream));
  CBX_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CBX_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  *us_out = ms * 1000.0f / (float)(reps * n_weights);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  for (int i = 1; i < n_weights; ++i) cudaFree(Ws[i].w16);
  free_weight(Ws[0]);
  CBX_GUARD_END(h)
}

/* hardware probe: D[128][64] = A[shift .. shift+127][64] . W[64][64]^T with A (bf16 [160][64], device) staged ONCE in shared
 * memory and addressed through a row-shifted SWIZZLE_128B descriptor; mode 1 also sets the descriptor's base-offset field */
int cbx_test_umma_rowshift(cbx_handle* h, const void* A_bf16, const void* W_bf16, int shift, int mode, float* C, cbx_stream stream) {
  if (!h) return CBX_ERR_INVALID;
  CBX_GUARD_BEGIN
  Ctx c = make_ctx(h, nullptr, 0, stream);
  umma_rowshift_probe(c, static_cast<const __nv_bfloat16*>(A_bf16), static_cast<const __nv_bfloat16*>(W_bf16), shift, mode, C);
  CBX_CHECK(cudaStreamSynchronize(c.stream));
  CBX_GUARD_END(h)
}

}  // extern "C"
