// T3: Llama-style 0.5B speech-token LM (reference src/chatterbox/models/t3/t3.py, transformers LlamaModel).
//   prefill  : packed tokens -> 30 x [RMSNorm, QKV GEMM, RoPE + KV append, causal flash attention, O GEMM(+res),
//              RMSNorm, gate/up GEMM with fused SiLU*mul, down GEMM(+res)] -> final norm -> speech head
//   decode   : per step  sampler kernel -> same layer stack with paged decode attention (one token per row)
// Turbo (reference t3.py:392-468, tts_turbo.py:151-166) swaps the backbone for transformers GPT2Model: LayerNorm,
// fused c_attn with bias, learned absolute positions (wpe) added to the input embeddings, gelu_new MLP, no RoPE, no CFG,
// speech head with bias over 6563 ids.  Same kernels, same paged KV cache; the RoPE tables are the identity.
#include "engine.h"
#include <cstdlib>

namespace cbx {

static std::vector<float> concat_rows(const std::vector<const HostTensor*>& ts) {
  std::vector<float> out;
  for (auto* t : ts) out.insert(out.end(), t->data.begin(), t->data.end());
  return out;
}

// GPT-2 `Conv1D` keeps its weight as [in, out]; the GEMM packs K-major [out, in]
static std::vector<float> transposed(const HostTensor& t) {
  const int K = (int)t.shape[0], N = (int)t.shape[1];
  std::vector<float> out((size_t)N * K);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) out[(size_t)n * K + k] = t.data[(size_t)k * N + n];
  return out;
}

static void t3_finalize_gpt(cbx_handle* h) {
  T3Model& m = h->t3;
  m.gpt = true;
  int L = 0;
  while (has_tensor(h, "t3.tfmr.h." + std::to_string(L) + ".attn.c_attn.weight")) ++L;
  m.n_layers = L;
  m.layers.resize(L);
  CBX_REQUIRE(host_tensor(h, "t3.tfmr.h.0.attn.c_attn.weight").shape[0] == 1024, "only the 1024-wide GPT2_medium Turbo backbone is supported");
  for (int i = 0; i < L; ++i) {
    const std::string p = "t3.tfmr.h." + std::to_string(i) + ".";
    T3Layer& ly = m.layers[i];
    pack_linear(ly.qkv, transposed(host_tensor(h, p + "attn.c_attn.weight")).data(), host_tensor(h, p + "attn.c_attn.bias").data.data(), 3072, 1024, true);
    pack_linear(ly.o, transposed(host_tensor(h, p + "attn.c_proj.weight")).data(), host_tensor(h, p + "attn.c_proj.bias").data.data(), 1024, 1024, true);
    pack_linear(ly.gu, transposed(host_tensor(h, p + "mlp.c_fc.weight")).data(), host_tensor(h, p + "mlp.c_fc.bias").data.data(), 4096, 1024, true);
    pack_linear(ly.down, transposed(host_tensor(h, p + "mlp.c_proj.weight")).data(), host_tensor(h, p + "mlp.c_proj.bias").data.data(), 1024, 4096, true);
    ly.ln1 = upload_tensor(h, p + "ln_1.weight"); ly.ln1_b = upload_tensor(h, p + "ln_1.bias");
    ly.ln2 = upload_tensor(h, p + "ln_2.weight"); ly.ln2_b = upload_tensor(h, p + "ln_2.bias");
  }
  m.final_norm = upload_tensor(h, "t3.tfmr.ln_f.weight");
  m.final_norm_b = upload_tensor(h, "t3.tfmr.ln_f.bias");
  m.wpe = upload_tensor(h, "t3.tfmr.wpe.weight");
  m.text_emb = upload_tensor(h, "t3.text_emb.weight");
  m.text_vocab = (int)host_tensor(h, "t3.text_emb.weight").shape[0];
  m.speech_emb = upload_tensor(h, "t3.speech_emb.weight");
  m.vocab = (int)host_tensor(h, "t3.speech_head.weight").shape[0];
  CBX_REQUIRE(m.vocab <= 8194 && (int)host_tensor(h, "t3.speech_emb.weight").shape[0] == m.vocab, "speech vocab");
  m.rope_cos = upload_tensor(h, "t3.rope_cos");      // identity tables supplied by the host shim: cos = 1, sin = 0
  m.rope_sin = upload_tensor(h, "t3.rope_sin");
  m.max_pos = (int)host_tensor(h, "t3.tfmr.wpe.weight").shape[0];
  CBX_REQUIRE((int)host_tensor(h, "t3.rope_cos").shape[0] >= m.max_pos, "rope identity table shorter than wpe");
  pack_linear(m.head, host_tensor(h, "t3.speech_head.weight").data.data(), host_tensor(h, "t3.speech_head.bias").data.data(),
              m.vocab, 1024, true);
  pack_linear(m.spkr, host_tensor(h, "t3.cond_enc.spkr_enc.weight").data.data(),
              host_tensor(h, "t3.cond_enc.spkr_enc.bias").data.data(), 1024, 256);
  m.ready = true;
}

void t3_finalize(cbx_handle* h) {
  T3Model& m = h->t3;
  if (has_tensor(h, "t3.tfmr.h.0.attn.c_attn.weight")) { t3_finalize_gpt(h); return; }
  m.gpt = false; m.vocab = 8194;
  int L = 0;
  while (has_tensor(h, "t3.tfmr.layers." + std::to_string(L) + ".self_attn.q_proj.weight")) ++L;
  CBX_REQUIRE(L > 0, "no T3 layers loaded");
  m.n_layers = L;
  m.layers.resize(L);
  for (int i = 0; i < L; ++i) {
    const std::string p = "t3.tfmr.layers." + std::to_string(i) + ".";
    T3Layer& ly = m.layers[i];
    auto qkv = concat_rows({&host_tensor(h, p + "self_attn.q_proj.weight"), &host_tensor(h, p + "self_attn.k_proj.weight"),
                            &host_tensor(h, p + "self_attn.v_proj.weight")});
    pack_linear(ly.qkv, qkv.data(), nullptr, 3072, 1024, true);    // + fp16 copy: operand of the fp16-activation decode mode
    pack_linear(ly.o, host_tensor(h, p + "self_attn.o_proj.weight").data.data(), nullptr, 1024, 1024, true);
    const auto& g = host_tensor(h, p + "mlp.gate_proj.weight").data;
    const auto& u = host_tensor(h, p + "mlp.up_proj.weight").data;
    std::vector<float> gu((size_t)8192 * 1024);
    for (int j = 0; j < 4096; ++j) {       // interleave so that the GEMM epilogue sees (gate_j, up_j) side by side
      memcpy(&gu[(size_t)(2 * j) * 1024], &g[(size_t)j * 1024], 4096);
      memcpy(&gu[(size_t)(2 * j + 1) * 1024], &u[(size_t)j * 1024], 4096);
    }
    pack_linear(ly.gu, gu.data(), nullptr, 8192, 1024, true);
    pack_linear(ly.down, host_tensor(h, p + "mlp.down_proj.weight").data.data(), nullptr, 1024, 4096, true);
    ly.ln1 = upload_tensor(h, p + "input_layernorm.weight");
    ly.ln2 = upload_tensor(h, p + "post_attention_layernorm.weight");
  }
  m.final_norm = upload_tensor(h, "t3.tfmr.norm.weight");
  m.text_emb = upload_tensor(h, "t3.text_emb.weight");
  m.text_vocab = (int)host_tensor(h, "t3.text_emb.weight").shape[0];
  m.speech_emb = upload_tensor(h, "t3.speech_emb.weight");
  m.text_pos = upload_tensor(h, "t3.text_pos_emb.emb.weight");
  m.speech_pos = upload_tensor(h, "t3.speech_pos_emb.emb.weight");
  m.rope_cos = upload_tensor(h, "t3.rope_cos");
  m.rope_sin = upload_tensor(h, "t3.rope_sin");
  m.max_pos = (int)host_tensor(h, "t3.rope_cos").shape[0];
  pack_linear(m.head, host_tensor(h, "t3.speech_head.weight").data.data(), nullptr, 8194, 1024, true);
  // conditioning encoder
  pack_linear(m.spkr, host_tensor(h, "t3.cond_enc.spkr_enc.weight").data.data(),
              host_tensor(h, "t3.cond_enc.spkr_enc.bias").data.data(), 1024, 256);
  const std::string pa = "t3.cond_enc.perceiver.attn.";
  pack_linear(m.pq, host_tensor(h, pa + "to_q.weight").data.data(), host_tensor(h, pa + "to_q.bias").data.data(), 1024, 1024);
  pack_linear(m.pk, host_tensor(h, pa + "to_k.weight").data.data(), host_tensor(h, pa + "to_k.bias").data.data(), 1024, 1024);
  pack_linear(m.pv, host_tensor(h, pa + "to_v.weight").data.data(), host_tensor(h, pa + "to_v.bias").data.data(), 1024, 1024);
  pack_linear(m.pproj, host_tensor(h, pa + "proj_out.weight").data.data(), host_tensor(h, pa + "proj_out.bias").data.data(), 1024, 1024);
  m.emotion_w = upload_tensor(h, "t3.cond_enc.emotion_adv_fc.weight");
  m.perc_query = upload_tensor(h, "t3.cond_enc.perceiver.pre_attention_query");
  m.perc_ln_w = upload_tensor(h, pa + "norm.weight");
  m.perc_ln_b = upload_tensor(h, pa + "norm.bias");
  m.ready = true;
}

// ---- small kernels -------------------------------------------------------------------------------
__global__ void scale_vec_kernel(const float* w, const float* scalar, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = w[i] * scalar[0];
}
__global__ void iota_kernel(int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
__global__ void last_index_kernel(const int* row_start, const int* row_len, int* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = row_start[i] + row_len[i] - 1;
}

// AttentionBlock2 of the perceiver (modules/perceiver.py:156-170): x1 [n1,1024] attends x2 [n2,1024]
static void perceiver_block(cbx_handle* h, Ctx& ctx, const float* x1, int n1, const float* x2, int n2, float* out) {
  T3Model& m = h->t3;
  const size_t mark = ctx.ws.mark();
  float* x1n = ctx.ws.get<float>((size_t)n1 * 1024);
  float* x2n = ctx.ws.get<float>((size_t)n2 * 1024);
  float* q = ctx.ws.get<float>((size_t)n1 * 1024);
  float* k = ctx.ws.get<float>((size_t)n2 * 1024);
  float* v = ctx.ws.get<float>((size_t)n2 * 1024);
  float* a = ctx.ws.get<float>((size_t)n1 * 1024);
  layernorm(ctx, x1, 1024, m.perc_ln_w.p, m.perc_ln_b.p, x1n, 1024, n1, 1024, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  layernorm(ctx, x2, 1024, m.perc_ln_w.p, m.perc_ln_b.p, x2n, 1024, n2, 1024, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  gemm(ctx, gemm_args_linear(x1n, 1024, n1, m.pq, q, 1024), m.pq);
  gemm(ctx, gemm_args_linear(x2n, 1024, n2, m.pk, k, 1024), m.pk);
  gemm(ctx, gemm_args_linear(x2n, 1024, n2, m.pv, v, 1024), m.pv);
  attention_generic(ctx, q, k, v, a, n1, n2, 4, 256, 1024, 1024, 1024, 1024, 1.0f / 16.0f);
  GemmDev g = gemm_args_linear(a, 1024, n1, m.pproj, out, 1024);
  g.res = x1; g.ldr = 1024;
  gemm(ctx, g, m.pproj);
  ctx.ws.reset(mark);
}

void t3_cond_encode(cbx_handle* h, Ctx& ctx, const float* spk, const int* prompt, int n_prompt, const float* emo,
                    int n_voices, float* cond_out) {
  T3Model& m = h->t3;
  CBX_REQUIRE(m.ready, "t3 weights not finalized");
  if (m.gpt) {
    // Turbo: [spkr_enc(speaker_emb) | speech_emb(prompt tokens)], no position table, no perceiver, no emotion row
    // (t3.py:96-100 with is_gpt; cond_enc.py:64-97 with use_perceiver_resampler / emotion_adv off)
    const int lc = 1 + n_prompt;
    for (int v = 0; v < n_voices; ++v) {
      float* out = cond_out + (size_t)v * lc * 1024;
      gemm(ctx, gemm_args_linear(spk + (size_t)v * 256, 256, 1, m.spkr, out, 1024), m.spkr);
      gather_rows(ctx, m.speech_emb.p, 1024, prompt + (size_t)v * n_prompt, out + 1024, 1024, n_prompt, 1024, nullptr, 0,
                  nullptr, m.vocab);
    }
    return;
  }
  const int len_cond = 34;
  float* emb = ctx.ws.get<float>((size_t)n_prompt * 1024);
  float* pre = ctx.ws.get<float>((size_t)32 * 1024);
  for (int v = 0; v < n_voices; ++v) {
    float* out = cond_out + (size_t)v * len_cond * 1024;
    // speech_emb(prompt) + speech_pos_emb(0..n-1)   (t3.py:97-99)
    gather_rows(ctx, m.speech_emb.p, 1024, prompt + (size_t)v * n_prompt, emb, 1024, n_prompt, 1024, m.speech_pos.p, 1024,
                nullptr, 8194);
    gemm(ctx, gemm_args_linear(spk + (size_t)v * 256, 256, 1, m.spkr, out, 1024), m.spkr);         // cond_enc.py:70
    perceiver_block(h, ctx, m.perc_query.p, 32, emb, n_prompt, pre);                                  // perceiver.py:209
    perceiver_block(h, ctx, pre, 32, pre, 32, out + 1024);                                            // perceiver.py:211
    if (!ctx.dry) {
      ctx.launches++;
      scale_vec_kernel<<<4, 256, 0, ctx.stream>>>(m.emotion_w.p, emo + v, out + (size_t)33 * 1024, 1024);  // cond_enc.py:88
    }
  }
}

static PagedKV paged_of(const cbx_t3_state& st, int n_layers) {
  PagedKV kv;
  kv.pages = st.kv_pages; kv.n_pages = st.n_pages; kv.kv_fp32 = st.kv_dtype; kv.n_layers = n_layers; kv.n_heads = 16;
  kv.page_tokens = st.page_tokens; kv.page_table = st.page_table; kv.max_pages_per_row = st.max_pages_per_row;
  return kv;
}

// ---- Turbo: one transformers GPT2Block on n rows (modeling_gpt2.py GPT2Block.forward) -----------------
//   x += c_proj(attn(c_attn(ln_1(x)))) ; x += mlp.c_proj(gelu_new(mlp.c_fc(ln_2(x))))
// `attend(att_hi, att_lo)` appends K/V to the paged cache and writes the attention output either as fp32 `att`
// (att_hi == nullptr) or as bf16 planes.
template <class Attend>
static void gpt_block(Ctx& ctx, T3Layer& ly, float* x, int n, float* xn, float* qkv, float* att, float* act, bool planes,
                      bool att_planes, Attend&& attend) {
  __nv_bfloat16* xn_hi = reinterpret_cast<__nv_bfloat16*>(xn);   __nv_bfloat16* xn_lo = xn_hi + (size_t)n * 1024;
  __nv_bfloat16* at_hi = reinterpret_cast<__nv_bfloat16*>(att);  __nv_bfloat16* at_lo = at_hi + (size_t)n * 1024;
  __nv_bfloat16* ac_hi = reinterpret_cast<__nv_bfloat16*>(act);  __nv_bfloat16* ac_lo = ac_hi + (size_t)n * 4096;
  GemmDev gq = gemm_args_linear(xn, 1024, n, ly.qkv, qkv, 3072);
  if (planes) {
    layernorm(ctx, x, 1024, ly.ln1.p, ly.ln1_b.p, nullptr, 1024, n, 1024, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr, xn_hi, xn_lo);
    gq.A = nullptr; gq.Ahi = xn_hi; gq.Alo = xn_lo; gq.ldab = 1024;
  } else {
    layernorm(ctx, x, 1024, ly.ln1.p, ly.ln1_b.p, xn, 1024, n, 1024, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  }
  gemm(ctx, gq, ly.qkv);
  const bool ap = planes && att_planes;
  attend(ap ? at_hi : nullptr, ap ? at_lo : nullptr);
  GemmDev go = gemm_args_linear(att, 1024, n, ly.o, x, 1024);
  if (ap) { go.A = nullptr; go.Ahi = at_hi; go.Alo = at_lo; go.ldab = 1024; }
  go.res = x; go.ldr = 1024;
  gemm(ctx, go, ly.o);
  GemmDev gf = gemm_args_linear(xn, 1024, n, ly.gu, act, 4096);
  GemmDev gd = gemm_args_linear(act, 4096, n, ly.down, x, 1024);
  gf.act = ACT_GELU_TANH;
  if (planes) {
    layernorm(ctx, x, 1024, ly.ln2.p, ly.ln2_b.p, nullptr, 1024, n, 1024, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr, xn_hi, xn_lo);
    gf.A = nullptr; gf.Ahi = xn_hi; gf.Alo = xn_lo; gf.ldab = 1024;
    gf.C = nullptr; gf.Chi = ac_hi; gf.Clo = ac_lo; gf.ldcb = 4096;
    gd.A = nullptr; gd.Ahi = ac_hi; gd.Alo = ac_lo; gd.ldab = 4096;
  } else {
    layernorm(ctx, x, 1024, ly.ln2.p, ly.ln2_b.p, xn, 1024, n, 1024, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  }
  gemm(ctx, gf, ly.gu);
  gd.res = x; gd.ldr = 1024;
  gemm(ctx, gd, ly.down);
}

void t3_prefill(cbx_handle* h, Ctx& ctx, const cbx_t3_state& st, int n_tok, const int* tok_row, const int* tok_pos,
                const int* row_start, const int* row_len, int max_row_len, const float* cond, const int* row_voice,
                int len_cond, const int* text_flat, const int* text_start, const int* n_text, const int* row_uncond) {
  T3Model& m = h->t3;
  CBX_REQUIRE(m.ready, "t3 weights not finalized");
  const int R = st.n_rows;
  float* x = ctx.ws.get<float>((size_t)n_tok * 1024);
  float* xn = ctx.ws.get<float>((size_t)n_tok * 1024);
  float* qkv = ctx.ws.get<float>((size_t)n_tok * 3072);
  float* att = ctx.ws.get<float>((size_t)n_tok * 1024);
  float* act = ctx.ws.get<float>((size_t)n_tok * 4096);
  float* hn = ctx.ws.get<float>((size_t)R * 1024);
  int* last = ctx.ws.get<int>(R);
  t3_embed(ctx, x, n_tok, tok_row, tok_pos, cond, row_voice, len_cond, text_flat, text_start, n_text, row_uncond,
           m.text_emb.p, m.text_vocab, m.text_pos.p, m.speech_emb.p, m.speech_pos.p, 6561, m.gpt ? m.wpe.p : nullptr);
  PagedKV kv = paged_of(st, m.n_layers);
  // tensor-core batches: GEMM operands that a norm / SwiGLU epilogue produces travel as bf16 hi/lo planes (see t3_decode)
  const bool planes = (n_tok > 8 && ctx.gemm_impl == 0);
  __nv_bfloat16* xn_hi = reinterpret_cast<__nv_bfloat16*>(xn);   __nv_bfloat16* xn_lo = xn_hi + (size_t)n_tok * 1024;
  __nv_bfloat16* ac_hi = reinterpret_cast<__nv_bfloat16*>(act);  __nv_bfloat16* ac_lo = ac_hi + (size_t)n_tok * 4096;
  for (int l = 0; l < m.n_layers; ++l) {
    T3Layer& ly = m.layers[l];
    if (m.gpt) {
      gpt_block(ctx, ly, x, n_tok, xn, qkv, att, act, planes, false, [&](__nv_bfloat16*, __nv_bfloat16*) {
        rope_and_store_kv(ctx, qkv, 3072, kv, l, tok_row, tok_pos, 0, n_tok, m.rope_cos.p, m.rope_sin.p);   // identity rotation
        AttnArgs a;
        a.Q = qkv; a.K = qkv + 1024; a.V = qkv + 2048; a.ldq = a.ldk = a.ldv = 3072; a.O = att; a.ldo = 1024;
        a.n_seq = R; a.n_heads = 16; a.q_start = row_start; a.q_len = row_len; a.kv_start = row_start; a.kv_len = row_len;
        a.max_q_len = max_row_len; a.scale = 0.125f; a.causal = 1;
        attention(ctx, a);
      });
      continue;
    }
    if (planes) {
      rmsnorm(ctx, x, 1024, ly.ln1.p, nullptr, 1024, n_tok, 1024, 1e-5f, nullptr, xn_hi, xn_lo);
      GemmDev gq = gemm_args_linear(nullptr, 1024, n_tok, ly.qkv, qkv, 3072);
      gq.Ahi = xn_hi; gq.Alo = xn_lo; gq.ldab = 1024;
      gemm(ctx, gq, ly.qkv);
    } else {
      rmsnorm(ctx, x, 1024, ly.ln1.p, xn, 1024, n_tok, 1024, 1e-5f, nullptr);
      gemm(ctx, gemm_args_linear(xn, 1024, n_tok, ly.qkv, qkv, 3072), ly.qkv);
    }
    rope_and_store_kv(ctx, qkv, 3072, kv, l, tok_row, tok_pos, 0, n_tok, m.rope_cos.p, m.rope_sin.p);
    AttnArgs a;
    a.Q = qkv; a.K = qkv + 1024; a.V = qkv + 2048; a.ldq = a.ldk = a.ldv = 3072; a.O = att; a.ldo = 1024;
    a.n_seq = R; a.n_heads = 16; a.q_start = row_start; a.q_len = row_len; a.kv_start = row_start; a.kv_len = row_len;
    a.max_q_len = max_row_len; a.scale = 0.125f; a.causal = 1;
    attention(ctx, a);
    GemmDev go = gemm_args_linear(att, 1024, n_tok, ly.o, x, 1024);
    go.res = x; go.ldr = 1024;
    gemm(ctx, go, ly.o);
    if (planes) {
      rmsnorm(ctx, x, 1024, ly.ln2.p, nullptr, 1024, n_tok, 1024, 1e-5f, nullptr, xn_hi, xn_lo);
      GemmDev gg = gemm_args_linear(nullptr, 1024, n_tok, ly.gu, nullptr, 0);
      gg.Ahi = xn_hi; gg.Alo = xn_lo; gg.ldab = 1024;
      gg.swiglu = 1; gg.Chi = ac_hi; gg.Clo = ac_lo; gg.ldcb = 4096;
      gemm(ctx, gg, ly.gu);
      GemmDev gd = gemm_args_linear(nullptr, 4096, n_tok, ly.down, x, 1024);
      gd.Ahi = ac_hi; gd.Alo = ac_lo; gd.ldab = 4096;
      gd.res = x; gd.ldr = 1024;
      gemm(ctx, gd, ly.down);
      continue;
    }
    rmsnorm(ctx, x, 1024, ly.ln2.p, xn, 1024, n_tok, 1024, 1e-5f, nullptr);
    GemmDev gg = gemm_args_linear(xn, 1024, n_tok, ly.gu, act, 4096);
    gg.swiglu = 1;
    gemm(ctx, gg, ly.gu);
    GemmDev gd = gemm_args_linear(act, 4096, n_tok, ly.down, x, 1024);
    gd.res = x; gd.ldr = 1024;
    gemm(ctx, gd, ly.down);
  }
  if (!ctx.dry) {
    ctx.launches++;
    last_index_kernel<<<(R + 127) / 128, 128, 0, ctx.stream>>>(row_start, row_len, last, R);
  }
  if (m.gpt) {        // ln_f on the last position of every row, then the speech head (with bias)
    float* hl = ctx.ws.get<float>((size_t)R * 1024);
    gather_rows(ctx, x, 1024, last, hl, 1024, R, 1024, nullptr, 0, nullptr, n_tok);
    layernorm(ctx, hl, 1024, m.final_norm.p, m.final_norm_b.p, hn, 1024, R, 1024, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  } else {
    rmsnorm(ctx, x, 1024, m.final_norm.p, hn, 1024, R, 1024, 1e-5f, last);
  }
  gemm(ctx, gemm_args_linear(hn, 1024, R, m.head, st.logits, st.ldl), m.head);
}

// ---- decode step -------------------------------------------------------------------------------------
// One step of the sampling loop of T3.inference (t3.py:338-386) / T3.inference_turbo (t3.py:426-461) for `cap` slots:
//   t3_compact (device-side retirement) -> sampler (token i, next input embedding) ->
//   L x [ resid_norm -> QKV GEMM -> paged attention with fused RoPE + KV append (bulk-copy staged K/V) ->
//         O GEMM (split-K partials) -> resid_norm -> gate/up GEMM (SwiGLU / gelu_new epilogue) -> down GEMM (split-K) ]
//   -> resid_norm (final norm) -> speech head.
// Every launch argument of a step is constant across steps (the moving parts -- active list, positions, tokens, done
// flags -- live in device memory), so a step is captured once into a CUDA graph per (state, capacity) and replayed.
struct DecodeTiles { int qkv_bn, qkv_dual, o_bn, o_split, gu_bn, gu_dual, down_bn, down_split, head_bn; };

static DecodeTiles decode_tiles(int S, bool f16) {
  // one-tile-per-CTA kernel (fp32-faithful planes): tuned by sweeps in round 2's first sessions; env overrides for sweeps
  DecodeTiles t;
  const int mt = (S + 127) / 128;
  t.qkv_bn = mt >= 3 ? 64 : 64;  t.qkv_dual = mt >= 4 ? 1 : 0;
  t.o_bn = 64;   t.o_split = mt >= 3 ? 2 : 4;
  t.gu_bn = mt >= 2 ? 128 : 64;  t.gu_dual = mt >= 3 ? 1 : 0;
  t.down_bn = 64; t.down_split = mt >= 3 ? 4 : 8;
  t.head_bn = 0;
  if (f16) {
    // persistent streaming kernel (fp16 plane): tools/decode_gemm_bench.py at 512 rows (profiles/r2_decode_gemm_bench.txt):
    // qkv 12.4 us at BN 128 (16.5 at 64), gate/up 18.5 at BN 256 (20.5 at 128), down 13.9 at BN 128 split 4 (17.7 at BN 64)
    // and at 128 / 256 / 384 rows (session 22): gate/up 14.4 us at BN 128 up to 256 rows (18.5 at 256), down 10.6 us with split 8 (12.5 with 4)
    t.qkv_bn = 128; t.gu_bn = mt >= 3 ? 256 : 128; t.down_bn = 128; t.down_split = mt >= 3 ? 4 : 8; t.o_split = mt >= 3 ? 2 : 4;
  }
  static const char* ov = getenv("CBX_DECODE_TILES");     // "qkv_bn,qkv_dual,o_bn,o_split,gu_bn,gu_dual,down_bn,down_split"
  if (ov) {
    int v[8];
    if (sscanf(ov, "%d,%d,%d,%d,%d,%d,%d,%d", v, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6, v + 7) == 8) {
      t.qkv_bn = v[0]; t.qkv_dual = v[1]; t.o_bn = v[2]; t.o_split = v[3]; t.gu_bn = v[4]; t.gu_dual = v[5]; t.down_bn = v[6]; t.down_split = v[7];
    }
  }
  return t;
}

static void decode_step(cbx_handle* h, Ctx& ctx, const cbx_t3_state& st, int cap) {
  T3Model& m = h->t3;
  const int rows_per = st.cfg ? 2 : 1;
  const int S = cap * rows_per;
  const size_t mark = ctx.ws.mark();
  const DecodeTiles tl = decode_tiles(S, S > 8 && ctx.gemm_impl == 0 && st.act_fp16 != 0);
  const bool pdl_saved = ctx.pdl;
  ctx.pdl = h->decode_pdl != 0;     // every kernel of the step waits (griddepcontrol.wait) before its first global access
  const int max_split = 8;
  float* xn = ctx.ws.get<float>((size_t)S * 1024);
  float* qkv = ctx.ws.get<float>((size_t)S * 3072);
  float* att = ctx.ws.get<float>((size_t)S * 1024);
  float* act = ctx.ws.get<float>((size_t)S * 4096);
  float* part = ctx.ws.get<float>((size_t)max_split * S * 1024);
  int nsplit = 1;
  if (S * 16 < 444) { nsplit = (444 + S * 16 - 1) / (S * 16); if (nsplit > 16) nsplit = 16; }     // 3 CTAs per SM
  float* scratch = ctx.ws.get<float>((size_t)S * 16 * nsplit * 66);
  PagedKV kv = paged_of(st, m.n_layers);
  T3SampleDev sp;
  memset(&sp, 0, sizeof(sp));
  sp.logits = st.logits; sp.ldl = st.ldl; sp.act_utt = st.act_utt; sp.n_act = st.n_act; sp.src_slot = st.src_slot;
  sp.force_tokens = st.force_tokens; sp.sampled_out = st.sampled_out;
  sp.cfg = st.cfg; sp.n_utts = st.n_utts;
  sp.cfg_weight = st.cfg_weight; sp.rep_penalty = st.rep_penalty; sp.temperature = st.temperature;
  sp.min_p = st.min_p; sp.top_p = st.top_p; sp.eos_id = 6562;
  sp.tokens = st.tokens; sp.max_tokens = st.max_tokens; sp.n_gen = st.n_gen; sp.max_new = st.max_new; sp.done = st.done;
  sp.seen = st.seen; sp.positions = st.positions; sp.base_pos = st.base_pos; sp.x = st.x;
  sp.speech_emb = m.speech_emb.p; sp.speech_pos = m.speech_pos.p; sp.q_noise = st.q_noise; sp.seed = st.seed;
  sp.vocab = m.vocab; sp.turbo = m.gpt ? 1 : 0; sp.top_k = st.top_k; sp.bos_id = 6561; sp.wpe = m.gpt ? m.wpe.p : nullptr;
  float* x = st.x;
  // Tensor-core batches (S > 8): every GEMM operand travels as bf16 hi/lo planes written by its producer (resid_norm,
  // paged attention, SwiGLU epilogue) and is loaded by TMA.  Small batches keep fp32 activations for the GEMV kernels.
  const bool planes = (S > 8 && ctx.gemm_impl == 0);
  // st.act_fp16 (throughput mode, together with a bf16 / fp8 KV cache): the same dataflow with ONE fp16 plane per activation
  // against the fp16 copy of the weights -- half the MMAs and half the activation bytes; its error (2^-12 relative per
  // element) is of the order of the bf16 rounding of K / V that mode already accepts.  fp32-faithful hi/lo planes otherwise.
  const bool f16 = planes && st.act_fp16 != 0;
  const int* m_live = planes ? st.m_live : nullptr;
  __half* xn16 = reinterpret_cast<__half*>(xn);
  __half* at16 = reinterpret_cast<__half*>(att);
  __half* ac16 = reinterpret_cast<__half*>(act);
  __nv_bfloat16* xn_hi = reinterpret_cast<__nv_bfloat16*>(xn);   __nv_bfloat16* xn_lo = xn_hi + (size_t)S * 1024;
  __nv_bfloat16* at_hi = reinterpret_cast<__nv_bfloat16*>(att);  __nv_bfloat16* at_lo = at_hi + (size_t)S * 1024;
  __nv_bfloat16* ac_hi = reinterpret_cast<__nv_bfloat16*>(act);  __nv_bfloat16* ac_lo = ac_hi + (size_t)S * 4096;
  PagedOpts po;
  po.fuse_rope = 1; po.cos_t = m.rope_cos.p; po.sin_t = m.rope_sin.p; po.n_live = st.m_live;

  t3_compact(ctx, st.act_utt, st.n_act, st.src_slot, st.slot_row, st.m_live, st.done, rows_per);
  t3_sample(ctx, sp, cap);
  // norm feeding a GEMM: x (+= split-K partials of the previous projection) -> norm -> planes / fp32
  // GEMV path (<= 8 rows): no separate norm launches -- the o / down GEMVs add their result into x in the epilogue and
  // the qkv / gate-up / head GEMVs normalise the raw residual stream in their prologue
  const DevVec* fused_w = nullptr; const DevVec* fused_b = nullptr;
  auto norm_in = [&](const float* prt, int ns, const float* bias, const DevVec& w, const DevVec& b) {
    if (!planes) { fused_w = &w; fused_b = &b; return; }
    ResidNormDev rn;
    memset(&rn, 0, sizeof(rn));
    rn.x = x; rn.ldx = 1024; rn.part = prt; rn.nsplit = ns; rn.split_stride = (long)S * 1024; rn.ldp = 1024; rn.bias = bias;
    rn.w = w.p; rn.b = b.p; rn.layernorm = m.gpt ? 1 : 0; rn.eps = 1e-5f;
    if (f16) rn.y16 = xn16; else if (planes) { rn.yhi = xn_hi; rn.ylo = xn_lo; } else rn.y = xn;
    rn.ldy = 1024; rn.dim = 1024; rn.m_live = st.m_live;
    resid_norm(ctx, rn, S);
  };
  auto feed = [&](GemmDev& g, const float* a32, __nv_bfloat16* hi, __nv_bfloat16* lo, int ld) {
    if (f16) { g.A = nullptr; g.A16 = reinterpret_cast<const __half*>(hi); g.lda16 = ld; }
    else if (planes) { g.A = nullptr; g.Ahi = hi; g.Alo = lo; g.ldab = ld; }
    else if (a32 == xn) {          // the consumer of a norm: raw residual stream + fused row norm
      g.A = x; g.lda = 1024; g.norm_w = fused_w->p; g.norm_b = fused_b->p; g.norm_ln = m.gpt ? 1 : 0; g.norm_eps = 1e-5f;
    } else g.A = a32;
    g.m_live = m_live;
  };
  const float* pend = nullptr; int pend_ns = 0; const float* pend_bias = nullptr;   // projection output not yet added to x
  for (int l = 0; l < m.n_layers; ++l) {
    T3Layer& ly = m.layers[l];
    norm_in(pend, pend_ns, pend_bias, ly.ln1, ly.ln1_b);
    GemmDev gq = gemm_args_linear(xn, 1024, S, ly.qkv, qkv, 3072);
    feed(gq, xn, xn_hi, xn_lo, 1024);
    gq.tile_bn = tl.qkv_bn; gq.tile_dual = tl.qkv_dual;
    gemm(ctx, gq, ly.qkv);
    po.out16 = f16 ? at16 : nullptr;
    paged_decode_attention(ctx, qkv, 3072, kv, l, st.slot_row, S, st.positions, planes ? nullptr : att, 1024, scratch, nsplit,
                           (planes && !f16) ? at_hi : nullptr, (planes && !f16) ? at_lo : nullptr, &po);
    GemmDev go = gemm_args_linear(att, 1024, S, ly.o, planes ? part : x, 1024);
    feed(go, att, at_hi, at_lo, 1024);
    const float* o_bias = m.gpt ? ly.o.bias : nullptr;
    if (planes) { go.bias = nullptr; go.splitk = tl.o_split; go.split_stride = (long)S * 1024; go.tile_bn = tl.o_bn; }
    else { go.res = x; go.ldr = 1024; }            // GEMV: x += o(att) (+ bias) in the epilogue
    gemm(ctx, go, ly.o);
    norm_in(part, planes ? tl.o_split : 1, o_bias, ly.ln2, ly.ln2_b);
    GemmDev gg = gemm_args_linear(xn, 1024, S, ly.gu, act, 4096);
    feed(gg, xn, xn_hi, xn_lo, 1024);
    if (m.gpt) gg.act = ACT_GELU_TANH; else gg.swiglu = 1;
    if (planes) { gg.C = nullptr; gg.Chi = ac_hi; gg.Clo = ac_lo; gg.ldcb = 4096; gg.c_half = f16 ? 1 : 0; }
    gg.tile_bn = tl.gu_bn; gg.tile_dual = tl.gu_dual;
    gemm(ctx, gg, ly.gu);
    GemmDev gd = gemm_args_linear(act, 4096, S, ly.down, planes ? part : x, 1024);
    feed(gd, act, ac_hi, ac_lo, 4096);
    pend_bias = m.gpt ? ly.down.bias : nullptr;
    if (planes) { gd.bias = nullptr; gd.splitk = tl.down_split; gd.split_stride = (long)S * 1024; gd.tile_bn = tl.down_bn; }
    else { gd.res = x; gd.ldr = 1024; }
    gemm(ctx, gd, ly.down);
    pend = part; pend_ns = planes ? tl.down_split : 1;
  }
  norm_in(pend, pend_ns, pend_bias, m.final_norm, m.final_norm_b);
  GemmDev gh = gemm_args_linear(xn, 1024, S, m.head, st.logits, st.ldl);
  feed(gh, xn, xn_hi, xn_lo, 1024);
  gemm(ctx, gh, m.head);
  ctx.pdl = pdl_saved;
  ctx.ws.reset(mark);
}

static uint64_t fnv1a(const void* p, size_t n, uint64_t hsh = 1469598103934665603ull) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) { hsh ^= b[i]; hsh *= 1099511628211ull; }
  return hsh;
}

void t3_decode(cbx_handle* h, Ctx& ctx, const cbx_t3_state& st, int cap, int n_steps) {
  T3Model& m = h->t3;
  CBX_REQUIRE(m.ready, "t3 weights not finalized");
  CBX_REQUIRE((st.sampler != 0) == m.gpt, "cbx_t3_state.sampler does not match the loaded backbone (1 = Turbo)");
  CBX_REQUIRE(!(m.gpt && st.cfg), "the Turbo backbone runs without CFG rows");
  CBX_REQUIRE(cap >= 1 && cap <= st.n_utts, "decode capacity");
  if (ctx.dry) { decode_step(h, ctx, st, cap); return; }
  CBX_REQUIRE(st.act_utt && st.n_act && st.src_slot && st.slot_row && st.m_live, "cbx_t3_state: device-side slot bookkeeping buffers");
  // One executable graph per (state, capacity, workspace): captured from a single step, replayed for every step.
  // Needs a capturable (non-legacy) stream and no per-launch event timer.
  const bool use_graph = h->decode_graph && !ctx.timer && ctx.stream != nullptr && ctx.stream != cudaStreamLegacy;
  if (!use_graph) {
    for (int step = 0; step < n_steps; ++step) decode_step(h, ctx, st, cap);
    return;
  }
  uint64_t key = fnv1a(&st, sizeof(st));
  const void* wsb = ctx.ws.base; const size_t wsc = ctx.ws.cap;
  key = fnv1a(&cap, sizeof(cap), key); key = fnv1a(&wsb, sizeof(wsb), key); key = fnv1a(&wsc, sizeof(wsc), key);
  key = fnv1a(&ctx.gemm_impl, sizeof(int), key); key = fnv1a(&h->decode_pdl, sizeof(int), key);
  auto it = h->decode_graphs.find(key);
  if (it == h->decode_graphs.end()) {
    if (h->decode_graphs.size() >= 48) {          // bound the cache: drop everything (states of finished batches)
      for (auto& kvp : h->decode_graphs) cudaGraphExecDestroy(kvp.second.exec);
      h->decode_graphs.clear();
    }
    const long before = ctx.launches;
    cudaGraph_t graph = nullptr;
    CBX_CHECK(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal));
    try {
      decode_step(h, ctx, st, cap);
    } catch (...) {
      cudaStreamEndCapture(ctx.stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    CBX_CHECK(cudaStreamEndCapture(ctx.stream, &graph));
    DecodeGraph dg;
    dg.launches = ctx.launches - before;
    ctx.launches = before;
    cudaError_t ge = cudaGraphInstantiate(&dg.exec, graph, 0);
    cudaGraphDestroy(graph);
    CBX_CHECK(ge);
    it = h->decode_graphs.emplace(key, dg).first;
  }
  for (int step = 0; step < n_steps; ++step) CBX_CHECK(cudaGraphLaunch(it->second.exec, ctx.stream));
  ctx.launches += it->second.launches * (long)n_steps;
}

}  // namespace cbx
