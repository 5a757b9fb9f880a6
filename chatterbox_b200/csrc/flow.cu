// S3Gen token->mel: upsampling conformer encoder + CFM Euler solver over the causal-conv U-Net estimator.
// Reference: src/chatterbox/models/s3gen/{flow.py, transformer/upsample_encoder.py, transformer/attention.py,
// flow_matching.py, decoder.py, matcha/decoder.py, matcha/transformer.py}.
// Everything runs on packed channel-last buffers [rows, C] (see cbx_layout); convs are implicit GEMMs.
#include "engine.h"
#include <cmath>
#include <cstdlib>

namespace cbx {

GemmDev conv_args(const float* A, int lda, const Weight& W, int c_in, int ntaps, int dil, int pad, int stride,
                  const cbx_layout& out, const cbx_layout& in, float* C, int ldc) {
  GemmDev g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.lda = lda; g.M = out.rows; g.M_in = in.rows;
  g.a_mode = A_TAPS; g.ntaps = ntaps; g.ctap = (c_in + 63) / 64 * 64; g.c_in = c_in; g.dil = dil; g.pad = pad; g.stride = stride;
  g.k_total = ntaps * g.ctap;
  g.has_seq = 1; g.seq = seqmap(out, in);
  g.Wp = W.w; g.Kpad = W.Kpad; g.Npad = W.Npad;
  g.C = C; g.ldc = ldc; g.n_out = W.N; g.bias = W.bias; g.alpha = 1.0f; g.act = ACT_NONE; g.out_scale = 1.0f;
  return g;
}
static GemmDev lin_args(const float* A, int lda, int c_in, int M, const Weight& W, float* C, int ldc) {
  GemmDev g = gemm_args_linear(A, lda, M, W, C, ldc);
  g.c_in = c_in;
  return g;
}

static void pack_lin(cbx_handle* h, Weight& W, const std::string& name, bool bias = true, bool half_copy = false) {
  const HostTensor& w = host_tensor(h, name + ".weight");
  const float* b = (bias && has_tensor(h, name + ".bias")) ? host_tensor(h, name + ".bias").data.data() : nullptr;
  pack_linear(W, w.data.data(), b, (int)w.shape[0], (int)w.shape[1], half_copy);
}
static void pack_conv(cbx_handle* h, Weight& W, const std::string& name) {
  const HostTensor& w = host_tensor(h, name + ".weight");
  const float* b = has_tensor(h, name + ".bias") ? host_tensor(h, name + ".bias").data.data() : nullptr;
  pack_conv_taps(W, w.data.data(), b, (int)w.shape[0], (int)w.shape[1], (int)w.shape[2]);
}
static void pack_cat3(cbx_handle* h, Weight& W, const std::string& a, const std::string& b, const std::string& c, bool bias,
                      bool half_copy = false) {
  const HostTensor &wa = host_tensor(h, a + ".weight"), &wb = host_tensor(h, b + ".weight"), &wc = host_tensor(h, c + ".weight");
  std::vector<float> w;
  w.insert(w.end(), wa.data.begin(), wa.data.end());
  w.insert(w.end(), wb.data.begin(), wb.data.end());
  w.insert(w.end(), wc.data.begin(), wc.data.end());
  std::vector<float> bb;
  if (bias) {
    for (const std::string& n : {a, b, c}) { const auto& t = host_tensor(h, n + ".bias").data; bb.insert(bb.end(), t.begin(), t.end()); }
  }
  const int N = (int)(wa.shape[0] + wb.shape[0] + wc.shape[0]);
  pack_linear(W, w.data(), bias ? bb.data() : nullptr, N, (int)wa.shape[1], half_copy);
}

static void build_enc_layer(cbx_handle* h, EncLayer& L, const std::string& p) {
  pack_cat3(h, L.qkv, p + "self_attn.linear_q", p + "self_attn.linear_k", p + "self_attn.linear_v", true);
  pack_lin(h, L.out, p + "self_attn.linear_out");
  pack_lin(h, L.pos, p + "self_attn.linear_pos", false);
  pack_lin(h, L.w1, p + "feed_forward.w_1");
  pack_lin(h, L.w2, p + "feed_forward.w_2");
  L.ln_mha_w = upload_tensor(h, p + "norm_mha.weight"); L.ln_mha_b = upload_tensor(h, p + "norm_mha.bias");
  L.ln_ff_w = upload_tensor(h, p + "norm_ff.weight"); L.ln_ff_b = upload_tensor(h, p + "norm_ff.bias");
  L.bias_u = upload_tensor(h, p + "self_attn.pos_bias_u"); L.bias_v = upload_tensor(h, p + "self_attn.pos_bias_v");
}
static void build_resnet(cbx_handle* h, CfmResnet& r, const std::string& p) {
  pack_conv(h, r.conv1, p + "block1.block.0");
  pack_conv(h, r.conv2, p + "block2.block.0");
  pack_conv(h, r.res, p + "res_conv");
  pack_lin(h, r.mlp, p + "mlp.1");
  r.ln1_w = upload_tensor(h, p + "block1.block.2.weight"); r.ln1_b = upload_tensor(h, p + "block1.block.2.bias");
  r.ln2_w = upload_tensor(h, p + "block2.block.2.weight"); r.ln2_b = upload_tensor(h, p + "block2.block.2.bias");
}
static void build_tfmr(cbx_handle* h, CfmTfmr& t, const std::string& p) {
  // fp16 copies: operands of the single-plane fp16-activation format of the block GEMMs (Engine.set_cfm_activation_precision)
  pack_cat3(h, t.qkv, p + "attn1.to_q", p + "attn1.to_k", p + "attn1.to_v", false, true);
  pack_lin(h, t.out, p + "attn1.to_out.0", true, true);
  pack_lin(h, t.ff1, p + "ff.net.0.proj", true, true);
  pack_lin(h, t.ff2, p + "ff.net.2", true, true);
  t.ln1_w = upload_tensor(h, p + "norm1.weight"); t.ln1_b = upload_tensor(h, p + "norm1.bias");
  t.ln3_w = upload_tensor(h, p + "norm3.weight"); t.ln3_b = upload_tensor(h, p + "norm3.bias");
}
static void build_stage(cbx_handle* h, CfmStage& s, const std::string& p) {
  build_resnet(h, s.res, p + "0.");
  for (int j = 0; j < 4; ++j) build_tfmr(h, s.t[j], p + "1." + std::to_string(j) + ".");
}

void flow_finalize(cbx_handle* h) {
  FlowModel& m = h->flow;
  const std::string f = "flow.";
  m.input_embedding = upload_tensor(h, f + "input_embedding.weight");
  pack_lin(h, m.spk_affine, f + "spk_embed_affine_layer");
  pack_lin(h, m.enc_proj, f + "encoder_proj");
  auto emb = [&](EncEmbed& e, const std::string& p) {
    pack_lin(h, e.lin, p + "out.0");
    e.ln_w = upload_tensor(h, p + "out.1.weight"); e.ln_b = upload_tensor(h, p + "out.1.bias");
  };
  emb(m.embed, f + "encoder.embed.");
  emb(m.up_embed, f + "encoder.up_embed.");
  pack_conv(h, m.pre_conv1, f + "encoder.pre_lookahead_layer.conv1");
  pack_conv(h, m.pre_conv2, f + "encoder.pre_lookahead_layer.conv2");
  pack_conv(h, m.up_conv, f + "encoder.up_layer.conv");
  for (int i = 0; i < 6; ++i) build_enc_layer(h, m.enc[i], f + "encoder.encoders." + std::to_string(i) + ".");
  for (int i = 0; i < 4; ++i) build_enc_layer(h, m.up_enc[i], f + "encoder.up_encoders." + std::to_string(i) + ".");
  m.pe_table = upload_tensor(h, f + "pe_table");                 // [2*max_len-1][512], built by the Python shim
  m.pe_center = ((int)host_tensor(h, f + "pe_table").shape[0] - 1) / 2;
  m.after_w = upload_tensor(h, f + "encoder.after_norm.weight"); m.after_b = upload_tensor(h, f + "encoder.after_norm.bias");
  const std::string e = f + "decoder.estimator.";
  pack_lin(h, m.time1, e + "time_mlp.linear_1");
  pack_lin(h, m.time2, e + "time_mlp.linear_2");
  m.meanflow = has_tensor(h, e + "time_embed_mixer.weight");
  if (m.meanflow) pack_lin(h, m.time_mixer, e + "time_embed_mixer", false);
  build_stage(h, m.down, e + "down_blocks.0.");
  pack_conv(h, m.down_conv, e + "down_blocks.0.2");
  for (int i = 0; i < 12; ++i) build_stage(h, m.mid[i], e + "mid_blocks." + std::to_string(i) + ".");
  build_stage(h, m.up, e + "up_blocks.0.");
  pack_conv(h, m.up_conv2, e + "up_blocks.0.2");
  pack_conv(h, m.final_conv, e + "final_block.block.0");
  m.final_ln_w = upload_tensor(h, e + "final_block.block.2.weight"); m.final_ln_b = upload_tensor(h, e + "final_block.block.2.bias");
  pack_conv(h, m.final_proj, e + "final_proj");
  m.ready = true;
}

// ---- small kernels -------------------------------------------------------------------------------
__global__ void l2norm_rows_kernel(const float* x, float* y, int dim) {   // F.normalize(dim=1, eps=1e-12)
  const int r = blockIdx.x;
  float ss = 0.f;
  for (int i = threadIdx.x; i < dim; i += 32) ss += x[(long)r * dim + i] * x[(long)r * dim + i];
  ss = warp_sum(ss);
  const float d = fmaxf(sqrtf(ss), 1e-12f);
  for (int i = threadIdx.x; i < dim; i += 32) y[(long)r * dim + i] = x[(long)r * dim + i] / d;
}
struct TVals { float v[32]; };
__global__ void tvals_kernel(TVals tv, float* out, int n) { if (threadIdx.x < n) out[threadIdx.x] = tv.v[threadIdx.x]; }

// ---- conformer encoder ---------------------------------------------------------------------------
static void encoder_layer(cbx_handle* h, Ctx& ctx, EncLayer& L, float* x, const cbx_layout& lay, const float* pe, int Tmax) {
  const int rows = lay.rows;
  const size_t mark = ctx.ws.mark();
  float* hn = ctx.ws.get<float>((size_t)rows * 512);
  float* qkv = ctx.ws.get<float>((size_t)rows * 1536);
  float* qu = ctx.ws.get<float>((size_t)rows * 512);
  float* qv = ctx.ws.get<float>((size_t)rows * 512);
  float* att = ctx.ws.get<float>((size_t)rows * 512);
  const int NP = 2 * Tmax - 1;
  float* P = ctx.ws.get<float>((size_t)NP * 512);
  layernorm(ctx, x, 512, L.ln_mha_w.p, L.ln_mha_b.p, hn, 512, rows, 512, 1e-12f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  gemm(ctx, gemm_args_linear(hn, 512, rows, L.qkv, qkv, 1536), L.qkv);
  gemm(ctx, gemm_args_linear(pe, 512, NP, L.pos, P, 512), L.pos);                    // linear_pos(pos_emb)
  add_pos_bias(ctx, qkv, 1536, L.bias_u.p, L.bias_v.p, qu, qv, rows);
  // matrix_bd = (q + v) . P^T per head, P as an on-the-fly hi/lo bf16 "weight"  (attention.py:314-318)
  const int Npad = (NP + 63) / 64 * 64;
  __nv_bfloat16* Pcat = ctx.ws.get<__nv_bfloat16>((size_t)8 * Npad * 128);      // per head [Npad][hi(64) | lo(64)]
  // chunk the sequences so the materialised bias stays below ~1.5 GB
  const long budget_rows = std::max<long>(kTileM, (long)(1.5e9 / ((double)8 * Npad * 4)) / kTileM * kTileM);
  const int chunk_rows_cap = (int)std::min<long>(rows, budget_rows);
  float* bd = ctx.ws.get<float>((size_t)8 * chunk_rows_cap * Npad);
  for (int hd = 0; hd < 8; ++hd) pack_hilo_cat(ctx, P + hd * 64, 512, NP, 64, Pcat + (size_t)hd * Npad * 128, Npad);
  int s0 = 0;
  while (s0 < lay.n_seq) {
    int s1 = s0; long r0 = lay.h_start[s0], r1 = r0;
    while (s1 < lay.n_seq) {
      long e = (long)lay.h_start[s1] + ((lay.h_len[s1] + kTileM - 1) / kTileM) * kTileM;
      if (s1 > s0 && e - r0 > chunk_rows_cap) break;
      r1 = e; ++s1;
    }
    const int crow = (int)(r1 - r0);
    CBX_REQUIRE(crow <= chunk_rows_cap, "sequence longer than the rel-pos bias chunk");
    for (int hd = 0; hd < 8; ++hd) {
      // one GEMM with K = 128: the 64 q_v channels are fed twice (tap stride 0) against [P_hi | P_lo]
      Weight W;
      W.w = Pcat + (size_t)hd * Npad * 128;
      W.N = NP; W.K = 128; W.Npad = Npad; W.Kpad = 128; W.bias = nullptr;
      if (!ctx.dry) make_tmaps_for(W);
      GemmDev g = gemm_args_linear(qv + r0 * 512 + hd * 64, 512, crow, W, bd + (size_t)hd * crow * Npad, Npad);
      g.bias = nullptr; g.ntaps = 2; g.ctap = 64; g.c_in = 64; g.dil = 0; g.k_total = 128;
      gemm(ctx, g, W);
    }
    AttnArgs a;
    a.Q = qu; a.K = qkv + 512; a.V = qkv + 1024; a.ldq = 512; a.ldk = a.ldv = 1536; a.O = att; a.ldo = 512;
    a.n_seq = s1 - s0; a.n_heads = 8; a.q_start = lay.start + s0; a.q_len = lay.len + s0;
    a.kv_start = lay.start + s0; a.kv_len = lay.len + s0; a.max_q_len = lay.max_len; a.scale = 0.125f; a.causal = 0;
    a.bias = bd; a.bias_head_stride = (long)crow * Npad; a.bias_ld = Npad; a.bias_row0 = r0; a.bias_rel = 1;
    a.bias_center = Tmax - 1;
    attention(ctx, a);
    s0 = s1;
  }
  GemmDev go = gemm_args_linear(att, 512, rows, L.out, x, 512);
  go.res = x; go.ldr = 512;
  gemm(ctx, go, L.out);
  ctx.ws.reset(mark);      // stream order protects the buffers that are still being read
  hn = ctx.ws.get<float>((size_t)rows * 512);
  float* ff = ctx.ws.get<float>((size_t)rows * 2048);
  layernorm(ctx, x, 512, L.ln_ff_w.p, L.ln_ff_b.p, hn, 512, rows, 512, 1e-12f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  GemmDev g1 = gemm_args_linear(hn, 512, rows, L.w1, ff, 2048);
  g1.act = ACT_SILU;
  gemm(ctx, g1, L.w1);
  GemmDev g2 = gemm_args_linear(ff, 2048, rows, L.w2, x, 512);
  g2.res = x; g2.ldr = 512;
  gemm(ctx, g2, L.w2);
  ctx.ws.reset(mark);
}

static void enc_embed(Ctx& ctx, EncEmbed& e, const float* in, float* out, int rows) {
  const size_t mark = ctx.ws.mark();
  float* t = ctx.ws.get<float>((size_t)rows * 512);
  gemm(ctx, gemm_args_linear(in, 512, rows, e.lin, t, 512), e.lin);
  layernorm(ctx, t, 512, e.ln_w.p, e.ln_b.p, out, 512, rows, 512, 1e-5f, ACT_NONE, sqrtf(512.0f), nullptr, 0, nullptr);
  ctx.ws.reset(mark);
}

void flow_encode(cbx_handle* h, Ctx& ctx, const int* tokens, const cbx_layout& L1, const cbx_layout& L2,
                 const float* xvec, float* mu, float* spk) {
  FlowModel& m = h->flow;
  CBX_REQUIRE(m.ready, "flow weights not finalized");
  CBX_REQUIRE(L1.h_start && L1.h_len && L2.h_start && L2.h_len, "layouts need host copies");
  const int B = L1.n_seq, rows1 = L1.rows, rows2 = L2.rows;
  // speaker: F.normalize(xvec) -> Linear(192->80)   (flow.py:149-151)
  float* xn = ctx.ws.get<float>((size_t)B * 192);
  if (!ctx.dry) { ctx.launches++; l2norm_rows_kernel<<<B, 32, 0, ctx.stream>>>(xvec, xn, 192); }
  gemm(ctx, gemm_args_linear(xn, 192, B, m.spk_affine, spk, 80), m.spk_affine);
  // token embedding + encoder
  float* e0 = ctx.ws.get<float>((size_t)rows1 * 512);
  float* x = ctx.ws.get<float>((size_t)rows1 * 512);
  float* c1 = ctx.ws.get<float>((size_t)rows1 * 512);
  gather_rows(ctx, m.input_embedding.p, 512, tokens, e0, 512, rows1, 512, nullptr, 0, nullptr, 6561);
  enc_embed(ctx, m.embed, e0, x, rows1);
  // PreLookaheadLayer (upsample_encoder.py:84-96): conv k4 looking 3 frames ahead, leaky_relu(0.01), causal conv k3, + x
  GemmDev gp1 = conv_args(x, 512, m.pre_conv1, 512, 4, 1, 0, 1, L1, L1, c1, 512);
  gp1.act = ACT_LRELU; gp1.act_p = 0.01f;
  gemm(ctx, gp1, m.pre_conv1);
  GemmDev gp2 = conv_args(c1, 512, m.pre_conv2, 512, 3, 1, 2, 1, L1, L1, e0, 512);
  gp2.res = x; gp2.ldr = 512;
  gemm(ctx, gp2, m.pre_conv2);
  float* xs = e0;   // encoder stream now lives in e0
  // rel-pos table rows [center-T+1, center+T) of the host-built espnet table (embedding.py:283-294)
  CBX_REQUIRE(L2.max_len <= m.pe_center + 1, "sequence longer than the positional table");
  for (int i = 0; i < 6; ++i)
    encoder_layer(h, ctx, m.enc[i], xs, L1, m.pe_table.p + (size_t)(m.pe_center - L1.max_len + 1) * 512, L1.max_len);
  // Upsample1D (upsample_encoder.py:59-63): nearest x2, left pad 4, conv k5
  float* u = ctx.ws.get<float>((size_t)rows2 * 512);
  float* x2 = ctx.ws.get<float>((size_t)rows2 * 512);
  upsample2(ctx, xs, u, 512, L2.tile_seq, L2.start, L2.len, L1.start, rows2);
  gemm(ctx, conv_args(u, 512, m.up_conv, 512, 5, 1, 4, 1, L2, L2, x2, 512), m.up_conv);
  enc_embed(ctx, m.up_embed, x2, u, rows2);
  for (int i = 0; i < 4; ++i)
    encoder_layer(h, ctx, m.up_enc[i], u, L2, m.pe_table.p + (size_t)(m.pe_center - L2.max_len + 1) * 512, L2.max_len);
  layernorm(ctx, u, 512, m.after_w.p, m.after_b.p, x2, 512, rows2, 512, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  // encoder_proj (flow.py:173), zero on layout padding rows
  GemmDev gm = conv_args(x2, 512, m.enc_proj, 512, 1, 0, 0, 1, L2, L2, mu, 80);
  gemm(ctx, gm, m.enc_proj);
}

// ---- CFM estimator ---------------------------------------------------------------------------------
struct EstBufs { float *h1, *h2, *hn, *qkv, *att, *ff; __nv_bfloat16 *qkv_hi, *qkv_lo; CUtensorMap tm_hi, tm_lo; bool tc; bool f16; bool a16;
                 double attn_work;
                 // plane-fed convs (b.cp): conv input [rows][<=512] and the LayerNorm+Mish output [rows][256] as bf16 hi/lo planes
                 bool cp; __nv_bfloat16 *in_hi, *in_lo, *h_hi, *h_lo; };

// stride-1 conv as a plane-fed GEMM: the input travels as bf16 hi/lo planes (zero on layout padding rows), every tap is the same
// TMA box shifted by one row -- no fp32 tile, no converter warps in the main loop; same split, same products as the converter path
static GemmDev conv_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int ldp, const Weight& W, int c_in, int ntaps, int pad,
                           const cbx_layout& L, float* C, int ldc) {
  GemmDev g = conv_args(nullptr, ldp, W, c_in, ntaps, 1, pad, 1, L, L, C, ldc);
  g.Ahi = hi; g.Alo = lo; g.ldab = ldp;
  return g;
}

static void cfm_resnet(Ctx& ctx, CfmResnet& r, const float* in, int lda, int cin, float* out, int ldo, const float* tvec,
                       const cbx_layout& L, EstBufs& b) {
  SeqMap sm = seqmap(L, L);
  if (b.cp) {
    pack_planes_seq(ctx, in, lda, L.rows, cin, b.in_hi, b.in_lo, 512, sm);
    gemm(ctx, conv_planes(b.in_hi, b.in_lo, 512, r.conv1, cin, 3, 2, L, b.h1, 256), r.conv1);
    layernorm(ctx, b.h1, 256, r.ln1_w.p, r.ln1_b.p, nullptr, 256, L.rows, 256, 1e-5f, ACT_MISH, 1.f, tvec, 0, &sm, b.h_hi, b.h_lo);
    gemm(ctx, conv_planes(b.h_hi, b.h_lo, 256, r.conv2, 256, 3, 2, L, b.h1, 256), r.conv2);
    layernorm(ctx, b.h1, 256, r.ln2_w.p, r.ln2_b.p, b.h2, 256, L.rows, 256, 1e-5f, ACT_MISH, 1.f, nullptr, 0, &sm);
    GemmDev g = conv_planes(b.in_hi, b.in_lo, 512, r.res, cin, 1, 0, L, out, ldo);
    g.res = b.h2; g.ldr = 256;
    gemm(ctx, g, r.res);
    return;
  }
  gemm(ctx, conv_args(in, lda, r.conv1, cin, 3, 1, 2, 1, L, L, b.h1, 256), r.conv1);
  layernorm(ctx, b.h1, 256, r.ln1_w.p, r.ln1_b.p, b.h2, 256, L.rows, 256, 1e-5f, ACT_MISH, 1.f, tvec, 0, &sm);
  gemm(ctx, conv_args(b.h2, 256, r.conv2, 256, 3, 1, 2, 1, L, L, b.h1, 256), r.conv2);
  layernorm(ctx, b.h1, 256, r.ln2_w.p, r.ln2_b.p, b.h2, 256, L.rows, 256, 1e-5f, ACT_MISH, 1.f, nullptr, 0, &sm);
  GemmDev g = conv_args(in, lda, r.res, cin, 1, 0, 0, 1, L, L, out, ldo);
  g.res = b.h2; g.ldr = 256;
  gemm(ctx, g, r.res);
}
static void cfm_tfmr(Ctx& ctx, CfmTfmr& t, float* x, int ldx, const cbx_layout& L, EstBufs& b) {
  const int rows = L.rows;
  if (b.tc) {
    // Every activation between two GEMMs of the block travels as bf16 hi/lo planes (the producer splits once, in its
    // epilogue) so that the GEMM loads its A operand by TMA straight into the UMMA layout: LN -> qkv -> tcgen05
    // attention -> out(+x) ; LN -> ff1(GELU) -> ff2(+x).  Same split as the in-kernel converter: results are unchanged.
    const size_t R = (size_t)rows;
    __nv_bfloat16* hn_hi = reinterpret_cast<__nv_bfloat16*>(b.hn);  __nv_bfloat16* hn_lo = hn_hi + R * 256;
    __nv_bfloat16* at_hi = reinterpret_cast<__nv_bfloat16*>(b.att); __nv_bfloat16* at_lo = at_hi + R * 512;
    __nv_bfloat16* ff_hi = reinterpret_cast<__nv_bfloat16*>(b.ff);  __nv_bfloat16* ff_lo = ff_hi + R * 1024;
    // b.a16 (opt-in, tools/attn_precision_study.py: all block GEMM inputs as ONE fp16 value keep the mel RMS at 1.4e-4):
    // the same dataflow with one fp16 plane per activation -- half the activation bytes, one MMA term per K step.
    __half* hn16 = reinterpret_cast<__half*>(b.hn);
    __half* at16 = reinterpret_cast<__half*>(b.att);
    __half* ff16 = reinterpret_cast<__half*>(b.ff);
    auto feed = [&](GemmDev& g, __nv_bfloat16* hi, __nv_bfloat16* lo, __half* h16, int ld) {
      if (b.a16) { g.A16 = h16; g.lda16 = ld; } else { g.Ahi = hi; g.Alo = lo; g.ldab = ld; }
    };
    if (b.a16) layernorm(ctx, x, ldx, t.ln1_w.p, t.ln1_b.p, nullptr, 256, rows, 256, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr, nullptr, nullptr, hn16);
    else layernorm(ctx, x, ldx, t.ln1_w.p, t.ln1_b.p, nullptr, 256, rows, 256, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr, hn_hi, hn_lo);
    GemmDev gq = gemm_args_linear(nullptr, 256, rows, t.qkv, nullptr, 0);
    feed(gq, hn_hi, hn_lo, hn16, 256);
    gq.Chi = b.qkv_hi; gq.Clo = b.qkv_lo; gq.ldcb = 1536;
    gq.c_half = b.f16 ? 1 : 0;        // one fp16 plane at qkv_hi (tm_hi maps the same bytes: 2-byte elements, 1536 per row)
    gemm(ctx, gq, t.qkv);
    AttnTcArgs a;
    a.tm_hi = &b.tm_hi; a.tm_lo = &b.tm_lo; a.q_col = 0; a.k_col = 512; a.v_col = 1024; a.O = nullptr; a.ldo = 512;
    if (b.a16) a.O16 = at16; else { a.Ohi = at_hi; a.Olo = at_lo; }
    a.f16 = b.f16 ? 1 : 0;
    a.n_seq = L.n_seq; a.n_heads = 8; a.q_start = L.start; a.q_len = L.len; a.kv_start = L.start; a.kv_len = L.len;
    a.max_q_len = L.max_len; a.scale = 0.125f; a.work = b.attn_work;
    attention_tc(ctx, a);
    GemmDev go = gemm_args_linear(nullptr, 512, rows, t.out, x, ldx);
    feed(go, at_hi, at_lo, at16, 512);
    go.res = x; go.ldr = ldx;
    gemm(ctx, go, t.out);
    if (b.a16) layernorm(ctx, x, ldx, t.ln3_w.p, t.ln3_b.p, nullptr, 256, rows, 256, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr, nullptr, nullptr, hn16);
    else layernorm(ctx, x, ldx, t.ln3_w.p, t.ln3_b.p, nullptr, 256, rows, 256, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr, hn_hi, hn_lo);
    GemmDev g1 = gemm_args_linear(nullptr, 256, rows, t.ff1, nullptr, 0);
    feed(g1, hn_hi, hn_lo, hn16, 256);
    g1.Chi = ff_hi; g1.Clo = ff_lo; g1.ldcb = 1024;
    g1.c_half = b.a16 ? 1 : 0;        // ff16 aliases ff_hi
    g1.act = ACT_GELU;
    gemm(ctx, g1, t.ff1);
    GemmDev g2 = gemm_args_linear(nullptr, 1024, rows, t.ff2, x, ldx);
    feed(g2, ff_hi, ff_lo, ff16, 1024);
    g2.res = x; g2.ldr = ldx;
    gemm(ctx, g2, t.ff2);
    return;
  }
  layernorm(ctx, x, ldx, t.ln1_w.p, t.ln1_b.p, b.hn, 256, rows, 256, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  {
    gemm(ctx, gemm_args_linear(b.hn, 256, rows, t.qkv, b.qkv, 1536), t.qkv);
    AttnArgs a;
    a.Q = b.qkv; a.K = b.qkv + 512; a.V = b.qkv + 1024; a.ldq = a.ldk = a.ldv = 1536; a.O = b.att; a.ldo = 512;
    a.n_seq = L.n_seq; a.n_heads = 8; a.q_start = L.start; a.q_len = L.len; a.kv_start = L.start; a.kv_len = L.len;
    a.max_q_len = L.max_len; a.scale = 0.125f; a.causal = 0;
    attention(ctx, a);
  }
  GemmDev go = gemm_args_linear(b.att, 512, rows, t.out, x, ldx);
  go.res = x; go.ldr = ldx;
  gemm(ctx, go, t.out);
  layernorm(ctx, x, ldx, t.ln3_w.p, t.ln3_b.p, b.hn, 256, rows, 256, 1e-5f, ACT_NONE, 1.f, nullptr, 0, nullptr);
  GemmDev g1 = gemm_args_linear(b.hn, 256, rows, t.ff1, b.ff, 1024);
  g1.act = ACT_GELU;
  gemm(ctx, g1, t.ff1);
  GemmDev g2 = gemm_args_linear(b.ff, 1024, rows, t.ff2, x, ldx);
  g2.res = x; g2.ldr = ldx;
  gemm(ctx, g2, t.ff2);
}

void cfm_solve(cbx_handle* h, Ctx& ctx, const float* mu, const float* spk, const float* cond, float* x,
               const cbx_layout& L2, const cbx_layout& L3, int n_steps, float cfg_rate, int meanflow) {
  FlowModel& m = h->flow;
  CBX_REQUIRE(m.ready, "flow weights not finalized");
  CBX_REQUIRE(n_steps >= 1 && n_steps <= 15, "n_steps");
  CBX_REQUIRE((meanflow != 0) == m.meanflow, "meanflow flag does not match the loaded estimator");
  const int B = L2.n_seq;
  const int cfg = (L3.n_seq == 2 * B) ? 1 : 0;
  CBX_REQUIRE(cfg || L3.n_seq == B, "L3 must hold B or 2B sequences");
  const long rows3 = L3.rows, rows2 = L2.rows;
  // time grid (flow_matching.py:222-225): linspace(0,1,n+1) [torch's symmetric formula], cosine schedule unless meanflow
  float tspan[16];
  {
    const float step = 1.0f / (float)n_steps;
    const int steps = n_steps + 1, halfway = steps / 2;
    for (int k = 0; k < steps; ++k) {
      float t = (k < halfway) ? 0.0f + step * (float)k : 1.0f - step * (float)(steps - k - 1);
      if (!meanflow) t = 1.0f - cosf(t * 0.5f * 3.14159265358979323846f);
      tspan[k] = t;
    }
  }
  // time embeddings for all steps: sinusoid(320) -> Linear -> SiLU -> Linear (matcha/decoder.py:20-29,103-117)
  const int nt = meanflow ? n_steps + 1 : n_steps;     // meanflow also embeds r = t_{k+1}
  float* tdev = ctx.ws.get<float>(32);
  float* sinus = ctx.ws.get<float>((size_t)nt * 320);
  float* t1 = ctx.ws.get<float>((size_t)nt * 1024);
  float* temb = ctx.ws.get<float>((size_t)nt * 1024);
  if (!ctx.dry) {
    TVals tv; for (int k = 0; k < 32; ++k) tv.v[k] = k <= n_steps ? tspan[k] : 0.f;
    ctx.launches++;
    tvals_kernel<<<1, 32, 0, ctx.stream>>>(tv, tdev, nt);
  }
  time_sinusoid(ctx, tdev, sinus, nt, 320, 1000.0f);
  GemmDev gt1 = gemm_args_linear(sinus, 320, nt, m.time1, t1, 1024);
  gt1.act = ACT_SILU;
  gemm(ctx, gt1, m.time1);
  gemm(ctx, gemm_args_linear(t1, 1024, nt, m.time2, temb, 1024), m.time2);
  float* tstep = temb;   // [n_steps][1024] embedding used at step k
  if (meanflow) {        // decoder.py:264-268: mixer(cat(emb(t), emb(r)))
    float* cat = ctx.ws.get<float>((size_t)n_steps * 2048);
    copy2d(ctx, temb, 1024, cat, 2048, n_steps, 1024);
    copy2d(ctx, temb + 1024, 1024, cat + 1024, 2048, n_steps, 1024);
    tstep = ctx.ws.get<float>((size_t)n_steps * 1024);
    gemm(ctx, gemm_args_linear(cat, 2048, n_steps, m.time_mixer, tstep, 1024), m.time_mixer);
  }
  // per-resnet additive vectors: Linear(Mish(temb))  (matcha/decoder.py:48,58)
  float* tm = ctx.ws.get<float>((size_t)n_steps * 1024);
  ew_act(ctx, tstep, 1024, tm, 1024, n_steps, 1024, ACT_MISH, 0.f, nullptr);
  float* tvec = ctx.ws.get<float>((size_t)14 * n_steps * 256);        // [resnet][step][256]
  CfmResnet* resnets[14];
  resnets[0] = &m.down.res; for (int i = 0; i < 12; ++i) resnets[1 + i] = &m.mid[i].res; resnets[13] = &m.up.res;
  for (int i = 0; i < 14; ++i)
    gemm(ctx, gemm_args_linear(tm, 1024, n_steps, resnets[i]->mlp, tvec + (size_t)i * n_steps * 256, 256), resnets[i]->mlp);

  // buffers of the estimator
  float* xin = ctx.ws.get<float>((size_t)rows3 * 320);
  float* xcat = ctx.ws.get<float>((size_t)rows3 * 512);
  float* bufX = ctx.ws.get<float>((size_t)rows3 * 256);
  float* bufY = ctx.ws.get<float>((size_t)rows3 * 256);
  float* v = ctx.ws.get<float>((size_t)rows3 * 80);
  EstBufs b;
  b.h1 = ctx.ws.get<float>((size_t)rows3 * 256); b.h2 = ctx.ws.get<float>((size_t)rows3 * 256);
  b.hn = ctx.ws.get<float>((size_t)rows3 * 256); b.qkv = ctx.ws.get<float>((size_t)rows3 * 1536);
  b.att = ctx.ws.get<float>((size_t)rows3 * 512); b.ff = ctx.ws.get<float>((size_t)rows3 * 1024);
  b.attn_work = 0.0;
  if (L3.h_len) for (int i = 0; i < L3.n_seq; ++i) b.attn_work += 4.0 * 64.0 * 8.0 * (double)L3.h_len[i] * (double)L3.h_len[i];
  b.tc = (ctx.attn_impl == 0 && ctx.gemm_impl == 0);
  static const bool conv_planes_on = !(getenv("CBX_CONV_PLANES") && atoi(getenv("CBX_CONV_PLANES")) == 0);
  b.cp = b.tc && conv_planes_on && L3.h_start && L3.h_len;
  // the plane path has no per-row tap mask: it needs >= 2 (zero) padding rows between consecutive sequences of the layout
  if (b.cp) for (int i = 0; i + 1 < L3.n_seq; ++i) if (L3.h_start[i + 1] - (L3.h_start[i] + L3.h_len[i]) < 2) { b.cp = false; break; }
  b.in_hi = b.in_lo = b.h_hi = b.h_lo = nullptr;
  if (b.cp) {
    b.in_hi = ctx.ws.get<__nv_bfloat16>((size_t)rows3 * 512); b.in_lo = ctx.ws.get<__nv_bfloat16>((size_t)rows3 * 512);
    b.h_hi = ctx.ws.get<__nv_bfloat16>((size_t)rows3 * 256); b.h_lo = ctx.ws.get<__nv_bfloat16>((size_t)rows3 * 256);
  }
  b.f16 = b.tc && ctx.attn_f16 != 0;
  b.a16 = b.f16 && ctx.cfm_act_f16 != 0;      // fp16 activations ride on the fp16 attention variant
  b.qkv_hi = reinterpret_cast<__nv_bfloat16*>(b.qkv);                 // the planes reuse the fp32 qkv buffer
  b.qkv_lo = b.qkv_hi + (size_t)rows3 * 1536;
  if (b.tc && !ctx.dry) { make_plane_tmap(&b.tm_hi, b.qkv_hi, rows3, 1536); make_plane_tmap(&b.tm_lo, b.qkv_lo, rows3, 1536); }
  SeqMap sm3 = seqmap(L3, L3);

  for (int k = 0; k < n_steps; ++k) {
    cfm_assemble(ctx, xin, x, mu, spk, cond, L3.tile_seq, L3.start, L3.len, L2.start, B, rows3, k == 0 ? 1 : 0);
    auto tv = [&](int i) { return tvec + ((size_t)i * n_steps + k) * 256; };
    // down block (decoder.py:280-295); its output doubles as the skip connection: keep it in xcat[:, 256:512]
    float* skip = xcat + 256;
    cfm_resnet(ctx, m.down.res, xin, 320, 320, skip, 512, tv(0), L3, b);
    for (int j = 0; j < 4; ++j) cfm_tfmr(ctx, m.down.t[j], skip, 512, L3, b);
    auto causal_conv = [&](const float* src, int lds, Weight& W, float* dst) {      // CausalConv1d k3 (decoder.py:26-34)
      if (b.cp) {
        pack_planes_seq(ctx, src, lds, rows3, 256, b.in_hi, b.in_lo, 512, sm3);
        gemm(ctx, conv_planes(b.in_hi, b.in_lo, 512, W, 256, 3, 2, L3, dst, 256), W);
      } else {
        gemm(ctx, conv_args(src, lds, W, 256, 3, 1, 2, 1, L3, L3, dst, 256), W);
      }
    };
    causal_conv(skip, 512, m.down_conv, bufX);
    // 12 mid blocks (decoder.py:299-312); the last one writes into xcat[:, 0:256]
    float* cur = bufX; float* nxt = bufY;
    for (int i = 0; i < 12; ++i) {
      float* out = (i == 11) ? xcat : nxt;
      const int ldo = (i == 11) ? 512 : 256;
      cfm_resnet(ctx, m.mid[i].res, cur, 256, 256, out, ldo, tv(1 + i), L3, b);
      for (int j = 0; j < 4; ++j) cfm_tfmr(ctx, m.mid[i].t[j], out, ldo, L3, b);
      if (i != 11) { float* t = cur; cur = nxt; nxt = t; }
    }
    // up block on cat[x, skip] (decoder.py:314-330)
    cfm_resnet(ctx, m.up.res, xcat, 512, 512, bufX, 256, tv(13), L3, b);
    for (int j = 0; j < 4; ++j) cfm_tfmr(ctx, m.up.t[j], bufX, 256, L3, b);
    causal_conv(bufX, 256, m.up_conv2, bufY);
    // final block + projection (decoder.py:331-333)
    causal_conv(bufY, 256, m.final_conv, b.h1);
    if (b.cp) {
      layernorm(ctx, b.h1, 256, m.final_ln_w.p, m.final_ln_b.p, nullptr, 256, (int)rows3, 256, 1e-5f, ACT_MISH, 1.f, nullptr, 0, &sm3, b.h_hi, b.h_lo);
      gemm(ctx, conv_planes(b.h_hi, b.h_lo, 256, m.final_proj, 256, 1, 0, L3, v, 80), m.final_proj);
    } else {
      layernorm(ctx, b.h1, 256, m.final_ln_w.p, m.final_ln_b.p, b.h2, 256, (int)rows3, 256, 1e-5f, ACT_MISH, 1.f, nullptr, 0, &sm3);
      gemm(ctx, conv_args(b.h2, 256, m.final_proj, 256, 1, 0, 0, 1, L3, L3, v, 80), m.final_proj);
    }
    cfm_euler(ctx, x, v, L2.tile_seq, L2.start, L2.len, L3.start, B, tspan[k + 1] - tspan[k], cfg_rate, cfg, rows2);
  }
}

}  // namespace cbx
