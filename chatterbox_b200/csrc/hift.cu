// HiFT vocoder (reference src/chatterbox/models/s3gen/hifigan.py + f0_predictor.py):
//   F0 predictor -> harmonic+noise source -> STFT(source) -> conv_pre -> 3 x [lrelu, polyphase ConvTranspose,
//   + source_down/source_resblock, mean of 3 Snake ResBlocks] -> conv_post -> exp/sin -> iSTFT -> clamp.
// Every conv is an implicit GEMM on packed channel-last buffers; weight-norm is folded at load time; Snake
// activations are fused into the epilogue of the conv that produces their input.
#include "engine.h"
#include <cmath>
#include <cstdlib>

namespace cbx {

// w = v * (g / ||v||), norm over all dims except 0 (torch weight_norm dim=0; for ConvTranspose1d dim 0 = in-channels)
static std::vector<float> fold_wn(cbx_handle* h, const std::string& name) {
  if (has_tensor(h, name + ".weight")) return host_tensor(h, name + ".weight").data;
  const HostTensor& g = host_tensor(h, name + ".parametrizations.weight.original0");
  const HostTensor& v = host_tensor(h, name + ".parametrizations.weight.original1");
  const size_t n0 = (size_t)v.shape[0], inner = v.data.size() / n0;
  std::vector<float> w(v.data.size());
  for (size_t i = 0; i < n0; ++i) {
    double ss = 0.0;
    for (size_t j = 0; j < inner; ++j) ss += (double)v.data[i * inner + j] * (double)v.data[i * inner + j];
    const float scale = g.data[i] / (float)std::sqrt(ss);
    for (size_t j = 0; j < inner; ++j) w[i * inner + j] = v.data[i * inner + j] * scale;
  }
  return w;
}
static std::vector<int64_t> wn_shape(cbx_handle* h, const std::string& name) {
  if (has_tensor(h, name + ".weight")) return host_tensor(h, name + ".weight").shape;
  return host_tensor(h, name + ".parametrizations.weight.original1").shape;
}
static void pack_wn_taps(cbx_handle* h, Weight& W, const std::string& name) {
  auto w = fold_wn(h, name); auto sh = wn_shape(h, name);
  pack_conv_taps(W, w.data(), host_tensor(h, name + ".bias").data.data(), (int)sh[0], (int)sh[1], (int)sh[2]);
}
static void pack_wn_window(cbx_handle* h, Weight& W, const std::string& name) {
  auto w = fold_wn(h, name); auto sh = wn_shape(h, name);
  pack_conv_window(W, w.data(), host_tensor(h, name + ".bias").data.data(), (int)sh[0], (int)sh[1], (int)sh[2]);
}
// ConvTranspose1d(Cin, Cout, k, stride u, padding p) as a 3-tap conv with N = u*Cout (polyphase):
// out[q*u + phi] = sum_d x[q + d] . Wt[:, :, phi + p - d*u],  d in {-1, 0, +1}
static void pack_upconv(cbx_handle* h, Weight& W, const std::string& name, int u) {
  auto w = fold_wn(h, name); auto sh = wn_shape(h, name);
  const int cin = (int)sh[0], cout = (int)sh[1], k = (int)sh[2], p = (k - u) / 2;
  const auto& b = host_tensor(h, name + ".bias").data;
  std::vector<float> wp((size_t)u * cout * cin * 3, 0.0f);   // [N = u*cout][cin][3]
  std::vector<float> bp((size_t)u * cout);
  for (int phi = 0; phi < u; ++phi)
    for (int co = 0; co < cout; ++co) {
      bp[(size_t)phi * cout + co] = b[co];
      for (int dd = 0; dd < 3; ++dd) {
        const int d = dd - 1, j = phi + p - d * u;
        if (j < 0 || j >= k) continue;
        for (int ci = 0; ci < cin; ++ci)
          wp[(((size_t)phi * cout + co) * cin + ci) * 3 + dd] = w[((size_t)ci * cout + co) * k + j];
      }
    }
  pack_conv_taps(W, wp.data(), bp.data(), u * cout, cin, 3);
}
static void build_resblock(cbx_handle* h, HiftResBlock& rb, const std::string& p) {
  for (int j = 0; j < 3; ++j) {
    pack_wn_taps(h, rb.c1[j], p + "convs1." + std::to_string(j));
    pack_wn_taps(h, rb.c2[j], p + "convs2." + std::to_string(j));
    rb.a1[j] = upload_tensor(h, p + "activations1." + std::to_string(j) + ".alpha");
    rb.a2[j] = upload_tensor(h, p + "activations2." + std::to_string(j) + ".alpha");
  }
  rb.k = (int)wn_shape(h, p + "convs1.0")[2];
}

void hift_finalize(cbx_handle* h) {
  HiftModel& m = h->hift;
  const std::string f = "hift.";
  pack_wn_window(h, m.f0conv[0], f + "f0_predictor.condnet.0");
  for (int i = 1; i < 5; ++i) pack_wn_taps(h, m.f0conv[i], f + "f0_predictor.condnet." + std::to_string(2 * i));
  m.f0_w = upload_tensor(h, f + "f0_predictor.classifier.weight");
  m.f0_b = host_tensor(h, f + "f0_predictor.classifier.bias").data[0];
  m.src_w = upload_tensor(h, f + "m_source.l_linear.weight");
  m.src_b = host_tensor(h, f + "m_source.l_linear.bias").data[0];
  pack_wn_window(h, m.conv_pre, f + "conv_pre");
  const int us[3] = {8, 5, 3};
  for (int i = 0; i < 3; ++i) {
    pack_upconv(h, m.ups[i], f + "ups." + std::to_string(i), us[i]);
    const HostTensor& w = host_tensor(h, f + "source_downs." + std::to_string(i) + ".weight");
    pack_conv_window(m.src_down[i], w.data.data(), host_tensor(h, f + "source_downs." + std::to_string(i) + ".bias").data.data(),
                     (int)w.shape[0], (int)w.shape[1], (int)w.shape[2]);
    build_resblock(h, m.src_rb[i], f + "source_resblocks." + std::to_string(i) + ".");
  }
  for (int i = 0; i < 9; ++i) build_resblock(h, m.rb[i], f + "resblocks." + std::to_string(i) + ".");
  pack_wn_taps(h, m.conv_post, f + "conv_post");
  m.ready = true;
}

static GemmDev window_args(const float* A, const Weight& W, int c_in, int ntaps, int pad, int stride,
                           const cbx_layout& out, const cbx_layout& in, float* C, int ldc) {
  GemmDev g = conv_args(A, c_in, W, c_in, ntaps, 0, pad, stride, out, in, C, ldc);
  g.a_mode = A_WINDOW; g.ctap = 64; g.k_total = ntaps * c_in;
  return g;
}

void hift_source_run(cbx_handle* h, Ctx& ctx, const float* mel, const cbx_hift_geom& g, const float* phase_vec,
                     const float* noise, unsigned long long seed, float* s_out, const float* f0_in, float* f0_out) {
  HiftModel& m = h->hift;
  CBX_REQUIRE(m.ready, "hift weights not finalized");
  const cbx_layout& LT = g.LT;
  const int rows = LT.rows;
  float* a = ctx.ws.get<float>((size_t)rows * 512);
  float* b = ctx.ws.get<float>((size_t)rows * 512);
  float* f0 = f0_out ? f0_out : ctx.ws.get<float>(rows);
  // ConvRNNF0Predictor (f0_predictor.py:27-55): 5 x [conv k3 pad 1, ELU] -> |Linear(512->1)|
  GemmDev g0 = window_args(mel, m.f0conv[0], 80, 3, 1, 1, LT, LT, a, 512);
  g0.act = ACT_ELU;
  gemm(ctx, g0, m.f0conv[0]);
  float* cur = a; float* nxt = b;
  for (int i = 1; i < 5; ++i) {
    GemmDev gi = conv_args(cur, 512, m.f0conv[i], 512, 3, 1, 1, 1, LT, LT, nxt, 512);
    gi.act = ACT_ELU;
    gemm(ctx, gi, m.f0conv[i]);
    float* t = cur; cur = nxt; nxt = t;
  }
  f0_head(ctx, cur, 512, m.f0_w.p, m.f0_b, f0, rows);
  if (f0_in) f0 = const_cast<float*>(f0_in);
  // per-frame phase table: 9 x (double c_0 + flag byte) per frame row
  float* cumf = ctx.ws.get<float>((size_t)rows * 9 * 3 + 64);
  hift_source(ctx, f0, cumf, phase_vec, noise, m.src_w.p, m.src_b, s_out, LT.start, LT.len,
              reinterpret_cast<const long*>(g.sample_start), LT.n_seq, LT.max_len, seed, rows);
}

struct RbBufs { float *xt, *t1, *xr; __nv_bfloat16 *p_hi, *p_lo, *t_hi, *t_lo; bool planes; };

// ResBlock.forward (hifigan.py:154-161) on the shared-memory-staged conv kernel (hift_conv.cu): every conv input travels
// as bf16 hi/lo planes written by its producer's epilogue; Snake, bias, residual, 1/3 scaling and the stage sum are all
// epilogue work.  in: x (read only).  out: dst (+)= resblock(x) * out_scale
static void run_resblock_planes(Ctx& ctx, HiftResBlock& rb, const float* x, int C, const cbx_layout& L, float* dst,
                                float out_scale, int accumulate, RbBufs& b) {
  const int dil[3] = {1, 3, 5};
  const int k = rb.k;
  snake_planes(ctx, x, C, rb.a1[0].p, b.p_hi, b.p_lo, L);                                   // snake(x; a1_0)
  for (int j = 0; j < 3; ++j) {
    // conv1_j (dilated) + Snake(a2_j) -> planes t
    hift_conv(ctx, rb.c1[j], C, k, dil[j], L, b.p_hi, b.p_lo, 0, rb.a2[j].p, nullptr, nullptr, 0, 1.f, b.t_hi, b.t_lo);
    const float* res = (j == 0) ? x : b.xr;
    if (j < 2) {    // conv2_j + residual -> xr (fp32) and the next branch's snake(xr; a1_{j+1}) as planes
      hift_conv(ctx, rb.c2[j], C, k, 1, L, b.t_hi, b.t_lo, 1, rb.a1[j + 1].p, res, b.xr, 0, 1.f, b.p_hi, b.p_lo);
    } else {        // last branch: into the stage sum
      hift_conv(ctx, rb.c2[j], C, k, 1, L, b.t_hi, b.t_lo, 2, nullptr, res, dst, accumulate, out_scale, nullptr, nullptr);
    }
  }
}

// ResBlock.forward (hifigan.py:154-161).  in: x (read only).  out: dst (+)= (resblock(x)) * out_scale
static void run_resblock(Ctx& ctx, HiftResBlock& rb, const float* x, int C, const cbx_layout& L, float* dst,
                         float out_scale, int accumulate, RbBufs& b) {
  if (b.planes) { run_resblock_planes(ctx, rb, x, C, L, dst, out_scale, accumulate, b); return; }
  const int dil[3] = {1, 3, 5};
  const int k = rb.k;
  ew_act(ctx, x, C, b.xt, C, L.rows, C, ACT_SNAKE, 0.f, rb.a1[0].p);
  for (int j = 0; j < 3; ++j) {
    GemmDev g1 = conv_args(b.xt, C, rb.c1[j], C, k, dil[j], dil[j] * (k - 1) / 2, 1, L, L, b.t1, C);
    g1.act = ACT_SNAKE; g1.act_vec = rb.a2[j].p;
    gemm(ctx, g1, rb.c1[j]);
    const float* res = (j == 0) ? x : b.xr;
    if (j < 2) {
      GemmDev g2 = conv_args(b.t1, C, rb.c2[j], C, k, 1, (k - 1) / 2, 1, L, L, b.xr, C);
      g2.res = res; g2.ldr = C;
      g2.C2 = b.xt; g2.ldc2 = C; g2.act2 = ACT_SNAKE; g2.act2_vec = rb.a1[j + 1].p;
      gemm(ctx, g2, rb.c2[j]);
    } else {
      GemmDev g2 = conv_args(b.t1, C, rb.c2[j], C, k, 1, (k - 1) / 2, 1, L, L, dst, C);
      g2.res = res; g2.ldr = C; g2.out_scale = out_scale; g2.accumulate = accumulate;
      gemm(ctx, g2, rb.c2[j]);
    }
  }
}

void hift_decode_run(cbx_handle* h, Ctx& ctx, const float* mel, const float* s, const cbx_hift_geom& g, float* wav,
                     int trim_fade) {
  HiftModel& m = h->hift;
  CBX_REQUIRE(m.ready, "hift weights not finalized");
  const cbx_layout* Ls[4] = {&g.LT, &g.L8, &g.L40, &g.L120};
  const int chans[4] = {512, 256, 128, 64};
  const int us[3] = {8, 5, 3};
  const int sd_k[3] = {30, 6, 1}, sd_s[3] = {15, 3, 1}, sd_p[3] = {7, 1, 0};
  for (int i = 0; i < 3; ++i) CBX_REQUIRE(Ls[i + 1]->rows == Ls[i]->rows * us[i], "upsampled layouts must be exact multiples");
  const long* sstart = reinterpret_cast<const long*>(g.sample_start);
  // STFT of the source (hifigan.py:413-414)
  float* s_stft = ctx.ws.get<float>((size_t)g.L120.rows * 18);
  hift_stft(ctx, s, s_stft, g.L120.start, g.LT.len, sstart, g.LT.n_seq, g.LT.max_len);
  // conv_pre (hifigan.py:416)
  float* x = ctx.ws.get<float>((size_t)g.LT.rows * 512);
  gemm(ctx, window_args(mel, m.conv_pre, 80, 7, 3, 1, g.LT, g.LT, x, 512), m.conv_pre);
  for (int i = 0; i < 3; ++i) {
    const cbx_layout& Lin = *Ls[i];
    const cbx_layout& Lout = *Ls[i + 1];
    const int Cin = chans[i], C = chans[i + 1];
    const size_t nout = (size_t)(Lout.rows + 1) * C;
    float* xl = ctx.ws.get<float>((size_t)Lin.rows * Cin);
    float* xu = ctx.ws.get<float>(nout);
    float* si = ctx.ws.get<float>(nout);
    float* xs = ctx.ws.get<float>(nout);
    RbBufs b; b.xt = ctx.ws.get<float>(nout); b.t1 = ctx.ws.get<float>(nout); b.xr = ctx.ws.get<float>(nout);
    // the planes of the staged-conv path live in the same two buffers (hi + lo = 4 bytes per element)
    static const bool legacy = getenv("CBX_HIFT") && std::string(getenv("CBX_HIFT")) == "legacy";
    b.planes = !legacy && ctx.gemm_impl == 0;
    b.p_hi = reinterpret_cast<__nv_bfloat16*>(b.xt); b.p_lo = b.p_hi + (size_t)Lout.rows * C;
    b.t_hi = reinterpret_cast<__nv_bfloat16*>(b.t1); b.t_lo = b.t_hi + (size_t)Lout.rows * C;
    ew_act(ctx, x, Cin, xl, Cin, Lin.rows, Cin, ACT_LRELU, 0.1f, nullptr);                          // :418
    // ups[i]: polyphase ConvTranspose -> rows q*u + phi (shifted by one row at the last stage for the reflection pad)
    const int shift = (i == 2) ? 1 : 0;
    GemmDev gu = conv_args(xl, Cin, m.ups[i], Cin, 3, 1, 1, 1, Lin, Lin, xu + (size_t)shift * C, us[i] * C);
    gemm(ctx, gu, m.ups[i]);
    if (shift) reflect_row0(ctx, xu, C, Lout.start, Lout.n_seq);                                     // :421-422
    // source branch (hifigan.py:425-427): x = x + source_resblock(source_down(s_stft))
    gemm(ctx, window_args(s_stft, m.src_down[i], 18, sd_k[i], sd_p[i], sd_s[i], Lout, g.L120, si, C), m.src_down[i]);
    run_resblock(ctx, m.src_rb[i], si, C, Lout, xu, 1.0f, 1, b);
    // mean of the three ResBlocks (hifigan.py:429-435)
    for (int j = 0; j < 3; ++j) run_resblock(ctx, m.rb[i * 3 + j], xu, C, Lout, xs, 1.0f / 3.0f, j > 0, b);
    x = xs;
  }
  // conv_post + iSTFT (hifigan.py:437-443)
  float* xl = ctx.ws.get<float>((size_t)g.L120.rows * 64);
  float* y = ctx.ws.get<float>((size_t)g.L120.rows * 18);
  ew_act(ctx, x, 64, xl, 64, g.L120.rows, 64, ACT_LRELU, 0.01f, nullptr);
  gemm(ctx, conv_args(xl, 64, m.conv_post, 64, 7, 1, 3, 1, g.L120, g.L120, y, 18), m.conv_post);
  hift_istft(ctx, y, wav, g.L120.start, g.LT.len, sstart, g.LT.n_seq, g.LT.max_len, trim_fade);
}

}  // namespace cbx
