// Declarations of the small kernels in elem.cu (model-specific glue that is not a GEMM or attention).
#pragma once
#include "ops.h"
#include <cuda_fp16.h>

namespace cbx {

struct T3SampleDev {
  const float* logits; int ldl;          // [n_slots][ldl]; slots (2j, 2j+1) = (cond, uncond) rows of active utt j
  const int* act_utt;                    // [n_act] active utterance ids
  const int* n_act;                      // device scalar: live entries of act_utt (CTAs beyond it exit); may be null
  const int* src_slot;                   // [n_act] slot whose logits row belongs to active entry j (set by t3_compact); may be null
  const int* force_tokens;               // optional [B][max_tokens]: teacher forcing (feed these ids, tests only)
  int* sampled_out;                      // with force_tokens: the ids the sampler itself picked
  int cfg; int n_utts;
  float cfg_weight, rep_penalty, temperature, min_p, top_p;
  int eos_id;
  int* tokens; int max_tokens;           // [B][max_tokens]
  int* n_gen; const int* max_new; int* done;   // [B]
  unsigned char* seen;                   // [B][8194] repetition-penalty history (BOS pre-set)
  int* positions; const int* base_pos;   // [R]
  float* x;                              // [n_slots][1024] next input embedding
  const float* speech_emb; const float* speech_pos;
  const float* q_noise;                  // optional [steps][B][8194] Exp(1) noise (parity mode), else counter RNG
  unsigned long long seed;
  int vocab;                             // speech ids scored (8194; Turbo 6563)
  int turbo;                             // 1: T3.inference_turbo processor order (temperature, top-k, top-p, repetition penalty)
  int top_k; int bos_id;
  const float* wpe;                      // Turbo: learned absolute positions, added to the next input embedding
};

void ew_act(Ctx& ctx, const float* x, int ldx, float* y, int ldy, long rows, int cols, int act, float p, const float* vec);
void copy2d(Ctx& ctx, const float* src, int lds, float* dst, int ldd, long rows, int cols);
void gather_rows(Ctx& ctx, const float* table, int ld, const int* ids, float* out, int ldo, int rows, int dim,
                 const float* add_table, int add_ld, const int* add_ids, int id_limit);
void pack_hilo(Ctx& ctx, const float* src, int ld, int N, int K, __nv_bfloat16* hi, __nv_bfloat16* lo, int Npad, int Kpad);
void pack_hilo_cat(Ctx& ctx, const float* src, int ld, int N, int K, __nv_bfloat16* out, int Npad);
void pack_planes_seq(Ctx& ctx, const float* src, int ld, long rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, int ldp,
                     const SeqMap& seq);
void t3_embed(Ctx& ctx, float* out, int n_tok, const int* tok_row, const int* tok_pos, const float* cond,
              const int* row_voice, int len_cond, const int* text_flat, const int* text_start, const int* n_text,
              const int* row_uncond, const float* text_emb, int text_vocab, const float* text_pos,
              const float* speech_emb, const float* speech_pos, int bos_id, const float* wpe = nullptr);
void t3_sample(Ctx& ctx, const T3SampleDev& p, int n_act);
void t3_sample_init();
void t3_compact(Ctx& ctx, int* act_utt, int* n_act, int* src_slot, int* slot_row, int* m_live, const int* done, int rows_per);
struct ResidNormDev {
  float* x; int ldx;                     // residual stream, updated in place when nsplit > 0
  const float* part; int nsplit; long split_stride; int ldp;   // split-K partial sums [nsplit][rows][ldp]
  const float* bias;                     // bias of the projection that produced `part` (GPT-2), or null
  const float* w; const float* b;        // norm weight (null: residual update only) and LayerNorm bias
  int layernorm; float eps;
  __nv_bfloat16* yhi; __nv_bfloat16* ylo; float* y; int ldy;   // bf16 hi/lo planes, or fp32 when yhi == null
  __half* y16;                           // one fp16 plane instead (fp16-activation decode mode)
  int dim;
  const int* m_live;                     // optional device scalar: rows >= *m_live exit
};
void resid_norm(Ctx& ctx, const ResidNormDev& p, int rows);
void add_pos_bias(Ctx& ctx, const float* qkv, int ld, const float* u, const float* v, float* qu, float* qv, long rows);
void relpos_table(Ctx& ctx, float* pe, int T, int d_model);
void upsample2(Ctx& ctx, const float* x, float* y, int C, const int* tile_seq2, const int* start2, const int* len2,
               const int* start1, long rows2);
void time_sinusoid(Ctx& ctx, const float* t, float* out, int n, int dim, float scale);
void cfm_assemble(Ctx& ctx, float* xin, const float* x, const float* mu, const float* spk, const float* cond,
                  const int* tile_seq3, const int* start3, const int* len3, const int* start2, int B, long rows3,
                  int write_static);
void cfm_euler(Ctx& ctx, float* x, const float* v, const int* tile_seq2, const int* start2, const int* len2,
               const int* start3, int B, float dt, float w, int cfg, long rows2);
void hift_source(Ctx& ctx, const float* f0, float* cumf, const float* phase_vec, const float* noise, const float* lin_w,
                 float lin_b, float* s_out, const int* startT, const int* lenT, const long* startS, int n_seq, int maxT,
                 unsigned long long seed, long n_frame_rows);
void f0_head(Ctx& ctx, const float* x, int ld, const float* w, float b, float* f0, long rows);
void hift_stft(Ctx& ctx, const float* s, float* out, const int* startF, const int* lenT, const long* startS, int n_seq,
               int maxT);
void reflect_row0(Ctx& ctx, float* x, int C, const int* start, int n_seq);
void hift_istft(Ctx& ctx, const float* y, float* wav, const int* startF, const int* lenT, const long* startS, int n_seq,
                int maxT, int trim_fade);

}  // namespace cbx
