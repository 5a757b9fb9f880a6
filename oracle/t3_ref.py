"""CPU fp32 restatement of the reference T3 inference path (TEST INFRASTRUCTURE - the oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file.  The product path (chatterbox_b200/) never does.

Each function cites the reference lines it follows (paths relative to
/root/reference/src/chatterbox/models/t3/ unless stated).  `modeling_llama.py` /
`logits_process.py` / `modeling_rope_utils.py` are the un-vendored `transformers` files (reference
pins 5.2.0, pyproject.toml:22; behaviour restated from the installed 5.5.0).

Pinned: tests/golden/t3_*.pt were produced by the *real* reference `T3.inference`
(oracle/make_golden.py, run in the authoring container) and tests/test_oracle_pinned.py checks this
restatement against them.  The reference itself ships no golden vectors.
"""
import math

import torch
import torch.nn.functional as F

START_TEXT, STOP_TEXT = 255, 0          # modules/t3_config.py:6-7
START_SPEECH, STOP_SPEECH = 6561, 6562  # modules/t3_config.py:11-12
N_HEADS, HEAD_DIM = 16, 64
RMS_EPS = 1e-5                          # llama_configs.py:21


def llama3_inv_freq(head_dim=64, base=500000.0, factor=8.0, low=1.0, high=4.0, orig=8192):
    """modeling_rope_utils.py:_compute_llama3_parameters (installed 5.5.0: lines 606-624) with the
    reference's rope_scaling (llama_configs.py:23-30)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(torch.float) / head_dim))
    low_w = orig / low
    high_w = orig / high
    wavelen = 2 * math.pi / inv_freq
    inv_l = torch.where(wavelen > low_w, inv_freq / factor, inv_freq)
    smooth = (orig / wavelen - low) / (high - low)
    smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
    is_med = ~(wavelen < high_w) * ~(wavelen > low_w)
    return torch.where(is_med, smoothed, inv_l)


def rope_tables(n_pos, head_dim=64):
    """cos/sin [n_pos, head_dim/2] in fp32 (modeling_llama.py LlamaRotaryEmbedding.forward :124-135:
    freqs = inv_freq @ position (fp32), cos/sin of it; attention_scaling = 1)."""
    inv = llama3_inv_freq(head_dim)
    pos = torch.arange(n_pos, dtype=torch.float32)
    freqs = (inv[None, :, None].float() @ pos[None, None, :].float()).transpose(1, 2)[0]  # [n_pos, hd/2]
    return freqs.cos(), freqs.sin()


def rms_norm(x, w, eps=RMS_EPS):
    """modeling_llama.py LlamaRMSNorm.forward :62-67."""
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + eps))


def apply_rope(x, cos, sin):
    """modeling_llama.py rotate_half/apply_rotary_pos_emb :138-167.  x [B,H,S,64]; cos/sin [S,32]."""
    c = torch.cat((cos, cos), -1)[None, None]
    s = torch.cat((sin, sin), -1)[None, None]
    x1, x2 = x[..., :32], x[..., 32:]
    rot = torch.cat((-x2, x1), -1)
    return x * c + rot * s


class T3Oracle:
    def __init__(self, sd, n_layers=30, max_pos=4096):
        self.sd = sd
        self.n_layers = n_layers
        self.cos, self.sin = rope_tables(max_pos)

    # ---- conditioning ------------------------------------------------------------------
    def _attn_block2(self, x1, x2):
        """modules/perceiver.py AttentionBlock2.forward :156-170 with AttentionQKV flash path :92-100
        (F.scaled_dot_product_attention default scale 1/sqrt(256); 4 heads of 256)."""
        sd, p = self.sd, "cond_enc.perceiver.attn."
        ln = lambda t: F.layer_norm(t, (1024,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
        x1n, x2n = ln(x1), ln(x2)
        q = F.linear(x1n, sd[p + "to_q.weight"], sd[p + "to_q.bias"])
        k = F.linear(x2n, sd[p + "to_k.weight"], sd[p + "to_k.bias"])
        v = F.linear(x2n, sd[p + "to_v.weight"], sd[p + "to_v.bias"])
        sp = lambda t: t.view(t.shape[0], t.shape[1], 4, 256).permute(0, 2, 1, 3)
        o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))
        o = o.permute(0, 2, 1, 3).reshape(x1.shape[0], x1.shape[1], 1024)
        return x1 + F.linear(o, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])

    def prepare_conditioning(self, speaker_emb, cond_prompt_speech_tokens, emotion_adv):
        """t3.py:92-100 + modules/cond_enc.py:64-97 + modules/perceiver.py:200-212 -> [1, 34, 1024]."""
        sd = self.sd
        emb = sd["speech_emb.weight"][cond_prompt_speech_tokens]                       # t3.py:97
        n = cond_prompt_speech_tokens.shape[1]
        emb = emb + sd["speech_pos_emb.emb.weight"][torch.arange(n)]                   # t3.py:99
        spk = F.linear(speaker_emb.view(-1, 256), sd["cond_enc.spkr_enc.weight"],
                       sd["cond_enc.spkr_enc.bias"])[:, None]                          # cond_enc.py:70
        query = sd["cond_enc.perceiver.pre_attention_query"].expand(emb.shape[0], -1, -1)
        pre = self._attn_block2(query, emb)                                            # perceiver.py:209
        perc = self._attn_block2(pre, pre)                                             # perceiver.py:211
        emo = F.linear(emotion_adv.view(-1, 1, 1).float(), sd["cond_enc.emotion_adv_fc.weight"])  # :88
        return torch.cat((spk, perc, emo), dim=1)                                      # cond_enc.py:91-96

    def prepare_input_embeds(self, cond_emb, text_tokens, cfg_weight):
        """t3.py:102-130 (+ the appended BOS of t3.py:305-313). text_tokens [rows, n] incl. SOT/EOT."""
        sd = self.sd
        text_emb = sd["text_emb.weight"][text_tokens].clone()
        if cfg_weight > 0.0:
            text_emb[1].zero_()                                                        # t3.py:113-114
        nt = text_tokens.shape[1]
        text_emb = text_emb + sd["text_pos_emb.emb.weight"][torch.arange(nt)]          # t3.py:118
        bos = sd["speech_emb.weight"][START_SPEECH] + sd["speech_pos_emb.emb.weight"][0]  # t3.py:116-119
        rows = text_tokens.shape[0]
        bos = bos[None, None].expand(rows, 1, -1)
        cond = cond_emb.expand(rows, -1, -1)
        # [cond | text | speech(BOS)] then BOS again (t3.py:126-129, 305-313)
        return torch.cat((cond, text_emb, bos, bos), dim=1)

    # ---- backbone ----------------------------------------------------------------------
    def _layer(self, i, x, pos0, cache):
        """modeling_llama.py LlamaDecoderLayer/LlamaAttention/LlamaMLP forward.  KV cache is grown with
        torch.cat exactly like transformers DynamicLayer.update (SURVEY.md 2a, 42% of CPU decode time)."""
        sd, p = self.sd, f"tfmr.layers.{i}."
        B, S, _ = x.shape
        h = rms_norm(x, sd[p + "input_layernorm.weight"])
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(B, S, N_HEADS, HEAD_DIM).transpose(1, 2)
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(B, S, N_HEADS, HEAD_DIM).transpose(1, 2)
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(B, S, N_HEADS, HEAD_DIM).transpose(1, 2)
        cos, sin = self.cos[pos0:pos0 + S], self.sin[pos0:pos0 + S]
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        if cache[i] is not None:
            k = torch.cat([cache[i][0], k], dim=-2)
            v = torch.cat([cache[i][1], v], dim=-2)
        cache[i] = (k, v)
        o = F.scaled_dot_product_attention(q, k, v, is_causal=(S > 1))                 # scale 1/8
        o = o.transpose(1, 2).reshape(B, S, N_HEADS * HEAD_DIM)
        x = x + F.linear(o, sd[p + "self_attn.o_proj.weight"])
        h = rms_norm(x, sd[p + "post_attention_layernorm.weight"])
        g = F.linear(h, sd[p + "mlp.gate_proj.weight"])
        u = F.linear(h, sd[p + "mlp.up_proj.weight"])
        return x + F.linear(F.silu(g) * u, sd[p + "mlp.down_proj.weight"])

    def backbone(self, x, pos0, cache):
        """LlamaModel.forward + final norm + speech_head (inference/t3_hf_backend.py:93-104)."""
        for i in range(self.n_layers):
            x = self._layer(i, x, pos0, cache)
        h = rms_norm(x, self.sd["tfmr.norm.weight"])
        return F.linear(h, self.sd["speech_head.weight"]), h

    # ---- sampling ----------------------------------------------------------------------
    @staticmethod
    def process_logits(cond, uncond, generated_ids, cfg_weight, repetition_penalty, temperature, min_p, top_p):
        """t3.py:339-356 with transformers logits_process.py RepetitionPenaltyLogitsProcessor,
        MinPLogitsWarper, TopPLogitsWarper restated.  cond/uncond [1,V]."""
        logits = cond + cfg_weight * (cond - uncond)                                  # t3.py:344
        ids = generated_ids
        score = torch.gather(logits, 1, ids)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        logits = logits.scatter(1, ids, score)
        if temperature != 1.0:
            logits = logits / temperature
        # MinP (min_tokens_to_keep=1): logits_process.py MinPLogitsWarper.__call__ (5.5.0 :752-768)
        probs = torch.softmax(logits, dim=-1)
        top_probs = probs.amax(dim=-1, keepdim=True)
        remove = probs < (min_p * top_probs)
        remove.scatter_(-1, torch.topk(probs, 1, dim=-1).indices, False)
        logits = logits.masked_fill(remove, -float("inf"))
        # TopP (ascending sort, keep >=1)
        sl, si = torch.sort(logits, descending=False)
        cum = sl.softmax(dim=-1).cumsum(dim=-1)
        srem = cum <= (1 - top_p)
        srem[..., -1:] = False
        rem = srem.scatter(1, si, srem)
        return logits.masked_fill(rem, -float("inf"))

    @torch.inference_mode()
    def inference(self, t3_cond, text_tokens, max_new_tokens, temperature=0.8, top_p=0.95, min_p=0.05,
                  repetition_penalty=1.2, cfg_weight=0.5, return_logits=False, q_noise=None):
        """t3.py:225-390.  text_tokens [2, n] (CFG pair, SOT/EOT already added, tts.py:237-243).
        Sampling uses torch.multinomial on the *global* CPU generator like the reference (t3.py:360)
        unless q_noise [steps, V] is given (multinomial(p,1) == argmax(p/q), q ~ Exp(1))."""
        text_tokens = torch.atleast_2d(text_tokens).long()
        cond_emb = self.prepare_conditioning(t3_cond["speaker_emb"], t3_cond["cond_prompt_speech_tokens"],
                                             t3_cond["emotion_adv"])
        x = self.prepare_input_embeds(cond_emb, text_tokens, cfg_weight)
        cache = [None] * self.n_layers
        logits, _ = self.backbone(x, 0, cache)
        pos = x.shape[1]
        generated = torch.tensor([[START_SPEECH]], dtype=torch.long)
        predicted, all_logits = [], []
        sd = self.sd
        for i in range(max_new_tokens):
            step = logits[:, -1, :]
            if return_logits:
                all_logits.append(step.clone())
            proc = self.process_logits(step[0:1], step[1:2], generated, cfg_weight, repetition_penalty,
                                       temperature, min_p, top_p)
            probs = torch.softmax(proc, dim=-1)
            if q_noise is not None:
                nxt = torch.argmax(probs / q_noise[i][None], dim=-1, keepdim=True)
            else:
                nxt = torch.multinomial(probs, num_samples=1)
            predicted.append(nxt)
            generated = torch.cat([generated, nxt], dim=1)
            if nxt.view(-1) == STOP_SPEECH:
                break
            e = sd["speech_emb.weight"][nxt] + sd["speech_pos_emb.emb.weight"][i + 1]  # t3.py:371-372
            e = torch.cat([e, e])
            logits, _ = self.backbone(e, pos, cache)
            pos += 1
        toks = torch.cat(predicted, dim=1)
        if return_logits:
            return toks, torch.stack(all_logits)
        return toks
