"""CPU restatement of the Turbo T3 path: GPT-2 backbone + `T3.inference_turbo`.

TEST INFRASTRUCTURE (oracle/): imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs; never by
the product path.  Pinned against the unmodified reference (oracle/make_golden.py -> tests/golden/turbo_golden.pt,
checked by tests/test_oracle_pinned.py).

Reference call sites restated here:
  * reference src/chatterbox/models/t3/t3.py:392-468      T3.inference_turbo (processor order temperature -> top-k ->
    top-p -> repetition penalty, first token sampled from the prefill, EOS stripped)
  * reference src/chatterbox/models/t3/t3.py:92-130       prepare_conditioning / prepare_input_embeds with is_gpt
    (no learned position tables, no CFG zeroing)
  * reference src/chatterbox/models/t3/modules/cond_enc.py:64-97  T3CondEnc without perceiver / emotion
  * reference src/chatterbox/tts_turbo.py:151-166         hyper-parameters of the Turbo T3
The backbone arithmetic lives in a third-party dependency that is not vendored under /root/reference:
`transformers` (pyproject pin 5.2.0, installed 5.5.0), `models/gpt2/modeling_gpt2.py` GPT2Model / GPT2Block /
GPT2Attention / GPT2MLP, `pytorch_utils.Conv1D` (y = x @ W[in,out] + b), `activations.NewGELUActivation`, and
`generation/logits_process.py` Temperature / TopK / TopP / RepetitionPenalty processors.  Their published
algorithm is restated below; parity is anchored on the reference's own call sites by running the real
`T3(hp).inference_turbo` on the same seeded weights (fixture generator committed).
"""
import math

import torch
import torch.nn.functional as F

START_SPEECH, STOP_SPEECH = 6561, 6562
N_HEADS, HEAD_DIM, DIM = 16, 64, 1024
LN_EPS = 1e-5


def gelu_new(x):
    """transformers activations.NewGELUActivation (tanh approximation, 'gelu_new' in llama_configs.py:36)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class TurboOracle:
    def __init__(self, sd, n_layers=24):
        self.sd = sd
        self.n_layers = n_layers

    # ---- conditioning / embeddings -------------------------------------------------------
    def prepare_conditioning(self, speaker_emb, cond_prompt_speech_tokens):
        """t3.py:92-100 with is_gpt (no speech_pos_emb) + cond_enc.py:64-97 without perceiver/emotion:
        [spkr_enc(speaker_emb) | speech_emb(prompt tokens)] -> [1, 1 + n_prompt, 1024]."""
        sd = self.sd
        emb = sd["speech_emb.weight"][cond_prompt_speech_tokens]
        spk = F.linear(speaker_emb.view(-1, 256), sd["cond_enc.spkr_enc.weight"], sd["cond_enc.spkr_enc.bias"])[:, None]
        return torch.cat((spk, emb), dim=1)

    def prepare_input_embeds(self, cond_emb, text_tokens):
        """t3.py:102-130 with is_gpt, cfg_weight 0: [cond | text_emb(text) | speech_emb(BOS)] (t3.py:407-413)."""
        sd = self.sd
        text_emb = sd["text_emb.weight"][text_tokens]
        bos = sd["speech_emb.weight"][START_SPEECH][None, None].expand(text_tokens.shape[0], 1, -1)
        return torch.cat((cond_emb.expand(text_tokens.shape[0], -1, -1), text_emb, bos), dim=1)

    # ---- backbone (transformers GPT2Model with inputs_embeds) ------------------------------
    def _block(self, i, x, cache):
        sd, p = self.sd, f"tfmr.h.{i}."
        B, S, _ = x.shape
        h = F.layer_norm(x, (DIM,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], LN_EPS)
        qkv = h @ sd[p + "attn.c_attn.weight"] + sd[p + "attn.c_attn.bias"]            # Conv1D
        q, k, v = qkv.split(DIM, dim=2)
        sp = lambda t: t.view(B, S, N_HEADS, HEAD_DIM).transpose(1, 2)
        q, k, v = sp(q), sp(k), sp(v)
        if cache[i] is not None:
            k = torch.cat([cache[i][0], k], dim=-2)
            v = torch.cat([cache[i][1], v], dim=-2)
        cache[i] = (k, v)
        o = F.scaled_dot_product_attention(q, k, v, is_causal=(S > 1))                  # scale 1/sqrt(64)
        o = o.transpose(1, 2).reshape(B, S, DIM)
        x = x + (o @ sd[p + "attn.c_proj.weight"] + sd[p + "attn.c_proj.bias"])
        h = F.layer_norm(x, (DIM,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], LN_EPS)
        h = gelu_new(h @ sd[p + "mlp.c_fc.weight"] + sd[p + "mlp.c_fc.bias"])
        return x + (h @ sd[p + "mlp.c_proj.weight"] + sd[p + "mlp.c_proj.bias"])

    def backbone(self, x, pos0, cache):
        """GPT2Model.forward(inputs_embeds=x, past_key_values): adds wpe[pos0 : pos0+S], runs the blocks and ln_f;
        then the reference's speech_head (Linear with bias when is_gpt, t3.py:83)."""
        sd = self.sd
        S = x.shape[1]
        h = x + sd["tfmr.wpe.weight"][pos0:pos0 + S][None]
        for i in range(self.n_layers):
            h = self._block(i, h, cache)
        h = F.layer_norm(h, (DIM,), sd["tfmr.ln_f.weight"], sd["tfmr.ln_f.bias"], LN_EPS)
        return F.linear(h, sd["speech_head.weight"], sd["speech_head.bias"]), h

    # ---- sampling ---------------------------------------------------------------------------
    @staticmethod
    def process_logits(logits, input_ids, temperature, top_k, top_p, repetition_penalty):
        """The LogitsProcessorList of t3.py:396-404 in its order.  logits [1,V], input_ids [1,n]."""
        if temperature > 0 and temperature != 1.0:
            logits = logits / temperature                                           # TemperatureLogitsWarper
        if top_k > 0:
            k = min(top_k, logits.shape[-1])                                        # TopKLogitsWarper
            kth = torch.topk(logits, k)[0][..., -1, None]
            logits = logits.masked_fill(logits < kth, -float("inf"))
        if top_p < 1.0:                                                             # TopPLogitsWarper (keep >= 1)
            sl, si = torch.sort(logits, descending=False)
            cum = sl.softmax(dim=-1).cumsum(dim=-1)
            srem = cum <= (1 - top_p)
            srem[..., -1:] = False
            logits = logits.masked_fill(srem.scatter(1, si, srem), -float("inf"))
        if repetition_penalty != 1.0:                                               # RepetitionPenaltyLogitsProcessor
            score = torch.gather(logits, 1, input_ids)
            score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
            logits = logits.scatter(1, input_ids, score)
        return logits

    @torch.inference_mode()
    def inference_turbo(self, t3_cond, text_tokens, temperature=0.8, top_k=1000, top_p=0.95, repetition_penalty=1.2,
                        max_gen_len=1000, q_noise=None, return_logits=False):
        """t3.py:392-468.  text_tokens [1, n] (GPT-2 tokenizer ids, no SOT/EOT).  Sampling draws from torch's global
        CPU generator like the reference unless q_noise [steps, V] is given (multinomial(p,1) == argmax(p/q))."""
        text_tokens = torch.atleast_2d(text_tokens).long()
        cond = self.prepare_conditioning(t3_cond["speaker_emb"], t3_cond["cond_prompt_speech_tokens"])
        x = self.prepare_input_embeds(cond, text_tokens)
        cache = [None] * self.n_layers
        logits, _ = self.backbone(x, 0, cache)
        pos = x.shape[1]
        generated, all_logits = [], []
        hist = torch.tensor([[START_SPEECH]], dtype=torch.long)      # first call sees the BOS id only (t3.py:428)
        for i in range(max_gen_len + 1):
            step = logits[:, -1, :]
            if return_logits:
                all_logits.append(step.clone())
            proc = self.process_logits(step, hist, temperature, top_k, top_p, repetition_penalty)
            if i > 0 and torch.all(proc == -float("inf")):           # t3.py:448-450
                break
            probs = torch.softmax(proc, dim=-1)
            if q_noise is not None:
                nxt = torch.argmax(probs / q_noise[i][None], dim=-1, keepdim=True)
            else:
                nxt = torch.multinomial(probs, num_samples=1)
            generated.append(nxt)
            hist = torch.cat(generated, dim=1)                       # later calls: generated tokens only (t3.py:446)
            if i > 0 and torch.all(nxt == STOP_SPEECH):              # the loop body checks EOS, the prefill token not
                break
            if i == max_gen_len:
                break
            e = self.sd["speech_emb.weight"][nxt]                    # t3.py:437 (wpe is added inside the backbone)
            logits, _ = self.backbone(e, pos, cache)
            pos += 1
        toks = torch.cat(generated, dim=1)
        if toks.size(1) > 0 and toks[0, -1] == STOP_SPEECH:          # t3.py:465-466
            toks = toks[:, :-1]
        if return_logits:
            return toks, torch.stack(all_logits)
        return toks
