"""T3 module boundary (reference src/chatterbox/models/t3/t3.py, modules/cond_enc.py:11-38)."""
from dataclasses import dataclass
from typing import Optional

import torch

from .engine import Engine, START_SPEECH, STOP_SPEECH


@dataclass
class T3Cond:
    """Same fields as the reference dataclass (modules/cond_enc.py:11-38)."""
    speaker_emb: torch.Tensor
    clap_emb: Optional[torch.Tensor] = None
    cond_prompt_speech_tokens: Optional[torch.Tensor] = None
    cond_prompt_speech_emb: Optional[torch.Tensor] = None
    emotion_adv: Optional[torch.Tensor] = 0.5

    def to(self, *, device=None, dtype=None):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                is_fp = v.is_floating_point()
                setattr(self, k, v.to(device=device, dtype=dtype if is_fp else None))
        return self


class T3:
    """Drop-in for the reference T3 inference surface; weights live in the engine."""

    start_speech_token, stop_speech_token = START_SPEECH, STOP_SPEECH
    start_text_token, stop_text_token = 255, 0

    def __init__(self, engine: Engine, state_dict=None):
        self.engine = engine
        if state_dict is not None:
            engine.load_t3(state_dict)

    def prepare_conditioning(self, t3_cond: T3Cond):
        """reference t3.py:92-100 -> (1, 34, 1024) on the device (Turbo: (1, 1 + n_prompt, 1024))."""
        emo = t3_cond.emotion_adv if t3_cond.emotion_adv is not None else 0.0
        emo = emo if torch.is_tensor(emo) else torch.tensor(float(emo))
        return self.engine.t3_cond(t3_cond.speaker_emb, t3_cond.cond_prompt_speech_tokens, emo.reshape(-1)[:1])

    @torch.inference_mode()
    def inference(self, *, t3_cond: T3Cond, text_tokens, initial_speech_tokens=None, prepend_prompt_speech_tokens=None,
                  num_return_sequences=1, max_new_tokens=None, stop_on_eos=True, do_sample=True, temperature=0.8,
                  top_p=0.95, min_p=0.05, length_penalty=1.0, repetition_penalty=1.2, cfg_weight=0.5,
                  q_noise=None, seed=0, kv_dtype="bf16"):
        """reference t3.py:225-390: same keyword surface (num_return_sequences / stop_on_eos / do_sample /
        length_penalty are accepted and ignored there too).  text_tokens: (2, n) CFG pair or (n,); returns
        LongTensor (1, n_generated) incl. EOS if hit.  Extra kwargs: q_noise [steps, 8194] Exp(1) draws for
        torch.multinomial parity, seed for the device RNG otherwise, kv_dtype 'bf16' | 'fp32'."""
        assert prepend_prompt_speech_tokens is None, "not implemented"
        text_tokens = torch.atleast_2d(text_tokens).to(torch.long)
        assert (text_tokens == self.start_text_token).int().sum() >= text_tokens.size(0), "missing start_text_token"
        assert (text_tokens == self.stop_text_token).int().sum() >= text_tokens.size(0), "missing stop_text_token"
        cond = self.prepare_conditioning(t3_cond)
        qn = q_noise[:, None, :] if q_noise is not None else None
        toks = self.engine.t3_generate([text_tokens[0].cpu()], cond, max_new_tokens=max_new_tokens or 4096,
                                       cfg_weight=cfg_weight if text_tokens.size(0) > 1 else 0.0,
                                       temperature=temperature, top_p=top_p, min_p=min_p,
                                       repetition_penalty=repetition_penalty, q_noise=qn, seed=seed, kv_dtype=kv_dtype)
        return toks[0][None]

    @torch.inference_mode()
    def inference_turbo(self, t3_cond: T3Cond, text_tokens, temperature=0.8, top_k=1000, top_p=0.95,
                        repetition_penalty=1.2, max_gen_len=1000, q_noise=None, seed=0, kv_dtype="bf16"):
        """reference t3.py:392-468 (Turbo checkpoint: GPT-2 backbone, no CFG).  text_tokens (1, n) tokenizer ids;
        returns LongTensor (1, n_generated) with a trailing EOS stripped, like the reference (t3.py:465-466).
        Extra kwargs as in inference(): q_noise [steps, V] Exp(1) draws for torch.multinomial parity."""
        assert self.engine.t3_turbo, "inference_turbo needs a Turbo (GPT-2 backbone) checkpoint"
        text_tokens = torch.atleast_2d(text_tokens).to(torch.long)
        cond = self.prepare_conditioning(t3_cond)
        qn = None
        if q_noise is not None:          # the engine's noise rows keep the 8194 stride of the speech vocabulary
            qn = torch.ones(q_noise.shape[0], 1, 8194, dtype=torch.float32)
            qn[:, 0, :q_noise.shape[1]] = q_noise
        toks = self.engine.t3_generate([text_tokens[0].cpu()], cond, max_new_tokens=max_gen_len + 1, cfg_weight=0.0,
                                       temperature=temperature, top_p=top_p, min_p=0.0,
                                       repetition_penalty=repetition_penalty, q_noise=qn, seed=seed, kv_dtype=kv_dtype,
                                       top_k=top_k)
        t = toks[0]
        self._last_turbo_draws = int(t.numel())      # tokens sampled incl. a trailing EOS (one multinomial draw each)
        if t.numel() > 0 and int(t[-1]) == self.stop_speech_token:
            t = t[:-1]
        return t[None]
