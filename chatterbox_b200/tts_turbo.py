"""Public API boundary of the Turbo model (reference src/chatterbox/tts_turbo.py:104-321)."""
from pathlib import Path

import numpy as np
import torch

from .engine import Engine
from .t3 import T3
from .s3gen import S3Gen, S3GEN_SR, SPEECH_VOCAB_SIZE
from .tts import Conditionals, punc_norm, synthesize_batch, apply_watermark

S3GEN_SIL = 4299            # reference models/s3gen/const.py:2
TURBO_SPEECH_VOCAB = 6563   # tts_turbo.py:155


class ChatterboxTurboTTS:
    """Drop-in for reference ChatterboxTurboTTS: GPT2_medium T3 (`T3.inference_turbo`, no CFG) + 2-step meanflow S3Gen
    + HiFT; every FLOP of generate() runs in libcbx.  Nano (GPT2_small, 768 wide) is not supported."""

    def __init__(self, t3: T3, s3gen: S3Gen, tokenizer, device, conds: Conditionals = None):
        self.sr = S3GEN_SR
        self.t3, self.s3gen, self.tokenizer, self.device, self.conds = t3, s3gen, tokenizer, device, conds
        self.engine = t3.engine

    @classmethod
    def from_state_dicts(cls, t3_sd, flow_sd, hift_sd, conds=None, tokenizer=None, device="cuda"):
        idx = torch.device(device).index or 0
        eng = Engine(idx)
        t3 = T3(eng, t3_sd)
        assert eng.t3_turbo, "not a Turbo T3 checkpoint (no tfmr.h.* keys)"
        s3 = S3Gen(eng, flow_sd, hift_sd, meanflow=True)
        return cls(t3, s3, tokenizer, device, conds)

    @classmethod
    def from_local(cls, ckpt_dir, device="cuda"):
        """reference tts_turbo.py:133-187: t3_turbo_v1.safetensors, s3gen_meanflow.safetensors, HF tokenizer files,
        conds.pt."""
        from safetensors.torch import load_file
        ckpt_dir = Path(ckpt_dir)
        t3_sd = load_file(ckpt_dir / "t3_turbo_v1.safetensors")
        if "model" in t3_sd:
            t3_sd = t3_sd["model"][0]
        s3 = load_file(ckpt_dir / "s3gen_meanflow.safetensors")
        flow_sd = {k[len("flow."):]: v for k, v in s3.items() if k.startswith("flow.")}
        hift_sd = {k[len("mel2wav."):]: v for k, v in s3.items() if k.startswith("mel2wav.")}
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(ckpt_dir)
        if tok.pad_token is None:
            tok.pad_token = tok.eos_token
        conds = Conditionals.load(ckpt_dir / "conds.pt") if (ckpt_dir / "conds.pt").exists() else None
        return cls.from_state_dicts(t3_sd, flow_sd, hift_sd, conds, tok, device)

    def prepare_conditionals(self, wav_fpath, exaggeration=0.0, norm_loudness=True):
        raise NotImplementedError("voice-prompt analysis (reference tts_turbo.py:223-270) is outside the B200 hot path; "
                                  "load a Conditionals object produced by the reference instead")

    @torch.inference_mode()
    def generate(self, text, repetition_penalty=1.2, min_p=0.0, top_p=0.95, audio_prompt_path=None, exaggeration=0.0,
                 cfg_weight=0.0, temperature=0.8, top_k=1000, norm_loudness=True, rng="torch_cpu", kv_dtype=None,
                 watermark=True):
        """reference tts_turbo.py:272-321 (CFG / min_p / exaggeration are ignored there too).  The output is watermarked
        on the host like the reference's (tts_turbo.py:319) unless watermark=False."""
        assert audio_prompt_path is None, "prepare_conditionals is outside the hot path; set .conds"
        assert self.tokenizer is not None, "no tokenizer loaded; pass token ids to generate_tokens()"
        ids = self.tokenizer(punc_norm(text), return_tensors="pt", padding=True, truncation=True).input_ids
        wav = self.generate_tokens(ids, repetition_penalty=repetition_penalty, top_p=top_p, temperature=temperature,
                                   top_k=top_k, rng=rng, kv_dtype=kv_dtype)
        return apply_watermark(wav, self.sr, watermark)

    @torch.inference_mode()
    def generate_batch(self, text_tokens, max_gen_len=1000, repetition_penalty=1.2, top_p=0.95, temperature=0.8,
                       top_k=1000, seed=0, kv_dtype="bf16", to_host=True, timings=None):
        """Batched generate(): equal to calling the reference's Turbo generate() once per utterance (each with its own
        device RNG stream, seed + utterance index).  text_tokens: list of 1-D tokenizer-id tensors; max_gen_len: int or
        per-utterance list.  One T3 row per utterance (no CFG), 2-step meanflow CFM, HiFT.  Returns float32 waveforms."""
        assert self.conds is not None, "Please set .conds (Conditionals)"
        eng = self.engine
        ev = lambda: torch.cuda.Event(enable_timing=True)
        marks = [ev() for _ in range(4)]
        marks[0].record()
        B = len(text_tokens)
        tts = [torch.as_tensor(t).reshape(-1).to(torch.long).cpu() for t in text_tokens]
        c = self.conds.t3
        cond = eng.t3_cond(c.speaker_emb.reshape(1, 256), c.cond_prompt_speech_tokens.reshape(1, -1), torch.zeros(1))
        budgets = [int(max_gen_len) + 1] * B if np.isscalar(max_gen_len) else [int(m) + 1 for m in max_gen_len]
        toks = eng.t3_generate(tts, cond, max_new_tokens=budgets, cfg_weight=0.0, temperature=temperature, top_p=top_p,
                               min_p=0.0, repetition_penalty=repetition_penalty, seed=seed, kv_dtype=kv_dtype, top_k=top_k)
        marks[1].record()
        sil = torch.tensor([S3GEN_SIL] * 3, dtype=torch.long)
        speech = []
        for t in toks:                                        # tts_turbo.py:307-311 per utterance
            if t.numel() > 0 and int(t[-1]) == 6562:
                t = t[:-1]                                    # t3.py:465-466
            speech.append(torch.cat([t[t < SPEECH_VOCAB_SIZE], sil]))
        refs = [self.conds.gen] * B
        wavs = synthesize_batch(eng, speech, refs, seed=seed, n_cfm_timesteps=2, marks=marks[2:4])
        if to_host:
            wavs = [w.cpu() for w in wavs]
        torch.cuda.synchronize()
        if timings is not None:
            timings.update(t3_ms=marks[0].elapsed_time(marks[1]), flow_ms=marks[1].elapsed_time(marks[2]),
                           hift_ms=marks[2].elapsed_time(marks[3]),
                           audio_s=sum(int(x.numel()) for x in speech) / 25.0)
        return wavs

    @torch.inference_mode()
    def generate_tokens(self, text_tokens, repetition_penalty=1.2, top_p=0.95, temperature=0.8, top_k=1000,
                        max_gen_len=1000, rng="torch_cpu", kv_dtype=None, return_intermediates=False):
        """generate() from tokenizer ids (1, n).  rng='torch_cpu' draws every random tensor from torch's global CPU
        generator in the reference's order (multinomial per token -> meanflow noise s3gen.py:316 -> randn_like(mu)
        flow_matching.py:216 -> SineGen phases -> SineGen noise); rng='device' uses the engine's counter RNG."""
        assert self.conds is not None, "Please set .conds (Conditionals)"
        if kv_dtype is None:
            kv_dtype = "fp32" if rng == "torch_cpu" else "bf16"
        tt = torch.atleast_2d(text_tokens).to(torch.long).cpu()
        q = None
        if rng == "torch_cpu":
            state = torch.get_rng_state()
            q = torch.stack([torch.empty(TURBO_SPEECH_VOCAB).exponential_(1) for _ in range(max_gen_len + 1)])
        toks = self.t3.inference_turbo(self.conds.t3, tt, temperature=temperature, top_k=top_k, top_p=top_p,
                                       repetition_penalty=repetition_penalty, max_gen_len=max_gen_len, q_noise=q,
                                       kv_dtype=kv_dtype)
        if rng == "torch_cpu":      # leave the generator where the reference's loop would have left it:
            torch.set_rng_state(state)      # one multinomial draw per sampled token, incl. a trailing EOS stripped above
            for _ in range(self.t3._last_turbo_draws):
                torch.empty(TURBO_SPEECH_VOCAB).exponential_(1)
        st = toks[0]
        st = st[st < SPEECH_VOCAB_SIZE]                                            # tts_turbo.py:308
        st = torch.cat([st, torch.tensor([S3GEN_SIL] * 3, dtype=st.dtype)])        # tts_turbo.py:310-311
        z = phase = noise = None
        n_p = int(self.conds.gen["prompt_token"].shape[-1])
        if rng == "torch_cpu":
            noised = torch.randn(1, 80, 2 * st.numel())                            # s3gen.py:316
            z = torch.randn(1, 80, 2 * (n_p + st.numel()))[0]                      # flow_matching.py:216
            z[:, 2 * n_p:] = noised[0]                                             # flow_matching.py:218-220
        mel = self.s3gen.flow_inference(st, ref_dict=self.conds.gen, z=z, n_cfm_timesteps=2)
        if rng == "torch_cpu":
            from torch.distributions.uniform import Uniform
            phase = Uniform(low=-np.pi, high=np.pi).sample(sample_shape=(1, 9, 1))   # hifigan.py:212-214
            phase[:, 0, :] = 0
            noise = torch.randn(1, 9, 480 * mel.shape[-1])                      # hifigan.py:226
            torch.randn(1, 480 * mel.shape[-1], 1)                              # hifigan.py:282 (unused draw)
            phase, noise = phase.reshape(9), noise[0]
        wav, src = self.s3gen.hift_inference(mel, None, phase_vec=phase, noise=noise, trim_fade=True)
        out = wav.detach().cpu()
        if return_intermediates:
            return out, dict(tokens=toks, speech_tokens=st, mel=mel, source=src)
        return out
