"""Public API boundary (reference src/chatterbox/tts.py:106-272)."""
import math
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

from .engine import Engine, SPEECH_VOCAB, STOP_SPEECH
from .t3 import T3, T3Cond
from .s3gen import S3Gen, S3GEN_SR, SPEECH_VOCAB_SIZE

SOT, EOT = 255, 0


class nvtx_range:
    """NVTX range around a pipeline stage (nsys / ncu --nvtx timelines)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        torch.cuda.nvtx.range_pop()


_watermarker = None


def apply_watermark(wav, sr, enabled=True):
    """reference tts.py:271 / mtl_tts.py:354 / tts_turbo.py:319: every generate() output passes through
    perth.PerthImplicitWatermarker on the host.  The `perth` wheel is a host-side dependency outside the CUDA path; when
    it is missing the call FAILS (unwatermarked audio must be asked for explicitly with watermark=False)."""
    global _watermarker
    if not enabled:
        return wav
    if _watermarker is None:
        try:
            import perth
        except ImportError as e:
            raise RuntimeError("the reference watermarks every generate() output with perth.PerthImplicitWatermarker, but the "
                               "`perth` package is not installed; install it or pass watermark=False to get unwatermarked "
                               "audio") from e
        _watermarker = perth.PerthImplicitWatermarker()
    out = _watermarker.apply_watermark(wav.squeeze(0).detach().cpu().numpy(), sample_rate=sr)
    return torch.from_numpy(out).unsqueeze(0)


def punc_norm(text: str) -> str:
    """Restates reference tts.py:22-61 (quick punctuation cleanup of the input text)."""
    if len(text) == 0:
        return "You need to add some text for me to talk."
    if text[0].islower():
        text = text[0].upper() + text[1:]
    text = " ".join(text.split())
    for old, new in [("...", ", "), ("…", ", "), (":", ","), (" - ", ", "), (";", ", "), ("—", "-"), ("–", "-"),
                     (" ,", ","), ("“", "\""), ("”", "\""), ("‘", "'"), ("’", "'")]:
        text = text.replace(old, new)
    text = text.rstrip(" ")
    if not any(text.endswith(p) for p in {".", "!", "?", "-", ","}):
        text += "."
    return text


@dataclass
class Conditionals:
    """reference tts.py:64-103: t3 = T3Cond, gen = S3Gen ref dict."""
    t3: T3Cond
    gen: dict

    def to(self, device):
        self.t3 = self.t3.to(device=device)
        for k, v in self.gen.items():
            if torch.is_tensor(v):
                self.gen[k] = v.to(device=device)
        return self

    def save(self, fpath):
        torch.save(dict(t3=self.t3.__dict__, gen=self.gen), fpath)

    @classmethod
    def load(cls, fpath, map_location="cpu"):
        kw = torch.load(fpath, map_location=map_location, weights_only=True)
        return cls(T3Cond(**kw["t3"]), kw["gen"])


def drop_invalid_tokens(x):
    """reference models/s3tokenizer/__init__.py:16-30 for a 1-D tensor."""
    x = x.reshape(-1)
    sos = (x == SPEECH_VOCAB_SIZE).nonzero()
    s = int(sos[0]) + 1 if len(sos) else 0
    eos = (x == SPEECH_VOCAB_SIZE + 1).nonzero()
    e = int(eos[0]) if len(eos) else None
    return x[s:e]


def synthesize_batch(eng, speech, refs, seed=0, flow_frames_per_chunk=120000, hift_frames_per_chunk=24000,
                     n_cfm_timesteps=None, marks=None):
    """speech tokens (list of 1-D id tensors) -> list of device waveforms: flow (encoder + CFM) and HiFT over packed
    chunks, longest utterances first (stage part of generate_batch; shared with the Turbo front-end)."""
    B = len(speech)
    order = sorted(range(B), key=lambda b: -speech[b].numel())
    mels = [None] * B
    i = 0
    while i < B:                                          # chunk the packed batch by total mel frames
        j, frames = i, 0
        while j < B:
            f = 2 * (int(refs[order[j]]["prompt_token"].shape[-1]) + speech[order[j]].numel())
            if j > i and frames + f > flow_frames_per_chunk:
                break
            frames += f
            j += 1
        idx = order[i:j]
        with nvtx_range("cbx.flow"):
            out = eng.flow_mel([speech[b] for b in idx], [refs[b] for b in idx], n_timesteps=n_cfm_timesteps)
        for b, m in zip(idx, out):
            mels[b] = m
        i = j
    if marks is not None:
        marks[0].record()
    wavs = [None] * B
    i = 0
    while i < B:
        j, frames = i, 0
        while j < B:
            f = int(mels[order[j]].shape[-1])
            if j > i and frames + f > hift_frames_per_chunk:
                break
            frames += f
            j += 1
        idx = [b for b in order[i:j] if mels[b].shape[-1] > 0]
        if idx:
            with nvtx_range("cbx.hift"):
                w, _ = eng.hift([mels[b] for b in idx], seed=seed + i, trim_fade=True)
            for b, x in zip(idx, w):
                wavs[b] = x.clone()
        for b in order[i:j]:
            if wavs[b] is None:
                wavs[b] = torch.zeros(0, device=eng.device)
        i = j
    if marks is not None:
        marks[1].record()
    return wavs


class ChatterboxTTS:
    """Drop-in for reference ChatterboxTTS (tts.py:106-272); every FLOP of generate() runs in libcbx."""

    def __init__(self, t3: T3, s3gen: S3Gen, tokenizer, device, conds: Conditionals = None):
        self.sr = S3GEN_SR
        self.t3, self.s3gen, self.tokenizer, self.device, self.conds = t3, s3gen, tokenizer, device, conds
        self.engine = t3.engine

    @classmethod
    def from_state_dicts(cls, t3_sd, flow_sd, hift_sd, conds=None, tokenizer=None, device="cuda"):
        idx = torch.device(device).index or 0
        eng = Engine(idx)
        return cls(T3(eng, t3_sd), S3Gen(eng, flow_sd, hift_sd), tokenizer, device, conds)

    @classmethod
    def from_local(cls, ckpt_dir, device="cuda"):
        """reference tts.py:133-163: t3_cfg.safetensors, s3gen.safetensors, tokenizer.json, conds.pt."""
        from safetensors.torch import load_file
        ckpt_dir = Path(ckpt_dir)
        t3_sd = load_file(ckpt_dir / "t3_cfg.safetensors")
        if "model" in t3_sd:
            t3_sd = t3_sd["model"][0]
        s3 = load_file(ckpt_dir / "s3gen.safetensors")
        flow_sd = {k[len("flow."):]: v for k, v in s3.items() if k.startswith("flow.")}
        hift_sd = {k[len("mel2wav."):]: v for k, v in s3.items() if k.startswith("mel2wav.")}
        tok = None
        if (ckpt_dir / "tokenizer.json").exists():
            from tokenizers import Tokenizer
            tok = Tokenizer.from_file(str(ckpt_dir / "tokenizer.json"))
        conds = Conditionals.load(ckpt_dir / "conds.pt") if (ckpt_dir / "conds.pt").exists() else None
        return cls.from_state_dicts(t3_sd, flow_sd, hift_sd, conds, tok, device)

    @classmethod
    def from_pretrained(cls, device="cuda"):
        """reference tts.py:165-180 (needs the HF hub cache; there is no network in the build sandbox)."""
        from huggingface_hub import hf_hub_download
        for f in ["t3_cfg.safetensors", "s3gen.safetensors", "tokenizer.json", "conds.pt"]:
            local = hf_hub_download(repo_id="ResembleAI/chatterbox", filename=f)
        return cls.from_local(Path(local).parent, device)

    def prepare_conditionals(self, wav_fpath, exaggeration=0.5):
        raise NotImplementedError("voice-prompt analysis (reference tts.py:182-206) is outside the B200 hot path; "
                                  "load a Conditionals object produced by the reference instead")

    def text_to_tokens(self, text):
        """reference models/tokenizers/tokenizer.py:30-42 (EnTokenizer)."""
        assert self.tokenizer is not None, "no tokenizer.json loaded; pass token ids to generate_tokens()"
        ids = self.tokenizer.encode(text.replace(" ", "[SPACE]")).ids
        return torch.IntTensor(ids).unsqueeze(0)

    @torch.inference_mode()
    def generate(self, text, repetition_penalty=1.2, min_p=0.05, top_p=1.0, audio_prompt_path=None, exaggeration=0.5,
                 cfg_weight=0.5, temperature=0.8, max_new_tokens=1000, rng="torch_cpu", kv_dtype=None, watermark=True):
        """reference tts.py:208-272.  The output is watermarked on the host like the reference's (tts.py:271) unless
        watermark=False.  kv_dtype: None = fp32 KV cache in the reference-reproducing mode (rng='torch_cpu'), bf16 in
        the throughput mode (rng='device')."""
        assert audio_prompt_path is None, "prepare_conditionals is out of scope; set .conds"
        text_tokens = self.text_to_tokens(punc_norm(text))
        wav = self.generate_tokens(text_tokens, repetition_penalty, min_p, top_p, exaggeration, cfg_weight, temperature,
                                   max_new_tokens, rng, kv_dtype)
        return apply_watermark(wav, self.sr, watermark)

    @torch.inference_mode()
    def generate_batch(self, text_tokens, max_new_tokens=1000, repetition_penalty=1.2, min_p=0.05, top_p=1.0,
                       cfg_weight=0.5, temperature=0.8, seed=0, kv_dtype="bf16", to_host=True, voice_ids=None,
                       conds_list=None, flow_frames_per_chunk=120000, hift_frames_per_chunk=24000, timings=None):
        """Batched generate(): equal to calling the reference's generate() once per utterance (each with its own RNG
        stream; device counter RNG, seed + utterance index).  text_tokens: list of 1-D id tensors without SOT/EOT;
        max_new_tokens: int or per-utterance list.  Returns a list of float32 waveforms [960*N_b] (CPU if to_host)."""
        import os, sys, time
        trace = (lambda *a: print("[cbx]", f"{time.time():.1f}", *a, file=sys.stderr, flush=True)) if os.environ.get("CBX_TRACE") else (lambda *a: None)
        conds_list = conds_list or [self.conds]
        eng = self.engine
        ev = lambda: torch.cuda.Event(enable_timing=True)
        marks = [ev() for _ in range(5)]
        marks[0].record()
        B = len(text_tokens)
        tts = [F.pad(F.pad(torch.as_tensor(t).reshape(-1).to(torch.long).cpu(), (1, 0), value=SOT), (0, 1), value=EOT)
               for t in text_tokens]
        spk = torch.cat([c.t3.speaker_emb.reshape(1, 256) for c in conds_list])
        ptk = torch.cat([c.t3.cond_prompt_speech_tokens.reshape(1, -1) for c in conds_list])
        emo = torch.stack([torch.as_tensor(c.t3.emotion_adv).reshape(-1)[0] for c in conds_list]).float()
        with nvtx_range("cbx.t3"):
            cond = eng.t3_cond(spk, ptk, emo)
            toks = eng.t3_generate(tts, cond, voice_ids=voice_ids, max_new_tokens=max_new_tokens, cfg_weight=cfg_weight,
                                   temperature=temperature, top_p=top_p, min_p=min_p,
                                   repetition_penalty=repetition_penalty, seed=seed, kv_dtype=kv_dtype)
        marks[1].record()
        trace("t3 done", [int(t.numel()) for t in toks][:8])
        speech = []
        for t in toks:                                        # tts.py:257-262 per utterance
            # (the batch API always filters ids >= 6561: one stray special id must not take the whole batch down, which
            # is what the multilingual reference's unfiltered path would do, mtl_tts.py:341 -> flow.py:166)
            speech.append(ChatterboxTTS.clean_speech_tokens(t))
        vid = [0] * B if voice_ids is None else list(voice_ids)
        refs = [conds_list[v].gen for v in vid]
        wavs = synthesize_batch(eng, speech, refs, seed=seed, flow_frames_per_chunk=flow_frames_per_chunk,
                                hift_frames_per_chunk=hift_frames_per_chunk, marks=(marks[2], marks[3]))
        wavs = [self._post_wav(w, int(st.numel())) for w, st in zip(wavs, speech)]
        if to_host:
            total = sum(int(w.numel()) for w in wavs)
            host = torch.empty(total, dtype=torch.float32, pin_memory=True)
            o = 0
            outs = []
            for w in wavs:
                n = int(w.numel())
                host[o:o + n].copy_(w, non_blocking=True)
                outs.append(host[o:o + n])
                o += n
            wavs = outs
        marks[4].record()
        torch.cuda.synchronize()
        if timings is not None:
            timings.update(t3_ms=marks[0].elapsed_time(marks[1]), flow_ms=marks[1].elapsed_time(marks[2]),
                           hift_ms=marks[2].elapsed_time(marks[3]), d2h_ms=marks[3].elapsed_time(marks[4]),
                           audio_s=sum(int(x.numel()) for x in speech) / 25.0,
                           d2h_bytes=sum(int(w.numel()) for w in wavs) * 4,
                           h2d_bytes=sum(int(t.numel()) for t in tts) * 4 + int(spk.numel() + ptk.numel() + emo.numel()) * 4)
        return wavs

    @torch.inference_mode()
    def _post_wav(self, wav, n_speech_tokens):
        """per-utterance waveform post-processing of the batch API (identity here; the multilingual class trims)."""
        return wav

    @staticmethod
    def clean_speech_tokens(toks):
        """reference tts.py:257-262: cut at SOS / EOS, then keep ids < 6561."""
        st = drop_invalid_tokens(toks)
        return st[st < SPEECH_VOCAB_SIZE]

    def generate_tokens(self, text_tokens, repetition_penalty=1.2, min_p=0.05, top_p=1.0, exaggeration=0.5,
                        cfg_weight=0.5, temperature=0.8, max_new_tokens=1000, rng="torch_cpu", kv_dtype=None,
                        return_intermediates=False):
        """generate() from text token ids (1, n) without SOT/EOT.  rng='torch_cpu' draws every random tensor from
        torch's global CPU generator in the reference's order (multinomial -> randn_like(mu) -> SineGen phases ->
        SineGen noise), so that the same torch.manual_seed gives the reference's output; rng='device' uses the
        engine's counter RNG (throughput mode)."""
        assert self.conds is not None, "Please set .conds (Conditionals)"
        if kv_dtype is None:          # reproducing the reference (its RNG stream, fp32 arithmetic) keeps the KV cache in fp32
            kv_dtype = "fp32" if rng == "torch_cpu" else "bf16"
        if float(exaggeration) != float(self.conds.t3.emotion_adv.reshape(-1)[0]):
            c = self.conds.t3
            self.conds.t3 = T3Cond(speaker_emb=c.speaker_emb, cond_prompt_speech_tokens=c.cond_prompt_speech_tokens,
                                   emotion_adv=exaggeration * torch.ones(1, 1, 1))
        tt = torch.atleast_2d(text_tokens).to(torch.long).cpu()
        if cfg_weight > 0.0:
            tt = torch.cat([tt, tt], dim=0)
        tt = F.pad(F.pad(tt, (1, 0), value=SOT), (0, 1), value=EOT)
        q = None
        if rng == "torch_cpu":
            state = torch.get_rng_state()
            q = torch.stack([torch.empty(SPEECH_VOCAB).exponential_(1) for _ in range(max_new_tokens)])
        toks = self.t3.inference(t3_cond=self.conds.t3, text_tokens=tt, max_new_tokens=max_new_tokens,
                                 temperature=temperature, cfg_weight=cfg_weight, repetition_penalty=repetition_penalty,
                                 min_p=min_p, top_p=top_p, q_noise=q, kv_dtype=kv_dtype)
        if rng == "torch_cpu":      # leave the generator where the reference's loop would have left it
            torch.set_rng_state(state)
            for _ in range(toks.shape[1]):
                torch.empty(SPEECH_VOCAB).exponential_(1)
        st = self.clean_speech_tokens(toks[0])
        z = phase = noise = None
        n_p = int(self.conds.gen["prompt_token"].shape[-1])
        T = 2 * (n_p + st.numel())
        if rng == "torch_cpu":
            z = torch.randn(1, 80, T)[0]                                       # flow_matching.py:216
        mel = self.s3gen.flow_inference(st, ref_dict=self.conds.gen, z=z)
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


# reference mtl_tts.py:31-55
SUPPORTED_LANGUAGES = {
    "ar": "Arabic", "da": "Danish", "de": "German", "el": "Greek", "en": "English", "es": "Spanish", "fi": "Finnish",
    "fr": "French", "he": "Hebrew", "hi": "Hindi", "it": "Italian", "ja": "Japanese", "ko": "Korean", "ms": "Malay",
    "nl": "Dutch", "no": "Norwegian", "pl": "Polish", "pt": "Portuguese", "ru": "Russian", "sv": "Swedish",
    "sw": "Swahili", "tr": "Turkish", "zh": "Chinese",
}


def mtl_tail_trim(wav, n_speech_tokens):
    """reference mtl_tts.py:346-351: drop the audio of the final speech token (960 samples per token, keep >= 1)."""
    keep = max(1, int(n_speech_tokens) - 1) * (S3GEN_SR // 25)
    return wav[..., :keep]


def mtl_punc_norm(text: str) -> str:
    """Restates reference mtl_tts.py:70-107: the English clean-up plus the CJK sentence enders."""
    if len(text) == 0:
        return "You need to add some text for me to talk."
    if text[0].islower():
        text = text[0].upper() + text[1:]
    text = " ".join(text.split())
    for old, new in [("...", ", "), ("\u2026", ", "), (":", ","), (" - ", ", "), (";", ", "), ("\u2014", "-"), ("\u2013", "-"),
                     (" ,", ","), ("\u201c", "\""), ("\u201d", "\""), ("\u2018", "'"), ("\u2019", "'")]:
        text = text.replace(old, new)
    text = text.rstrip(" ")
    if not any(text.endswith(p) for p in {".", "!", "?", "-", ",", "\u3001", "\uff0c", "\u3002", "\uff1f", "\uff01"}):
        text += "."
    return text


class MTLTokenizer:
    """Text front-end of the multilingual model (reference models/tokenizers/tokenizer.py:256-312): lower-case, NFKD,
    `[lang]` prefix token, `[SPACE]`, then the HF tokenizer file `grapheme_mtl_merged_expanded_v1.json`.  The
    language-specific normalisers of zh / ja / he / ko (Cangjie, kakasi, dicta, jamo decomposition) live in third-party
    packages outside this repo: those languages need the reference's tokenizer object passed in instead."""

    NEEDS_NORMALISER = {"zh", "ja", "he", "ko"}

    def __init__(self, vocab_file_path):
        from tokenizers import Tokenizer
        self.tokenizer = Tokenizer.from_file(str(vocab_file_path))
        voc = self.tokenizer.get_vocab()
        assert "[START]" in voc and "[STOP]" in voc

    def encode(self, txt, language_id=None, lowercase=True, nfkd_normalize=True):
        from unicodedata import normalize
        if lowercase:
            txt = txt.lower()
        if nfkd_normalize:
            txt = normalize("NFKD", txt)
        if language_id in self.NEEDS_NORMALISER:
            raise NotImplementedError(f"language '{language_id}' needs the reference's text normaliser (un-vendored "
                                      "dependency); pass the reference MTLTokenizer as `tokenizer`")
        if language_id:
            txt = f"[{language_id.lower()}]{txt}"
        return self.tokenizer.encode(txt.replace(" ", "[SPACE]")).ids

    def text_to_tokens(self, text, language_id=None, lowercase=True, nfkd_normalize=True):
        return torch.IntTensor(self.encode(text, language_id, lowercase, nfkd_normalize)).unsqueeze(0)


class ChatterboxMultilingualTTS(ChatterboxTTS):
    """Drop-in for reference ChatterboxMultilingualTTS (mtl_tts.py:137-355): the same hot path with the multilingual
    T3 checkpoint (text vocabulary 2454, t3_config.py:28-41); language only changes the token ids
    (`[lang]` prefix token, models/tokenizers/tokenizer.py:301-302), plus the tail-trim rule of mtl_tts.py:346-351."""

    @classmethod
    def get_supported_languages(cls):
        return SUPPORTED_LANGUAGES.copy()

    @classmethod
    def from_local(cls, ckpt_dir, device="cuda", t3_model=None):
        """reference mtl_tts.py:176-222: `t3_mtl23ls_v2.safetensors` (or the file named by t3_model), `s3gen.pt`,
        `grapheme_mtl_merged_expanded_v1.json`, `conds.pt`."""
        from safetensors.torch import load_file
        ckpt_dir = Path(ckpt_dir)
        t3_file = t3_model if (t3_model and t3_model.endswith(".safetensors")) else "t3_mtl23ls_v2.safetensors"
        t3_sd = load_file(ckpt_dir / t3_file)
        if "model" in t3_sd:
            t3_sd = t3_sd["model"][0]
        s3 = torch.load(ckpt_dir / "s3gen.pt", map_location="cpu", weights_only=True)
        flow_sd = {k[len("flow."):]: v for k, v in s3.items() if k.startswith("flow.")}
        hift_sd = {k[len("mel2wav."):]: v for k, v in s3.items() if k.startswith("mel2wav.")}
        tok = MTLTokenizer(ckpt_dir / "grapheme_mtl_merged_expanded_v1.json")
        conds = Conditionals.load(ckpt_dir / "conds.pt") if (ckpt_dir / "conds.pt").exists() else None
        return cls.from_state_dicts(t3_sd, flow_sd, hift_sd, conds, tok, device)

    @classmethod
    def from_pretrained(cls, device="cuda", t3_model=None):
        """reference mtl_tts.py:224-247 (HF hub snapshot of ResembleAI/chatterbox; no network in the build sandbox)."""
        from huggingface_hub import snapshot_download
        ckpt = snapshot_download(repo_id="ResembleAI/chatterbox", repo_type="model",
                                 allow_patterns=["ve.pt", "t3_mtl23ls_v2.safetensors", "s3gen.pt",
                                                 "grapheme_mtl_merged_expanded_v1.json", "conds.pt", "Cangjie5_TC.json"])
        return cls.from_local(ckpt, device, t3_model)

    @staticmethod
    def clean_speech_tokens(toks):
        """reference mtl_tts.py:339-341: cut at SOS / EOS only -- unlike the English class there is NO `< 6561` filter, so
        a surviving special id (6563..8193) reaches `flow.input_embedding` (6561 rows) and the reference dies with an
        IndexError there; the same input is rejected here."""
        st = drop_invalid_tokens(toks)
        if (st >= SPEECH_VOCAB_SIZE).any():
            raise IndexError(f"speech token id {int(st.max())} out of range for the flow embedding ({SPEECH_VOCAB_SIZE} rows), "
                             "as in the reference (mtl_tts.py:341 -> flow.py:166)")
        return st

    def _post_wav(self, wav, n_speech_tokens):
        return mtl_tail_trim(wav, n_speech_tokens) if n_speech_tokens > 0 else wav

    def text_to_tokens(self, text, language_id=None):
        assert self.tokenizer is not None, "no multilingual tokenizer loaded; pass token ids to generate_tokens()"
        return self.tokenizer.text_to_tokens(text, language_id=language_id)

    @torch.inference_mode()
    def generate(self, text, language_id, audio_prompt_path=None, exaggeration=0.5, cfg_weight=0.5, temperature=0.8,
                 repetition_penalty=1.2, min_p=0.05, top_p=1.0, max_new_tokens=1000, rng="torch_cpu", kv_dtype=None,
                 watermark=True):
        """reference mtl_tts.py:280-355 (same keyword defaults; the reference always runs the CFG pair, which at
        cfg_weight = 0 equals the single conditional row run here: `cond + 0 * (cond - uncond)`, t3.py:343)."""
        if language_id and language_id.lower() not in SUPPORTED_LANGUAGES:
            raise ValueError(f"Unsupported language_id '{language_id}'. Supported languages: "
                             + ", ".join(SUPPORTED_LANGUAGES.keys()))
        assert audio_prompt_path is None, "prepare_conditionals is outside the hot path; set .conds"
        ids = self.text_to_tokens(mtl_punc_norm(text), language_id=language_id.lower() if language_id else None)
        wav = self.generate_tokens(ids, repetition_penalty=repetition_penalty, min_p=min_p, top_p=top_p,
                                   exaggeration=exaggeration, cfg_weight=cfg_weight, temperature=temperature,
                                   max_new_tokens=max_new_tokens, rng=rng, kv_dtype=kv_dtype)
        return apply_watermark(wav, self.sr, watermark)

    @torch.inference_mode()
    def generate_tokens(self, text_tokens, *args, return_intermediates=False, **kw):
        wav, mid = super().generate_tokens(text_tokens, *args, return_intermediates=True, **kw)
        wav = mtl_tail_trim(wav, mid["speech_tokens"].numel())
        return (wav, mid) if return_intermediates else wav
