"""S3Gen module boundary (reference src/chatterbox/models/s3gen/s3gen.py:232-362)."""
import torch

from .engine import Engine

S3GEN_SR = 24000
SPEECH_VOCAB_SIZE = 6561


class S3Gen:
    """token -> mel (CFM) -> waveform (HiFT).  `embed_ref` (voice prompt analysis) is out of scope of the hot path
    (SURVEY.md 2 rows 5-8): pass `ref_dict` as produced by the reference's embed_ref / stored in conds.pt."""

    def __init__(self, engine: Engine, flow_state_dict=None, hift_state_dict=None, meanflow=False):
        self.engine = engine
        self.meanflow = meanflow
        if flow_state_dict is not None:
            engine.load_flow(flow_state_dict)
        if hift_state_dict is not None:
            engine.load_hift(hift_state_dict)

    def embed_ref(self, ref_wav, ref_sr, device="auto", ref_fade_out=True):
        raise NotImplementedError("embed_ref runs once per voice and is not part of the B200 hot path; "
                                  "use a precomputed ref_dict (conds.pt / Conditionals.gen)")

    @torch.inference_mode()
    def flow_inference(self, speech_tokens, ref_wav=None, ref_sr=None, ref_dict=None, n_cfm_timesteps=None,
                       finalize=True, speech_token_lens=None, z=None):
        """reference s3gen.py:301-321 -> mel (1, 80, 2N).  `z` optionally injects the CFM noise [80, 2(Np+N)]."""
        assert ref_dict is not None, "ref_dict required (embed_ref is out of scope)"
        toks = torch.atleast_2d(speech_tokens)[0]
        # finalize=False (streaming chunk, flow.py:170-171): the mel of the last 3 tokens (6 frames) is withheld
        mel = self.engine.flow_mel([toks], ref_dict, z=None if z is None else [z], n_timesteps=n_cfm_timesteps,
                                   finalize=finalize)[0]
        return mel[None]

    @torch.inference_mode()
    def hift_inference(self, speech_feat, cache_source=None, phase_vec=None, noise=None, seed=0, trim_fade=False, f0=None):
        """reference s3gen.py:324-327 -> (wav (1, 480T), source (1, 1, 480T))."""
        src = None
        if cache_source is not None and cache_source.numel() > 0:      # full or partial head of the source (hifigan.py:470-472)
            assert cache_source.shape[-1] <= 480 * speech_feat.shape[-1]
            src = [cache_source]
        wavs, srcs = self.engine.hift([speech_feat[0]], source=src,
                                      phase_vec=None if phase_vec is None else [phase_vec],
                                      noise=None if noise is None else [noise], seed=seed, trim_fade=trim_fade,
                                      f0=None if f0 is None else [f0])
        return wavs[0][None], srcs[0][None, None]

    @torch.inference_mode()
    def inference(self, speech_tokens, ref_wav=None, ref_sr=None, ref_dict=None, drop_invalid_tokens=True,
                  n_cfm_timesteps=None, speech_token_lens=None, z=None, phase_vec=None, noise=None, seed=0):
        """reference s3gen.py:330-362 -> (wav (1, 960N) with trim-fade, source)."""
        mel = self.flow_inference(speech_tokens, ref_dict=ref_dict, n_cfm_timesteps=n_cfm_timesteps, finalize=True, z=z)
        return self.hift_inference(mel, None, phase_vec=phase_vec, noise=noise, seed=seed, trim_fade=True)
