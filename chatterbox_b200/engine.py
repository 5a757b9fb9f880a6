"""Host side of the B200 engine: owns the cbx handle, packs batches into the layouts libcbx expects,
and drives the stage-level C-ABI calls.  PyTorch is only the container for device memory / streams.

Batch semantics: the reference is batch-1; a batch here equals running the reference once
per utterance (per-utterance noise / RNG streams are explicit inputs).
"""
import ctypes as C
import math

import numpy as np
import torch

from ._lib import Handle, Layout, T3State, HiftGeom, CbxError

TILE = 128
SPEECH_VOCAB = 8194
START_SPEECH, STOP_SPEECH = 6561, 6562
LEN_COND = 34
LDL = 8256          # logits row stride (8194 padded to a multiple of 64)
PAGE_TOKENS = 32


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class PackedLayout:
    """Sequences back to back, each starting at a multiple of 128 rows (cbx_layout)."""

    def __init__(self, lens, device, alloc=None, starts=None):
        lens = np.asarray(lens, dtype=np.int32)
        if starts is None:
            alloc = lens if alloc is None else np.asarray(alloc, dtype=np.int64)
            padded = ((np.asarray(alloc, dtype=np.int64) + TILE - 1) // TILE) * TILE
            padded = np.maximum(padded, TILE)
            starts = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int32)
            rows = int(padded.sum())
        else:
            starts, rows = starts
        self.lens, self.starts, self.rows = lens, np.asarray(starts, dtype=np.int32), int(rows)
        self.n_seq = len(lens)
        self.max_len = int(lens.max()) if len(lens) else 0
        tile_seq = np.full(self.rows // TILE, -1, dtype=np.int32)
        ends = np.concatenate([self.starts[1:], [self.rows]])
        for s in range(self.n_seq):
            tile_seq[self.starts[s] // TILE: ends[s] // TILE] = s
        self.d_tile = torch.from_numpy(tile_seq).to(device)
        self.d_start = torch.from_numpy(self.starts.copy()).to(device)
        self.d_len = torch.from_numpy(self.lens.copy()).to(device)
        self._h_start = np.ascontiguousarray(self.starts)
        self._h_len = np.ascontiguousarray(self.lens)
        self.c = Layout(self.n_seq, self.rows, self.max_len, _ptr(self.d_tile), _ptr(self.d_start), _ptr(self.d_len),
                        C.c_void_p(self._h_start.ctypes.data), C.c_void_p(self._h_len.ctypes.data))

    def scaled(self, factor, lens, device):
        """Layout whose starts / rows are exact multiples of this one (polyphase up-convs write through it)."""
        return PackedLayout(lens, device, starts=(self.starts.astype(np.int64) * factor, self.rows * factor))

    def concat_twice(self, device):
        """CFG layout: the same sequences twice (conditional copies then unconditional copies)."""
        starts = np.concatenate([self.starts, self.starts + self.rows])
        return PackedLayout(np.concatenate([self.lens, self.lens]), device, starts=(starts, 2 * self.rows))


def espnet_pe_table(max_len=5000, d_model=512):
    """The reference's relative positional table (transformer/embedding.py:229-258), built with the same torch
    ops so that the engine reads bit-identical values: row p <-> relative position (max_len-1-p)."""
    position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe_pos = torch.zeros(max_len, d_model)
    pe_neg = torch.zeros(max_len, d_model)
    pe_pos[:, 0::2] = torch.sin(position * div_term)
    pe_pos[:, 1::2] = torch.cos(position * div_term)
    pe_neg[:, 0::2] = torch.sin(-1 * position * div_term)
    pe_neg[:, 1::2] = torch.cos(-1 * position * div_term)
    return torch.cat([torch.flip(pe_pos, [0]), pe_neg[1:]], dim=0).contiguous()


def llama3_rope_tables(n_pos, head_dim=64, base=500000.0, factor=8.0, low=1.0, high=4.0, orig=8192):
    """cos/sin [n_pos, 32] exactly as transformers computes them (modeling_rope_utils._compute_llama3_parameters +
    LlamaRotaryEmbedding.forward) for the reference's rope_scaling (t3/llama_configs.py:23-30)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(torch.float) / head_dim))
    low_w, high_w = orig / low, orig / high
    wavelen = 2 * math.pi / inv_freq
    inv_l = torch.where(wavelen > low_w, inv_freq / factor, inv_freq)
    smooth = (orig / wavelen - low) / (high - low)
    smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
    is_med = ~(wavelen < high_w) * ~(wavelen > low_w)
    inv = torch.where(is_med, smoothed, inv_l)
    pos = torch.arange(n_pos, dtype=torch.float32)
    freqs = (inv[None, :, None].float() @ pos[None, None, :].float()).transpose(1, 2)[0]
    return freqs.cos().contiguous(), freqs.sin().contiguous()


class Engine:
    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise CbxError("chatterbox_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.h = Handle(device)
        self._ws = None
        self.t3_layers = 0
        self.t3_turbo = False       # GPT-2 backbone (reference tts_turbo.py): no CFG, learned absolute positions
        self._decode_stream = torch.cuda.Stream(self.device)   # decode steps replay CUDA graphs: needs a capturable stream
        self._t3_bufs = {}          # persistent T3 state buffers keyed by shape (stable pointers -> decode graphs are reused)
        self._pinned = []           # pinned host scalars for the asynchronous n_act snapshots
        self.decode_steps_per_call = 16
        self.meanflow = False
        # algorithmic-traffic bookkeeping for bench.py's roofline (bytes the paged decode attention must read)
        self.stats = dict(paged_bytes=0.0, paged_launches=0, decode_steps=0, decode_row_steps=0)

    # ------------------------------------------------------------------ weights
    def _load(self, prefix, sd):
        for k, v in sd.items():
            t = v.detach().to(torch.float32).contiguous().cpu()
            shape = (C.c_int64 * t.dim())(*t.shape)
            self.h.call("cbx_load_tensor", (prefix + k).encode(), C.c_void_p(t.data_ptr()), t.dim(), shape)

    def load_t3(self, sd, max_pos=6400):
        """State dict with the reference key names: Llama backbone (`tfmr.layers.*`, t3.py:49-85) or the Turbo GPT-2
        backbone (`tfmr.h.*`, `tfmr.wpe`, tts_turbo.py:151-166; `tfmr.wte` is ignored like the reference deletes it)."""
        sd = {k: v for k, v in sd.items() if not k.startswith("text_head") and not k.startswith("tfmr.embed_tokens")
              and not k.startswith("tfmr.wte")}
        self.t3_turbo = any(k.startswith("tfmr.h.") for k in sd)
        self._load("t3.", sd)
        if self.t3_turbo:
            n_pos = int(sd["tfmr.wpe.weight"].shape[0])
            cos, sin = torch.ones(n_pos, 32), torch.zeros(n_pos, 32)      # GPT-2 has no rotary embedding: identity
            self.t3_layers = len({k.split(".")[2] for k in sd if k.startswith("tfmr.h.")})
        else:
            cos, sin = llama3_rope_tables(max_pos)
            self.t3_layers = len({k.split(".")[2] for k in sd if k.startswith("tfmr.layers.")})
        self._load("t3.", {"rope_cos": cos, "rope_sin": sin})
        self.t3_max_pos = int(cos.shape[0])
        self.t3_speech_pos_rows = int(sd["speech_pos_emb.emb.weight"].shape[0]) if "speech_pos_emb.emb.weight" in sd else 1 << 30
        self.t3_text_pos_rows = int(sd["text_pos_emb.emb.weight"].shape[0]) if "text_pos_emb.emb.weight" in sd else 1 << 30
        self.h.call("cbx_finalize_weights", b"t3")

    def load_flow(self, sd, max_len=5000):
        self._load("flow.", sd)
        self._load("flow.", {"pe_table": espnet_pe_table(max_len)})
        self.h.call("cbx_finalize_weights", b"flow")
        self.meanflow = any("time_embed_mixer" in k for k in sd)

    def load_hift(self, sd):
        self._load("hift.", sd)
        self.h.call("cbx_finalize_weights", b"hift")

    def workspace(self, nbytes):
        if nbytes == 0:
            raise CbxError("workspace query failed: " + self.h.lib.cbx_last_error(self.h.h).decode())
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_attention_precision(self, fmt="fp16"):
        """Operand format of the CFM (flow) attention: 'fp16' = one fp16 plane per operand, one MMA term (default; measured
        mel RMS 1.1e-5 vs the reference at T = 2040 frames, bar 1e-3); 'bf16x3' = bf16 hi/lo planes, three terms (fp32-faithful,
        6.7e-6)."""
        assert fmt in ("bf16x3", "fp16")
        self.h.set_option("attn_prec", fmt)

    def set_cfm_activation_precision(self, fmt="fp16"):
        """Operand format of the CFM transformer-block GEMM inputs (LayerNorm outputs, attention output, GELU output):
        'fp16' = one fp16 plane against an fp16 copy of the weights, one MMA term (default; measured mel RMS 1.5e-4 at
        T = 2040 frames with fp16 attention as well); 'bf16x2' = bf16 hi/lo planes, two terms.  The residual stream stays fp32."""
        assert fmt in ("bf16x2", "fp16")
        self.h.set_option("cfm_act", fmt)

    def set_decode_graph(self, on=True):
        """Default on: every decode step replays a CUDA graph captured once per (state buffers, capacity); stream
        capture needs a non-default stream, so decode runs on a side stream.  Off = direct launches (debugging)."""
        self.h.set_option("decode_graph", "1" if on else "0")

    def _t3_state_buffers(self, key, make):
        """State tensors of a T3 batch; the most recent shape is kept so that a repeated batch shape reuses the same
        device pointers (and with them the cached decode graphs and the KV pool allocation)."""
        if key not in self._t3_bufs:
            self._t3_bufs.clear()            # one shape at a time: the KV pool can be tens of GB
            with torch.inference_mode(False):    # persistent buffers outlive the caller's inference_mode block
                self._t3_bufs[key] = make()
        return self._t3_bufs[key]

    @staticmethod
    def decode_capacity(n, rows_per):
        """Slots launched per decode step for <= n live utterances: exact up to 8 rows (GEMV kernels), then powers of two
        up to one 128-row GEMM tile, then whole tiles."""
        n = max(1, int(n))
        rows = n * rows_per
        if rows <= 8:
            r = 2 if rows <= 2 else 4 if rows <= 4 else 8
            return max(1, r // rows_per)
        per_tile = 128 // rows_per
        if n <= per_tile:
            c = 1
            while c < n:
                c *= 2
            return c
        return ((n + per_tile - 1) // per_tile) * per_tile

    # ------------------------------------------------------------------ T3
    def t3_cond(self, speaker_emb, prompt_tokens, emotion_adv):
        """[n_voices,256], [n_voices,n_prompt] int, [n_voices] -> cond [n_voices, 34, 1024]
        (reference T3.prepare_conditioning, t3.py:92-100)."""
        dev = self.device
        spk = speaker_emb.to(dev, torch.float32).reshape(-1, 256).contiguous()
        nv = spk.shape[0]
        ptok = prompt_tokens.to(dev, torch.int32).reshape(nv, -1).contiguous()
        emo = emotion_adv.to(dev, torch.float32).reshape(nv).contiguous()
        len_cond = 1 + ptok.shape[1] if self.t3_turbo else LEN_COND      # Turbo: [spkr | prompt embeddings]
        out = torch.empty(nv, len_cond, 1024, device=dev, dtype=torch.float32)
        ws = self.workspace(self.h.lib.cbx_t3_workspace_bytes(self.h.h, 256, 2))
        self.h.call("cbx_t3_cond_encode", _ptr(spk), _ptr(ptok), ptok.shape[1], _ptr(emo), nv, _ptr(out), _ptr(ws),
                    ws.numel(), self._stream())
        return out

    def t3_generate(self, text_tokens, cond, voice_ids=None, max_new_tokens=1000, cfg_weight=0.5, temperature=0.8,
                    top_p=1.0, min_p=0.05, repetition_penalty=1.2, q_noise=None, seed=0, kv_dtype="bf16",
                    max_sync_steps=None, return_state=False, top_k=0, force_tokens=None, act_dtype=None):
        """Batched equivalent of T3.inference (t3.py:225-390) or, with a Turbo checkpoint loaded, of
        T3.inference_turbo (t3.py:392-468: no CFG, processors temperature -> top_k -> top_p -> repetition penalty,
        min_p unused; max_new_tokens counts the token sampled from the prefill, i.e. max_gen_len + 1).
        text_tokens: list of 1-D int tensors incl. SOT/EOT (Turbo: raw tokenizer ids), one per utterance.
        cond: [n_voices, len_cond, 1024] from t3_cond.  max_new_tokens: int or per-utterance list.
        Returns a list of 1-D int64 CPU tensors (EOS included if hit)."""
        dev = self.device
        B = len(text_tokens)
        turbo = self.t3_turbo
        len_cond = int(cond.shape[1])
        cfg = 1 if (cfg_weight > 0.0 and not turbo) else 0
        rp = 2 if cfg else 1
        R = B * rp
        voice_ids = [0] * B if voice_ids is None else list(voice_ids)
        max_new = [int(max_new_tokens)] * B if np.isscalar(max_new_tokens) else [int(m) for m in max_new_tokens]
        n_text = np.array([len(t) for t in text_tokens], dtype=np.int32)
        s0 = len_cond + n_text + (1 if turbo else 2)       # [cond | text | BOS (| BOS)]  (t3.py:126-129, 305-313, 407-413)
        row_len = np.repeat(s0, rp).astype(np.int32)
        row_start = np.concatenate([[0], np.cumsum(row_len)[:-1]]).astype(np.int32)
        n_tok = int(row_len.sum())
        tok_row = np.repeat(np.arange(R, dtype=np.int32), row_len)
        tok_pos = np.concatenate([np.arange(l, dtype=np.int32) for l in row_len])
        text_flat = np.concatenate([np.asarray(t, dtype=np.int32).reshape(-1) for t in text_tokens])
        utt_text_start = np.concatenate([[0], np.cumsum(n_text)[:-1]]).astype(np.int32)
        row_text_start = np.repeat(utt_text_start, rp)
        row_ntext = np.repeat(n_text, rp)
        row_voice = np.repeat(np.asarray(voice_ids, dtype=np.int32), rp)
        row_uncond = np.tile(np.array([0, 1], dtype=np.int32), B) if cfg else np.zeros(R, dtype=np.int32)
        # paged KV cache: just enough pages per row for prefill + budget
        pages_per_row = (np.repeat(s0 + np.asarray(max_new, dtype=np.int32), rp) + PAGE_TOKENS - 1) // PAGE_TOKENS
        max_pages = int(pages_per_row.max())
        first_page = np.concatenate([[0], np.cumsum(pages_per_row)[:-1]]).astype(np.int64)
        page_table = first_page[:, None] + np.arange(max_pages, dtype=np.int64)[None, :]
        page_table = np.where(np.arange(max_pages)[None, :] < pages_per_row[:, None], page_table, 0).astype(np.int32)
        n_pages = int(pages_per_row.sum())
        kvt = (torch.float32 if kv_dtype in ("fp32", "f32", torch.float32) else
               torch.float8_e4m3fn if kv_dtype in ("fp8", "e4m3", torch.float8_e4m3fn) else torch.bfloat16)
        kv_code = {torch.bfloat16: 0, torch.float32: 1, torch.float8_e4m3fn: 2}[kvt]
        # decode-step operand format: fp32-faithful bf16 hi/lo planes with the fp32 (parity) cache, one fp16 plane in the
        # throughput configurations (bf16 / fp8 cache) unless asked otherwise
        if act_dtype is None:
            act_dtype = "bf16x2" if kvt == torch.float32 else "fp16"
        assert act_dtype in ("bf16x2", "fp16")
        L = self.t3_layers
        max_tokens = int(max(max_new))
        # position tables of the checkpoint bound what may be generated (reference: learned tables of 2050 / 4100 rows,
        # RoPE table built for max_pos rows)
        if int(row_len.max()) + max_tokens > self.t3_max_pos:
            raise CbxError(f"prefill length {int(row_len.max())} + max_new_tokens {max_tokens} exceeds the {self.t3_max_pos} "
                           "positions of the T3 position tables")
        if not turbo and (max_tokens + 1 > self.t3_speech_pos_rows or int(n_text.max()) > self.t3_text_pos_rows):
            raise CbxError(f"max_new_tokens {max_tokens} / text length {int(n_text.max())} exceed the learned position tables "
                           f"({self.t3_speech_pos_rows} speech, {self.t3_text_pos_rows} text rows)")
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

        def make():
            return dict(
                kv=torch.empty(L, n_pages, 2, 16, PAGE_TOKENS, 64, dtype=kvt, device=dev),      # layer-major (include/cbx.h)
                page_table=torch.empty(R, max_pages, dtype=torch.int32, device=dev),
                positions=torch.empty(R, dtype=torch.int32, device=dev), base_pos=torch.empty(R, dtype=torch.int32, device=dev),
                tokens=torch.empty(B, max_tokens, dtype=torch.int32, device=dev),
                n_gen=torch.empty(B, dtype=torch.int32, device=dev), max_new=torch.empty(B, dtype=torch.int32, device=dev),
                done=torch.empty(B, dtype=torch.int32, device=dev),
                seen=torch.empty(B, SPEECH_VOCAB, dtype=torch.uint8, device=dev),
                x=torch.empty(R, 1024, dtype=torch.float32, device=dev),
                logits=torch.empty(R, LDL, dtype=torch.float32, device=dev),
                act_utt=torch.empty(B, dtype=torch.int32, device=dev), n_act=torch.empty(1, dtype=torch.int32, device=dev),
                src_slot=torch.empty(B, dtype=torch.int32, device=dev), slot_row=torch.empty(R, dtype=torch.int32, device=dev),
                m_live=torch.empty(1, dtype=torch.int32, device=dev),
                force=torch.empty(B, max_tokens, dtype=torch.int32, device=dev),
                sampled=torch.empty(B, max_tokens, dtype=torch.int32, device=dev))

        st_t = self._t3_state_buffers((B, R, max_tokens, n_pages, max_pages, kvt, L), make)
        kv = st_t["kv"]
        st_t["page_table"].copy_(t(page_table))
        st_t["base_pos"].copy_(t(row_len))
        st_t["max_new"].copy_(t(np.asarray(max_new, dtype=np.int32)))
        for k in ("positions", "tokens", "n_gen", "done", "seen", "x", "logits", "src_slot", "sampled"):
            st_t[k].zero_()
        st_t["seen"][:, START_SPEECH] = 1          # repetition penalty history starts with BOS (t3.py:316,347)
        st_t["act_utt"].copy_(torch.arange(B, dtype=torch.int32, device=dev))
        st_t["slot_row"].copy_(torch.arange(R, dtype=torch.int32, device=dev))
        st_t["n_act"].fill_(B)
        st_t["m_live"].fill_(R)
        forced = None
        if force_tokens is not None:               # teacher forcing (parity tests): [B][<= max_tokens] ids
            st_t["force"].zero_()
            for b in range(B):
                ft = torch.as_tensor(force_tokens[b]).reshape(-1).to(torch.int32)
                st_t["force"][b, :ft.numel()] = ft.to(dev)
            forced = st_t["force"]
        d = dict(tok_row=t(tok_row), tok_pos=t(tok_pos), row_start=t(row_start), row_len=t(row_len),
                 text_flat=t(text_flat), row_text_start=t(row_text_start), row_ntext=t(row_ntext),
                 row_voice=t(row_voice), row_uncond=t(row_uncond))
        qn = q_noise.to(dev, torch.float32).contiguous() if q_noise is not None else None
        st = T3State(B, R, cfg, _ptr(kv), kv_code, PAGE_TOKENS, _ptr(st_t["page_table"]), max_pages, n_pages,
                     _ptr(st_t["positions"]), _ptr(st_t["base_pos"]), _ptr(st_t["tokens"]), max_tokens,
                     _ptr(st_t["n_gen"]), _ptr(st_t["max_new"]), _ptr(st_t["done"]), _ptr(st_t["seen"]),
                     _ptr(st_t["x"]), _ptr(st_t["logits"]), LDL, float(cfg_weight), float(repetition_penalty),
                     float(temperature), float(min_p), float(top_p), _ptr(qn), int(seed), 1 if turbo else 0, int(top_k),
                     _ptr(st_t["act_utt"]), _ptr(st_t["n_act"]), _ptr(st_t["src_slot"]), _ptr(st_t["slot_row"]),
                     _ptr(st_t["m_live"]), _ptr(forced), _ptr(st_t["sampled"]) if forced is not None else C.c_void_p(0),
                     1 if act_dtype == "fp16" else 0)
        ws = self.workspace(self.h.lib.cbx_t3_workspace_bytes(self.h.h, n_tok, R))
        cond = cond.to(dev, torch.float32).contiguous()
        self.h.call("cbx_t3_prefill", C.byref(st), n_tok, _ptr(d["tok_row"]), _ptr(d["tok_pos"]), _ptr(d["row_start"]),
                    _ptr(d["row_len"]), int(row_len.max()), _ptr(cond), _ptr(d["row_voice"]), len_cond,
                    _ptr(d["text_flat"]), _ptr(d["row_text_start"]), _ptr(d["row_ntext"]), _ptr(d["row_uncond"]),
                    _ptr(ws), ws.numel(), self._stream())
        if return_state == "prefill":
            return st_t
        # decode: graph replays on the side stream, finished utterances retire on the device
        self._decode_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._decode_stream):
            self._decode_loop(st, st_t, B, rp, max_new, ws)
        torch.cuda.current_stream(self.device).wait_stream(self._decode_stream)
        toks = st_t["tokens"].cpu()
        n_gen = st_t["n_gen"].cpu()
        out = [toks[b, :int(n_gen[b])].to(torch.int64) for b in range(B)]
        # algorithmic traffic of the paged decode attention (bench.py roofline): step i of an utterance reads its
        # s0 + i + 1 cached tokens, K and V, all heads, every layer, both CFG rows
        elt = {torch.bfloat16: 2, torch.float32: 4, torch.float8_e4m3fn: 1}[kvt]
        ng = n_gen.numpy().astype(np.int64)
        ctx_tok = (s0.astype(np.int64) * ng + ng * (ng + 1) // 2).sum()
        self.stats["paged_bytes"] += float(ctx_tok) * rp * 2 * 1024 * elt * L
        self.stats["decode_row_steps"] += int(ng.sum()) * rp
        if return_state:
            return out, st_t
        return out

    def _decode_loop(self, st, st_t, B, rp, max_new, ws):
        """Host side of the decode loop: no device->host read on the critical path.  The device retires finished utterances
        every step (t3_compact_kernel); the host only picks the launch capacity from what it knows for sure -- the budgets --
        and from asynchronous snapshots of the device's live count (pinned memory + event) taken two calls earlier."""
        budget = np.asarray(max_new, dtype=np.int64)
        total = int(budget.max())
        steps_done, live_seen = 0, B
        pending = []
        lag = 2        # snapshots are consumed with a FIXED lag of `lag` calls (their copies finished long ago: no stall), never
                       # by polling: the capacity sequence -- and with it tile / split choices and the rounding of every logit --
                       # is then a function of the device state alone, not of host timing (run-to-run reproducible ids)
        while steps_done < total:
            while len(pending) > lag:
                ev, buf = pending.pop(0)
                ev.synchronize()
                live_seen = min(live_seen, int(buf[0]))
                self._pinned.append((ev, buf))
            live = min(live_seen, int((budget > steps_done).sum()))
            if live <= 0:
                break
            cap = self.decode_capacity(live, rp)
            k = min(self.decode_steps_per_call, total - steps_done)
            self.h.call("cbx_t3_decode", C.byref(st), min(cap, B), k, _ptr(ws), ws.numel(), self._stream())
            steps_done += k
            self.stats["decode_steps"] += k
            self.stats["paged_launches"] += k * self.t3_layers
            if self._pinned:
                ev, buf = self._pinned.pop()
            else:
                with torch.inference_mode(False):
                    ev, buf = torch.cuda.Event(), torch.zeros(1, dtype=torch.int32).pin_memory()
            buf.copy_(st_t["n_act"], non_blocking=True)
            ev.record()
            pending.append((ev, buf))
        for ev, buf in pending:
            self._pinned.append((ev, buf))

    # ------------------------------------------------------------------ flow
    def flow_mel(self, tokens, ref_dicts, z=None, n_timesteps=None, cfg_rate=0.7, return_mu=False, finalize=True):
        """Batched equivalent of S3Token2Wav.flow_inference (s3gen.py:301-321 -> flow.py:131-198).
        tokens: list of 1-D int tensors; ref_dicts: one dict (shared voice) or a list of dicts with
        prompt_token [1,Np], prompt_feat [1,2Np,80], embedding [1,192].  z: optional list of [80, 2(Np+N)] noise
        tensors (what flow_matching.py:216 would draw).  Returns a list of mel tensors [80, 2N] on the device."""
        dev = self.device
        B = len(tokens)
        refs = ref_dicts if isinstance(ref_dicts, (list, tuple)) else [ref_dicts] * B
        n_timesteps = n_timesteps or (2 if self.meanflow else 10)
        np_len = np.array([int(r["prompt_token"].shape[-1]) for r in refs], dtype=np.int32)
        n_gen = np.array([int(t.numel()) for t in tokens], dtype=np.int32)
        n = np_len + n_gen
        L1 = PackedLayout(n, dev)
        L2 = PackedLayout(2 * n, dev, alloc=2 * n + 4)     # >= 4 zero rows after every sequence: the halo of the plane-fed causal convs
        # streaming chunk (finalize=False, reference flow.py:170-171): the encoder sees every token, the decoder drops the
        # last pre_lookahead_len * token_mel_ratio = 6 frames (they still depend on tokens that have not arrived yet)
        cut = 0 if finalize else 6
        Ld = L2 if finalize else PackedLayout(2 * n - cut, dev, starts=(L2.starts, L2.rows))
        L3 = Ld if self.meanflow else Ld.concat_twice(dev)
        tok = torch.zeros(L1.rows, dtype=torch.int32)
        cond = torch.zeros(L2.rows, 80, dtype=torch.float32)
        xvec = torch.zeros(B, 192, dtype=torch.float32)
        for b in range(B):
            r = refs[b]
            s = int(L1.starts[b])
            tok[s:s + np_len[b]] = r["prompt_token"].reshape(-1).to(torch.int32).cpu()
            tok[s + np_len[b]:s + n[b]] = tokens[b].reshape(-1).to(torch.int32).cpu()
            pf = r["prompt_feat"].reshape(-1, 80).to(torch.float32).cpu()
            s2 = int(L2.starts[b])
            cond[s2:s2 + pf.shape[0]] = pf                     # flow.py:178-180
            xvec[b] = r["embedding"].reshape(-1).to(torch.float32).cpu()
        tok, cond, xvec = tok.to(dev), cond.to(dev), xvec.to(dev)
        mu = torch.zeros(L2.rows, 80, dtype=torch.float32, device=dev)
        spk = torch.zeros(B, 80, dtype=torch.float32, device=dev)
        x = torch.zeros(L2.rows, 80, dtype=torch.float32, device=dev)
        for b in range(B):
            s2, T = int(L2.starts[b]), int(2 * n[b]) - cut
            if z is not None:
                x[s2:s2 + T] = z[b].reshape(80, T).t().to(dev, torch.float32)
            else:
                x[s2:s2 + T] = torch.randn(80, T, device=dev).t()
        ws = self.workspace(self.h.lib.cbx_flow_workspace_bytes(self.h.h, C.byref(L1.c), C.byref(L2.c), C.byref(L3.c)))
        self.h.call("cbx_flow_encode", _ptr(tok), C.byref(L1.c), C.byref(L2.c), _ptr(xvec), _ptr(mu), _ptr(spk),
                    _ptr(ws), ws.numel(), self._stream())
        if return_mu:
            return [mu[int(L2.starts[b]):int(L2.starts[b]) + int(2 * n[b])].clone() for b in range(B)], spk
        self.h.call("cbx_cfm_solve", _ptr(mu), _ptr(spk), _ptr(cond), _ptr(x), C.byref(Ld.c), C.byref(L3.c),
                    int(n_timesteps), float(cfg_rate), 1 if self.meanflow else 0, _ptr(ws), ws.numel(), self._stream())
        out = []
        for b in range(B):
            s2 = int(L2.starts[b])
            out.append(x[s2 + 2 * np_len[b]:s2 + 2 * n[b] - cut].t().contiguous())      # drop prompt frames (flow.py:196)
        return out

    # ------------------------------------------------------------------ HiFT
    def _hift_geom(self, T):
        dev = self.device
        T = np.asarray(T, dtype=np.int32)
        # +4 frames: room for the 120T+1-th row of the last stage, and a zero gap of >= 32 rows at the 8T level between two
        # sequences -- the staged-tile ResBlock convolutions (hift_conv.cu) read their dilated halo (up to 25 rows) from it
        LT = PackedLayout(T, dev, alloc=T + 4)
        L8 = LT.scaled(8, 8 * T, dev)
        L40 = LT.scaled(40, 40 * T, dev)
        L120 = LT.scaled(120, 120 * T + 1, dev)
        sstart = torch.from_numpy(LT.starts.astype(np.int64) * 480).to(dev)
        g = HiftGeom(LT.c, L8.c, L40.c, L120.c, _ptr(sstart), int(LT.rows) * 480)
        keep = (LT, L8, L40, L120, sstart)
        return g, keep

    def hift(self, mels, source=None, phase_vec=None, noise=None, seed=0, trim_fade=True, f0=None):
        """Batched equivalent of S3Token2Wav.hift_inference + trim-fade (s3gen.py:324-327,359-360).
        mels: list of [80, T] tensors.  source: optional list of [1, 480T] (reference cache_source hook);
        phase_vec: optional list of [9]; noise: optional list of [9, 480T] (SineGen draws, hifigan.py:212-226).
        Returns (list of wav [480T], list of source [480T]) on the device."""
        dev = self.device
        B = len(mels)
        T = np.array([int(m.shape[-1]) for m in mels], dtype=np.int32)
        g, keep = self._hift_geom(T)
        LT = keep[0]
        mel = torch.zeros(LT.rows, 80, dtype=torch.float32, device=dev)
        for b in range(B):
            mel[int(LT.starts[b]):int(LT.starts[b]) + int(T[b])] = mels[b].reshape(80, -1).t().to(dev, torch.float32)
        total = int(LT.rows) * 480
        s = torch.zeros(total, dtype=torch.float32, device=dev)
        wav = torch.zeros(total, dtype=torch.float32, device=dev)
        ws = self.workspace(self.h.lib.cbx_hift_workspace_bytes(self.h.h, C.byref(g)))
        full_cache = source is not None and all(c is not None and c.numel() >= 480 * int(t) for c, t in zip(source, T))
        if not full_cache:
            pv = None
            if phase_vec is not None:
                pv = torch.stack([p.reshape(9).to(torch.float32) for p in phase_vec]).to(dev).contiguous()
            nz = None
            if noise is not None:
                nz = torch.zeros(total * 9, dtype=torch.float32, device=dev)
                for b in range(B):
                    o = int(LT.starts[b]) * 480 * 9
                    nz[o:o + 9 * 480 * int(T[b])] = noise[b].reshape(-1).to(dev, torch.float32)
            f0d = None
            if f0 is not None:         # inject the reference's f0 (unit-test hook)
                f0d = torch.zeros(LT.rows, dtype=torch.float32, device=dev)
                for b in range(B):
                    f0d[int(LT.starts[b]):int(LT.starts[b]) + int(T[b])] = f0[b].reshape(-1).to(dev, torch.float32)
            self.h.call("cbx_hift_source", _ptr(mel), C.byref(g), _ptr(pv), _ptr(nz), int(seed), _ptr(s), _ptr(f0d),
                        C.c_void_p(0), _ptr(ws), ws.numel(), self._stream())
        if source is not None:        # cache_source (hifigan.py:470-472): overwrites the head of the source, full or partial
            for b in range(B):
                if source[b] is None:
                    continue
                o = int(LT.starts[b]) * 480
                cs = source[b].reshape(-1).to(dev, torch.float32)
                m = min(int(cs.numel()), 480 * int(T[b]))
                s[o:o + m] = cs[:m]
        self.h.call("cbx_hift_decode", _ptr(mel), _ptr(s), C.byref(g), _ptr(wav), 1 if trim_fade else 0, _ptr(ws),
                    ws.numel(), self._stream())
        wavs, srcs = [], []
        for b in range(B):
            o = int(LT.starts[b]) * 480
            wavs.append(wav[o:o + 480 * int(T[b])])
            srcs.append(s[o:o + 480 * int(T[b])])
        return wavs, srcs

    def hift_f0(self, mels):
        """F0 predictor only (unit-test hook; f0_predictor.py:52-55)."""
        dev = self.device
        T = np.array([int(m.shape[-1]) for m in mels], dtype=np.int32)
        g, keep = self._hift_geom(T)
        LT = keep[0]
        mel = torch.zeros(LT.rows, 80, dtype=torch.float32, device=dev)
        for b in range(len(mels)):
            mel[int(LT.starts[b]):int(LT.starts[b]) + int(T[b])] = mels[b].reshape(80, -1).t().to(dev, torch.float32)
        s = torch.zeros(int(LT.rows) * 480, dtype=torch.float32, device=dev)
        f0 = torch.zeros(LT.rows, dtype=torch.float32, device=dev)
        ws = self.workspace(self.h.lib.cbx_hift_workspace_bytes(self.h.h, C.byref(g)))
        self.h.call("cbx_hift_source", _ptr(mel), C.byref(g), C.c_void_p(0), C.c_void_p(0), 0, _ptr(s), C.c_void_p(0),
                    _ptr(f0), _ptr(ws), ws.numel(), self._stream())
        return [f0[int(LT.starts[b]):int(LT.starts[b]) + int(T[b])] for b in range(len(mels))]
