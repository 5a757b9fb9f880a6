"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/cbx.h declares,
the product path fails loudly without a GPU, and the host-side batching logic is sound."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cbx.h")).read()
    return sorted(set(re.findall(r"\b(cbx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "chatterbox_b200", "libcbx.so")):
        ge.build()
    from chatterbox_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cbx.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature"
    assert lib.cbx_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from chatterbox_b200 import Engine, CbxError
    with pytest.raises(CbxError):
        Engine(0)


def test_product_path_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "chatterbox_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle is only", ""), f"{f} references oracle/"


def test_packed_layout_invariants():
    from chatterbox_b200.engine import PackedLayout, TILE
    lens = [1, 128, 129, 700]
    L = PackedLayout(lens, torch.device("cpu"))
    assert L.rows % TILE == 0 and (L.starts % TILE == 0).all()
    for s, n in enumerate(lens):
        nxt = L.starts[s + 1] if s + 1 < len(lens) else L.rows
        assert nxt - L.starts[s] >= n
        assert (L.d_tile[L.starts[s] // TILE: nxt // TILE] == s).all()
    L8 = L.scaled(8, [8 * n for n in lens], torch.device("cpu"))
    assert (L8.starts == 8 * L.starts).all() and L8.rows == 8 * L.rows
    L3 = L.concat_twice(torch.device("cpu"))
    assert L3.n_seq == 8 and L3.rows == 2 * L.rows and (L3.starts[4:] == L.starts + L.rows).all()


def test_tables_match_oracle():
    from chatterbox_b200.engine import llama3_rope_tables, espnet_pe_table
    from oracle.t3_ref import rope_tables
    from oracle.flow_ref import espnet_rel_pos_emb
    c, s = llama3_rope_tables(300)
    co, so = rope_tables(300)
    assert torch.equal(c, co) and torch.equal(s, so)
    pe = espnet_pe_table(5000)
    T = 37
    assert torch.equal(pe[4999 - T + 1: 4999 + T], espnet_rel_pos_emb(T)[0])


def test_punc_norm_and_token_cleanup():
    from chatterbox_b200.tts import punc_norm, drop_invalid_tokens
    assert punc_norm("hello world") == "Hello world."
    assert punc_norm("") == "You need to add some text for me to talk."
    x = torch.tensor([6561, 5, 7, 6562, 9])
    assert drop_invalid_tokens(x).tolist() == [5, 7]


def test_ctypes_structs_match_the_c_header(tmp_path):
    """include/cbx.h is plain C: compile a probe with gcc and compare sizes / field offsets with the ctypes mirrors."""
    import ctypes as C
    import subprocess
    from chatterbox_b200._lib import T3State, Layout, HiftGeom
    fields = ["n_utts", "kv_pages", "page_table", "n_pages", "positions", "tokens", "max_new", "seen", "logits", "ldl",
              "cfg_weight", "top_p", "q_noise", "seed", "sampler", "top_k", "act_utt", "n_act", "src_slot", "slot_row", "m_live",
              "force_tokens", "sampled_out", "act_fp16"]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "cbx.h"', 'int main(void) {',
           'printf("%zu %zu %zu\\n", sizeof(cbx_t3_state), sizeof(cbx_layout), sizeof(cbx_hift_geom));']
    src += [f'printf("%zu\\n", offsetof(cbx_t3_state, {f}));' for f in fields]
    src += ['return 0; }']
    c = tmp_path / "probe.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert [int(x) for x in out[:3]] == [C.sizeof(T3State), C.sizeof(Layout), C.sizeof(HiftGeom)]
    assert [int(x) for x in out[3:]] == [getattr(T3State, f).offset for f in fields]


def test_turbo_shim_surface():
    """Turbo boundary mirrors the reference's (tts_turbo.py:104-321, t3.py:392-394) without needing a GPU."""
    import inspect
    from chatterbox_b200 import ChatterboxTurboTTS, T3
    sig = inspect.signature(T3.inference_turbo)
    for name, default in [("temperature", 0.8), ("top_k", 1000), ("top_p", 0.95), ("repetition_penalty", 1.2),
                          ("max_gen_len", 1000)]:
        assert sig.parameters[name].default == default
    gsig = inspect.signature(ChatterboxTurboTTS.generate)
    for name, default in [("repetition_penalty", 1.2), ("min_p", 0.0), ("top_p", 0.95), ("exaggeration", 0.0),
                          ("cfg_weight", 0.0), ("temperature", 0.8), ("top_k", 1000), ("norm_loudness", True)]:
        assert gsig.parameters[name].default == default


def test_multilingual_shim_rules():
    """mtl_tts.py:293-298 (language validation) and :346-351 (tail trim) without a GPU."""
    from chatterbox_b200.tts import ChatterboxMultilingualTTS, mtl_tail_trim, SUPPORTED_LANGUAGES
    assert len(SUPPORTED_LANGUAGES) == 23 and "sw" in ChatterboxMultilingualTTS.get_supported_languages()
    wav = torch.arange(5 * 960, dtype=torch.float32)[None]
    assert mtl_tail_trim(wav, 5).shape == (1, 4 * 960)
    assert mtl_tail_trim(wav[:, :960], 1).shape == (1, 960)          # keeps at least one token of audio
    tts = ChatterboxMultilingualTTS.__new__(ChatterboxMultilingualTTS)
    with pytest.raises(ValueError):
        tts.generate("hola", language_id="xx")


def test_synthesize_batch_chunking_and_order():
    """Host logic of the batched flow/vocoder stages (no GPU): chunks respect the frame budgets, every utterance comes
    back in its own slot, empty utterances yield empty waveforms."""
    from chatterbox_b200.tts import synthesize_batch

    class FakeEngine:
        device = torch.device("cpu")

        def __init__(self):
            self.flow_calls, self.hift_calls = [], []

        def flow_mel(self, speech, refs, n_timesteps=None):
            self.flow_calls.append([int(s.numel()) for s in speech])
            return [torch.full((80, 2 * int(s.numel())), float(s.numel())) for s in speech]

        def hift(self, mels, seed=0, trim_fade=True):
            self.hift_calls.append([int(m.shape[-1]) for m in mels])
            return [torch.full((480 * int(m.shape[-1]),), float(m[0, 0])) for m in mels], None

    eng = FakeEngine()
    lens = [5, 0, 40, 17, 33, 8]
    speech = [torch.zeros(n, dtype=torch.long) for n in lens]
    refs = [dict(prompt_token=torch.zeros(1, 10, dtype=torch.long))] * len(lens)
    wavs = synthesize_batch(eng, speech, refs, flow_frames_per_chunk=150, hift_frames_per_chunk=100)
    for n, w in zip(lens, wavs):
        assert w.numel() == 960 * n
        assert n == 0 or float(w[0]) == float(n)                 # slot b holds utterance b
    for call in eng.flow_calls:                                   # 2 * (prompt + tokens) frames per utterance
        assert len(call) == 1 or sum(2 * (10 + n) for n in call) <= 150
    assert sorted(sum(eng.flow_calls, [])) == sorted(lens)
    for call in eng.hift_calls:
        assert len(call) == 1 or sum(call) <= 100
    assert sorted(sum(eng.hift_calls, [])) == sorted(2 * n for n in lens if n > 0)


def test_decode_capacity_buckets():
    """Launch capacity of a decode step: never below the live count, exact GEMV shapes up to 8 rows, then powers of two up
    to one 128-row tile, then whole tiles (few distinct values -> few cached CUDA graphs)."""
    from chatterbox_b200.engine import Engine
    for rp in (1, 2):
        seen = set()
        for n in range(1, 600):
            c = Engine.decode_capacity(n, rp)
            assert c >= n or c * rp <= 8 and c >= min(n, 8 // rp), (n, rp, c)
            seen.add(c)
            if n * rp > 128:
                assert (c * rp) % 128 == 0 and c * rp - n * rp < 128
        assert len(seen) <= 16
    assert [Engine.decode_capacity(n, 2) for n in (1, 2, 3, 4, 5, 33, 64, 65, 256)] == [1, 2, 4, 4, 8, 64, 64, 128, 256]


def test_multilingual_text_front_end(tmp_path):
    """mtl_tts.py:70-107 sentence enders, models/tokenizers/tokenizer.py:256-312 language prefix, mtl_tts.py:339-341 token
    clean-up (no `< 6561` filter: an out-of-range id is an error like in the reference), watermark failure mode."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from chatterbox_b200.tts import (mtl_punc_norm, punc_norm, MTLTokenizer, ChatterboxMultilingualTTS, ChatterboxTTS,
                                     apply_watermark)
    assert mtl_punc_norm("你好。") == "你好。" and punc_norm("你好。") == "你好。."
    assert mtl_punc_norm("hello") == "Hello." and mtl_punc_norm("quoi？") == "Quoi？"
    vocab = {"[START]": 0, "[STOP]": 1, "[UNK]": 2, "[SPACE]": 3, "[fr]": 4, "a": 5, "b": 6, "e": 7, "́": 8}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="[UNK]"))
    from tokenizers import Regex
    tok.pre_tokenizer = pre_tokenizers.Split(Regex(r"\[[A-Za-z]+\]|."), behavior="isolated")
    f = tmp_path / "grapheme_mtl_merged_expanded_v1.json"
    tok.save(str(f))
    mt = MTLTokenizer(f)
    assert mt.text_to_tokens("A bÉ", language_id="fr").tolist() == [[4, 5, 3, 6, 7, 8]]      # lower-case, NFKD, prefix, [SPACE]
    with pytest.raises(NotImplementedError):
        mt.encode("x", language_id="zh")
    x = torch.tensor([6561, 5, 6563, 7, 6562, 9])
    assert ChatterboxTTS.clean_speech_tokens(x).tolist() == [5, 7]
    with pytest.raises(IndexError):
        ChatterboxMultilingualTTS.clean_speech_tokens(x)
    assert ChatterboxMultilingualTTS.clean_speech_tokens(torch.tensor([5, 7, 6562, 6563])).tolist() == [5, 7]
    wav = torch.zeros(1, 100)
    assert apply_watermark(wav, 24000, enabled=False) is wav
    try:
        import perth  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError):
            apply_watermark(wav, 24000)
