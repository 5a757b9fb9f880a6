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
