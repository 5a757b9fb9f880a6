"""CPU checks of bench.py's host-side pieces: the synthetic workload is the one SURVEY.md 8d prescribes and is reproducible,
the roofline work model follows BASELINE.md 4, every tool script at least compiles."""
import glob
import os
import py_compile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_workload_is_the_survey_config_and_reproducible():
    b = _bench()
    t1, n1 = b.make_workload(256, b.SEED, 0)
    t2, n2 = b.make_workload(256, b.SEED, 0)
    assert n1 == n2 and all(torch.equal(a, c) for a, c in zip(t1, t2))
    assert len(t1) == 256 and all(16 <= int(t.numel()) < 160 for t in t1) and all(75 <= n < 1000 for n in n1)
    assert all(int(t.min()) >= 1 and int(t.max()) < 255 for t in t1)
    t3, n3 = b.make_workload(256, b.SEED, 1)                 # another rank owns other utterances (weak scaling)
    assert n3 != n1


def test_stage_work_model_matches_the_baseline_formulas():
    b = _bench()
    texts = [torch.zeros(100, dtype=torch.long)]
    w = b.stage_work(texts, [500])
    n, s0 = 500.0, 138.0
    kv = 122880.0
    assert np.isclose(w["t3_bytes"], n * 1.0234e9 + 2 * (s0 * n + n * (n + 1) / 2) * kv + 2 * n * kv)
    T = 2.0 * (500 + 250)
    assert np.isclose(w["flow_flops"], 113.2e6 * 750 + 90.1e3 * 750 * 750 + 10 * 2 * T * (132161536.0 + 114688.0 * T))
    assert np.isclose(w["hift_flops"], 612.3e6 * 1000) and np.isclose(w["hift_bytes"], 0.30e6 * 1000)


def test_tool_scripts_compile():
    for f in glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        py_compile.compile(f, doraise=True)
