#!/usr/bin/env python
"""Headline benchmark: audio-seconds generated per wall-second (and RTF) for Chatterbox 0.5B on 256-utterance
synthetic batches, one process per GPU (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this engine (torchrun launches N ranks)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port) timed
                                                             # on the box's host cores on a bounded sample

A "step" = one full pass of the hot path over one batch: T3 prefill + AR decode with CFG -> token clean-up ->
flow encoder -> 10-step CFM with CFG -> HiFT vocoder, for 256 mixed-length utterances per GPU (weak scaling:
every rank owns its own 256 utterances; the only collective is the broadcast of the voice conditionals).
Weights are seeded random-init tensors of the exact reference architecture (no checkpoints / network here);
utterance length is set by per-utterance max_new_tokens ~ U(75, 1000) as SURVEY.md 8d prescribes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_T0 = time.time()


def log(*a):
    if os.environ.get("CBX_BENCH_VERBOSE", "1") != "0" and int(os.environ.get("RANK", 0)) == 0:
        print(f"[bench {time.time() - _T0:7.1f}s]", *a, file=sys.stderr, flush=True)

METRIC = "audio_seconds_per_second"
WORKLOAD = "Chatterbox 0.5B en, batch=256 mixed-length utterances per GPU, CFG, 10-step CFM, paged bf16 KV"
UNIT = "audio-s/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1400.0, "fallback"


def make_workload(batch, seed, rank, budget_max=1000):
    """N_text ~ randint(16,160), ids randint(1,255), N ~ randint(75,1000)."""
    g = torch.Generator().manual_seed(seed + 7919 * rank)
    n_text = torch.randint(16, 160, (batch,), generator=g)
    texts = [torch.randint(1, 255, (int(n),), generator=g) for n in n_text]
    budgets = torch.randint(min(75, budget_max - 1), budget_max, (batch,), generator=g).tolist()
    return texts, budgets


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads():
    """Threads the CPU arm may really use: scheduler affinity and cgroup quota, not the raw core count of the host."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = min(n, max(1, int(float(q[0]) / float(q[1]))))
    except Exception:
        pass
    return max(1, min(n, 64))


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(n_text=40, n_tokens=100, threads=None):
    """The reference's CPU algorithm (oracle port: same op sequence as the reference modules, fp32, torch CPU) on ONE
    utterance: T3.inference (CFG pair, KV cache grown with torch.cat) + flow (10 NFE) + HiFT.  Returns
    (audio_seconds, wall_seconds, split)."""
    from oracle import weights as W
    from oracle.t3_ref import T3Oracle
    from oracle.flow_ref import FlowOracle
    from oracle.hift_ref import HiFTOracle
    threads = threads or host_threads()
    torch.set_num_threads(threads)
    st = cpu_reference_sample.__dict__.setdefault("state", {})
    if not st:
        st["t3"] = T3Oracle(W.make_t3_weights(0))
        st["flow"] = FlowOracle(W.make_flow_weights(0))
        st["hift"] = HiFTOracle(W.make_hift_weights(0))
        st["conds"] = W.make_conds(1234)
    c3, cg = st["conds"]
    g = torch.Generator().manual_seed(5)
    text = torch.randint(1, 255, (n_text,), generator=g)
    tt = F.pad(F.pad(text, (1, 0), value=255), (0, 1), value=0)
    tt = torch.stack([tt, tt])
    torch.manual_seed(0)
    t0 = time.perf_counter()
    toks = st["t3"].inference(c3, tt, n_tokens, temperature=0.8, top_p=1.0, min_p=0.05, repetition_penalty=1.2, cfg_weight=0.5)
    t1 = time.perf_counter()
    sp = toks[0]
    sp = sp[sp < 6561]
    mel = st["flow"].inference(sp, cg, 10)
    t2 = time.perf_counter()
    wav, _ = st["hift"].inference(mel)
    t3 = time.perf_counter()
    audio = sp.numel() / 25.0
    return audio, t3 - t0, dict(t3_s=t1 - t0, flow_s=t2 - t1, hift_s=t3 - t2, tokens=int(sp.numel()))


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    sample = "1 utterance: 40 text tokens, 100 speech tokens (CFG pair), 250-token voice prompt, 10 NFE, HiFT"
    cpu_reference_sample(threads=threads)      # builds weights (untimed)
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_sample(threads=threads)
    audio = wall = 0.0
    for _ in range(args.steps):
        a, w, split = cpu_reference_sample(threads=threads)
        audio += a
        wall += w
    v = audio / wall
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "rtf": wall / audio,
            "config": {"workload": WORKLOAD, "arm": "CPU oracle port of the reference algorithm (fp32, torch CPU ops in the "
                       "reference's order), bounded sample of the workload: one utterance per step",
                       "sample": sample, "split_s": split},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ this engine
def run_engine(args, rank, world, local_rank):
    import torch.distributed as dist
    from oracle import weights as W            # only the seeded synthetic checkpoint generator + cpu_baseline leg
    from chatterbox_b200 import ChatterboxTTS, Conditionals, T3, T3Cond, S3Gen, Engine
    torch.cuda.set_device(local_rank)
    torch.set_num_threads(max(1, host_threads() // max(1, world)))     # host-side weight synthesis: do not oversubscribe
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    t3 = T3(eng, W.make_t3_weights(0))
    s3 = S3Gen(eng, W.make_flow_weights(0), W.make_hift_weights(0))
    # voice conditionals: rank 0 owns them, NCCL-broadcast to the other ranks (north_star "speaker-embedding broadcast")
    from chatterbox_b200.dist import broadcast_conditionals
    c3, cg = W.make_conds(1234) if rank == 0 else (None, None)
    c3, cg = broadcast_conditionals(c3, cg, torch.device("cuda", local_rank), src=0)
    tts = ChatterboxTTS(t3, s3, None, f"cuda:{local_rank}", Conditionals(T3Cond(**c3), cg))
    texts, budgets = make_workload(args.batch, 20260922, rank, args.budget_max)
    log(f"models loaded; batch={args.batch} sum_budget={sum(budgets)}")

    def one_pass(to_host, timings):
        return tts.generate_batch(texts, max_new_tokens=budgets, seed=1000 * rank, kv_dtype="bf16", to_host=to_host,
                                  timings=timings)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(to_host, steps):
        tm_all = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            tm = {}
            one_pass(to_host, tm)
            tm_all.append(tm)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1000.0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1000.0
        return ms, wall, tm_all

    for i in range(args.warmup):
        tmw = {}
        one_pass(False, tmw)
        log(f"warmup {i}: " + json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in tmw.items()}))
    launches0 = eng.h.launch_count()
    eng.stats.update(paged_bytes=0.0, paged_launches=0, decode_steps=0, decode_row_steps=0)
    eng.h.set_option("time_kernel", "gemm_tc")        # dominant kernel family of the step
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, wall, tms = timed(False, args.steps)
    clk = clocks.stop() if rank == 0 else None
    log(f"timed: {ms:.1f} ms for {args.steps} steps")
    gemm_ms, gemm_n, gemm_flops = eng.h.timer_read()
    gemm_bytes = eng.h.timer_read_bytes()
    stats_timed = dict(eng.stats)
    eng.h.set_option("time_kernel", "paged")          # second family, measured during the e2e pass below
    launches = eng.h.launch_count() - launches0
    audio = sum(t["audio_s"] for t in tms)
    audio_t = torch.tensor([audio], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(audio_t)
    audio_total = float(audio_t[0])
    # e2e: the public batch API with host buffers (token ids in host memory, waveforms copied back to pinned host memory)
    eng.stats.update(paged_bytes=0.0, paged_launches=0, decode_steps=0, decode_row_steps=0)
    ms_e, wall_e, tms_e = timed(True, 1)
    paged_ms, paged_n, _ = eng.h.timer_read()
    eng.h.set_option("time_kernel", "none")
    stats_e2e = dict(eng.stats)            # frozen here: later passes (B=1 latency) must not leak into the ratio
    audio_e = torch.tensor([sum(t["audio_s"] for t in tms_e)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(audio_e)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # BASELINE config[1]: one utterance (B=1, CFG pair), 150 speech tokens = 6 s of audio: latency / RTF
    g1 = torch.Generator().manual_seed(1234)
    text1 = [torch.randint(1, 255, (100,), generator=g1)]
    tts.generate_batch(text1, max_new_tokens=[150], seed=7, kv_dtype="bf16", to_host=True, timings={})
    tm1 = {}
    tts.generate_batch(text1, max_new_tokens=[150], seed=7, kv_dtype="bf16", to_host=True, timings=tm1)
    b1_ms = tm1["t3_ms"] + tm1["flow_ms"] + tm1["hift_ms"] + tm1["d2h_ms"]
    b1 = {"audio_s": tm1["audio_s"], "latency_ms": b1_ms, "rtf": b1_ms / 1000.0 / max(tm1["audio_s"], 1e-9),
          "t3_ms": tm1["t3_ms"], "flow_ms": tm1["flow_ms"], "hift_ms": tm1["hift_ms"]}
    log("B=1 latency: " + json.dumps(b1))
    hbm_peak, tf_peak, peak_src = load_peaks()
    paged_bytes_per_launch = stats_e2e["paged_bytes"] / max(1, stats_e2e["paged_launches"])
    achieved = (stats_e2e["paged_bytes"] / 1e9) / (paged_ms / 1e3) if paged_ms > 0 else 0.0
    stage = {k: sum(t[k] for t in tms) / len(tms) for k in ("t3_ms", "flow_ms", "hift_ms")}
    log(f"cpu baseline on {host_threads()} threads (os.cpu_count={os.cpu_count()})")
    cpu_audio, cpu_wall, cpu_split = cpu_reference_sample()
    log(f"cpu baseline done: {cpu_wall:.1f}s")
    value = audio_total / (ms / 1000.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16 weights/KV, fp32 activations+accumulate (activations split hi+lo bf16 on the tensor cores)",
        "data": "synthetic",
        "rtf": (ms / 1000.0) / audio_total * world,
        "config": {"workload": WORKLOAD,
                   "utterances_per_gpu": args.batch, "global_batch": args.batch * world, "parallelism": f"utterance-sharded dp{world}",
                   "weights": "seeded random init of the reference architecture (532M T3 + 112M flow + 21M HiFT)",
                   "l2_policy": "working set >> L2 (KV pages ~40 GB, activations GBs); no explicit flush needed",
                   "audio_s_per_step_per_gpu": audio / args.steps, "stage_ms": stage,
                   "decode_steps_per_step": stats_timed["decode_steps"] / args.steps, "peaks": peak_src,
                   "b1_latency": b1},
        "clocks": clk,
        "gpu_launches": int(launches),
        "e2e": {"value": float(audio_e[0]) / (ms_e / 1000.0), "unit": UNIT,
                "h2d_bytes_per_step": int(tms_e[0]["h2d_bytes"]), "d2h_bytes_per_step": int(tms_e[0]["d2h_bytes"]),
                "wall_s": wall_e},
        "roofline": {"kernel": "gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv family: T3 projections, CFM, HiFT convs)",
                     "bound": "tensor", "achieved": gemm_flops / 1e12 / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0,
                     "peak": tf_peak, "unit": "TFLOP/s", "frac": (gemm_flops / 1e12 / (gemm_ms / 1e3)) / tf_peak if gemm_ms > 0 else 0.0,
                     "traffic": None, "algorithmic_flops_per_launch": gemm_flops / max(1, gemm_n), "launches": gemm_n,
                     "avg_launch_ms": gemm_ms / max(1, gemm_n), "share_of_step": gemm_ms / ms,
                     "note": "algorithmic flops = 2*M*N*K per launch (the bf16 hi/lo activation split issues 2x that on the tensor pipe)",
                     # the same launches against the other roof: weights, activations and results once each (DESIGN.md 7:
                     # the K=256 CFM block GEMMs sit below the ridge at fp32-precision activations)
                     "hbm_view": {"algorithmic_bytes_per_launch": gemm_bytes / max(1, gemm_n),
                                  "achieved": gemm_bytes / 1e9 / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0, "peak": hbm_peak,
                                  "unit": "GB/s", "frac": (gemm_bytes / 1e9 / (gemm_ms / 1e3)) / hbm_peak if gemm_ms > 0 else 0.0}},
        "roofline_secondary": {"kernel": "paged_decode_kernel<bf16> (T3 decode attention over the paged KV cache)", "bound": "hbm",
                               "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                               "algorithmic_bytes_per_launch": paged_bytes_per_launch, "launches": paged_n,
                               "avg_launch_ms": paged_ms / max(1, paged_n), "share_of_step": paged_ms / ms_e,
                               "measured_in": "e2e pass"},
        "cpu_baseline": {"value": cpu_audio / cpu_wall, "unit": UNIT, "cores": host_threads(), "kind": "port",
                         "sample": "1 utterance (40 text tokens, 100 speech tokens, 250-token prompt) through the oracle "
                                   "port of the reference's CPU path", "split_s": cpu_split},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("CBX_BENCH_WATCHDOG", 1500)), exit=True, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--budget-max", type=int, default=1000, help="upper bound of the per-utterance token budget (debug)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
