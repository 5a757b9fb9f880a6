"""Summarise an `ncu --set full` report here (no GPU needed): key metrics per launch -> profiles/<name>.txt, and the
per-launch DRAM traffic of the kernel -> profiles/r2_traffic.json (read by bench.py's roofline.traffic).

    python tools/ncu_summary.py gpurun_out/r2_gemm_tc.ncu-rep profiles/r2_ncu_gemm_tc.txt gemm_tc_kernel "header text"
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * mult.get(unit, 1)


def main():
    rep, out, kname = sys.argv[1], sys.argv[2], sys.argv[3]
    header = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    lines = [f"# {header}", f"# source: {os.path.basename(rep)} (ncu --set full --clock-control none; binary report not committed)"]
    traffic = []
    for r in data:
        lines.append("--- launch")
        for k in KEYS:
            if k in col:
                lines.append(f"  {k} = {r[col[k]]} {units[col[k]]}")
        st = []
        tot = 0.0
        for h in stall_cols:
            try:
                v = float(r[col[h]].replace(",", ""))
            except ValueError:
                continue
            st.append((v, h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            tot += v
        st.sort(reverse=True)
        if tot > 0:
            lines.append("  stall reasons (pc sampling): " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:7]))
        if "dram__bytes_read.sum" in col:
            traffic.append(to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) +
                           to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]))
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    open(out, "w").write("\n".join(lines) + "\n")
    tp = os.path.join(os.path.dirname(os.path.abspath(out)), "r2_traffic.json")
    tj = json.load(open(tp)) if os.path.exists(tp) else {}
    if traffic:
        tj[kname] = {"dram_bytes_per_launch": sum(traffic) / len(traffic), "launches_captured": len(traffic),
                     "from": os.path.basename(out)}
        json.dump(tj, open(tp, "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
