"""Condense an `ncu --set full` details page (stdout log) to its metric tables: drops the OPT/INF prose and separators.
    python tools/ncu_log_trim.py gpurun_out/s9_ncu_r2_attn_tc.log profiles/r2_ncu_attn_tc.txt "header"
"""
import re, sys
src, dst = sys.argv[1], sys.argv[2]
header = sys.argv[3] if len(sys.argv) > 3 else ""
out = [f"# {header}", f"# condensed from the ncu --set full --clock-control none details page ({src.split('/')[-1]}); prose removed"]
keep = False
skip_block = False
for line in open(src, errors="ignore"):
    s = line.rstrip("\n")
    if re.match(r"^  \S.*\(\d+, \d+, \d+\)x\(\d+, \d+, \d+\)", s):      # kernel header
        out.append(""); out.append(s.strip()); keep = True; skip_block = False; continue
    if not keep:
        continue
    if re.match(r"^\s+(OPT|INF|WRN)\s", s):
        skip_block = True; continue
    if re.match(r"^\s+Section:", s):
        skip_block = False; out.append(s.strip()); continue
    if skip_block or not s.strip() or re.match(r"^\s+-{5,}", s) or "Metric Name" in s or s.strip().startswith("Warning:"):
        continue
    out.append("  " + re.sub(r"\s{2,}", "  ", s.strip()))
open(dst, "w").write("\n".join(out) + "\n")
print(dst, len(out), "lines")
