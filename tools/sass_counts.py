"""cuobjdump -sass of libcbx.so -> per-kernel counts of the mnemonics that prove the tcgen05 / TMA / bulk-copy paths
(B200_PROFILING.md): UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor load / store), LDTM / STTM (tcgen05.ld / st),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), MUFU.EX2, FFMA2, FMNMX3.   python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chatterbox_b200", "libcbx.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UBLKCP", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "FMNMX3", "HMMA", "STG.E.ENL2.256", "LDL", "STL"]
cur, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("cbx::", "")
        order.append(cur); continue
    if cur is None: continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m: continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    for k in keys:
        if op.startswith(k): counts[cur][k] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.basename(so)} (sm_100a); total = instructions in the kernel")
print(f"{'kernel':52s} {'total':>7s} " + " ".join(f"{k[:9]:>9s}" for k in keys))
for k in order:
    c = counts[k]
    if not any(c[x] for x in keys[:6]) and c["_total"] < 2000: continue
    print(f"{k[:52]:52s} {c['_total']:7d} " + " ".join(f"{c[x]:9d}" for x in keys))
