"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum": continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("cbx::", "")
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print(f"# {sys.argv[1]}: {sum(cnt.values())} launches, {total/1000:.2f} ms device time (ncu: cold-cache, serialised -> compare shares)")
print(f"{'kernel':60s} {'launches':>9s} {'total_us':>12s} {'avg_us':>9s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k[:60]:60s} {cnt[k]:9d} {v:12.1f} {v/cnt[k]:9.2f} {100*v/total:6.1f}%")
