"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share.
    python tools/launch_summary.py launches.csv "command that was profiled" > summary.txt"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
k_i, m_i, v_i, u_i = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
for r in rows[hdr_i + 1:]:
    if len(r) <= v_i or r[m_i] != "gpu__time_duration.sum":
        continue
    ms = float(r[v_i].replace(",", "")) * scale.get(r[u_i], 1e-6)
    tot[r[k_i]] += ms
    cnt[r[k_i]] += 1
total = sum(tot.values())
print("kernel launches of `%s` (ncu --metrics gpu__time_duration.sum, serialised, cold cache)" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
for k, ms in tot.most_common():
    print("%-72s launches %3d  total %9.3f ms  share %5.1f%%  mean %8.3f ms" % (k[:72], cnt[k], ms, 100 * ms / total, ms / cnt[k]))
print("total %.3f ms in %d launches" % (total, sum(cnt.values())))
