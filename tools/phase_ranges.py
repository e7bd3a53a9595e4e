"""Sums the dyn% column of tools/static_by_phase.py output over named line ranges of the anchor function.
    PHASE_TOP=100000 python tools/static_by_phase.py <kernel> <anchor> <rep> | python tools/phase_ranges.py file.cuh 389-399:name ..."""
import collections
import re
import sys

fname = sys.argv[1]
ranges = []
for spec in sys.argv[2:]:
    r, name = spec.split(":")
    a, b = r.split("-")
    ranges.append((int(a), int(b), name))
agg, smp = collections.Counter(), collections.Counter()
for ln in sys.stdin:
    m = re.match(r"(\S+):(\d+)\s+(\S+)\s+(\d+)\s+(\d+)\s+([\d.]+)%\s+([\d.]+)%", ln)
    if not m:
        continue
    f, l, dyn, sm = m.group(1), int(m.group(2)), float(m.group(6)), float(m.group(7))
    if f == fname:
        for a, b, name in ranges:
            if a <= l <= b:
                agg[name] += dyn
                smp[name] += sm
                break
        else:
            agg["%s:%d" % (f, l)] += dyn
            smp["%s:%d" % (f, l)] += sm
    else:
        agg["outside (%s)" % f] += dyn
        smp["outside (%s)" % f] += sm
print("%-44s %7s %7s" % ("phase", "inst%", "time%"))
for k, v in agg.most_common(int(sys.argv[0] and 14)):
    print("%-44s %6.2f%% %6.2f%%" % (k, v, smp[k]))
print("%-44s %6.2f%% %6.2f%%" % ("total", sum(agg.values()), sum(smp.values())))
