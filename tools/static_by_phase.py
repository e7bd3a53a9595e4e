"""Static SASS size of a kernel attributed through the inlining chain (instruction-cache footprint by phase).

    python tools/static_by_phase.py [kernel-substring] [anchor-function] [ncu-report]

Every instruction carries its chain of inlined frames (`nvdisasm -gi`).  It is attributed to the frame that lies
in `anchor-function` (default v0_step: the line of the step that the code was expanded from) and to the function
called from that line; instructions outside the anchor go to the kernel line.  With an ncu report of the SAME
library the executed-instruction counts are joined by offset.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gym_futbol_b200", "csrc")
LIB = os.path.join(CSRC, os.environ.get("FUTBOL_B200_LIB", "libfutbol_b200.so"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from static_size import functions_of  # noqa: E402

_funcs = {}


def func_at(f, l):
    if f not in _funcs:
        _funcs[f] = functions_of(os.path.join(CSRC, f))
    name = "?"
    for first, fn in _funcs[f]:
        if first <= l:
            name = fn
    return name


def chains(sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        active, fn, chain, pending = False, None, [], []
        for ln in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln) or re.match(r"\.text\.(\S+):", ln)
            if m:
                fn = m.group(1)
                active = sub in fn
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at)?', ln)
            if m:
                pending.append((os.path.basename(m.group(1)), int(m.group(2))))
                if not m.group(3):          # outermost frame closes the chain
                    chain, pending = pending, []
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.setdefault(fn, {})[int(m.group(1), 16)] = (tuple(chain), m.group(2).strip())
    return out


def main():
    sub = sys.argv[1] if len(sys.argv) > 1 else "v0_rollout_kernelILb0"
    anchor = sys.argv[2] if len(sys.argv) > 2 else "v0_step"
    rep = sys.argv[3] if len(sys.argv) > 3 else None
    tabs = chains(sub)
    fn = max(tabs, key=lambda k: len(tabs[k]))
    tab = tabs[fn]
    execd, samples = {}, {}
    if rep:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        ix = {h: i for i, h in enumerate(rows[1])}
        data = rows[2:]
        base = int(data[0][ix["Address"]], 16)
        assert abs(len(data) - len(tab)) < 16, "report does not match the library (%d vs %d instructions)" % (len(data), len(tab))
        for r in data:
            execd[int(r[ix["Address"]], 16) - base] = int(r[ix["Instructions Executed"]])
            samples[int(r[ix["Address"]], 16) - base] = int(r[ix["# Samples"]])
    per = collections.defaultdict(lambda: [0, 0, 0, 0])      # key -> [static, static executed, dynamic, stall samples]
    for off, (chain, _) in tab.items():
        key = None
        for i, (f, l) in enumerate(chain):
            if func_at(f, l) == anchor:
                callee = func_at(*chain[i - 1]) if i > 0 else "-"
                key = ("%s:%d" % (f, l), callee)
                break
        if key is None:
            f, l = chain[-1] if chain else ("?", 0)
            inner = func_at(*chain[0]) if chain else "?"
            key = ("%s:%d" % (f, l), inner)
        p = per[key]
        p[0] += 1
        e = execd.get(off, 0)
        p[1] += 1 if e > 0 else 0
        p[2] += e
        p[3] += samples.get(off, 0)
    tot = sum(p[0] for p in per.values())
    dyn = sum(p[2] for p in per.values()) or 1
    smp = sum(p[3] for p in per.values()) or 1
    print("kernel %s: %d SASS instructions (%.1f KB)%s" % (fn[:50], tot, tot / 64.0, ", executed %d (%.1f KB)" % (
        sum(p[1] for p in per.values()), sum(p[1] for p in per.values()) / 64.0) if rep else ""))
    print("%-22s %-22s %6s %6s %7s %7s" % ("line in " + anchor, "callee", "static", "exec'd", "dyn%", "samp%"))
    for key, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("PHASE_TOP", 60))]:
        print("%-22s %-22s %6d %6d %6.2f%% %6.2f%%" % (key[0], key[1], p[0], p[1], 100.0 * p[2] / dyn, 100.0 * p[3] / smp))


if __name__ == "__main__":
    main()
