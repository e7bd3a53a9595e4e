"""Per-function share of executed warp-instructions of a kernel (functions of v0_step.cuh / philox.cuh / v0_kernels.cu).
    python tools/ncu_by_func.py rep.ncu-rep kernel-substring warp_steps"""
import collections, csv, io, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_by_line as T

rep, sub = sys.argv[1], sys.argv[2]
warp_steps = float(sys.argv[3]) if len(sys.argv) > 3 else 524288.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
base = int(data[0][ix["Address"]], 16)
tables = T.line_table(sub); fn = min(tables, key=lambda k: abs(len(tables[k]) - len(data))); tab = tables[fn]
marks = {}
for f in ("v0_step.cuh", "philox.cuh", "v0_kernels.cu"):
    src = open(os.path.join(T.ROOT, "gym_futbol_b200", "csrc", f)).read().splitlines()
    m_ = []
    for i, l in enumerate(src, 1):
        m = re.match(r'(?:static |inline |template <[^>]*> )*(?:__device__|__global__|__host__ __device__)[^(]*?(\w+)\(', l)
        if m: m_.append((i, m.group(1)))
    marks[f] = m_
def func(key):
    f, l = key
    name = f
    for i, n in marks.get(f, []):
        if i <= l: name = n
    return name
wr = {'dmul', 'dadd', 'dsub', 'ddiv', 'sqsum', 'hyp', 'pick', 'f', 'draw', 'make_lane'}
inst, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
cur = 'prologue'
for r in data:
    off = int(r[ix["Address"]], 16) - base
    key = tab.get(off, (("?", 0), ""))[0]
    f = func(key) if key else '?'
    if f not in wr and not f.startswith('sm_'): cur = f
    inst[cur] += int(r[ix["Instructions Executed"]]); thr[cur] += int(r[ix["Thread Instructions Executed"]]); smp[cur] += int(r[ix["# Samples"]])
tot, ts = sum(inst.values()), sum(smp.values())
print("static SASS %d, executed warp-instructions %d = %.0f per warp-step" % (len(data), tot, tot / warp_steps))
for k, n in inst.most_common():
    print("%-24s %6.2f%%  samples %5.1f%%  %7.1f inst/warp-step  avg active threads %.1f" % (k, 100 * n / tot, 100 * smp[k] / ts, n / warp_steps, thr[k] / max(1, n)))
