"""Executed warp-instructions per SASS symbol (kernel body vs out-of-line subroutines) + hottest call sites.
    python tools/ncu_by_symbol.py rep.ncu-rep kernel-substring [warp_steps]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gym_futbol_b200", "csrc", "libfutbol_b200.so")
rep, sub = sys.argv[1], sys.argv[2]
ws = float(sys.argv[3]) if len(sys.argv) > 3 else 524288.0
tmp = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
dis = []
for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin") and os.path.getsize(os.path.join(tmp, f)) > 10000):
    dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
best = None
active, label, tab, tabs = False, 'main', {}, {}
for ln in dis:
    m = re.match(r"\.text\.(\S+):", ln)
    if m:
        active = sub in m.group(1); label = 'kernel body'; tab = tabs.setdefault(m.group(1), {}) if active else {}
        continue
    if not active: continue
    m = re.match(r"^(\$\S+|_Z\S+):", ln.strip())
    if m and not m.group(1).startswith('.L'): label = m.group(1)
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: tab[int(m.group(1), 16)] = (label, m.group(2))
tab = tabs[min(tabs, key=lambda k: abs(len(tabs[k]) - len(data)))]
base = int(data[0][ix["Address"]], 16)
c, t, st = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    lab = tab[int(r[ix["Address"]], 16) - base][0]
    c[lab] += int(r[ix["Instructions Executed"]]); t[lab] += int(r[ix["Thread Instructions Executed"]]); st[lab] += 1
tot = sum(c.values())
def short(s):
    m = re.search(r"futbol(\d+)(\w+)", s)
    if s.startswith('$') and m: return m.group(2)[:int(m.group(1))] if False else s[:70]
    return s[:70]
for k, v in c.most_common():
    print("%-72s static %4d  %6.2f%%  %7.1f/warp-step  threads %.1f" % (short(k), st[k], 100 * v / tot, v / ws, t[k] / max(1, v)))
