"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.

    python tools/ncu_by_line.py gpurun_out/prof.ncu-rep [kernel-substring] [top-N]

Joins `ncu --page source --csv` (SASS view: executed instructions, stall samples) with `nvdisasm -g`
line info of the in-tree library by instruction offset.  Needs the .so the report was taken with.
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
LIB = os.path.join(ROOT, "gym_futbol_b200", "csrc", "libfutbol_b200.so")


def line_table(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    table = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin") or os.path.getsize(os.path.join(tmp, f)) < 10000:
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, fn, active = None, None, False
        for ln in dis.splitlines():
            m = re.match(r"\.text\.(\S+):", ln)
            if m:
                fn = m.group(1)
                active = kernel_sub in fn
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                table.setdefault(fn, {})[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else "rollout"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    base = int(data[0][ix["Address"]], 16)
    tables = line_table(sub)
    # pick the function whose instruction count matches
    fn = min(tables, key=lambda k: abs(len(tables[k]) - len(data)))
    tab = tables[fn]
    print("kernel", fn, "sass", len(data), "disasm", len(tab))
    inst, thr, samp, noinst = (collections.Counter() for _ in range(4))
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        key = tab.get(off, (("?", 0), ""))[0]
        inst[key] += int(r[ix["Instructions Executed"]])
        thr[key] += int(r[ix["Thread Instructions Executed"]])
        samp[key] += int(r[ix["# Samples"]])
        noinst[key] += int(r[ix["stall_no_inst"]]) if "stall_no_inst" in ix else 0
    ti, ts = sum(inst.values()), sum(samp.values())
    print("total warp-instructions %d, samples %d" % (ti, ts))
    src = {}
    print("%-22s %7s %7s %6s %7s  %s" % ("file:line", "inst%", "samp%", "thr", "noinst%", "source"))
    for key, n in inst.most_common(top):
        f, l = key
        if f not in src:
            try:
                src[f] = open(os.path.join(ROOT, "gym_futbol_b200", "csrc", f)).read().splitlines()
            except OSError:
                src[f] = []
        text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
        print("%-22s %6.2f%% %6.2f%% %6.1f %6.2f%%  %s" % ("%s:%d" % (f, l), 100 * n / ti, 100 * samp[key] / max(1, ts),
                                                        thr[key] / max(1, n), 100 * noinst[key] / max(1, ts), text))


if __name__ == "__main__":
    main()
