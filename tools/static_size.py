"""Static SASS size of a kernel by source line / enclosing function (instruction-cache footprint).

    python tools/static_size.py [kernel-substring] [top-N]

`nvdisasm -g` line info of the in-tree library; instructions are attributed to the innermost source line, lines to
the function whose definition precedes them in the file.  16 bytes per SASS instruction.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gym_futbol_b200", "csrc")
LIB = os.path.join(CSRC, os.environ.get("FUTBOL_B200_LIB", "libfutbol_b200.so"))


def functions_of(path):
    """[(first_line, name)] of function definitions, by a loose regex good enough for these sources."""
    out = []
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__).*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(")
    try:
        for n, ln in enumerate(open(path), 1):
            m = pat.match(ln)
            if m:
                out.append((n, m.group(1)))
    except OSError:
        pass
    return out


def main():
    sub = sys.argv[1] if len(sys.argv) > 1 else "v0_rollout_kernelILb0"
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    per_line = collections.Counter()
    per_section = collections.Counter()
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        active, cur, sec = False, None, None
        for ln in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln) or re.match(r"\.text\.(\S+):", ln)
            if m:
                active = sub in m.group(1)
                sec = m.group(1)
                continue
            if not active:
                continue
            m = re.match(r"\s*(\$?[A-Za-z_][\w$]*):\s*$", ln)
            if m and m.group(1).startswith("$"):
                sec = m.group(1)       # out-of-line subroutine label inside the kernel's section
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                per_line[cur] += 1
                per_section[sec[:60]] += 1
    total = sum(per_line.values())
    print("kernel *%s*: %d SASS instructions = %.1f KB" % (sub, total, total * 16 / 1024))
    for s, n in per_section.most_common():
        print("  section %-62s %5d" % (s, n))
    funcs = {}
    per_func = collections.Counter()
    for (f, l), n in per_line.items():
        if f not in funcs:
            funcs[f] = functions_of(os.path.join(CSRC, f))
        name = "?"
        for first, fn in funcs[f]:
            if first <= l:
                name = fn
        per_func[(f, name)] += n
    print("by enclosing function:")
    for (f, name), n in per_func.most_common(top):
        print("  %-18s %-28s %5d  %5.1f%%" % (f, name, n, 100 * n / total))


if __name__ == "__main__":
    main()
