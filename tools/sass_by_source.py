"""Static code size of one kernel by source region: SASS instructions per (file, source function), from nvdisasm -g
line annotations of the built library.  Tells where the instruction-cache footprint of the step loop comes from.

    python tools/sass_by_source.py '<0, 16, 16, 1, 4>'     # template arguments of zs_sim_kernel as cuobjdump prints them
"""
import bisect
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "libzombsole_b200", "csrc", "libzs_b200.so")
CSRC = os.path.join(ROOT, "libzombsole_b200", "csrc")


def function_starts(path):
    """[(line, name)] of the __device__ / __global__ functions of a source file (first line of the definition)."""
    out = []
    with open(path) as f:
        for i, line in enumerate(f, 1):
            m = re.search(r"(?:__device__|__global__)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", line)
            if m and not line.strip().startswith("//"):
                out.append((i, m.group(1)))
    return out


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else "<0, 16, 16, 1, 4>"
    mangled = "_Z13zs_sim_kernelILi%sELi%sELi%sELi%sELi%sEEv8ZsParams4ZsIO" % tuple(x.strip() for x in want.strip("<>").split(","))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    starts = {}
    for fn in os.listdir(CSRC):
        if fn.endswith((".cu", ".cuh")):
            starts[fn] = function_starts(os.path.join(CSRC, fn))
    counts = collections.Counter()
    lines = collections.Counter()
    inside, cur = False, ("?", "?")
    total = 0
    for line in txt.splitlines():
        if line.startswith(".text."):
            inside = line.startswith(".text." + mangled + ":")
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            fn, ln = os.path.basename(m.group(1)), int(m.group(2))
            st = starts.get(fn)
            if st:
                k = bisect.bisect_right([s[0] for s in st], ln) - 1
                cur = (fn, st[k][1] if k >= 0 else "?")
            else:
                cur = (fn, "")
            curline = (fn, ln)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            counts[cur] += 1
            total += 1
    print("kernel zs_sim_kernel%s: %d SASS instructions (%.1f KB)" % (want, total, total * 16 / 1024.0))
    for (fn, name), n in counts.most_common(40):
        print("  %6d  %5.1f%%  %s  %s" % (n, 100.0 * n / total, fn, name))


if __name__ == "__main__":
    main()
