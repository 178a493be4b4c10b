"""Opcode histogram of every kernel in the built library (cuobjdump -sass): evidence for what the kernels are made of —
UBLKCP (cp.async.bulk, the TMA bulk copies), SYNCS (mbarrier), MATCH / REDUX / VOTE / SHFL (warp primitives), no tensor-core
or TMEM opcodes (nothing here is a contraction).

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "libzombsole_b200", "csrc", "libzs_b200.so")
KEY = ["UBLKCP", "SYNCS", "MATCH", "REDUX", "VOTE", "VOTEU", "SHFL", "LDS", "STS", "LDG", "STG", "LDC", "ATOMS", "ATOMG", "RED",
       "BAR", "WARPSYNC", "BSSY", "BSYNC", "CALL", "IMAD", "HMMA", "UTCMMA", "UTMALDG", "UTMASTG", "LDTM"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n")
    return dict(zip(names, out))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = re.findall(r"arch = (sm_\w+)", txt)
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    dm = demangle(list(kernels))
    print("library: %s   cubin architectures: %s" % (os.path.relpath(LIB, ROOT), sorted(set(arch))))
    print("%d kernels; columns: total SASS instructions, then the opcodes of interest (base opcode, all suffixes summed)\n" % len(kernels))
    tot_all = collections.Counter()
    for k, c in kernels.items():
        base = collections.Counter()
        for op, n in c.items():
            base[op.split(".")[0]] += n
        tot_all.update(base)
        name = dm.get(k, k)
        name = re.sub(r"^void ", "", name).replace("(ZsParams, ZsIO)", "")
        cols = "  ".join("%s %d" % (op, base[op]) for op in KEY if base[op])
        print("%-58s %6d  %s" % (name[:58], sum(c.values()), cols))
    print("\nwhole library: " + "  ".join("%s %d" % (op, tot_all[op]) for op in KEY))
    ublk = collections.Counter()
    for c in kernels.values():
        for op, n in c.items():
            if op.startswith("UBLKCP") or op.startswith("SYNCS") or op.startswith("MATCH") or op.startswith("REDUX"):
                ublk[op] += n
    print("full mnemonics: " + "  ".join("%s %d" % kv for kv in sorted(ublk.items())))


if __name__ == "__main__":
    main()
