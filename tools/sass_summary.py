#!/usr/bin/env python
"""Per-kernel SASS evidence of libri_b200.so:  python tools/sass_summary.py [out.md]

Counts, for every kernel of the built library (cuobjdump -sass), the instructions that prove which Blackwell engines it
drives: UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st: TMEM), UTCBAR (tcgen05.commit), UTCATOMSWS
(TMEM allocation), UBLKCP (cp.async.bulk: the TMA engine's linear form), UTMALDG / UTMASTG (tensor-map TMA), SYNCS
(mbarrier), LDGSTS (cp.async), plus registers from the ELF section info.  Library (CUB) kernels are listed by count only.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "point-cloud-registration-based-on-rotation-invariant-feature_b200", "libri_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS",
             "LDGSTS", "HMMA", "REDUX", "DFMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"^void ", "", n)
    depth, cut = 0, len(n)
    for i, ch in enumerate(n):                 # drop the argument list, keep template arguments
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return n[:cut]


def main():
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.md")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", subprocess.run(
            ["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout):
        regs[m.group(1)] = int(m.group(2))
    counts, total, cur = collections.OrderedDict(), collections.Counter(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if cur and m:
            total[cur] += 1
            op = m.group(1)
            for k in MNEMONICS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
    names = demangle(list(counts))
    ours = [(short(names[k]), k) for k in counts if "cub::" not in names[k]]
    lib = [k for k in counts if "cub::" in names[k]]
    cols = [c for c in MNEMONICS if any(counts[k][c] for _, k in ours)]
    with open(dst, "w") as f:
        f.write("# SASS summary of `libri_b200.so` (sm_100a) — `python tools/sass_summary.py`\n\n")
        f.write("Instruction counts per kernel from `cuobjdump -sass`; `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld (TMEM), "
                "`UTCBAR` = tcgen05.commit, `UTCATOMSWS` = TMEM alloc, `UBLKCP` = cp.async.bulk (TMA engine), `SYNCS` = mbarrier, "
                "`LDGSTS` = cp.async, `REDUX` = redux.sync, `DFMA` = fp64 FMA.\n\n")
        f.write("| kernel | regs | SASS insts | " + " | ".join(cols) + " |\n|---|---|---|" + "---|" * len(cols) + "\n")
        for s, k in sorted(ours):
            f.write("| `%s` | %s | %d | %s |\n" % (s, regs.get(k, ""), total[k],
                                                 " | ".join(str(counts[k][c]) if counts[k][c] else "" for c in cols)))
        f.write("\n%d kernels of this repo; %d CUB kernels (radix sort / scan of the scan-sized voxelizer and grid "
                "subsampling) are linked in as library code.\n" % (len(ours), len(lib)))
    print("wrote", dst, len(ours), "kernels")


if __name__ == "__main__":
    main()
