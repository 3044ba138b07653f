"""Per-kernel census of the SASS mnemonics that prove which hardware paths the built library uses
(B200_PROFILING.md: UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTMALDG = tensor-map
TMA, DMMA = FP64 tensor pipe, SYNCS = mbarrier).  Runs on the CPU box: `python scripts/sass_census.py [lib.so]`.
Prints a markdown table (kernels with at least one tensor / bulk-copy instruction, then totals)."""
import collections
import os
import re
import subprocess
import sys

MNEMONICS = ["UTCIMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "LDGSTS", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name, width=110):
    name = re.sub(r"\((anonymous namespace)\)::", "", name)
    name = re.sub(r"\(.*\)$", "", name)           # drop the argument list
    name = re.sub(r"^void ", "", name)
    return name if len(name) <= width else name[: width - 3] + "..."


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "reachabilityanalysis.jl_b200", "libreach_cuda.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    pat = re.compile(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            counts[cur] = collections.Counter()
            continue
        m = pat.match(line)
        if m and cur is not None:
            op = m.group(1)
            counts[cur][op.split(".")[0]] += 1
            counts[cur]["_all"] += 1
    names = demangle(list(counts))
    print("| kernel | instr | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|" + "---|" * len(MNEMONICS))
    total = collections.Counter()
    for k, c in counts.items():
        total.update(c)
        if not (c["UTCIMMA"] or c["DMMA"] or c["UBLKCP"] or c["UTMALDG"]):
            continue
        print("| `%s` | %d | " % (short(names[k]), c["_all"]) + " | ".join(str(c[m]) for m in MNEMONICS) + " |")
    print("| **all %d kernels** | %d | " % (len(counts), total["_all"]) + " | ".join(str(total[m]) for m in MNEMONICS) + " |")


if __name__ == "__main__":
    main()
