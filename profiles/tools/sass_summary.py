#!/usr/bin/env python
"""Per-kernel SASS evidence of the shipped library: counts of the mnemonics that prove the Blackwell-native paths
(B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP / UTMA* = bulk TMA,
LDGMC / multimem = NVLS, REDG = red.global, SYNCS = mbarrier), plus registers / shared memory per kernel from cuobjdump.
    python profiles/tools/sass_summary.py > profiles/sass_summary.md      (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "single-stable-dreamfusion_b200", "ngp_b200", "lib", "libngp_b200.so")
KEYS = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"), ("UBLKCP", r"\bUBLKCP"),
        ("UTMA", r"\bUTMA(LDG|STG)"), ("SYNCS", r"\bSYNCS"), ("LDGMC", r"\bLDGMC|MULTIMEM"), ("REDG", r"\bREDG?\.E"), ("REDG.x2/x4", r"\bREDG?\.E\.ADD\.F32\.?(x2|x4|\.64|\.128)|RED\.E\.ADD\.F32x[24]"),
        ("ATOMG", r"\bATOMG"), ("HMMA", r"\bHMMA"), ("SHFL", r"\bSHFL"), ("MUFU", r"\bMUFU")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    archs = set()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            archs.add(m.group(1))
        if cur is None or "/*" not in line:
            continue
        counts[cur]["instr"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
        for key, pat in KEYS:
            if re.search(pat, line):
                counts[cur][key] += 1
    dm = demangle(list(counts))
    print("# SASS summary of libngp_b200.so (cuobjdump -sass; architectures: %s)\n" % ", ".join(sorted(archs)))
    print("What each column proves: UTCHMMA = `tcgen05.mma`, LDTM = `tcgen05.ld`, UTCBAR = `tcgen05.commit`, UBLKCP = bulk (TMA) copies,\n"
          "SYNCS = mbarrier ops, LDGMC = `multimem.ld_reduce` / `multimem.st` (NVLS), REDG = `red.global.add` (x2/x4 = vector reds).\n")
    hdr = ["kernel", "instr", "regs", "smem"] + [k for k, _ in KEYS]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    tot = collections.Counter()
    for fn, c in counts.items():
        name = dm.get(fn, fn)
        name = re.sub(r"\(.*", "", name).replace("ngp::", "")
        if len(name) > 70:
            name = name[:67] + "..."
        r, s = usage.get(fn, ("", ""))
        row = [name, str(c["instr"]), str(r), str(s)] + [str(c[k]) if c[k] else "" for k, _ in KEYS]
        print("| " + " | ".join(row) + " |")
        tot.update(c)
    print("| **total** | %d | | | %s |" % (tot["instr"], " | ".join(str(tot[k]) for k, _ in KEYS)))


if __name__ == "__main__":
    sys.exit(main())
