#!/usr/bin/env python
"""SASS evidence per kernel of robo_b200/libgpk.so (runs without a GPU): for every kernel the count of the instructions
that prove what it is made of — DMMA (fp64 tensor pipe), UTCIMMA (tcgen05.mma kind::i8; .2CTA = cta_group::2), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), DFMA/DADD/DMUL (fp64 vector pipe), UTMALDG (TMA loads),
SYNCS (mbarrier), LDS/STS, LDG/STG, MUFU, BAR, plus registers from the ELF.  Output: profiles/<tag>_sass_summary.txt
    python tools/sass_summary.py r02"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "robo_b200", "libgpk.so")
KEYS = ["DMMA", "UTCIMMA", "UTCIMMA.2CTA", "LDTM", "UTCBAR", "DFMA", "DADD", "DMUL", "UTMALDG", "SYNCS", "LDS", "STS", "LDG", "STG", "MUFU", "BAR",
        "LDGSTS", "ATOM", "RED"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    regs = {}
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    arch = None
    name = None
    for line in sass.splitlines():
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[name]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[name][k] += 1
    try:
        demangle = subprocess.run(["c++filt"] + list(counts), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    except OSError:
        demangle = []
    if len(demangle) != len(counts):
        demangle = list(counts)                          # no c++filt on this box: keep the mangled names
    out = ["# SASS summary of robo_b200/libgpk.so (%s), cuobjdump -sass; one line per kernel" % arch,
           "# %-70s %6s %5s %6s " % ("kernel", "instr", "regs", "smem") + " ".join("%7s" % k[-7:] for k in KEYS)]
    tot = collections.Counter()
    for (mangled, c), nice in zip(counts.items(), demangle):
        nice = re.sub(r"\(.*", "", nice).replace("void ", "")
        r = regs.get(mangled, (0, 0))
        out.append("%-72s %6d %5d %6d " % (nice[:72], c["_total"], r[0], r[1]) + " ".join("%7d" % c[k] for k in KEYS))
        tot.update(c)
    out.append("%-72s %6d %5s %6s " % ("TOTAL", tot["_total"], "", "") + " ".join("%7d" % tot[k] for k in KEYS))
    path = os.path.join(ROOT, "profiles", "%s_sass_summary.txt" % tag)
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out))
    print("->", path)


if __name__ == "__main__":
    main()
