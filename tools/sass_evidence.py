#!/usr/bin/env python3
"""profiles/sass_<tag>.txt: per kernel of libviorb_b200.so, the counts of the SASS mnemonics that show how the kernel moves
data and does its arithmetic (TMA, cp.async, DPX min/max, integer dot products, warp reductions, population counts, tensor
cores), plus the architecture of every cubin.  usage: tools/sass_evidence.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "viorb_b200", "lib", "libviorb_b200.so")
KEYS = ["UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDG.E.128", "STG.E.128", "LDS.128", "VIMNMX3", "VIMNMX", "VIADDMNMX", "VABSDIFF4", "IDP.4A", "IDP.2A",
        "REDUX", "POPC", "PRMT", "SHFL", "VOTE", "ATOMS", "ATOMG", "RED.E", "BAR.SYNC", "ACQBULK", "PREEXIT", "UTCHMMA", "UTCIMMA", "UTCQMMA", "HMMA", "IMMA", "LDTM", "DFMA", "DMUL"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    per = collections.OrderedDict()
    name = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", ln)
        if m and name:
            op = m.group(1)
            per[name]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k.count(".") and op.startswith(k)):
                    per[name][k] += 1
    out = ["# SASS evidence (%s): cuobjdump -sass viorb_b200/lib/libviorb_b200.so, static instruction counts per kernel" % tag,
           "# cubin architectures: %s" % ", ".join(archs),
           "# TMA = UTMALDG (+ SYNCS mbarrier waits); cp.async = LDGSTS; DPX = VIMNMX3 / VIADDMNMX; programmatic dependent launch = ACQBULK / PREEXIT;",
           "# tensor cores would show as UTC*MMA / HMMA / IMMA / LDTM: none, by design (nothing on this path is a contraction)", ""]
    for name, c in per.items():
        keys = [k for k in KEYS if c[k]]
        out.append("%-58s %6d instr | %s" % (name[:58], c["_total"], "  ".join("%s %d" % (k, c[k]) for k in keys)))
    path = os.path.join(ROOT, "profiles", "sass_%s.txt" % tag)
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
