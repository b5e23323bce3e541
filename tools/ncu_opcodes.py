#!/usr/bin/env python3
"""Opcode histogram of one kernel from an ncu report with source counters (--set full --import-source on):
executed warp instructions and stall samples per SASS opcode, plus the pipe each opcode issues on (B300_MICROARCH:
IMAD/FFMA/IDP on the fma pipe; IADD3/LOP3/SHF/PRMT/VIMNMX/VABSDIFF on the alu pipe).
usage: tools/ncu_opcodes.py report.ncu-rep kernel_regex [top]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    ops, samp = collections.Counter(), collections.Counter()
    lines = []
    tot = tsamp = 0
    nk = 0
    for row in rows:
        if row and row[0] == "Kernel Name":
            nk += 1
            if nk > 1:
                break           # first captured launch only
            continue
        if row and row[0] == "Address":
            hdr = row
            si, ie, sm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(row) <= ie:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", row[si])
        if not m:
            continue
        base = m.group(2).split(".")[0]
        n, s = int(row[ie]), int(row[sm])
        ops[base] += n
        samp[base] += s
        tot += n
        tsamp += s
        lines.append((s, n, row[si].strip()))
    print("kernel %s: %d warp instructions, %d samples" % (kern, tot, tsamp))
    for k, v in ops.most_common(top):
        print("%-12s %10d %5.1f%%   samples %5.1f%%" % (k, v, 100.0 * v / tot, 100.0 * samp[k] / max(tsamp, 1)))
    print("-- hottest instructions by samples")
    for s, n, src in sorted(lines, reverse=True)[:top]:
        print("%6d %9d  %s" % (s, n, src[:110]))


if __name__ == "__main__":
    main()
