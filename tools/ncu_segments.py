#!/usr/bin/env python3
"""Executed warp instructions of one kernel split at its CTA barriers / mbarrier waits (first captured launch):
which phase of a multi-phase kernel the instructions go to.  usage: tools/ncu_segments.py report.ncu-rep kernel_regex"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, nk, L = None, 0, []
for row in rows:
    if row and row[0] == "Kernel Name":
        nk += 1
        if nk > 1:
            break
        continue
    if row and row[0] == "Address":
        hdr = row
        si, ie, te, sm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(row) <= te or not row[ie].isdigit():
        continue
    L.append((int(row[ie]), int(row[te]), int(row[sm]), row[si].strip()))
tot = sum(x[0] for x in L)
cur = [0, 0, 0, 0]
print("kernel %s: %d warp instructions" % (kern, tot))
for n, t, s, src in L + [(0, 0, 0, "END")]:
    cur[0] += n; cur[1] += t; cur[2] += s; cur[3] += 1
    if "BAR.SYNC" in src or "SYNCS.PHASECHK" in src or src == "END":
        print("%11d warp-inst %5.1f%%  threads/inst %4.1f  samples %5d  static %4d   up to %s" % (cur[0], 100.0 * cur[0] / tot, cur[1] / max(cur[0], 1), cur[2], cur[3], src[:44]))
        cur = [0, 0, 0, 0]
