#!/usr/bin/env python3
"""Turns one gpurun profiling call (tools/gpu_round.sh) into the tracked files under profiles/:

  profiles/<tag>_launches.csv   the ncu launch list (gpu__time_duration.sum per launch), copied
  profiles/<tag>_summary.md     per-kernel shares, the --set full key metrics, SASS phase breakdown
  profiles/traffic.json         dram bytes per launch of each kernel (read by bench.py for roofline.traffic)

usage: tools/ncu_summary.py <tag> "<command that was profiled>"     (reads gpurun_out/<tag>_*)
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def short(name):
    name = re.sub(r"^void\s+", "", name.strip())
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    m = re.match(r"([A-Za-z0-9_:]+)", name)
    return m.group(1).split("::")[-1] if m else name[:40]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    tag, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    go = os.path.join(ROOT, "gpurun_out")
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    md = ["# %s -- ncu summary" % tag, "", "Command (1 GPU, B200): `%s`" % cmd,
          "Plain run first (exit 0), then `ncu --metrics gpu__time_duration.sum --clock-control none` (launch list, `profiles/%s_launches.csv`)"
          % tag, "and `ncu --set full --clock-control none --import-source on` on the hot kernels.", ""]
    # ---- launch list
    ll = os.path.join(go, tag + "_launches.csv")
    if os.path.exists(ll):
        lines = [l for l in open(ll) if l.startswith('"')]
        shutil.copy(ll, os.path.join(out, tag + "_launches.csv"))
        agg = collections.OrderedDict()
        for r in csv.DictReader(lines):
            if r["Metric Name"] != "gpu__time_duration.sum":
                continue
            a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
            a[0] += 1
            a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
        tot = sum(a[1] for a in agg.values())
        md += ["## Launch list (cold-cache, serialised; compare shares)", "", "| kernel | launches | total us | share | avg us |",
               "|---|---|---|---|---|"]
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            md.append("| %s | %d | %.1f | %.1f%% | %.1f |" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
        md.append("")
        step = {"pyramid": ("pyr_level0_kernel", "pyr_resize_kernel"), "fast": ("fast_cells_kernel",), "octree": ("octree_kernel",),
                "describe": ("orient_describe_kernel", "blur_levels_kernel", "describe_blurred_kernel")}
        stot = sum(agg[k][1] for ks in step.values() for k in ks if k in agg)
        if stot > 0:
            md += ["Shares inside the extraction step (what `bench.py` times; the matcher launches above belong to its separate",
                   "`matcher` block): " + ", ".join("%s %.1f%%" % (n, 100 * sum(agg[k][1] for k in ks if k in agg) / stot)
                                                    for n, ks in step.items()) + ".", ""]
            bj = os.path.join(go, tag + "_bench.json")
            if os.path.exists(bj):
                try:
                    st = json.loads(open(bj).read().strip().splitlines()[-1])["roofline"]["stage_ms_per_step"]
                    t2 = sum(st.values())
                    md += ["Live CUDA-event stage timers of the default 4096-frame bench of the same call: " +
                           ", ".join("%s %.1f%%" % (n, 100 * st[n] / t2) for n in step) + ".", ""]
                except Exception:
                    pass
    # ---- full capture
    rep = os.path.join(go, tag + "_prof.ncu-rep")
    traffic = {}
    tpath = os.path.join(out, "traffic.json")
    if os.path.exists(tpath):
        traffic = {k: v for k, v in json.load(open(tpath)).items() if v.get("source") != tag and k != "void"}
    if os.path.exists(rep):
        seen = {}
        for rp in (rep, os.path.join(go, tag + "_prof_match.ncu-rep")):
            if not os.path.exists(rp):
                continue
            raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(raw.splitlines()))
            hdr, units = rows[0], rows[1]
            for r in rows[2:]:
                k = short(r[hdr.index("Kernel Name")])
                key = (k, r[hdr.index("launch__grid_size")])
                seen.setdefault(key, (r, hdr, units))       # first launch of every (kernel, grid) pair
        md += ["## `--set full` (first captured launch per kernel and grid)", ""]
        for (k, grid), (r, hdr, units) in seen.items():
            md.append("### %s (grid %s)" % (k, grid))
            md.append("")
            for m in KEYS:
                if m in hdr:
                    md.append("- %s = %s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            e = traffic.setdefault(k, {"source": tag, "launches": []})
            if e.get("source") != tag:
                e.clear()
                e.update({"source": tag, "launches": []})
            def pct(name):
                return float(r[hdr.index(name)].replace(",", "")) if name in hdr else None
            e["launches"].append({"grid": int(grid), "dram_bytes": rd + wr,
                                  "us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")),
                                  "alu_pipe_pct": pct("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                                  "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                  "dram_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")})
            md.append("")
        # ---- SASS phase breakdown between barriers for the two big kernels
        for kname in ("fast_cells_kernel", "describe_blurred_kernel", "blur_levels_kernel", "pyr_resize_kernel"):
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname, "-c", "1"],
                                 capture_output=True, text=True).stdout
            srows = [r for r in csv.reader(src.splitlines()) if len(r) > 6 and r[0].startswith("0x")]
            if not srows:
                continue
            for i in range(1, len(srows)):          # the page may list several instances: keep the first
                if int(srows[i][0], 16) <= int(srows[i - 1][0], 16):
                    srows = srows[:i]
                    break
            seg, acc, ops = 0, collections.defaultdict(lambda: [0, 0, 0]), collections.defaultdict(collections.Counter)
            for r in srows:
                s = r[1].strip()
                acc[seg][0] += int(r[5]); acc[seg][1] += int(r[2]); acc[seg][2] += int(r[6])
                op = s.split()[1] if s.startswith("@") else s.split()[0]
                ops[seg][op.split(".")[0]] += int(r[5])
                if "BAR.SYNC" in s:
                    seg += 1
            ti, ts = sum(a[0] for a in acc.values()) or 1, sum(a[1] for a in acc.values()) or 1
            md += ["### %s: SASS between barriers (segment = code up to the n-th BAR.SYNC)" % kname, "",
                   "| segment | warp inst | stall samples | threads/inst | top opcodes |", "|---|---|---|---|---|"]
            for s_, a in acc.items():
                md.append("| %d | %.1f%% | %.1f%% | %.1f | %s |" % (s_, 100 * a[0] / ti, 100 * a[1] / ts, a[2] / max(a[0], 1),
                                                                   ", ".join("%s %.1f%%" % (o, 100 * c / ti) for o, c in ops[s_].most_common(5))))
            md.append("")
        json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    notes = os.path.join(out, tag + "_notes.md")
    if os.path.exists(notes):
        md += ["## Reading", ""] + open(notes).read().splitlines()
    open(os.path.join(out, tag + "_summary.md"), "w").write("\n".join(md) + "\n")
    print("wrote", os.path.join(out, tag + "_summary.md"))


if __name__ == "__main__":
    main()
