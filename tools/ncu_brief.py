#!/usr/bin/env python3
"""Key metrics of every captured launch of an ncu report: tools/ncu_brief.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print('---', r[h.index('Kernel Name')][:60], r[h.index('Grid Size')] if 'Grid Size' in h else '')
    for w in want:
        if w in h:
            print('  %-75s %s' % (w, r[h.index(w)]))
    st = [(float(r[i]), n.split('issue_stalled_')[1].split('_per')[0]) for i, n in enumerate(h)
          if 'average_warps_issue_stalled' in n and n.endswith('per_issue_active.ratio') and r[i]]
    st.sort(reverse=True)
    print('  stalls/issue:', ', '.join('%s %.2f' % (n, v) for v, n in st[:7]))
