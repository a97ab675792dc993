#!/usr/bin/env python
"""Print the judged subset of an `ncu --page raw --csv` export (one block per launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; units = rows[1]
idx = {h:i for i,h in enumerate(hdr)}
want = ['Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum','smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    if len(r) != len(hdr): continue
    print('----')
    for w in want:
        if w in idx: print(f"  {w:75s} {r[idx[w]][:70]} {units[idx[w]]}")
    st = [(float(r[i].replace(',','')), h) for h,i in idx.items() if h.startswith('smsp__average_warps_issue_stalled') and r[i] not in ('', 'n/a')]
    for v,h in sorted(st, reverse=True)[:6]: print(f"   stall {v:8.2f} {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}")
