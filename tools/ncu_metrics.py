#!/usr/bin/env python
"""Print the metrics of an ncu raw-page CSV that matter for the ROIAlign kernels (one row per captured launch).
usage: ncu -i X.ncu-rep --page raw --csv > X.csv ; python tools/ncu_metrics.py X.csv [extra substrings...]"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct', 'l1tex__data_bank_conflicts_pipe_lsu',
        'l1tex__data_pipe_lsu_wavefronts', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'per_issue_active',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sectors_op_atom.sum', 'launch__registers_per_thread', 'sm__throughput.avg.pct',
        'smsp__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fma', 'sm__inst_executed_pipe_alu',
        'l1tex__lsu_writeback_active', 'smsp__inst_executed_op_shared', 'smsp__inst_executed_op_global',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared', 'gpu__dram_throughput.avg.pct']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
extra = sys.argv[2:]
for vals in rows[2:]:
    print('==', vals[hdr.index('Kernel Name')][:90] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in WANT + extra):
            try:
                fv = float(v.replace(',', ''))
                if fv == 0:
                    continue
            except ValueError:
                pass
            print(f"  {h} [{u}] = {v}")
