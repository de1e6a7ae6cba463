#!/usr/bin/env python
"""profiles/ncu_full_r2_roi_summary.csv + profiles/traffic_r2.json from an `ncu --set full` report of the two ROIAlign
kernels.  usage: ncu -i X.ncu-rep --page raw --csv > X.csv ; python tools/ncu_roi_summary.py X.csv "<capture id>" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
H, U = rows[0], rows[1]
EXACT = {'Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
         'smsp__inst_executed.sum', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
         'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum',
         'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_srcunit_tex_op_red.sum',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
         'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
         'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
         'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
         'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
         'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
         'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'}
keep = [i for i, h in enumerate(H) if h in EXACT or
        (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'))]
out = [[H[i] for i in keep], [U[i] for i in keep]] + [[r[i] for i in keep] for r in rows[2:]]
csv.writer(open(os.path.join(ROOT, 'profiles', 'ncu_full_r2_roi_summary.csv'), 'w')).writerows(out)
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
tr = {"capture": sys.argv[2] if len(sys.argv) > 2 else "profiles/ncu_full_r2_roi_summary.csv",
      "algorithmic_bytes": 6733725696, "kernels": {}}
for r in rows[2:]:
    d = dict(zip(H, r))
    name = 'roi_align_fwd_pr' if 'fwd_pr' in d['Kernel Name'] else 'roi_align_bwd_rw'
    g = lambda k: float(d[k].replace(',', ''))
    rd = g('dram__bytes_read.sum') * scale[U[H.index('dram__bytes_read.sum')]]
    wr = g('dram__bytes_write.sum') * scale[U[H.index('dram__bytes_write.sum')]]
    tr['kernels'][name] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                           "ms_under_ncu": g('gpu__time_duration.sum'),
                           "warp_instructions": g('smsp__inst_executed.sum'),
                           "issue_active_pct": g('smsp__issue_active.avg.pct_of_peak_sustained_active')}
    print(name, tr['kernels'][name])
json.dump(tr, open(os.path.join(ROOT, 'profiles', 'traffic_r2.json'), 'w'), indent=1)
