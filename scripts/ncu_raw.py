"""Key raw metrics of every kernel in an ncu report: python scripts/ncu_raw.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_active.avg', 'sm__cycles_active.max', 'sm__cycles_active.min',
        'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
idx = [(w, h.index(w)) for w in want if w in h]
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:70])
    print("   ", {w.split('.')[0].replace('launch__', '').replace('smsp__', '').replace('sm__', ''): r[i][:12] + units[i][:6] for w, i in idx})
    st = [(float(r[i]), h[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
          for i in range(len(h)) if 'smsp__average_warps_issue_stalled' in h[i] and r[i] not in ('', 'n/a')]
    print("    stalls:", [(round(a, 2), b) for a, b in sorted(st, reverse=True)[:7]])
