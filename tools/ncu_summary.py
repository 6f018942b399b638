#!/usr/bin/env python
"""Write a compact text summary of an .ncu-rep (raw page metrics that matter for the roofline) + SASS hot spots."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def main(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    print(f'# {rep}')
    for r in rows[2:]:
        print(f'kernel: {r[kn][:100]}')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w:90s} {r[i]:>16s} {units[i]}')
    sys.stdout.flush()
    subprocess.run([sys.executable, __file__.replace('ncu_summary.py', 'ncu_sass.py'), rep, '20'])


if __name__ == '__main__':
    main(sys.argv[1])
