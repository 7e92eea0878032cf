#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics the
roofline discussion needs.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'sm__inst_executed.sum',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']


def main():
    for path in sys.argv[1:]:
        txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
        names, units = rows[hdr], rows[hdr + 1]
        print(f'== {path}')
        for r in rows[hdr + 2:]:
            if len(r) != len(names):
                continue
            print('--', r[names.index('Kernel Name')][:70], 'grid', r[names.index('Grid Size')])
            for w in WANT:
                if w in names:
                    i = names.index(w)
                    print(f'  {w:82s} {r[i]:>16s} {units[i]}')
        if '--stall' in sys.argv:
            pass


if __name__ == '__main__':
    main()
