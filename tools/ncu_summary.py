"""Key figures of an ncu --set full report (first kernel, or every kernel with --all).
    python tools/ncu_summary.py report.ncu-rep [--all]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_tensor.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for vals in rows[2:] if '--all' in sys.argv else rows[2:3]:
    print('==', vals[hdr.index('Kernel Name')][:110])
    for w in WANT:
        if w in hdr:
            print(f'  {w:72s} {vals[hdr.index(w)]}')
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and 'per_warp_active' in h and 'not_issued' not in h:
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if v > 0.25:
                print(f'     stall {h.split("warp_issue_stalled_")[1]:58s} {v}')
