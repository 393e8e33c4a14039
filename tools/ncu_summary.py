#!/usr/bin/env python3
"""Extract the metrics DESIGN.md / bench.py cite from an .ncu-rep (read on the CPU box) into a small CSV.
usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x.csv"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum']


def main(rep, out):
    raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv'], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'metric', 'unit', 'value'])
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            for h, u, v in zip(hdr, units, r):
                if h in KEYS or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
                    try:
                        if float(v.replace(',', '')) == 0 and 'stalled' in h:
                            continue
                    except ValueError:
                        pass
                    w.writerow([name, h, u, v])
    print('wrote', out)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
