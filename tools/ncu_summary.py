"""Condense an ncu report (--set full, one kernel) into the handful of metrics profiles/*.csv keep.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-name-substring] > profiles/<tag>_ncu_full_<kernel>.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'launch__grid_size', 'launch__block_size',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
]


def main(path, only=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(names)}
    for vals in rows[2:]:
        if len(vals) < len(names) or (only and (only not in vals[col['Kernel Name']] or 'k_pack' in vals[col['Kernel Name']])):
            continue
        print('# kernel: %s' % vals[col['Kernel Name']][:120])
        for k in KEEP:
            if k in col:
                print('%s,%s,%s' % (k, vals[col[k]], units[col[k]]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
