"""Development aid: the metrics of an ncu report (--set full) that the roofline discussion uses, as text.
usage: python tools/ncu_summary.py <report.ncu-rep> [title]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'launch__block_size', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__icc_request_hit_rate.pct', 'gcc__average_cache_request_hit_rate.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum')
for row in rows[2:]:
    d = dict(zip(hdr, row))
    print(f"# {title}: {d.get('Kernel Name', '?')}")
    for h, u in zip(hdr, units):
        if h in KEEP or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            print(f'{h} [{u}] = {d[h]}')
