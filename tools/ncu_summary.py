"""Summarise an ncu report (exported with `ncu -i X.ncu-rep --page raw --csv`): per kernel launch the duration, pipe
utilisations, DRAM traffic and the top stall reasons.  usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
want = [('time_us', 'gpu__time_duration.sum'), ('tensor%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('alu%', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
        ('fma%', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
        ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        ('lsu_wavefronts%', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
        ('smem_wavefronts%', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
        ('warps_active%', 'sm__warps_active.avg.pct_of_peak_sustained_active'), ('regs', 'launch__registers_per_thread'),
        ('dram_rd', 'dram__bytes_read.sum'), ('dram_wr', 'dram__bytes_write.sum'), ('inst', 'smsp__inst_executed.sum'),
        ('cycles', 'sm__cycles_elapsed.avg'), ('smem_bank_conflicts', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')]
units = rows[1]
for r in data:
    name = r[hdr.index('Kernel Name')]
    line = {}
    for short, full in want:
        if full in hdr:
            i = hdr.index(full)
            line[short] = r[i] + ((' ' + units[i]) if short.startswith('dram') else '')
    st = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
    vals = sorted([(float(r[hdr.index(h)] or 0), h.split('stalled_')[1].split('_per')[0]) for h in st], reverse=True)[:6]
    print(name)
    print('   ', line)
    print('    stalls (cycles per issue):', ', '.join('%s %.2f' % (n, v) for v, n in vals))
