"""Key metrics of every kernel in an .ncu-rep (raw page).  usage: python tools/ncu_keys.py report.ncu-rep"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('-----', r[hdr.index('Kernel Name')][:100])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f'  {w:75s} {r[i]} {units[i]}')
    st = sorted(((float(r[hdr.index(s)].replace(',', '') or 0), s) for s in stall), reverse=True)[:6]
    for v, s in st:
        print(f'  stall {s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):30s} {v:.2f}')
