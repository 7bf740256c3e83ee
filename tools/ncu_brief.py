"""Brief of an .ncu-rep (`ncu --set full`): per captured launch the numbers the roofline discussion needs.
    python tools/ncu_brief.py report.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']
stalls = [h for h in H if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
for r in rows[2:]:
    d = dict(zip(H, r))
    print('=' * 100)
    for k in want:
        if k in d:
            print(f'{k:72s} {U[H.index(k)]:16s} {d[k][:90]}')
    tot = sum(float(d[s]) for s in stalls if d[s])
    top = sorted(((float(d[s]), s.replace('smsp__pcsamp_warps_issue_stalled_', '')) for s in stalls if d[s]), reverse=True)[:8]
    print('stall samples: ' + ', '.join(f'{n} {100 * v / tot:.1f}%' for v, n in top))
