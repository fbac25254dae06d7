"""Text summary of an ncu report (raw page metrics + top stall sites): python tools/ncu_summary.py rep.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f'{h:70s} {v} {u}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']]) for r in data)
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print('\nwarp stall samples:', tot)
print('  ' + ' '.join(f'{k[6:]}={100 * v / tot:.0f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print('top SASS sites:')
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:12]:
    s = int(r[ix['# Samples']])
    main = max(stalls, key=lambda st: int(r[ix[st]] or 0))
    print(f"  {100 * s / tot:5.1f}% {main:22s} exec={r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip()[:90]}")
