"""Turn the ncu captures of one round (gpurun_out/ncu/<case>.raw.csv + <case>.source.csv.gz, exported on the GPU box by
tools/ncu_capture.sh from `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <n> -c 1 python
tools/kernel_cases.py <case>`) into the committed evidence: profiles/<round>_ncu_<case>.txt (counters + top stall sites) and profiles/ncu_traffic.json
(per case: DRAM bytes per launch, duration, tensor-pipe and DRAM utilisation), which bench.py reads for `roofline.traffic`.

    python tools/ncu_collect.py r02
"""
import csv
import glob
import gzip
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else 'r02'
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(unit, 1)


traffic_path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
traffic = {}
if os.path.exists(traffic_path):
    with open(traffic_path) as f:
        traffic = {k: v for k, v in json.load(f).items() if isinstance(v, dict)}
for rep in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', 'ncu', '*.raw.csv'))):
    case = os.path.basename(rep)[:-8]
    raw = open(rep).read()
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print('skip (empty report):', rep)
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = [f'ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 python tools/kernel_cases.py {case}',
           f'Kernel Name: {m.get("Kernel Name", ("?", ""))[0]}']
    for h in WANT:
        if h in m:
            out.append(f'{h:80s} {m[h][0]} {m[h][1]}')
    try:
        src = gzip.open(rep[:-8] + '.source.csv.gz', 'rt').read()
        srows = list(csv.reader(io.StringIO(src)))
        shdr, data = srows[1], srows[2:]
        ix = {h: i for i, h in enumerate(shdr)}
        stalls = [h for h in shdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(int(r[ix['# Samples']]) for r in data)
        agg = {s_: sum(int(r[ix[s_]] or 0) for r in data) for s_ in stalls}
        out.append(f'\nwarp stall samples: {tot}')
        out.append('  ' + ' '.join(f'{k[6:]}={100 * v / max(tot, 1):.0f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        out.append('top SASS sites:')
        for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:10]:
            n = int(r[ix['# Samples']])
            main = max(stalls, key=lambda st: int(r[ix[st]] or 0))
            out.append(f"  {100 * n / max(tot, 1):5.1f}% {main:22s} exec={r[ix['Instructions Executed']]:>10s}  {r[ix['Source']].strip()[:90]}")
    except Exception as e:                                    # no source page in the report
        out.append(f'(no source page: {e})')
    name = f'{rnd}_ncu_{case}.txt'
    with open(os.path.join(ROOT, 'profiles', name), 'w') as f:
        f.write('\n'.join(out) + '\n')
    rd, wr = m.get('dram__bytes_read.sum'), m.get('dram__bytes_write.sum')
    traffic[case] = {
        'dram_bytes': int(to_bytes(*rd) + to_bytes(*wr)) if rd and wr else None,
        'duration_us': to_us(*m['gpu__time_duration.sum']) if 'gpu__time_duration.sum' in m else None,
        'tensor_pipe_pct': float(m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'][0]) if 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active' in m else None,
        'dram_pct': float(m['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0]) if 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed' in m else None,
        'kernel': m.get('Kernel Name', ('?', ''))[0][:120], 'file': 'profiles/' + name}
    print(case, traffic[case])
with open(traffic_path, 'w') as f:
    json.dump(traffic, f, indent=1, sort_keys=True)
