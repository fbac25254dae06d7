"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one step (between pack_volume launches)."""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))


def us(row):
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    return v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v


seq = [(r['Kernel Name'].split('(')[0].replace('void <unnamed>::', '').replace('<unnamed>::', ''), us(r)) for r in rows]
starts = [i for i, s in enumerate(seq) if 'pack_volume' in s[0]]
a = starts[0] if starts else 0
b = starts[1] if len(starts) > 1 else len(seq)
step = seq[a:b]
tot = sum(s[1] for s in step)
print(f'one step: {len(step)} launches, {tot:.1f} us (cold-cache, serialised)')
agg, cnt = collections.defaultdict(float), collections.Counter()
for n, t in step:
    agg[n[:64]] += t
    cnt[n[:64]] += 1
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:24]:
    print(f'{v:9.1f} us {100 * v / tot:5.1f}%  n={cnt[k]:4d}  {k}')
