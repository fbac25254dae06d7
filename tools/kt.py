"""One-line digest of a bench.py JSON line: python tools/kt.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: round(d[k], 3) for k in ('value', 'ms_per_step')}, 'e2e', round(d['e2e']['value'], 1), 'launches/step', d.get('gpu_launches', 0) // max(d['steps'], 1))
for k in d.get('kernels', []):
    print(f"  {k['kernel']:28s} {k['sec'] * 1e6:7.1f} us  {k['bound']:6s} hbm {k.get('frac_hbm', k['frac']):.3f}  tensor {k.get('frac_tensor', 0):.3f}")
