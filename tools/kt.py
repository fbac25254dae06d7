import json,sys
d=json.load(open(sys.argv[1])); print({k:round(d[k],3) for k in ("value","ms_per_step")}); print(' | '.join(f'{k["sec"]*1e6:.1f}' for k in d["kernels"]))
