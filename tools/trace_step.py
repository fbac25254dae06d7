"""Kernel timeline of ONE captured training step (torch.profiler / CUPTI; nsys is not in the image): writes
gpurun_out/trace_step.json = [[kernel, stream, start_us, dur_us], ...] and prints span, busy (union) time, summed kernel
time and the largest idle gaps.  Timings under the profiler are for attribution only, never bench numbers.
usage: python tools/trace_step.py [out.json] [batch]"""
import contextlib, io, json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
from torch.profiler import profile, ProfilerActivity
from __graft_entry__ import import_mirror
cfg, fusion_nets, loss_mod, weight_init = import_mirror()
from ffpn.trainer import FusionTrainer
from oracle import fusion_fpn_oracle as O   # synthetic batch generator only

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'trace_step.json')
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    model = fusion_nets.factory_classes['FPNHybridFusion']()
model.apply(weight_init.weight_init)
model = model.cuda().train()
crit = loss_mod.Mix({'Dice': loss_mod.Dice_loss_jointv2('prediction', 'mask'), 'BCE': loss_mod.BCE_Lossv2('prediction', 'mask')})
dev = {k: v.cuda() for k, v in O.synthetic_batch(batch, 32, 128, 128, 320, 128, seed=1234).items()}
tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
tr.capture(dev, warmup=3)
for _ in range(5):
    tr.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.replay()
    torch.cuda.synchronize()
tmp = tempfile.mktemp(suffix='.json')
prof.export_chrome_trace(tmp)
ev = json.load(open(tmp))['traceEvents']
ks = [e for e in ev if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy') and 'dur' in e]
ks.sort(key=lambda e: e['ts'])
t0 = ks[0]['ts']
rows = [[e['name'].replace('void ', '').replace('(anonymous namespace)::', '').split('(')[0][:60], e.get('args', {}).get('stream', -1),
         round(e['ts'] - t0, 2), round(e['dur'], 2)] for e in ks]
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(rows, open(out, 'w'))
span = max(r[2] + r[3] for r in rows)
busy, end, gaps = 0.0, 0.0, []
for i, r in enumerate(rows):
    s, e = r[2], r[2] + r[3]
    if s > end:
        gaps.append((s - end, i))
        busy += e - s
    elif e > end:
        busy += e - end
    end = max(end, e)
work = sum(r[3] for r in rows)
print(f'{len(rows)} kernels, span {span:.0f} us, busy {busy:.0f} us, idle {span - busy:.0f} us, summed kernel time {work:.0f} us, '
      f'streams {len(set(r[1] for r in rows))}')
gaps.sort(reverse=True)
for g, i in gaps[:10]:
    print(f'  idle {g:6.1f} us before #{i} {rows[i][0]} (after {rows[i - 1][0]})')
