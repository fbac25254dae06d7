"""A/B timing of the captured training step under different environment toggles (FFPN_STREAMS, FFPN_PDL, ...):
prints ms/step (CUDA events around K graph replays), the loss trajectory end point and an fp64 checksum of the
parameters after the run, so two settings can be compared for speed AND for bitwise-equal results.
usage: python tools/ab_step.py [steps] [batch] [S H W S2 W2]"""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
from __graft_entry__ import import_mirror
cfg, fusion_nets, loss_mod, weight_init = import_mirror()
from ffpn.trainer import FusionTrainer
from ffpn import lib
from oracle import fusion_fpn_oracle as O   # synthetic batch generator only

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
# optional shape: S H W S2 W2 (B-scans, depth, width, SLO rows, SLO cols); default = the C2 bench workload
S_, H_, W_, S2_, W2_ = [int(v) for v in sys.argv[3:8]] if len(sys.argv) >= 8 else (32, 128, 128, 320, 128)
torch.manual_seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    model = fusion_nets.factory_classes['FPNHybridFusion']()
model.apply(weight_init.weight_init)
model = model.cuda().train()
crit = loss_mod.Mix({'Dice': loss_mod.Dice_loss_jointv2('prediction', 'mask'), 'BCE': loss_mod.BCE_Lossv2('prediction', 'mask')})
dev = {k: v.cuda() for k, v in O.synthetic_batch(batch, S_, H_, W_, S2_, W2_, seed=1234).items()}
tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
tr.capture(dev, warmup=3)
tr.replay()
torch.cuda.synchronize()
p1 = tr.flat_p.double()
print(f'AB1 after 1 step: psum {float(p1.sum()):.12e} pl2 {float(p1.norm()):.12e} loss {float(tr._static_loss):.8f}')
for _ in range(4):
    tr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = tr.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
p = tr.flat_p.double()
toggles = {k: v for k, v in os.environ.items() if k.startswith('FFPN_')}
print(f'AB {toggles} B{batch} S{S_} H{H_} W{W_} slo {S2_}x{W2_} ms/step {ms:.3f} samples/s {batch / ms * 1e3:.1f} loss {float(loss):.6f} '
      f'psum {float(p.sum()):.10e} pl2 {float(p.norm()):.10e} mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB')
