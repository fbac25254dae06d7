"""Launch count per training step (eager), before / after the packed-weight arena is sealed."""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
from __graft_entry__ import import_mirror
cfg, fusion_nets, loss_mod, weight_init = import_mirror()
import ffpn
from ffpn.trainer import FusionTrainer
from oracle import fusion_fpn_oracle as O
torch.manual_seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    model = fusion_nets.factory_classes['FPNHybridFusion']()
model.apply(weight_init.weight_init)
model = model.cuda().train()
crit = loss_mod.Mix({'Dice': loss_mod.Dice_loss_jointv2('prediction', 'mask'), 'BCE': loss_mod.BCE_Lossv2('prediction', 'mask')})
dev = {k: v.cuda() for k, v in O.synthetic_batch(8, 32, 128, 128, 320, 128, seed=1234).items()}
tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
for i in range(4):
    n0 = ffpn.lib.launch_count(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = tr.step(dev)
    e1.record()
    torch.cuda.synchronize()
    print(f'step {i}: {ffpn.lib.launch_count(0) - n0} launches, loss {loss.item():.5f}, {e0.elapsed_time(e1):.2f} ms (eager)')
from ffpn import ops
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.weight_arena_pack(tr.flat_p)
e1.record()
torch.cuda.synchronize()
print(f'weight_arena_pack: {e0.elapsed_time(e1) * 100:.1f} us per call')
