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
loss = tr.forward_backward(dev)
torch.cuda.synchronize()
print('loss', loss.item(), 'grad finite', bool(torch.isfinite(tr.flat_g).all()), 'grad norm', float(tr.flat_g.norm()))
off = 0
bad = []
for k, p in model.named_parameters():
    g = tr.flat_g[off:off + p.numel()]
    if not torch.isfinite(g).all() or float(g.abs().max()) > 1e4:
        bad.append((k, float(g.abs().max()) if torch.isfinite(g).all() else 'nan'))
    off += p.numel()
print('bad grads', bad[:12], len(bad))
for k, b in model.named_buffers():
    if b.dtype.is_floating_point and not torch.isfinite(b).all():
        print('bad buffer', k)
