"""Bitwise reproducibility of one forward+backward (same weights, same batch): run it several times with the branch
streams off and on and list the parameters whose gradient differs between any two runs.  A race between streams would
show up here as a difference between the 'streams on' runs and the serial ones."""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
from __graft_entry__ import import_mirror
cfg, fusion_nets, loss_mod, weight_init = import_mirror()
from ffpn.trainer import FusionTrainer
from oracle import fusion_fpn_oracle as O   # synthetic batch generator only

torch.manual_seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    model = fusion_nets.factory_classes['FPNHybridFusion']()
model.apply(weight_init.weight_init)
model = model.cuda().train()
crit = loss_mod.Mix({'Dice': loss_mod.Dice_loss_jointv2('prediction', 'mask'), 'BCE': loss_mod.BCE_Lossv2('prediction', 'mask')})
dev = {k: v.cuda() for k, v in O.synthetic_batch(8, 32, 128, 128, 320, 128, seed=1234).items()}
tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
names = [k for k, _ in model.named_parameters()]
sizes = [p.numel() for _, p in model.named_parameters()]
runs = []
for mode in sys.argv[1:] or ['0', '0', '1', '1', '0', '1']:
    os.environ['FFPN_STREAMS'] = mode
    tr.flat_g.zero_()
    loss = tr.forward_backward(dev)
    torch.cuda.synchronize()
    runs.append((mode, float(loss), tr.flat_g.clone()))
ref = runs[0][2]
for i, (mode, loss, g) in enumerate(runs):
    diff, off = [], 0
    for n, k in zip(names, sizes):
        if not torch.equal(g[off:off + k], ref[off:off + k]):
            d = (g[off:off + k] - ref[off:off + k]).abs().max().item() / (ref[off:off + k].abs().max().item() + 1e-30)
            diff.append((n, f'{d:.1e}'))
        off += k
    print(f'run {i} FFPN_STREAMS={mode} loss {loss!r} params differing from run 0: {len(diff)} {diff[:8]}')
