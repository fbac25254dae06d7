"""Which conv calls of one training step leave the warp-specialised tcgen05 kernels (FFPN_VERBOSE_FALLBACK lines,
de-duplicated with counts).  usage: python tools/route_check.py [batch S H W S2 W2]"""
import collections, contextlib, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get('ROUTE_CHILD'):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
    import torch
    from __graft_entry__ import import_mirror
    cfg, fusion_nets, loss_mod, weight_init = import_mirror()
    from ffpn.trainer import FusionTrainer
    from oracle import fusion_fpn_oracle as O   # synthetic batch generator only
    a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [8, 32, 128, 128, 320, 128]
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        model = fusion_nets.factory_classes['FPNHybridFusion']()
    model.apply(weight_init.weight_init)
    model = model.cuda().train()
    crit = loss_mod.Mix({'Dice': loss_mod.Dice_loss_jointv2('prediction', 'mask'), 'BCE': loss_mod.BCE_Lossv2('prediction', 'mask')})
    dev = {k: v.cuda() for k, v in O.synthetic_batch(*a, seed=1234).items()}
    tr = FusionTrainer(model, crit)
    tr.forward_backward(dev)
    torch.cuda.synchronize()
else:
    env = dict(os.environ, ROUTE_CHILD='1', FFPN_VERBOSE_FALLBACK='1')
    r = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    c = collections.Counter(l for l in r.stderr.splitlines() if '->' in l)
    print(f'{sum(c.values())} conv calls outside the warp-specialised kernels (rc {r.returncode})')
    for l, n in c.most_common():
        print(f'{n:3d} x {l}')
    if r.returncode:
        print(r.stderr[-2000:])
