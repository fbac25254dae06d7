"""Generate the golden fixtures that pin ``oracle/fusion_fpn_oracle.py`` to the reference.

Runs ONLY in the build container (needs ``/root/reference``, read-only, imported unmodified with
the recipe of SURVEY.md section 8c).  The GPU box never runs this; it reads the committed ``.npz``.

    python tests/golden/make_golden.py            # fusion_*.npz, index_ops.npz, weight_init.npz
    python tests/golden/make_golden.py wirings    # wiring_<id>.npz (other factory entries / direct bodies)

Fixtures written next to this file:
  fusion_<crop>.npz   FPNHybridFusion fwd+bwd on a tiny batch with weights from
                      ``oracle.make_state_dict(seed)`` loaded ``strict=True`` into the reference:
                      prediction, loss, per-stage activation checksums, full gradients of a few small
                      tensors, (sum, l2) of every gradient, BN running-stat updates of a few layers.
  index_ops.npz       MaxPool3d / adaptive_max_pool3d argmax tables and Upsample_Custom3d_nearest
                      index tables from torch / the reference module, incl. ties and NaN.
  wiring_<id>.npz     the other seven factory entries and the directly constructed bodies (4-level, 'add' fusion,
                      original=True) of tests/golden/wiring_cases.py: output, loss, every gradient's (sum, l2),
                      small gradients in full, BN running-stat checksums.
  weight_init.npz     per-tensor (sum, abs-sum) after ``torch.manual_seed(1234)`` + construction +
                      ``weight_init`` (train.py:42,53-56) -- pins the mirror's RNG consumption order.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'


def import_reference(crop='relative_2d_max', modality='slo'):
    os.chdir(REF)                                    # the .ini is read cwd-relative (fusion_nets.py:24-26)
    sys.path.insert(0, REF)
    sys.argv = ['x', '--training-dataset', 'hrf_fusion', '--model', 'FPNHybridFusion',
                '--fusion-modality', modality, '--crop', crop]
    with contextlib.redirect_stdout(io.StringIO()):
        import config as cfg                          # noqa: F401  (argparse at import, config.py:76)
        from models.fusion_nets import factory_classes
        from common import loss, weight_init
    return cfg.config, factory_classes, loss, weight_init


def make_wirings(cfg, ref_loss, O):
    """wiring_<id>.npz: every case of tests/golden/wiring_cases.py run on the unmodified reference in fp64."""
    sys.path.insert(0, HERE)
    import wiring_cases as WC
    for case in WC.CASES:
        res = WC.run(case, cfg, ref_loss, O, device='cpu', dtype=torch.float64)
        fx = WC.to_fixture(res)
        np.savez_compressed(os.path.join(HERE, f"wiring_{case['id']}.npz"), **fx)
        print(case['id'], 'loss', fx['loss'], 'out', fx['out'].shape, 'params', len(fx['names']),
              'without grad', int((fx['grad_l2'] < 0).sum()))


def main():
    torch.set_num_threads(8)
    mode = sys.argv[1] if len(sys.argv) > 1 else 'all'          # read before import_reference() rewrites sys.argv
    cfg, factory, ref_loss, ref_init = import_reference()
    sys.path.insert(0, REPO)
    from oracle import fusion_fpn_oracle as O
    if mode == 'wirings':
        return make_wirings(cfg, ref_loss, O)

    shapes = dict(B=2, S=8, H=64, W=32, S2=20, W2=48)          # small; S%4==0, W%16==0, H>=64
    for crop in ('relative_2d_max', 'relative_2d', 'oct'):
        cfg.crop = crop
        with contextlib.redirect_stdout(io.StringIO()):
            model = factory['FPNHybridFusion']()
        sd = O.make_state_dict(seed=1234, dtype=torch.float64)
        assert list(model.state_dict().keys()) == list(sd.keys()), 'key order differs from reference'
        model.double()                                # fp64: fixtures are free of fp32 reduction-order noise
        model.load_state_dict(sd, strict=True)
        model.train()
        sh = dict(shapes)
        if crop == 'oct':
            sh['S2'], sh['W2'] = sh['S'], sh['W']
        batch = O.synthetic_batch(sh['B'], sh['S'], sh['H'], sh['W'], sh['S2'], sh['W2'], seed=7, dtype=torch.float64)
        acts = {}
        hooks = []
        body = model.resensnet
        for name in ['conv1', 'conv2', 'conv3', 'conv4', 'conv5', 'zdimRed1', 'zdimRed2', 'zdimRed3', 'zdimRed4',
                     'zdimRed5', 'conv1_2d', 'conv2_2d', 'conv3_2d', 'conv4_2d', 'conv5_2d', 'up_concat4',
                     'up_concat3', 'up_concat2', 'up_concat1', 'final1']:
            hooks.append(getattr(body, name).register_forward_hook(
                lambda m, i, o, name=name: acts.__setitem__(name, o.detach().clone())))
        out = model({k: v.clone() for k, v in batch.items()})
        crit = ref_loss.Mix({'Dice': ref_loss.Dice_loss_jointv2('prediction', 'mask'),
                             'BCE': ref_loss.BCE_Lossv2('prediction', 'mask')})
        loss, parts = crit(batch, out)
        loss.backward()
        for h in hooks:
            h.remove()
        fx = {'shape': np.array([sh[k] for k in ('B', 'S', 'H', 'W', 'S2', 'W2')]), 'seed_weights': 1234,
              'seed_batch': 7, 'prediction': out['prediction'].detach().numpy(), 'loss': loss.item(),
              'dice': parts['Dice'].item(), 'bce': parts['BCE'].item()}
        for k, a in acts.items():
            a64 = a.double()
            fx[f'act/{k}'] = np.array([a64.sum().item(), a64.abs().sum().item(), (a64 ** 2).sum().item()] + list(a.shape))
        small = ['resensnet.final1.weight', 'resensnet.final1.bias', 'resensnet.conv1.0.convBlock.0.0.weight',
                 'resensnet.conv1.0.convBlock.0.1.weight', 'resensnet.conv1.0.convBlock.0.1.bias',
                 'resensnet.conv1_2d.0.convBlock.0.0.weight', 'resensnet.zdimRed1.0.convBlock.0.0.weight',
                 'resensnet.zdimRed1.0.downsample.0.weight', 'resensnet.zdimRed1.1.convBlock.0.0.weight',
                 'resensnet.up_concat1.conv.convBlock.0.0.weight', 'resensnet.up_concat1.conv.downsample.0.weight',
                 'resensnet.conv1.1.convBlock.2.0.weight', 'resensnet.conv2.0.convBlock.0.0.weight']
        names, gsum, gl2 = [], [], []
        for k, p in model.named_parameters():
            g = p.grad.double()
            names.append(k)
            gsum.append(g.sum().item())
            gl2.append(g.norm().item())
            if k in small:
                fx[f'grad/{k}'] = p.grad.numpy()
        fx['grad_names'] = np.array(names)
        fx['grad_sum'] = np.array(gsum)
        fx['grad_l2'] = np.array(gl2)
        after = model.state_dict()
        for k in ['resensnet.conv1.0.convBlock.0.1', 'resensnet.zdimRed2.0.downsample.1',
                  'resensnet.up_concat3.conv.convBlock.1.1', 'resensnet.conv3_2d.1.convBlock.2.1']:
            for s in ('running_mean', 'running_var', 'num_batches_tracked'):
                fx[f'bn/{k}.{s}'] = after[f'{k}.{s}'].numpy()
        np.savez_compressed(os.path.join(HERE, f'fusion_{crop}.npz'), **fx)
        print(crop, 'loss', loss.item(), 'pred', tuple(out['prediction'].shape))

    # ---- index ops -------------------------------------------------------------------
    import torch.nn.functional as F
    from models.fpn.components import Upsample_Custom3d_nearest
    g = torch.Generator().manual_seed(3)
    ix = {}
    x = torch.randn(1, 1, 4, 6, 8, generator=g)
    x = (x * 2).round() / 2                                   # many ties
    x[0, 0, 1, 2, 3] = float('nan')
    x[0, 0, 3, 0, 0] = float('nan')
    x[0, 0, 3, 1, 1] = float('nan')                          # two NaNs in one window: the last is indexed
    for name, k in (('p122', (1, 2, 2)), ('p222', (2, 2, 2))):
        v, i = F.max_pool3d(x, k, return_indices=True)
        ix[f'{name}/x'] = x.numpy()
        ix[f'{name}/val'] = v.numpy()
        ix[f'{name}/idx'] = i.numpy()
    const = torch.ones(1, 1, 2, 4, 4)
    ix['const/idx'] = F.max_pool3d(const, (2, 2, 2), return_indices=True)[1].numpy()
    y = ((torch.randn(1, 1, 20, 48, 1, generator=g) * 2).round() / 2)
    for name, o in (('a8x32', (8, 32, 1)), ('a8x16', (8, 16, 1)), ('a3x7', (3, 7, 1))):
        v, i = F.adaptive_max_pool3d(y, o, return_indices=True)
        ix[f'{name}/x'] = y.numpy()
        ix[f'{name}/val'] = v.numpy()
        ix[f'{name}/idx'] = i.numpy()
    ar = torch.arange(7.).view(1, 1, 7, 1, 1)
    ix['ar7to3/idx'] = F.adaptive_max_pool3d(ar, (3, 1, 1), return_indices=True)[1].numpy()
    for f in ((2, 2, 1), (1, 2, 1)):
        src = torch.arange(4 * 6 * 1, dtype=torch.float32).view(1, 1, 4, 6, 1)
        ix[f'up{f[0]}{f[1]}{f[2]}/out'] = Upsample_Custom3d_nearest(scale_factor=f)(src).numpy()
    z = torch.rand(1, 2, 5, 7, 1, generator=g)
    ix['tri/x'] = z.numpy()
    ix['tri/out_4x16'] = F.interpolate(z, size=(4, 16, 1), mode='trilinear').numpy()
    ix['tri/out_8x5'] = F.interpolate(z, size=(8, 5, 1), mode='trilinear').numpy()
    np.savez_compressed(os.path.join(HERE, 'index_ops.npz'), **ix)

    # ---- weight_init -----------------------------------------------------------------
    cfg.crop = 'relative_2d_max'
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        model = factory['FPNHybridFusion']()
    model.apply(ref_init.weight_init)
    names, s1, s2 = [], [], []
    for k, v in model.state_dict().items():
        names.append(k)
        s1.append(v.double().sum().item())
        s2.append(v.double().abs().sum().item())
    np.savez_compressed(os.path.join(HERE, 'weight_init.npz'), names=np.array(names), sum=np.array(s1),
                        abssum=np.array(s2), shapes=np.array([str(tuple(v.shape)) for v in model.state_dict().values()]))
    print('weight_init fixture:', len(names), 'entries')


if __name__ == '__main__':
    main()
