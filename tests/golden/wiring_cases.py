"""The other registered wirings and directly constructed bodies of the reference (SURVEY.md section 8 rows a3, a6, a14,
a17, a18), as data + ONE runner that drives either implementation through the same public API:

  * ``tests/golden/make_golden.py wirings`` runs it on the UNMODIFIED reference (CPU, fp64) and writes
    ``tests/golden/wiring_<id>.npz``;
  * ``tests/test_gpu_wirings.py`` runs it on this repository's modules (cuda, kernels underneath) and compares;
  * ``tests/test_oracle_golden.py`` pins the oracle's functional restatements of the same wirings (CPU).

Both implementations expose the same import paths (``models.fusion_nets.factory_classes``,
``models.fpn.fusion3D2D.ModifiedUnet3D2D`` ...), which is the drop-in contract being tested.
Weights come from ``oracle.fill_like`` (deterministic per key), so a fixture only stores results.
"""
import contextlib
import importlib
import io

import numpy as np
import torch

SMALL = dict(B=2, S=8, H=64, W=32, S2=20, W2=48)
ONGRID = dict(B=2, S=8, H=64, W=32, S2=8, W2=32)              # crop 'oct': the 2-D image is on the en-face grid
H128 = dict(B=2, S=8, H=128, W=32, S2=8, W2=32)               # original=True: kernel-8 tail needs H/16 == 8

# kind 'factory': models.fusion_nets.factory_classes[name]() called with the batch dict
# kind 'body'   : a body class constructed directly (the 4-level bases are reachable no other way), called with tensors
CASES = [
    dict(id='fpn', kind='factory', name='FPN', crop='oct', shape=ONGRID, loss='mix'),
    dict(id='fpn_regression', kind='factory', name='FPNRegression', crop='oct', shape=ONGRID, loss='mse'),
    dict(id='fpn_classification', kind='factory', name='FPNClassification', crop='oct', shape=ONGRID, loss='nll', n_out=3),
    dict(id='hybrid_regression', kind='factory', name='FPNHybridFusionRegression', crop='relative_2d_max', shape=SMALL,
         loss='mse'),
    dict(id='fpn2d', kind='factory', name='FPN2D', crop='oct', shape=ONGRID, loss='mix'),
    dict(id='fpn2d_resized', kind='factory', name='FPN2D', crop='relative_2d', shape=SMALL, loss='mix'),
    dict(id='late_fusion_max', kind='factory', name='FPNLateFusion', crop='relative_2d_max', shape=SMALL, loss='mix'),
    dict(id='late_fusion_bilinear', kind='factory', name='FPNLateFusion', crop='relative_2d', shape=SMALL, loss='mix'),
    dict(id='late_fusion_regression', kind='factory', name='FPNLateFusionRegression', crop='oct', shape=ONGRID, loss='mse'),
    dict(id='body3d2d_4level', kind='body', module='models.fpn.fusion3D2D', cls='ModifiedUnet3D2D',
         kwargs=dict(interpolate='2d_max'), inputs=('oct', 'slo'), shape=SMALL, loss='mse'),
    dict(id='body3d2d_4level_add', kind='body', module='models.fpn.fusion3D2D', cls='ModifiedUnet3D2D',
         kwargs=dict(interpolate='2d', feature_fusion='add'), inputs=('oct', 'slo'), shape=SMALL, loss='mse'),
    dict(id='body3d2d_level5_add', kind='body', module='models.fpn.fusion3D2D', cls='ModifiedUnet3D2DLevel5',
         kwargs=dict(interpolate=None, feature_fusion='add'), inputs=('oct', 'slo'), shape=ONGRID, loss='mse'),
    dict(id='body3d_original', kind='body', module='models.fpn.unets3D', cls='ModifiedUnet3D', kwargs=dict(original=True),
         inputs=('oct',), shape=H128, loss='mse'),
    dict(id='body2d_4level', kind='body', module='models.fpn.unets2D', cls='ModifiedUnet2D', kwargs=dict(),
         inputs=('slo',), shape=SMALL, loss='mse'),
    dict(id='body2d_4level_features', kind='body', module='models.fpn.unets2D', cls='ModifiedUnet2D',
         kwargs=dict(output_features=True), inputs=('slo',), shape=SMALL, loss='mse'),
]
CASE_IDS = [c['id'] for c in CASES]
SEED_WEIGHTS, SEED_BATCH = 77, 5
FULL_GRAD_MAX_NUMEL = 2304                                      # gradients stored in full up to a 16->16 3x3 kernel


def case_by_id(cid):
    return next(c for c in CASES if c['id'] == cid)


def build(case, cfg):
    """Construct the model of ``case`` from whatever implementation is importable as ``models`` / ``config``."""
    cfg.crop, cfg.fusion_modality = case.get('crop', 'oct'), 'slo'
    cfg.number_of_outputs = case.get('n_out', 1)
    fusion_nets = importlib.import_module('models.fusion_nets')
    with contextlib.redirect_stdout(io.StringIO()):
        if case['kind'] == 'factory':
            return fusion_nets.factory_classes[case['name']]()
        ini = fusion_nets.FPNConfig().config                   # the architecture .ini (fusion_nets.py:21-26)
        return getattr(importlib.import_module(case['module']), case['cls'])(ini, **case['kwargs'])


def _target(shape, dtype, device):
    g = torch.Generator().manual_seed(99)
    return torch.randn(tuple(shape), generator=g).to(dtype).to(device)


def forward_loss(case, model, batch, loss_mod):
    """-> (output tensor, scalar loss); the same calls a user of the reference makes."""
    if case['kind'] == 'factory':
        out = model(batch)['prediction']
    else:
        args = {'oct': batch['image'].permute(0, 1, 2, 4, 3), 'slo': batch['slo'][:, :, :, 0, :]}
        out = model(*[args[k] for k in case['inputs']])
    if case['loss'] == 'mix':
        crit = loss_mod.Mix({'Dice': loss_mod.Dice_loss_jointv2('prediction', 'mask'),
                             'BCE': loss_mod.BCE_Lossv2('prediction', 'mask')})
        loss = crit(batch, {'prediction': out})[0]
    elif case['loss'] == 'mse':
        loss = ((out - _target(out.shape, out.dtype, out.device)) ** 2).mean()
    else:                                                       # 'nll' on the softmax of FPNClassification
        labels = torch.arange(out.shape[0], device=out.device) % out.shape[1]
        loss = -torch.log(out[torch.arange(out.shape[0], device=out.device), labels]).mean()
    return out, loss


def run(case, cfg, loss_mod, O, device='cpu', dtype=torch.float64):
    """Build, load the deterministic weights, forward, loss, backward.  Returns a dict of torch tensors / lists."""
    model = build(case, cfg)
    sd = O.fill_like(model.state_dict(), SEED_WEIGHTS, dtype)
    if dtype == torch.float64:
        model.double()
    model.load_state_dict(sd, strict=True)
    model.to(device).train()
    sh = case['shape']
    batch = O.synthetic_batch(sh['B'], sh['S'], sh['H'], sh['W'], sh['S2'], sh['W2'], seed=SEED_BATCH, dtype=dtype)
    batch = {k: v.to(device) for k, v in batch.items()}
    out, loss = forward_loss(case, model, batch, loss_mod)
    loss.backward()
    names = [k for k, _ in model.named_parameters()]
    grads = {k: p.grad for k, p in model.named_parameters()}
    return dict(model=model, names=names, trainable=[p.requires_grad for _, p in model.named_parameters()],
                out=out.detach(), loss=loss.detach(), grads=grads, state_keys=list(sd.keys()),
                buffers={k: v.detach() for k, v in model.state_dict().items() if 'running_' in k})


NPROJ = 8


def grad_projections(key, g):
    """<g, r_j> for NPROJ fixed standard-normal vectors r_j seeded by the parameter name: large gradients are pinned
    without storing them (E[(<g-g', r>)^2] = |g-g'|^2, so the projections estimate the error norm)."""
    h = 1469598103934665603
    for ch in key.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    gen = torch.Generator().manual_seed(h & 0x7FFFFFFF)
    r = torch.randn(NPROJ, g.numel(), generator=gen, dtype=torch.float64)
    return (r @ g.double().reshape(-1).cpu()).numpy()


def to_fixture(res):
    fx = {'names': np.array(res['names']), 'trainable': np.array(res['trainable']), 'state_keys': np.array(res['state_keys']),
          'out': res['out'].cpu().numpy(), 'loss': float(res['loss'])}
    gsum, gl2, gproj = [], [], []
    for k in res['names']:
        g = res['grads'][k]
        if g is None:
            gsum.append(0.0); gl2.append(-1.0)                  # -1: the reference produced no gradient (frozen / unused)
            gproj.append(np.zeros(NPROJ))
            continue
        g = g.double().cpu()
        gsum.append(g.sum().item()); gl2.append(g.norm().item())
        gproj.append(grad_projections(k, g))
        if g.numel() <= FULL_GRAD_MAX_NUMEL:
            fx['grad/' + k] = g.numpy().astype(np.float32)
    fx['grad_sum'], fx['grad_l2'], fx['grad_proj'] = np.array(gsum), np.array(gl2), np.stack(gproj)
    bk = sorted(res['buffers'])
    fx['buffer_keys'] = np.array(bk)
    fx['buffer_sum'] = np.array([res['buffers'][k].double().sum().item() for k in bk])
    return fx
