"""Stage- and model-level parity on the B200: the reference-shaped modules (CUDA kernels underneath) against
the CPU oracle and the committed golden fixtures, same weights, same seeded inputs.

Tolerances: fp32 path -- prediction rel-L2 <= 1e-4, gradients <= 3x the fp32-vs-fp64 floor of
torch itself (we use 1e-2 global / 3e-2 for conv1-4 stages whose ReLU/pool decisions flip); bf16 path -- loss
within 2e-2 of fp32, prediction rel-L2 <= 0.15 (PyTorch's own bf16 autocast is at 7.5e-2 on this model)."""
import os

import numpy as np
import pytest
import torch

from oracle import fusion_fpn_oracle as O

pytestmark = pytest.mark.gpu
CROPS = ['relative_2d_max', 'relative_2d', 'oct']


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(autouse=True)
def _exact_mode():
    import gc
    import ffpn
    gc.collect()                                       # trainers of earlier tests release the packed-weight arena of the device
    ffpn.set_compute_dtype(torch.float32)
    yield
    ffpn.set_compute_dtype(torch.bfloat16)


def _loss(mirror, batch, out):
    crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
    return crit(batch, out)[0]


def _run(mirror, sd, batch, crop, train=True, acts=None):
    model = mirror.build('FPNHybridFusion', crop).cuda()
    model.load_state_dict(sd, strict=True)
    model.train(train)
    cb = {k: v.cuda() for k, v in batch.items()}
    hooks = []
    if acts is not None:
        import stagehooks
        from ffpn import functional as FF
        a, hooks = stagehooks.attach(model.resensnet)
        FF.fuse_head_activation(False)                  # the hook on final1 must see logits, like the reference's
    try:
        out = model(cb)
    finally:
        if acts is not None:
            FF.fuse_head_activation(True)
    for h in hooks:
        h.remove()
    if acts is not None:
        acts.update(a)
    return model, cb, out


@pytest.mark.parametrize('crop', CROPS)
def test_model_matches_golden_and_oracle_fp32(mirror, golden_dir, crop):
    fx = np.load(os.path.join(golden_dir, f'fusion_{crop}.npz'))
    B, S, H, W, S2, W2 = [int(v) for v in fx['shape']]
    sd = O.make_state_dict(seed=int(fx['seed_weights']))
    batch = O.synthetic_batch(B, S, H, W, S2, W2, seed=int(fx['seed_batch']))
    acts = {}
    model, cb, out = _run(mirror, sd, batch, crop, acts=acts)
    pred = out['prediction']
    assert tuple(pred.shape) == (B, 1, S, 1, W) and pred.dtype == torch.float32
    # per-stage activation checksums of the reference (sum, abs-sum, square-sum, shape): SURVEY.md App. D.2
    n_checked = 0
    for k in fx.files:
        if not k.startswith('act/'):
            continue
        name, ref = k[4:], fx[k]
        ours = {'zdimRed': 'proj', 'up_concat': 'up'}
        key = name
        for a, b in ours.items():
            key = key.replace(a, b)
        t = acts[key].double()
        if name.startswith('zdimRed'):
            # the depth mean is fused behind the projection: sum(mean) * depth == sum of the reference's pre-mean tensor (>= 0)
            hrem = int(ref[-1])
            assert list(t.shape) == [int(v) for v in ref[3:-1]] + [1]
            assert abs(t.sum().item() * hrem - ref[0]) <= 1e-4 * abs(ref[1]), (name, t.sum().item() * hrem, ref[0])
        else:
            assert list(t.shape) == [int(v) for v in ref[3:]], (name, tuple(t.shape))
            assert abs(t.sum().item() - ref[0]) <= 1e-4 * ref[1], (name, t.sum().item(), ref[0])
            assert abs(t.abs().sum().item() - ref[1]) <= 1e-4 * ref[1], name
            assert abs((t ** 2).sum().item() - ref[2]) <= 2e-4 * ref[2], name
        n_checked += 1
    assert n_checked == 20
    # golden = the unmodified reference in fp64
    assert rel(pred.cpu(), torch.from_numpy(fx['prediction'])) <= 1e-4
    loss = _loss(mirror, cb, out)
    assert abs(loss.item() - float(fx['loss'])) <= 1e-4
    loss.backward()
    names = [str(n) for n in fx['grad_names']]
    grads = dict((k, p.grad) for k, p in model.named_parameters())
    assert list(grads) == names and all(g is not None for g in grads.values())
    l2 = np.array([grads[n].double().norm().item() for n in names])
    tot = np.sqrt((fx['grad_l2'] ** 2).sum())
    assert abs(np.sqrt((l2 ** 2).sum()) - tot) <= 1e-2 * tot
    for k in fx.files:
        if k.startswith('grad/'):
            g, r = grads[k[5:]].cpu().numpy(), fx[k]
            denom = max(np.linalg.norm(r), 1e-3 * tot)
            assert np.linalg.norm(g - r) <= 3e-2 * denom, (k, np.linalg.norm(g - r), denom)
    # BN running statistics and counters
    after = model.state_dict()
    for k in fx.files:
        if k.startswith('bn/'):
            np.testing.assert_allclose(after[k[3:]].cpu().numpy(), fx[k], rtol=1e-4, atol=1e-5, err_msg=k)


def test_all_gradients_against_oracle_fp32(mirror):
    sd = O.make_state_dict(seed=21)
    batch = O.synthetic_batch(2, 8, 64, 32, 20, 48, seed=9, smooth=True)
    model, cb, out = _run(mirror, sd, batch, 'relative_2d_max')
    _loss(mirror, cb, out).backward()
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    loss64, pred64, g64 = O.loss_and_grads(sd64, {k: v.double() for k, v in batch.items()})
    assert rel(out['prediction'].detach().cpu(), pred64) <= 1e-4
    num = den = 0.0
    per_stage = {}
    for k, p in model.named_parameters():
        d = (p.grad.double().cpu() - g64[k]).norm().item() ** 2
        n = g64[k].norm().item() ** 2
        num, den = num + d, den + n
        st = k.split('.')[1]
        a = per_stage.setdefault(st, [0.0, 0.0])
        a[0] += d
        a[1] += n
    assert (num / den) ** 0.5 <= 1e-2, (num / den) ** 0.5
    for st, (d, n) in per_stage.items():
        assert (d / max(n, 1e-30)) ** 0.5 <= 3e-2, (st, (d / n) ** 0.5)


def test_eval_mode_uses_running_stats(mirror):
    sd = O.make_state_dict(seed=4, randomize_running=True)
    batch = O.synthetic_batch(1, 4, 64, 16, 4, 16, seed=2)
    model, cb, out = _run(mirror, sd, batch, 'oct', train=False)
    ref = O.fpn_hybrid_fusion_forward(sd, batch, 'oct', train=False)['prediction']
    assert rel(out['prediction'].cpu(), ref) <= 1e-4
    after = model.state_dict()
    assert all(torch.equal(after[k].cpu(), sd[k]) for k in sd)          # eval must not touch the buffers


def test_odd_depth_496(mirror):
    """Full axial depth H=496 (not a multiple of 16): pools floor, projection halves with ceil (C3 shape)."""
    sd = O.make_state_dict(seed=8)
    batch = O.synthetic_batch(1, 4, 496, 16, 8, 32, seed=3)
    model, cb, out = _run(mirror, sd, batch, 'relative_2d_max')
    ref = O.fpn_hybrid_fusion_forward(sd, batch, 'relative_2d_max')['prediction']
    assert rel(out['prediction'].cpu(), ref) <= 1e-4
    _loss(mirror, cb, out).backward()


def test_bf16_path_close_to_fp32(mirror):
    import ffpn
    sd = O.make_state_dict(seed=21)
    batch = O.synthetic_batch(2, 8, 64, 32, 20, 48, seed=9, smooth=True)
    loss64, pred64, g64 = O.loss_and_grads({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()},
                                           {k: v.double() for k, v in batch.items()})
    ffpn.set_compute_dtype(torch.bfloat16)
    model, cb, out = _run(mirror, sd, batch, 'relative_2d_max')
    loss = _loss(mirror, cb, out)
    loss.backward()
    assert abs(loss.item() - loss64.item()) <= 2e-2
    assert rel(out['prediction'].detach().cpu(), pred64) <= 0.15
    dot = sum((p.grad.double().cpu() * g64[k]).sum().item() for k, p in model.named_parameters())
    na = sum(p.grad.double().norm().item() ** 2 for p in model.parameters()) ** 0.5
    nb = sum(g.norm().item() ** 2 for g in g64.values()) ** 0.5
    assert dot / (na * nb) >= 0.4, dot / (na * nb)       # PyTorch bf16 autocast: 0.45


@pytest.mark.parametrize('name', ['FPN', 'FPN2D', 'FPNLateFusion'])
def test_other_wirings_forward(mirror, name):
    import torch.nn.functional as F
    batch = O.synthetic_batch(1, 8, 64, 32, 8, 32, seed=5)
    model = mirror.build(name, 'oct').cuda().train()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    out = model({k: v.cuda() for k, v in batch.items()})['prediction']
    oct = batch['image'].permute(0, 1, 2, 4, 3)
    slo = batch['slo'][:, :, :, 0, :]
    if name == 'FPN':
        ref = torch.sigmoid(O.unet3d_body_forward(sd, oct).permute(0, 1, 2, 4, 3))
    elif name == 'FPN2D':
        ref = torch.sigmoid(O.unet2d_body_forward(sd, slo).permute(0, 1, 2, 4, 3))
    else:
        a = O.unet3d_body_forward(sd, oct, prefix='resensnet3d', use_1x1=False)
        b = O.unet2d_body_forward(sd, slo, prefix='resensnet2d', output_features=True)
        ref = torch.sigmoid(F.conv3d(torch.cat([a, b], 1), sd['fusion_module.weight'], sd['fusion_module.bias'])
                            .permute(0, 1, 2, 4, 3))
    assert out.shape == ref.shape and rel(out.detach().cpu(), ref) <= 1e-4
    out.sum().backward()
    assert all(p.grad is not None for k, p in model.named_parameters() if 'final1' not in k or name != 'FPNLateFusion')


def test_training_wrapper_step(mirror):
    model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
    crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    wrap = mirror.wrapper.Model(model, crit, None, None, None, [opt])
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 16, 8, 32, seed=1).items()}
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = wrap.training_step(batch, 0)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and wrap.configure_optimizers() == [opt]
    assert 'Training/Dice' in wrap.logged and len(wrap.logged['Training/BCE']) == 3
    assert int(model.state_dict()['resensnet.conv1.0.convBlock.0.1.num_batches_tracked']) == 3


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_trainer_gradient_sink_and_fused_sgd(mirror, dtype):
    """FusionTrainer (backward kernels writing straight into the flat gradient buffer, fused SGD kernel) against the
    plain autograd path + torch.optim.SGD on the same weights and batch (train.py:126-133 semantics)."""
    import ffpn
    from ffpn.trainer import FusionTrainer
    ffpn.set_compute_dtype(dtype)
    sd = O.make_state_dict(seed=5)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 32, 8, 32, seed=3).items()}
    crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
    ref = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
    ref.load_state_dict(sd, strict=True)
    ref.train()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    loss_ref, _ = crit(batch, ref(batch))
    loss_ref.backward()
    g_ref = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
    opt.step()

    model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
    model.load_state_dict(sd, strict=True)
    model.train()
    tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
    assert len(tr._sink) > 200                                   # conv weights and BatchNorm affine parameters
    loss = tr.forward_backward(batch)
    # fp32: same kernels, same data.  bf16: also deterministic since the last float atomics were removed (see
    # test_branch_streams_are_bitwise_equal_to_serial); the looser bound is kept for the eager-autograd reference path
    assert abs(loss.item() - loss_ref.item()) <= (1e-6 if dtype == torch.float32 else 5e-3) * max(1.0, abs(loss_ref.item()))
    # same kernels on the same data: the sink path must reproduce autograd's gradients up to the summation-order noise
    # of the kernels that still accumulate with float atomics (CUDA-core fp32 path, stems, strided projection convs);
    # the warp-specialised tcgen05 wgrad reduces its partial tiles in a fixed order and is bitwise reproducible
    if dtype == torch.float32:
        for k, p in model.named_parameters():
            assert rel(p.grad, g_ref[k]) <= 2e-4, (k, rel(p.grad, g_ref[k]))
    else:
        fa = torch.cat([p.grad.reshape(-1) for _, p in model.named_parameters()]).double()
        fb = torch.cat([g_ref[k].reshape(-1) for k, _ in model.named_parameters()]).double()
        assert float((fa * fb).sum() / (fa.norm() * fb.norm())) >= 0.995            # run-to-run bf16 noise only
        assert all(float(p.grad.abs().sum()) > 0 for k, p in model.named_parameters() if p.dim() > 1)
    tr.optimizer_step()
    if dtype == torch.float32:
        for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            assert rel(p.data, q.data) <= 1e-5, k
    assert float(tr.flat_g.abs().max()) == 0.0
    tr.close()


def _trainer_grads(mirror, sd, batch, env):
    """One FusionTrainer forward+backward under the given environment toggles -> (loss, flat gradient copy)."""
    from ffpn.trainer import FusionTrainer
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                                'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
        model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
        model.load_state_dict(sd, strict=True)
        model.train()
        tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
        loss = tr.forward_backward(batch)
        torch.cuda.synchronize()
        out = (loss.item(), tr.flat_g.clone(), {k: b.clone() for k, b in model.named_buffers()})
        tr.close()
        return out
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_branch_streams_are_bitwise_equal_to_serial(mirror):
    """The forward/backward DAG on side streams (2-D encoder, per-level projections, weight gradients) must give the
    SAME BITS as the single-stream order -- every kernel is deterministic, so any difference is a race between streams.
    bf16 path (the tcgen05 kernels), a shape with all five levels and several tiles per CTA."""
    import ffpn
    ffpn.set_compute_dtype(torch.bfloat16)
    sd = O.make_state_dict(seed=11)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 8, 64, 64, 40, 64, seed=4).items()}
    l0, g0, b0 = _trainer_grads(mirror, sd, batch, {'FFPN_STREAMS': '0'})
    for mode in ('1', 'branches', 'wgrad', '0', '1'):
        l1, g1, b1 = _trainer_grads(mirror, sd, batch, {'FFPN_STREAMS': mode})
        assert l1 == l0, (mode, l1, l0)
        assert torch.equal(g1, g0), (mode, float((g1 - g0).abs().max()))
        for k in b0:
            assert torch.equal(b1[k], b0[k]), (mode, k)             # BatchNorm running statistics too


def test_captured_graph_step_matches_eager_bitwise(mirror):
    """FusionTrainer.capture()/replay() (multi-stream CUDA graph with the fused SGD) against eager step() on a twin model:
    identical parameters after three optimisation steps."""
    import ffpn
    from ffpn.trainer import FusionTrainer
    ffpn.set_compute_dtype(torch.bfloat16)
    sd = O.make_state_dict(seed=12)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 32, 16, 32, seed=6).items()}
    crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
    params, buffers = [], []
    for graph in (False, True):
        model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
        model.load_state_dict(sd, strict=True)
        model.train()
        tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
        if graph:
            tr.capture(batch, warmup=2)
            for _ in range(3):
                tr.replay()
        else:
            for _ in range(3):
                tr.step(batch)
        torch.cuda.synchronize()
        params.append(tr.flat_p.clone())
        buffers.append({k: b.clone() for k, b in model.named_buffers()})
        tr.close()
    assert torch.equal(params[0], params[1]), float((params[0] - params[1]).abs().max())
    # the capture's warm-up passes must not leave a trace in the BatchNorm running statistics / counters
    for k in buffers[0]:
        assert torch.equal(buffers[0][k], buffers[1][k]), k
    assert int(buffers[1]['resensnet.conv1.0.convBlock.0.1.num_batches_tracked']) == 3


def test_prefetched_replay_equals_direct_replay(mirror):
    """FusionTrainer.prefetch() (copy stream + staging buffers) followed by replay(prefetched=True) must feed the captured step
    exactly the batch that replay(batch) would: two different batches, identical parameters afterwards."""
    import ffpn
    from ffpn.trainer import FusionTrainer
    ffpn.set_compute_dtype(torch.bfloat16)
    sd = O.make_state_dict(seed=13)
    hosts = [{k: v.pin_memory() for k, v in O.synthetic_batch(2, 4, 64, 32, 16, 32, seed=s).items()} for s in (7, 8)]
    crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
    params, losses = [], []
    for pre in (False, True):
        model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
        model.load_state_dict(sd, strict=True)
        model.train()
        tr = FusionTrainer(model, crit, lr=0.1, momentum=0.9, weight_decay=1e-4)
        tr.capture({k: v.cuda() for k, v in hosts[0].items()}, warmup=2)
        ls = []
        if pre:
            tr.prefetch(hosts[0])
            for i in range(2):
                loss = tr.replay(prefetched=True)
                if i == 0:
                    tr.prefetch(hosts[1])
                ls.append(loss.item())
        else:
            for i in range(2):
                ls.append(tr.replay(hosts[i]).item())
        torch.cuda.synchronize()
        params.append(tr.flat_p.clone())
        losses.append(ls)
        tr.close()
    assert losses[0] == losses[1] and losses[0][0] != losses[0][1]
    assert torch.equal(params[0], params[1])


def _crit(mirror):
    return mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})


def test_eval_after_training_sees_the_updated_running_statistics(mirror):
    """The eval-mode BatchNorm coefficients are cached per module; the library's kernels (and CUDA-graph replays) update the
    running statistics through raw pointers, which torch's version counters do not see -- an eval forward after an eager step
    and after a graph replay must nevertheless equal that of a fresh model loaded with the current state_dict."""
    import ffpn
    from ffpn.trainer import FusionTrainer
    ffpn.set_compute_dtype(torch.bfloat16)
    sd = O.make_state_dict(seed=33)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 32, 16, 32, seed=7).items()}

    def evaluate(m):
        m.eval()
        with torch.no_grad():
            out = m(batch)['prediction'].clone()
        m.train()
        return out

    def fresh_copy(m):
        f = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
        f.load_state_dict({k: v.clone() for k, v in m.state_dict().items()}, strict=True)
        return f

    model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
    model.load_state_dict(sd, strict=True)
    model.train()
    first = evaluate(model)                                      # fills the cache
    tr = FusionTrainer(model, _crit(mirror))
    tr.step(batch)
    torch.cuda.synchronize()
    after_eager = evaluate(model)
    assert not torch.equal(after_eager, first)
    assert torch.equal(after_eager, evaluate(fresh_copy(model)))
    tr.capture(batch, warmup=1)
    evaluate(model)                                              # cache again, then replay behind it
    tr.replay()
    torch.cuda.synchronize()
    assert torch.equal(evaluate(model), evaluate(fresh_copy(model)))
    tr.close()


def test_weights_loaded_behind_the_trainer_are_never_stale(mirror):
    """The packed bf16 weight arena (recorded by the trainer's first step, baked into its CUDA graph) must not serve stale
    images: (a) conv calls made outside FusionTrainer.forward_backward -- an eval forward of the same model -- pack from the
    current fp32 weights; (b) load_state_dict while a trainer is alive marks the arena dirty, so the next eager step or graph
    replay runs on the loaded weights."""
    import ffpn
    from ffpn.trainer import FusionTrainer
    ffpn.set_compute_dtype(torch.bfloat16)
    sd_a, sd_b = O.make_state_dict(seed=31), O.make_state_dict(seed=32)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 32, 16, 32, seed=6).items()}

    def fresh(sd, train):
        m = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
        m.load_state_dict(sd, strict=True)
        m.train(train)
        return m

    with torch.no_grad():
        want_eval_b = fresh(sd_b, False)(batch)['prediction'].clone()
    ref_tr = FusionTrainer(fresh(sd_b, True), _crit(mirror))
    want_loss_b = ref_tr.step(batch).item()
    torch.cuda.synchronize()
    want_params_b = ref_tr.flat_p.clone()
    ref_tr.close()

    model = fresh(sd_a, True)
    tr = FusionTrainer(model, _crit(mirror))
    tr.capture(batch, warmup=2)                                  # arena sealed with A's images; the graph reads the arena
    assert tr._arena_state == 2
    model.load_state_dict(sd_b, strict=True)                     # in place: same parameter pointers
    model.eval()
    with torch.no_grad():
        got_eval = model(batch)['prediction']
    assert torch.equal(got_eval, want_eval_b)                    # (a) an eval forward outside the trainer bypasses the arena
    model.train()
    loss = tr.replay()                                           # (b) the replay repacks first, then steps on B's weights
    torch.cuda.synchronize()
    assert loss.item() == want_loss_b
    assert torch.equal(tr.flat_p, want_params_b)
    tr.close()


def test_accumulate_grad_batches_sums_micro_batches(mirror):
    """train.py:161 accumulate_grad_batches=k: the optimiser sees the MEAN of k micro-batch gradients (each with its own
    BatchNorm statistics), then one SGD step.  Checked in the fp32 exact mode against plain autograd on a twin model."""
    from ffpn.trainer import FusionTrainer
    sd = O.make_state_dict(seed=41)
    micro = [{k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 32, 16, 32, seed=s).items()} for s in (1, 2)]
    ref = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
    ref.load_state_dict(sd, strict=True)
    ref.train()
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    for b in micro:
        (_crit(mirror)(b, ref(b))[0] / 2).backward()
    g_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()]).clone()
    opt.step()
    model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
    model.load_state_dict(sd, strict=True)
    model.train()
    tr = FusionTrainer(model, _crit(mirror), lr=0.1, momentum=0.9, weight_decay=1e-4, accumulate_grad_batches=2)
    assert not tr._sink
    tr.step(micro[0])
    assert tr.steps == 0 and float(tr.flat_g.abs().max()) > 0    # no optimiser step after the first micro-batch
    g_half = tr.flat_g.clone()
    tr.forward_backward(micro[1])
    assert rel(tr.flat_g, g_ref) <= 1e-5 and rel(g_half, g_ref) > 1e-2
    tr.micro += 1
    tr.optimizer_step()
    assert tr.steps == 1 and float(tr.flat_g.abs().max()) == 0.0
    for (k, p), q in zip(model.named_parameters(), ref.parameters()):
        assert rel(p.data, q.data) <= 1e-5, k
    tr.close()


def test_eval_mode_backward_is_frozen_batchnorm(mirror):
    """Gradients under model.eval() (frozen-BatchNorm fine-tuning, saliency): BatchNorm uses the running statistics in
    forward AND backward (dy = g * gamma * invstd, no batch-statistics terms), like torch's eval-mode
    native_batch_norm_backward.  Against the oracle's autograd in eval mode."""
    sd = O.make_state_dict(seed=4, randomize_running=True)
    batch = O.synthetic_batch(2, 4, 64, 16, 4, 16, seed=2)
    model, cb, out = _run(mirror, sd, batch, 'oct', train=False)
    _loss(mirror, cb, out).backward()
    work = {k: (v.clone().requires_grad_(True) if k in set(O.param_keys(sd)) else v) for k, v in sd.items()}
    pred = O.fpn_hybrid_fusion_forward(work, batch, 'oct', train=False)['prediction']
    keys = O.param_keys(sd)
    grads = dict(zip(keys, torch.autograd.grad(O.mix_loss(pred, batch['mask']), [work[k] for k in keys])))
    assert rel(out['prediction'].detach().cpu(), pred.detach()) <= 1e-4
    num = sum((p.grad.cpu().double() - grads[k].double()).norm().item() ** 2 for k, p in model.named_parameters())
    den = sum(g.double().norm().item() ** 2 for g in grads.values())
    assert (num / den) ** 0.5 <= 1e-2, (num / den) ** 0.5
    after = model.state_dict()
    assert all(torch.equal(after[k].cpu(), sd[k]) for k in sd)


def test_fused_finalize_is_safe_with_branch_streams(mirror):
    """FFPN_FUSED_FIN=1 (BatchNorm finalize by the conv's last CTA) with convs in flight on several streams: every launch
    has its own arrival counter, so the result is bit-identical to the separate finalize kernel on one stream."""
    import ffpn
    ffpn.set_compute_dtype(torch.bfloat16)
    sd = O.make_state_dict(seed=11)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 8, 64, 64, 40, 64, seed=4).items()}
    l0, g0, b0 = _trainer_grads(mirror, sd, batch, {'FFPN_STREAMS': '0', 'FFPN_FUSED_FIN': '0'})
    for _ in range(3):
        l1, g1, b1 = _trainer_grads(mirror, sd, batch, {'FFPN_STREAMS': '1', 'FFPN_FUSED_FIN': '1'})
        assert l1 == l0 and torch.equal(g1, g0)
        assert all(torch.equal(b1[k], b0[k]) for k in b0)


def test_fused_head_activation_equals_separate_sigmoid(mirror):
    """final1 + sigmoid in one kernel (default) against final1 -> torch.sigmoid: same prediction, same gradients."""
    from ffpn import functional as FF
    sd = O.make_state_dict(seed=51)
    batch = O.synthetic_batch(2, 4, 64, 32, 16, 32, seed=5)
    res = []
    for fused in (True, False):
        FF.fuse_head_activation(fused)
        try:
            model, cb, out = _run(mirror, sd, batch, 'relative_2d_max')
            _loss(mirror, cb, out).backward()
        finally:
            FF.fuse_head_activation(True)
        res.append((out['prediction'].detach().clone(), torch.cat([p.grad.reshape(-1) for p in model.parameters()])))
    assert rel(res[0][0], res[1][0]) <= 1e-6
    assert rel(res[0][1], res[1][1]) <= 1e-5


def test_no_concat_copies_and_bucketed_step(mirror, monkeypatch):
    """(a) NS-1: the hybrid model's step launches no slice_copy (producers write into the concat slots, gradients are read in
    place).  (b) the trainer's two gradient buckets: after the verification step the first bucket is reduced + stepped from
    inside backward; parameters equal the single-bucket trainer's bit for bit."""
    import ffpn
    from ffpn import ops
    from ffpn.trainer import FusionTrainer
    ffpn.set_compute_dtype(torch.bfloat16)
    sd = O.make_state_dict(seed=61)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 32, 16, 32, seed=6).items()}

    def boom(*a, **k):
        raise AssertionError('slice_copy launched: a concat member was copied')
    params = []
    for buckets in (True, False):
        monkeypatch.setenv('FFPN_BUCKETS', '1' if buckets else '0')
        model = mirror.build('FPNHybridFusion', 'relative_2d_max').cuda()
        model.load_state_dict(sd, strict=True)
        model.train()
        tr = FusionTrainer(model, _crit(mirror))
        assert (tr.n_early > 0) == buckets
        if buckets:
            monkeypatch.setattr(ops, 'slice_copy', boom)
        for i in range(3):
            tr.step(batch)
            if buckets and i == 0:
                assert tr._bucket_checked and 0 < tr.n_early < tr.flat_p.numel()       # layout verified, two buckets kept
        torch.cuda.synchronize()
        params.append({k: p.detach().clone() for k, p in model.named_parameters()})
        tr.close()
    for k in params[0]:
        assert torch.equal(params[0][k], params[1][k]), k


def test_ensemble_of_five_checkpoints_full_volume(mirror, tmp_path):
    """the evaluation regime of the reference (validate_ensemble.py:221-263, test_utils.py:21-38,
    354-360): five checkpoints in the reference's .ckpt layout, members in eval() mode (BatchNorm on running statistics), one
    full-volume batch-1 forward each, sigmoid outputs averaged -- against the oracle's average, in the fp32 exact mode and in
    bf16.  Also: eval forwards cache the BatchNorm coefficients (no finalize launch from the second forward on)."""
    import ffpn
    from ffpn import checkpoint as ck
    S, H, W, S2, W2 = 16, 128, 128, 64, 128                       # batch-1 volume on a 16-multiple en-face grid (training_config.py:103-106)
    batch = O.synthetic_batch(1, S, H, W, S2, W2, seed=77)
    paths, sds = [], []
    for i in range(5):
        sd = O.make_state_dict(seed=200 + i, randomize_running=True)
        net = mirror.build('FPNHybridFusion', 'relative_2d_max')
        net.load_state_dict(sd, strict=True)
        path = str(tmp_path / f'epoch={i}-Dice=0.9{i}.ckpt')
        ck.save_checkpoint(mirror.wrapper.Model(net, None, None, None, None, []), path, epoch=i)
        paths.append(path)
        sds.append(sd)
    want = sum(O.fpn_hybrid_fusion_forward(sd, batch, 'relative_2d_max', train=False)['prediction'] for sd in sds) / 5
    cb = {k: v.cuda() for k, v in batch.items()}
    for dtype, tol in ((torch.float32, 1e-4), (torch.bfloat16, 5e-2)):
        ffpn.set_compute_dtype(dtype)
        ens = ck.Ensemble.from_checkpoints(lambda: mirror.build('FPNHybridFusion', 'relative_2d_max'), paths, device='cuda')
        assert not ens.training and all(not m.training for m in ens.members)
        out = ens(cb)['prediction']
        n0 = ffpn.lib.launch_count(0)
        out2 = ens(cb)['prediction']
        n1 = ffpn.lib.launch_count(0)
        ens(cb)
        n2 = ffpn.lib.launch_count(0)
        assert tuple(out.shape) == (1, 1, S, 1, W)
        assert rel(out.cpu(), want) <= tol, (dtype, rel(out.cpu(), want))
        assert torch.equal(out, out2)
        assert n2 - n1 == n1 - n0                                  # steady state
        # per member: 92 convs, block ends, pools, tails -- no BatchNorm finalize (coefficients cached), and in bf16 no per-conv
        # weight packing either (the ensemble's packed-weight arena)
        assert (n1 - n0) // 5 <= (170 if dtype == torch.bfloat16 else 260), (dtype, (n1 - n0) // 5)
        for m, sd in zip(ens.members, sds):                        # eval must not touch the buffers
            after = m.state_dict()
            assert all(torch.equal(after[k].cpu(), sd[k]) for k in sd)
        # weights replaced under the ensemble: the arena is regenerated (parameter version counters), never stale
        ens.members[0].load_state_dict(sds[1], strict=True)
        want2 = (want * 5 - O.fpn_hybrid_fusion_forward(sds[0], batch, 'relative_2d_max', train=False)['prediction']
                 + O.fpn_hybrid_fusion_forward(sds[1], batch, 'relative_2d_max', train=False)['prediction']) / 5
        assert rel(ens(cb)['prediction'].cpu(), want2) <= tol
        ens._arena.close()
    ffpn.set_compute_dtype(torch.bfloat16)
