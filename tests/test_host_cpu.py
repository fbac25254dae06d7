"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol of include/ffpn.h, the
reference-shaped modules keep the reference's state_dict / constructor / registry contract, seeded
initialisation reproduces the reference's, and the product refuses to run without CUDA (no fallback)."""
import inspect
import os
import re

import numpy as np
import pytest
import torch

from oracle import fusion_fpn_oracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ffpn import lib
    header = open(os.path.join(REPO, 'include', 'ffpn.h')).read()
    declared = set(re.findall(r'\b(ffpn_[a-z0-9_]+)\s*\(', header))
    declared -= {'ffpn_last_error'} - set(lib.EXPORTS)
    handle = lib.load()
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.ffpn_abi_version() == 3
    d = lib.ConvDesc()
    assert handle.ffpn_conv_workspace_bytes(d) >= 0          # pure host call, no GPU needed


def test_binding_matches_the_header_argument_for_argument():
    """ffpn/lib.py declares the ctypes signature of every entry point by hand: the number of arguments and their kinds
    (pointer / 32-bit / 64-bit / float / double / size_t) must be those of the prototype in include/ffpn.h -- a binding that is one
    argument off would still load and then hand the kernels garbage."""
    import ctypes as C
    from ffpn import lib
    header = re.sub(r'/\*.*?\*/', ' ', open(os.path.join(REPO, 'include', 'ffpn.h')).read(), flags=re.S)
    protos = dict(re.findall(r'\b(ffpn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', header))

    def kind(decl):
        decl = decl.strip()
        if '*' in decl:
            return 'ptr'
        t = decl.rsplit(None, 1)[0] if ' ' in decl else decl
        return {'int': 'i32', 'int32_t': 'i32', 'int64_t': 'i64', 'float': 'f32', 'double': 'f64', 'size_t': 'size'}[t.replace('const ', '').strip()]

    def ckind(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, 'contents') or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return 'ptr'
        return {C.c_int: 'i32', C.c_int64: 'i64', C.c_float: 'f32', C.c_double: 'f64', C.c_size_t: 'size'}[t]

    checked = 0
    for name, args in list(lib.SIGNATURES.items()) + [(n, a) for n, (a, _) in lib.NO_CTX.items()]:
        params = [q for q in protos[name].split(',') if q.strip() and q.strip() != 'void']
        if name in lib.SIGNATURES:
            assert 'ffpn_ctx' in params[0], name
            params = params[1:]
        want = [kind(q) for q in params]
        got = [ckind(t) for t in args]
        # size_t and int64 have the same width on this ABI; everything else must agree exactly
        norm = lambda ks: ['i64' if k == 'size' else k for k in ks]
        assert norm(want) == norm(got), (name, want, got)
        checked += 1
    assert checked == len(lib.EXPORTS)


def test_sass_has_no_foreign_arch():
    import subprocess
    from ffpn import lib
    out = subprocess.run(['cuobjdump', '-lelf', lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_no_cpu_fallback(mirror):
    model = mirror.build('FPNHybridFusion')
    batch = O.synthetic_batch(1, 4, 64, 16, 8, 32)
    from ffpn.lib import FfpnError
    with pytest.raises(FfpnError):
        model(batch)


def test_state_dict_contract(mirror, golden_dir):
    model = mirror.build('FPNHybridFusion')
    sd = model.state_dict()
    osd = O.make_state_dict()
    assert list(sd.keys()) == list(osd.keys())
    assert all(sd[k].shape == osd[k].shape and sd[k].dtype == osd[k].dtype for k in sd)
    model.load_state_dict(osd, strict=True)
    assert sum(p.numel() for p in model.parameters()) == 6142481 and len(list(model.parameters())) == 275
    assert all(p.requires_grad for p in model.parameters())
    kinds = {}
    for m in model.modules():
        kinds[type(m).__mro__[1].__name__ if type(m).__module__.startswith('models') and type(m).__mro__[1].__module__.startswith('torch') else type(m).__name__] = \
            kinds.get(type(m).__mro__[1].__name__ if type(m).__module__.startswith('models') and type(m).__mro__[1].__module__.startswith('torch') else type(m).__name__, 0) + 1
    n = lambda cls: sum(isinstance(m, cls) for m in model.modules())
    # SURVEY.md App. A leaf-module census
    assert (n(torch.nn.Conv3d), n(torch.nn.BatchNorm3d), n(torch.nn.Conv2d), n(torch.nn.BatchNorm2d)) == (62, 61, 30, 30)
    assert (n(torch.nn.ReLU), n(torch.nn.MaxPool3d), n(torch.nn.MaxPool2d)) == (73, 4, 4)
    # wrapper prefix (pl_model_wrapper.py:123)
    wrap = mirror.wrapper.Model(model, None, None, None, None, [])
    assert next(iter(wrap.state_dict())) == 'model.resensnet.conv1.0.convBlock.0.0.weight'


def test_seeded_weight_init_matches_reference(mirror, golden_dir):
    fx = np.load(os.path.join(golden_dir, 'weight_init.npz'))
    torch.manual_seed(1234)
    model = mirror.build('FPNHybridFusion')
    model.apply(mirror.weight_init.weight_init)
    sd = model.state_dict()
    assert [str(n) for n in fx['names']] == list(sd.keys())
    s1 = np.array([v.double().sum().item() for v in sd.values()])
    s2 = np.array([v.double().abs().sum().item() for v in sd.values()])
    np.testing.assert_allclose(s1, fx['sum'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(s2, fx['abssum'], rtol=0, atol=1e-9)


def test_factory_and_signatures(mirror):
    fc = mirror.fusion_nets.factory_classes
    assert sorted(fc) == ['FPN', 'FPN2D', 'FPNClassification', 'FPNHybridFusion', 'FPNHybridFusionRegression',
                          'FPNLateFusion', 'FPNLateFusionRegression', 'FPNRegression']
    counts = {'FPN': 4369553, 'FPN2D': 2077745, 'FPNLateFusion': 6447314}
    for name, c in counts.items():
        assert sum(p.numel() for p in mirror.build(name).parameters()) == c
    from models.fpn import fusion3D2D, unets3D, unets2D, components
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(fusion3D2D.ModifiedUnet3D2D.__init__) == ['self', 'config', 'interpolate', 'feature_fusion']
    assert sig(fusion3D2D.ModifiedUnet3D2DLevel5.forward) == ['self', 'oct', 'slo']
    assert sig(unets3D.ModifiedUnet3D.__init__) == ['self', 'config', 'original', 'classification']
    assert sig(unets2D.ModifiedUnet2D.__init__) == ['self', 'config', 'output_features']
    assert sig(fusion3D2D.unet3dConvX.__init__) == ['self', 'in_size', 'out_size', 'kernel_size', 'stride', 'padding',
                                                    'is_batchnorm', 'is_residual', 'dropout', 'downsample']
    assert sig(components.unet3dUp2modified.__init__) == ['self', 'lowlayer_channels', 'currlayer_channels', 'upfactor',
                                                          'is_deconv', 'is_residual', 'dropout', 'is_batchnorm']
    assert sig(fusion3D2D.unet3dUp2modified.forward) == ['self', 'inputs1', 'inputs1_b', 'inputs2']
    m = mirror.build('FPNHybridFusion', crop='relative_2d')
    assert m.interpolate == '2d' and mirror.build('FPNHybridFusion', crop='oct').interpolate is None
    assert mirror.build('FPNHybridFusion', crop='relative_2d_max').resensnet.interpolate == '2d_max'
    body = m.resensnet
    assert body.channels == [16, 32, 64, 128, 256] and len(body.dropout) == 9 and body.model_name == 'ModifiedUnet3D'
    with pytest.raises(ValueError):
        fusion3D2D.ModifiedUnet3D2D(m.config, None, 'mul')
    assert 'resensnet.final1.0.weight' in mirror.build('FPN2D').state_dict()
    late = mirror.build('FPNLateFusion')
    assert late.resensnet3d.use_1x1 is False and 'fusion_module.weight' in late.state_dict()
    cls = mirror.build('FPNClassification')
    assert not any(p.requires_grad for p in cls.resensnet.zdimRed1.parameters())


def test_config_flags(mirror):
    cfg = mirror.config
    for k, v in dict(batch_size=8, virtual_batch_size=1, learning_rate=0.1, number_of_outputs=1, epochs=40,
                     force_mem_cache_release='ReleaseMemCache', threads=8, base_channels=64).items():
        assert getattr(cfg, k) == v, k
    assert cfg.use_complementary is True and cfg.layers == [1, 1, 2, 4] and cfg.number_of_channels == [32, 64, 128, 256]


def test_losses_match_oracle(mirror):
    g = torch.Generator().manual_seed(0)
    pred = torch.rand(2, 1, 8, 1, 32, generator=g)
    mask = (torch.rand(2, 1, 8, 1, 32, generator=g) > 0.5).float()
    crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
    total, parts = crit({'mask': mask}, {'prediction': pred})
    assert abs(total.item() - O.mix_loss(pred, mask).item()) < 1e-7
    assert abs(parts['Dice'].item() - O.dice_loss(pred, mask).item()) < 1e-7
    with pytest.raises(AssertionError):
        crit({'mask': mask[:1]}, {'prediction': pred})


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ffpn.trainer import flatten_parameters, allreduce_mean_
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv3d(2, 3, 1), torch.nn.BatchNorm3d(3))
    flat_p, flat_g = flatten_parameters(net)
    g = torch.Generator().manual_seed(100 + rank)
    for p in net.parameters():
        p.grad.copy_(torch.randn(p.shape, generator=g))          # rank-local gradients (own BN stats / loss)
    local = {k: p.grad.clone().numpy() for k, p in net.named_parameters()}
    scale = allreduce_mean_(flat_g)
    # numpy arrays are pickled by value: torch tensors would travel as file descriptors that die with this process
    q.put((rank, local, {k: (p.grad.clone() * scale).numpy() for k, p in net.named_parameters()},
           all(p.data_ptr() >= flat_p.data_ptr() for p in net.parameters())))
    dist.destroy_process_group()


def test_dp_gradient_average_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    want = O.dp_average_gradients([{k: torch.from_numpy(v) for k, v in res[i][1].items()} for i in range(2)])
    for r in res:
        assert r[3]
        for k in want:
            assert torch.allclose(torch.from_numpy(r[2][k]), want[k], atol=1e-7)


def _bcast_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ffpn.trainer import FusionTrainer
    torch.manual_seed(1000 + rank)                               # ranks initialise DIFFERENTLY (unseeded init, per-rank checkpoint ...)
    net = torch.nn.Sequential(torch.nn.Conv3d(2, 3, 1), torch.nn.BatchNorm3d(3))
    with torch.no_grad():
        net[1].running_mean.add_(rank + 1.0)
        net[1].num_batches_tracked.add_(7 * (rank + 1))
    before = float(sum(p.double().sum() for p in net.parameters()))
    tr = FusionTrainer(net, criterion=None)
    q.put((rank, before, tr.flat_p.clone().numpy(), {k: v.clone().numpy() for k, v in net.state_dict().items()}))
    dist.destroy_process_group()


def test_trainer_broadcasts_rank0_state_world2_gloo():
    """Replicas that start from different weights / BatchNorm buffers are made equal to rank 0 at FusionTrainer construction
    (the reference's DataParallel re-replicates device 0's module every step, so its replicas cannot differ)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] != res[1][1]                                # they did start out different
    assert np.array_equal(res[0][2], res[1][2])
    for k in res[0][3]:
        assert np.array_equal(res[0][3][k], res[1][3][k]), k
    assert float(res[1][3]['1.running_mean'][0]) == 1.0 and int(res[1][3]['1.num_batches_tracked']) == 7    # rank 0's values


def test_every_kernel_waits_for_its_predecessor_grid():
    """Programmatic dependent launch is only safe if EVERY kernel launched with the attribute begins with
    griddepcontrol.wait before it touches global memory, and if no launch bypasses ffpn_launch().  Source-level guard:
    each __global__ body starts with pdl_prologue() (or pdl_trigger() + a later pdl_wait()), no raw <<< >>> launches."""
    import glob
    csrc = os.path.join(REPO, 'multimodal-fusion-fpn_b200', 'csrc')
    nkernels = 0
    for path in sorted(glob.glob(os.path.join(csrc, '*.cu'))):
        src = open(path).read()
        assert '<<<' not in src, f'{os.path.basename(path)}: raw kernel launch bypasses ffpn_launch()'
        pos = 0
        while True:
            j = src.find('__global__', pos)
            if j < 0:
                break
            k = j + len('__global__')
            m = re.compile(r'\s*(?:void\s+)?__launch_bounds__\s*\(').match(src, k)
            if m:                                   # skip the balanced __launch_bounds__(...)
                depth, k = 1, m.end()
                while depth:
                    depth += (src[k] == '(') - (src[k] == ')')
                    k += 1
            k = src.find('(', k)                    # parameter list
            depth, k = 1, k + 1
            while depth:
                depth += (src[k] == '(') - (src[k] == ')')
                k += 1
            b = src.find('{', k)
            first = src[b + 1:b + 200].strip()
            assert first.startswith('pdl_prologue();') or first.startswith('pdl_trigger();'), (os.path.basename(path), first[:60])
            if first.startswith('pdl_trigger();'):
                end = src.find('\n}\n', b)
                assert 'pdl_wait();' in src[b:end], os.path.basename(path)
            nkernels += 1
            pos = b + 1
    assert nkernels >= 40


def test_branch_stream_toggle(monkeypatch):
    from ffpn import functional as FF
    for value, branches, wgrad in (('1', True, True), ('0', False, False), ('branches', True, False), ('wgrad', False, True)):
        monkeypatch.setenv('FFPN_STREAMS', value)
        assert FF.streams_enabled() is branches and FF.streams_enabled('wgrad') is wgrad
    monkeypatch.delenv('FFPN_STREAMS')
    assert FF.streams_enabled() and FF.streams_enabled('wgrad')


def test_eval_batchnorm_coefficient_cache_validity(monkeypatch):
    """Host logic of ops.eval_bn_coefficients (no GPU: the finalize call is stubbed): an entry is reused only for the same
    tensor OBJECTS with unchanged version counters within the same training epoch -- in-place updates through torch, a
    training-mode finalize or a graph replay (ops.note_bn_statistics_update) and a different module with equal values all miss."""
    from ffpn import ops
    calls = []

    def fake_finalize(partial, rows, count, gamma, beta, rm, rv, momentum, eps, training):
        calls.append(1)
        return tuple(torch.full((gamma.numel(),), float(len(calls))) for _ in range(4))

    monkeypatch.setattr(ops, 'bn_finalize', fake_finalize)
    monkeypatch.setattr(torch.cuda, 'is_current_stream_capturing', lambda: False)      # (raises without a GPU)
    g, b, rm, rv = torch.ones(4), torch.zeros(4), torch.zeros(4), torch.ones(4)
    first = ops.eval_bn_coefficients(g, b, rm, rv, 1e-5)
    assert ops.eval_bn_coefficients(g, b, rm, rv, 1e-5)[0] is first[0] and len(calls) == 1      # hit
    rm.add_(1.0)                                                                                  # torch in-place update: version bump
    assert ops.eval_bn_coefficients(g, b, rm, rv, 1e-5)[0] is not first[0] and len(calls) == 2
    ops.eval_bn_coefficients(g, b, rm, rv, 1e-5)
    assert len(calls) == 2
    ops.note_bn_statistics_update()                                                               # raw-pointer update (kernels, graph replay)
    ops.eval_bn_coefficients(g, b, rm, rv, 1e-5)
    assert len(calls) == 3
    ops.eval_bn_coefficients(g, b, rm, rv, 1e-3)                                                  # another eps
    assert len(calls) == 4
    g2, b2, rm2, rv2 = g.clone(), b.clone(), rm.clone(), rv.clone()                               # another module, equal values
    ops.eval_bn_coefficients(g2, b2, rm2, rv2, 1e-3)
    assert len(calls) == 5
    ops.eval_bn_coefficients(g2, b, rm, rv, 1e-3)                                                 # sibling tensor swapped
    assert len(calls) == 6


def test_checkpoint_roundtrip_reference_layout(mirror, tmp_path):
    """ffpn.checkpoint: pytorch-lightning 1.5.10 `save_weights_only` layout (wrapper keys, `model.` prefix), the loading
    rules of train.py:146-153 (with / without 'state_dict') and the legacy key fix of validate_ensemble.py:251-256."""
    from ffpn import checkpoint as ck
    torch.manual_seed(3)
    net = mirror.build('FPNHybridFusion')
    net.apply(mirror.weight_init.weight_init)
    wrap = mirror.wrapper.Model(net, None, None, None, None, None)
    path = str(tmp_path / 'epoch=3-Dice=0.9000.ckpt')
    ck.save_checkpoint(net, path, epoch=3, global_step=120)            # from the bare network ...
    blob = torch.load(path)
    assert set(blob) == {'epoch', 'global_step', 'pytorch-lightning_version', 'state_dict'} and blob['epoch'] == 3
    assert list(blob['state_dict']) == list(wrap.state_dict())          # ... same keys and order as the wrapper's
    # wrapper <- file, bare network <- file, bare network <- bare state dict, legacy 'resensenet' spelling
    for target, source in ((mirror.wrapper.Model(mirror.build('FPNHybridFusion'), None, None, None, None, None), path),
                           (mirror.build('FPNHybridFusion'), path),
                           (mirror.build('FPNHybridFusion'), net.state_dict()),
                           (mirror.build('FPNHybridFusion'),
                            {'state_dict': {k.replace('resensnet', 'resensenet'): v for k, v in wrap.state_dict().items()}})):
        ck.load_checkpoint(target, source, strict=True)
        got = target.model.state_dict() if isinstance(target, mirror.wrapper.Model) else target.state_dict()
        for k, v in net.state_dict().items():
            assert torch.equal(got[k], v), k
    with pytest.raises(RuntimeError):                                   # strict: a foreign key is an error, as in the reference
        ck.load_checkpoint(mirror.build('FPNHybridFusion'), {**net.state_dict(), 'bogus.weight': torch.zeros(1)})


def test_ensemble_averages_like_the_reference():
    """test_utils.average_outputs semantics (dict key by key, tensors sum/n, strings: first) and eval-mode members."""
    from ffpn import checkpoint as ck

    class Member(torch.nn.Module):
        def __init__(self, bias):
            super().__init__()
            self.bn = torch.nn.BatchNorm1d(1)
            self.bias = bias

        def forward(self, batch):
            assert not self.training                                    # ensemble members run on running statistics
            return {'prediction': batch['image'] + self.bias, 'id': f'scan-{self.bias}'}

    ens = ck.Ensemble([Member(b) for b in (0.0, 1.0, 5.0)])
    ens.train()                                                         # stays in eval mode
    x = torch.arange(6.).view(2, 3)
    out = ens({'image': x})
    assert torch.allclose(out['prediction'], x + 2.0) and out['id'] == 'scan-0.0'
    assert not out['prediction'].requires_grad
    with pytest.raises(AssertionError):
        ck.average_outputs([1, 2], int)
    with pytest.raises(ValueError):
        ck.Ensemble([])


def _conv_geometries(B, S, H, W, S2, W2, ch=(16, 32, 64, 128, 256)):
    """Every convolution of one FPNHybridFusion forward as (phys input shape, cout, kernel, stride, pad), from the shape algebra
    of SURVEY.md App. A / B (pools floor, projection halves with ceil)."""
    out = []
    sp3 = [(S, W, H), (S, W // 2, H // 2), (S, W // 4, H // 4), (S // 2, W // 8, H // 8), (S // 4, W // 16, H // 16)]
    sp2 = [(S2, W2), (S2, W2 // 2), (S2, W2 // 4), (S2 // 2, W2 // 8), (S2 // 4, W2 // 16)]
    for l in range(5):
        cin, c = (1 if l == 0 else ch[l - 1]), ch[l]
        s, w, h = sp3[l]
        out += [((B, s, w, h, cin), c, (1, 3, 3), (1, 1, 1), (0, 1, 1)), ((B, s, w, h, cin), c, (1, 1, 1), (1, 1, 1), (0, 0, 0))]
        out += [((B, s, w, h, c), c, (1, 3, 3), (1, 1, 1), (0, 1, 1)), ((B, s, w, h, c), c, (3, 1, 1), (1, 1, 1), (1, 0, 0))]
        n, hh = 4 - l, h
        if n > 0:
            out.append(((B, s, w, h, c), c, (1, 1, 1), (1, 1, 2 ** n), (0, 0, 0)))
        for _ in range(n):
            out.append(((B, s, w, hh, c), c, (1, 1, 3), (1, 1, 2), (0, 0, 1)))
            hh = (hh - 1) // 2 + 1
        out.append(((B, s, w, hh, c), c, (1, 1, 4), (1, 1, 1), (0, 0, 0)))
        s2, w2 = sp2[l]
        out += [((B, s2, w2, 1, cin), c, (1, 3, 1), (1, 1, 1), (0, 1, 0)), ((B, s2, w2, 1, cin), c, (1, 1, 1), (1, 1, 1), (0, 0, 0))]
        out += [((B, s2, w2, 1, c), c, (1, 3, 1), (1, 1, 1), (0, 1, 0)), ((B, s2, w2, 1, c), c, (3, 1, 1), (1, 1, 1), (1, 0, 0))]
    for l in (4, 3, 2, 1):
        low, cur = (ch[4] * 2 if l == 4 else ch[l]), ch[l - 1]
        s, w, _ = sp3[l - 1]
        out += [((B, s, w, 1, low + 2 * cur), cur, (3, 3, 1), (1, 1, 1), (1, 1, 0)), ((B, s, w, 1, cur), cur, (3, 3, 1), (1, 1, 1), (1, 1, 0)),
                ((B, s, w, 1, low + 2 * cur), cur, (1, 1, 1), (1, 1, 1), (0, 0, 0))]
    return out


@pytest.mark.parametrize('shape', [(1, 64, 128, 128, 128, 128), (8, 32, 128, 128, 320, 128), (8, 32, 496, 128, 320, 128),
                                   (16, 32, 128, 128, 320, 128), (2, 8, 64, 32, 20, 48)],
                         ids=['C1', 'C2', 'C3', 'C4', 'toy'])
def test_every_conv_of_the_model_has_a_tensor_core_plan(shape):
    """Host-only plan introspection (ffpn_conv_plan_info): at every BASELINE.json shape -- and at the toy shape of the golden
    fixtures -- forward, dgrad and weight gradient of every convolution with Cin >= 16 are taken by the warp-specialised tcgen05
    kernels (the Cin == 1 stems have their own streaming kernels); nothing on the bf16 path falls to the CUDA-core kernels."""
    import ctypes as C
    from ffpn import lib, ops
    handle = lib.load()
    buf = C.create_string_buffer(800)
    missing = []
    for (x_shape, cout, k, s, p) in _conv_geometries(*shape):
        if x_shape[-1] == 1:
            continue
        d = ops.make_desc(x_shape, cout, k, s, p, torch.bfloat16)
        for mode, what in ((0, 'fwd'), (1, 'dgrad'), (2, 'wgrad')):
            assert handle.ffpn_conv_plan_info(C.byref(d), mode, buf, 800) == 0
            if buf.value.decode().startswith('not on'):
                missing.append((what, x_shape, cout, k, s))
    assert not missing, missing
