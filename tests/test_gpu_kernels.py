"""Kernel-level parity on the B200: every C-ABI op against the corresponding torch op / oracle function on
identical inputs.  Tolerances: fp32 path rel-L2 <= 1e-5; bf16 path rel-L2 <= 1e-2
against torch fed the same bf16-rounded inputs with fp32 accumulation; index outputs bit-exact."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.bfloat16]


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def tol(dt):
    return 1e-5 if dt == torch.float32 else 1e-2


@pytest.fixture(scope='module')
def ops():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from ffpn import ops as o
    return o


def phys(x):      # logical (B,C,S,W,H) -> physical (B,S,W,H,C)
    return x.permute(0, 2, 3, 4, 1).contiguous()


def logical(p):
    return p.permute(0, 4, 1, 2, 3)


CONV_CASES = [
    # Cin, Cout, kernel, stride, pad, (S, W, H)
    (1, 16, (1, 3, 3), (1, 1, 1), (0, 1, 1), (3, 10, 12)),
    (16, 16, (1, 3, 3), (1, 1, 1), (0, 1, 1), (3, 9, 14)),
    (16, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 8, 8)),
    (32, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), (5, 6, 7)),
    (16, 32, (1, 1, 1), (1, 1, 1), (0, 0, 0), (2, 5, 9)),
    (16, 16, (1, 1, 3), (1, 1, 2), (0, 0, 1), (2, 5, 31)),
    (16, 16, (1, 1, 3), (1, 1, 2), (0, 0, 1), (2, 5, 16)),
    (32, 32, (1, 1, 1), (1, 1, 8), (0, 0, 0), (2, 3, 62)),
    (64, 64, (1, 1, 4), (1, 1, 1), (0, 0, 0), (2, 3, 8)),
    (96, 32, (3, 3, 1), (1, 1, 1), (1, 1, 0), (6, 7, 1)),
    (16, 16, (1, 3, 1), (1, 1, 1), (0, 1, 0), (9, 11, 1)),
    (40, 24, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 5, 6)),       # channel counts that are not tile multiples
    (1, 16, (1, 3, 1), (1, 1, 1), (0, 1, 0), (9, 11, 1)),        # 2-D stem (1,3)
    (1, 16, (1, 1, 1), (1, 1, 1), (0, 0, 0), (3, 10, 12)),       # stem shortcut
    (16, 16, (1, 1, 3), (1, 1, 2), (0, 0, 1), (2, 5, 25)),       # odd depth: dgrad falls back to the CUDA-core kernel
]


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_fwd_dgrad_wgrad(ops, dt, case):
    cin, cout, k, s, p, (S, W, H) = case
    if dt == torch.bfloat16 and (cin % 8 or cout % 8) and cin != 1:
        pytest.skip('bf16 activations need channel multiples of 8')
    g = torch.Generator().manual_seed(hash(case) & 0xffff)
    B = 2
    x = torch.randn(B, cin, S, W, H, generator=g).cuda()
    w = (torch.randn(cout, cin, *k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5).cuda()
    sc = (0.5 + torch.rand(cin, generator=g)).cuda()
    sh = (0.3 * torch.randn(cin, generator=g)).cuda()
    xq = x.to(dt).float()                                     # what the kernel actually reads
    for affine in (False, True):
        xin = torch.relu(xq * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)) if affine else xq
        xin = xin.clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        ref = F.conv3d(xin, wr, None, s, p)
        y, partial, rows = ops.conv_fwd(phys(x).to(dt), w, k, s, p, sc if affine else None, sh if affine else None, affine)
        yl = logical(y.float())
        assert yl.shape == ref.shape
        assert rel(yl, ref.detach()) <= tol(dt), ('fwd', affine)
        # statistics are those of the stored (rounded) output
        st = partial.view(-1, 2, cout)[:rows].double().sum(0)
        ys = y.float().double().reshape(-1, cout)
        assert torch.allclose(st[0], ys.sum(0), rtol=1e-4, atol=1e-3 * ys.abs().sum(0).max().item())
        assert torch.allclose(st[1], (ys * ys).sum(0), rtol=1e-4)
        dy = torch.randn(ref.shape, generator=g).cuda()
        dyq = dy.to(dt).float()
        ref.backward(dyq)
        dx = ops.conv_dgrad(phys(dy).to(dt), w, tuple(phys(x).shape), k, s, p)
        assert rel(logical(dx.float()), xin.grad) <= tol(dt), ('dgrad', affine)
        dw = ops.conv_wgrad(phys(x).to(dt), phys(dy).to(dt), w.shape, k, s, p, sc if affine else None,
                            sh if affine else None, affine)
        assert rel(dw, wr.grad) <= tol(dt), ('wgrad', affine)
    add = torch.randn(B, cin, S, W, H, generator=g).cuda().to(dt)
    dx2 = ops.conv_dgrad(phys(dy).to(dt), w, tuple(phys(x).shape), k, s, p, addend=phys(add))
    assert rel(logical(dx2.float()), logical(dx.float()) + add.float()) <= tol(dt)


@pytest.mark.parametrize('dt', DT)
def test_bn_finalize_and_backward(ops, dt):
    g = torch.Generator().manual_seed(3)
    B, C, S, W, H = 2, 16, 3, 5, 7
    y = (torch.randn(B, C, S, W, H, generator=g) * 2 + 0.5).cuda().to(dt)
    yq = y.float()
    bn = torch.nn.BatchNorm3d(C).cuda()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(C, generator=g))
        bn.bias.copy_(0.1 * torch.randn(C, generator=g))
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    yr = yq.clone().requires_grad_(True)
    out = torch.relu(bn(yr))
    # partial sums straight from the tensor (what a conv epilogue would have written)
    yp = phys(y)
    flat = yp.float().reshape(-1, C)
    partial = torch.zeros(4, 2, C, device='cuda')
    partial[1, 0], partial[1, 1] = flat.sum(0), (flat * flat).sum(0)
    a, b, mean, invstd = ops.bn_finalize(partial.flatten(), 4, flat.shape[0], bn.weight.detach(), bn.bias.detach(), rm, rv,
                                         0.1, 1e-5, True)
    assert torch.allclose(rm, bn.running_mean, rtol=1e-5, atol=1e-6) and torch.allclose(rv, bn.running_var, rtol=1e-5)
    z = ops.block_end_fwd(yp, a, b)
    assert rel(logical(z.float()), out.detach()) <= tol(dt)
    dA = torch.randn(out.shape, generator=g).cuda().to(dt)
    out.backward(dA.float())
    part, rows = ops.bn_bwd_reduce(phys(dA), yp, a, b, True)
    dg, db, cA, cP, cQ = ops.bn_bwd_finalize(part, rows, 2, 1, flat.shape[0], bn.weight.detach(), mean, invstd)
    dy = ops.bn_bwd_apply(phys(dA), yp, a, b, True, cA, cP, cQ)
    assert rel(dg, bn.weight.grad) <= 10 * tol(dt) and rel(db, bn.bias.grad) <= 10 * tol(dt)
    assert rel(logical(dy.float()), yr.grad) <= 3 * tol(dt)
    # a block with a conv + BN shortcut: both applies in one pass over G (ffpn_bn_bwd_apply2) == two separate applies, bit for bit
    y2 = phys((torch.randn(B, C, S, W, H, generator=g) - 0.3).cuda().to(dt))
    c2 = (cA * 0.7 + 0.1, cP * 1.3 - 0.05, cQ + 0.2)
    one_a = ops.bn_bwd_apply(phys(dA), yp, a, b, False, cA, cP, cQ)
    one_b = ops.bn_bwd_apply(phys(dA), y2, a, b, False, *c2)
    two_a, two_b = ops.bn_bwd_apply2(phys(dA), yp, y2, (cA, cP, cQ), c2)
    assert torch.equal(one_a, two_a) and torch.equal(one_b, two_b)
    # eval mode: running statistics
    bn.eval()
    a2, b2, _, _ = ops.bn_finalize(None, 0, flat.shape[0], bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                   bn.running_var, 0.1, 1e-5, False)
    z2 = ops.block_end_fwd(yp, a2, b2)
    assert rel(logical(z2.float()), torch.relu(bn(yq)).detach()) <= tol(dt)


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('kernel', [(1, 2, 2), (2, 2, 2)])
def test_maxpool_indices_bit_exact_and_routing(ops, dt, kernel):
    g = torch.Generator().manual_seed(5)
    B, C, S, W, H = 2, 16, 4, 6, 10
    x = ((torch.randn(B, C, S, W, H, generator=g) * 2).round() / 2)          # many ties, exactly representable
    x[0, 3, 1, 2, 3] = float('nan')
    x[1, 5, 2, 0, 0] = float('nan'); x[1, 5, 2, 1, 1] = float('nan')
    xc = x.cuda()
    ref_v, ref_i = F.max_pool3d(x, kernel, return_indices=True)               # CPU torch = the reference's op
    zp, idx = ops.maxpool_fwd(phys(xc).to(dt), kernel, want_idx=True)
    assert torch.equal(logical(idx).cpu(), ref_i)                              # bit-exact argmax
    assert torch.equal(torch.nan_to_num(logical(zp.float()).cpu(), nan=123.), torch.nan_to_num(ref_v, nan=123.))
    # gradient routing (finite input): stand-alone backward and the fused block-end backward
    x2 = ((torch.randn(B, C, S, W, H, generator=g) * 2).round() / 2).abs()
    xr = x2.clone().requires_grad_(True)
    pooled = F.max_pool3d(xr, kernel)
    dzp = torch.randn(pooled.shape, generator=g)
    pooled.backward(dzp)
    dz = ops.maxpool_bwd(phys(x2.cuda()).to(dt), phys(dzp.cuda()).to(dt), kernel)
    assert rel(logical(dz.float()).cpu(), xr.grad.to(dt).float()) <= 1e-6
    dzs = torch.randn(x2.shape, generator=g)
    y = torch.randn(x2.shape, generator=g)
    G, partial, rows, ncols = ops.block_end_bwd(phys(dzs.cuda()).to(dt), phys(dzp.cuda()).to(dt), phys(x2.cuda()).to(dt),
                                                phys(y.cuda()).to(dt), None, kernel)
    want = (dzs.to(dt).float() + xr.grad.to(dt).float()) * (x2 > 0)
    assert rel(logical(G.float()).cpu(), want) <= tol(dt)
    st = partial.view(-1, ncols, C)[:rows].sum(0).cpu()
    Gf = logical(G.float()).cpu()
    assert torch.allclose(st[0], Gf.sum(dim=(0, 2, 3, 4)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(st[1], (Gf * y.to(dt).float()).sum(dim=(0, 2, 3, 4)), rtol=1e-3, atol=1e-2)


def test_maxpool_golden_fixture(ops, golden_dir):
    ix = np.load(os.path.join(golden_dir, 'index_ops.npz'))
    for name, k in (('p122', (1, 2, 2)), ('p222', (2, 2, 2))):
        x = torch.from_numpy(ix[f'{name}/x']).repeat(1, 4, 1, 1, 1)          # C=4 for the fp32 vector width
        zp, idx = ops.maxpool_fwd(phys(x.cuda()), k, want_idx=True)
        assert np.array_equal(logical(idx)[:, :1].cpu().numpy(), ix[f'{name}/idx'])


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('mode', [None, '2d_max', '2d'])
def test_resize2d(ops, dt, mode):
    g = torch.Generator().manual_seed(11)
    B, C, Si, Wi = 2, 8, 20, 48
    So, Wo = (Si, Wi) if mode is None else (8, 32)
    x = ((torch.randn(B, C, Si, Wi, generator=g) * 2).round() / 2) if mode == '2d_max' else torch.randn(B, C, Si, Wi, generator=g)
    xq = x.to(dt).float()
    xr = xq.clone().requires_grad_(True)
    x5 = xr[:, :, :, :, None]
    if mode == '2d_max':
        ref, ref_i = F.adaptive_max_pool3d(x5, (So, Wo, 1), return_indices=True)
    elif mode == '2d':
        ref = F.interpolate(x5, size=(So, Wo, 1), mode='trilinear')
    else:
        ref = x5
    xp = x.cuda().permute(0, 2, 3, 1).unsqueeze(3).contiguous().to(dt)
    out, idx = ops.resize2d_fwd(xp, So, Wo, mode)
    assert rel(logical(out.float()).cpu(), ref.detach()) <= (1e-6 if mode != '2d' else tol(dt))
    if mode == '2d_max':
        assert torch.equal(idx.permute(0, 3, 1, 2).cpu().long(), ref_i[..., 0])   # bit-exact argmax
    dout = torch.randn(ref.shape, generator=g).to(dt).float()
    ref.backward(dout)
    dx = ops.resize2d_bwd(phys(dout.cuda()).to(dt), tuple(xp.shape), mode, idx)
    assert rel(logical(dx.float()).cpu()[..., 0], xr.grad) <= tol(dt)


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('out', [(5, 48), (10, 24), (20, 12)])
def test_resize2d_max_tiled(ops, dt, out):
    """Adaptive-max resize whose windows tile the input exactly (the shapes of the model: 320x128 -> 32x128): the
    vectorised backward fast path, against torch's adaptive_max_pool3d backward; ties included."""
    g = torch.Generator().manual_seed(17)
    B, C, Si, Wi = 2, 16, 20, 48
    x = (torch.randn(B, C, Si, Wi, generator=g) * 2).round() / 2
    xr = x.to(dt).float().clone().requires_grad_(True)
    ref, ref_i = F.adaptive_max_pool3d(xr[:, :, :, :, None], (out[0], out[1], 1), return_indices=True)
    xp = x.cuda().permute(0, 2, 3, 1).unsqueeze(3).contiguous().to(dt)
    o, idx = ops.resize2d_fwd(xp, out[0], out[1], '2d_max')
    assert torch.equal(idx.permute(0, 3, 1, 2).cpu().long(), ref_i[..., 0])
    dout = torch.randn(ref.shape, generator=g).to(dt).float()
    ref.backward(dout)
    dx = ops.resize2d_bwd(phys(dout.cuda()).to(dt), tuple(xp.shape), '2d_max', idx)
    assert torch.equal(logical(dx.float()).cpu()[..., 0], xr.grad.to(dt).float())


def test_resize_golden_fixture(ops, golden_dir):
    ix = np.load(os.path.join(golden_dir, 'index_ops.npz'))
    for name, o in (('a8x32', (8, 32)), ('a8x16', (8, 16)), ('a3x7', (3, 7))):
        x = torch.from_numpy(ix[f'{name}/x'])                                  # (1,1,20,48,1)
        xp = x.cuda().permute(0, 2, 3, 4, 1).contiguous().repeat(1, 1, 1, 1, 4)
        out, idx = ops.resize2d_fwd(xp, o[0], o[1], '2d_max')
        assert np.array_equal(idx[..., 0].cpu().numpy().astype(np.int64), ix[f'{name}/idx'][0, 0, :, :, 0][None])
    z = torch.from_numpy(ix['tri/x'])                                          # (1,2,5,7,1)
    zp = z.cuda().permute(0, 2, 3, 4, 1).contiguous().repeat(1, 1, 1, 1, 2)
    for key, o in (('tri/out_4x16', (4, 16)), ('tri/out_8x5', (8, 5))):
        out, _ = ops.resize2d_fwd(zp, o[0], o[1], '2d')
        np.testing.assert_allclose(logical(out)[:, :2].cpu().numpy(), ix[key], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('dt', DT)
@pytest.mark.parametrize('factor', [(2, 2), (1, 2)])
def test_upsample_and_cat(ops, dt, factor, golden_dir):
    from oracle import fusion_fpn_oracle as O
    g = torch.Generator().manual_seed(13)
    B, C, Si, Wi = 2, 8, 4, 6
    x = torch.randn(B, C, Si, Wi, 1, generator=g).to(dt).float()
    xr = x.clone().requires_grad_(True)
    ref = O.upsample_nearest(xr, (factor[0], factor[1], 1))
    cat = torch.zeros(B, Si * factor[0], Wi * factor[1], 1, C + 8, device='cuda', dtype=dt)
    ops.upsample_fwd(phys(x.cuda()).to(dt), factor[0], factor[1], out=cat, coff=8)
    assert torch.equal(logical(cat.float())[:, 8:].cpu(), ref.detach())
    assert cat[..., :8].abs().max().item() == 0
    dcat = torch.randn(cat.shape, generator=g).to(dt)
    ref.backward(logical(dcat.float())[:, 8:])
    dx = ops.upsample_bwd(dcat.cuda(), tuple(phys(x).shape), factor[0], factor[1], coff=8)
    assert rel(logical(dx.float()).cpu(), xr.grad) <= tol(dt)
    ix = np.load(os.path.join(golden_dir, 'index_ops.npz'))
    src = torch.arange(24, dtype=torch.float32).view(1, 1, 4, 6, 1).repeat(1, 4, 1, 1, 1)
    out = ops.upsample_fwd(phys(src.cuda()), factor[0], factor[1])
    assert np.array_equal(logical(out)[:, :1].cpu().numpy(), ix[f'up{factor[0]}{factor[1]}1/out'])


@pytest.mark.parametrize('dt', DT)
def test_proj_tail_and_head(ops, dt):
    g = torch.Generator().manual_seed(17)
    B, C, S, W, H = 2, 16, 3, 4, 5
    y = torch.randn(B, C, S, W, H, generator=g).to(dt).float()
    a, b = (0.5 + torch.rand(C, generator=g)), 0.2 * torch.randn(C, generator=g)
    ref = torch.relu(y * a.view(1, -1, 1, 1, 1) + b.view(1, -1, 1, 1, 1)).mean(dim=4, keepdim=True)
    out = ops.proj_tail_fwd(phys(y.cuda()).to(dt), a.cuda(), b.cuda())
    assert rel(logical(out.float()).cpu(), ref) <= tol(dt)
    dout = torch.randn(ref.shape, generator=g).to(dt)
    dA = ops.proj_tail_bwd(phys(dout.cuda()), tuple(phys(y).shape))
    assert rel(logical(dA.float()).cpu(), (dout.float() / H).expand(-1, -1, -1, -1, H)) <= tol(dt)
    # head
    n = 2
    x = torch.randn(B, C, S, W, 1, generator=g).to(dt).float()
    w = torch.randn(n, C, 1, 1, 1, generator=g)
    bias = torch.randn(n, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    ref = F.conv3d(xr, wr, br)
    logits = ops.head_fwd(phys(x.cuda()).to(dt), w.cuda(), bias.cuda())
    assert rel(logits.cpu(), ref.detach()) <= 1e-5
    dl = torch.randn(ref.shape, generator=g)
    ref.backward(dl)
    dx, dw, db = ops.head_bwd(phys(x.cuda()).to(dt), w.cuda(), dl.cuda())
    assert rel(logical(dx.float()).cpu(), xr.grad) <= tol(dt)
    assert rel(dw.cpu(), wr.grad) <= 1e-5 and rel(db.cpu(), br.grad) <= 1e-5
    # head with the wrappers' sigmoid fused in (fusion_nets.py:110,118), forward and backward
    xr.grad = wr.grad = br.grad = None
    refp = torch.sigmoid(F.conv3d(xr, wr, br))
    pred = ops.head_fwd(phys(x.cuda()).to(dt), w.cuda(), bias.cuda(), act=1)
    assert rel(pred.cpu(), refp.detach()) <= 1e-6
    refp.backward(dl)
    dx, dw, db = ops.head_bwd(phys(x.cuda()).to(dt), w.cuda(), dl.cuda(), pred)
    assert rel(logical(dx.float()).cpu(), xr.grad) <= tol(dt)
    assert rel(dw.cpu(), wr.grad) <= 1e-5 and rel(db.cpu(), br.grad) <= 1e-5
    # a head over more positions than one reduction block handles (two-stage weight-gradient reduction), run-to-run identical
    x2 = torch.randn(3, C, 40, 64, 1, generator=g).to(dt).float()
    dl2 = torch.randn(3, n, 40, 64, 1, generator=g)
    xr2, wr2, br2 = x2.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    F.conv3d(xr2, wr2, br2).backward(dl2)
    r1 = ops.head_bwd(phys(x2.cuda()).to(dt), w.cuda(), dl2.cuda())
    r2 = ops.head_bwd(phys(x2.cuda()).to(dt), w.cuda(), dl2.cuda())
    assert rel(r1[1].cpu(), wr2.grad) <= 1e-5 and rel(r1[2].cpu(), br2.grad) <= 1e-5
    assert torch.equal(r1[1], r2[1]) and torch.equal(r1[2], r2[2])


@pytest.mark.parametrize('n', [1, 3])
def test_fused_mix_loss_matches_oracle(ops, mirror, n):
    """Mix({Dice_loss_jointv2, BCE_Lossv2}) (common/loss.py:9-90) as three kernels: value and gradient against the oracle's
    restatement to 1e-6, the reference-shaped Mix module takes the fused path on CUDA tensors, saturated predictions hit
    torch's clamps (log >= -100, p(1-p) >= 1e-12)."""
    from oracle import fusion_fpn_oracle as O
    g = torch.Generator().manual_seed(23 + n)
    pred = torch.rand(4, n, 16, 1, 64, generator=g)
    pred[0, 0, 0, 0, :4] = torch.tensor([0.0, 1.0, 1e-30, 1 - 1e-7])
    mask = (torch.rand(4, n, 16, 1, 64, generator=g) > 0.5).float()
    pr = pred.clone().requires_grad_(True)
    ref = O.mix_loss(pr, mask)
    (ref * 0.5).backward()
    crit = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask'),
                            'BCE': mirror.loss.BCE_Lossv2('prediction', 'mask')})
    pc = pred.clone().cuda().requires_grad_(True)
    n0 = __import__('ffpn').lib.launch_count(0)
    total, parts = crit({'mask': mask.cuda()}, {'prediction': pc})
    (total * 0.5).backward()
    torch.cuda.synchronize()
    assert __import__('ffpn').lib.launch_count(0) - n0 == 3            # partial sums, finalize, backward
    assert abs(total.item() - ref.item()) <= 1e-6
    assert abs(parts['Dice'].item() - O.dice_loss(pred, mask).item()) <= 1e-6
    assert abs(parts['BCE'].item() - O.bce_loss(pred, mask).item()) <= 1e-6 * max(1.0, O.bce_loss(pred, mask).item())
    assert torch.allclose(pc.grad.cpu(), pr.grad, rtol=1e-5, atol=1e-9)
    # generic path (not the training pair): unchanged semantics
    only = mirror.loss.Mix({'Dice': mirror.loss.Dice_loss_jointv2('prediction', 'mask')})
    t2, _ = only({'mask': mask.cuda()}, {'prediction': pred.cuda()})
    assert abs(t2.item() - O.dice_loss(pred, mask).item()) <= 1e-6


def test_pack_volume_and_sgd(ops):
    from oracle import fusion_fpn_oracle as O
    g = torch.Generator().manual_seed(19)
    img = torch.randn(2, 1, 3, 37, 45, generator=g)
    for dt in DT:
        out = ops.pack_volume(img.cuda(), dt)
        assert torch.equal(out.float().cpu(), img.permute(0, 1, 2, 4, 3).to(dt).float())
    n = 10007
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g)
    params, bufs = {'p': p.clone()}, {}
    pc, mom = p.clone().cuda(), torch.zeros(n, device='cuda')
    for step in range(3):
        O.sgd_step(params, {'p': gr * (step + 1)}, bufs, 0.1, 0.9, 1e-4)
        ops.sgd_step(pc, (gr * (step + 1) * 4).cuda(), mom, 0.1, 0.9, 1e-4, 0.25, step == 0)
    assert rel(pc.cpu(), params['p']) <= 1e-6


def test_zscore_bscans_and_device_dice(ops):
    """GPU input pipeline / metric kernels against the reference's formulas restated in numpy: ZScoreNormalization(axis=(2,3))
    (mytransforms.py:277-296: population std, eps 1e-8 added to the std) and metrics.Dice.calculate_batch (metrics.py:232-253:
    thresholds 0.5, per sample, 1 where both are empty)."""
    from ffpn.pipeline import DeviceDice, GpuInputPipeline
    g = torch.Generator().manual_seed(31)
    img = (torch.randn(2, 1, 5, 37, 45, generator=g) * 800 + 9000)
    img[1, 0, 3] = 7.0                                          # a constant B-scan: std 0 -> (x - mean) / 1e-8 = 0
    x = img.numpy()[:, 0]                                       # the dataloader's [1, S, H, W] per sample
    want = np.stack([(v - v.mean(axis=(1, 2), keepdims=True)) / (v.std(axis=(1, 2), keepdims=True) + 1e-8) for v in x])[:, None]
    got = ops.zscore_bscans(img.cuda())
    assert np.allclose(got.cpu().numpy(), want, rtol=1e-4, atol=2e-4)
    pipe = GpuInputPipeline()
    host = {'image': img.pin_memory(), 'mask': (torch.rand(2, 1, 5, 1, 45, generator=g) > 0.5).float().pin_memory()}
    for _ in range(3):                                          # both buffer sets, reuse
        pipe.prepare(host)
        dev = pipe.get()
        assert np.allclose(dev['image'].cpu().numpy(), want, rtol=1e-4, atol=2e-4) and torch.equal(dev['mask'].cpu(), host['mask'])
        pipe.release()
    pred = torch.rand(3, 2, 6, 1, 40, generator=g)
    mask = (torch.rand(3, 2, 6, 1, 40, generator=g) > 0.6).float()
    pred[2], mask[2] = 0.1, 0.0                                 # empty prediction and mask -> 1
    for ch in (0, 1):
        p, t = (pred[:, ch] > 0.5).float().view(3, -1), (mask[:, ch] > 0.5).float().view(3, -1)
        num, den = (p * t).sum(1).numpy(), (p + t).sum(1).numpy()
        r = 2 * num / np.where(den == 0, 1, den)
        r[den == 0] = 1
        m = DeviceDice('prediction', 'mask', slice=ch)
        m.update({'mask': mask.cuda()}, {'prediction': pred.cuda()})
        m.update({'mask': mask.cuda()}, {'prediction': pred.cuda()})
        assert np.allclose(m.accumulator[0].cpu().numpy(), r, atol=1e-6)
        assert abs(m.get() - float(np.nanmean(np.concatenate([r, r])))) <= 1e-6
        m.reset()
        assert m.accumulator == []
