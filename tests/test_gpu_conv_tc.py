"""The warp-specialised tcgen05/TMEM implicit-GEMM convolutions (csrc/conv_ws.cu, conv_wgrad_ws.cu) against torch's fp32
convolution on the same bf16-rounded inputs: forward (with the fused BN+ReLU prologue and the BatchNorm partial sums), dgrad
(with the fused residual addend) and the weight gradient.  impl=2 forces the tensor-core kernels and errors if they do not
take the geometry; the cases in GENERIC (positions not a multiple of the flat 1x1x1 tile, channel counts off the 16-grid,
odd line splits) are the ones the library routes to its CUDA-core kernels instead, and are tested on that route.
Tolerance: rel-L2 <= 1e-2 against the unrounded reference, <= 5e-4 against the reference rounded like the stored result."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def qbf(t):
    return t.to(torch.bfloat16).float()


def phys(x):
    return x.permute(0, 2, 3, 4, 1).contiguous()


def logical(p):
    return p.permute(0, 4, 1, 2, 3)


CASES = [
    # name, Cin, Cout, kernel, pad, (B, S, W, H)
    ('tiny133', 16, 16, (1, 3, 3), (0, 1, 1), (1, 2, 8, 16)),
    ('l1_133', 16, 16, (1, 3, 3), (0, 1, 1), (2, 3, 128, 128)),
    ('l2_133_widen', 16, 32, (1, 3, 3), (0, 1, 1), (2, 3, 64, 64)),
    ('l3_133', 64, 64, (1, 3, 3), (0, 1, 1), (2, 4, 32, 32)),
    ('l4_133', 128, 128, (1, 3, 3), (0, 1, 1), (2, 4, 16, 16)),
    ('l5_133', 256, 256, (1, 3, 3), (0, 1, 1), (2, 3, 8, 8)),
    ('l1_311', 16, 16, (3, 1, 1), (1, 0, 0), (2, 8, 16, 128)),
    ('l3_311', 64, 64, (3, 1, 1), (1, 0, 0), (2, 12, 32, 32)),
    ('l5_311', 256, 256, (3, 1, 1), (1, 0, 0), (2, 8, 8, 8)),
    ('sc_111', 16, 32, (1, 1, 1), (0, 0, 0), (2, 3, 20, 24)),            # 2880 positions: ragged tail of the flat tiles
    ('sc_111_tiny', 64, 32, (1, 1, 1), (0, 0, 0), (1, 2, 5, 7)),           # 70 positions: less than one TMA box
    ('sc_111_big', 128, 256, (1, 1, 1), (0, 0, 0), (1, 4, 8, 8)),
    # flat mode with > 2^16 positions (regression: reciprocal division in the epilogue must stay exact)
    ('sc_111_p160k', 16, 32, (1, 1, 1), (0, 0, 0), (8, 320, 64, 1)),
    ('proj_114', 64, 64, (1, 1, 4), (0, 0, 0), (2, 4, 16, 8)),
    ('proj_114_l5', 256, 256, (1, 1, 4), (0, 0, 0), (2, 4, 8, 8)),
    ('dec_331', 96, 32, (3, 3, 1), (1, 1, 0), (2, 16, 24, 1)),
    ('dec_331_up4', 768, 128, (3, 3, 1), (1, 1, 0), (2, 8, 16, 1)),
    ('dec_111_up4', 768, 128, (1, 1, 1), (0, 0, 0), (2, 8, 16, 1)),
    ('enc2d_13', 32, 32, (1, 3, 1), (0, 1, 0), (2, 40, 64, 1)),
    ('enc2d_31', 64, 64, (3, 1, 1), (1, 0, 0), (2, 40, 32, 1)),
    ('odd_133', 32, 48, (1, 3, 3), (0, 1, 1), (1, 2, 31, 62)),
    # lines wider than one TMA box (C3: depth 496): X is cut into segments (even split: also the 5-D dy map of the wgrad;
    # uneven split: tail segment validity in the fwd/dgrad epilogue)
    ('wide_133_496', 16, 16, (1, 3, 3), (0, 1, 1), (1, 2, 12, 496)),
    ('wide_133_301', 32, 32, (1, 3, 3), (0, 1, 1), (1, 2, 9, 301)),
    # projection: depth-strided convs (de-interleaved residue planes) and strided 1x1x1 shortcuts
    ('proj_s2_l1', 16, 16, (1, 1, 3), (0, 0, 1), (2, 3, 16, 128), (1, 1, 2)),
    ('proj_s2_l3', 64, 64, (1, 1, 3), (0, 0, 1), (2, 3, 8, 32), (1, 1, 2)),
    ('proj_s2_62', 32, 32, (1, 1, 3), (0, 0, 1), (1, 2, 4, 62), (1, 1, 2)),
    ('proj_s2_l4', 128, 128, (1, 1, 3), (0, 0, 1), (2, 4, 16, 16), (1, 1, 2)),
    # strided 1x1x1 shortcuts: a flat conv over a strided view of the input (tensor-map row pitch = stride * Cin)
    ('sc_s16', 16, 16, (1, 1, 1), (0, 0, 0), (2, 4, 8, 128), (1, 1, 16)),
    ('sc_s8', 32, 32, (1, 1, 1), (0, 0, 0), (2, 4, 8, 64), (1, 1, 8)),
    ('sc_s2', 128, 128, (1, 1, 1), (0, 0, 0), (2, 4, 8, 16), (1, 1, 2)),
    ('sc_s16_496', 16, 16, (1, 1, 1), (0, 0, 0), (1, 8, 32, 496), (1, 1, 16)),
]


# (case, which calls) the tcgen05 kernels decline -> CUDA-core route (impl 0); everything else must run on tensor cores (impl 2)
GENERIC = {'odd_133': ('wgrad',), 'wide_133_301': ('wgrad',)}


def _impl(name, call):
    return 0 if call in GENERIC.get(name, ()) else 2


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_conv_tc(case):
    from ffpn import lib, ops
    torch.backends.cudnn.allow_tf32 = False
    name, cin, cout, k, p, (B, S, W, H) = case[:6]
    s1 = case[6] if len(case) > 6 else (1, 1, 1)
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(len(name) * 131 + cin)
    x = torch.randn(B, cin, S, W, H, generator=g).cuda()
    w = (torch.randn(cout, cin, *k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5).cuda()
    sc = (0.5 + torch.rand(cin, generator=g)).cuda()
    sh = (0.3 * torch.randn(cin, generator=g)).cuda()
    xq = x.to(dt).float()
    old = ops.get_conv_impl()
    try:
        for affine in (False, True):
            xin = torch.relu(xq * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)) if affine else xq
            ref = F.conv3d(xin, w.to(dt).float(), None, s1, p)
            ops.set_conv_impl(_impl(name, 'fwd'))
            y, partial, rows = ops.conv_fwd(phys(x).to(dt), w, k, s1, p, sc if affine else None, sh if affine else None, affine)
            torch.cuda.synchronize()
            yl = logical(y.float())
            assert yl.shape == ref.shape
            assert rel(yl, ref) <= 1e-2, ('fwd', affine, rel(yl, ref))
            # tight form: torch fed the SAME bf16-rounded operands (incl. the re-rounded BN+ReLU prologue
            # output), result rounded to bf16 like the stored one -- what is left is fp32 summation order and the rounding ties it
            # flips (measured <= 2e-4 over all cases, profiles/r02_diag_tight.txt)
            xin_q = qbf(torch.relu(torch.addcmul(sh.view(1, -1, 1, 1, 1), xq, sc.view(1, -1, 1, 1, 1)))) if affine else xq
            w_used = w if _impl(name, 'fwd') == 0 else qbf(w)        # the CUDA-core route reads the fp32 master weights as they are
            ref_q = qbf(F.conv3d(xin_q.double(), w_used.double(), None, s1, p).float())
            assert rel(yl, ref_q) <= 5e-4, ('fwd tight', affine, rel(yl, ref_q))
            st = partial.view(-1, 2, cout)[:rows].double().sum(0)
            ys = y.float().double().reshape(-1, cout)
            assert torch.allclose(st[0], ys.sum(0), rtol=1e-3, atol=1e-3 * ys.abs().sum(0).max().item())
            assert torch.allclose(st[1], (ys * ys).sum(0), rtol=1e-3)
            # run-to-run determinism (also a race detector for the mbarrier pipelines): same call, bitwise-equal output
            y1, partial1, rows1 = ops.conv_fwd(phys(x).to(dt), w, k, s1, p, sc if affine else None, sh if affine else None, affine)
            torch.cuda.synchronize()
            assert torch.equal(y, y1)
            if s1 == (1, 1, 1) and cin in (16, 32, 64, 128, 256, 768) and name != 'sc_111':
                assert rows1 == rows and torch.equal(partial[:rows * 2 * cout], partial1[:rows * 2 * cout])
            # one-call conv + BatchNorm finalize: default (separate finalize kernel) and fused into the conv's last CTA
            # (FFPN_FUSED_FIN=1: measured slower on B200, so it is off by default, but it must stay correct)
            gm, bt = (0.5 + torch.rand(cout, generator=g)).cuda(), (0.2 * torch.randn(cout, generator=g)).cuda()
            ref_aff = ops.bn_finalize(partial, rows, y.numel() // cout, gm, bt, torch.zeros(cout).cuda(), torch.ones(cout).cuda(),
                                      0.1, 1e-5, True)
            for fused in ('0', '1'):
                os.environ['FFPN_FUSED_FIN'] = fused
                try:
                    rm, rv = torch.zeros(cout).cuda(), torch.ones(cout).cuda()
                    yb, aff = ops.conv_fwd_bn(phys(x).to(dt), w, k, s1, p, sc if affine else None, sh if affine else None, affine,
                                              gm, bt, rm, rv, 0.1, 1e-5, True)
                finally:
                    os.environ.pop('FFPN_FUSED_FIN', None)
                torch.cuda.synchronize()
                assert torch.equal(yb, y)
                for a_, b_ in zip(aff, ref_aff):
                    assert torch.allclose(a_, b_, rtol=1e-5, atol=1e-6), (fused, (a_ - b_).abs().max())
            ops.set_conv_impl(1)
            y2, _, _ = ops.conv_fwd(phys(x).to(dt), w, k, s1, p, sc if affine else None, sh if affine else None, affine)
            assert rel(y.float(), y2.float()) <= 1e-2
        dy = torch.randn(ref.shape, generator=g).cuda().to(dt)
        xr = xq.clone().requires_grad_(True)
        F.conv3d(xr, w.to(dt).float(), None, s1, p).backward(dy.float())
        add = torch.randn(B, cin, S, W, H, generator=g).cuda().to(dt)
        ops.set_conv_impl(_impl(name, 'dgrad'))
        dx = ops.conv_dgrad(phys(dy), w, tuple(phys(x).shape), k, s1, p)
        dx2 = ops.conv_dgrad(phys(dy), w, tuple(phys(x).shape), k, s1, p, addend=phys(add))
        torch.cuda.synchronize()
        assert rel(logical(dx.float()), xr.grad) <= 1e-2, ('dgrad', rel(logical(dx.float()), xr.grad))
        assert rel(logical(dx2.float()), xr.grad + add.float()) <= 1e-2
        # dgrad fused with pass 1 of the backward of the BatchNorm + ReLU that produced the conv's input (ffpn_conv_dgrad_bnr): the
        # stored gradient is the plain dgrad, bit for bit; the sums are those of bn_bwd_reduce on it (ReLU mask applied)
        yprev = phys(torch.randn(B, cin, S, W, H, generator=g).cuda()).to(dt)
        # (default: dgrad + bn_bwd_reduce as two launches; FFPN_WS_BNR_MAXPOS switches the one-kernel epilogue variant on,
        # measured slower on B200 -- both must give the same numbers)
        G0, part0, prow0 = ops.conv_dgrad_bnr(phys(dy), w, yprev, sc, sh, k, s1, p)
        os.environ['FFPN_WS_BNR_MAXPOS'] = str(1 << 40)
        try:
            n0 = lib.launch_count()
            G, part, prow = ops.conv_dgrad_bnr(phys(dy), w, yprev, sc, sh, k, s1, p)
            fused = lib.launch_count() - n0 <= 2                    # (a per-call weight-packing launch + the conv kernel)
            G1, part1, prow1 = ops.conv_dgrad_bnr(phys(dy), w, yprev, sc, sh, k, s1, p)
        finally:
            os.environ.pop('FFPN_WS_BNR_MAXPOS', None)
        torch.cuda.synchronize()
        if s1 == (1, 1, 1) and cin % 16 == 0:
            assert fused, 'the one-kernel dgrad + BatchNorm-backward-sums path declined a stride-1 geometry'
        assert torch.equal(G0, dx)
        s0 = part0.view(-1, 2, cin)[:prow0].double().sum(0)
        mask = torch.addcmul(sh, yprev.float(), sc) > 0
        assert torch.equal(G, dx)
        Gd, yd_ = torch.where(mask, dx.float(), torch.zeros((), device='cuda')).double().reshape(-1, cin), yprev.double().reshape(-1, cin)
        st = part.view(-1, 2, cin)[:prow].double().sum(0)
        assert torch.allclose(st[0], Gd.sum(0), rtol=1e-3, atol=1e-3 * Gd.abs().sum(0).max().item()), ('dgrad_bnr sum G', (st[0] - Gd.sum(0)).abs().max())
        assert torch.allclose(st[1], (Gd * yd_).sum(0), rtol=1e-3, atol=1e-3 * (Gd * yd_).abs().sum(0).max().item())
        assert torch.equal(G, G1) and prow == prow1 and torch.equal(part[:prow * 2 * cin], part1[:prow * 2 * cin])
        assert torch.allclose(s0, st, rtol=1e-4, atol=1e-4 * Gd.abs().sum(0).max().item())     # two launches == one launch
        # wgrad on tensor cores (fused BN+ReLU prologue on x), against torch autograd
        for affine in (False, True):
            xin = torch.relu(xq * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1)) if affine else xq
            wr = w.clone().requires_grad_(True)
            F.conv3d(xin, wr, None, s1, p).backward(dy.float())
            # e.g. an odd line of 301 positions has no equal segments: no tensor-core weight gradient, the library picks the CUDA-core one
            ops.set_conv_impl(_impl(name, 'wgrad'))
            dw = ops.conv_wgrad(phys(x).to(dt), phys(dy), w.shape, k, s1, p, sc if affine else None,
                                sh if affine else None, affine)
            torch.cuda.synchronize()
            assert rel(dw, wr.grad) <= 1e-2, ('wgrad', affine, rel(dw, wr.grad))
            if s1 == (1, 1, 1) and cin in (16, 32, 64, 128, 256, 768) and cout in (16, 32, 64, 128, 256) and name not in ('sc_111', 'wide_133_301'):
                dw1 = ops.conv_wgrad(phys(x).to(dt), phys(dy), w.shape, k, s1, p, sc if affine else None,
                                     sh if affine else None, affine)
                assert torch.equal(dw, dw1)                # fixed-order partial-tile reduction: bitwise reproducible
    finally:
        ops.set_conv_impl(old)


@pytest.mark.parametrize('cin,cout', [(16, 16), (16, 32), (32, 32)])
def test_conv_pair_view_stride1(cin, cout):
    """FFPN_WS_PAIR2=1: narrow (1,3,3) convs on the pair view of input and output (remapped weights, all-zero K halves of the
    outer pair taps skipped, statistics folded).  Off by default (measured slower) but kept correct:
    forward (+BN prologue, statistics) and dgrad (+addend) against torch, and against the default path."""
    from ffpn import ops
    torch.backends.cudnn.allow_tf32 = False
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(100 + cin + cout)
    B, S, W, H = 2, 3, 20, 64
    x = torch.randn(B, cin, S, W, H, generator=g).cuda()
    w = (torch.randn(cout, cin, 1, 3, 3, generator=g) / (cin * 9) ** 0.5).cuda()
    sc, sh = (0.5 + torch.rand(cin, generator=g)).cuda(), (0.3 * torch.randn(cin, generator=g)).cuda()
    xq = x.to(dt).float()
    xin = torch.relu(xq * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))
    ref = F.conv3d(xin, w.to(dt).float(), None, 1, (0, 1, 1))
    dy = torch.randn(ref.shape, generator=g).cuda().to(dt)
    xr = xq.clone().requires_grad_(True)
    F.conv3d(xr, w.to(dt).float(), None, 1, (0, 1, 1)).backward(dy.float())
    add = torch.randn(B, cin, S, W, H, generator=g).cuda().to(dt)
    outs = {}
    old = ops.get_conv_impl()
    try:
        ops.set_conv_impl(2)
        for flag in ('0', '1'):
            os.environ['FFPN_WS_PAIR2'] = flag
            y, partial, rows = ops.conv_fwd(phys(x).to(dt), w, (1, 3, 3), (1, 1, 1), (0, 1, 1), sc, sh, True)
            dx = ops.conv_dgrad(phys(dy), w, tuple(phys(x).shape), (1, 3, 3), (1, 1, 1), (0, 1, 1), addend=phys(add))
            torch.cuda.synchronize()
            assert rel(logical(y.float()), ref) <= 1e-2, (flag, rel(logical(y.float()), ref))
            assert rel(logical(dx.float()), xr.grad + add.float()) <= 1e-2
            st = partial.view(-1, 2, cout)[:rows].double().sum(0)
            ys = y.float().double().reshape(-1, cout)
            assert torch.allclose(st[0], ys.sum(0), rtol=1e-3, atol=1e-3 * ys.abs().sum(0).max().item())
            assert torch.allclose(st[1], (ys * ys).sum(0), rtol=1e-3)
            outs[flag] = (y.float(), dx.float())
    finally:
        os.environ.pop('FFPN_WS_PAIR2', None)
        ops.set_conv_impl(old)
    assert rel(outs['1'][0], outs['0'][0]) <= 4e-3 and rel(outs['1'][1], outs['0'][1]) <= 4e-3
