"""Diagnostic: every conv kernel case of tests/test_gpu_conv_tc.py against torch fed the SAME bf16-rounded operands with the
output rounded to bf16 too (what is left is fp32 summation order + the rounding ties it flips: expect <= ~2e-4), plus the
BatchNorm partial sums against sums of the stored values.  Prints one line per case; usage: python tools/diag_tight.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
from test_gpu_conv_tc import CASES, phys, logical, rel
from ffpn import ops

torch.backends.cudnn.allow_tf32 = False
q = lambda t: t.to(torch.bfloat16).float()
for case in CASES:
    name, cin, cout, k, p, (B, S, W, H) = case[:6]
    s1 = case[6] if len(case) > 6 else (1, 1, 1)
    g = torch.Generator().manual_seed(len(name) * 131 + cin)
    x = torch.randn(B, cin, S, W, H, generator=g).cuda()
    w = (torch.randn(cout, cin, *k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5).cuda()
    sc = (0.5 + torch.rand(cin, generator=g)).cuda()
    sh = (0.3 * torch.randn(cin, generator=g)).cuda()
    xq = q(x)
    out = [f'{name:14s}']
    for impl in (2, 0):
        ops.set_conv_impl(impl)
        for affine in (False, True):
            xin = q(torch.relu(torch.addcmul(sh.view(1, -1, 1, 1, 1), xq, sc.view(1, -1, 1, 1, 1)))) if affine else xq
            ref = q(F.conv3d(xin.double(), q(w).double(), None, s1, p).float())
            try:
                y, partial, rows = ops.conv_fwd(phys(x).to(torch.bfloat16), w, k, s1, p, sc if affine else None, sh if affine else None, affine)
            except Exception as e:
                out.append(f'impl{impl} aff{int(affine)} n/a')
                continue
            torch.cuda.synchronize()
            yl = logical(y.float())
            st = partial.view(-1, 2, cout)[:rows].double().sum(0)
            ys = y.float().double().reshape(-1, cout)
            e_sum = ((st[0] - ys.sum(0)).abs().max() / ys.abs().sum(0).max()).item()
            e_sq = ((st[1] - (ys * ys).sum(0)).abs() / (ys * ys).sum(0)).max().item()
            out.append(f'impl{impl} aff{int(affine)} y {rel(yl, ref):.1e} maxabs {(yl - ref).abs().max().item():.1e} sum {e_sum:.1e} sq {e_sq:.1e}')
    print(' | '.join(out), flush=True)
ops.set_conv_impl(0)
