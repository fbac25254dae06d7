"""Full-size en-face (2-D encoder level 1/2) conv shapes: WS kernels vs torch fp32 reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
import torch.nn.functional as F
from ffpn import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
phys = lambda t: t.permute(0, 2, 3, 4, 1).contiguous()
logical = lambda p: p.permute(0, 4, 1, 2, 3)
g = torch.Generator(device='cuda').manual_seed(0)
cases = [('13_16', 16, 16, (1, 3, 1), (0, 1, 0), (8, 320, 128, 1)), ('31_16', 16, 16, (3, 1, 1), (1, 0, 0), (8, 320, 128, 1)),
         ('13_16_32', 16, 32, (1, 3, 1), (0, 1, 0), (8, 320, 64, 1)), ('11_16_32', 16, 32, (1, 1, 1), (0, 0, 0), (8, 320, 64, 1)),
         ('13_32', 32, 32, (1, 3, 1), (0, 1, 0), (8, 320, 64, 1)), ('31_32', 32, 32, (3, 1, 1), (1, 0, 0), (8, 320, 64, 1)),
         ('31_64', 64, 64, (3, 1, 1), (1, 0, 0), (8, 320, 32, 1)), ('13_64', 64, 64, (1, 3, 1), (0, 1, 0), (8, 320, 32, 1))]
ops.set_conv_impl(2)
for name, cin, cout, k, p, (B, S, W, H) in cases:
    x = torch.randn(B, cin, S, W, H, device='cuda', generator=g)
    w = torch.randn(cout, cin, *k, device='cuda', generator=g) / (cin * k[0] * k[1]) ** 0.5
    sc = torch.rand(cin, device='cuda', generator=g) + 0.5
    sh = torch.randn(cin, device='cuda', generator=g) * 0.3
    xq = x.bfloat16().float()
    xin = torch.relu(xq * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))
    wq = w.bfloat16().float()
    ref = F.conv3d(xin, wq, None, 1, p)
    y, partial, rows = ops.conv_fwd(phys(x).bfloat16(), w, k, (1, 1, 1), p, sc, sh, True)
    dy = torch.randn(ref.shape, device='cuda', generator=g).bfloat16()
    xr = xq.clone().requires_grad_(True)
    F.conv3d(xr, wq, None, 1, p).backward(dy.float())
    add = torch.randn(B, cin, S, W, H, device='cuda', generator=g).bfloat16()
    dx = ops.conv_dgrad(phys(dy), w, tuple(phys(x).shape), k, (1, 1, 1), p, addend=phys(add))
    wr = w.clone().requires_grad_(True)
    F.conv3d(xin, wr, None, 1, p).backward(dy.float())
    dw = ops.conv_wgrad(phys(x).bfloat16(), phys(dy), w.shape, k, (1, 1, 1), p, sc, sh, True)
    dw2 = ops.conv_wgrad(phys(x).bfloat16(), phys(dy), w.shape, k, (1, 1, 1), p, sc, sh, True)
    torch.cuda.synchronize()
    print(f'{name}: fwd {rel(logical(y.float()), ref):.2e} dgrad {rel(logical(dx.float()), xr.grad + add.float()):.2e} '
          f'wgrad {rel(dw, wr.grad):.2e} finite {bool(torch.isfinite(dw).all())} deterministic {torch.equal(dw, dw2)}', flush=True)
