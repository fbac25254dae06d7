"""Run single conv launches at a C2-level shape for ncu (python tools/prof_conv.py <level> <kind>)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
from ffpn import ops

level = int(sys.argv[1]) if len(sys.argv) > 1 else 1
kind = sys.argv[2] if len(sys.argv) > 2 else 'fwd133'
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
C = [16, 32, 64, 128, 256][level - 1]
B, S = 8, [32, 32, 32, 16, 8][level - 1]
W = H = 128 >> (level - 1)
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(B, S, W, H, C, device='cuda', generator=g).to(torch.bfloat16)
sc, sh = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
if kind == 'fwd133':
    w = torch.randn(C, C, 1, 3, 3, device='cuda', generator=g) * 0.1
    fn = lambda: ops.conv_fwd(x, w, (1, 3, 3), (1, 1, 1), (0, 1, 1), sc, sh, True)
elif kind == 'fwd311':
    w = torch.randn(C, C, 3, 1, 1, device='cuda', generator=g) * 0.1
    fn = lambda: ops.conv_fwd(x, w, (3, 1, 1), (1, 1, 1), (1, 0, 0), sc, sh, True)
elif kind == 'dgrad133':
    w = torch.randn(C, C, 1, 3, 3, device='cuda', generator=g) * 0.1
    fn = lambda: ops.conv_dgrad(x, w, tuple(x.shape), (1, 3, 3), (1, 1, 1), (0, 1, 1))
elif kind == 'wgrad133':
    w = torch.randn(C, C, 1, 3, 3, device='cuda', generator=g) * 0.1
    dy = torch.randn_like(x)
    fn = lambda: ops.conv_wgrad(x, dy, w.shape, (1, 3, 3), (1, 1, 1), (0, 1, 1), sc, sh, True)
elif kind == 'proj':
    w = torch.randn(C, C, 1, 1, 3, device='cuda', generator=g) * 0.1
    fn = lambda: ops.conv_fwd(x, w, (1, 1, 3), (1, 1, 2), (0, 0, 1), sc, sh, True)
for _ in range(iters):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fn()
e1.record()
torch.cuda.synchronize()
print(f'level {level} {kind}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us per call, '
      f'{2 * x.numel() * 2 / (e0.elapsed_time(e1) / iters * 1e-3) / 1e9:.0f} GB/s (in+out)')
