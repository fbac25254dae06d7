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
elif kind == 'dgrad133add':
    w = torch.randn(C, C, 1, 3, 3, device='cuda', generator=g) * 0.1
    add = torch.randn_like(x)
    fn = lambda: ops.conv_dgrad(x, w, tuple(x.shape), (1, 3, 3), (1, 1, 1), (0, 1, 1), addend=add)
elif kind in ('dgrad111', 'dgrad113', 'dgrad131'):
    k = {'dgrad111': (1, 1, 1), 'dgrad113': (1, 1, 3), 'dgrad131': (1, 3, 1)}[kind]
    pd = tuple((v - 1) // 2 for v in k)
    w = torch.randn(C, C, *k, device='cuda', generator=g) * 0.1
    fn = lambda: ops.conv_dgrad(x, w, tuple(x.shape), k, (1, 1, 1), pd)
elif kind == 'wgrad133':
    w = torch.randn(C, C, 1, 3, 3, device='cuda', generator=g) * 0.1
    dy = torch.randn_like(x)
    fn = lambda: ops.conv_wgrad(x, dy, w.shape, (1, 3, 3), (1, 1, 1), (0, 1, 1), sc, sh, True)
elif kind == 'proj':
    w = torch.randn(C, C, 1, 1, 3, device='cuda', generator=g) * 0.1
    fn = lambda: ops.conv_fwd(x, w, (1, 1, 3), (1, 1, 2), (0, 0, 1), sc, sh, True)
elif kind == 'proj_dgrad':
    w = torch.randn(C, C, 1, 1, 3, device='cuda', generator=g) * 0.1
    dyp = torch.randn(B, S, W, H // 2, C, device='cuda', generator=g).to(torch.bfloat16)
    fn = lambda: ops.conv_dgrad(dyp, w, tuple(x.shape), (1, 1, 3), (1, 1, 2), (0, 0, 1))
elif kind == 'proj_wgrad':
    w = torch.randn(C, C, 1, 1, 3, device='cuda', generator=g) * 0.1
    dyp = torch.randn(B, S, W, H // 2, C, device='cuda', generator=g).to(torch.bfloat16)
    fn = lambda: ops.conv_wgrad(x, dyp, w.shape, (1, 1, 3), (1, 1, 2), (0, 0, 1), sc, sh, True)
for _ in range(3):
    fn()
torch.cuda.synchronize()
if os.environ.get('PROF_NO_GRAPH'):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / iters
else:
    # capture `iters` back-to-back calls in a CUDA graph: pure device time, no Python / launch gaps
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(iters):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / iters
print(f'level {level} {kind}: {per * 1e3:.1f} us per call, {2 * x.numel() * 2 / (per * 1e-3) / 1e9:.0f} GB/s (in+out)')
