"""Graph-replayed timing of the latency-bound BatchNorm finalize kernels (rows = 148 partial rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
from ffpn import ops
for C in (16, 64, 256):
    partial = torch.randn(1184 * 2 * C, device='cuda')
    g, b = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
    rm, rv = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
    mean, inv = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
    fns = {'bn_finalize': lambda: ops.bn_finalize(partial, 148, 1e6, g, b, rm, rv, 0.1, 1e-5, True),
           'bn_bwd_finalize': lambda: ops.bn_bwd_finalize(partial, 148, 2, 1, 1e6, g, mean, inv)}
    for name, fn in fns.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(50):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record()
        torch.cuda.synchronize()
        print(f'C {C:3d} {name}: {e0.elapsed_time(e1) / 50 * 1e3:.2f} us per launch (graph replay, back to back)')
