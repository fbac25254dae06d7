"""Role ablation of the warp-specialised conv kernel in ONE process: FFPN_TC_DEBUG bits switch off the MMA issue (1),
the epilogue body (2), the TMA loads (4), the BN transform (8), the output stores (16), the statistics (32); each
setting is captured as a CUDA graph of `iters` launches and timed.  Results are wrong by construction -- timing only.
usage: python tools/ablate_conv.py [levels] [kinds] [dbg,dbg,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
from ffpn import ops

levels = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '1,2,3').split(',')]
kinds = (sys.argv[2] if len(sys.argv) > 2 else 'fwd133').split(',')
dbgs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else '0,14,13,3,11,7,12,2,1').split(',')]
iters = 20
names = {1: 'mma', 2: 'epi', 4: 'tma', 8: 'xf', 16: 'st', 32: 'stat'}
for level in levels:
    C = [16, 32, 64, 128, 256][level - 1]
    B, S = 8, [32, 32, 32, 16, 8][level - 1]
    W = H = 128 >> (level - 1)
    g = torch.Generator(device='cuda').manual_seed(0)
    x = torch.randn(B, S, W, H, C, device='cuda', generator=g).to(torch.bfloat16)
    sc, sh = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
    w = torch.randn(C, C, 1, 3, 3, device='cuda', generator=g) * 0.1
    add = torch.randn_like(x)
    for kind in kinds:
        if kind == 'fwd133':
            fn = lambda: ops.conv_fwd(x, w, (1, 3, 3), (1, 1, 1), (0, 1, 1), sc, sh, True)
        elif kind == 'fwd133stats':
            gam, bet, rm, rv = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
            fn = lambda: ops.conv_fwd_bn(x, w, (1, 3, 3), (1, 1, 1), (0, 1, 1), sc, sh, True, gam, bet, rm, rv, 0.1, 1e-5, True)
        elif kind == 'proj':
            wp = torch.randn(C, C, 1, 1, 3, device='cuda', generator=g) * 0.1
            fn = lambda: ops.conv_fwd(x, wp, (1, 1, 3), (1, 1, 2), (0, 0, 1), sc, sh, True)
        elif kind == 'fwd311':
            w3 = torch.randn(C, C, 3, 1, 1, device='cuda', generator=g) * 0.1
            fn = lambda: ops.conv_fwd(x, w3, (3, 1, 1), (1, 1, 1), (1, 0, 0), sc, sh, True)
        elif kind == 'dgrad133':
            fn = lambda: ops.conv_dgrad(x, w, tuple(x.shape), (1, 3, 3), (1, 1, 1), (0, 1, 1))
        elif kind == 'dgrad133add':
            fn = lambda: ops.conv_dgrad(x, w, tuple(x.shape), (1, 3, 3), (1, 1, 1), (0, 1, 1), addend=add)
        else:
            raise SystemExit(kind)
        out = []
        for d in dbgs:
            if d >= 1000:                                   # 1128 / 1160: full kernel with FFPN_WS_NT = 128 / 160 transform threads
                os.environ['FFPN_WS_NT'] = str(d - 1000)
                d = 0
            os.environ['FFPN_TC_DEBUG'] = str(d)
            fn(); torch.cuda.synchronize()
            side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(iters):
                    fn()
            graph.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / iters * 1e3
            off = '+'.join(n for b, n in names.items() if d & b) or 'none'
            out.append(f'off[{off}] nt{os.environ.get("FFPN_WS_NT", "128")} {us:.1f}')
        os.environ['FFPN_TC_DEBUG'] = '0'
        print(f'level {level} {kind}: ' + ' | '.join(out), flush=True)
