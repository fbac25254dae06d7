"""conv_fwd_bn (finalize fused into the conv kernel) against conv_fwd + bn_finalize at the C2 level shapes."""
import os, sys
os.environ.setdefault("FFPN_FUSED_FIN", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-fusion-fpn_b200'))
import torch
from ffpn import ops
g = torch.Generator(device='cuda').manual_seed(0)
for level in (1, 2, 3, 4, 5):
    C = [16, 32, 64, 128, 256][level - 1]
    B, S = 8, [32, 32, 32, 16, 8][level - 1]
    W = H = 128 >> (level - 1)
    x = torch.randn(B, S, W, H, C, device='cuda', generator=g).to(torch.bfloat16)
    w = torch.randn(C, C, 1, 3, 3, device='cuda', generator=g) * 0.1
    sc, sh = torch.rand(C, device='cuda') + 0.5, torch.randn(C, device='cuda') * 0.1
    gm, bt = torch.rand(C, device='cuda') + 0.5, torch.randn(C, device='cuda') * 0.1
    for rep in range(3):
        rm1, rv1 = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
        rm2, rv2 = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
        y1, partial, rows = ops.conv_fwd(x, w, (1, 3, 3), (1, 1, 1), (0, 1, 1), sc, sh, True)
        aff1 = ops.bn_finalize(partial, rows, y1.numel() // C, gm, bt, rm1, rv1, 0.1, 1e-5, True)
        y2, aff2 = ops.conv_fwd_bn(x, w, (1, 3, 3), (1, 1, 1), (0, 1, 1), sc, sh, True, gm, bt, rm2, rv2, 0.1, 1e-5, True)
        torch.cuda.synchronize()
        errs = [float((a - b).abs().max()) for a, b in zip(aff1, aff2)] + [float((rm1 - rm2).abs().max()), float((rv1 - rv2).abs().max())]
        print(f'level {level} rep {rep}: y equal {torch.equal(y1, y2)}; max abs diff scale/shift/mean/invstd/rmean/rvar = ' + ' '.join(f'{e:.2e}' for e in errs), flush=True)
