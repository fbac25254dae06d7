"""Named single-kernel workloads at the C2 (bench) shapes: ONE table used by bench.py (CUDA-event timing -> the `kernels`
list and `roofline`), by the ncu captures committed under profiles/ and by tools/ncu_collect.py (which turns the .ncu-rep
files into profiles/ncu_traffic.json, keyed by the same case names).

    python tools/kernel_cases.py <case> [iters]        # run one case (what ncu wraps)
    python tools/kernel_cases.py --list

Each case: fn() launching the kernel(s), algorithmic bytes and flops per call (SURVEY.md section 8d conventions: activations
read + written once at bf16, weights / BatchNorm vectors ignored), the roofline that bounds it, and a regex selecting the
dominant kernel for ncu -k.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'multimodal-fusion-fpn_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

B, S = 8, 32                                   # C2: batch 8, 32 B-scans
CH = [16, 32, 64, 128, 256]
SL = [32, 32, 32, 16, 8]                       # B-scans per level
EF = [128, 64, 32, 16, 8]                      # W = H per level


def build(torch, ops):
    """-> {name: dict(fn, bytes, flops, bound, kernel)} (tensors are created lazily per case)."""
    dt = torch.bfloat16
    g = torch.Generator(device='cuda').manual_seed(0)
    cases = {}

    def act(level, C=None, H=None):
        C = CH[level - 1] if C is None else C
        H = EF[level - 1] if H is None else H
        return torch.randn(B, SL[level - 1], EF[level - 1], H, C, device='cuda', generator=g).to(dt)

    def vec(C):
        return torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')

    def conv_case(name, level, kind, kernel, pad, stride=(1, 1, 1), cin=None, cout=None, H=None, note=''):
        def make():
            Cin = CH[level - 1] if cin is None else cin
            Cout = CH[level - 1] if cout is None else cout
            x = act(level, Cin, H)
            w = torch.randn(Cout, Cin, *kernel, device='cuda', generator=g) * 0.05
            sc, sh = vec(Cin)
            oS = (x.shape[1] + 2 * pad[0] - kernel[0]) // stride[0] + 1
            oW = (x.shape[2] + 2 * pad[1] - kernel[1]) // stride[1] + 1
            oH = (x.shape[3] + 2 * pad[2] - kernel[2]) // stride[2] + 1
            yshape = (B, oS, oW, oH, Cout)
            npos_out = B * oS * oW * oH
            taps = kernel[0] * kernel[1] * kernel[2]
            nbytes = (x.numel() + npos_out * Cout) * 2
            flops = 2.0 * npos_out * taps * Cin * Cout
            if kind == 'fwd':
                fn = lambda: ops.conv_fwd(x, w, kernel, stride, pad, sc, sh, True)
            elif kind == 'dgrad':
                dy = torch.randn(yshape, device='cuda', generator=g).to(dt)
                fn = lambda: ops.conv_dgrad(dy, w, tuple(x.shape), kernel, stride, pad)
            elif kind == 'dgrad_add':
                # dgrad with the residual-branch gradient added in the epilogue: reads dy and the addend, writes dx
                dy = torch.randn(yshape, device='cuda', generator=g).to(dt)
                add = torch.randn_like(x)
                fn = lambda: ops.conv_dgrad(dy, w, tuple(x.shape), kernel, stride, pad, addend=add)
                nbytes = (2 * x.numel() + npos_out * Cout) * 2
            elif kind == 'dgrad_bnr':
                # dgrad + BatchNorm-backward sums of the input's BN in the epilogue: reads dy and the BN's raw input, writes dx
                dy = torch.randn(yshape, device='cuda', generator=g).to(dt)
                fn = lambda: ops.conv_dgrad_bnr(dy, w, x, sc, sh, kernel, stride, pad)
                nbytes = (2 * x.numel() + npos_out * Cout) * 2
            else:
                dy = torch.randn(yshape, device='cuda', generator=g).to(dt)
                dw = torch.zeros_like(w)
                fn = lambda: ops.conv_wgrad(x, dy, w.shape, kernel, stride, pad, sc, sh, True, out=dw)
            return fn, nbytes, flops
        ai = None
        cases[name] = dict(make=make, kernel='conv_wgrad_ws_kernel' if kind == 'wgrad' else 'conv_ws_kernel', note=note)

    for lvl in (1, 2, 3, 4, 5):
        C = CH[lvl - 1]
        conv_case(f'conv_fwd_l{lvl}', lvl, 'fwd', (1, 3, 3), (0, 1, 1), note=f'(1,3,3) {C}->{C}, BN+ReLU on load, statistics epilogue')
        conv_case(f'conv_dgrad_l{lvl}', lvl, 'dgrad', (1, 3, 3), (0, 1, 1), note=f'(1,3,3) {C}->{C}')
        conv_case(f'conv_wgrad_l{lvl}', lvl, 'wgrad', (1, 3, 3), (0, 1, 1), note=f'(1,3,3) {C}->{C}, + partial-tile reduce')
        conv_case(f'conv_dgrad_add_l{lvl}', lvl, 'dgrad_add', (1, 3, 3), (0, 1, 1), note=f'(1,3,3) {C}->{C} dgrad + residual-branch gradient added in the epilogue')
        conv_case(f'conv_dgrad_bnr_l{lvl}', lvl, 'dgrad_bnr', (1, 3, 3), (0, 1, 1), note=f'(1,3,3) {C}->{C} dgrad + ReLU mask + BatchNorm-backward sums of the input BN (replaces dgrad + bn_bwd_reduce)')
    conv_case('proj_conv_l1', 1, 'fwd', (1, 1, 3), (0, 0, 1), stride=(1, 1, 2), note='projection (1,1,3) s(1,1,2) 16->16 on the pair view')
    conv_case('proj_wgrad_l1', 1, 'wgrad', (1, 1, 3), (0, 0, 1), stride=(1, 1, 2), note='projection wgrad')
    conv_case('up4_fwd', 4, 'fwd', (3, 3, 1), (1, 1, 0), cin=768, cout=128, H=1, note='up_concat4 first conv (3,3,1) 768->128 @16x16 en-face')
    conv_case('up4_dgrad', 4, 'dgrad', (3, 3, 1), (1, 1, 0), cin=768, cout=128, H=1, note='up_concat4 dgrad')
    conv_case('up4_wgrad', 4, 'wgrad', (3, 3, 1), (1, 1, 0), cin=768, cout=128, H=1, note='up_concat4 wgrad')

    def simple(name, kernel, make, note=''):
        cases[name] = dict(make=make, kernel=kernel, note=note)

    def mk_block_end_fwd():
        x, y = act(1), act(1)
        sc, sh = vec(16)
        return (lambda: ops.block_end_fwd(y, sc, sh, x)), 3 * x.numel() * 2, 4.0 * x.numel()
    simple('block_end_fwd_l1', 'block_end_fwd_kernel', mk_block_end_fwd, 'BN apply + residual + ReLU, level 1')

    def mk_block_end_bwd():
        z, y, dz = act(1), act(1), act(1)
        dzp = torch.randn(B, 32, 64, 64, 16, device='cuda', generator=g).to(dt)
        return (lambda: ops.block_end_bwd(dz, dzp, z, y, None, (1, 2, 2))), (4 * z.numel() + dzp.numel()) * 2, 6.0 * z.numel()
    simple('block_end_bwd_l1', 'block_end_bwd', mk_block_end_bwd, 'ReLU bwd + pool routing + BN-bwd sums, level 1')

    def mk_block_end_bwd_plain():
        z, y, dz, yres = act(1), act(1), act(1), act(1)
        return (lambda: ops.block_end_bwd(dz, None, z, y, yres, None)), 5 * z.numel() * 2, 8.0 * z.numel()
    simple('block_end_bwd_plain_l1', 'block_end_bwd_kernel', mk_block_end_bwd_plain,
           'ReLU bwd + BN-bwd sums of the block\'s last BN and of its shortcut BN, no pooled branch, level 1')

    def mk_bn_bwd_reduce():
        x, y = act(1), act(1)
        sc, sh = vec(16)
        return (lambda: ops.bn_bwd_reduce(y, x, sc, sh, True)), 2 * x.numel() * 2, 4.0 * x.numel()
    simple('bn_bwd_reduce_l1', 'bn_bwd_reduce_kernel', mk_bn_bwd_reduce, 'BatchNorm backward sums, level 1')

    def mk_bn_bwd_apply():
        x, y = act(1), act(1)
        sc, sh = vec(16)
        cA, cP = vec(16)
        cQ = torch.zeros(16, device='cuda')
        return (lambda: ops.bn_bwd_apply(y, x, sc, sh, True, cA, cP, cQ, out=y)), 3 * x.numel() * 2, 5.0 * x.numel()
    simple('bn_bwd_apply_l1', 'bn_bwd_apply_kernel', mk_bn_bwd_apply, 'BatchNorm backward apply (in place), level 1')

    def mk_proj_tail_fwd():
        y = torch.randn(B, 32, 128, 5, 16, device='cuda', generator=g).to(dt)       # level-1 projection tail: depth 128/16 - 3 = 5
        sc, sh = vec(16)
        return (lambda: ops.proj_tail_fwd(y, sc, sh)), (y.numel() + y.numel() // 5) * 2, 3.0 * y.numel()
    simple('proj_tail_fwd_l1', 'proj_tail_fwd_kernel', mk_proj_tail_fwd, 'BN + ReLU + mean over depth (5 taps), level 1')

    def mk_proj_tail_bwd():
        shape = (B, 32, 128, 5, 16)
        dout = torch.randn(B, 32, 128, 1, 16, device='cuda', generator=g).to(dt)
        n = B * 32 * 128 * 5 * 16
        return (lambda: ops.proj_tail_bwd(dout, shape)), (n + n // 5) * 2, 1.0 * n
    simple('proj_tail_bwd_l1', 'proj_tail_bwd_kernel', mk_proj_tail_bwd, 'mean backward (broadcast / depth), level 1')

    def mk_upsample():
        x = torch.randn(B, 32, 64, 1, 32, device='cuda', generator=g).to(dt)         # up1: level-2 map (1,2,1) -> level-1 grid
        out = torch.empty(B, 32, 128, 1, 64, device='cuda', dtype=dt)
        return (lambda: ops.upsample_fwd(x, 1, 2, out=out, coff=32)), 3 * x.numel() * 2, 0.0
    simple('upsample_into_slot_l1', 'upsample_fwd_kernel', mk_upsample, 'nearest upsample (1,2,1) written into the concat slot of up_concat1')

    def mk_resize():
        x = torch.randn(B, 320, 128, 1, 16, device='cuda', generator=g).to(dt)       # conv1_2d (320x128) -> en-face 32x128, adaptive max
        out = torch.empty(B, 32, 128, 1, 64, device='cuda', dtype=dt)
        return (lambda: ops.resize2d_fwd(x, 32, 128, '2d_max', out=out, coff=16)), (x.numel() + x.numel() // 10) * 2, 1.0 * x.numel()
    simple('resize2d_max_into_slot_l1', 'resize2d_fwd_kernel', mk_resize, 'adaptive max 320x128 -> 32x128 into the concat slot')

    def mk_stem_wgrad():
        x = torch.randn(B, 32, 128, 128, 1, device='cuda', generator=g).to(dt)       # the OCT volume, one channel
        dy = act(1)
        dw = torch.zeros(16, 1, 1, 3, 3, device='cuda')
        return (lambda: ops.conv_wgrad(x, dy, dw.shape, (1, 3, 3), (1, 1, 1), (0, 1, 1), out=dw)), (x.numel() + dy.numel()) * 2, 2.0 * dy.numel() * 9
    simple('stem_wgrad_l1', 'stem_wgrad_kernel', mk_stem_wgrad, 'weight gradient of the first conv (Cin = 1, CUDA cores), level 1')

    def mk_stem_fwd():
        x = torch.randn(B, 32, 128, 128, 1, device='cuda', generator=g).to(dt)
        w = torch.randn(16, 1, 1, 3, 3, device='cuda', generator=g) * 0.3
        return (lambda: ops.conv_fwd(x, w, (1, 3, 3), (1, 1, 1), (0, 1, 1))), (x.numel() + 16 * x.numel()) * 2, 2.0 * 16 * x.numel() * 9
    simple('stem_fwd_l1', 'stem_fwd_kernel', mk_stem_fwd, 'first conv (Cin = 1, CUDA cores) + statistics, level 1')

    def mk_maxpool():
        z = act(1)
        return (lambda: ops.maxpool_fwd(z, (1, 2, 2))), (z.numel() + z.numel() // 4) * 2, 1.0 * z.numel()
    simple('maxpool_fwd_l1', 'maxpool_fwd_kernel', mk_maxpool, 'MaxPool3d (1,2,2), level 1')
    return cases


DEFAULT_BENCH = ['conv_fwd_l1', 'proj_conv_l1', 'conv_wgrad_l1', 'conv_dgrad_l1', 'conv_dgrad_add_l2', 'conv_fwd_l2', 'conv_fwd_l3', 'conv_fwd_l4',
                 'conv_fwd_l5', 'up4_fwd', 'conv_wgrad_l4', 'block_end_fwd_l1', 'block_end_bwd_l1', 'block_end_bwd_plain_l1', 'bn_bwd_reduce_l1',
                 'bn_bwd_apply_l1', 'proj_tail_fwd_l1', 'proj_tail_bwd_l1', 'resize2d_max_into_slot_l1', 'upsample_into_slot_l1']
RIDGE_FLOP_PER_BYTE = 212.0                    # MEASURED_PEAKS: 1389 TFLOP/s sustained / 6.55 TB/s


def main():
    import torch
    from ffpn import ops
    cases = build(torch, ops)
    if len(sys.argv) < 2 or sys.argv[1] == '--list':
        for k, c in cases.items():
            print(f'{k:28s} -k regex:{c["kernel"]:24s} {c["note"]}')
        return
    name = sys.argv[1]
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    fn, nbytes, flops = cases[name]['make']()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f'{name}: {us:.1f} us per call (eager, incl. per-call weight packing for convs), {nbytes / us / 1e3:.0f} GB/s, '
          f'{flops / us / 1e6:.1f} TFLOP/s, {flops / max(nbytes, 1):.0f} FLOP/B')


if __name__ == '__main__':
    main()
