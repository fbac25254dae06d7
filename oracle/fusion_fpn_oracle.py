"""CPU oracle for the projective multimodal fusion FPN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product path (``multimodal-fusion-fpn_b200/``) never
imports this module and fails loudly when its CUDA library is missing.

What it is: a *functional* restatement (no ``nn.Module``; weights come from a
``state_dict`` keyed exactly like the reference's) of the forward pass of
``FPNHybridFusion`` and its sibling wirings, written with plain ``torch`` CPU ops
(fp32 or fp64) plus numpy for the integer index tables.  Gradients come from
``torch.autograd`` over this functional graph.  Each function cites the reference
``file:line`` it restates (paths relative to ``/root/reference``).

Parity pin: the reference ships no golden vectors, so the
oracle is pinned against the *reference itself*, imported unmodified in the build
container by ``tests/golden/make_golden.py``; the resulting fixtures live in
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this file
against them.  The third-party arithmetic that cannot be imported
(pytorch-lightning 1.5.10's DP loss/grad averaging, ``requirements.txt:47``) is
restated in ``dp_average_gradients`` and is "parity unpinned".
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

BN_EPS = 1e-5          # torch.nn.BatchNorm default, used by every BN in fusion3D2D.py
BN_MOMENTUM = 0.1
CHANNELS = (16, 32, 64, 128, 256)   # models/fpn/modifiedUnet3D_red-convPlusFully_dropout00.ini:4


# --------------------------------------------------------------------------------------
# integer / index oracles (numpy)
# --------------------------------------------------------------------------------------

def nearest_index_table(n_in: int, scale: float) -> np.ndarray:
    """Index table of Upsample_Custom3d_nearest for one axis.

    models/fpn/components.py:261-265: ``ceil(arange(1, 1+int(n_in*scale)) / scale) - 1``.
    """
    n_out = int(n_in * scale)
    return (np.ceil(np.arange(1, 1 + n_out) / scale) - 1).astype(np.int64)


def adaptive_window(i: int, n_in: int, n_out: int) -> Tuple[int, int]:
    """Window [start, end) of F.adaptive_max_pool3d along one axis (fusion3D2D.py:559-564).

    start = floor(i*in/out), end = ceil((i+1)*in/out)  (SURVEY.md App. B).
    """
    start = (i * n_in) // n_out
    end = -((-(i + 1) * n_in) // n_out)
    return start, end


def maxpool_argmax_firstmax(x: np.ndarray, kernel: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """nn.MaxPool3d(kernel) (stride=kernel, floor) with PyTorch's tie rule, pure loops.

    fusion3D2D.py:87-90.  Tie -> first element in row-major window scan; NaN wins and
    propagates, and with several NaNs in a window the LAST one is indexed (checked against torch).  ``x`` is (S, W, H); returns (values, flat indices into
    S*W*H).  Small cases only.
    """
    kS, kW, kH = kernel
    S, W, H = x.shape
    oS, oW, oH = S // kS, W // kW, H // kH
    val = np.empty((oS, oW, oH), x.dtype)
    idx = np.empty((oS, oW, oH), np.int64)
    for s in range(oS):
        for w in range(oW):
            for h in range(oH):
                best = None
                bi = -1
                for ds in range(kS):
                    for dw in range(kW):
                        for dh in range(kH):
                            ss, ww, hh = s * kS + ds, w * kW + dw, h * kH + dh
                            v = x[ss, ww, hh]
                            if best is None or v > best or np.isnan(v):      # aten rule: (val > max) || isnan(val)
                                best, bi = v, (ss * W + ww) * H + hh
                val[s, w, h] = best
                idx[s, w, h] = bi
    return val, idx


def adaptive_maxpool2d_argmax(x: np.ndarray, out_size: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """F.adaptive_max_pool3d on a (S', W', 1) map -> (S, W, 1), first-max rule, pure loops.

    fusion3D2D.py:559-564.  ``x`` is (S', W'); returns (values, flat indices into S'*W').
    """
    Si, Wi = x.shape
    So, Wo = out_size
    val = np.empty((So, Wo), x.dtype)
    idx = np.empty((So, Wo), np.int64)
    for s in range(So):
        s0, s1 = adaptive_window(s, Si, So)
        for w in range(Wo):
            w0, w1 = adaptive_window(w, Wi, Wo)
            best = None
            bi = -1
            for ss in range(s0, s1):
                for ww in range(w0, w1):
                    v = x[ss, ww]
                    if best is None or v > best or np.isnan(v):      # aten rule: (val > max) || isnan(val)
                        best, bi = v, ss * Wi + ww
            val[s, w] = best
            idx[s, w] = bi
    return val, idx


def projection_depth(h: int, n_red: int, final_kernel: int = 4) -> int:
    """Depth left after a projection chain (fusion3D2D.py:296-377)."""
    for _ in range(n_red):
        h = (h - 1) // 2 + 1            # (1,1,3) stride 2 pad 1
    return h - (final_kernel - 1)       # (1,1,4) no pad


# --------------------------------------------------------------------------------------
# float building blocks
# --------------------------------------------------------------------------------------

class BNRecorder:
    """Collects the running-stat updates a train-mode forward would have applied."""

    def __init__(self) -> None:
        self.updates: Dict[str, Tensor] = {}


def batch_norm(x: Tensor, sd: StateDict, prefix: str, train: bool,
               rec: Optional[BNRecorder] = None) -> Tensor:
    """nn.BatchNorm2d/3d restated.

    Train: per-channel batch mean and *biased* variance over all non-channel axes;
    running stats take momentum 0.1 with the *unbiased* variance.  Eval: running stats.
    """
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    red = [d for d in range(x.dim()) if d != 1]
    shape = [1, -1] + [1] * (x.dim() - 2)
    if train:
        mean = x.mean(dim=red)
        var = ((x - mean.view(shape)) ** 2).mean(dim=red)
        if rec is not None:
            n = x.numel() // x.shape[1]
            with torch.no_grad():
                rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
                rec.updates[prefix + '.running_mean'] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach().to(rm.dtype)
                rec.updates[prefix + '.running_var'] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * (var.detach() * n / max(n - 1, 1)).to(rv.dtype)
                rec.updates[prefix + '.num_batches_tracked'] = sd[prefix + '.num_batches_tracked'] + 1
    else:
        mean = sd[prefix + '.running_mean'].to(x.dtype)
        var = sd[prefix + '.running_var'].to(x.dtype)
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mean.view(shape)) * (inv * w).view(shape) + b.view(shape)


# ---- bf16-storage emulation -------------------------------------------------------------------------------------
# The CUDA path stores activations (and the packed conv weights) as bf16 and accumulates in fp32.  ``bf16_storage()``
# makes this oracle round at exactly the points where that path stores: the input images, every raw conv output (the
# BatchNorm statistics are then taken from the rounded values), the BN+ReLU'd operand a conv stages, each block output,
# the projection's depth mean, the resized 2-D feature; conv weights except the Cin == 1 stems (kept fp32 by the stem
# kernels) and the head.  Rounding is straight-through for autograd (gradients are those of the rounded forward).
# This is SURVEY.md App. D.1's "torch op fed the same bf16-rounded inputs" at model level: what is left between this
# and the kernels is fp32 summation order (and the rounding ties it flips), not storage precision.
_EMULATE_BF16 = False


class bf16_storage:
    def __enter__(self):
        global _EMULATE_BF16
        self.prev, _EMULATE_BF16 = _EMULATE_BF16, True
        return self

    def __exit__(self, *exc):
        global _EMULATE_BF16
        _EMULATE_BF16 = self.prev


def _q(t: Tensor) -> Tensor:
    if not _EMULATE_BF16:
        return t
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def _conv(x: Tensor, w: Tensor, stride, padding) -> Tensor:
    if w.shape[1] != 1:
        w = _q(w)
    y = F.conv3d(x, w, None, stride, padding) if w.dim() == 5 else F.conv2d(x, w, None, stride, padding)
    return _q(y)


def convx_block(x: Tensor, sd: StateDict, prefix: str, strides: Sequence, paddings: Sequence,
                is_residual: bool, train: bool, rec: Optional[BNRecorder] = None,
                shortcut_stride=None, store_out: bool = True) -> Tensor:
    """unet3dConvX / unet2dConvX forward (fusion3D2D.py:717-732, :878-893).

    k x [conv(bias=False) -> BN -> ReLU] with the last one conv -> BN only (:597-648);
    residual = downsample(x) (1x1x1 conv + BN, :231-233 / :317-326 / :924-933) when the
    state_dict has ``<prefix>.downsample.0.weight`` else x; ``out += residual``; ReLU.
    Dropout rates are all 0.0 in the .ini, so there is no dropout.
    """
    k = len(strides)
    out = x
    for i in range(k):
        p = f'{prefix}.convBlock.{i}'
        out = _conv(out, sd[p + '.0.weight'], strides[i], paddings[i])
        out = batch_norm(out, sd, p + '.1', train, rec)
        if i < k - 1:
            out = _q(torch.relu(out))
    if is_residual:
        res = x
        if prefix + '.downsample.0.weight' in sd:
            wd = sd[prefix + '.downsample.0.weight']
            st = shortcut_stride if shortcut_stride is not None else 1
            res = _conv(x, wd, st, 0)
            res = batch_norm(res, sd, prefix + '.downsample.1', train, rec)
        out = out + res
    out = torch.relu(out)
    return _q(out) if store_out else out


def encoder_level_3d(x: Tensor, sd: StateDict, prefix: str, train: bool,
                     rec: Optional[BNRecorder] = None) -> Tensor:
    """_make_layer_2plus3 (fusion3D2D.py:226-258): two residual blocks."""
    x = convx_block(x, sd, prefix + '.0', [(1, 1, 1)] * 2, [(0, 1, 1)] * 2, True, train, rec)
    x = convx_block(x, sd, prefix + '.1', [(1, 1, 1)] * 3, [(0, 1, 1), (0, 1, 1), (1, 0, 0)], True, train, rec)
    return x


def encoder_level_2d(x: Tensor, sd: StateDict, prefix: str, train: bool,
                     rec: Optional[BNRecorder] = None) -> Tensor:
    """_make_layer_2plus3_2d (fusion3D2D.py:261-293)."""
    x = convx_block(x, sd, prefix + '.0', [(1, 1)] * 2, [(0, 1)] * 2, True, train, rec)
    x = convx_block(x, sd, prefix + '.1', [(1, 1)] * 3, [(0, 1), (0, 1), (1, 0)], True, train, rec)
    return x


def projection_block(x: Tensor, sd: StateDict, prefix: str, n_red: int, train: bool,
                     rec: Optional[BNRecorder] = None, take_mean: bool = True) -> Tensor:
    """_make_zdimReductionConvPlusFully + torch.mean (fusion3D2D.py:296-377, :527-536).

    n_red x (1,1,3) stride (1,1,2) pad (0,0,1) convs with a 1x1x1 stride (1,1,2**n_red)
    shortcut, then a non-residual (1,1,4) block, then the mean over what is left of depth.
    """
    if n_red > 0:
        x = convx_block(x, sd, prefix + '.0', [(1, 1, 2)] * n_red, [(0, 0, 1)] * n_red, True, train, rec,
                        shortcut_stride=(1, 1, 2 ** n_red))
        x = convx_block(x, sd, prefix + '.1', [(1, 1, 1)], [(0, 0, 0)], False, train, rec, store_out=not take_mean)
    else:
        x = convx_block(x, sd, prefix + '.0', [(1, 1, 1)], [(0, 0, 0)], False, train, rec, store_out=not take_mean)
    if take_mean:
        x = _q(x.mean(dim=4, keepdim=True))        # the kernels fuse BN + ReLU + mean and store only the mean
    return x


def upsample_nearest(x: Tensor, factor: Sequence[float]) -> Tensor:
    """Upsample_Custom3d_nearest.forward (components.py:259-268): three chained gathers."""
    d = torch.from_numpy(nearest_index_table(x.shape[-3], factor[-3]))
    r = torch.from_numpy(nearest_index_table(x.shape[-2], factor[-2]))
    c = torch.from_numpy(nearest_index_table(x.shape[-1], factor[-1]))
    return x[:, :, d, :, :][:, :, :, r, :][:, :, :, :, c]


def resize_2d_feature(f2d: Tensor, size: Sequence[int], interpolate: Optional[str]) -> Tensor:
    """2-D feature -> '2-D in 3-D' and resize to the en-face grid (fusion3D2D.py:544-564)."""
    f = f2d[:, :, :, :, None]
    if interpolate == '2d':
        f = _q(F.interpolate(f, size=tuple(size), mode='trilinear'))
    elif interpolate == '2d_max':
        f = F.adaptive_max_pool3d(f, output_size=tuple(size))
    return f


def up_block(inputs: List[Tensor], deeper: Tensor, sd: StateDict, prefix: str, upfactor, train: bool,
             rec: Optional[BNRecorder] = None, feature_fusion: str = 'concat') -> Tensor:
    """unet3dUp2modified.forward, 3-input (fusion3D2D.py:956-966) and 2-input
    (components.py:72-76) forms, plus the 'add' variant (fusion3D2D.py:1028-1039)."""
    up = upsample_nearest(deeper, upfactor)
    if feature_fusion == 'add' and len(inputs) == 2:
        inputs = [inputs[0] + inputs[1]]
    cat = torch.cat(list(inputs) + [up], 1)
    return convx_block(cat, sd, prefix + '.conv', [(1, 1, 1)] * 2, [(1, 1, 0)] * 2, True, train, rec)


POOLS_3D = ((1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2))   # fusion3D2D.py:87-90
POOLS_2D = ((1, 2), (1, 2), (2, 2), (2, 2))                # fusion3D2D.py:168-171
UPFACTORS = {4: (2, 2, 1), 3: (2, 2, 1), 2: (1, 2, 1), 1: (1, 2, 1)}   # fusion3D2D.py:185-220


def fusion_body_forward(sd: StateDict, oct: Tensor, slo: Tensor, interpolate: Optional[str],
                        train: bool = True, prefix: str = 'resensnet', level5: bool = True,
                        rec: Optional[BNRecorder] = None, feature_fusion: str = 'concat',
                        stages: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ModifiedUnet3D2DLevel5.forward (fusion3D2D.py:499-581) / ModifiedUnet3D2D.forward (:380-469).

    oct: (B,1,S,W,H) depth last; slo: (B,1,S',W').  Returns (B,n_classes,S,W,1).
    ``stages`` (optional dict) receives every named intermediate for stage-level parity.
    """
    P = prefix + '.' if prefix else ''
    n2d = 5 if level5 else 4
    f2d, x = [], _q(slo)
    for l in range(1, n2d + 1):
        x = encoder_level_2d(x, sd, f'{P}conv{l}_2d', train, rec)
        f2d.append(x)
        if l < n2d:
            x = F.max_pool2d(x, POOLS_2D[l - 1])
    f3d, x = [], _q(oct)
    for l in range(1, 6):
        x = encoder_level_3d(x, sd, f'{P}conv{l}', train, rec)
        f3d.append(x)
        if l < 5:
            x = F.max_pool3d(x, POOLS_3D[l - 1])
    proj = [projection_block(f3d[l - 1], sd, f'{P}zdimRed{l}', 5 - l, train, rec) for l in range(1, 6)]
    r2d = [resize_2d_feature(f2d[l], proj[l].shape[2:], interpolate) for l in range(n2d)]
    if stages is not None:
        for l in range(5):
            stages[f'conv{l+1}'] = f3d[l]
            stages[f'proj{l+1}'] = proj[l]
        for l in range(n2d):
            stages[f'conv{l+1}_2d'] = f2d[l]
            stages[f'res{l+1}_2d'] = r2d[l]
    deeper = torch.cat([proj[4], r2d[4]], 1) if level5 else proj[4]      # fusion3D2D.py:572
    for l in (4, 3, 2, 1):
        deeper = up_block([proj[l - 1], r2d[l - 1]], deeper, sd, f'{P}up_concat{l}', UPFACTORS[l], train, rec,
                          feature_fusion)
        if stages is not None:
            stages[f'up{l}'] = deeper
    out = F.conv3d(deeper, sd[P + 'final1.weight'], sd[P + 'final1.bias'])   # fusion3D2D.py:223,579
    return out


def fpn_hybrid_fusion_forward(sd: StateDict, batch: Dict[str, Tensor], crop: str = 'relative_2d_max',
                              fusion_modality: str = 'slo', train: bool = True, sigmoid: bool = True,
                              rec: Optional[BNRecorder] = None,
                              stages: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
    """FPNHybridFusion.forward (models/fusion_nets.py:99-121); crop -> interpolate (:102-107)."""
    interpolate = '2d' if 'relative_2d' in crop else None
    if 'max' in crop and interpolate is not None:
        interpolate += '_max'
    oct = batch['image'].permute(0, 1, 2, 4, 3)
    slo = batch[fusion_modality][:, :, :, 0, :]
    seg = fusion_body_forward(sd, oct, slo, interpolate, train, rec=rec, stages=stages)
    seg = seg.permute(0, 1, 2, 4, 3)
    if sigmoid:
        seg = torch.sigmoid(seg)
    return {'prediction': seg}


# --------------------------------------------------------------------------------------
# other wirings of the same bodies
# --------------------------------------------------------------------------------------

def unet3d_body_forward(sd: StateDict, oct: Tensor, train: bool = True, prefix: str = 'resensnet',
                        use_1x1: bool = True, rec: Optional[BNRecorder] = None, original: bool = False,
                        classification: bool = False) -> Tensor:
    """ModifiedUnet3D.forward (models/fpn/unets3D.py:441-485).  ``classification`` returns conv5 (:453-454);
    ``original`` keeps the depth axis after the kernel-8 projection tail instead of averaging it (:458-471; the
    tail's kernel size comes with the weights, :79-82)."""
    P = prefix + '.' if prefix else ''
    f3d, x = [], _q(oct)
    for l in range(1, 6):
        x = encoder_level_3d(x, sd, f'{P}conv{l}', train, rec)
        f3d.append(x)
        if l < 5:
            x = F.max_pool3d(x, POOLS_3D[l - 1])
    if classification:
        return f3d[4]
    proj = [projection_block(f3d[l - 1], sd, f'{P}zdimRed{l}', 5 - l, train, rec, take_mean=not original)
            for l in range(1, 6)]
    deeper = proj[4]
    for l in (4, 3, 2, 1):
        deeper = up_block([proj[l - 1]], deeper, sd, f'{P}up_concat{l}', UPFACTORS[l], train, rec)
    if use_1x1:
        deeper = F.conv3d(deeper, sd[P + 'final1.weight'], sd[P + 'final1.bias'])
    return deeper


def unet2d_body_forward(sd: StateDict, img: Tensor, train: bool = True, prefix: str = 'resensnet',
                        level5: bool = True, output_features: bool = False,
                        rec: Optional[BNRecorder] = None) -> Tensor:
    """ModifiedUnet2DLevel5.forward (models/fpn/unets2D.py:172-213) / ModifiedUnet2D.forward (:108-144)."""
    P = prefix + '.' if prefix else ''
    n2d = 5 if level5 else 4
    f2d, x = [], _q(img)
    for l in range(1, n2d + 1):
        x = encoder_level_2d(x, sd, f'{P}conv{l}_2d', train, rec)
        f2d.append(x[:, :, :, :, None])
        if l < n2d:
            x = F.max_pool2d(x, POOLS_2D[l - 1])
    deeper = f2d[-1]
    for l in range(n2d - 1, 0, -1):
        deeper = up_block([f2d[l - 1]], deeper, sd, f'{P}up_concat{l}', UPFACTORS[l], train, rec)
    if output_features:
        return deeper
    return F.conv3d(deeper, sd[P + 'final1.0.weight'], sd[P + 'final1.0.bias'])   # unets2D.py:102-106


# --------------------------------------------------------------------------------------
# loss, optimiser and data-parallel semantics
# --------------------------------------------------------------------------------------

def dice_loss(pred: Tensor, gt: Tensor) -> Tensor:
    """Dice_loss_jointv2.forward (common/loss.py:73-90)."""
    s = gt.shape
    p, g = pred.reshape(s[0], s[1], -1), gt.reshape(s[0], s[1], -1)
    inter = (p * g).sum(dim=(0, 2)) + 1e-6
    union = (p ** 2 + g).sum(dim=(0, 2)) + 2e-6
    return 1.0 - torch.mean(2.0 * inter / union)


def bce_loss(pred: Tensor, gt: Tensor) -> Tensor:
    """BCE_Lossv2.forward (common/loss.py:47-56)."""
    return F.binary_cross_entropy(pred.reshape(-1), gt.reshape(-1), reduction='mean')


def mix_loss(pred: Tensor, gt: Tensor) -> Tensor:
    """Mix.forward with {Dice, BCE}, unit coefficients (common/loss.py:18-28)."""
    return (dice_loss(pred, gt) + bce_loss(pred, gt)) / 2


def sgd_step(params: Dict[str, Tensor], grads: Dict[str, Tensor], bufs: Dict[str, Tensor], lr: float = 0.1,
             momentum: float = 0.9, weight_decay: float = 1e-4) -> None:
    """torch.optim.SGD as configured by train.py:126-133 (no dampening, no nesterov), in place."""
    for k, p in params.items():
        g = grads[k] + weight_decay * p
        if k in bufs:
            bufs[k].mul_(momentum).add_(g)
        else:
            bufs[k] = g.clone()
        p.sub_(lr * bufs[k])


def dp_average_gradients(per_rank_grads: List[Dict[str, Tensor]]) -> Dict[str, Tensor]:
    """DP semantics of train.py:155-167 (strategy='dp', sync_batchnorm=False): each replica has its own
    BN batch statistics and its own loss; Lightning averages the replica losses, hence the gradient is
    the mean of per-replica gradients.  PARITY UNPINNED (pytorch-lightning 1.5.10 is not installed)."""
    n = len(per_rank_grads)
    return {k: sum(g[k] for g in per_rank_grads) / n for k in per_rank_grads[0]}


# --------------------------------------------------------------------------------------
# synthetic batches and a state_dict factory (no nn.Module needed)
# --------------------------------------------------------------------------------------

def synthetic_batch(B: int, S: int, H: int, W: int, S2: int, W2: int, seed: int = 1234,
                    modality: str = 'slo', dtype=torch.float32, smooth: bool = False) -> Dict[str, Tensor]:
    """Synthetic batch with the dataloader's keys/shapes (SURVEY.md section 8d;
    common/dataloader_hrf_spec_aligned_seg.py:56-57,90-112)."""
    g = torch.Generator().manual_seed(seed)
    if smooth:
        coarse = torch.randn(B, 1, max(S // 4, 1), max(H // 8, 1), max(W // 8, 1), generator=g)
        img = F.interpolate(coarse, size=(S, H, W), mode='trilinear', align_corners=False)
    else:
        img = torch.randn(B, 1, S, H, W, generator=g)
    aux = torch.rand(B, 1, S2, 1, W2, generator=g)
    mask = (torch.rand(B, 1, S, 1, W, generator=g) > 0.5).float()
    return {'image': img.to(dtype), modality: aux.to(dtype), 'mask': mask.to(dtype)}


def conv_mac_count(sd: StateDict, S: int, H: int, W: int, S2: int, W2: int) -> int:
    """Forward conv MACs per sample of FPNHybridFusion, from shapes."""
    total = 0
    C = CHANNELS
    sp3 = [(S, W, H), (S, W // 2, H // 2), (S, W // 4, H // 4), (S // 2, W // 8, H // 8), (S // 4, W // 16, H // 16)]
    sp2 = [(S2, W2), (S2, W2 // 2), (S2, W2 // 4), (S2 // 2, W2 // 8), (S2 // 4, W2 // 16)]
    for l in range(5):
        cin = 1 if l == 0 else C[l - 1]
        c = C[l]
        n3 = sp3[l][0] * sp3[l][1] * sp3[l][2]
        n2 = sp2[l][0] * sp2[l][1]
        total += n3 * (cin * c * 9 + c * c * 9 * 3 + c * c * 3 + cin * c)
        total += n2 * (cin * c * 3 + c * c * 3 * 3 + c * c * 3 + cin * c)
        h = sp3[l][2]
        ew = sp3[l][0] * sp3[l][1]
        for _ in range(4 - l):
            h = (h - 1) // 2 + 1
            total += ew * h * c * c * 3
        if l < 4:
            total += ew * h * c * c
        total += ew * (h - 3) * c * c * 4
    for l in (4, 3, 2, 1):
        low = C[4] * 2 if l == 4 else C[l]
        cur = C[l - 1]
        ew = sp3[l - 1][0] * sp3[l - 1][1]
        total += ew * ((low + 2 * cur) * cur * 9 + cur * cur * 9 + (low + 2 * cur) * cur)
    total += sp3[0][0] * sp3[0][1] * C[0]
    return total


# --------------------------------------------------------------------------------------
# deterministic state_dict factory: keys/shapes of FPNHybridFusion
# --------------------------------------------------------------------------------------

def _key_seed(seed: int, key: str) -> int:
    h = 1469598103934665603
    for ch in key.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ seed) & 0x7FFFFFFF


def _add_conv(sd, order, key, shape, seed, dtype):
    fan_in = shape[1] * int(np.prod(shape[2:]))
    fan_out = shape[0] * int(np.prod(shape[2:]))
    std = math.sqrt(2.0 / (fan_in + fan_out))            # xavier_normal_, gain 1 (weight_init.py:19,25)
    g = torch.Generator().manual_seed(_key_seed(seed, key))
    sd[key] = (torch.randn(shape, generator=g) * std).to(dtype)
    order.append(key)


def _add_bn(sd, order, prefix, c, seed, dtype, randomize_running):
    g = torch.Generator().manual_seed(_key_seed(seed, prefix))
    sd[prefix + '.weight'] = (1 + 0.02 * torch.randn(c, generator=g)).to(dtype)   # weight_init.py:44-48
    sd[prefix + '.bias'] = (0.05 * torch.randn(c, generator=g)).to(dtype)          # non-zero on purpose
    if randomize_running:
        sd[prefix + '.running_mean'] = (0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[prefix + '.running_var'] = (1 + 0.2 * torch.rand(c, generator=g)).to(dtype)
    else:
        sd[prefix + '.running_mean'] = torch.zeros(c, dtype=dtype)
        sd[prefix + '.running_var'] = torch.ones(c, dtype=dtype)
    sd[prefix + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)
    order.extend([prefix + s for s in ('.weight', '.bias', '.running_mean', '.running_var', '.num_batches_tracked')])


def _add_convx(sd, order, prefix, cin, cout, kernels, downsample, seed, dtype, rr):
    for i, k in enumerate(kernels):
        ci = cin if i == 0 else cout
        _add_conv(sd, order, f'{prefix}.convBlock.{i}.0.weight', (cout, ci) + tuple(k), seed, dtype)
        _add_bn(sd, order, f'{prefix}.convBlock.{i}.1', cout, seed, dtype, rr)
    if downsample:
        ones = (1,) * len(kernels[0])
        _add_conv(sd, order, f'{prefix}.downsample.0.weight', (cout, cin) + ones, seed, dtype)
        _add_bn(sd, order, f'{prefix}.downsample.1', cout, seed, dtype, rr)


def make_state_dict(seed: int = 1234, dtype=torch.float32, n_classes: int = 1, prefix: str = 'resensnet',
                    randomize_running: bool = False) -> StateDict:
    """state_dict of FPNHybridFusion with reference key names, shapes and *registration order*
    (fusion3D2D.py:51-223 then :482-497), filled deterministically per key (independent of torch's
    global RNG).  ``tests/golden/make_golden.py`` loads it into the unmodified reference with
    ``strict=True`` and asserts the key order, which pins this grammar."""
    sd: StateDict = {}
    order: List[str] = []
    C = CHANNELS
    P = prefix + '.' if prefix else ''
    rr = randomize_running
    for l in range(5):                                                     # conv1..5
        cin = 1 if l == 0 else C[l - 1]
        _add_convx(sd, order, f'{P}conv{l+1}.0', cin, C[l], [(1, 3, 3)] * 2, cin != C[l], seed, dtype, rr)
        _add_convx(sd, order, f'{P}conv{l+1}.1', C[l], C[l], [(1, 3, 3), (1, 3, 3), (3, 1, 1)], False, seed, dtype, rr)
    for l in range(5):                                                     # zdimRed1..5
        n = 4 - l
        if n > 0:
            _add_convx(sd, order, f'{P}zdimRed{l+1}.0', C[l], C[l], [(1, 1, 3)] * n, True, seed, dtype, rr)
            _add_convx(sd, order, f'{P}zdimRed{l+1}.1', C[l], C[l], [(1, 1, 4)], False, seed, dtype, rr)
        else:
            _add_convx(sd, order, f'{P}zdimRed{l+1}.0', C[l], C[l], [(1, 1, 4)], False, seed, dtype, rr)
    for l in range(4):                                                     # conv1..4_2d
        cin = 1 if l == 0 else C[l - 1]
        _add_convx(sd, order, f'{P}conv{l+1}_2d.0', cin, C[l], [(1, 3)] * 2, cin != C[l], seed, dtype, rr)
        _add_convx(sd, order, f'{P}conv{l+1}_2d.1', C[l], C[l], [(1, 3), (1, 3), (3, 1)], False, seed, dtype, rr)
    for l in (4, 3, 2, 1):                                                 # up_concat4..1 (4 is Level5-widened)
        low = C[4] * 2 if l == 4 else C[l]
        cur = C[l - 1]
        _add_convx(sd, order, f'{P}up_concat{l}.conv', low + 2 * cur, cur, [(3, 3, 1)] * 2, True, seed, dtype, rr)
    _add_conv(sd, order, P + 'final1.weight', (n_classes, C[0], 1, 1, 1), seed, dtype)
    g = torch.Generator().manual_seed(_key_seed(seed, P + 'final1.bias'))
    sd[P + 'final1.bias'] = (0.1 * torch.randn(n_classes, generator=g)).to(dtype)
    order.append(P + 'final1.bias')
    _add_convx(sd, order, f'{P}conv5_2d.0', C[3], C[4], [(1, 3)] * 2, True, seed, dtype, rr)   # registered last
    _add_convx(sd, order, f'{P}conv5_2d.1', C[4], C[4], [(1, 3), (1, 3), (3, 1)], False, seed, dtype, rr)
    return {k: sd[k] for k in order}


def fill_like(template: StateDict, seed: int = 1234, dtype=torch.float32, randomize_running: bool = False) -> StateDict:
    """Deterministic weights for ANY state_dict of the model family (other wirings, directly constructed bodies): same
    per-key generators as ``make_state_dict`` (``fill_like(make_state_dict(s), s) == make_state_dict(s)``, tested), keyed
    on the names and shapes of ``template``: >=4-D ``weight`` = conv (xavier normal, weight_init.py:19,25), a
    ``weight`` with a sibling ``running_mean`` = BatchNorm, other 1-D tensors = conv bias."""
    sd: StateDict = {}
    order: List[str] = []
    for key, v in template.items():
        if key in sd:
            continue
        if key.endswith('.weight') and (key[:-7] + '.running_mean') in template:
            _add_bn(sd, order, key[:-7], v.numel(), seed, dtype, randomize_running)
        elif v.dim() >= 4:
            _add_conv(sd, order, key, tuple(v.shape), seed, dtype)
        elif v.is_floating_point():
            g = torch.Generator().manual_seed(_key_seed(seed, key))
            sd[key] = (0.1 * torch.randn(tuple(v.shape), generator=g)).to(dtype)
            order.append(key)
        else:
            raise KeyError(f'fill_like: unexpected entry {key}')
    assert list(template.keys()) == [k for k in template if k in sd]
    return {k: sd[k] for k in template}


def wiring_forward(case: dict, sd: StateDict, batch: Dict[str, Tensor], train: bool = True) -> Tensor:
    """Functional restatement of the other registered wirings / directly constructed bodies, keyed by the case table of
    ``tests/golden/wiring_cases.py``: FPN / FPNRegression (fusion_nets.py:29-50), FPNClassification (:53-77),
    FPNHybridFusionRegression (:125-127), FPN2D (:131-147), FPNLateFusion(+Regression) (:150-222), and the bodies
    ModifiedUnet3D2D (fusion3D2D.py:380-469, feature_fusion 'add' :969-1039), ModifiedUnet3D(original=True)
    (unets3D.py:441-485) and ModifiedUnet2D (unets2D.py:108-144)."""
    oct = batch['image'].permute(0, 1, 2, 4, 3)
    slo = batch['slo'][:, :, :, 0, :]
    crop = case.get('crop', 'oct')
    interp = '2d' if 'relative_2d' in crop else None
    if 'max' in crop and interp is not None:
        interp += '_max'
    if case['kind'] == 'body':
        kw = case['kwargs']
        if case['cls'] in ('ModifiedUnet3D2D', 'ModifiedUnet3D2DLevel5'):
            return fusion_body_forward(sd, oct, slo, kw.get('interpolate'), train, prefix='',
                                       level5=case['cls'].endswith('Level5'), feature_fusion=kw.get('feature_fusion', 'concat'))
        if case['cls'] == 'ModifiedUnet3D':
            return unet3d_body_forward(sd, oct, train, prefix='', original=kw.get('original', False))
        return unet2d_body_forward(sd, slo, train, prefix='', level5=case['cls'].endswith('Level5'),
                                   output_features=kw.get('output_features', False))
    name = case['name']
    if name in ('FPN', 'FPNRegression'):
        seg = unet3d_body_forward(sd, oct, train).permute(0, 1, 2, 4, 3)
        return torch.sigmoid(seg) if name == 'FPN' else seg
    if name == 'FPNClassification':
        feat = unet3d_body_forward(sd, oct, train, classification=True)
        pred = F.conv3d(feat, sd['one_one.weight']).mean(dim=(2, 3, 4))        # AdaptiveAvgPool3d((1,1,1)) + squeezes
        return torch.softmax(pred, dim=-1)
    if name == 'FPNHybridFusionRegression':
        return fpn_hybrid_fusion_forward(sd, batch, crop, 'slo', train, sigmoid=False)['prediction']
    if name == 'FPN2D':
        seg = torch.sigmoid(unet2d_body_forward(sd, slo, train).permute(0, 1, 2, 4, 3))
        if seg.shape != batch['mask'].shape:
            seg = F.interpolate(seg, size=batch['mask'].shape[2:], mode='trilinear')
        return seg
    if name in ('FPNLateFusion', 'FPNLateFusionRegression'):
        a = unet3d_body_forward(sd, oct, train, prefix='resensnet3d', use_1x1=False).permute(0, 1, 2, 4, 3)
        b = unet2d_body_forward(sd, slo, train, prefix='resensnet2d', output_features=True).permute(0, 1, 2, 4, 3)
        if interp == '2d':
            b = F.interpolate(b, size=a.shape[2:], mode='trilinear')
        elif interp == '2d_max':
            b = F.adaptive_max_pool3d(b, output_size=a.shape[2:])
        seg = F.conv3d(torch.cat([a, b], 1), sd['fusion_module.weight'], sd['fusion_module.bias'])
        return torch.sigmoid(seg) if name == 'FPNLateFusion' else seg
    raise KeyError(name)


def param_keys(sd: StateDict) -> List[str]:
    return [k for k in sd if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))]


def loss_and_grads(sd: StateDict, batch: Dict[str, Tensor], crop: str = 'relative_2d_max', modality: str = 'slo',
                   stages: Optional[Dict[str, Tensor]] = None, rec: Optional[BNRecorder] = None):
    """Forward + Mix loss + backward (pl_model_wrapper.py:243-254 training_step semantics)."""
    work = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k in set(param_keys(sd)) else v)
            for k, v in sd.items()}
    out = fpn_hybrid_fusion_forward(work, batch, crop, modality, True, True, rec, stages)
    loss = mix_loss(out['prediction'], batch['mask'])
    keys = param_keys(sd)
    grads = torch.autograd.grad(loss, [work[k] for k in keys])
    return loss.detach(), out['prediction'].detach(), dict(zip(keys, grads))
