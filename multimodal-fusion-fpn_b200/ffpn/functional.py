"""torch.autograd Functions that run the fusion-FPN stages on the sm_100a kernels.

Tensors crossing these Functions are LOGICAL reference-shaped tensors -- (B, C, S, W, H) for 3-D features
and (B, C, S', W') for 2-D ones -- stored channels-last (torch.channels_last_3d / channels_last), so a
caller sees the same shapes the reference produces while the kernels see (B, S, W, H, C).
"""
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import ops

_COMPUTE_DTYPE = torch.bfloat16


def set_compute_dtype(dtype: torch.dtype) -> None:
    """Activation storage type of the kernels: torch.bfloat16 (default, fp32 accumulate) or torch.float32."""
    global _COMPUTE_DTYPE
    assert dtype in (torch.bfloat16, torch.float32)
    _COMPUTE_DTYPE = dtype


def get_compute_dtype() -> torch.dtype:
    return _COMPUTE_DTYPE


# ---- gradient sink -------------------------------------------------------------------------------------
# FusionTrainer keeps all gradients in one flat fp32 buffer that is zeroed once per step.  When it installs the
# sink (parameter data_ptr -> gradient view), the backward kernels write weight / BatchNorm gradients straight into
# that buffer and the Functions return None for them, instead of autograd allocating a tensor per parameter and
# launching one `grad += new` kernel each (~350 tiny launches per step).
_GRAD_SINK = {}


def set_grad_sink(mapping) -> None:
    global _GRAD_SINK
    _GRAD_SINK = dict(mapping) if mapping else {}


def _sink(t):
    return _GRAD_SINK.get(t.data_ptr()) if _GRAD_SINK else None


# ---- branch streams --------------------------------------------------------------------------------------
# The model is a DAG, not a chain: the 2-D encoder is independent of the 3-D one, the projective block of level l is
# independent of encoder levels > l, and most of the step's ~800 launches are small (deep levels, en-face maps: a few
# CTAs each).  Branches are therefore forked onto side streams (event fork / join); autograd replays each node's
# backward on the stream its forward ran on, so the backward gets the mirrored DAG, and a captured CUDA graph keeps the
# parallel edges.  FFPN_STREAMS=0 runs everything on the caller's stream.
_SIDE_STREAMS = {}
_USED_SIDE = set()


def streams_enabled(kind: str = 'branches') -> bool:
    """FFPN_STREAMS: '1' (default) everything, '0' nothing, 'branches' / 'wgrad' one of the two mechanisms."""
    v = os.environ.get('FFPN_STREAMS', '1')
    return v == '1' or v == kind


def side_stream(device, i: int) -> 'torch.cuda.Stream':
    key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), i)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=key[0])
    return st


def fork(side: 'torch.cuda.Stream', *tensors) -> 'torch.cuda.Stream':
    """``side`` continues after everything enqueued so far on the current stream; ``tensors`` (allocated on the current
    stream) will be read there."""
    side.wait_stream(torch.cuda.current_stream())
    _USED_SIDE.add(side)
    for t in tensors:
        t.record_stream(side)
    return side


def join(side: 'torch.cuda.Stream', *tensors) -> None:
    """The current stream continues after ``side``; ``tensors`` (allocated on ``side``) will be read here."""
    cur = torch.cuda.current_stream()
    cur.wait_stream(side)
    for t in tensors:
        t.record_stream(cur)


def join_side_streams() -> None:
    """Join every side stream used since the last call into the current stream (end of backward: the gradient sink is
    written from all of them; a CUDA-graph capture must not end with unjoined work)."""
    cur = torch.cuda.current_stream()
    for st in list(_USED_SIDE):
        if st.device == cur.device:
            cur.wait_stream(st)
            _USED_SIDE.discard(st)


def fork_from(side: 'torch.cuda.Stream', event: 'torch.cuda.Event', *tensors) -> 'torch.cuda.Stream':
    """``side`` continues after ``event`` (recorded earlier on the current stream) instead of after everything enqueued so
    far: a branch issued late in program order still starts early on the device (and in a captured graph)."""
    side.wait_event(event)
    _USED_SIDE.add(side)
    for t in tensors:
        t.record_stream(side)
    return side


def used_side_streams(device):
    return [st for st in _USED_SIDE if st.device == device]


# ---- gradient buckets ------------------------------------------------------------------------------------
# FusionTrainer overlaps the gradient all-reduce + SGD of everything but the first two encoder levels with the backward of
# those levels.  The body marks the boundary in its forward: the marker's backward runs when the
# gradient has flowed back through level 3, i.e. after every kernel writing an "early" gradient has been issued.
_BUCKET_HOOK = [None]


def set_bucket_hook(fn) -> None:
    _BUCKET_HOOK[0] = fn


class _BucketMarker(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if _BUCKET_HOOK[0] is not None:
            _BUCKET_HOOK[0]()
        return g


def bucket_marker(x: torch.Tensor) -> torch.Tensor:
    return _BucketMarker.apply(x) if (_BUCKET_HOOK[0] is not None and x.requires_grad) else x


# ---- layout helpers ------------------------------------------------------------------------------------
def to_phys(x: torch.Tensor) -> torch.Tensor:
    """logical (B,C,S,W,H) / (B,C,S,W) -> physical (B,S,W,H,C) contiguous in the compute dtype."""
    if x.dim() == 5:
        p = x.permute(0, 2, 3, 4, 1)
    elif x.dim() == 4:
        p = x.permute(0, 2, 3, 1).unsqueeze(3)
    else:
        raise ValueError(f'expected a 4-D or 5-D feature tensor, got {tuple(x.shape)}')
    if not p.is_contiguous():
        p = p.contiguous()          # boundary only: a caller handed us a non-channels-last tensor
    if p.dtype != _COMPUTE_DTYPE:
        p = ops.cast(p.float() if p.dtype != torch.float32 else p, _COMPUTE_DTYPE)
    return p


class CatSlot:
    """A channel slice [coff, coff + C) of a pre-allocated concat buffer (physical (B, S, W, 1, Ctot)).  The producers of a
    decoder level's inputs -- the projection tail, the 2-D feature resize, the nearest upsample of the deeper level -- write
    their result straight into their slot, so ``torch.cat([proj3D, feat2D, up], 1)`` (fusion3D2D.py:572,966) is never executed
    as a copy; in backward the consumers' gradient is read back slice by slice, in place."""

    def __init__(self, buf: torch.Tensor, coff: int, C: int):
        self.buf, self.coff, self.C = buf, int(coff), int(C)

    def view(self) -> torch.Tensor:
        """Logical (B, C, S, W, 1) view of the slot (channel-strided, no copy), tagged so that UpCat/Cat recognise it."""
        return to_logical(self.buf[..., self.coff:self.coff + self.C], 5)


def tag_slot(t: torch.Tensor, slot):
    """Mark ``t`` (the tensor an autograd Function returned for a slot) so that cat()/upcat() recognise it as already in place."""
    if slot is not None:
        t._ffpn_slot = slot
    return t


def resize2d(x, size, mode, slot=None):
    """2-D feature -> en-face grid (fusion3D2D.py:544-564), optionally straight into a concat slot."""
    return tag_slot(Resize2DFunction.apply(x, size, mode, slot), slot)


def new_cat_buffer(like_phys_shape, Ctot: int, dtype, device) -> torch.Tensor:
    B, S, W = like_phys_shape[0], like_phys_shape[1], like_phys_shape[2]
    return torch.empty((B, S, W, 1, Ctot), dtype=dtype, device=device)


def phys_slice(x: torch.Tensor):
    """logical (B,C,S,W,1) tensor that is either channels-last contiguous or a channel slice of a wider channels-last tensor
    -> (physical tensor to hand to a kernel, row stride in elements or None when contiguous).  Anything else is made
    contiguous (boundary case)."""
    p = x.permute(0, 2, 3, 4, 1)
    if p.dtype == _COMPUTE_DTYPE and p.shape[3] == 1 and p.stride(4) == 1 and not p.is_contiguous():
        ct = p.stride(2)
        if ct >= p.shape[4] and p.stride(1) == p.shape[2] * ct and p.stride(0) == p.shape[1] * p.shape[2] * ct \
                and (p.storage_offset() * p.element_size()) % 16 == 0 and ct % 8 == 0:
            return p, ct
    return to_phys(x), None


def to_logical(p: torch.Tensor, ndim: int) -> torch.Tensor:
    if ndim == 5:
        return p.permute(0, 4, 1, 2, 3)
    return p.squeeze(3).permute(0, 3, 1, 2)


def k3(t) -> Tuple[int, int, int]:
    t = tuple(int(v) for v in (t if isinstance(t, (tuple, list)) else (t,) * 3))
    return t if len(t) == 3 else (t[0], t[1], 1)       # 2-D (kS', kW') acts on (S, W) with H == 1


def p3(t) -> Tuple[int, int, int]:
    t = tuple(int(v) for v in (t if isinstance(t, (tuple, list)) else (t,) * 3))
    return t if len(t) == 3 else (t[0], t[1], 0)


@dataclass
class ConvXSpec:
    kernels: tuple          # per conv (kS,kW,kH)
    strides: tuple
    pads: tuple
    residual: bool
    has_ds: bool
    ds_stride: tuple
    pool: Optional[tuple]   # fused max-pool kernel or None
    tail: str               # 'relu' (block end) | 'mean' (projection tail: BN+ReLU+mean over depth)
    training: bool
    momentum: float
    eps: float
    need_dx: bool
    ndim: int
    out: Optional[object] = None     # CatSlot: tail == 'mean' writes the projected map straight into a concat buffer


def _wgrad(inp, dy, w_shape, kernel, stride, pad, a_in, b_in, relu, out):
    """Weight gradient of one conv.  With the gradient sink installed nothing downstream in this backward reads the
    result, so the launch (+ its partial-tile reduce) goes to a side stream paired with the current one and leaves the
    dgrad -> BatchNorm-backward chain, which is the critical path; the trainer joins the side streams before the
    optimiser reads the flat gradient."""
    if out is None or not streams_enabled('wgrad'):
        return ops.conv_wgrad(inp, dy, w_shape, kernel, stride, pad, a_in, b_in, relu, out=out)
    cur = torch.cuda.current_stream()
    wg = fork(side_stream(inp.device, ('wgrad', cur.cuda_stream)), inp, dy, *([a_in, b_in] if a_in is not None else []))
    with torch.cuda.stream(wg):
        ops.conv_wgrad(inp, dy, w_shape, kernel, stride, pad, a_in, b_in, relu, out=out)
    return None


class ConvXFunction(torch.autograd.Function):
    """One residual block (unet3dConvX / unet2dConvX, reference fusion3D2D.py:717-732, :878-893):
    k x [conv -> BN -> ReLU] (last without ReLU), optional 1x1x1 conv+BN shortcut, add, ReLU; optionally
    fused with the max-pool that follows it (fusion3D2D.py:515-521) or with the projection's depth mean
    (:527-536).  tensors = per conv (w, gamma, beta, running_mean, running_var), then the shortcut's."""

    @staticmethod
    def forward(ctx, spec: ConvXSpec, x, *tensors):
        k = len(spec.kernels)
        xp = to_phys(x)
        ys, affs = [], []
        cur, cur_aff = xp, None
        for i in range(k):
            w, g, b, rm, rv = tensors[5 * i: 5 * i + 5]
            y, aff = ops.conv_fwd_bn(cur, w, spec.kernels[i], spec.strides[i], spec.pads[i],
                                     None if cur_aff is None else cur_aff[0], None if cur_aff is None else cur_aff[1],
                                     cur_aff is not None, g, b, rm, rv, spec.momentum, spec.eps, spec.training)
            ys.append(y)
            affs.append(aff)
            cur, cur_aff = y, aff
        yd, affd = None, None
        if spec.residual and spec.has_ds:
            wd, gd, bd, rmd, rvd = tensors[5 * k: 5 * k + 5]
            yd, affd = ops.conv_fwd_bn(xp, wd, (1, 1, 1), spec.ds_stride, (0, 0, 0), None, None, False, gd, bd, rmd, rvd,
                                       spec.momentum, spec.eps, spec.training)
        a, b = affs[-1][0], affs[-1][1]
        zp = None
        if spec.tail == 'mean':
            assert not spec.residual and spec.pool is None
            if spec.out is not None:
                ops.proj_tail_fwd(ys[-1], a, b, out=spec.out.buf, coff=spec.out.coff)
                z = None
            else:
                z = ops.proj_tail_fwd(ys[-1], a, b)
        else:
            if spec.residual:
                z = ops.block_end_fwd(ys[-1], a, b, yd if yd is not None else xp, None if affd is None else affd[0],
                                      None if affd is None else affd[1])
            else:
                z = ops.block_end_fwd(ys[-1], a, b)
            if spec.pool is not None:
                zp = ops.maxpool_fwd(z, spec.pool)
        ctx.spec = spec
        ctx.x_shape = tuple(xp.shape)
        ctx.nt = len(tensors)
        flat = [xp, z if z is not None else xp] + ys + [t for aff in affs for t in aff]     # (the mean tail's backward does not read z)
        if yd is not None:
            flat += [yd] + list(affd)
        ctx.save_for_backward(*flat, *tensors)
        ctx.nflat = len(flat)
        if z is None:
            return spec.out.view()
        zl = to_logical(z, spec.ndim)
        if spec.pool is not None:
            return zl, to_logical(zp, spec.ndim)
        return zl

    @staticmethod
    def backward(ctx, dz, dzp=None):
        spec: ConvXSpec = ctx.spec
        k = len(spec.kernels)
        saved = ctx.saved_tensors
        flat, tensors = saved[:ctx.nflat], saved[ctx.nflat:]
        xp, z = flat[0], flat[1]
        ys = list(flat[2:2 + k])
        affs = [flat[2 + k + 4 * i: 2 + k + 4 * i + 4] for i in range(k)]
        yd, affd = None, None
        if spec.residual and spec.has_ds:
            yd = flat[2 + 5 * k]
            affd = flat[3 + 5 * k: 7 + 5 * k]
        grads = [None] * ctx.nt
        dzp_p = to_phys(dzp) if dzp is not None else None
        dz_stride = None
        if spec.tail == 'mean':
            dz_p, dz_stride = phys_slice(dz)           # the gradient of a concat buffer's slice is read in place
        else:
            dz_p = to_phys(dz) if dz is not None else None
        y_last = ys[-1]
        # eval mode: BatchNorm normalised with the running statistics, which do not depend on the batch -> its backward is
        # dy = g * gamma * invstd without the batch-statistics terms.  The coefficients cP, cQ are proportional to 1/count, so
        # an infinite count gives exactly torch's eval-mode native_batch_norm_backward (dgamma, dbeta do not involve count).
        inf = float('inf')
        count = y_last.numel() // y_last.shape[-1] if spec.training else inf
        g_last = tensors[5 * (k - 1) + 1]
        if spec.tail == 'mean':
            dA = ops.proj_tail_bwd(dz_p, y_last.shape, ostride=dz_stride)
            partial, rows = ops.bn_bwd_reduce(dA, y_last, affs[-1][0], affs[-1][1], True)
            dg, db, cA, cP, cQ = ops.bn_bwd_finalize(partial, rows, 2, 1, count, g_last, affs[-1][2], affs[-1][3],
                                                     _sink(g_last), _sink(tensors[5 * (k - 1) + 2]))
            dy = ops.bn_bwd_apply(dA, y_last, affs[-1][0], affs[-1][1], True, cA, cP, cQ, out=dA)
            G = None
        else:
            G, partial, rows, ncols = ops.block_end_bwd(dz_p, dzp_p, z, y_last, yd, spec.pool)
            dg, db, cA, cP, cQ = ops.bn_bwd_finalize(partial, rows, ncols, 1, count, g_last, affs[-1][2], affs[-1][3],
                                                     _sink(g_last), _sink(tensors[5 * (k - 1) + 2]))
            if not (spec.residual and yd is not None):
                dy = ops.bn_bwd_apply(G, y_last, affs[-1][0], affs[-1][1], False, cA, cP, cQ)
        grads[5 * (k - 1) + 1], grads[5 * (k - 1) + 2] = dg, db
        # shortcut branch
        dx_short = None
        if spec.residual:
            if yd is not None:
                wd, gd = tensors[5 * k], tensors[5 * k + 1]
                dgd, dbd, cAd, cPd, cQd = ops.bn_bwd_finalize(partial, rows, ncols, 2, yd.numel() // yd.shape[-1] if spec.training else inf, gd,
                                                              affd[2], affd[3], _sink(gd), _sink(tensors[5 * k + 2]))
                # the block's last BN and the shortcut's BN take the same G: one pass writes both conv gradients
                dy, dyd = ops.bn_bwd_apply2(G, y_last, yd, (cA, cP, cQ), (cAd, cPd, cQd))
                grads[5 * k] = _wgrad(xp, dyd, wd.shape, (1, 1, 1), spec.ds_stride, (0, 0, 0), None, None, False, _sink(wd))
                grads[5 * k + 1], grads[5 * k + 2] = dgd, dbd
                if spec.need_dx:
                    dx_short = ops.conv_dgrad(dyd, wd, xp.shape, (1, 1, 1), spec.ds_stride, (0, 0, 0))
            else:
                dx_short = G
        # main branch, last conv to first
        dx = None
        for i in range(k - 1, -1, -1):
            w = tensors[5 * i]
            if i > 0:
                inp, a_in, b_in = ys[i - 1], affs[i - 1][0], affs[i - 1][1]
            else:
                inp, a_in, b_in = xp, None, None
            grads[5 * i] = _wgrad(inp, dy, w.shape, spec.kernels[i], spec.strides[i], spec.pads[i], a_in, b_in, i > 0,
                                  _sink(w))
            if i > 0:
                # dgrad with the ReLU mask and the BatchNorm-backward sums of the previous conv's BN taken in its epilogue
                dA, partial, rows = ops.conv_dgrad_bnr(dy, w, inp, a_in, b_in, spec.kernels[i], spec.strides[i], spec.pads[i])
                cnt = inp.numel() // inp.shape[-1] if spec.training else inf
                dg, db, cA, cP, cQ = ops.bn_bwd_finalize(partial, rows, 2, 1, cnt, tensors[5 * (i - 1) + 1],
                                                         affs[i - 1][2], affs[i - 1][3], _sink(tensors[5 * (i - 1) + 1]),
                                                         _sink(tensors[5 * (i - 1) + 2]))
                grads[5 * (i - 1) + 1], grads[5 * (i - 1) + 2] = dg, db
                dy = ops.bn_bwd_apply(dA, inp, a_in, b_in, True, cA, cP, cQ, out=dA)
            elif spec.need_dx:
                dx = ops.conv_dgrad(dy, w, inp.shape, spec.kernels[i], spec.strides[i], spec.pads[i], addend=dx_short)
                dx_short = None
        if spec.need_dx and dx is None:
            dx = dx_short
        dxl = to_logical(dx, spec.ndim) if (spec.need_dx and dx is not None) else None
        return (None, dxl) + tuple(grads)


class MaxPoolFunction(torch.autograd.Function):
    """Stand-alone nn.MaxPool3d / nn.MaxPool2d (kernel = stride, floor); fusion3D2D.py:87-90,168-171."""

    @staticmethod
    def forward(ctx, x, kernel, ndim):
        xp = to_phys(x)
        zp = ops.maxpool_fwd(xp, kernel)
        ctx.save_for_backward(xp)
        ctx.kernel, ctx.ndim = kernel, ndim
        return to_logical(zp, ndim)

    @staticmethod
    def backward(ctx, dzp):
        (xp,) = ctx.saved_tensors
        return to_logical(ops.maxpool_bwd(xp, to_phys(dzp), ctx.kernel), ctx.ndim), None, None


class MeanDepthFunction(torch.autograd.Function):
    """torch.mean(x, dim=4, keepdim=True) on a channels-last feature (fusion3D2D.py:528-536), for callers that
    use the projection modules stand-alone; the fused path is ConvXSpec.tail == 'mean'."""

    @staticmethod
    def forward(ctx, x):
        xp = to_phys(x)
        C = xp.shape[-1]
        one = torch.ones(C, dtype=torch.float32, device=xp.device)
        zero = torch.zeros(C, dtype=torch.float32, device=xp.device)
        ctx.shape = tuple(xp.shape)
        # inputs here are post-ReLU (>= 0): relu(1*x+0) == x
        ctx.nonneg = True
        return to_logical(ops.proj_tail_fwd(xp, one, zero), 5)

    @staticmethod
    def backward(ctx, dout):
        return to_logical(ops.proj_tail_bwd(to_phys(dout), ctx.shape), 5)


class Resize2DFunction(torch.autograd.Function):
    """conv_2d[:,:,:,:,None] then None / F.interpolate(trilinear) / F.adaptive_max_pool3d to the en-face grid
    (fusion3D2D.py:544-564).  Input (B,C,S',W') logical; output (B,C,S,W,1)."""

    @staticmethod
    def forward(ctx, x, size, mode, slot=None):
        xp = to_phys(x)
        ctx.mode, ctx.x_shape, ctx.in_ndim = mode, tuple(xp.shape), x.dim()
        if slot is not None:
            _, idx = ops.resize2d_fwd(xp, int(size[0]), int(size[1]), mode, out=slot.buf, coff=slot.coff)
            ctx.save_for_backward(idx if idx is not None else torch.empty(0, device=xp.device))
            return slot.view()
        out, idx = ops.resize2d_fwd(xp, int(size[0]), int(size[1]), mode)
        ctx.save_for_backward(idx if idx is not None else torch.empty(0, device=xp.device))
        return to_logical(out, 5)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        dp, stride = phys_slice(dout)
        dx = ops.resize2d_bwd(dp, ctx.x_shape, ctx.mode, idx if idx.numel() else None, ostride=stride)
        return to_logical(dx, ctx.in_ndim), None, None, None


def _slots_in_place(tensors, buf, first_off=0):
    """True when every tensor is the tagged view of consecutive slots of ``buf`` starting at channel ``first_off``."""
    off = first_off
    for t in tensors:
        slot = getattr(t, '_ffpn_slot', None)
        if slot is None or slot.buf is not buf or slot.coff != off or t.shape[1] != slot.C:
            return False
        off += slot.C
    return True


def common_cat_buffer(tensors):
    """The first CatSlot of the concat buffer shared by all of ``tensors`` if each is the tagged view of consecutive slots from
    channel 0, else None."""
    slot = getattr(tensors[0], '_ffpn_slot', None) if tensors else None
    if slot is None or not _slots_in_place(tensors, slot.buf):
        return None
    return slot


def _slice_grads(dc, shapes, in_place):
    """Gradient of the leading concat members: channel-slice VIEWS of dc when the producers read them in place, else copies."""
    outs, off = [], 0
    for sh in shapes:
        if in_place:
            outs.append(to_logical(dc[..., off:off + sh[-1]], 5))
        else:
            g = torch.empty(sh, dtype=dc.dtype, device=dc.device)
            ops.slice_copy(dc, off, g, 0, sh[-1])
            outs.append(to_logical(g, 5))
        off += sh[-1]
    return outs, off


class UpCatFunction(torch.autograd.Function):
    """Upsample_Custom3d_nearest(deeper) and torch.cat([skip..., up], 1) (fusion3D2D.py:956-966, components.py:72-76,
    :259-268).  ``buf``: the level's pre-allocated concat buffer whose leading slots the skips' producers have ALREADY filled
    (CatSlot) -- then only the upsample runs, into the last slot, and the backward hands the skips' gradients out as views.
    Without it the skips are copied into a fresh buffer."""

    @staticmethod
    def forward(ctx, factor, deeper, holder, *skips):
        buf = holder.buf if holder is not None else None      # a CatSlot (plain object): the buffer is not an autograd input
        dp = to_phys(deeper)
        fS, fW = int(factor[0]), int(factor[1])
        B, Si, Wi, H, Cd = dp.shape
        if H != 1 or int(factor[2]) != 1:
            raise ValueError('the decoder upsamples en-face maps only (depth 1, factor (fS, fW, 1))')
        So, Wo = Si * fS, Wi * fW
        in_place = buf is not None
        if in_place:
            sshapes = [(B, So, Wo, 1, int(s.shape[1])) for s in skips]
            for s in skips:
                if tuple(s.shape) != (B, s.shape[1], So, Wo, 1):
                    raise RuntimeError(f'Sizes of tensors must match except in dimension 1: skip {tuple(s.shape)} vs '
                                       f'upsampled {(B, Cd, So, Wo, 1)}')
            off = sum(sh[-1] for sh in sshapes)
            if tuple(buf.shape) != (B, So, Wo, 1, off + Cd):
                raise RuntimeError(f'concat buffer {tuple(buf.shape)} does not fit {(B, So, Wo, 1, off + Cd)}')
            cat = buf
        else:
            sp = [to_phys(s) for s in skips]
            for s in sp:
                if tuple(s.shape[:4]) != (B, So, Wo, 1):
                    raise RuntimeError(f'Sizes of tensors must match except in dimension 1: skip {tuple(s.shape)} vs '
                                       f'upsampled {(B, So, Wo, 1, Cd)}')
            sshapes = [tuple(s.shape) for s in sp]
            cat = torch.empty((B, So, Wo, 1, sum(s.shape[-1] for s in sp) + Cd), dtype=dp.dtype, device=dp.device)
            off = 0
            for s in sp:
                ops.slice_copy(s, 0, cat, off, s.shape[-1])
                off += s.shape[-1]
        ops.upsample_fwd(dp, fS, fW, out=cat, coff=off)
        ctx.meta = (fS, fW, tuple(dp.shape), sshapes, in_place)
        return to_logical(cat, 5)

    @staticmethod
    def backward(ctx, dcat):
        fS, fW, dshape, sshapes, in_place = ctx.meta
        dc = to_phys(dcat)
        outs, off = _slice_grads(dc, sshapes, in_place)
        dd = ops.upsample_bwd(dc, dshape, fS, fW, coff=off)
        return (None, to_logical(dd, 5), None) + tuple(outs)


class CatFunction(torch.autograd.Function):
    """torch.cat(tensors, 1) for channels-last en-face maps (fusion3D2D.py:572).  ``buf``: see UpCatFunction (all members
    already in place -> no kernel at all)."""

    @staticmethod
    def forward(ctx, holder, *xs):
        buf = holder.buf if holder is not None else None
        if buf is not None:
            ctx.shapes = [tuple(buf.shape[:4]) + (int(x.shape[1]),) for x in xs]
            ctx.in_place = True
            return to_logical(buf, 5)
        ps = [to_phys(x) for x in xs]
        Ctot = sum(p.shape[-1] for p in ps)
        cat = torch.empty(tuple(ps[0].shape[:4]) + (Ctot,), dtype=ps[0].dtype, device=ps[0].device)
        off = 0
        for p in ps:
            if tuple(p.shape[:4]) != tuple(ps[0].shape[:4]):
                raise RuntimeError('Sizes of tensors must match except in dimension 1')
            ops.slice_copy(p, 0, cat, off, p.shape[-1])
            off += p.shape[-1]
        ctx.shapes = [tuple(p.shape) for p in ps]
        ctx.in_place = False
        return to_logical(cat, 5)

    @staticmethod
    def backward(ctx, dcat):
        outs, _ = _slice_grads(to_phys(dcat), ctx.shapes, ctx.in_place)
        return (None,) + tuple(outs)


def cat(*xs):
    """torch.cat(xs, 1); free when the members are the consecutive slots of one concat buffer."""
    return CatFunction.apply(common_cat_buffer(xs), *xs)


def upcat(factor, deeper, *skips):
    """cat([*skips, upsample(deeper)], 1); the skips are not copied when they are the leading slots of one concat buffer."""
    return UpCatFunction.apply(factor, deeper, common_cat_buffer(skips) if skips else None, *skips)


_FUSE_HEAD_ACT = True


def fuse_head_activation(on: bool) -> None:
    """The wrappers' sigmoid (fusion_nets.py:110,118) runs inside the head kernel by default; switch it off to get logits out of
    ``final1`` (forward hooks on that module, stage-level parity tests)."""
    global _FUSE_HEAD_ACT
    _FUSE_HEAD_ACT = bool(on)


def head_activation_fused() -> bool:
    return _FUSE_HEAD_ACT


class HeadFunction(torch.autograd.Function):
    """final1 = nn.Conv3d(C, n_classes, 1) with bias (fusion3D2D.py:223,579) -> fp32 (B,n,S,W,1): logits, or with
    ``act='sigmoid'`` the prediction itself (fusion_nets.py:110,118 fused into the same kernel, forward and backward)."""

    @staticmethod
    def forward(ctx, x, w, bias, act=None):
        xp = to_phys(x)
        out = ops.head_fwd(xp, w, bias, 1 if act == 'sigmoid' else 0)
        ctx.save_for_backward(xp, w, *([out] if act == 'sigmoid' else []))
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        xp, w, *rest = ctx.saved_tensors
        pred = rest[0] if rest else None
        dx, dw, db = ops.head_bwd(xp, w, dout.contiguous().float(), pred, need_dx=ctx.needs_input_grad[0])
        return (to_logical(dx, 5) if dx is not None else None), dw, (db if ctx.has_bias else None), None


class MixDiceBCEFunction(torch.autograd.Function):
    """Mix({Dice_loss_jointv2, BCE_Lossv2}) with unit coefficients (common/loss.py:9-90) as two kernels forward and one
    backward -> (total, dice, bce); only ``total`` carries a gradient (to the prediction)."""

    @staticmethod
    def forward(ctx, pred, mask):
        pred_c, mask_c = pred.contiguous(), mask.contiguous().float()
        stats = ops.mix_loss_fwd(pred_c, mask_c)
        ctx.save_for_backward(pred_c, mask_c, stats)
        total, dice, bce = stats[0], stats[1], stats[2]
        ctx.mark_non_differentiable(dice, bce)
        return total, dice, bce

    @staticmethod
    def backward(ctx, g_total, _g_dice, _g_bce):
        pred_c, mask_c, stats = ctx.saved_tensors
        return ops.mix_loss_bwd(pred_c, mask_c, stats, g_total.contiguous().float()), None


def pack_oct(oct: torch.Tensor) -> torch.Tensor:
    """(B,1,S,W,H) logical OCT volume as handed over by FPNHybridFusion.forward (a permuted *view* of the
    dataloader's (B,1,S,H,W) batch, fusion_nets.py:114) -> channels-last compute-dtype tensor, same logical
    shape.  The H<->W transpose is done by one kernel instead of a strided read in the first conv."""
    if oct.dim() != 5 or oct.shape[1] != 1:
        raise ValueError(f'expected a (B,1,S,W,H) volume, got {tuple(oct.shape)}')
    if oct.requires_grad:
        raise NotImplementedError('gradients with respect to the input volume are not computed by the CUDA path')
    src = oct.permute(0, 1, 2, 4, 3)                       # back to (B,1,S,H,W)
    if src.is_contiguous() and src.dtype == torch.float32:
        p = ops.pack_volume(src, _COMPUTE_DTYPE)           # memory order (B,1,S,W,H); C == 1
        B, _, S, W, H = oct.shape
        return p.view(B, S, W, H, 1).permute(0, 4, 1, 2, 3)
    return to_logical(to_phys(oct.float()), 5)


def pack_image2d(img: torch.Tensor) -> torch.Tensor:
    """(B,1,S',W') fp32 -> channels-last compute dtype (same logical shape)."""
    if img.requires_grad:
        raise NotImplementedError('gradients with respect to the 2-D input image are not computed by the CUDA path')
    return to_logical(to_phys(img.float().contiguous() if img.dtype != torch.float32 else img), 4)
