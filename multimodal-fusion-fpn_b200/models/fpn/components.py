"""Shared building blocks of the FPN bodies, API-compatible with the reference's
``models/fpn/components.py`` (class names, constructor signatures, attribute names, ``state_dict`` keys),
with every forward running on the sm_100a kernels of libfusionfpn.so through ``ffpn.functional``.

The residual block exists once here (``ConvXBase``); the reference keeps separate 3-D / 2-D copies in
``components.py:80-227``, ``fusion3D2D.py:585-732`` and ``:735-893`` that differ only in the Conv/BN
dimensionality and in how the ``is_batchnorm=False`` variant wraps a lone conv.
"""
import numpy as np
import torch
from torch import nn

from ffpn import functional as FF


class SegmentationNetwork(nn.Module):
    """Base class holding the four architecture switches (reference components.py:7-19)."""

    def __init__(self, n_classes=1, is_batchnorm=True, in_channels=1, is_deconv=False):
        super().__init__()
        self.n_classes = n_classes        # number of output classes
        self.is_deconv = is_deconv        # transposed convs instead of nearest upsampling (unsupported, see up block)
        self.in_channels = in_channels
        self.is_batchnorm = is_batchnorm


class ConvXBase(nn.Module):
    """k x [Conv -> BN -> ReLU] with the last pair lacking the ReLU, a residual branch (identity or
    ``downsample`` = 1x1 conv + BN), an in-place add and a final ReLU (reference forward
    fusion3D2D.py:717-732).  Sub-module tree and parameter names follow the reference exactly:
    ``convBlock.<i>.0`` conv, ``convBlock.<i>.1`` BN, ``downsample.0/.1``."""

    conv_cls = nn.Conv3d
    bn_cls = nn.BatchNorm3d
    wrap_lone_conv = False   # components.py:164,185 wrap the bias conv in nn.Sequential; fusion3D2D.py:669,690 do not
    ndim = 5

    def __init__(self, in_size, out_size, kernel_size, stride, padding, is_batchnorm, is_residual, dropout, downsample):
        super().__init__()
        k = len(kernel_size)
        layers = []
        for i in range(k):
            conv = self.conv_cls(in_channels=in_size if i == 0 else out_size, out_channels=out_size,
                                 kernel_size=kernel_size[i], stride=stride[i], padding=padding[i],
                                 bias=not is_batchnorm)
            last = i == k - 1
            if is_batchnorm:
                mods = [conv, self.bn_cls(out_size)] + ([] if last else [nn.ReLU()])
                layers.append(nn.Sequential(*mods))
            elif not last:
                layers.append(nn.Sequential(conv, nn.ReLU()))
            else:
                layers.append(nn.Sequential(conv) if self.wrap_lone_conv else conv)
        if k == 0:
            raise AssertionError('UserException: in module "%s". Error when value of iterator is "0' % type(self).__name__)
        self.convBlock = nn.Sequential(*layers)
        self.is_residual = is_residual
        self.downsample = downsample
        self.relu = nn.ReLU(inplace=True)
        self.drop = nn.Dropout(dropout) if dropout > 0.0 else None
        self._is_batchnorm = is_batchnorm
        self._nbt_managed = False          # True when the owning body bumps num_batches_tracked in one launch

    # -- kernel path ------------------------------------------------------------------------------------
    def _bn_modules(self):
        bns = [blk[1] for blk in self.convBlock]
        if self.is_residual and self.downsample is not None:
            bns.append(self.downsample[1])
        return bns

    def forward(self, x, pool=None, tail='relu', need_dx=True, out=None):
        """``pool``: a MaxPool module to fuse behind the block (returns ``(out, pooled)``);
        ``tail='mean'``: fuse the projection's BN+ReLU+mean(dim=4) (returns the (B,C,S,W,1) mean), written straight into
        the concat slot ``out`` (ffpn.functional.CatSlot) when one is given."""
        if not self._is_batchnorm:
            raise NotImplementedError('is-batchnorm=False has no CUDA path (the reference .ini pins is-batchnorm=True)')
        if self.drop is not None:
            raise NotImplementedError('dropout > 0 has no CUDA path (all nine rates are 0.0 in the reference .ini)')
        convs = [blk[0] for blk in self.convBlock]
        bns = [blk[1] for blk in self.convBlock]
        has_ds = self.is_residual and self.downsample is not None
        tensors = []
        for c, b in zip(convs, bns):
            tensors += [c.weight, b.weight, b.bias, b.running_mean, b.running_var]
        ds_stride = (1, 1, 1)
        if has_ds:
            dc, db = self.downsample[0], self.downsample[1]
            tensors += [dc.weight, db.weight, db.bias, db.running_mean, db.running_var]
            ds_stride = FF.k3(dc.stride)
        bn0 = bns[0]
        spec = FF.ConvXSpec(
            kernels=tuple(FF.k3(c.kernel_size) for c in convs), strides=tuple(FF.k3(c.stride) for c in convs),
            pads=tuple(FF.p3(c.padding) for c in convs), residual=bool(self.is_residual), has_ds=has_ds,
            ds_stride=ds_stride, pool=None if pool is None else FF.k3(pool.kernel_size), tail=tail,
            training=self.training, momentum=0.1 if bn0.momentum is None else float(bn0.momentum), eps=float(bn0.eps),
            need_dx=bool(need_dx and x.requires_grad), ndim=x.dim(), out=out if tail == 'mean' else None)
        if self.training and not self._nbt_managed:
            torch._foreach_add_([b.num_batches_tracked for b in self._bn_modules()], 1)
        res = FF.ConvXFunction.apply(spec, x, *tensors)
        return FF.tag_slot(res, spec.out) if spec.out is not None else res


class unet3dConvX(ConvXBase):
    '''Convolution-Block with X convolutions in 3D (reference components.py:80-227).'''
    wrap_lone_conv = True


class unet3dUp2modified(nn.Module):
    """Two-input up block: nearest-upsample the deeper level, concatenate with the skip, ConvX
    [(3,3,1),(3,3,1)] with a 1x1x1 conv+BN shortcut (reference components.py:23-76)."""

    n_skips = 1
    convx_cls = unet3dConvX

    def __init__(self, lowlayer_channels, currlayer_channels, upfactor, is_deconv, is_residual, dropout, is_batchnorm):
        super().__init__()
        if is_deconv:
            self.up = nn.ConvTranspose3d(lowlayer_channels, currlayer_channels, kernel_size=upfactor, stride=upfactor)
        else:
            self.up = Upsample_Custom3d_nearest(scale_factor=upfactor, mode='nearest')
        cin = lowlayer_channels + currlayer_channels * self.n_skips
        shortcut = nn.Conv3d(cin, currlayer_channels, kernel_size=1, stride=1, bias=not is_batchnorm)
        if is_batchnorm:
            downsample = nn.Sequential(shortcut, nn.BatchNorm3d(currlayer_channels))
        else:
            downsample = self._wrap_bias_shortcut(shortcut)
        self.conv = self.convx_cls(in_size=cin, out_size=currlayer_channels, kernel_size=[(3, 3, 1), (3, 3, 1)],
                                   stride=[(1, 1, 1), (1, 1, 1)], padding=[(1, 1, 0), (1, 1, 0)],
                                   is_batchnorm=is_batchnorm, is_residual=is_residual, dropout=dropout,
                                   downsample=downsample)

    @staticmethod
    def _wrap_bias_shortcut(conv):
        return nn.Sequential(conv)                  # components.py:51-57

    def _cat(self, skips, deeper):
        if not isinstance(self.up, Upsample_Custom3d_nearest):
            raise NotImplementedError('is-deconv=True has no CUDA path (shape-inconsistent in the reference too)')
        return FF.upcat(self.up.int_factor(), deeper, *skips)

    def forward(self, inputs1, inputs2):
        return self.conv(self._cat([inputs1], inputs2))


class Upsample_Custom3d_nearest(nn.Module):
    """Nearest-neighbour upsampling by ``scale_factor`` over the last three axes (reference
    components.py:230-276): per axis ``idx[i] = ceil((i+1)/scale) - 1`` for ``i < int(n*scale)``."""

    def __init__(self, scale_factor=None, mode='nearest'):
        super().__init__()
        self.size = None
        self.scale_factor = scale_factor
        self.mode = mode
        assert self.scale_factor is not None, "scale_factor must be set"

    @staticmethod
    def index_table(n_in, scale):
        return (np.ceil(np.arange(1, 1 + int(n_in * scale)) / scale) - 1).astype(int)

    def int_factor(self):
        f = tuple(self.scale_factor)
        if any(int(v) != v or v < 1 for v in f):
            raise NotImplementedError(f'only integer up-factors have a CUDA path, got {f}')
        return tuple(int(v) for v in f)

    def forward(self, input: torch.Tensor):
        f = self.int_factor()
        if f[2] != 1 or input.shape[-1] != 1:
            raise NotImplementedError('the CUDA path upsamples en-face maps (depth 1) only')
        # idx = ceil((i+1)/f) - 1 == i // f for integer f: identical to the reference's gather
        return FF.upcat(f, input)

    def __repr__(self):
        info = 'scale_factor=' + str(self.scale_factor) if self.scale_factor is not None else 'size=' + str(self.size)
        return self.__class__.__name__ + '(' + info + ', mode=' + self.mode + ')'


class Upsample_Custom2d_nearest(nn.Module):
    """2-D twin (reference components.py:281-323; unused by every registered model)."""

    def __init__(self, scale_factor=None, mode='nearest'):
        super().__init__()
        self.size = None
        self.scale_factor = scale_factor
        self.mode = mode

    def forward(self, input):
        f = tuple(int(v) for v in self.scale_factor)
        out = FF.upcat((f[0], f[1], 1), input[:, :, :, :, None])
        return out[:, :, :, :, 0]

    def __repr__(self):
        info = 'scale_factor=' + str(self.scale_factor) if self.scale_factor is not None else 'size=' + str(self.size)
        return self.__class__.__name__ + '(' + info + ', mode=' + self.mode + ')'


class MaxPool3d(nn.MaxPool3d):
    """nn.MaxPool3d (kernel = stride, no padding) on the CUDA kernels; isinstance-compatible."""

    def forward(self, input):
        k = FF.k3(self.kernel_size)
        if FF.k3(self.stride) != k or any(FF.p3(self.padding)) or self.ceil_mode or FF.k3(self.dilation) != (1, 1, 1):
            raise NotImplementedError('only kernel == stride, unpadded, floor-mode pooling has a CUDA path')
        return FF.MaxPoolFunction.apply(input, k, 5)


class MaxPool2d(nn.MaxPool2d):
    def forward(self, input):
        k = FF.k3(self.kernel_size)
        st = self.stride if isinstance(self.stride, (tuple, list)) else (self.stride, self.stride)
        if FF.k3(st) != k or self.ceil_mode:
            raise NotImplementedError('only kernel == stride, unpadded, floor-mode pooling has a CUDA path')
        return FF.MaxPoolFunction.apply(input, k, 4)


class HeadConv3d(nn.Conv3d):
    """final1: Conv3d(C -> n_classes, 1x1x1, bias) producing fp32 logits (reference fusion3D2D.py:223).  A task wrapper
    whose ``last_activation`` is the stock sigmoid sets ``fused_activation = 'sigmoid'`` for the duration of its forward: the
    same kernel then emits the prediction (``fusion_nets.py:110,118``), and the backward folds p(1-p) in."""

    fused_activation = None

    def forward(self, input):
        if FF.k3(self.kernel_size) != (1, 1, 1):
            raise NotImplementedError('the head kernel is 1x1x1')
        return FF.HeadFunction.apply(input, self.weight, self.bias, self.fused_activation)
