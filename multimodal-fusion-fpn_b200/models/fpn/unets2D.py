"""2-D-only FPN bodies, API-compatible with the reference's ``models/fpn/unets2D.py``: ``ModifiedUnet2D``
(:9-144) and ``ModifiedUnet2DLevel5`` (:147-213) -- the fusion body's 2-D encoder with a two-input decoder."""
from torch import nn

from config import config as global_config
from models.fpn import components as K
from models.fpn.components import unet3dUp2modified
from models.fpn.fusion3D2D import ModifiedUnet3D2D, UPFACTORS


class ModifiedUnet2D(ModifiedUnet3D2D):
    def __init__(self, config, output_features: bool = False):
        # like the reference (:11) this skips ModifiedUnet3D2D.__init__: no 3-D encoder is built
        super(ModifiedUnet3D2D, self).__init__(n_classes=global_config.number_of_outputs,
                                               is_batchnorm=config.getboolean('architecture', 'is-batchnorm'),
                                               in_channels=1, is_deconv=config.getboolean('architecture', 'is-deconv'))
        self.output_features = output_features
        self._read_architecture(config)
        ch = self.channels
        self._build_2d_encoder(levels=4)
        for l in (3, 2, 1):
            setattr(self, f'up_concat{l}', unet3dUp2modified(ch[l], ch[l - 1], upfactor=UPFACTORS[l],
                                                             is_deconv=self.is_deconv, is_residual=True,
                                                             dropout=self.dropout[9 - l], is_batchnorm=self.is_batchnorm))
        if not self.output_features:
            # wrapped in Sequential in the reference (:102-106) => key 'final1.0.weight'
            self.final1 = nn.Sequential(K.HeadConv3d(in_channels=ch[0], out_channels=self.n_classes, kernel_size=1))

    def _decode_2d(self, f2d):
        feats = [f[:, :, :, :, None] for f in f2d]            # 2-D -> "2-D within 3-D" (a view)
        deeper = feats[-1]
        for l in range(len(feats) - 1, 0, -1):
            deeper = getattr(self, f'up_concat{l}')(feats[l - 1], deeper)
        return deeper if self.output_features else self.final1(deeper)

    def forward(self, input_):  # type: ignore
        self._bump_bn_counters()
        return self._decode_2d(self._encode_2d(input_, 4))


class ModifiedUnet2DLevel5(ModifiedUnet2D):
    def __init__(self, config, output_features: bool = False):
        super().__init__(config, output_features)
        ch = self.channels
        self.conv5_2d = self._make_layer_2plus3_2d(ch[3], ch[4], is_batchnorm=self.is_batchnorm, is_residual=True,
                                                   dropout=self.dropout[4])
        self.pool4_2d = K.MaxPool2d(kernel_size=(2, 2))
        self.up_concat4 = unet3dUp2modified(ch[4], ch[3], upfactor=(2, 2, 1), is_deconv=self.is_deconv,
                                            is_residual=True, dropout=self.dropout[5], is_batchnorm=self.is_batchnorm)

    def forward(self, input_):
        self._bump_bn_counters()
        return self._decode_2d(self._encode_2d(input_, 5))
