"""``model.apply(weight_init)`` with the reference's per-type rules (``common/weight_init.py:17-82``):
Conv2d/3d + transposed 2d/3d + Linear weights xavier-normal, Conv1d/ConvTranspose1d normal; biases normal
except Conv3d bias = 0; BatchNorm weight ~ N(1, 0.02), bias 0; recurrent cells orthogonal matrices / normal
vectors.  The order of the random draws is what makes seeded initialisation reproduce the reference's."""
import torch.nn as nn
import torch.nn.init as init

_XAVIER = (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d, nn.Linear)
_NORMAL = (nn.Conv1d, nn.ConvTranspose1d)
_NORMS = (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)
_RECURRENT = (nn.LSTM, nn.LSTMCell, nn.GRU, nn.GRUCell)


def weight_init(m):
    if isinstance(m, _NORMAL) or isinstance(m, _XAVIER):
        (init.normal_ if isinstance(m, _NORMAL) else init.xavier_normal_)(m.weight.data)
        if m.bias is not None:
            if isinstance(m, nn.Conv3d):
                init.zeros_(m.bias.data)
            else:
                init.normal_(m.bias.data)
    elif isinstance(m, _NORMS):
        init.normal_(m.weight.data, mean=1, std=0.02)
        init.constant_(m.bias.data, 0)
    elif isinstance(m, _RECURRENT):
        for param in m.parameters():
            if len(param.shape) >= 2:
                init.orthogonal_(param.data)
            else:
                init.normal_(param.data)
