"""SharedMLP / SE3d — the two small dense layers PVConv is built from (reference: PVCNN/modules/shared_mlp.py:6-36,
se.py:6-17).  Dense layers stay on stock PyTorch / cuDNN (out of the hot-path scope); they are restated here only
so that PVConv can be constructed with identical state_dict keys (`layers.*`, `fc.*`)."""
import torch.nn as nn

__all__ = ['SharedMLP', 'SE3d']


class SharedMLP(nn.Module):
    def __init__(self, in_channels, out_channels, dim=1):
        super().__init__()
        if dim not in (1, 2):
            raise ValueError
        conv, bn = (nn.Conv1d, nn.BatchNorm1d) if dim == 1 else (nn.Conv2d, nn.BatchNorm2d)
        widths = out_channels if isinstance(out_channels, (list, tuple)) else [out_channels]
        layers = []
        for oc in widths:
            if oc < 1:                       # a fractional entry is a dropout probability
                layers.append(nn.Dropout(oc))
                continue
            layers += [conv(in_channels, oc, 1), bn(oc), nn.ReLU(True)]
            in_channels = oc
        self.layers = nn.Sequential(*layers)

    def forward(self, inputs):
        if isinstance(inputs, (list, tuple)):
            return (self.layers(inputs[0]), *inputs[1:])
        return self.layers(inputs)


class SE3d(nn.Module):
    def __init__(self, channel, reduction=8):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())

    def forward(self, inputs):
        b, c = inputs.shape[:2]
        gate = self.fc(inputs.mean(-1).mean(-1).mean(-1))
        return inputs * gate.view(b, c, 1, 1, 1)
