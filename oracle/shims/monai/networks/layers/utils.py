import torch
import torch.nn as nn


class Swish(nn.Module):
    def __init__(self, alpha=1.0):
        super().__init__()
        self.alpha = alpha

    def forward(self, x):
        return x * torch.sigmoid(self.alpha * x)


def _split(name):
    if isinstance(name, (tuple, list)):
        return str(name[0]).lower(), dict(name[1]) if len(name) > 1 else {}
    return str(name).lower(), {}


def get_act_layer(name):
    kind, kw = _split(name)
    if kind == "swish":
        return Swish(**kw)
    if kind == "leakyrelu":
        return nn.LeakyReLU(**kw)
    if kind == "relu":
        return nn.ReLU(**kw)
    if kind == "gelu":
        return nn.GELU(**kw)
    raise NotImplementedError(kind)


def get_norm_layer(name, spatial_dims=1, channels=1):
    kind, kw = _split(name)
    if kind == "group":
        return nn.GroupNorm(num_channels=channels, **kw)
    if kind == "batch":
        return (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)[spatial_dims - 1](channels, **kw)
    if kind == "instance":
        return (nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)[spatial_dims - 1](channels, **kw)
    raise NotImplementedError(kind)


def get_dropout_layer(name, dropout_dim=1):
    if isinstance(name, (int, float)):
        p = float(name)
    else:
        kind, kw = _split(name)
        p = kw.get("p", 0.5)
    return (nn.Dropout, nn.Dropout2d, nn.Dropout3d)[dropout_dim - 1](p=p)
