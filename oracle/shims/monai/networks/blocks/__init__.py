import torch.nn as nn
from ..layers.factories import Conv


class UnetOutBlock(nn.Module):
    """1x1 conv head; parameter keys `conv.conv.{weight,bias}` as in MONAI's Convolution wrapper."""

    def __init__(self, spatial_dims, in_channels, out_channels, dropout=None):
        super().__init__()
        inner = nn.Sequential()
        inner.add_module("conv", Conv["conv", spatial_dims](in_channels, out_channels, kernel_size=1, stride=1, bias=True))
        self.conv = inner

    def forward(self, x):
        return self.conv(x)


class TransformerBlock(nn.Module):  # imported by the reference, never instantiated on the hot path
    def __init__(self, *a, **k):
        super().__init__()
