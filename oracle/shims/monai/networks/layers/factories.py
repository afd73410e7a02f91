import torch.nn as nn


class _Factory:
    def __init__(self, table):
        self._table = table
        for key in table:
            setattr(self, key.upper(), key.upper())

    def __getitem__(self, args):
        name, dim = args
        return self._table[str(name).lower()][int(dim) - 1]


Conv = _Factory({
    "conv": (nn.Conv1d, nn.Conv2d, nn.Conv3d),
    "convtrans": (nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d),
})
Pool = _Factory({
    "avg": (nn.AvgPool1d, nn.AvgPool2d, nn.AvgPool3d),
    "max": (nn.MaxPool1d, nn.MaxPool2d, nn.MaxPool3d),
})
