"""Stand-in for pytorch_lightning 1.x: just enough for the reference modules to import (test infra)."""
import torch
import torch.nn as nn


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.global_step = 0

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        return torch.device("cpu")


class LightningDataModule:
    def __init__(self, *a, **k):
        pass
