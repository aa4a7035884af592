"""Minimal stand-in for lightning.pytorch (see ../__init__.py)."""
import torch
from torch import nn


class LightningModule(nn.Module):
    @property
    def device(self):
        for p in self.parameters():
            return p.device
        for b in self.buffers():
            return b.device
        return torch.device("cpu")

    def save_hyperparameters(self, *args, **kwargs):
        return None

    def log(self, *args, **kwargs):
        return None


class LightningDataModule:
    pass
