"""Lightning hooks around the VQ-VAE-Patch model -- mirror of the reference's
``model/autencoder_lightning_base.py`` (hooks :80-124), kept so that training scripts and
checkpoints written for the reference keep working with the B200 encode/quantise path.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from .._lightning_compat import LightningModule


class Autoencoder(LightningModule):
    """Base of VQVAEPatch: stores the hyper-parameters (-> checkpoint ``hyper_parameters``),
    provides the reconstruction loss, the train/val/test steps and the optimiser."""

    def __init__(self, hidden_dim: int, input_dim: int, num_embeddings: int, embedding_dim: int,
                 n_resblocks: int, learning_rate: float, seq_len: int = 200, dropout_p: float = 0.1):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.input_dim = input_dim
        self.num_embeddings: int = num_embeddings
        self.embedding_dim = embedding_dim
        self.n_resblocks = n_resblocks
        self.learning_rate = learning_rate
        self.seq_len = seq_len
        self.dropout_p = dropout_p
        self.last_recon = (0, 0)
        self.betas = (0.9, 0.95)      # kept for parity with the reference attributes (:37-38); unused there too
        self.weight_decay = 0.1
        self.save_hyperparameters()   # :41

    def forward(self, x: torch.Tensor):
        raise NotImplementedError

    def loss(self, preds: torch.Tensor, labels: torch.Tensor):
        return F.mse_loss(preds, labels)

    @staticmethod
    def weights_init(m):
        """Xavier-uniform weights and zero bias for every module whose class name contains
        'Conv' (:70-78).  The codebook is deliberately not touched."""
        if "Conv" in m.__class__.__name__:
            try:
                nn.init.xavier_uniform_(m.weight.data)
                m.bias.data.fill_(0)
            except AttributeError:
                print("Skipping initialization of ", m.__class__.__name__)

    def _forward_setp(self, x: torch.Tensor):   # (sic) name kept: the reference's hooks call it
        embedding_loss, data_recon, _perplexity = self(x)
        recon_error = F.mse_loss(data_recon, x)
        return recon_error + embedding_loss, recon_error, data_recon

    def training_step(self, batch, batch_idx):
        loss, recon_error, data_recon = self._forward_setp(batch)
        self.log("train/loss", loss, prog_bar=True)
        self.log("train/recon_error", recon_error)
        pick = torch.randint(0, len(batch), (1,))
        self.last_recon = (batch[pick], data_recon[pick])
        return {"loss": loss, "recon_error": recon_error}

    def _eval_step(self, batch, prefix: str):
        loss, recon_error, data_recon = self._forward_setp(batch)
        self.log(f"{prefix}/loss", loss, sync_dist=True, on_epoch=True, prog_bar=True)
        self.log(f"{prefix}/recon_error", recon_error, sync_dist=True, on_epoch=True)
        return {"loss": loss, "recon_error": recon_error, "data_recon": data_recon}

    def validation_step(self, batch, batch_idx):
        return self._eval_step(batch, "val")

    def test_step(self, batch, batch_idx):
        return self._eval_step(batch, "test")

    def configure_optimizers(self):
        return torch.optim.RAdam(self.parameters(), lr=self.learning_rate)   # :122-124
