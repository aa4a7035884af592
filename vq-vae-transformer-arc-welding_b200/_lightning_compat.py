"""Base class selection for the drop-in modules.

The reference derives its modules from ``lightning.pytorch.LightningModule``
(model/vector_quantizer.py:59, model/autencoder_lightning_base.py:8).  When Lightning is
installed the drop-ins derive from the real class, so Trainer/DDP/checkpoint hooks work
unchanged; in images without Lightning (this one) a small nn.Module with the handful of
members the reference code touches (`device`, `log`, `save_hyperparameters`, `hparams`)
takes its place.
"""
from __future__ import annotations

import inspect
from types import SimpleNamespace

import torch
from torch import nn

try:  # pragma: no cover - depends on the environment
    import lightning.pytorch as _pl
    LightningModule = _pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # ImportError or a broken install
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):  # type: ignore[no-redef]
        """Stand-in used only when Lightning is absent."""

        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            self.hparams = SimpleNamespace()
            self.logged = {}

        @property
        def device(self) -> torch.device:
            for p in self.parameters():
                return p.device
            for b in self.buffers():
                return b.device
            return torch.device("cpu")

        def save_hyperparameters(self, *args, **kwargs) -> None:
            """Records the constructor arguments of the calling frame (what Lightning stores
            as ``hyper_parameters`` in a checkpoint)."""
            frame = inspect.currentframe().f_back
            merged = {}
            try:
                # walk outwards through the chain of __init__ frames of this object: the leaf
                # class's arguments win, like Lightning's collect_init_args
                while frame is not None and frame.f_code.co_name == "__init__" \
                        and frame.f_locals.get("self") is self:
                    names, _, _, values = inspect.getargvalues(frame)
                    merged.update({n: values[n] for n in names if n != "self"})
                    frame = frame.f_back
                self.hparams = SimpleNamespace(**merged)
            finally:
                del frame

        def log(self, name, value, *args, **kwargs) -> None:
            self.logged[name] = value.detach() if isinstance(value, torch.Tensor) else value
