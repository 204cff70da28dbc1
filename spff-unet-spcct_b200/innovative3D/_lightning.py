"""Minimal stand-in for the parts of pytorch_lightning the hot path touches, used only when the real
package is not installed (it is a third-party dependency of the reference, `requirements.txt:70`,
and absent from this image). With Lightning present, `innovative3D.models` subclasses the real
`pl.LightningModule` and this file is never imported."""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn


class _HParams(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class LightningModule(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()
        self._hparams = _HParams()
        self.logged = {}

    @property
    def hparams(self):
        return self._hparams

    def save_hyperparameters(self, *args, **kwargs):
        for a in args:
            if isinstance(a, dict):
                self._hparams.update(a)
        self._hparams.update(kwargs)

    def log(self, name, value, *a, **kw):
        self.logged[name] = value

    def log_dict(self, d, *a, **kw):
        self.logged.update(d)

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


class LightningDataModule:
    pass


def seed_everything(seed: int, workers: bool = False):
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


pl = SimpleNamespace(LightningModule=LightningModule, LightningDataModule=LightningDataModule,
                     seed_everything=seed_everything)
