"""Global auxiliary-loss registry — same surface as the reference singleton
(``vlnce_baselines/common/aux_losses.py:4-44``) so trainers that import ``AuxLosses`` keep
working: ``clear / register_loss / get_loss / reduce / is_active / activate / deactivate``.
When the reference package is importable its singleton is re-used, so both see the same
registered losses."""
from __future__ import annotations

import torch


class _AuxLosses:
    def __init__(self):
        self._losses = {}
        self._loss_alphas = {}
        self._is_active = False

    def clear(self):
        self._losses.clear()
        self._loss_alphas.clear()

    def register_loss(self, name, loss, alpha=1.0):
        assert self.is_active()
        assert name not in self._losses
        self._losses[name] = loss
        self._loss_alphas[name] = alpha

    def get_loss(self, name):
        return self._losses[name]

    def reduce(self, mask):
        assert self.is_active()
        total = 0.0
        for k, v in self._losses.items():
            total = total + self._loss_alphas[k] * torch.masked_select(v, mask).mean()
        return total

    def is_active(self):
        return self._is_active

    def activate(self):
        self._is_active = True

    def deactivate(self):
        self._is_active = False


try:  # pragma: no cover - only when the reference package is on the path
    from vlnce_baselines.common.aux_losses import AuxLosses  # type: ignore
except Exception:
    AuxLosses = _AuxLosses()
