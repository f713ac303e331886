"""Process-wide registry of auxiliary losses with the surface the reference's trainers and nets call
(``vlnce_baselines/common/aux_losses.py:4-44``): ``activate / deactivate / is_active``, ``clear``,
``register_loss(name, loss, alpha)``, ``get_loss(name)``, ``reduce(mask)``.  The nets register the
progress-monitor loss here (``cma_policy.py:296-307``); ``_update_agent`` reduces it over the valid rows
(``base_il_trainer.py:167``).  When the reference package is importable its own singleton is used, so both
code bases see the same entries."""
from __future__ import annotations

from typing import Dict, NamedTuple

import torch


class _Entry(NamedTuple):
    loss: torch.Tensor
    alpha: float


class AuxLossRegistry:
    """name -> (unreduced loss tensor, weight); only usable while activated."""

    def __init__(self):
        self.entries: Dict[str, _Entry] = {}
        self.enabled = False

    # ---- switch
    def activate(self):
        self.enabled = True

    def deactivate(self):
        self.enabled = False

    def is_active(self) -> bool:
        return self.enabled

    # ---- entries
    def clear(self):
        self.entries = {}

    def register_loss(self, name: str, loss: torch.Tensor, alpha: float = 1.0):
        if not self.enabled:
            raise AssertionError("AuxLosses.register_loss() while the registry is not active")
        if name in self.entries:
            raise AssertionError(f"auxiliary loss {name!r} is already registered (clear() once per update)")
        self.entries[name] = _Entry(loss, alpha)

    def get_loss(self, name: str) -> torch.Tensor:
        return self.entries[name].loss

    def reduce(self, mask: torch.Tensor):
        """sum_k alpha_k * mean(loss_k[mask]) — the mask broadcasts against each loss."""
        if not self.enabled:
            raise AssertionError("AuxLosses.reduce() while the registry is not active")
        return sum((e.alpha * torch.masked_select(e.loss, mask).mean() for e in self.entries.values()), 0.0)


try:  # pragma: no cover - only when the reference package is on the path
    from vlnce_baselines.common.aux_losses import AuxLosses  # type: ignore
except Exception:
    AuxLosses = AuxLossRegistry()
