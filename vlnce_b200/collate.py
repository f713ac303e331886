"""Batch assembly of the imitation-learning trainers on the device:

* ``collate_fn`` — ``dagger_trainer.py:39-114`` with the same input (list of ``(obs, prev_actions, oracle_actions,
  inflec_weights)`` trajectories) and the same five outputs in the time-major ``(T*N, ...)`` layout (observations
  padded with 1.0 — SURVEY Q7 —, actions / weights with 0, ``not_done_masks`` zero on the first row block).  It
  only moves data (pad, stack, view), so it works on device tensors as it does on host tensors: trajectories that
  already live in HBM (a device-resident feature cache) never bounce through the host.
* ``inflection_weights`` — ``dagger_trainer.py:199-211`` / ``recollection_dataset.py:251-263`` for a whole padded
  ``(T, N)`` action matrix in one kernel launch (``vlnce_inflection_weights``).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Optional

import torch

from . import lib
from .lib import C, ptr


def _pad(t: torch.Tensor, max_len: int, fill=0):
    d = max_len - t.size(0)
    if d == 0:
        return t
    return torch.cat([t, torch.full_like(t[0:1], fill).expand(d, *t.size()[1:])], dim=0)


def collate_fn(batch):
    obs_l, prev_l, act_l, w_l = (list(x) for x in zip(*batch))
    tmax = max(p.size(0) for p in prev_l)
    stacked: Dict[str, list] = defaultdict(list)
    for o in obs_l:
        for k, v in o.items():
            stacked[k].append(_pad(v, tmax, 1.0))
    obs = {}
    for k, v in stacked.items():
        s = torch.stack(v, dim=1)
        obs[k] = s.view(-1, *s.size()[2:])
    prev = torch.stack([_pad(p, tmax) for p in prev_l], dim=1)
    actions = torch.stack([_pad(a, tmax) for a in act_l], dim=1)
    weights = torch.stack([_pad(w, tmax) for w in w_l], dim=1)
    masks = torch.ones_like(actions, dtype=torch.uint8)
    masks[0] = 0
    return obs, prev.view(-1, 1), masks.view(-1, 1), actions, weights


def inflection_weights(oracle_actions: torch.Tensor, coef: float, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """oracle_actions (T,) or padded time-major (T,N) int64 on the device -> float weights of the same shape."""
    lib.require_cuda(oracle_actions, "oracle_actions")
    a = oracle_actions if oracle_actions.dtype == torch.int64 else oracle_actions.long()
    a = a.contiguous()
    T = a.shape[0]
    N = 1 if a.dim() == 1 else a.shape[1]
    w = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    lens = None if lengths is None else lengths.to(device=a.device, dtype=torch.int32).contiguous()
    C.vlnce_inflection_weights(ptr(a), ptr(lens), ptr(w), T, N, float(coef), lib.stream())
    return w
