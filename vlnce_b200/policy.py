"""``ILPolicy`` — the reference's imitation-learning policy surface
(``vlnce_baselines/models/policy.py:10-58``): ``act()`` and ``build_distribution()`` with the
same arguments, return values and ``state_dict`` keys (``net.*``,
``action_distribution.linear.*``), computing on the CUDA library only."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import lib
from .compat import CustomFixedCategorical, Policy
from .lib import C, ptr


class CategoricalNet(nn.Module):
    """habitat_baselines v0.1.7 ``utils/common.py`` CategoricalNet: ``linear`` attribute,
    orthogonal(gain 0.01) weight, zero bias (``policy.py:19-21``)."""

    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.linear = nn.Linear(num_inputs, num_outputs)
        nn.init.orthogonal_(self.linear.weight, gain=0.01)
        nn.init.constant_(self.linear.bias, 0)


class ILPolicy(Policy):
    def __init__(self, net, dim_actions):
        nn.Module.__init__(self)
        self.net = net
        self.dim_actions = dim_actions
        self.action_distribution = CategoricalNet(self.net.output_size, self.dim_actions)

    def forward(self, *x):
        raise NotImplementedError

    # ---- reference surface ---------------------------------------------------------------
    @lib.guarded
    def act(self, observations, rnn_states, prev_actions, masks, deterministic=False):
        """``policy.py:26-44``: features -> Categorical -> mode() / sample()."""
        with torch.no_grad():
            logits, rnn_states_out = self.net.logits_and_states(
                self.action_distribution.linear, observations, rnn_states, prev_actions, masks)
            R, A = logits.shape
            if deterministic:  # CustomFixedCategorical.mode(): probs.argmax(-1, keepdim=True)
                action = torch.empty(R, 1, dtype=torch.int64, device=logits.device)
                C.vlnce_argmax_rows(ptr(logits), ptr(action), R, A, lib.stream())
            else:
                action = CustomFixedCategorical(logits=logits).sample()
        return action, rnn_states_out

    def get_value(self, *args, **kwargs):
        raise NotImplementedError

    def evaluate_actions(self, *args, **kwargs):
        raise NotImplementedError

    @lib.guarded
    def build_distribution(self, observations, rnn_states, prev_actions, masks):
        """``policy.py:52-58``; ``.logits`` are normalised log-probabilities and carry the
        hand-written backward when grad is enabled."""
        logits, _ = self.net.logits_and_states(self.action_distribution.linear, observations,
                                               rnn_states, prev_actions, masks)
        return CustomFixedCategorical(logits=logits)
