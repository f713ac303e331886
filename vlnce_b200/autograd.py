"""Bridge between the launch plan and torch.autograd: the whole CMA net (+ optional action
head) is ONE autograd node whose backward is the hand-written CUDA backward — this is what
lets ``loss.backward()`` in the reference's ``_update_agent``
(``common/base_il_trainer.py:159-172``) work unchanged on ``distribution.logits``."""
from __future__ import annotations

import torch


def cma_features(net, head_linear, observations, rnn_states, prev_actions, masks):
    """Returns (x or log-prob logits, rnn_states_out, progress_hat|None)."""
    eng = net.engine()
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in net.parameters())
    if need_grad:
        from .backward import CMAFunction

        return CMAFunction.run(net, head_linear, observations, rnn_states, prev_actions, masks)
    x, h_out, hat = eng.forward(observations, rnn_states, prev_actions, masks, save=False)
    out = x if head_linear is None else eng.action_logits(head_linear, x, x.shape[0])
    return out.clone(), h_out.clone(), (hat.clone() if hat is not None else None)
