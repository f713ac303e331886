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


def act_graph_fingerprint(policy) -> tuple:
    """Everything a captured ``act`` forward bakes in besides the input signature: the precision mode, the BatchNorm mode
    of the trunks (it selects the trunk variant), the versions of the conv weights the packed trunk operands were built
    from (``load_state_dict`` bumps them -> the trunks repack into new buffers) and the addresses of the parameters
    the head kernels read in place (``.to()`` / ``.float()`` re-create them).  A changed fingerprint drops the graphs."""
    net = policy.net
    fp = [getattr(net, "precision", None)]
    for enc, mod in (("rgb_encoder", "cnn"), ("depth_encoder", "visual_encoder")):
        m = getattr(getattr(net, enc, None), mod, None)
        if m is None:
            continue
        fp.append(bool(m.training))
        for tr in list(getattr(m, "_trunks", {}).values()):
            fp.append(tuple(cv.conv.weight._version for cv in tr._all()))
    first = next(iter(policy.parameters()), None)
    last = None
    for last in policy.parameters():
        pass
    fp.append((first.data_ptr() if first is not None else 0, last.data_ptr() if last is not None else 0))
    return tuple(fp)


def act_graph_cache(policy) -> dict:
    """The policy's {input signature: (static inputs, graph, outputs)} cache, emptied when the fingerprint moved."""
    fp = act_graph_fingerprint(policy)
    if policy.__dict__.get("_act_graph_fp") != fp:
        policy.__dict__["_act_graphs"] = {}
        policy.__dict__["_act_graph_fp"] = fp
    return policy.__dict__["_act_graphs"]


class ILPolicy(Policy):
    def __init__(self, net, dim_actions):
        nn.Module.__init__(self)
        self.net = net
        self.dim_actions = dim_actions
        self.action_distribution = CategoricalNet(self.net.output_size, self.dim_actions)

    def forward(self, *x):
        raise NotImplementedError

    # opt-in: replay the forward of ``act`` (both trunks + head, ~150 launches) from a CUDA graph keyed by the input
    # signature (an evaluation / rollout loop calls ``act`` with fixed shapes; at most ``GRAPH_ACT_MAX`` signatures are
    # kept, e.g. for the shrinking batch of an evaluation run).  Sampling stays outside the graph.
    graph_act = False
    GRAPH_ACT_MAX = 8

    def _logits_and_states(self, observations, rnn_states, prev_actions, masks):
        from .aux_losses import AuxLosses

        head = self.action_distribution.linear
        if (not self.graph_act or not rnn_states.is_cuda or torch.cuda.is_current_stream_capturing()
                or AuxLosses.is_active()):
            return self.net.logits_and_states(head, observations, rnn_states, prev_actions, masks)
        flat = [observations[k] for k in sorted(observations)] + [rnn_states, prev_actions, masks]
        key = tuple((tuple(t.shape), t.dtype) for t in flat) + tuple(sorted(observations))
        graphs = act_graph_cache(self)
        entry = graphs.get(key)
        if entry is None:
            if len(graphs) >= self.GRAPH_ACT_MAX:
                graphs.pop(next(iter(graphs)))
            s_obs = {k: v.clone() for k, v in observations.items()}
            s_rest = [rnn_states.clone(), prev_actions.clone(), masks.clone()]
            self.net.logits_and_states(head, s_obs, *s_rest)     # eager pass: plans, tile selection, workspaces
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    out = self.net.logits_and_states(head, s_obs, *s_rest)
            torch.cuda.current_stream().wait_stream(side)
            entry = graphs[key] = ([s_obs[k] for k in sorted(s_obs)] + s_rest, graph, out)
            self.__dict__["_act_graph_fp"] = act_graph_fingerprint(self)   # (the eager pass may have built the trunks)
        static, graph, out = entry
        for dst, src in zip(static, flat):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        return out[0].clone(), out[1].clone()

    # ---- reference surface ---------------------------------------------------------------
    @lib.guarded
    def act(self, observations, rnn_states, prev_actions, masks, deterministic=False):
        """``policy.py:26-44``: features -> Categorical -> mode() / sample()."""
        with torch.no_grad():
            logits, rnn_states_out = self._logits_and_states(observations, rnn_states, prev_actions, masks)
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
