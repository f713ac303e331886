"""DD-PPO side of the waypoint path on the device:

* ``ActionDictRolloutStorage`` — the reference's rollout container (``common/rollout_storage.py:11-288``) with
  the same constructor, ``insert`` / ``after_update`` / ``compute_returns`` / ``recurrent_generator`` surface;
  tensors live on the GPU, the GAE recursion is one kernel (``vlnce_gae_returns``) instead of a Python loop of
  T x 6 tiny launches, and the image sensors may be kept in uint8 (``obs_dtypes``) instead of the reference's
  float32 (12 x 224 x 224 x 3 x 4 B per env-step).
* ``WDDPPOStep`` — ``WDDPPO.update`` (``common/ddppo_alg.py:37-149``) as a fused launch sequence per
  mini-batch: forward (trunks + head) -> ``vlnce_wp_ppo_loss`` (evaluate_actions + clipped surrogate + clipped
  value loss + entropies + offset regulariser AND their gradients, one kernel) -> hand-written backward ->
  [one NCCL all-reduce of the flat gradient, mirroring the DD-PPO reducer's average] -> ``clip_grad_norm_`` ->
  Adam.  Same constructor keywords and the same six returned scalars.  The reference's own ``WDDPPO`` also
  runs unchanged on ``WaypointPolicy.evaluate_actions`` (one autograd node) — this class is the fast path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import lib
from .il_step import FlatParams
from .lib import C, ptr
from .waypoint_policy import WaypointPolicy, _host_floats, trainable_named


class ActionDictRolloutStorage:
    KEYS = ("pano", "offset", "distance")

    def __init__(self, num_steps, num_envs, observation_space, recurrent_hidden_state_size, num_recurrent_layers=1,
                 continuous_offset=True, continuous_distance=True, obs_dtypes: Optional[Dict[str, torch.dtype]] = None):
        T, N = num_steps, num_envs
        obs_dtypes = obs_dtypes or {}
        self.observations = {k: torch.zeros(T + 1, N, *sp.shape, dtype=obs_dtypes.get(k, torch.float32))
                             for k, sp in observation_space.spaces.items()}
        self.recurrent_hidden_states = torch.zeros(T + 1, N, num_recurrent_layers, recurrent_hidden_state_size)
        self.rewards = torch.zeros(T, N, 1)
        self.value_preds = torch.zeros(T + 1, N, 1)
        self.returns = torch.zeros(T + 1, N, 1)
        self.advantages = torch.zeros(T, N, 1)
        self.action_log_probs = torch.zeros(T, N, 1)
        self.actions = {k: torch.zeros(T, N, 1) for k in self.KEYS}
        self.prev_actions = {k: torch.zeros(T + 1, N, 1) for k in self.KEYS}
        self.prev_actions["pano"] = self.prev_actions["pano"].long()
        if not continuous_distance:
            self.prev_actions["distance"] = self.prev_actions["distance"].long()
        if not continuous_offset:
            self.prev_actions["offset"] = self.prev_actions["offset"].long()
        self.masks = torch.zeros(T + 1, N, 1, dtype=torch.uint8)
        self.num_steps, self.step = T, 0
        self.visual_cache = False

    RAW_FRAME_KEYS = ("rgb", "depth", "rgb_history", "depth_history")

    def enable_visual_cache(self, policy):
        """B200-first extension (not in the reference, whose update re-runs both ResNets on the stored frames,
        ``ddppo_alg.py:56-75`` -> ``waypoint_policy.py:265-333``): keep the FROZEN trunks' features of every stored step
        (``WaypointPolicy.visual_features()`` after ``act``) beside the observations; ``recurrent_generator`` then hands
        the update those features instead of the raw frames and ``evaluate_actions`` skips the trunks.  Identical
        results as long as the trunks are frozen and in eval mode — every waypoint recipe (``config/default.py:242,253``)."""
        from .waypoint_policy import VISUAL_CACHE_KEYS

        eng = policy.net.engine()
        T1, N = self.masks.shape[0], self.masks.shape[1]
        dev = self.masks.device
        self.observations[VISUAL_CACHE_KEYS[0]] = torch.zeros(T1, N, eng.P + 1, eng.c_rgb, device=dev)
        self.observations[VISUAL_CACHE_KEYS[1]] = torch.zeros(T1, N, eng.P + 1, eng.S, eng.c_dep, device=dev)
        self.visual_cache = True

    def store_visual_features(self, features: Dict[str, torch.Tensor]):
        """features of the observations in the CURRENT slot (call between ``act`` and ``insert``)."""
        for k, v in features.items():
            self.observations[k][self.step].copy_(v)

    def to(self, device):
        for d in (self.observations, self.actions, self.prev_actions):
            for k in d:
                d[k] = d[k].to(device)
        for name in ("recurrent_hidden_states", "rewards", "value_preds", "returns", "advantages", "action_log_probs", "masks"):
            setattr(self, name, getattr(self, name).to(device))

    def insert(self, observations, recurrent_hidden_states, action, action_log_probs, value_preds, rewards, masks):
        s = self.step
        for k, v in observations.items():
            self.observations[k][s + 1].copy_(v)
        self.recurrent_hidden_states[s + 1].copy_(recurrent_hidden_states)
        for k, v in action.items():
            self.actions[k][s].copy_(v)
            self.prev_actions[k][s + 1].copy_(v)
        self.action_log_probs[s].copy_(action_log_probs)
        self.value_preds[s].copy_(value_preds)
        self.rewards[s].copy_(rewards)
        self.masks[s + 1].copy_(masks)
        self.step = s + 1

    def after_update(self):
        s = self.step
        for v in self.observations.values():
            v[0].copy_(v[s])
        self.recurrent_hidden_states[0].copy_(self.recurrent_hidden_states[s])
        self.masks[0].copy_(self.masks[s])
        for v in self.prev_actions.values():
            v[0].copy_(v[s])
        self.step = 0

    def compute_returns(self, next_value, use_gae, gamma, tau):
        """``rollout_storage.py:127-152`` in one launch; also fills ``self.advantages`` (returns - value_preds,
        ``ddppo_alg.py:30-32``)."""
        lib.require_cuda(self.rewards, "rollout storage (call .to(device) first)")
        T, N = self.step, self.rewards.shape[1]
        nv = next_value.reshape(N).float().contiguous()
        C.vlnce_gae_returns(ptr(self.rewards), ptr(self.value_preds), ptr(self.masks), ptr(nv), ptr(self.returns),
                            ptr(self.advantages), T, N, float(gamma), float(tau), int(bool(use_gae)), lib.stream())

    def recurrent_generator(self, advantages, num_mini_batch, perm=None):
        """``rollout_storage.py:154-276``: env columns are split into ``num_mini_batch`` groups (random order) and
        every group is yielded as time-major rows (T*n, ...) — gathers only, no arithmetic."""
        N = self.rewards.size(1)
        assert N >= num_mini_batch, (
            f"Trainer requires the number of processes ({N}) to be greater than or equal to the number of trainer "
            f"mini batches ({num_mini_batch}).")
        per = N // num_mini_batch
        perm = torch.randperm(N) if perm is None else perm
        T = self.step
        dev = self.rewards.device
        for start in range(0, N, per):
            ind = perm[start:start + per].to(dev)

            def tm(x):
                y = x[:T].index_select(1, ind)
                return y.reshape(T * y.shape[1], *y.shape[2:])

            skip = self.RAW_FRAME_KEYS if self.visual_cache else ()
            yield ({k: tm(v) for k, v in self.observations.items() if k not in skip},
                   self.recurrent_hidden_states[0].index_select(0, ind),
                   {k: tm(v) for k, v in self.actions.items()},
                   {k: tm(v) for k, v in self.prev_actions.items()},
                   tm(self.value_preds), tm(self.returns), tm(self.masks), tm(self.action_log_probs), tm(advantages))


class FusedAdam(torch.optim.Optimizer):
    """The optimiser object the reference trainer talks to (``agent.optimizer``: ``state_dict()`` into checkpoints,
    ``LambdaLR(agent.optimizer)`` for the linear lr decay, ``ddppo_waypoint_trainer.py:425,485``).  The update itself
    is the fused ``vlnce_adam_step`` launch of ``WDDPPOStep.apply_grads`` on the flat parameter buffer; this class
    holds the hyper-parameters (``param_groups[0]["lr"]`` is read at every step, so schedulers work) and serialises
    the flat moments."""

    def __init__(self, params, lr, betas, eps, owner):
        super().__init__(list(params), dict(lr=lr, betas=betas, eps=eps))
        self._owner = owner

    def step(self, closure=None):  # the fused path applies the update inside WDDPPOStep.apply_grads
        raise RuntimeError("FusedAdam is stepped by WDDPPOStep.update / apply_grads")

    def state_dict(self):
        o = self._owner
        o._prepare()
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        return {"state": {"step": int(o.step_id), "exp_avg": o.fp.exp_avg.detach().clone(),
                          "exp_avg_sq": o.fp.exp_avg_sq.detach().clone(), "names": list(o.fp.named)},
                "param_groups": groups}

    def load_state_dict(self, sd):
        o = self._owner
        o._prepare()
        st = sd["state"]
        if list(st.get("names", list(o.fp.named))) != list(o.fp.named):
            raise ValueError("optimizer state belongs to a different parameter layout")
        o.fp.exp_avg.copy_(st["exp_avg"])
        o.fp.exp_avg_sq.copy_(st["exp_avg_sq"])
        o.step_id = int(st["step"])
        o._step_dev.fill_(o.step_id)
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in saved.items() if k != "params"})


class WDDPPOStep(torch.nn.Module):
    """``nn.Module`` like the reference's ``WDDPPO`` (habitat_baselines ``PPO``): ``state_dict()`` keys are prefixed
    ``actor_critic.`` (``ddppo_waypoint_trainer.py:141-143,484``), ``.optimizer`` exists, and ``clip_param`` may be
    reassigned between updates (``use_linear_clip_decay``)."""

    LOSS_KEYS = ("loss", "value_loss", "action_loss", "entropy_loss", "offset_loss", "pano_entropy", "offset_entropy",
                 "distance_entropy")

    def __init__(self, actor_critic: WaypointPolicy, clip_param=0.2, ppo_epoch=2, num_mini_batch=4, value_loss_coef=0.5,
                 entropy_coef=0.01, lr=2.0e-4, eps=1e-5, max_grad_norm=0.2, use_clipped_value_loss=True,
                 use_normalized_advantage=False, offset_regularize_coef=0.1146, pano_entropy_coef=1.0,
                 offset_entropy_coef=0.0, distance_entropy_coef=0.0, betas=(0.9, 0.999), process_group=None):
        super().__init__()
        self.actor_critic = actor_critic
        self.clip_param, self.ppo_epoch, self.num_mini_batch = clip_param, ppo_epoch, num_mini_batch
        self.eps, self.betas, self.max_grad_norm = eps, betas, max_grad_norm
        self.use_normalized_advantage = use_normalized_advantage
        self._loss_coefs = [value_loss_coef, entropy_coef, offset_regularize_coef, pano_entropy_coef,
                            offset_entropy_coef, distance_entropy_coef, float(bool(use_clipped_value_loss))]
        self.ppo_consts = _host_floats([clip_param] + self._loss_coefs)
        self.optimizer = FusedAdam((p for p in actor_critic.parameters() if p.requires_grad), lr, betas, eps, self)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.fp: Optional[FlatParams] = None
        self.step_id = 0
        self._graphs = None

    @property
    def lr(self):
        return float(self.optimizer.param_groups[0]["lr"])

    @lr.setter
    def lr(self, value):
        self.optimizer.param_groups[0]["lr"] = float(value)

    def forward(self, *x):
        raise NotImplementedError

    def _prepare(self):
        if self.fp is None or not self.fp.aliased():
            old = self.fp
            self.fp = FlatParams(trainable_named(self.actor_critic))
            if old is not None and old.total == self.fp.total:  # parameters were re-pointed (load_state_dict / .to()): keep the moments
                self.fp.exp_avg.copy_(old.exp_avg)
                self.fp.exp_avg_sq.copy_(old.exp_avg_sq)
            dev = self.fp.flat.device
            self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._step_dev.fill_(self.step_id)
            self._losses = torch.zeros(8, device=dev)
            self._clip_ws = torch.zeros(1026, device=dev)

    def get_advantages(self, rollouts):  # ddppo_alg.py:30-35
        adv = rollouts.advantages[: rollouts.step] if hasattr(rollouts, "advantages") else (
            rollouts.returns[:-1] - rollouts.value_preds[:-1])
        if not self.use_normalized_advantage:
            return adv
        return (adv - adv.mean()) / (adv.std() + 1e-5)  # optional, off in every shipped recipe: (T,N) torch ops

    @lib.guarded
    def update_minibatch(self, sample, apply: bool = True):
        """One mini-batch of ``WDDPPO.update`` (:56-140).  Returns the device tensor of ``LOSS_KEYS``."""
        obs, h0, actions, prev_actions, v_old, ret, masks, lp_old, adv = sample
        pol = self.actor_critic
        self._prepare()
        self.ppo_consts[0] = float(self.clip_param)  # the trainer may decay it between updates (use_linear_clip_decay)
        eng = pol.net.engine()
        fp = self.fp
        st = lib.stream()
        out = eng.forward(obs, h0, prev_actions, masks, save=True, critic=pol.critic)
        B, P = out["B"], eng.P
        acts = pol._action_rows(actions, B)
        flat = lambda t: t.reshape(B).float().contiguous()  # noqa: E731
        bufs = pol._grad_buffers(eng, B, P)
        d_values = eng.buf("wp_d_values", (B, 1))
        ev = eng.buf("wp_eval_out", (B, 4))
        C.vlnce_wp_ppo_loss(ptr(out["logp"]), ptr(out["o1"]), ptr(out["o2"]), ptr(out["d1"]), ptr(out["d2"]),
                            ptr(out["values"]), ptr(acts[0]), ptr(acts[1]), ptr(acts[2]), ptr(flat(v_old)), ptr(flat(ret)),
                            ptr(flat(lp_old)), ptr(flat(adv)), B, P, ctypes.addressof(eng.act_consts),
                            ctypes.addressof(self.ppo_consts), ptr(ev), ptr(self._losses), *(ptr(t) for t in bufs),
                            ptr(d_values), st)
        fp.grad.zero_()
        eng.backward(fp.grad_views, pol.critic, *bufs, d_values)
        if apply:
            self.apply_grads()
        return self._losses

    @lib.guarded
    def apply_grads(self):
        fp = self.fp
        st = lib.stream()
        if self.world > 1:
            torch.distributed.all_reduce(fp.grad, group=self.pg)  # NCCL sum; the 1/world average is folded into the clip
        C.vlnce_clip_grad_norm(ptr(fp.grad), fp.total, float(self.max_grad_norm), 1.0 / self.world, ptr(self._clip_ws),
                               self._clip_ws.numel() * 4, st)
        C.vlnce_counter_add(ptr(self._step_dev), 1, st)
        self.step_id += 1
        C.vlnce_adam_step(ptr(fp.flat), ptr(fp.grad), ptr(fp.exp_avg), ptr(fp.exp_avg_sq), fp.total, float(self.lr),
                          float(self.betas[0]), float(self.betas[1]), float(self.eps), 0, ptr(self._step_dev), 1.0, st)

    # ---- CUDA-graph replay of a mini-batch (the eager step is ~490 launches at ~13 us of host time each) ----------
    use_cuda_graph = True     # instance / class switch; VLNCE_PPO_GRAPH=0 in the environment turns it off

    @staticmethod
    def _tensors(sample):
        out = []
        for part in sample:
            if isinstance(part, dict):
                out.extend(part[k] for k in sorted(part))
            else:
                out.append(part)
        return out

    def _update_minibatch_graphed(self, sample):
        """Replays a captured ``update_minibatch`` (forward, fused loss, backward, [NCCL all-reduce], clip, Adam) on
        static copies of the sample; one graph per sample signature, built on first use after an eager pass without
        the optimiser step (plans, tile selection and workspace allocation happen there)."""
        tensors = self._tensors(sample)
        key = tuple((tuple(t.shape), t.dtype) for t in tensors)
        if self._graphs is None:
            self._graphs = {}
        entry = self._graphs.get(key)
        hyper = (float(self.lr), float(self.clip_param))   # launch arguments baked into the captured kernels
        if entry is None or entry[2] != hyper:
            if entry is None:
                static = tuple({k: v.clone() for k, v in part.items()} if isinstance(part, dict) else part.clone()
                               for part in sample)
                self.update_minibatch(static, apply=False)
                torch.cuda.synchronize()
            else:   # lr schedule / clip decay moved on: re-record with the same static buffers (no eager pass needed)
                static = entry[0]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    self.update_minibatch(static, apply=True)
            torch.cuda.current_stream().wait_stream(side)
            self.step_id -= 1          # the captured call was recorded, not run
            entry = (static, graph, hyper)
            self._graphs[key] = entry
        static, graph = entry[0], entry[1]
        for dst, src in zip(self._tensors(static), tensors):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        self.step_id += 1
        return self._losses

    @lib.guarded
    def update(self, rollouts):
        """``WDDPPO.update``: returns (value_loss, action_loss, entropy_loss, pano_entropy, offset_entropy,
        distance_entropy) averaged over epochs x mini-batches — ONE D2H read at the end instead of six
        ``.item()`` per mini-batch."""
        advantages = self.get_advantages(rollouts)
        self._prepare()
        acc = torch.zeros_like(self._losses)
        import os

        graphed = (self.use_cuda_graph and os.environ.get("VLNCE_PPO_GRAPH", "1") != "0" and self._losses.is_cuda
                   and not torch.cuda.is_current_stream_capturing())
        for _ in range(self.ppo_epoch):
            for sample in rollouts.recurrent_generator(advantages, self.num_mini_batch):
                if graphed:
                    try:
                        self._update_minibatch_graphed(sample)
                    except Exception as e:  # capture is an optimisation: fall back to eager launches for good
                        import warnings

                        warnings.warn(f"WDDPPOStep: CUDA-graph capture of the mini-batch failed ({type(e).__name__}: {e}); "
                                      "running eagerly")
                        torch.cuda.synchronize()
                        graphed = self.use_cuda_graph = False
                        self._graphs = None
                        self.update_minibatch(sample)
                else:
                    self.update_minibatch(sample)
                C.vlnce_axpy(ptr(acc), ptr(self._losses), 1.0, 8, lib.stream())
        n = self.ppo_epoch * self.num_mini_batch
        v = (acc / n).tolist()
        return v[1], v[2], v[3], v[5], v[6], v[7]


def select_history_frames(step_observation: Dict[str, torch.Tensor], pano_stop: torch.Tensor, num_panos: int,
                          keys=("rgb", "depth")) -> Dict[str, torch.Tensor]:
    """The observation-history frames of the next step (``ddppo_waypoint_trainer.py:190-199``): for every env that
    did not STOP, the pano frame it moved towards; zeros otherwise.  The reference loops over the envs with a
    ``.item()`` each; this is one batched gather per sensor (data movement only, no host sync)."""
    idx = pano_stop.reshape(-1).long()
    go = idx != num_panos
    sel = torch.where(go, idx, torch.zeros_like(idx))
    out = {}
    for k in keys:
        frames = step_observation[k]                                   # (N, P, H, W, C)
        n = frames.shape[0]
        picked = frames[torch.arange(n, device=frames.device), sel]     # (N, H, W, C)
        out[k] = torch.where(go.view(n, 1, 1, 1), picked, torch.zeros_like(picked))
    return out
