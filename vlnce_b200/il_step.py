"""Fused imitation-learning update — the per-step loss of the reference's
``BaseVLNCETrainer._update_agent`` (``common/base_il_trainer.py:134-180``) as ONE launch
sequence: forward (trunks + head) -> fused inflection-weighted CE + progress-monitor loss and
their gradients -> hand-written head backward -> [one NCCL all-reduce of the flat gradient]
-> fused Adam.  Same arguments, same three returned loss values, same accumulation semantics
(``step_grad`` / ``loss_accumulation_scalar``, ``recollect_trainer.py:161-185``).

Data parallelism: trajectories are sharded across ranks, parameters replicated;
the only collective is an all-reduce(sum) of the flat trainable-gradient buffer, the 1/world
factor is folded into the Adam kernel.  The reference's IL trainers are single-GPU; the
averaging mirrors its DD-PPO reducer (``ddppo_waypoint_trainer.py:370``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import lib
from .aux_losses import AuxLosses
from .backward import trainable_named
from .lib import C, ptr


import os as _os

HEAD_PRIORITY = _os.environ.get("VLNCE_HEAD_PRIORITY", "1") != "0"


class FlatParams:
    """All trainable parameters as views of one flat fp32 buffer (+ flat grad / Adam moments),
    so the optimiser and the gradient all-reduce are single launches."""

    def __init__(self, named: Dict[str, torch.nn.Parameter]):
        self.named = named
        dev = next(iter(named.values())).device
        self.sizes = {k: p.numel() for k, p in named.items()}
        total = sum(self.sizes.values())
        self.flat = torch.empty(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad_views: Dict[str, torch.Tensor] = {}
        o = 0
        with torch.no_grad():
            for k, p in named.items():
                n = self.sizes[k]
                self.flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)  # re-point: state_dict/load_state_dict keep working
                self.grad_views[k] = self.grad[o:o + n].view(p.shape)
                o += n
        self.total = total

    def aliased(self) -> bool:
        o = 0
        for k, p in self.named.items():
            if p.data_ptr() != self.flat.data_ptr() + 4 * o:
                return False
            o += self.sizes[k]
        return True


class ILTrainerStep:
    def __init__(self, policy, lr: float = 2.5e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None):
        self.policy = policy
        self.lr, self.betas, self.eps = lr, betas, eps
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.fp: Optional[FlatParams] = None
        self.step_id = 0
        self._step_dev = None
        self._accum = None
        self._pending = 0
        # opt-in: rows whose loss weight is 0 (the padding of a ragged teacher-forcing batch) bypass the frozen
        # eval-mode conv trunks (engine.encode_visual); costs one host read of the valid-row count per update
        self.skip_padded_rows = False
        # optional: the indices of the real rows of the NEXT update (int64, time-major row = t * N + n), e.g. built on
        # the host from the trajectory lengths the collate step already knows — saves the host read above
        self.valid_rows_hint = None

    def _prepare(self):
        if self.fp is None or not self.fp.aliased():
            named = trainable_named(self.policy.net, self.policy.action_distribution.linear)
            self.fp = FlatParams(named)
            dev = self.fp.flat.device
            self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._step_dev.fill_(self.step_id)
            self._accum = torch.zeros_like(self.fp.grad)
            self._losses = torch.zeros(3, device=dev)
            self._loss_acc = torch.zeros(3, device=dev)

    # ---------------------------------------------------------------- the reference signature
    @lib.guarded
    def update_agent(self, observations, prev_actions, not_done_masks, corrected_actions, weights,
                     step_grad: bool = True, loss_accumulation_scalar: int = 1, return_tensor: bool = False):
        self.compute_grads(observations, prev_actions, not_done_masks, corrected_actions, weights,
                           loss_accumulation_scalar)
        self.apply_grads(step_grad)
        if return_tensor:
            return self._losses
        loss, action_loss, aux_loss = self._losses.tolist()  # one D2H read (the reference does three .item())
        return loss, action_loss, aux_loss

    @lib.guarded
    def update_agent_from_host(self, observations, prev_actions, not_done_masks, corrected_actions, weights, **kw):
        """``update_agent`` for a batch that still lives in host memory (what a DataLoader hands the reference's
        ``_update_agent`` before its ``.to(device)`` calls, ``dagger_trainer.py:560-583``): every tensor travels to the
        policy's device with ONE non-blocking copy into a persistent staging buffer (pinned sources overlap with the
        kernels already queued), then the fused step runs on the staged batch.  Same return value as ``update_agent``."""
        dev = next(self.policy.parameters()).device
        stage = self.__dict__.setdefault("_host_stage", {})

        def put(name, t):
            buf = stage.get(name)
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype or buf.device != dev:
                buf = stage[name] = torch.empty(t.shape, dtype=t.dtype, device=dev)
            buf.copy_(t, non_blocking=True)
            return buf

        obs = {k: put("obs/" + k, v) for k, v in observations.items()}
        return self.update_agent(obs, put("prev_actions", prev_actions), put("not_done_masks", not_done_masks),
                                 put("corrected_actions", corrected_actions), put("weights", weights), **kw)

    # The two phases are separately CUDA-graph capturable; the NCCL all-reduce sits between them.
    # ---- pipelined training loop: the frozen trunks of batch i+1 run concurrently with the head
    # forward/backward/Adam of batch i (they depend on no trainable parameter); bit-identical results
    @lib.guarded
    def prefetch_visual(self, observations_next):
        """Trunk features of the NEXT batch into the 'next' slot (call inside a forked stream or
        let ``pipelined_step`` do it)."""
        eng = self.policy.net.engine()
        net = self.policy.net
        if getattr(net.rgb_encoder, "trainable", False) or getattr(net.depth_encoder, "trainable", False):
            raise lib.VlnceError("the pipelined loop runs the trunks ahead of the update: it needs frozen trunks "
                                 "(RGB_ENCODER.trainable / DEPTH_ENCODER.trainable = False)")
        rgb = observations_next["rgb"] if "rgb" in observations_next else observations_next["rgb_features"]
        eng.ensure_pool(rgb.device)
        R = rgb.shape[0]
        eng.encode_visual(observations_next, *eng.visual_buffers(R, "next"))

    @lib.guarded
    def rotate_visual(self, R):
        eng = self.policy.net.engine()
        for cur, nxt in zip(eng.visual_buffers(R, "cur"), eng.visual_buffers(R, "next")):
            cur.copy_(nxt)

    @lib.guarded
    def pipelined_step(self, observations, prev_actions, not_done_masks, corrected_actions, weights,
                       observations_next, apply: bool = True):
        """One update on the CURRENT batch (its trunk features were prefetched by the previous
        call) while the trunks of ``observations_next`` run on a forked stream."""
        from .engine import Branch

        R = prev_actions.shape[0]
        with Branch("visual_next", priority=0) as br:
            self.prefetch_visual(observations_next)
        # the head update is a long chain of short kernels next to the trunks' SM-filling ones: on a high-priority
        # stream its CTAs are placed first whenever a conv CTA retires (VLNCE_HEAD_PRIORITY=0: same priority)
        with Branch("head", enabled=None if HEAD_PRIORITY else False, priority=-1 if HEAD_PRIORITY else 0) as bh:
            self.compute_grads(observations, prev_actions, not_done_masks, corrected_actions, weights,
                               visual_slot="cur")
            if apply:
                self.apply_grads(True)
        bh.join()
        br.join()
        self.rotate_visual(R)
        return self._losses

    @lib.guarded
    def compute_grads(self, observations, prev_actions, not_done_masks, corrected_actions, weights,
                      loss_accumulation_scalar: int = 1, visual_slot=None):
        pol = self.policy
        net = pol.net
        self._prepare()
        eng = net.engine()
        fp = self.fp
        st = lib.stream()
        T, N = corrected_actions.shape
        dev = fp.flat.device
        h0 = eng.ensure_pool(dev).get("il_h0", (N, net.num_recurrent_layers, eng.H), torch.float32)
        h0.zero_()  # base_il_trainer.py:145-150
        AuxLosses.clear()
        R = T * N
        visual = eng.visual_buffers(R, visual_slot) if visual_slot is not None else None
        eng.valid_rows = None
        if self.skip_padded_rows and visual is None and not (weights.is_cuda and torch.cuda.is_current_stream_capturing()):
            idx = self.valid_rows_hint
            self.valid_rows_hint = None
            if idx is None:
                idx = (weights.reshape(-1) > 0).nonzero().reshape(-1)   # (T, N) time-major == the row order
            if 0 < idx.numel() < R:
                eng.valid_rows = idx.to(weights.device, non_blocking=True)
        try:
            x, _, hat = eng.forward(observations, h0, prev_actions, not_done_masks, save=True, visual=visual)
        finally:
            eng.valid_rows = None
        head = pol.action_distribution.linear
        A = head.weight.shape[0]
        logp = eng.action_logits(head, x, R)
        use_pm = bool(net.model_config.PROGRESS_MONITOR.use) and AuxLosses.is_active()
        mode, progress = 0, None
        if use_pm:
            progress = observations["progress"]
            lib.require_cuda(progress, "progress")
            mode = 2 if progress.dim() == 2 else 1
            progress = progress.contiguous().float()
        actions = corrected_actions.contiguous()
        if actions.dtype != torch.int64:
            actions = actions.long()
        w = weights.contiguous().float()
        d_raw = eng.buf("d_raw", (R, A))
        d_hat = eng.buf("d_hat", (R, 1))
        C.vlnce_il_loss(ptr(logp), ptr(actions), ptr(w), ptr(hat) if use_pm else None,
                        ptr(progress) if use_pm else None, mode,
                        float(net.model_config.PROGRESS_MONITOR.alpha), 1.0 / float(loss_accumulation_scalar),
                        T, N, A, ptr(d_raw), ptr(d_hat), ptr(self._losses), st)
        fp.grad.zero_()
        eng.run_backward(head, fp.grad_views, d_raw, None, d_hat if use_pm else None)

    @lib.guarded
    def apply_grads(self, step_grad: bool = True, allreduce: bool = True):
        fp = self.fp
        st = lib.stream()
        accumulating = self._pending > 0 or not step_grad
        if accumulating:
            C.vlnce_axpy(ptr(self._accum), ptr(fp.grad), 1.0, fp.total, st)
            self._pending += 1
        if step_grad:
            g = self._accum if accumulating else fp.grad
            if self.world > 1 and allreduce:
                torch.distributed.all_reduce(g, group=self.pg)  # NCCL sum over NVLink; 1/world folded below
            C.vlnce_counter_add(ptr(self._step_dev), 1, st)
            self.step_id += 1
            C.vlnce_adam_step(ptr(fp.flat), ptr(g), ptr(fp.exp_avg), ptr(fp.exp_avg_sq), fp.total, float(self.lr),
                              float(self.betas[0]), float(self.betas[1]), float(self.eps), 0, ptr(self._step_dev),
                              1.0 / self.world, st)
            if accumulating:
                self._accum.zero_()
                self._pending = 0
