"""Synthetic observation batches of the shapes BASELINE.json quotes the metric on
(the simulator is stubbed: north_star).  Shapes/dtypes follow what the reference
feeds ``policy.act`` / ``_update_agent``:

* ``rgb``  uint8 ``(R,H,W,3)`` 0..255, ``depth`` f32 ``(R,H,W,1)`` 0..1 — sensor
  layout, ``habitat_extensions/config/vlnce_task.yaml`` + ``resnet_encoders.py:198``;
* ``instruction`` int64 ``(R,L)`` token ids, 0 = PAD (``instruction_encoder.py:70-78``);
* time-major ``(T*N, ...)`` row layout, ``masks`` zero on the first row-block,
  inflection weights — ``dagger_trainer.py:39-114,199-211``;
* ``progress`` f32 ``(R,1)`` (``habitat_extensions/sensors.py:72,85-87``).

Generated on the CPU with a seeded ``torch.Generator`` so the oracle and the CUDA
path see bit-identical inputs.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def make_instruction(gen, rows: int, length: int = 80, vocab: int = 2504,
                     min_len: int = 40) -> torch.Tensor:
    ids = torch.randint(1, vocab, (rows, length), generator=gen, dtype=torch.int64)
    lens = torch.randint(min(min_len, length), length + 1, (rows,), generator=gen)
    pos = torch.arange(length).unsqueeze(0)
    return ids * (pos < lens.unsqueeze(1))


def make_rxr_instruction(gen, rows: int, length: int = 512, dim: int = 768,
                         min_len: int = 256) -> torch.Tensor:
    x = torch.randn(rows, length, dim, generator=gen)
    lens = torch.randint(min(min_len, length), length + 1, (rows,), generator=gen)
    pos = torch.arange(length).view(1, length, 1)
    return x * (pos < lens.view(rows, 1, 1))


def make_act_batch(seed: int, n: int, hw: int = 256, rgb_hw: int | None = None,
                   length: int = 80, hidden: int = 512, layers: int = 2,
                   vocab: int = 2504, num_actions: int = 4,
                   fresh: bool = False, rxr_dim: int = 0) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, torch.Tensor, torch.Tensor]:
    """One ``policy.act`` call: (observations, rnn_states, prev_actions, masks).  ``rxr_dim`` > 0: the
    instruction arrives as precomputed multilingual-BERT features ``rxr_instruction`` (R, L, rxr_dim), zero rows
    beyond each instruction's length (``rxr_cma_en.yaml:43-48``, ``instruction_encoder.py:74-78``)."""
    gen = torch.Generator().manual_seed(seed)
    rgb_hw = rgb_hw or hw
    obs = {
        "rgb": torch.randint(0, 256, (n, rgb_hw, rgb_hw, 3), generator=gen, dtype=torch.uint8),
        "depth": torch.rand(n, hw, hw, 1, generator=gen),
    }
    if rxr_dim:
        obs["rxr_instruction"] = make_rxr_instruction(gen, n, length, rxr_dim, min_len=length // 2)
    else:
        obs["instruction"] = make_instruction(gen, n, length, vocab)
    if fresh:
        h = torch.zeros(n, layers, hidden)
        prev = torch.zeros(n, 1, dtype=torch.int64)
        masks = torch.zeros(n, 1, dtype=torch.uint8)
    else:
        h = torch.randn(n, layers, hidden, generator=gen) * 0.5
        prev = torch.randint(0, num_actions, (n, 1), generator=gen, dtype=torch.int64)
        masks = (torch.rand(n, 1, generator=gen) > 0.2).to(torch.uint8)
    return obs, h, prev, masks


def make_train_batch(seed: int, t: int, n: int, hw: int = 256, rgb_hw: int | None = None,
                     length: int = 80, vocab: int = 2504, num_actions: int = 4,
                     coef: float = 3.2, ragged: bool = True, features: bool = False, rgb_feature_hw: int = 4,
                     rxr_dim: int = 0):
    """One ``_update_agent`` call in the collate_fn layout: returns
    (observations, prev_actions (T*N,1), not_done_masks (T*N,1) u8,
    corrected_actions (T,N) i64, weights (T,N) f32)."""
    gen = torch.Generator().manual_seed(seed)
    rgb_hw = rgb_hw or hw
    r = t * n
    # the same instruction for all T rows of a trajectory (dagger_trainer.py:342-362)
    if rxr_dim:
        feats = make_rxr_instruction(gen, n, length, rxr_dim, min_len=length // 2)
        obs: Dict[str, torch.Tensor] = {
            "rxr_instruction": feats.unsqueeze(0).expand(t, n, length, rxr_dim).reshape(r, length, rxr_dim).contiguous()}
    else:
        instr = make_instruction(gen, n, length, vocab).unsqueeze(0).expand(t, n, length).reshape(r, length)
        obs = {"instruction": instr.contiguous()}
    if features:  # DAgger caches frozen-CNN features (dagger_trainer.py:300-314)
        # spatial encoders (CMA) cache the 4x4 map, the Seq2Seq encoder its globally pooled 1x1 output
        obs["rgb_features"] = torch.rand(r, 2048, rgb_feature_hw, rgb_feature_hw, generator=gen)
        obs["depth_features"] = torch.rand(r, 128, 4, 4, generator=gen)
    else:
        obs["rgb"] = torch.randint(0, 256, (r, rgb_hw, rgb_hw, 3), generator=gen, dtype=torch.uint8)
        obs["depth"] = torch.rand(r, hw, hw, 1, generator=gen)
    obs["progress"] = torch.rand(r, 1, generator=gen)
    actions = torch.randint(0, num_actions, (t, n), generator=gen, dtype=torch.int64)
    lens = torch.randint(max(1, t // 2), t + 1, (n,), generator=gen) if ragged else torch.full((n,), t)
    lens[0] = t
    valid = torch.arange(t).unsqueeze(1) < lens.unsqueeze(0)  # (T,N)
    infl = torch.ones(t, n, dtype=torch.bool)
    infl[1:] = actions[1:] != actions[:-1]
    weights = torch.where(infl, torch.tensor(coef), torch.tensor(1.0)) * valid
    actions = actions * valid
    prev = torch.zeros(t, n, dtype=torch.int64)
    prev[1:] = actions[:-1]
    prev = prev * valid
    masks = torch.ones(t, n, dtype=torch.uint8)
    masks[0] = 0
    return obs, prev.view(-1, 1), masks.view(-1, 1), actions, weights.float()


def seeded_state_dict(module: torch.nn.Module, seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic "trained-like" weights for any of the policies, a pure function
    of (parameter names, shapes, seed): plain ``torch.randn`` / ``torch.rand`` on a
    seeded CPU generator (no LAPACK-dependent initialisers), so the build container,
    the GPU box and the golden fixtures all regenerate bit-identical tensors.

    There is no network for checkpoints (``resnet_encoders.py:136`` pretrained
    ImageNet weights, ``default.py:247-250`` DD-PPO depth weights), so every parity
    test and the benchmark use these.  Scales keep activations O(1) through the
    16 residual blocks (last norm of each block damped) and give the action head a
    usable logit spread (SURVEY hard-part 2: orthogonal gain 0.01 leaves ~1e-2)."""
    gen = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    sd = module.state_dict()
    basic_blocks = not any(".bn3." in k for k in sd)  # torchvision ResNet-18/34: bn2 closes a block
    for name, ref in sd.items():
        shape = tuple(ref.shape)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros_like(ref)
            continue
        if name.endswith("_scale"):
            out[name] = ref.clone()
            continue
        if leaf == "running_mean":
            t = torch.randn(shape, generator=gen) * 0.1
        elif leaf == "running_var":
            t = torch.rand(shape, generator=gen) + 0.5
        elif ref.dim() == 1 and leaf == "weight":  # norm scale
            damp = (".bn3." in name or name.endswith("convs.7.weight")
                    or ".downsample.1." in name or name.endswith("convs.4.weight")
                    or (basic_blocks and ".bn2." in name))
            t = torch.rand(shape, generator=gen) * (0.2 if damp else 1.0) + (0.2 if damp else 0.5)
        elif ref.dim() == 1 or "bias" in leaf:
            t = torch.randn(shape, generator=gen) * 0.05
        elif "embedding" in name:
            t = torch.randn(shape, generator=gen) * 0.5
            if name.endswith("embedding_layer.weight"):
                t[0] = 0  # PAD row (instruction_encoder.py:41-45 padding_idx=0)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 2.0 if ref.dim() == 4 else 1.0
            if "action_distribution" in name:
                gain = 4.0
            if "_var_linear." in name or "distance_linear." in name or "offset_linear." in name:
                gain = 8.0  # waypoint heads: spread the sigmoid / tanh outputs without saturating them
            t = torch.randn(shape, generator=gen) * (gain / fan_in) ** 0.5
        out[name] = t.to(ref.dtype)
    return out


# ---------------------------------------------------------------------------------------------
# Waypoint path: panoramic observations of ``1-wpn-cc.yaml``
# ---------------------------------------------------------------------------------------------
def angle_features(n: int, panos: int = 12) -> torch.Tensor:
    """``habitat_extensions/sensors.py:102-105``: [sin, cos, sin, cos] of each pano heading."""
    a = torch.arange(panos, dtype=torch.float32) * (2 * 3.141592653589793 / panos)
    f = torch.stack([torch.sin(a), torch.cos(a), torch.sin(a), torch.cos(a)], 1)
    return f.unsqueeze(0).expand(n, panos, 4).contiguous()


def make_waypoint_rows(gen, rows: int, rgb_hw: int = 224, hw: int = 256, length: int = 80, panos: int = 12,
                       vocab: int = 2504, rgb_u8: bool = False, instruction=None) -> Dict[str, torch.Tensor]:
    """Observations of ``rows`` policy steps in the layout ``ddppo_waypoint_trainer.py:222-246``
    feeds the policy: every sensor float32 (``batch_obs`` + ``.float()``), rgb 0..255."""
    rgb = torch.randint(0, 256, (rows, panos, rgb_hw, rgb_hw, 3), generator=gen, dtype=torch.uint8)
    rgb_h = torch.randint(0, 256, (rows, rgb_hw, rgb_hw, 3), generator=gen, dtype=torch.uint8)
    instr = make_instruction(gen, rows, length, vocab, min_len=length // 2) if instruction is None else instruction
    return {
        "rgb": rgb if rgb_u8 else rgb.float(),
        "depth": torch.rand(rows, panos, hw, hw, 1, generator=gen),
        "rgb_history": rgb_h if rgb_u8 else rgb_h.float(),
        "depth_history": torch.rand(rows, hw, hw, 1, generator=gen),
        "angle_features": angle_features(rows, panos),
        "instruction": instr if rgb_u8 else instr.float(),
    }


def make_waypoint_actions(gen, rows: int, panos: int = 12, dmin: float = 0.25, dmax: float = 4.0, classes=None):
    """Action components as ``WaypointPolicy.act`` emits them: pano in 0..panos (panos = STOP); continuous heads:
    offset in [-pi/panos, pi/panos], distance in [dmin, dmax]; discrete heads (``classes[k]`` > 0): class indices."""
    lim = 3.141592653589793 / panos
    classes = classes or {}
    out = {
        "pano": torch.randint(0, panos + 1, (rows, 1), generator=gen, dtype=torch.int64),
        "offset": (torch.rand(rows, 1, generator=gen) * 2 - 1) * lim * 0.98,
        "distance": dmin + torch.rand(rows, 1, generator=gen) * (dmax - dmin),
    }
    for k in ("offset", "distance"):
        if classes.get(k, 0):
            out[k] = torch.randint(0, classes[k], (rows, 1), generator=gen, dtype=torch.int64)
    return out


def make_waypoint_act_batch(seed: int, n: int, hidden: int = 256, layers: int = 2, rgb_u8: bool = False,
                            classes=None, **kw):
    """One ``WaypointPolicy.act`` call: (observations, rnn_states, prev_actions, masks)."""
    gen = torch.Generator().manual_seed(seed)
    obs = make_waypoint_rows(gen, n, rgb_u8=rgb_u8, **kw)
    h = torch.randn(n, layers, hidden, generator=gen) * 0.5
    prev = make_waypoint_actions(gen, n, classes=classes)
    masks = torch.ones(n, 1, dtype=torch.uint8)
    masks[0] = 0  # one env at an episode start (history frame and prev action zeroed)
    return obs, h, prev, masks


def make_waypoint_eval_batch(seed: int, t: int, n: int, hidden: int = 256, layers: int = 2, classes=None, **kw):
    """One PPO mini-batch as ``ActionDictRolloutStorage.recurrent_generator`` yields it
    (``rollout_storage.py:154-276``), time-major rows: (obs, h0 (N,Lr,H), actions, prev_actions,
    value_preds, returns, masks, old_action_log_probs, adv_targ)."""
    gen = torch.Generator().manual_seed(seed)
    r = t * n
    length = kw.pop("length", 80)
    instr = make_instruction(gen, n, length, min_len=length // 2)
    instr = instr.unsqueeze(0).expand(t, n, length).reshape(r, length).contiguous()
    obs = make_waypoint_rows(gen, r, length=length, instruction=instr, **kw)
    h0 = torch.randn(n, layers, hidden, generator=gen) * 0.5
    actions = make_waypoint_actions(gen, r, classes=classes)
    actions["pano"][min(1, r - 1)] = 12  # make sure a STOP row is covered
    prev = {k: torch.zeros_like(v) for k, v in actions.items()}
    for k in actions:
        prev[k][n:] = actions[k][:-n]
    masks = (torch.rand(r, 1, generator=gen) > 0.25).to(torch.uint8)
    prev = {k: v * masks.to(v.dtype) for k, v in prev.items()}
    actions = {k: v.float() for k, v in actions.items()}  # the storage keeps actions as float (rollout_storage.py:49-52)
    v_old = torch.randn(r, 1, generator=gen) * 0.5
    ret = v_old + torch.randn(r, 1, generator=gen) * 0.5
    lp_old = -2.0 + torch.randn(r, 1, generator=gen) * 0.3
    adv = torch.randn(r, 1, generator=gen)
    return obs, h0, actions, prev, v_old, ret, masks, lp_old, adv


# ---------------------------------------------------------------------------------------------
# Seeded policies of the benchmarked configuration, built from the package alone (the benchmark's
# product arm must not depend on the checker package ``oracle/``).
# ---------------------------------------------------------------------------------------------
def cma_model_config(progress_monitor: bool = True, rxr_dim: int = 0, **over):
    """``vlnce_baselines/config/default.py:214-285`` MODEL node with the ``cma.yaml:31-35`` /
    ``cma_pm.yaml`` overrides (bi-directional instruction encoder, progress monitor) and, for
    ``rxr_dim`` > 0, ``rxr_cma_en.yaml:43-48`` (``rxr_instruction`` sensor, 768-d features)."""
    from .compat import Config

    c = dict(
        policy_name="CMAPolicy", normalize_rgb=False, ablate_depth=False, ablate_rgb=False, ablate_instruction=False,
        INSTRUCTION_ENCODER=dict(sensor_uuid="instruction", vocab_size=2504, use_pretrained_embeddings=False,
                                 embedding_file="", dataset_vocab="", fine_tune_embeddings=False, embedding_size=50,
                                 hidden_size=128, rnn_type="LSTM", final_state_only=True, bidirectional=True),
        RGB_ENCODER=dict(cnn_type="TorchVisionResNet50", output_size=256, trainable=False),
        DEPTH_ENCODER=dict(cnn_type="VlnResnetDepthEncoder", output_size=128, backbone="resnet50",
                           ddppo_checkpoint="NONE", trainable=False),
        STATE_ENCODER=dict(hidden_size=512, rnn_type="GRU"),
        PROGRESS_MONITOR=dict(use=bool(progress_monitor), alpha=1.0),
        SEQ2SEQ=dict(use_prev_action=False),
    )
    if rxr_dim:
        c["INSTRUCTION_ENCODER"].update(sensor_uuid="rxr_instruction", embedding_size=int(rxr_dim))
    for k, v in over.items():
        if isinstance(v, dict) and k in c:
            c[k].update(v)
        else:
            c[k] = v
    return Config(c)


class _no_random_init:
    """Every tensor is overwritten by ``seeded_state_dict`` right after construction: skip the constructors'
    random initialisers (orthogonal QR / kaiming over ~30 M weights)."""

    NAMES = ("kaiming_normal_", "kaiming_uniform_", "orthogonal_", "normal_", "uniform_", "xavier_uniform_",
             "xavier_normal_", "trunc_normal_")

    def __enter__(self):
        self.saved = {n: getattr(torch.nn.init, n) for n in self.NAMES}
        for n in self.NAMES:
            setattr(torch.nn.init, n, lambda tensor, *a, **k: tensor)

    def __exit__(self, *exc):
        for n, f in self.saved.items():
            setattr(torch.nn.init, n, f)


def build_cma_policy(wseed: int = 13, hw: int = 256, rgb_hw: int | None = None, num_actions: int = 4,
                     progress_monitor: bool = True, rxr_dim: int = 0, **over):
    """``CMAPolicy.from_config`` on the benchmark's observation space with ``seeded_state_dict(wseed)`` weights.
    Defaults = BASELINE config C2 (the same construction + seed as the parity case ``cma_train_pixels``, so the CPU
    checker can rebuild the identical policy on its side)."""
    from .cma_policy import CMAPolicy
    from .compat import Box, Config, Dict as SpaceDict, Discrete

    rgb_hw = rgb_hw or hw
    space = SpaceDict({"rgb": Box(0, 255, (rgb_hw, rgb_hw, 3), "uint8"),
                       "depth": Box(0.0, 1.0, (hw, hw, 1), "float32")})
    cfg = Config(MODEL=cma_model_config(progress_monitor, rxr_dim, **over), TORCH_GPU_ID=0,
                 TASK_CONFIG=dict(TASK=dict(PANO_ROTATIONS=12)))
    with _no_random_init():
        pol = CMAPolicy.from_config(cfg, space, Discrete(num_actions))
    pol.load_state_dict(seeded_state_dict(pol, wseed), strict=True)
    return pol
