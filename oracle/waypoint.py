"""CPU fp32 restatement of the reference's waypoint path.
TEST INFRASTRUCTURE (see oracle/__init__).  Paths below are relative to
``/root/reference/vlnce_baselines``.

* ``TruncatedNormal``, ``single_query_attention``, ``PanoAttention``     <- ``models/utils.py:12-266``
* ``WaypointPredictionNet``                                              <- ``models/waypoint_predictors.py:29-625``
* ``WaypointPolicy`` (``act`` / ``get_value`` / ``evaluate_actions``)    <- ``models/waypoint_policy.py:19-347``
* ``wddppo_losses`` / ``WDDPPO``                                         <- ``common/ddppo_alg.py:9-149``
* ``ActionDictRolloutStorage``                                           <- ``common/rollout_storage.py:11-288``

Module / attribute names follow the reference so a reference ``state_dict`` loads
unchanged.  Pinned against the unmodified files by ``tests/test_oracle_vs_reference.py``
and frozen in ``tests/golden/wp_*.pt``.  Reference quirks that are kept on purpose:

* W1 the multiplicative, inverted instruction-attention mask;
* W2 ``text_q.repeat_interleave(12).view(B*12, -1)`` (``waypoint_predictors.py:459-464``)
  repeats *elements*, not rows: the query of pano frame ``r`` is
  ``q.flatten()[(r*D + j) // 12]`` — a scrambled vector, not a copy of row ``r // 12``;
* W3 ``single_spatial_filter=False`` keeps torchvision's global pool, so the 4x4 RGB "map"
  is one vector broadcast to 16 positions.
"""
from __future__ import annotations

import math
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import hb017, vlnce
from .hb017 import Config

PREV_ACTION_DIM = 4   # waypoint_predictors.py:24  sin, cos, offset, distance
PANO_KEY_DIM = 128    # :25
ANGLE_DIM = 4         # :26


def waypoint_model_config(**over) -> Config:
    """``config/default.py:214-285`` + ``config/r2r_waypoint/1-wpn-cc.yaml`` MODEL overrides
    (continuous offset & distance heads, ResNet-18 RGB, hidden 256)."""
    c = vlnce.default_model_config(
        policy_name="WaypointPolicy",
        INSTRUCTION_ENCODER=dict(bidirectional=True, final_state_only=False),
        RGB_ENCODER=dict(cnn_type="TorchVisionResNet18", output_size=128),
        STATE_ENCODER=dict(hidden_size=256),
        WAYPOINT=dict(predict_distance=True, continuous_distance=True, min_distance_var=0.01,
                      max_distance_var=3.516, max_distance_prediction=4.0,
                      min_distance_prediction=0.25, discrete_distances=6, predict_offset=True,
                      continuous_offset=True, min_offset_var=0.00030625,
                      max_offset_var=0.06853892, discrete_offsets=7, offset_temperature=4.0),
    )
    vlnce._merge(c, over)
    return Config(c)


# --------------------------------------------------------------------------- #
# models/utils.py:24-152
# --------------------------------------------------------------------------- #
_SQRT_2PI = math.sqrt(2.0 * math.pi)


def _phi(x):  # standard normal density, utils.py:72-74
    return torch.exp(-0.5 * x * x) / _SQRT_2PI


def _Phi(x):  # standard normal cdf (torch.distributions.Normal(0,1).cdf)
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


class TruncatedNormal:
    """Two-sided truncated normal on [lo, hi], parameterised by the location and scale of
    the untruncated normal.  ``mode`` is the location (utils.py:148-149)."""

    def __init__(self, loc, scale, lo, hi):
        assert lo < hi and math.isfinite(lo) and math.isfinite(hi)
        self.loc, self.scale, self.lo, self.hi = loc, scale, float(lo), float(hi)
        self._normal = torch.distributions.Normal(loc, scale, validate_args=False)
        a = (self.lo - loc) / scale
        b = (self.hi - loc) / scale
        pa, pb = _phi(a), _phi(b)
        z = _Phi(b) - _Phi(a)
        self.Z = self._normal.cdf(torch.as_tensor(self.hi)) - self._normal.cdf(torch.as_tensor(self.lo))
        self.mean = loc - scale * ((pb - pa) / z)                                   # :82-84
        self.variance = scale ** 2 * (1 - (b * pb - a * pa) / z - ((pb - pa) / z) ** 2)  # :86-88
        self._entropy = (0.5 * math.log(2 * math.pi * math.e) + torch.log(scale * z)
                         + (a * pa - b * pb) / (2 * z))                             # :90-94

    def mode(self):
        return self.loc

    def entropy(self):
        return self._entropy

    def log_prob(self, value):  # :126-146
        dens = torch.exp(-0.5 * ((value - self.loc) / self.scale) ** 2) / (self.scale * _SQRT_2PI)
        return (dens / self.Z).log()

    def sample(self, limit=10000):  # :106-124  rejection sampling, same RNG consumption
        s = self._normal.sample()
        bad = (s < self.lo) | (s > self.hi)
        n = 0
        while bad.any():
            assert n < limit
            n += 1
            s[bad] = self._normal.sample()[bad]
            bad = (s < self.lo) | (s > self.hi)
        return s


class PanoStopCategorical(torch.distributions.Categorical):
    """utils.py:269-292 (``log_prob`` spelling, (B,1) shapes)."""

    def sample(self, sample_shape=torch.Size()):  # noqa: B008
        return super().sample(sample_shape).unsqueeze(-1)

    def log_prob(self, actions):
        return super().log_prob(actions.squeeze(-1)).view(actions.size(0), -1).sum(-1).unsqueeze(-1)

    def mode(self):
        return self.probs.argmax(dim=-1, keepdim=True)


def single_query_attention(q, k, v, dk, mask=None):
    """utils.py:155-178: q (B,Dk), k (B,Dk,P), v (B,Dv,P); the optional mask MULTIPLIES the
    energies (W1)."""
    e = torch.einsum("bd,bdp->bp", q, k)
    if mask is not None:
        e = e * mask.float()
    a = F.softmax(e * (1.0 / dk ** 0.5), dim=1)
    return torch.einsum("bp,bdp->bd", a, v)


class PanoAttention(nn.Module):
    """utils.py:181-266 with the configuration the net uses (:131-139): one head, bias-free
    projections, no dropout, LayerNorm(eps=1e-6) on the output, no residual."""

    def __init__(self, d_q_in, d_kv_in, d_qk, d_v, d_out):
        super().__init__()
        self.q_linear = nn.Linear(d_q_in, d_qk, bias=False)
        self.k_linear = nn.Linear(d_kv_in, d_qk, bias=False)
        self.v_linear = nn.Linear(d_kv_in, d_v, bias=False)
        self.final_linear = nn.Linear(d_v, d_out, bias=False)
        self.layer_norm = nn.LayerNorm(d_out, eps=1e-6)
        self.d_qk = d_qk

    def forward(self, q, kv):  # q (B,Dq), kv (B,P,Dkv)
        k = self.k_linear(kv).transpose(1, 2)
        v = self.v_linear(kv).transpose(1, 2)
        att = single_query_attention(self.q_linear(q), k, v, self.d_qk)
        return self.layer_norm(self.final_linear(att))


class _TemperatureTanh(nn.Module):  # utils.py:12-21
    def __init__(self, temperature):
        super().__init__()
        self.t = float(temperature)

    def forward(self, x):
        return torch.tanh(x / self.t)


# --------------------------------------------------------------------------- #
# models/waypoint_predictors.py:29-625
# --------------------------------------------------------------------------- #
class WaypointPredictionNet(hb017.Net):
    def __init__(self, observation_space, model_config):
        super().__init__()
        cfg = self.model_config = model_config
        w = self.wypt_cfg = model_config.WAYPOINT
        H = self._hidden_size = cfg.STATE_ENCODER.hidden_size
        self._num_panos = cfg.num_panos
        rgb_out, dep_out = cfg.RGB_ENCODER.output_size, cfg.DEPTH_ENCODER.output_size

        self.instruction_encoder = vlnce.InstructionEncoder(cfg.INSTRUCTION_ENCODER)
        self.depth_encoder = vlnce.VlnResnetDepthEncoder(
            observation_space, output_size=dep_out, checkpoint=cfg.DEPTH_ENCODER.ddppo_checkpoint,
            backbone=cfg.DEPTH_ENCODER.backbone, spatial_output=True)                     # :42-50
        version = {"TorchVisionResNet50": "resnet50", "TorchVisionResNet18": "resnet18"}[
            cfg.RGB_ENCODER.cnn_type]
        self.rgb_encoder = vlnce.TorchVisionResNet(
            rgb_out, resnet_version=version, normalize_visual_inputs=cfg.normalize_rgb,
            spatial_output=True, single_spatial_filter=False)                             # :53-60 (W3)
        Di = self.instruction_encoder.output_size
        feat = rgb_out + dep_out + ANGLE_DIM

        self.visual_rnn = hb017.build_rnn_state_encoder(
            input_size=rgb_out + PREV_ACTION_DIM + dep_out + rgb_out, hidden_size=H,
            rnn_type=cfg.STATE_ENCODER.rnn_type, num_layers=1)                            # :63-74
        self.rgb_pool_linear = nn.Linear(self.rgb_encoder.resnet_layer_size, rgb_out)
        self.rgb_hist_linear = nn.Sequential(
            nn.AdaptiveAvgPool1d(1), nn.Flatten(),
            nn.Linear(self.rgb_encoder.output_shape[0], rgb_out), nn.ReLU(True))
        self.depth_hist_linear = nn.Sequential(
            nn.Flatten(), nn.Linear(int(np.prod(self.depth_encoder.output_shape)), dep_out),
            nn.ReLU(True))
        self.inst_attn_q = nn.Sequential(nn.Linear(H, H // 2), nn.ReLU(True))
        self.inst_attn_k = nn.Conv1d(Di, H // 2, 1)
        self.text_q_linear = nn.Linear(Di, H // 2)
        self.rgb_kv_spatial = nn.Conv1d(self.rgb_encoder.output_shape[0], H // 2 + rgb_out, 1)
        self.depth_kv_spatial = nn.Conv1d(self.depth_encoder.output_shape[0], H // 2 + dep_out, 1)
        self.pano_attn = PanoAttention(Di, feat, PANO_KEY_DIM, PANO_KEY_DIM, feat)        # :131-139
        self.main_state_compress = nn.Sequential(
            nn.Linear(Di + feat + H + PREV_ACTION_DIM, H), nn.ReLU(True))
        self.main_state_encoder = hb017.build_rnn_state_encoder(
            input_size=H, hidden_size=H, rnn_type=cfg.STATE_ENCODER.rnn_type, num_layers=1)
        self.stop_linear = nn.Linear(H, 1)
        nn.init.constant_(self.stop_linear.bias, 0)
        self.compress_x_linear = nn.Sequential(nn.Linear(H, feat), nn.ReLU(True))
        cat = H + feat
        if w.continuous_distance:                                                          # :223-243
            self.distance_linear = nn.Sequential(nn.Linear(cat, 1), nn.Sigmoid())
            self.distance_var_linear = nn.Sequential(nn.Linear(cat, 1), nn.Sigmoid())
        else:
            self.distance_linear = nn.Linear(cat, w.discrete_distances)
        if w.continuous_offset:                                                            # :245-266
            self.offset_linear = nn.Sequential(nn.Linear(cat, 1), _TemperatureTanh(w.offset_temperature))
            self.offset_scale = math.pi / self._num_panos
            self.offset_var_linear = nn.Sequential(nn.Linear(cat, 1), nn.Sigmoid())
        else:
            self.offset_linear = nn.Linear(cat, w.discrete_offsets)
        self.train()

    output_size = property(lambda self: self._hidden_size)
    num_recurrent_layers = property(
        lambda self: self.main_state_encoder.num_recurrent_layers + self.visual_rnn.num_recurrent_layers)

    def distance_to_continuous(self, d):  # :184-197
        w = self.wypt_cfg
        if w.continuous_distance:
            return d
        step = (w.max_distance_prediction - w.min_distance_prediction) / (w.discrete_distances - 1)
        return w.min_distance_prediction + d * step

    def offset_to_continuous(self, o):  # :199-209
        w = self.wypt_cfg
        if w.continuous_offset:
            return o
        per_pano = 2 * math.pi / self._num_panos
        return -per_pano / 2 + o * (per_pano / (w.discrete_offsets - 1))

    def _frames(self, observations, key, masks):
        """:327-366: the 12 pano frames + the history frame (zeroed at episode starts) as one
        (B*13, H, W, C) batch."""
        cur, hist = observations[key], observations[key + "_history"]
        hist = hist * masks.view(-1, 1, 1, 1).to(hist.dtype)
        allf = torch.cat([cur, hist.unsqueeze(1)], 1)
        return allf.flatten(0, 1), allf.shape[0], allf.shape[1]

    def forward(self, observations, rnn_states, prev_actions, masks):
        cfg, w, P = self.model_config, self.wypt_cfg, self._num_panos
        assert observations["rgb"].shape[1] == P and observations["depth"].shape[1] == P
        h_out = torch.zeros_like(rnn_states)
        instr = self.instruction_encoder(observations)                       # (B,Di,L)

        frames, B, F13 = self._frames(observations, "rgb", masks)
        rgb = self.rgb_encoder({"rgb": frames})
        rgb = rgb.view(B, F13, rgb.shape[1], -1)                             # (B,13,C,16)
        frames, _, _ = self._frames(observations, "depth", masks)
        dep = self.depth_encoder({"depth": frames})
        dep = dep.view(B, F13, dep.shape[1], -1)
        rgb_hist, rgb = rgb[:, P], rgb[:, :P].contiguous()
        dep_hist, dep = dep[:, P], dep[:, :P].contiguous()

        pa = {k: (v.unsqueeze(1) if v.dim() == 1 else v) for k, v in prev_actions.items()}
        heading = pa["pano"] * (2 * math.pi / P)                             # :268-275
        prev = torch.cat([torch.sin(heading), torch.cos(heading),
                          self.offset_to_continuous(pa["offset"]),
                          self.distance_to_continuous(pa["distance"])], 1).float() * masks

        if cfg.ablate_instruction:
            instr = instr * 0
        if cfg.ablate_rgb:
            rgb, rgb_hist = rgb * 0, rgb_hist * 0
        if cfg.ablate_depth:
            dep, dep_hist = dep * 0, dep_hist * 0

        # visual history GRU (:396-413); mean-pooled current RGB (:277-287)
        pooled = self.rgb_pool_linear(rgb[:, :, : self.rgb_encoder.resnet_layer_size].mean(3)).mean(1)
        n1 = self.visual_rnn.num_recurrent_layers
        hist_feats, h_out[:, 0:n1] = self.visual_rnn(
            torch.cat([pooled, prev, self.rgb_hist_linear(rgb_hist), self.depth_hist_linear(dep_hist)], 1),
            rnn_states[:, 0:n1], masks)

        # instruction attention (:419-424), W1
        H = self._hidden_size
        text = single_query_attention(self.inst_attn_q(hist_feats), self.inst_attn_k(instr), instr, H // 2,
                                      mask=(instr == 0.0).all(dim=1))

        # spatial attention inside every pano frame (:431-489), W2
        q = self.text_q_linear(text)
        q = q.repeat_interleave(P).view(B * P, -1)
        rgb_kv = self.rgb_kv_spatial(rgb.flatten(0, 1))
        dep_kv = self.depth_kv_spatial(dep.flatten(0, 1))
        att_rgb = single_query_attention(q, rgb_kv[:, : H // 2], rgb_kv[:, H // 2:], H // 2).view(B, P, -1)
        att_dep = single_query_attention(q, dep_kv[:, : H // 2], dep_kv[:, H // 2:], H // 2).view(B, P, -1)

        # attention over the pano frames (:495-509)
        frame_feats = torch.cat([att_rgb, att_dep, observations["angle_features"]], 2)   # (B,P,feat)
        pano = self.pano_attn(text, frame_feats)

        # main recurrence (:515-540)
        x = self.main_state_compress(torch.cat([text, pano, hist_feats, prev], 1))
        x, h_out[:, n1:] = self.main_state_encoder(x, rnn_states[:, n1:], masks)

        # heads (:546-615)
        dotted = (frame_feats * self.compress_x_linear(x).unsqueeze(1)).sum(2)
        pano_stop = PanoStopCategorical(logits=torch.cat([dotted, self.stop_linear(x)], 1))
        cat = torch.cat([frame_feats, x.unsqueeze(1).expand(-1, P, -1)], 2)
        if w.continuous_distance:
            span = w.max_distance_prediction - w.min_distance_prediction
            d1 = span * self.distance_linear(cat).squeeze(2) + w.min_distance_prediction
            d2 = ((w.max_distance_var - w.min_distance_var) * self.distance_var_linear(cat).squeeze(2)
                  + w.min_distance_var)
        else:
            d1, d2 = self.distance_linear(cat).squeeze(2), None
        if w.continuous_offset:
            o1 = self.offset_scale * self.offset_linear(cat).squeeze(2)
            o2 = ((w.max_offset_var - w.min_offset_var) * self.offset_var_linear(cat).squeeze(2)
                  + w.min_offset_var)
        else:
            o1, o2 = self.offset_linear(cat).squeeze(2), None
        return pano_stop, o1, o2, d1, d2, x, h_out


# --------------------------------------------------------------------------- #
# models/waypoint_policy.py:19-347
# --------------------------------------------------------------------------- #
def _select(x, idx):  # utils.py:295-317 batched_index_select along dim 1
    return torch.gather(x, 1, idx.view(-1, 1, *([1] * (x.dim() - 2))).expand(-1, 1, *x.shape[2:])).squeeze(1)


class WaypointPolicy(hb017.Policy):
    def __init__(self, observation_space, action_space, model_config):
        super().__init__(WaypointPredictionNet(observation_space, model_config), 1)
        self._config = model_config
        self.wypt_cfg = model_config.WAYPOINT
        self._offset_limit = math.pi / model_config.num_panos

    @classmethod
    def from_config(cls, config, observation_space, action_space):  # :335-347
        config.MODEL.num_panos = config.TASK_CONFIG.TASK.PANO_ROTATIONS
        return cls(observation_space, action_space, config.MODEL)

    def _component(self, which, v1, v2, pano):  # :41-81
        w = self.wypt_cfg
        if getattr(w, "continuous_" + which):
            lo, hi = ((w.min_distance_prediction, w.max_distance_prediction) if which == "distance"
                      else (-self._offset_limit, self._offset_limit))
            return TruncatedNormal(torch.gather(v1, 1, pano), torch.sqrt(torch.gather(v2, 1, pano)), lo, hi)
        return PanoStopCategorical(logits=_select(v1, pano))

    def _predict(self, which, dist, deterministic):  # :83-127
        w = self.wypt_cfg
        val = dist.mode() if deterministic else dist.sample()
        lp = dist.log_prob(val)
        to_cont = getattr(self.net, which + "_to_continuous")
        act, var, mode = to_cont(val), dist.variance, dist.mode()
        if not getattr(w, "predict_" + which):
            if which == "offset":
                act, val = torch.zeros_like(act), torch.zeros_like(val)
                if val.dtype == torch.int64:
                    val = val * (w.discrete_offsets // 2)
            else:
                act = torch.zeros_like(act) + 0.25
                val = torch.zeros_like(val) if val.dtype == torch.int64 else torch.zeros_like(val) + 0.25
            var = torch.zeros_like(var)
        return val, act, lp, var, mode

    def act(self, observations, rnn_states, prev_actions, masks, deterministic=False):  # :129-247
        w, P = self.wypt_cfg, self._config.num_panos
        pano_stop_dist, o1, o2, d1, d2, x, h_out = self.net(observations, rnn_states, prev_actions, masks)
        pano_stop = pano_stop_dist.mode() if deterministic else pano_stop_dist.sample()
        stop = (pano_stop == P).to(torch.uint8)
        pano = pano_stop % P
        d_dist = self._component("distance", d1, d2, pano)
        o_dist = self._component("offset", o1, o2, pano)
        dist, a_dist, d_lp, d_var, d_mode = self._predict("distance", d_dist, deterministic)
        ofs, a_ofs, o_lp, o_var, o_mode = self._predict("offset", o_dist, deterministic)
        theta = (pano * (2 * math.pi / P) + a_ofs) % (2 * math.pi)
        actions = []
        for i in range(pano_stop.shape[0]):
            if stop[i]:
                actions.append({"action": "STOP"})
            else:
                actions.append({"action": {"action": "GO_TOWARD_POINT",
                                           "action_args": {"r": a_dist[i].item(), "theta": theta[i].item()}}})
        lp = pano_stop_dist.log_prob(pano_stop)
        go = (pano_stop != P).to(lp.dtype)
        if w.predict_distance:
            lp = lp + go * w.predict_distance * d_lp
        if w.predict_offset:
            lp = lp + go * w.predict_offset * o_lp
        return (self.critic(x), actions, {"pano": pano_stop, "offset": ofs, "distance": dist},
                {"offset": o_mode, "distance": d_mode}, {"distance": d_var, "offset": o_var}, lp, h_out,
                pano_stop_dist)

    def get_value(self, observations, rnn_states, prev_actions, masks):  # :249-263
        return self.critic(self.net(observations, rnn_states, prev_actions, masks)[5])

    def evaluate_actions(self, observations, rnn_states, prev_actions, masks, action_components):  # :265-333
        w, P = self.wypt_cfg, self._config.num_panos
        pano_stop_dist, o1, o2, d1, d2, x, h_out = self.net(observations, rnn_states, prev_actions, masks)
        pano = action_components["pano"]
        lp_pano = pano_stop_dist.log_prob(pano)
        idx = pano.to(torch.int64) % P
        d_dist = self._component("distance", d1, d2, idx)
        o_dist = self._component("offset", o1, o2, idx)
        go = (pano != P).to(lp_pano.dtype)
        dm, om = go * w.predict_distance, go * w.predict_offset
        lp = lp_pano + dm * d_dist.log_prob(action_components["distance"]) \
            + om * o_dist.log_prob(action_components["offset"])
        ent = {"pano": pano_stop_dist.entropy(), "offset": (om * o_dist.entropy()).squeeze(1),
               "distance": (dm * d_dist.entropy()).squeeze(1)}
        return self.critic(x), lp, ent, h_out


# --------------------------------------------------------------------------- #
# common/ddppo_alg.py:9-149  (+ habitat_baselines v0.1.7 PPO: Adam over requires_grad
# parameters, clip_grad_norm_ before the step)
# --------------------------------------------------------------------------- #
def wddppo_losses(policy, sample, clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01,
                  use_clipped_value_loss=True, offset_regularize_coef=0.1146, pano_entropy_coef=1.0,
                  offset_entropy_coef=1.0, distance_entropy_coef=1.0):
    """One mini-batch of ``WDDPPO.update`` (:56-124) up to the scalar loss.  Returns
    (loss, {value_loss, action_loss, entropy_loss, offset_loss, entropies...})."""
    obs, h0, actions, prev, v_old, ret, masks, lp_old, adv = sample
    values, lp, ent, h_out = policy.evaluate_actions(obs, h0, prev, masks, actions)
    entropy_loss = (pano_entropy_coef * ent["pano"] + offset_entropy_coef * ent["offset"]
                    + distance_entropy_coef * ent["distance"]).mean() * entropy_coef
    ratio = torch.exp(lp - lp_old)
    action_loss = -torch.min(ratio * adv, torch.clamp(ratio, 1 - clip_param, 1 + clip_param) * adv).mean()
    if use_clipped_value_loss:
        v_clip = v_old + (values - v_old).clamp(-clip_param, clip_param)
        value_loss = 0.5 * torch.max((values - ret).pow(2), (v_clip - ret).pow(2)).mean()
    else:
        value_loss = 0.5 * (ret - values).pow(2).mean()
    value_loss = value_loss * value_loss_coef
    offset_loss = 0.0
    if "offset" in actions:  # :110-118 — an L1 penalty on the TAKEN offsets: a constant w.r.t. the parameters
        offset_loss = offset_regularize_coef * policy.net.offset_to_continuous(actions["offset"]).abs().mean()
    loss = value_loss + action_loss + offset_loss - entropy_loss
    return loss, dict(value_loss=value_loss, action_loss=action_loss, entropy_loss=entropy_loss,
                      offset_loss=offset_loss, pano_entropy=ent["pano"].mean(),
                      offset_entropy=ent["offset"].mean(), distance_entropy=ent["distance"].mean(),
                      values=values, action_log_probs=lp, rnn_states_out=h_out, entropy_pano=ent["pano"],
                      entropy_offset=ent["offset"], entropy_distance=ent["distance"])


def advantages_of(rollouts, normalize=True):  # :30-35
    adv = rollouts.returns[:-1] - rollouts.value_preds[:-1]
    return (adv - adv.mean()) / (adv.std() + 1e-5) if normalize else adv


# --------------------------------------------------------------------------- #
# common/rollout_storage.py:11-288
# --------------------------------------------------------------------------- #
class ActionDictRolloutStorage:
    KEYS = ("pano", "offset", "distance")

    def __init__(self, num_steps, num_envs, observation_space, hidden_size, num_recurrent_layers=1,
                 continuous_offset=True, continuous_distance=True):
        T, N = num_steps, num_envs
        self.observations = {k: torch.zeros(T + 1, N, *sp.shape) for k, sp in observation_space.spaces.items()}
        self.recurrent_hidden_states = torch.zeros(T + 1, N, num_recurrent_layers, hidden_size)
        self.rewards = torch.zeros(T, N, 1)
        self.value_preds = torch.zeros(T + 1, N, 1)
        self.returns = torch.zeros(T + 1, N, 1)
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

    def insert(self, observations, hidden, action, action_log_probs, value_preds, rewards, masks):  # :87-113
        s = self.step
        for k, v in observations.items():
            self.observations[k][s + 1].copy_(v)
        self.recurrent_hidden_states[s + 1].copy_(hidden)
        for k, v in action.items():
            self.actions[k][s].copy_(v)
            self.prev_actions[k][s + 1].copy_(v)
        self.action_log_probs[s].copy_(action_log_probs)
        self.value_preds[s].copy_(value_preds)
        self.rewards[s].copy_(rewards)
        self.masks[s + 1].copy_(masks)
        self.step = s + 1

    def after_update(self):  # :115-125
        s = self.step
        for v in self.observations.values():
            v[0].copy_(v[s])
        self.recurrent_hidden_states[0].copy_(self.recurrent_hidden_states[s])
        self.masks[0].copy_(self.masks[s])
        for v in self.prev_actions.values():
            v[0].copy_(v[s])
        self.step = 0

    def compute_returns(self, next_value, use_gae, gamma, tau):  # :127-152
        if use_gae:
            self.value_preds[self.step] = next_value
            gae = 0
            for t in reversed(range(self.step)):
                m = self.masks[t + 1]
                delta = self.rewards[t] + gamma * self.value_preds[t + 1] * m - self.value_preds[t]
                gae = delta + gamma * tau * m * gae
                self.returns[t] = gae + self.value_preds[t]
        else:
            self.returns[self.step] = next_value
            for t in reversed(range(self.step)):
                self.returns[t] = self.returns[t + 1] * gamma * self.masks[t + 1] + self.rewards[t]

    def recurrent_generator(self, advantages, num_mini_batch, perm=None):  # :154-276
        N = self.rewards.size(1)
        assert N >= num_mini_batch
        per = N // num_mini_batch
        perm = torch.randperm(N) if perm is None else perm
        T = self.step
        for start in range(0, N, per):
            ind = perm[start:start + per]

            def tm(x):  # (T, n, ...) columns ``ind`` -> time-major rows (T*n, ...)
                y = x[:T][:, ind]
                return y.reshape(T * y.shape[1], *y.shape[2:])

            yield ({k: tm(v) for k, v in self.observations.items()},
                   self.recurrent_hidden_states[0][ind],
                   {k: tm(v) for k, v in self.actions.items()},
                   {k: tm(v) for k, v in self.prev_actions.items()},
                   tm(self.value_preds), tm(self.returns), tm(self.masks), tm(self.action_log_probs),
                   tm(advantages))


def ppo_update(policy, optimizer, rollouts, ppo_epoch=2, num_mini_batch=4, max_grad_norm=0.2,
               use_normalized_advantage=False, **loss_kw):
    """``WDDPPO.update`` (ddppo_alg.py:37-149): epochs x recurrent mini-batches of
    evaluate_actions -> loss -> backward -> clip_grad_norm_ (hb PPO.before_step) -> Adam.
    Returns the six epoch-averaged scalars in the reference's order."""
    adv = advantages_of(rollouts, use_normalized_advantage)
    keys = ("value_loss", "action_loss", "entropy_loss", "pano_entropy", "offset_entropy", "distance_entropy")
    sums = dict.fromkeys(keys, 0.0)
    for _ in range(ppo_epoch):
        for sample in rollouts.recurrent_generator(adv, num_mini_batch):
            loss, parts = wddppo_losses(policy, sample, **loss_kw)
            optimizer.zero_grad()
            loss.backward()
            nn.utils.clip_grad_norm_(policy.parameters(), max_grad_norm)
            optimizer.step()
            for k in keys:
                sums[k] += float(parts[k].detach()) if torch.is_tensor(parts[k]) else float(parts[k])
    n = ppo_epoch * num_mini_batch
    return tuple(sums[k] / n for k in keys)


def storage_obs_space(rgb=(12, 224, 224, 3), depth=(12, 256, 256, 1), length=80):
    """The observation space the waypoint trainer hands to the rollout storage
    (``ddppo_waypoint_trainer.py:60-100``: policy sensors + the two history frames)."""
    return hb017.Dict({
        "rgb": hb017.Box(0, 255, rgb, "float32"), "depth": hb017.Box(0.0, 1.0, depth, "float32"),
        "rgb_history": hb017.Box(0, 255, rgb[1:], "float32"),
        "depth_history": hb017.Box(0.0, 1.0, depth[1:], "float32"),
        "angle_features": hb017.Box(-1.0, 1.0, (rgb[0], 4), "float32"),
        "instruction": hb017.Box(0, 2503, (length,), "float32"),
    })
