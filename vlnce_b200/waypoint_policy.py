"""``WaypointPolicy`` / ``WaypointPredictionNet`` — the waypoint policy of the reference
(``vlnce_baselines/models/waypoint_policy.py:19-347``, ``waypoint_predictors.py:29-625``) as a
drop-in: same registry name, constructor, ``from_config``, ``act`` / ``get_value`` /
``evaluate_actions`` signatures and return values, attribute surface
(``net.num_recurrent_layers``, ``net.rgb_encoder.eval()``, ``critic.fc``,
``net.distance_to_continuous`` ...) and ``state_dict`` keys.  The torch sub-modules only hold
parameters; the arithmetic runs in the CUDA library (``csrc/waypoint.cu`` + the head kernels the
CMA policy uses), sequenced by ``WaypointEngine``.

All six recipes of ``config/r2r_waypoint`` are covered: continuous (TruncatedNormal) and discrete (categorical)
distance / offset heads in any combination, with the ``predict_distance`` / ``predict_offset`` switches
(1-wpn-cc = BASELINE config C4, 2-wpn-dc, 3-wpn-dd, 4-wpn-d_, 5-hpn-_c, 6-hpn-__).

Algebra kept from the CMA engine: single-query attentions are re-associated (keys / values are never
materialised); additionally ``mean_frames(Linear(mean_pos(x))) = Linear(mean(x))`` for the pooled RGB
branch.  Reference quirks kept bug-for-bug: W1 (multiplicative inverted instruction mask, closed form
in ``vlnce_wp_inst_attn_fwd``), W2 (element-wise ``repeat_interleave`` of the spatial query), W3 (the
globally pooled RGB vector broadcast to the 16 "spatial" positions).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from . import lib
from .cma_policy import _GRUStateEncoder
from .compat import CustomFixedCategorical, Net, Policy, baseline_registry
from .encoders import resnet_encoders
from .encoders.instruction_encoder import InstructionEncoder
from .engine import Branch, HeadEngine, linear, linear_dw, linear_dx, off
from .lib import C, ptr
from .policy import CategoricalNet

PREV_ACTION_DIM = 4
PANO_ATTN_KEY_DIM = 128
ANGLE_FEATURE_SIZE = 4


# observation keys of the frozen-trunk feature cache (WaypointEngine.export_visual_features / forward)
VISUAL_CACHE_KEYS = ("pano_rgb_features", "pano_depth_features")


def _host_floats(vals):
    """Small constant block handed to a kernel wrapper by HOST pointer (kept alive by the caller)."""
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


class _PanoAttention(nn.Module):
    """Parameter container of ``MultiHeadDotProductAttention`` (``models/utils.py:181-266``) as the net
    configures it (``waypoint_predictors.py:131-139``): one head, bias-free, LayerNorm(eps 1e-6)."""

    def __init__(self, d_q_in, d_kv_in, d_qk, d_v, d_out):
        super().__init__()
        self.q_linear = nn.Linear(d_q_in, d_qk, bias=False)
        self.k_linear = nn.Linear(d_kv_in, d_qk, bias=False)
        self.v_linear = nn.Linear(d_kv_in, d_v, bias=False)
        self.final_linear = nn.Linear(d_v, d_out, bias=False)
        self.layer_norm = nn.LayerNorm(d_out, eps=1e-6)


class _CriticHead(nn.Module):
    """habitat_baselines v0.1.7 ``CriticHead``: ``fc`` attribute, orthogonal weight, zero bias
    (re-initialised by the trainer, ``ddppo_waypoint_trainer.py:118-120``)."""

    def __init__(self, input_size):
        super().__init__()
        self.fc = nn.Linear(input_size, 1)
        nn.init.orthogonal_(self.fc.weight)
        nn.init.constant_(self.fc.bias, 0)


class WaypointPredictionNet(Net):
    def __init__(self, observation_space, model_config):
        nn.Module.__init__(self)
        cfg = self.model_config = model_config
        w = self.wypt_cfg = model_config.WAYPOINT
        H = self._hidden_size = cfg.STATE_ENCODER.hidden_size
        self._num_panos = cfg.num_panos
        rgb_out, dep_out = cfg.RGB_ENCODER.output_size, cfg.DEPTH_ENCODER.output_size
        self.instruction_encoder = InstructionEncoder(cfg.INSTRUCTION_ENCODER)
        assert cfg.DEPTH_ENCODER.cnn_type in ["VlnResnetDepthEncoder"]
        frame = resnet_encoders_single_frame(observation_space)
        self.depth_encoder = getattr(resnet_encoders, cfg.DEPTH_ENCODER.cnn_type)(
            frame, output_size=dep_out, checkpoint=cfg.DEPTH_ENCODER.ddppo_checkpoint,
            backbone=cfg.DEPTH_ENCODER.backbone, spatial_output=True)
        assert cfg.RGB_ENCODER.cnn_type in ["TorchVisionResNet18", "TorchVisionResNet50"]
        self.rgb_encoder = getattr(resnet_encoders, cfg.RGB_ENCODER.cnn_type)(
            rgb_out, normalize_visual_inputs=cfg.normalize_rgb, spatial_output=True, single_spatial_filter=False)
        Di = self.instruction_encoder.output_size
        feat = rgb_out + dep_out + ANGLE_FEATURE_SIZE
        self.visual_rnn = _GRUStateEncoder(rgb_out + PREV_ACTION_DIM + dep_out + rgb_out, H,
                                           cfg.STATE_ENCODER.rnn_type, 1)
        self.rgb_pool_linear = nn.Linear(self.rgb_encoder.resnet_layer_size, rgb_out)
        self.rgb_hist_linear = nn.Sequential(nn.AdaptiveAvgPool1d(1), nn.Flatten(),
                                             nn.Linear(self.rgb_encoder.output_shape[0], rgb_out), nn.ReLU(True))
        self.depth_hist_linear = nn.Sequential(
            nn.Flatten(), nn.Linear(int(np.prod(self.depth_encoder.output_shape)), dep_out), nn.ReLU(True))
        self.inst_attn_q = nn.Sequential(nn.Linear(H, H // 2), nn.ReLU(True))
        self.inst_attn_k = nn.Conv1d(Di, H // 2, 1)
        self.text_q_linear = nn.Linear(Di, H // 2)
        self.rgb_kv_spatial = nn.Conv1d(self.rgb_encoder.output_shape[0], H // 2 + rgb_out, 1)
        self.depth_kv_spatial = nn.Conv1d(self.depth_encoder.output_shape[0], H // 2 + dep_out, 1)
        self.pano_attn = _PanoAttention(Di, feat, PANO_ATTN_KEY_DIM, PANO_ATTN_KEY_DIM, feat)
        self.main_state_compress = nn.Sequential(nn.Linear(Di + feat + H + PREV_ACTION_DIM, H), nn.ReLU(True))
        self.main_state_encoder = _GRUStateEncoder(H, H, cfg.STATE_ENCODER.rnn_type, 1)
        self.stop_linear = nn.Linear(H, 1)
        nn.init.constant_(self.stop_linear.bias, 0)
        self.compress_x_linear = nn.Sequential(nn.Linear(H, feat), nn.ReLU(True))
        cat = H + feat
        if w.continuous_distance:  # waypoint_predictors.py:223-243
            self.distance_linear = nn.Sequential(nn.Linear(cat, 1), nn.Sigmoid())
            self.distance_var_linear = nn.Sequential(nn.Linear(cat, 1), nn.Sigmoid())
        else:
            self.distance_linear = nn.Linear(cat, w.discrete_distances)
        self.offset_scale = math.pi / self._num_panos
        if w.continuous_offset:    # :245-266
            self.offset_linear = nn.Sequential(nn.Linear(cat, 1), nn.Tanh())  # TemperatureTanh: parameter-free
            self.offset_var_linear = nn.Sequential(nn.Linear(cat, 1), nn.Sigmoid())
        else:
            self.offset_linear = nn.Linear(cat, w.discrete_offsets)
        self.precision = "bf16"
        self.train()

    # ---- reference attribute surface --------------------------------------------------
    def to_continuous_consts(self):
        """(dist_a, dist_b, off_a, off_b): ``{distance,offset}_to_continuous(v) = a + b * v``
        (waypoint_predictors.py:184-209; identity for the continuous heads)."""
        w = self.wypt_cfg
        da, db, oa, ob = 0.0, 1.0, 0.0, 1.0
        if not w.continuous_distance:
            da = w.min_distance_prediction
            db = (w.max_distance_prediction - w.min_distance_prediction) / (w.discrete_distances - 1)
        if not w.continuous_offset:
            per_pano = 2 * math.pi / self._num_panos
            oa, ob = -per_pano / 2, per_pano / (w.discrete_offsets - 1)
        return da, db, oa, ob

    def distance_to_continuous(self, distance):  # waypoint_predictors.py:184-197 (tiny, off the hot path)
        da, db, _, _ = self.to_continuous_consts()
        return distance if self.wypt_cfg.continuous_distance else da + distance * db

    def offset_to_continuous(self, offset):  # :199-209
        _, _, oa, ob = self.to_continuous_consts()
        return offset if self.wypt_cfg.continuous_offset else oa + offset * ob

    @property
    def num_recurrent_layers(self):
        return self.main_state_encoder.num_recurrent_layers + self.visual_rnn.num_recurrent_layers

    @property
    def is_blind(self):
        return False

    @property
    def output_size(self):
        return self._hidden_size

    def engine(self) -> "WaypointEngine":
        eng = self.__dict__.get("_engine")
        if eng is None or eng.precision != self.precision:
            eng = WaypointEngine(self, self.precision)
            self.__dict__["_engine"] = eng
        return eng

    def forward(self, observations, rnn_states, prev_actions, masks):
        """The reference's 7-tuple (``waypoint_predictors.py:286-625``) without autograd (rollouts /
        inference); training goes through ``WaypointPolicy.evaluate_actions`` (one autograd node)."""
        out = self.engine().forward(observations, rnn_states, prev_actions, masks, save=False)
        B, P = out["B"], self._num_panos
        dist = CustomFixedCategorical(logits=out["logp"].clone())
        shaped = lambda t: None if t is None else t.view(B, P, -1).squeeze(-1).clone()  # noqa: E731
        return (dist, shaped(out["o1"]), shaped(out["o2"]), shaped(out["d1"]), shaped(out["d2"]), out["x"].clone(),
                out["h_out"].clone())


def resnet_encoders_single_frame(observation_space):
    """``common/utils.py:32-42`` ``single_frame_box_shape``: strip the leading pano dimension for the
    depth encoder (``resnet_encoders.py:34-36``)."""
    from .compat import Box, Dict

    spaces = {}
    for k, box in observation_space.spaces.items():
        if len(box.shape) >= 4:
            spaces[k] = Box(low=float(np.min(box.low)), high=float(np.max(box.high)), shape=box.shape[1:],
                            dtype=box.dtype)
        else:
            spaces[k] = box
    return Dict(spaces)


# ======================================================================================
class WaypointEngine(HeadEngine):
    """Forward / hand-written backward launch plan bound to one ``WaypointPredictionNet``."""

    def __init__(self, net, precision="bf16"):
        super().__init__(net, precision)
        cfg, w = net.model_config, net.wypt_cfg
        ie = net.instruction_encoder
        self.Hs, self.dirs, self.E = ie.config.hidden_size, (2 if ie.config.bidirectional else 1), ie.config.embedding_size
        self.Di = self.Hs * self.dirs
        self.P = net._num_panos
        self.S = 16
        self.rgb_out, self.dep_out = cfg.RGB_ENCODER.output_size, cfg.DEPTH_ENCODER.output_size
        self.c_rgb = net.rgb_encoder.resnet_layer_size
        self.Crgb = net.rgb_encoder.output_shape[0]
        self.c_dep = net.depth_encoder.visual_encoder.output_shape[0]
        self.Cdep = net.depth_encoder.output_shape[0]
        self.feat = self.rgb_out + self.dep_out + ANGLE_FEATURE_SIZE
        self.cat = self.H + self.feat                       # per-frame head input [frame feats | x]
        self.state_in = self.rgb_out + PREV_ACTION_DIM + self.rgb_out + self.dep_out
        self.xcat = self.Di + self.feat + self.H + PREV_ACTION_DIM
        self.col_pano, self.col_hist, self.col_prev = self.Di, self.Di + self.feat, self.Di + self.feat + self.H
        self.scale = 1.0 / math.sqrt(self.H // 2)
        self.pano_scale = 1.0 / math.sqrt(PANO_ATTN_KEY_DIM)
        self.head_consts = _host_floats([
            w.max_distance_prediction - w.min_distance_prediction, w.min_distance_prediction,
            w.max_distance_var - w.min_distance_var, w.min_distance_var, net.offset_scale,
            1.0 / w.offset_temperature, w.max_offset_var - w.min_offset_var, w.min_offset_var])
        self.nd = 0 if w.continuous_distance else int(w.discrete_distances)   # classes of a discrete head (0 = continuous)
        self.no = 0 if w.continuous_offset else int(w.discrete_offsets)
        self.to_cont = net.to_continuous_consts()
        self.act_consts = _host_floats([w.min_distance_prediction, w.max_distance_prediction, math.pi / self.P,
                                        float(bool(w.predict_distance)), float(bool(w.predict_offset)), self.nd, self.no,
                                        *self.to_cont])

    def heads(self):
        """(raw column, parameter prefix, Linear) of the scalar heads feeding ``wp_head_raw`` (continuous components)."""
        n, out = self.net, []
        if self.nd == 0:
            out += [(0, "distance_linear.0", n.distance_linear[0]), (1, "distance_var_linear.0", n.distance_var_linear[0])]
        if self.no == 0:
            out += [(2, "offset_linear.0", n.offset_linear[0]), (3, "offset_var_linear.0", n.offset_var_linear[0])]
        return out

    def class_heads(self):
        """(buffer name, parameter prefix, Linear, classes) of the categorical heads (discrete components)."""
        n, out = self.net, []
        if self.nd:
            out.append(("wp_dlog", "distance_linear", n.distance_linear, self.nd))
        if self.no:
            out.append(("wp_olog", "offset_linear", n.offset_linear, self.no))
        return out

    def export_visual_features(self) -> Dict[str, torch.Tensor]:
        """The frozen trunks' outputs of the LAST forward, per row: ``pano_rgb_features`` (B, P+1, c_rgb) — the RGB
        trunk's pooled vector of the 12 pano frames and of the (mask-reset) history frame — and ``pano_depth_features``
        (B, P+1, 16, c_dep).  Rows that carry these keys skip both ResNets in ``forward`` (strided copies, no arithmetic).
        Call right after ``act`` / ``get_value``: the buffers are reused by the next forward."""
        B, P, S = self._visual_rows, self.P, self.S
        F = B * (P + 1)
        rgb_e, dep_e = self.buf("wp_rgb_e", (F, S, self.Crgb)), self.buf("wp_dep_e", (F, S, self.Cdep))
        fr = torch.empty(B, P + 1, self.c_rgb, device=rgb_e.device)
        fd = torch.empty(B, P + 1, S, self.c_dep, device=rgb_e.device)
        fr[:, :P].copy_(rgb_e[:B * P, 0, :self.c_rgb].reshape(B, P, self.c_rgb))
        fr[:, P].copy_(rgb_e[B * P:, 0, :self.c_rgb])
        fd[:, :P].copy_(dep_e[:B * P, :, :self.c_dep].reshape(B, P, S, self.c_dep))
        fd[:, P].copy_(dep_e[B * P:, :, :self.c_dep])
        return {VISUAL_CACHE_KEYS[0]: fr, VISUAL_CACHE_KEYS[1]: fd}

    # ------------------------------------------------------------------ forward
    def forward(self, observations, rnn_states, prev_actions: Dict[str, torch.Tensor], masks, save: bool, critic=None):
        net = self.net
        # frozen-trunk feature cache (both conv trunks are frozen in every waypoint recipe, config/default.py:242,253):
        # rows that carry the features a previous forward exported (export_visual_features, kept in the rollout storage
        # by the trainer loop) skip frame stacking and both ResNets — the PPO update then re-reads 213 KB per env step
        # instead of pushing 13 frames through two ResNet-50s again
        cached = VISUAL_CACHE_KEYS[0] in observations and VISUAL_CACHE_KEYS[1] in observations
        for k in (VISUAL_CACHE_KEYS if cached else ("rgb", "depth", "rgb_history", "depth_history")) + ("angle_features",):
            assert k in observations, k
            lib.require_cuda(observations[k], k)
        lib.require_cuda(rnn_states, "rnn_states")
        dev = rnn_states.device
        self.ensure_pool(dev)
        st = lib.stream()
        H, P, S, hk = self.H, self.P, self.S, self.H // 2
        N = rnn_states.shape[0]
        B = observations["angle_features"].shape[0]
        if cached:
            assert tuple(observations[VISUAL_CACHE_KEYS[0]].shape) == (B, P + 1, self.c_rgb)
            assert tuple(observations[VISUAL_CACHE_KEYS[1]].shape) == (B, P + 1, S, self.c_dep)
        else:
            assert observations["rgb"].shape[0] == B
            assert observations["rgb"].shape[1] == P and observations["depth"].shape[1] == P
        assert B % N == 0, "rows must be T*N (time-major)"
        T = B // N
        F = B * (P + 1)
        rnn_states = rnn_states.contiguous().float()
        masks_f = self.buf("masks_f", (B,))
        masks_f.copy_(masks.reshape(-1))

        with Branch("instr") as br_instr:
            instr, _, lengths, L = self._encode_instruction(observations, B, save)

        # ---- 12 pano frames + masked history frame -> one trunk batch per modality (:327-366)
        rgb_e = self.buf("wp_rgb_e", (F, S, self.Crgb))
        dep_e = self.buf("wp_dep_e", (F, S, self.Cdep))
        self._visual_rows = B
        if cached:
            # frame order of the trunk batch: [B*P pano frames (row-major) | B history frames]; strided copies only
            fr = observations[VISUAL_CACHE_KEYS[0]].float()
            fd = observations[VISUAL_CACHE_KEYS[1]].float()
            rgb_e[:B * P, 0, :self.c_rgb].copy_(fr[:, :P].reshape(B * P, self.c_rgb))
            rgb_e[B * P:, 0, :self.c_rgb].copy_(fr[:, P])
            dep_e[:B * P, :, :self.c_dep].copy_(fd[:, :P].reshape(B * P, S, self.c_dep))
            dep_e[B * P:, :, :self.c_dep].copy_(fd[:, P])
            C.vlnce_spatial_embed_fill(ptr(net.depth_encoder.spatial_embeddings.weight), ptr(dep_e), S * self.Cdep,
                                       self.Cdep, self.c_dep, F, S, 64, st)
            br_depth = None
        else:
            rgb, rgb_h = observations["rgb"], observations["rgb_history"]
            if rgb.dtype != rgb_h.dtype or rgb.dtype not in (torch.uint8, torch.float32):
                rgb, rgb_h = rgb.float(), rgb_h.float()
            rgb, rgb_h = rgb.contiguous(), rgb_h.contiguous()
            dep = observations["depth"].contiguous().float()
            dep_h = observations["depth_history"].contiguous().float()
            fr_rgb = self.buf("wp_frames_rgb", (F,) + tuple(rgb.shape[2:]), rgb.dtype)
            fr_dep = self.buf("wp_frames_dep", (F,) + tuple(dep.shape[2:]), torch.float32)
            C.vlnce_wp_stack_frames(ptr(rgb), ptr(rgb_h), ptr(masks_f), ptr(fr_rgb), B, P,
                                    rgb[0, 0].numel() * rgb.element_size(), st)
            C.vlnce_wp_stack_frames(ptr(dep), ptr(dep_h), ptr(masks_f), ptr(fr_dep), B, P, dep[0, 0].numel() * 4, st)
            with Branch("depth") as br_depth:
                net.depth_encoder.features_into({"depth": fr_dep}, dep_e, self.Cdep, self.precision)
                C.vlnce_spatial_embed_fill(ptr(net.depth_encoder.spatial_embeddings.weight), ptr(dep_e), S * self.Cdep,
                                           self.Cdep, self.c_dep, F, S, 64, lib.stream())
            # W3: the trunk's global pool lands in position 0 of every frame, then is shown at all 16 positions
            net.rgb_encoder.features_into({"rgb": fr_rgb}, rgb_e, S * self.Crgb, self.precision)
        C.vlnce_broadcast_positions(ptr(rgb_e), S * self.Crgb, self.Crgb, F, S, self.c_rgb, st)
        C.vlnce_spatial_embed_fill(ptr(net.rgb_encoder.spatial_embeddings.weight), ptr(rgb_e), S * self.Crgb,
                                   self.Crgb, self.c_rgb, F, S, 64, st)
        if br_depth is not None:
            br_depth.join()
        cfg = net.model_config  # modality ablations (waypoint_predictors.py:400-407): embeddings AND history * 0
        if cfg.get("ablate_rgb", False):
            rgb_e.zero_()
        if cfg.get("ablate_depth", False):
            dep_e.zero_()
        hist_rgb = off(rgb_e, B * P * S * self.Crgb)   # history frames: rows [B*P, B*P+B)
        hist_dep = off(dep_e, B * P * S * self.Cdep)

        # ---- prev-action features into both concat slices (:372-384)
        state_in = self.buf("wp_state_in", (B, self.state_in))
        xcat = self.buf("wp_xcat", (B, self.xcat))
        pano = prev_actions["pano"].reshape(-1)
        pano = (pano if pano.dtype == torch.int64 else pano.long()).contiguous()
        p_off = prev_actions["offset"].reshape(-1).float().contiguous()
        p_dist = prev_actions["distance"].reshape(-1).float().contiguous()
        C.vlnce_wp_prev_action(ptr(pano), ptr(p_off), ptr(p_dist), ptr(masks_f), off(state_in, self.rgb_out),
                               self.state_in, off(xcat, self.col_prev), self.xcat, B, P, float(self.to_cont[2]),
                               float(self.to_cont[3]), float(self.to_cont[0]), float(self.to_cont[1]), st)

        # ---- visual-history GRU inputs (:396-405)
        rgb_pm = self.buf("wp_rgb_pm", (B, self.c_rgb))  # mean over the 12 pano frames (Linear and means commute)
        C.vlnce_mean_positions(ptr(rgb_e), P * S * self.Crgb, S * self.Crgb, ptr(rgb_pm), self.c_rgb, B, P, self.c_rgb, st)
        lin = net.rgb_pool_linear
        linear(ptr(rgb_pm), self.c_rgb, ptr(lin.weight), ptr(state_in), self.state_in, ptr(lin.bias), B, self.rgb_out,
               self.c_rgb)
        rgb_hm = self.buf("wp_rgb_hmean", (B, self.Crgb))
        C.vlnce_mean_positions(hist_rgb, S * self.Crgb, self.Crgb, ptr(rgb_hm), self.Crgb, B, S, self.Crgb, st)
        lin = net.rgb_hist_linear[2]
        c0 = self.rgb_out + PREV_ACTION_DIM
        linear(ptr(rgb_hm), self.Crgb, ptr(lin.weight), off(state_in, c0), self.state_in, ptr(lin.bias), B,
               self.rgb_out, self.Crgb, act=1)
        dep_hf = self.buf("wp_dep_hflat", (B, self.Cdep * S))
        C.vlnce_positions_to_nchw(hist_dep, S * self.Cdep, self.Cdep, ptr(dep_hf), self.Cdep * S, B, S, self.Cdep, st)
        lin = net.depth_hist_linear[1]
        linear(ptr(dep_hf), self.Cdep * S, ptr(lin.weight), off(state_in, c0 + self.rgb_out), self.state_in,
               ptr(lin.bias), B, self.dep_out, self.Cdep * S, act=1)

        # ---- visual-history GRU -> xcat[:, col_hist:col_hist+H] (:407-413)
        s1, Lr = self.slots(net.visual_rnn.rnn), net.num_recurrent_layers  # (N, Lr, H): visual_rnn | main encoder
        h_out = self.buf("wp_rnn_states_out", (N, Lr, H))
        self._gru_forward("wp_vgru", net.visual_rnn.rnn, ptr(state_in), self.state_in, self.state_in, ptr(rnn_states),
                          Lr * H, masks_f, off(xcat, self.col_hist), self.xcat, ptr(h_out), Lr * H, T, N, save)

        # ---- instruction attention (:419-424), quirk W1 -> xcat[:, 0:Di]
        br_instr.join()
        if cfg.get("ablate_instruction", False):
            instr.zero_()
        q1 = self.buf("wp_q1", (B, hk))
        lin = net.inst_attn_q[0]
        linear(off(xcat, self.col_hist), self.xcat, ptr(lin.weight), ptr(q1), hk, ptr(lin.bias), B, hk, H, act=1)
        ia_save = self.buf("wp_ia_save", (B, 2))
        C.vlnce_wp_inst_attn_fwd(ptr(q1), hk, ptr(net.inst_attn_k.bias), ptr(instr), ptr(lengths), self.scale,
                                 ptr(xcat), self.xcat, ptr(ia_save), B, L, self.Di, hk, st)

        # ---- spatial attention inside each pano frame (:431-489), quirk W2
        q2 = self.buf("wp_q2", (B, hk))
        lin = net.text_q_linear
        linear(ptr(xcat), self.xcat, ptr(lin.weight), ptr(q2), hk, ptr(lin.bias), B, hk, self.Di)
        qs = self.buf("wp_qs", (B * P, hk))
        C.vlnce_wp_scramble_query(ptr(q2), ptr(qs), B * hk, P, st)
        catted = self.buf("wp_catted", (B * P, self.cat))
        for name, E_, Cc, kv, v_out, col in (("rgb", rgb_e, self.Crgb, net.rgb_kv_spatial, self.rgb_out, 0),
                                             ("dep", dep_e, self.Cdep, net.depth_kv_spatial, self.dep_out, self.rgb_out)):
            qv = self.buf("wp_q_" + name, (B * P, Cc))
            linear_dx(ptr(qs), hk, ptr(kv.weight), ptr(qv), Cc, B * P, hk, Cc)          # W_k^T q
            probs = self.buf("wp_p_" + name, (B * P, S)) if save else None
            pooled = self.buf("wp_pooled_" + name, (B * P, Cc))
            C.vlnce_attn_pool_fwd(ptr(qv), Cc, ptr(E_), S * Cc, Cc, None, self.scale, ptr(probs), ptr(pooled), Cc,
                                  B * P, S, Cc, st)
            linear(ptr(pooled), Cc, off(kv.weight, hk * Cc), off(catted, col), self.cat, off(kv.bias, hk), B * P,
                   v_out, Cc)                                                            # W_v pooled + b_v
        ang = observations["angle_features"].contiguous().float()
        ones = self.buf("wp_ones", (B * P,))
        ones.fill_(1.0)
        C.vlnce_mask_rows(ptr(ang), ANGLE_FEATURE_SIZE, ptr(ones), off(catted, self.rgb_out + self.dep_out), self.cat,
                          B * P, ANGLE_FEATURE_SIZE, st)

        # ---- attention over the pano frames (:495-509) -> LayerNorm -> xcat[:, col_pano:col_pano+feat]
        pa = net.pano_attn
        D = self.feat
        qp = self.buf("wp_qp", (B, PANO_ATTN_KEY_DIM))
        linear(ptr(xcat), self.xcat, ptr(pa.q_linear.weight), ptr(qp), PANO_ATTN_KEY_DIM, None, B, PANO_ATTN_KEY_DIM,
               self.Di)
        qpk = self.buf("wp_qpk", (B, D))
        linear_dx(ptr(qp), PANO_ATTN_KEY_DIM, ptr(pa.k_linear.weight), ptr(qpk), D, B, PANO_ATTN_KEY_DIM, D)
        p_pano = self.buf("wp_p_pano", (B, P)) if save else None
        pooled_p = self.buf("wp_pooled_pano", (B, D))
        C.vlnce_attn_pool_fwd(ptr(qpk), D, ptr(catted), P * self.cat, self.cat, None, self.pano_scale, ptr(p_pano),
                              ptr(pooled_p), D, B, P, D, st)
        pv = self.buf("wp_pano_v", (B, PANO_ATTN_KEY_DIM))
        linear(ptr(pooled_p), D, ptr(pa.v_linear.weight), ptr(pv), PANO_ATTN_KEY_DIM, None, B, PANO_ATTN_KEY_DIM, D)
        fo = self.buf("wp_pano_fo", (B, D))
        linear(ptr(pv), PANO_ATTN_KEY_DIM, ptr(pa.final_linear.weight), ptr(fo), D, None, B, D, PANO_ATTN_KEY_DIM)
        ln_save = self.buf("wp_ln_save", (B, 2))
        C.vlnce_layernorm_fwd(ptr(fo), D, ptr(pa.layer_norm.weight), ptr(pa.layer_norm.bias), off(xcat, self.col_pano),
                              self.xcat, ptr(ln_save), B, D, float(pa.layer_norm.eps), st)

        # ---- main recurrence (:515-540)
        x2 = self.buf("wp_x2", (B, H))
        lin = net.main_state_compress[0]
        linear(ptr(xcat), self.xcat, ptr(lin.weight), ptr(x2), H, ptr(lin.bias), B, H, self.xcat, act=1)
        x_out = self.buf("wp_x_out", (B, H))
        self._gru_forward("wp_mgru", net.main_state_encoder.rnn, ptr(x2), H, H, off(rnn_states, s1 * H), Lr * H,
                          masks_f, ptr(x_out), H, off(h_out, s1 * H), Lr * H, T, N, save)

        # ---- heads (:546-615)
        xs = self.buf("wp_xs", (B, D))
        lin = net.compress_x_linear[0]
        linear(ptr(x_out), H, ptr(lin.weight), ptr(xs), D, ptr(lin.bias), B, D, H, act=1)
        C.vlnce_broadcast_rows(ptr(x_out), H, off(catted, D), P * self.cat, self.cat, B, P, H, st)
        logp = self.buf("wp_logits", (B, P + 1))
        C.vlnce_wp_dot_frames_fwd(ptr(catted), P * self.cat, self.cat, ptr(xs), D, ptr(logp), P + 1, B, P, D, st)
        lin = net.stop_linear
        linear(ptr(x_out), H, ptr(lin.weight), off(logp, P), P + 1, ptr(lin.bias), B, 1, H)
        C.vlnce_log_softmax_rows(ptr(logp), B, P + 1, st)
        raw = self.buf("wp_head_raw", (B * P, 4))
        if self.nd or self.no:
            raw.zero_()  # columns of a discrete component stay unused
        for h, _, lin in self.heads():
            linear(ptr(catted), self.cat, ptr(lin.weight), off(raw, h), 4, ptr(lin.bias), B * P, 1, self.cat)
        d1, d2 = self.buf("wp_d1", (B * P,)), self.buf("wp_d2", (B * P,))
        o1, o2 = self.buf("wp_o1", (B * P,)), self.buf("wp_o2", (B * P,))
        C.vlnce_wp_head_act_fwd(ptr(raw), ptr(d1), ptr(d2), ptr(o1), ptr(o2), B * P, ctypes.addressof(self.head_consts), st)
        for name, _, lin, ncls in self.class_heads():  # categorical heads: raw logits (B*P, classes)
            lg = self.buf(name, (B * P, ncls))
            linear(ptr(catted), self.cat, ptr(lin.weight), ptr(lg), ncls, ptr(lin.bias), B * P, ncls, self.cat)
        if self.nd:
            d1, d2 = self.buf("wp_dlog", (B * P, self.nd)), None
        if self.no:
            o1, o2 = self.buf("wp_olog", (B * P, self.no)), None
        values = None
        if critic is not None:
            values = self.buf("wp_values", (B, 1))
            linear(ptr(x_out), H, ptr(critic.fc.weight), ptr(values), 1, ptr(critic.fc.bias), B, 1, H)
        self.ctx.update(R=B, B=B, N=N, T=T, L=L, lengths=lengths, saved=save)
        return dict(logp=logp, o1=o1, o2=o2, d1=d1, d2=d2, x=x_out, h_out=h_out, values=values, B=B, N=N, T=T)

    # ------------------------------------------------------------------ backward
    def backward(self, g: Dict[str, torch.Tensor], critic, d_raw, d_o1, d_o2, d_d1, d_d2, d_values):
        """Hand-written backward of the last ``forward(save=True)``.  ``g``: parameter name -> ZEROED gradient
        view.  Inputs: gradients w.r.t. the raw pano/stop logits (B,P+1), the four head outputs (B*P) and the
        critic values (B,1).  Both conv trunks are frozen (``config/default.py:242,253``)."""
        from .backward import _Leaves, _colsum, _gru_backward, instruction_backward

        net, cx, b = self.net, self.ctx, self.buf
        assert cx.get("saved"), "backward needs the forward to have run with save=True"
        B, N, T, L = cx["B"], cx["N"], cx["T"], cx["L"]
        H, P, S, hk, D = self.H, self.P, self.S, self.H // 2, self.feat
        st = lib.stream()
        leaves = _Leaves()
        masks_f = b("masks_f", (B,))
        x_out, x2, xcat, state_in = b("wp_x_out", (B, H)), b("wp_x2", (B, H)), b("wp_xcat", (B, self.xcat)), b("wp_state_in", (B, self.state_in))
        catted, xs = b("wp_catted", (B * P, self.cat)), b("wp_xs", (B, D))
        dx_out = b("wp_d_x_out", (B, H))

        # 1. critic
        with leaves():
            linear_dw(ptr(d_values), 1, ptr(x_out), H, ptr(g["critic.fc.weight"]), H, B, 1, H)
            _colsum(ptr(d_values), 1, ptr(g["critic.fc.bias"]), B, 1)
        linear_dx(ptr(d_values), 1, ptr(critic.fc.weight), ptr(dx_out), H, B, 1, H)
        # 2. the four heads: activations, then Linear(cat -> 1) each
        raw, graw = b("wp_head_raw", (B * P, 4)), b("wp_d_head_raw", (B * P, 4))
        zeros = b("wp_zeros", (B * P,))
        if self.nd or self.no:
            zeros.zero_()
        C.vlnce_wp_head_act_bwd(ptr(raw), ptr(zeros if self.nd else d_d1), ptr(zeros if self.nd else d_d2),
                                ptr(zeros if self.no else d_o1), ptr(zeros if self.no else d_o2), ptr(graw), B * P,
                                ctypes.addressof(self.head_consts), st)
        d_catted = b("wp_d_catted", (B * P, self.cat))
        first = True
        for h, nm, lin in self.heads():
            with leaves():
                linear_dw(off(graw, h), 4, ptr(catted), self.cat, ptr(g[f"net.{nm}.weight"]), self.cat, B * P, 1, self.cat)
                _colsum(off(graw, h), 4, ptr(g[f"net.{nm}.bias"]), B * P, 1)
            linear_dx(off(graw, h), 4, ptr(lin.weight), ptr(d_catted), self.cat, B * P, 1, self.cat,
                      beta=0.0 if first else 1.0)
            first = False
        for (_, nm, lin, ncls), dlg in zip(self.class_heads(), [t for t, n_ in ((d_d1, self.nd), (d_o1, self.no)) if n_]):
            with leaves():
                linear_dw(ptr(dlg), ncls, ptr(catted), self.cat, ptr(g[f"net.{nm}.weight"]), self.cat, B * P, ncls, self.cat)
                _colsum(ptr(dlg), ncls, ptr(g[f"net.{nm}.bias"]), B * P, ncls)
            linear_dx(ptr(dlg), ncls, ptr(lin.weight), ptr(d_catted), self.cat, B * P, ncls, self.cat,
                      beta=0.0 if first else 1.0)
            first = False
        # 3. pano/stop logits
        lin = net.stop_linear
        with leaves():
            linear_dw(off(d_raw, P), P + 1, ptr(x_out), H, ptr(g["net.stop_linear.weight"]), H, B, 1, H)
            _colsum(off(d_raw, P), P + 1, ptr(g["net.stop_linear.bias"]), B, 1)
        linear_dx(off(d_raw, P), P + 1, ptr(lin.weight), ptr(dx_out), H, B, 1, H, beta=1.0)
        d_xs = b("wp_d_xs", (B, D))
        C.vlnce_wp_dot_frames_bwd(ptr(catted), P * self.cat, self.cat, ptr(xs), D, ptr(d_raw), P + 1, ptr(d_catted),
                                  P * self.cat, self.cat, 1, ptr(d_xs), D, B, P, D, st)
        # 4. x broadcast into the per-frame concat
        C.vlnce_sum_positions(off(d_catted, D), P * self.cat, self.cat, ptr(dx_out), H, 1, B, P, H, st)
        # 5. compress_x_linear = ReLU(Linear(x))
        C.vlnce_relu_bwd(ptr(d_xs), D, ptr(xs), D, B, D, st)
        lin = net.compress_x_linear[0]
        with leaves():
            linear_dw(ptr(d_xs), D, ptr(x_out), H, ptr(g["net.compress_x_linear.0.weight"]), H, B, D, H)
            _colsum(ptr(d_xs), D, ptr(g["net.compress_x_linear.0.bias"]), B, D)
        linear_dx(ptr(d_xs), D, ptr(lin.weight), ptr(dx_out), H, B, D, H, beta=1.0)
        # 6. main GRU, 7. main_state_compress
        d_x2 = b("wp_d_x2", (B, H))
        _gru_backward(self, "wp_mgru", net.main_state_encoder.rnn, "net.main_state_encoder.rnn.", g, ptr(dx_out), H,
                      ptr(x2), H, H, masks_f, T, N, ptr(d_x2), H, leaves)
        C.vlnce_relu_bwd(ptr(d_x2), H, ptr(x2), H, B, H, st)
        lin = net.main_state_compress[0]
        with leaves():
            linear_dw(ptr(d_x2), H, ptr(xcat), self.xcat, ptr(g["net.main_state_compress.0.weight"]), self.xcat, B, H, self.xcat)
            _colsum(ptr(d_x2), H, ptr(g["net.main_state_compress.0.bias"]), B, H)
        dxcat = b("wp_d_xcat", (B, self.xcat))
        linear_dx(ptr(d_x2), H, ptr(lin.weight), ptr(dxcat), self.xcat, B, H, self.xcat)
        # 8. pano attention: LayerNorm <- final_linear <- v_linear <- pooling <- (k_linear, q_linear)
        pa = net.pano_attn
        KD = PANO_ATTN_KEY_DIM
        fo, pv, pooled_p = b("wp_pano_fo", (B, D)), b("wp_pano_v", (B, KD)), b("wp_pooled_pano", (B, D))
        qp, qpk, p_pano = b("wp_qp", (B, KD)), b("wp_qpk", (B, D)), b("wp_p_pano", (B, P))
        d_fo, prod = b("wp_d_fo", (B, D)), b("wp_ln_prod", (B, D))
        C.vlnce_layernorm_bwd(off(dxcat, self.col_pano), self.xcat, ptr(fo), D, ptr(pa.layer_norm.weight),
                              ptr(b("wp_ln_save", (B, 2))), ptr(d_fo), D, ptr(prod), D, B, D, st)
        with leaves():
            _colsum(ptr(prod), D, ptr(g["net.pano_attn.layer_norm.weight"]), B, D)
            _colsum(off(dxcat, self.col_pano), self.xcat, ptr(g["net.pano_attn.layer_norm.bias"]), B, D)
            linear_dw(ptr(d_fo), D, ptr(pv), KD, ptr(g["net.pano_attn.final_linear.weight"]), KD, B, D, KD)
        d_pv = b("wp_d_pano_v", (B, KD))
        linear_dx(ptr(d_fo), D, ptr(pa.final_linear.weight), ptr(d_pv), KD, B, D, KD)
        with leaves():
            linear_dw(ptr(d_pv), KD, ptr(pooled_p), D, ptr(g["net.pano_attn.v_linear.weight"]), D, B, KD, D)
        d_pooled_p = b("wp_d_pooled_pano", (B, D))
        linear_dx(ptr(d_pv), KD, ptr(pa.v_linear.weight), ptr(d_pooled_p), D, B, KD, D)
        d_qpk = b("wp_d_qpk", (B, D))
        C.vlnce_attn_pool_bwd(ptr(qpk), D, ptr(catted), P * self.cat, self.cat, ptr(p_pano), self.pano_scale,
                              ptr(d_pooled_p), D, ptr(d_qpk), D, ptr(d_catted), P * self.cat, self.cat, 1, B, P, D, st)
        with leaves():
            linear_dw(ptr(qp), KD, ptr(d_qpk), D, ptr(g["net.pano_attn.k_linear.weight"]), D, B, KD, D)
        d_qp = b("wp_d_qp", (B, KD))
        linear(ptr(d_qpk), D, ptr(pa.k_linear.weight), ptr(d_qp), KD, None, B, KD, D)
        with leaves():
            linear_dw(ptr(d_qp), KD, ptr(xcat), self.xcat, ptr(g["net.pano_attn.q_linear.weight"]), self.Di, B, KD, self.Di)
        linear_dx(ptr(d_qp), KD, ptr(pa.q_linear.weight), ptr(dxcat), self.xcat, B, KD, self.Di, beta=1.0)
        # 9. spatial attentions (frame feats = d_catted[:, 0:feat])
        rgb_e = b("wp_rgb_e", (B * (P + 1), S, self.Crgb))
        dep_e = b("wp_dep_e", (B * (P + 1), S, self.Cdep))
        qs, d_qs = b("wp_qs", (B * P, hk)), b("wp_d_qs", (B * P, hk))
        dE = {}
        first = True
        for name, pname, E_, Cc, kv, v_out, col in (
                ("rgb", "rgb_kv_spatial", rgb_e, self.Crgb, net.rgb_kv_spatial, self.rgb_out, 0),
                ("dep", "depth_kv_spatial", dep_e, self.Cdep, net.depth_kv_spatial, self.dep_out, self.rgb_out)):
            pooled, qv, probs = b("wp_pooled_" + name, (B * P, Cc)), b("wp_q_" + name, (B * P, Cc)), b("wp_p_" + name, (B * P, S))
            gw, gb = g[f"net.{pname}.weight"], g[f"net.{pname}.bias"]
            d_att = off(d_catted, col)
            with leaves():
                linear_dw(d_att, self.cat, ptr(pooled), Cc, off(gw, hk * Cc), Cc, B * P, v_out, Cc)
                _colsum(d_att, self.cat, off(gb, hk), B * P, v_out)
            d_pooled = b("wp_d_pooled_" + name, (B * P, Cc))
            linear_dx(d_att, self.cat, off(kv.weight, hk * Cc), ptr(d_pooled), Cc, B * P, v_out, Cc)
            dq = b("wp_d_q_" + name, (B * P, Cc))
            dEn = b("wp_d_emb_" + name, (B * P, S, Cc))
            C.vlnce_attn_pool_bwd(ptr(qv), Cc, ptr(E_), S * Cc, Cc, ptr(probs), self.scale, ptr(d_pooled), Cc, ptr(dq),
                                  Cc, ptr(dEn), S * Cc, Cc, 0, B * P, S, Cc, st)
            dE[name] = dEn
            with leaves():
                linear_dw(ptr(qs), hk, ptr(dq), Cc, ptr(gw), Cc, B * P, hk, Cc)   # dW_k = qs^T dq' (key bias: exactly 0)
            linear(ptr(dq), Cc, ptr(kv.weight), ptr(d_qs), hk, None, B * P, hk, Cc, beta=0.0 if first else 1.0)
            first = False
        # 10. W2 adjoint, text_q_linear
        d_q2 = b("wp_d_q2", (B, hk))
        C.vlnce_wp_scramble_query_bwd(ptr(d_qs), ptr(d_q2), B * hk, P, st)
        lin = net.text_q_linear
        with leaves():
            linear_dw(ptr(d_q2), hk, ptr(xcat), self.xcat, ptr(g["net.text_q_linear.weight"]), self.Di, B, hk, self.Di)
            _colsum(ptr(d_q2), hk, ptr(g["net.text_q_linear.bias"]), B, hk)
        linear_dx(ptr(d_q2), hk, ptr(lin.weight), ptr(dxcat), self.xcat, B, hk, self.Di, beta=1.0)
        # 11. instruction attention (W1): d_instr, the key-bias path, inst_attn_q
        q1 = b("wp_q1", (B, hk))
        dcs, d_instr = b("wp_dcs", (B, 1)), b("d_instr", (B, L, self.Di))
        C.vlnce_wp_inst_attn_bwd(ptr(b("wp_ia_save", (B, 2))), ptr(xcat), self.xcat, ptr(dxcat), self.xcat,
                                 ptr(cx["lengths"]), self.scale, ptr(dcs), ptr(d_instr), B, L, self.Di, st)
        with leaves():
            linear_dw(ptr(dcs), 1, ptr(q1), hk, ptr(g["net.inst_attn_k.bias"]), hk, B, 1, hk)
        d_q1 = b("wp_d_q1", (B, hk))
        linear_dx(ptr(dcs), 1, ptr(net.inst_attn_k.bias), ptr(d_q1), hk, B, 1, hk)
        C.vlnce_relu_bwd(ptr(d_q1), hk, ptr(q1), hk, B, hk, st)
        lin = net.inst_attn_q[0]
        with leaves():
            linear_dw(ptr(d_q1), hk, off(xcat, self.col_hist), self.xcat, ptr(g["net.inst_attn_q.0.weight"]), H, B, hk, H)
            _colsum(ptr(d_q1), hk, ptr(g["net.inst_attn_q.0.bias"]), B, hk)
        linear_dx(ptr(d_q1), hk, ptr(lin.weight), off(dxcat, self.col_hist), self.xcat, B, hk, H, beta=1.0)
        cfg = net.model_config
        with Branch("lstm_bwd") as br_lstm:
            if not cfg.get("ablate_instruction", False):  # ``instruction_embedding *= 0``: nothing flows upstream
                instruction_backward(self, g, d_instr)
        # 12. visual-history GRU
        d_state_in = b("wp_d_state_in", (B, self.state_in))
        _gru_backward(self, "wp_vgru", net.visual_rnn.rnn, "net.visual_rnn.rnn.", g, off(dxcat, self.col_hist), self.xcat,
                      ptr(state_in), self.state_in, self.state_in, masks_f, T, N, ptr(d_state_in), self.state_in, leaves)
        # 13. its input branches (the trunks below them are frozen)
        c0 = self.rgb_out + PREV_ACTION_DIM
        with leaves():
            linear_dw(ptr(d_state_in), self.state_in, ptr(b("wp_rgb_pm", (B, self.c_rgb))), self.c_rgb,
                      ptr(g["net.rgb_pool_linear.weight"]), self.c_rgb, B, self.rgb_out, self.c_rgb)
            _colsum(ptr(d_state_in), self.state_in, ptr(g["net.rgb_pool_linear.bias"]), B, self.rgb_out)
        C.vlnce_relu_bwd(off(d_state_in, c0), self.state_in, off(state_in, c0), self.state_in, B,
                         self.rgb_out + self.dep_out, st)
        lin = net.rgb_hist_linear[2]
        d_rh = off(d_state_in, c0)
        with leaves():
            linear_dw(d_rh, self.state_in, ptr(b("wp_rgb_hmean", (B, self.Crgb))), self.Crgb,
                      ptr(g["net.rgb_hist_linear.2.weight"]), self.Crgb, B, self.rgb_out, self.Crgb)
            _colsum(d_rh, self.state_in, ptr(g["net.rgb_hist_linear.2.bias"]), B, self.rgb_out)
        d_mean_sp = b("wp_d_mean_sp", (B, 64))
        linear_dx(d_rh, self.state_in, off(lin.weight, self.c_rgb), ptr(d_mean_sp), 64, B, self.rgb_out, 64, ldw=self.Crgb)
        lin = net.depth_hist_linear[1]
        d_dh = off(d_state_in, c0 + self.rgb_out)
        with leaves():
            linear_dw(d_dh, self.state_in, ptr(b("wp_dep_hflat", (B, self.Cdep * S))), self.Cdep * S,
                      ptr(g["net.depth_hist_linear.1.weight"]), self.Cdep * S, B, self.dep_out, self.Cdep * S)
            _colsum(d_dh, self.state_in, ptr(g["net.depth_hist_linear.1.bias"]), B, self.dep_out)
        d_flat_sp = b("wp_d_flat_sp", (B, 64 * S))
        linear_dx(d_dh, self.state_in, off(lin.weight, self.c_dep * S), ptr(d_flat_sp), 64 * S, B, self.dep_out, 64 * S,
                  ldw=self.Cdep * S)
        # 14. spatial-embedding tables: attention source (all pano frames) + the history-frame branches
        g_rgb, g_dep = g["net.rgb_encoder.spatial_embeddings.weight"], g["net.depth_encoder.spatial_embeddings.weight"]
        if not cfg.get("ablate_rgb", False):
            C.vlnce_spatial_embed_grad(ptr(dE["rgb"]), S * self.Crgb, self.Crgb, self.c_rgb, None, 0, None, 0, ptr(g_rgb),
                                       0, B * P, S, 64, st)
            C.vlnce_spatial_embed_grad(None, 0, 0, 0, None, 0, ptr(d_mean_sp), 64, ptr(g_rgb), 1, B, S, 64, st)
        if not cfg.get("ablate_depth", False):
            C.vlnce_spatial_embed_grad(ptr(dE["dep"]), S * self.Cdep, self.Cdep, self.c_dep, None, 0, None, 0, ptr(g_dep),
                                       0, B * P, S, 64, st)
            C.vlnce_spatial_embed_grad(None, 0, 0, 0, ptr(d_flat_sp), 64 * S, None, 0, ptr(g_dep), 1, B, S, 64, st)
        br_lstm.join()
        leaves.join()


def trainable_named(policy) -> Dict[str, torch.nn.Parameter]:
    return {k: p for k, p in policy.named_parameters() if p.requires_grad}


class _EvaluateFunction(torch.autograd.Function):
    """``evaluate_actions`` as ONE autograd node, so the reference's ``WDDPPO.update``
    (``common/ddppo_alg.py:56-131``: loss arithmetic in torch, ``loss.backward()``) runs unchanged on top of it."""

    @staticmethod
    def forward(ctx, policy, observations, rnn_states, prev_actions, masks, actions, names, *params):
        eng = policy.net.engine()
        out = eng.forward(observations, rnn_states, prev_actions, masks, save=True, critic=policy.critic)
        B, P = out["B"], eng.P
        acts = policy._action_rows(actions, B)
        ev = eng.buf("wp_eval_out", (B, 4))
        C.vlnce_wp_ppo_loss(ptr(out["logp"]), ptr(out["o1"]), ptr(out["o2"]), ptr(out["d1"]), ptr(out["d2"]), None,
                            ptr(acts[0]), ptr(acts[1]), ptr(acts[2]), None, None, None, None, B, P,
                            ctypes.addressof(eng.act_consts), None, ptr(ev), None, None, None, None, None, None, None,
                            lib.stream())
        ctx.policy, ctx.names, ctx.acts = policy, names, acts
        ctx.shapes = [p.shape for p in params]
        ctx.token = object()
        eng.ctx["token"] = ctx.token
        h = out["h_out"].clone()
        ctx.mark_non_differentiable(h)
        evc = ev.clone()
        return out["values"].clone(), evc[:, 0:1].clone(), evc[:, 1].clone(), evc[:, 2].clone(), evc[:, 3].clone(), h

    @staticmethod
    def backward(ctx, d_v, d_lp, d_hp, d_ho, d_hd, _d_h):
        policy = ctx.policy
        eng = policy.net.engine()
        if eng.ctx.get("token") is not ctx.token:
            raise lib.VlnceError("backward() after another forward on the same policy: the saved workspace was overwritten")
        B, P = eng.ctx["B"], eng.P
        dev = d_lp.device if d_lp is not None else d_v.device
        sizes = [int(torch.Size(s).numel()) for s in ctx.shapes]
        flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
        g, o = {}, 0
        for name, shp, n in zip(ctx.names, ctx.shapes, sizes):
            g[name] = flat[o:o + n].view(shp)
            o += n
        up = eng.buf("wp_upstream", (B, 4))
        up.zero_()
        for j, d in enumerate((d_lp, d_hp, d_ho, d_hd)):
            if d is not None:
                up[:, j].copy_(d.reshape(B))
        dvals = eng.buf("wp_d_values", (B, 1))
        dvals.zero_()
        if d_v is not None:
            dvals.copy_(d_v.reshape(B, 1))
        policy._backward_from_upstream(eng, g, ctx.acts, up, dvals)
        return (None,) * 7 + tuple(g[n] for n in ctx.names)


@baseline_registry.register_policy
class WaypointPolicy(Policy):
    def __init__(self, observation_space, action_space, model_config):
        nn.Module.__init__(self)
        self.net = WaypointPredictionNet(observation_space=observation_space, model_config=model_config)
        self.dim_actions = 1  # waypoint_policy.py:33 "ignore the action dimension"
        self.action_distribution = CategoricalNet(self.net.output_size, self.dim_actions)  # unused, kept for the state_dict
        self.critic = _CriticHead(self.net.output_size)
        self._config = model_config
        self.wypt_cfg = model_config.WAYPOINT
        self._offset_limit = np.pi / self._config.num_panos
        self._sample_calls = 0
        # stream of the counter-based sampler: torch's seed (the trainer seeds per rank with
        # SEED + world_rank * NUM_ENVIRONMENTS, ddppo_waypoint_trainer.py:330-336) mixed with the rank, so DD-PPO
        # workers draw different exploration noise; ``_sample_calls`` advances per act() and can be restored on requeue
        rank = torch.distributed.get_rank() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
        self.sample_seed = (int(torch.initial_seed()) * 0x9E3779B1 + rank * 0x85EBCA77 + 0x5EED) & 0x7FFFFFFF

    def forward(self, *x):
        raise NotImplementedError

    @classmethod
    def from_config(cls, config, observation_space, action_space):
        config.defrost()
        config.MODEL.num_panos = config.TASK_CONFIG.TASK.PANO_ROTATIONS
        config.freeze()
        return cls(observation_space=observation_space, action_space=action_space, model_config=config.MODEL)

    # ------------------------------------------------------------------ act (waypoint_policy.py:129-247)
    # opt-in: replay the forward of ``act`` (stacking, both trunks, head: ~100 launches) from a CUDA graph keyed by the
    # input signature — the rollout collection calls ``act`` once per environment step with fixed shapes, and at 4 envs
    # per rank the eager launches cost as much host time as the kernels take.  The sampling kernel and the D2H read of
    # the env actions stay outside the graph (their RNG counter is a launch argument).
    graph_act = False

    def _forward_for_act(self, eng, observations, rnn_states, prev_actions, masks):
        if not (self.graph_act and rnn_states.is_cuda) or torch.cuda.is_current_stream_capturing():
            return eng.forward(observations, rnn_states, prev_actions, masks, save=False, critic=self.critic)
        parts = (observations, rnn_states, prev_actions, masks)
        flat = []
        for part in parts:
            flat.extend(part[k] for k in sorted(part)) if isinstance(part, dict) else flat.append(part)
        key = tuple((tuple(t.shape), t.dtype) for t in flat) + (id(eng),)
        from .policy import act_graph_cache, act_graph_fingerprint

        graphs = act_graph_cache(self)
        entry = graphs.get(key)
        if entry is None:
            if len(graphs) >= 8:
                graphs.pop(next(iter(graphs)))
            static = tuple({k: v.clone() for k, v in part.items()} if isinstance(part, dict) else part.clone()
                           for part in parts)
            eng.forward(*static, save=False, critic=self.critic)      # eager pass: plans, tile selection, workspaces
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    out = eng.forward(*static, save=False, critic=self.critic)
            torch.cuda.current_stream().wait_stream(side)
            sflat = []
            for part in static:
                sflat.extend(part[k] for k in sorted(part)) if isinstance(part, dict) else sflat.append(part)
            entry = graphs[key] = (sflat, graph, out)
            self.__dict__["_act_graph_fp"] = act_graph_fingerprint(self)   # (the eager pass may have built the trunks)
        sflat, graph, out = entry
        for dst, src in zip(sflat, flat):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        eng._visual_rows = out["B"]     # (python-side state of forward() that a replay does not re-run)
        return out

    @lib.guarded
    def act(self, observations, rnn_states, prev_actions, masks, deterministic=False):
        with torch.no_grad():
            eng = self.net.engine()
            out = self._forward_for_act(eng, observations, rnn_states, prev_actions, masks)
            B, P = out["B"], eng.P
            dev = rnn_states.device
            pano_stop = torch.empty(B, 1, dtype=torch.int64, device=dev)
            f = eng.buf("wp_act_out", (B, 10))
            self._sample_calls += 1
            C.vlnce_wp_act(ptr(out["logp"]), ptr(out["o1"]), ptr(out["o2"]), ptr(out["d1"]), ptr(out["d2"]),
                           ptr(pano_stop), ptr(f), B, P, int(bool(deterministic)), int(self.sample_seed),
                           int(self._sample_calls) << 32, ctypes.addressof(eng.act_consts), lib.stream())
            f = f.clone()
            host = f[:, 7:10].cpu() if f.is_cuda else f[:, 7:10]  # ONE D2H read for the env action dicts (r, theta, stop)
            actions = []
            for r, theta, stop in host.tolist():
                if stop:
                    actions.append({"action": "STOP"})
                else:
                    actions.append({"action": {"action": "GO_TOWARD_POINT", "action_args": {"r": r, "theta": theta}}})
            col = lambda j: f[:, j:j + 1].clone()  # noqa: E731
            cls = lambda t, n_: t.long() if n_ else t  # noqa: E731  (categorical heads emit int64 class indices)
            var = lambda t, n_: t.reshape(-1) if n_ else t  # noqa: E731  (torch's Categorical.variance is (B,))
            action_elements = {"pano": pano_stop, "offset": cls(col(0), eng.no), "distance": cls(col(1), eng.nd)}
            variances = {"distance": var(col(3), eng.nd), "offset": var(col(4), eng.no)}
            modes = {"offset": cls(col(5), eng.no), "distance": cls(col(6), eng.nd)}
            dist = CustomFixedCategorical(logits=out["logp"].clone())
            return (out["values"].clone(), actions, action_elements, modes, variances, col(2), out["h_out"].clone(), dist)

    def visual_features(self) -> Dict[str, torch.Tensor]:
        """Frozen-trunk features of the rows of the last ``act`` / ``get_value`` call (see
        ``WaypointEngine.export_visual_features``); a trainer loop stores them beside the observations so that
        ``evaluate_actions`` in the PPO update does not run the ResNets a second time."""
        return self.net.engine().export_visual_features()

    @lib.guarded
    def get_value(self, observations, rnn_states, prev_actions, masks):  # :249-263
        with torch.no_grad():
            out = self.net.engine().forward(observations, rnn_states, prev_actions, masks, save=False, critic=self.critic)
            return out["values"].clone()

    # ------------------------------------------------------------------ evaluate_actions (:265-333)
    @staticmethod
    def _action_rows(actions, B):
        return tuple(actions[k].reshape(B).float().contiguous() for k in ("pano", "offset", "distance"))

    @lib.guarded
    def evaluate_actions(self, observations, rnn_states, prev_actions, masks, action_components):
        named = trainable_named(self)
        need_grad = torch.is_grad_enabled() and len(named) > 0
        if need_grad:
            v, lp, hp, ho, hd, h = _EvaluateFunction.apply(self, observations, rnn_states, prev_actions, masks,
                                                           action_components, tuple(named.keys()), *named.values())
        else:
            eng = self.net.engine()
            out = eng.forward(observations, rnn_states, prev_actions, masks, save=False, critic=self.critic)
            B = out["B"]
            acts = self._action_rows(action_components, B)
            ev = eng.buf("wp_eval_out", (B, 4))
            C.vlnce_wp_ppo_loss(ptr(out["logp"]), ptr(out["o1"]), ptr(out["o2"]), ptr(out["d1"]), ptr(out["d2"]), None,
                                ptr(acts[0]), ptr(acts[1]), ptr(acts[2]), None, None, None, None, B, eng.P,
                                ctypes.addressof(eng.act_consts), None, ptr(ev), None, None, None, None, None, None,
                                None, lib.stream())
            evc = ev.clone()
            v, lp, hp, ho, hd, h = (out["values"].clone(), evc[:, 0:1], evc[:, 1], evc[:, 2], evc[:, 3],
                                    out["h_out"].clone())
        # W4 (waypoint_policy.py:318-325): for a categorical head the reference multiplies its (B,1) mask with the
        # (B,) entropy -> a (B,B) matrix mask_i * H_j (no-op squeeze); reproduced with a broadcasting view
        eng = self.net.engine()
        if eng.no or eng.nd:
            go = (action_components["pano"].reshape(-1, 1) != eng.P).to(lp.dtype)
            if eng.no:
                ho = (go * float(bool(self.wypt_cfg.predict_offset))) * ho.view(1, -1)
            if eng.nd:
                hd = (go * float(bool(self.wypt_cfg.predict_distance))) * hd.view(1, -1)
        return v, lp, {"pano": hp, "offset": ho, "distance": hd}, h

    def _backward_from_upstream(self, eng, g, acts, upstream, d_values):
        """(B,4) upstream gradients of {log-prob, entropies} -> net-output gradients -> the net's backward."""
        B, P = eng.ctx["B"], eng.P
        bufs = self._grad_buffers(eng, B, P)
        o1, o2, d1, d2 = self._head_outputs(eng, B, P)
        C.vlnce_wp_eval_bwd(ptr(eng.buf("wp_logits", (B, P + 1))), ptr(o1), ptr(o2), ptr(d1), ptr(d2), ptr(acts[0]),
                            ptr(acts[1]), ptr(acts[2]), ptr(upstream), B, P, ctypes.addressof(eng.act_consts),
                            *(ptr(t) for t in bufs), lib.stream())
        eng.backward(g, self.critic, *bufs, d_values)

    @staticmethod
    def _head_outputs(eng, B, P):
        """The head-output buffers of the last forward: (o1, o2, d1, d2); a categorical head has its logits in
        the first slot and no second tensor."""
        o1, o2 = ((eng.buf("wp_olog", (B * P, eng.no)), None) if eng.no
                  else (eng.buf("wp_o1", (B * P,)), eng.buf("wp_o2", (B * P,))))
        d1, d2 = ((eng.buf("wp_dlog", (B * P, eng.nd)), None) if eng.nd
                  else (eng.buf("wp_d1", (B * P,)), eng.buf("wp_d2", (B * P,))))
        return o1, o2, d1, d2

    @staticmethod
    def _grad_buffers(eng, B, P):
        """(d_raw, d_o1, d_o2, d_d1, d_d2); categorical heads: (B*P, classes) logit gradients in the first slot."""
        return (eng.buf("wp_d_raw", (B, P + 1)), eng.buf("wp_g_o1", (B * P, max(eng.no, 1))), eng.buf("wp_g_o2", (B * P,)),
                eng.buf("wp_g_d1", (B * P, max(eng.nd, 1))), eng.buf("wp_g_d2", (B * P,)))
