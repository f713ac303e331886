"""CPU fp32 restatement of the reference's policy hot path.  TEST INFRASTRUCTURE
(see oracle/__init__).  Module / attribute names follow the reference so that a
reference ``state_dict`` loads unchanged; every block cites the reference lines
it follows (paths relative to ``/root/reference/vlnce_baselines``).

Pinned against the unmodified reference files by
``tests/test_oracle_vs_reference.py`` (build container) and by the golden
vectors under ``tests/golden/`` (everywhere).
"""
from __future__ import annotations

import copy
from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import hb017
from .hb017 import Config


# --------------------------------------------------------------------------- #
# config/default.py:214-285  (MODEL node only; data, not code)
# --------------------------------------------------------------------------- #
def default_model_config(**over) -> Config:
    c = Config(
        policy_name="CMAPolicy",
        normalize_rgb=False,
        ablate_depth=False,
        ablate_rgb=False,
        ablate_instruction=False,
        INSTRUCTION_ENCODER=dict(
            sensor_uuid="instruction", vocab_size=2504, use_pretrained_embeddings=False,
            embedding_file="", dataset_vocab="", fine_tune_embeddings=False,
            embedding_size=50, hidden_size=128, rnn_type="LSTM", final_state_only=True,
            bidirectional=False),
        RGB_ENCODER=dict(cnn_type="TorchVisionResNet50", output_size=256, trainable=False),
        DEPTH_ENCODER=dict(cnn_type="VlnResnetDepthEncoder", output_size=128,
                           backbone="resnet50", ddppo_checkpoint="NONE", trainable=False),
        STATE_ENCODER=dict(hidden_size=512, rnn_type="GRU"),
        PROGRESS_MONITOR=dict(use=False, alpha=1.0),
        SEQ2SEQ=dict(use_prev_action=False),
    )

    _merge(c, over)
    return Config(c)


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and k in dst:
            _merge(dst[k], v)
        else:
            dst[k] = v


def cma_model_config(**over) -> Config:
    """``config/r2r_baselines/cma.yaml:31-35`` (+ optional overrides, e.g. PM on
    as in ``cma_pm.yaml``)."""
    c = default_model_config(policy_name="CMAPolicy", INSTRUCTION_ENCODER=dict(bidirectional=True))
    _merge(c, over)
    return Config(c)


# --------------------------------------------------------------------------- #
# common/aux_losses.py:4-44
# --------------------------------------------------------------------------- #
class AuxLossRegistry:
    def __init__(self):
        self.losses, self.alphas, self.active = {}, {}, False

    def clear(self):
        self.losses.clear()
        self.alphas.clear()

    def register_loss(self, name, loss, alpha=1.0):
        assert self.active and name not in self.losses
        self.losses[name], self.alphas[name] = loss, alpha

    def reduce(self, mask):
        assert self.active
        total = 0.0
        for k, v in self.losses.items():  # aux_losses.py:28-30
            total = total + self.alphas[k] * torch.masked_select(v, mask).mean()
        return total

    def is_active(self):
        return self.active

    def activate(self):
        self.active = True

    def deactivate(self):
        self.active = False


AuxLosses = AuxLossRegistry()


# --------------------------------------------------------------------------- #
# models/encoders/instruction_encoder.py:11-94
# --------------------------------------------------------------------------- #
class InstructionEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        cell = nn.GRU if config.rnn_type == "GRU" else nn.LSTM
        self.encoder_rnn = cell(input_size=config.embedding_size, hidden_size=config.hidden_size,
                                bidirectional=config.bidirectional)
        if config.sensor_uuid == "instruction":
            assert not config.use_pretrained_embeddings, "oracle uses a seeded random table"
            self.embedding_layer = nn.Embedding(config.vocab_size, config.embedding_size,
                                                padding_idx=0)

    @property
    def output_size(self):
        return self.config.hidden_size * (2 if self.config.bidirectional else 1)

    def forward(self, observations):
        if self.config.sensor_uuid == "instruction":  # :70-73
            x = self.embedding_layer(observations["instruction"].long())
        else:
            x = observations["rxr_instruction"]
        # :77-78  a token counts if its embedding row has any non-zero entry
        lengths = ((x != 0.0).long().sum(2) != 0).long().sum(1).cpu()
        packed = nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True,
                                                   enforce_sorted=False)
        out, final = self.encoder_rnn(packed)
        if self.config.rnn_type == "LSTM":
            final = final[0]
        if self.config.final_state_only:  # :89-90
            return final.squeeze(0)
        out = nn.utils.rnn.pad_packed_sequence(out, batch_first=True)[0]
        return out.permute(0, 2, 1)  # (R, C, Lmax)


# --------------------------------------------------------------------------- #
# models/encoders/resnet_encoders.py:17-115
# --------------------------------------------------------------------------- #
def _single_frame_box(box):  # common/utils.py:32-42
    if len(box.shape) < 4:
        return box
    return hb017.Box(low=float(box.low.min()), high=float(box.high.max()),
                     shape=box.shape[1:], dtype=box.dtype)


def _spatial_cat(x, table):
    """resnet_encoders.py:97-113 / 201-217: cat a (16,64) embedding table, viewed
    (1,64,4,4) in table-row-major order (a raw ``view``: element [pos, d] lands at
    flat index pos*64+d of the (64,4,4) block)."""
    b, _, h, w = x.shape
    idx = torch.arange(table.num_embeddings, device=x.device)
    sp = table(idx).view(1, -1, h, w).expand(b, table.embedding_dim, h, w)
    return torch.cat([x, sp], 1)


class VlnResnetDepthEncoder(nn.Module):
    def __init__(self, observation_space, output_size=128, checkpoint="NONE",
                 backbone="resnet50", resnet_baseplanes=32, trainable=False,
                 spatial_output=False):
        super().__init__()
        assert checkpoint == "NONE"
        self.visual_encoder = hb017.ResNetEncoder(
            hb017.Dict({"depth": _single_frame_box(observation_space.spaces["depth"])}),
            baseplanes=resnet_baseplanes, ngroups=resnet_baseplanes // 2,
            make_backbone=getattr(hb017.resnet, backbone), normalize_visual_inputs=False)
        for p in self.visual_encoder.parameters():
            p.requires_grad_(trainable)
        self.spatial_output = spatial_output
        if not spatial_output:
            self.output_shape = (output_size,)
            self.visual_fc = nn.Sequential(
                nn.Flatten(),
                nn.Linear(int(np.prod(self.visual_encoder.output_shape)), output_size),
                nn.ReLU(True))
        else:
            c, h, w = self.visual_encoder.output_shape
            self.spatial_embeddings = nn.Embedding(h * w, 64)
            self.output_shape = (c + 64, h, w)

    def forward(self, observations):
        if "depth_features" in observations:  # :92-93
            x = observations["depth_features"]
        else:
            x = self.visual_encoder(observations)
        if self.spatial_output:
            return _spatial_cat(x, self.spatial_embeddings)
        return self.visual_fc(x)


# --------------------------------------------------------------------------- #
# models/encoders/resnet_encoders.py:118-229
# --------------------------------------------------------------------------- #
class _SpatialAvgPool(nn.Module):
    def forward(self, x):
        return F.adaptive_avg_pool2d(x, (4, 4))


class TorchVisionResNet(nn.Module):
    def __init__(self, output_size, resnet_version="resnet50", normalize_visual_inputs=False,
                 trainable=False, spatial_output=False, single_spatial_filter=True):
        super().__init__()
        import torchvision.models as tvm

        self.normalize_visual_inputs = normalize_visual_inputs
        self.spatial_output = spatial_output
        ctor = getattr(tvm, resnet_version)
        try:
            net = ctor(weights=None)
        except TypeError:  # patched by reference_import (pretrained kwarg form)
            net = ctor(pretrained=False)
        mods = list(net.children())
        self.resnet_layer_size = mods[-1].in_features
        self.cnn = nn.Sequential(*mods[:-1])  # conv1 bn1 relu maxpool layer1-4 avgpool
        for p in self.cnn.parameters():
            p.requires_grad_(trainable)
        self.cnn.train(trainable)
        if not spatial_output:
            self.output_shape = (output_size,)
            self.fc = nn.Sequential(nn.Flatten(), nn.Linear(self.resnet_layer_size, output_size),
                                    nn.ReLU())
        else:
            if single_spatial_filter:  # Q2: drop the global pool, then append the 4x4 pool
                self.cnn = nn.Sequential(*list(self.cnn.children())[:-1])
            self.cnn.avgpool = _SpatialAvgPool()
            self.spatial_embeddings = nn.Embedding(16, 64)
            self.output_shape = (self.resnet_layer_size + 64, 4, 4)

    def forward(self, observations):
        if "rgb_features" in observations:  # :194-195
            y = observations["rgb_features"]
        else:
            x = observations["rgb"].permute(0, 3, 1, 2).contiguous() / 255.0  # :182,198
            if self.normalize_visual_inputs:  # :183-190
                mean = torch.tensor([0.485, 0.456, 0.406], device=x.device)[None, :, None, None]
                std = torch.tensor([0.229, 0.224, 0.225], device=x.device)[None, :, None, None]
                x = x.sub(mean).div(std)
            y = self.cnn(x)
        if self.spatial_output:
            return _spatial_cat(y, self.spatial_embeddings)
        return self.fc(y)


# --------------------------------------------------------------------------- #
# models/cma_policy.py:52-309
# --------------------------------------------------------------------------- #
def _prev_action_index(prev_actions, masks):  # Q6, cma_policy.py:233-235
    return ((prev_actions.float() + 1) * masks).long().view(-1)


def _progress_loss(head, x, observations):  # cma_policy.py:296-302 (Q3 broadcast kept)
    hat = torch.tanh(head(x))
    return F.mse_loss(hat.squeeze(1), observations["progress"], reduction="none")


class CMANet(hb017.Net):
    def __init__(self, observation_space, model_config, num_actions):
        super().__init__()
        self.model_config = model_config
        model_config.INSTRUCTION_ENCODER.final_state_only = False  # :62-64
        self.instruction_encoder = InstructionEncoder(model_config.INSTRUCTION_ENCODER)
        self.depth_encoder = VlnResnetDepthEncoder(
            observation_space, output_size=model_config.DEPTH_ENCODER.output_size,
            checkpoint=model_config.DEPTH_ENCODER.ddppo_checkpoint,
            backbone=model_config.DEPTH_ENCODER.backbone,
            trainable=model_config.DEPTH_ENCODER.trainable, spatial_output=True)
        version = {"TorchVisionResNet50": "resnet50", "TorchVisionResNet18": "resnet18"}[
            model_config.RGB_ENCODER.cnn_type]
        self.rgb_encoder = TorchVisionResNet(
            model_config.RGB_ENCODER.output_size, resnet_version=version,
            normalize_visual_inputs=model_config.normalize_rgb,
            trainable=model_config.RGB_ENCODER.trainable, spatial_output=True)
        self.prev_action_embedding = nn.Embedding(num_actions + 1, 32)
        hs = model_config.STATE_ENCODER.hidden_size
        self._hidden_size = hs
        rgb_out, depth_out = model_config.RGB_ENCODER.output_size, model_config.DEPTH_ENCODER.output_size
        self.rgb_linear = nn.Sequential(
            nn.AdaptiveAvgPool1d(1), nn.Flatten(),
            nn.Linear(self.rgb_encoder.output_shape[0], rgb_out), nn.ReLU(True))
        self.depth_linear = nn.Sequential(
            nn.Flatten(), nn.Linear(int(np.prod(self.depth_encoder.output_shape)), depth_out),
            nn.ReLU(True))
        self.state_encoder = hb017.build_rnn_state_encoder(
            input_size=depth_out + rgb_out + 32, hidden_size=hs,
            rnn_type=model_config.STATE_ENCODER.rnn_type, num_layers=1)
        cat_size = hs + rgb_out + depth_out + self.instruction_encoder.output_size
        self.rgb_kv = nn.Conv1d(self.rgb_encoder.output_shape[0], hs // 2 + rgb_out, 1)
        self.depth_kv = nn.Conv1d(self.depth_encoder.output_shape[0], hs // 2 + depth_out, 1)
        self.state_q = nn.Linear(hs, hs // 2)
        self.text_k = nn.Conv1d(self.instruction_encoder.output_size, hs // 2, 1)
        self.text_q = nn.Linear(self.instruction_encoder.output_size, hs // 2)
        self.register_buffer("_scale", torch.tensor(1.0 / ((hs // 2) ** 0.5)))
        self.second_state_compress = nn.Sequential(nn.Linear(cat_size + 32, hs), nn.ReLU(True))
        self.second_state_encoder = hb017.build_rnn_state_encoder(
            input_size=hs, hidden_size=hs, rnn_type=model_config.STATE_ENCODER.rnn_type,
            num_layers=1)
        self._output_size = hs
        self.progress_monitor = nn.Linear(hs, 1)
        if model_config.PROGRESS_MONITOR.use:  # :201-205
            nn.init.kaiming_normal_(self.progress_monitor.weight, nonlinearity="tanh")
            nn.init.constant_(self.progress_monitor.bias, 0)
        self.train()  # Q1: BatchNorm ends up in train mode (:184)

    output_size = property(lambda self: self._output_size)
    num_recurrent_layers = property(
        lambda self: self.state_encoder.num_recurrent_layers
        + self.second_state_encoder.num_recurrent_layers)

    def _attn(self, q, k, v, mask=None):  # :207-217
        logits = torch.einsum("nc,nci->ni", q, k)
        if mask is not None:
            logits = logits - mask.float() * 1e8
        return torch.einsum("ni,nci->nc", F.softmax(logits * self._scale, dim=1), v)

    def forward(self, observations, rnn_states, prev_actions, masks):
        cfg = self.model_config
        instr = self.instruction_encoder(observations)  # (R,256,L)
        depth = torch.flatten(self.depth_encoder(observations), 2)  # (R,192,16)
        rgb = torch.flatten(self.rgb_encoder(observations), 2)  # (R,2112,16)
        pa = self.prev_action_embedding(_prev_action_index(prev_actions, masks))
        if cfg.ablate_instruction:
            instr = instr * 0
        if cfg.ablate_depth:
            depth = depth * 0
        if cfg.ablate_rgb:
            rgb = rgb * 0
        n1 = self.state_encoder.num_recurrent_layers
        state_in = torch.cat([self.rgb_linear(rgb), self.depth_linear(depth), pa], 1)
        h_out = rnn_states.detach().clone()
        state, h_out[:, 0:n1] = self.state_encoder(state_in, rnn_states[:, 0:n1], masks)
        text_mask = (instr == 0.0).all(dim=1)
        text = self._attn(self.state_q(state), self.text_k(instr), instr, text_mask)
        hk = self._hidden_size // 2
        rgb_k, rgb_v = torch.split(self.rgb_kv(rgb), hk, dim=1)
        depth_k, depth_v = torch.split(self.depth_kv(depth), hk, dim=1)
        tq = self.text_q(text)
        rgb_att = self._attn(tq, rgb_k, rgb_v)
        depth_att = self._attn(tq, depth_k, depth_v)
        x = self.second_state_compress(torch.cat([state, text, rgb_att, depth_att, pa], 1))
        x, h_out[:, n1:] = self.second_state_encoder(x, rnn_states[:, n1:], masks)
        if cfg.PROGRESS_MONITOR.use and AuxLosses.is_active():
            AuxLosses.register_loss("progress_monitor",
                                    _progress_loss(self.progress_monitor, x, observations),
                                    cfg.PROGRESS_MONITOR.alpha)
        return x, h_out


# --------------------------------------------------------------------------- #
# models/seq2seq_policy.py:52-179
# --------------------------------------------------------------------------- #
class Seq2SeqNet(hb017.Net):
    def __init__(self, observation_space, model_config, num_actions):
        super().__init__()
        self.model_config = model_config
        self.instruction_encoder = InstructionEncoder(model_config.INSTRUCTION_ENCODER)
        self.depth_encoder = VlnResnetDepthEncoder(
            observation_space, output_size=model_config.DEPTH_ENCODER.output_size,
            checkpoint=model_config.DEPTH_ENCODER.ddppo_checkpoint,
            backbone=model_config.DEPTH_ENCODER.backbone,
            trainable=model_config.DEPTH_ENCODER.trainable)
        version = {"TorchVisionResNet50": "resnet50", "TorchVisionResNet18": "resnet18"}[
            model_config.RGB_ENCODER.cnn_type]
        self.rgb_encoder = TorchVisionResNet(
            model_config.RGB_ENCODER.output_size, resnet_version=version,
            normalize_visual_inputs=model_config.normalize_rgb,
            trainable=model_config.RGB_ENCODER.trainable, spatial_output=False)
        in_size = (self.instruction_encoder.output_size + model_config.DEPTH_ENCODER.output_size
                   + model_config.RGB_ENCODER.output_size)
        if model_config.SEQ2SEQ.use_prev_action:
            self.prev_action_embedding = nn.Embedding(num_actions + 1, 32)
            in_size += 32
        self.state_encoder = hb017.build_rnn_state_encoder(
            input_size=in_size, hidden_size=model_config.STATE_ENCODER.hidden_size,
            rnn_type=model_config.STATE_ENCODER.rnn_type, num_layers=1)
        self.progress_monitor = nn.Linear(model_config.STATE_ENCODER.hidden_size, 1)
        nn.init.kaiming_normal_(self.progress_monitor.weight, nonlinearity="tanh")
        nn.init.constant_(self.progress_monitor.bias, 0)
        self.train()

    output_size = property(lambda self: self.model_config.STATE_ENCODER.hidden_size)
    num_recurrent_layers = property(lambda self: self.state_encoder.num_recurrent_layers)

    def forward(self, observations, rnn_states, prev_actions, masks):
        cfg = self.model_config
        instr = self.instruction_encoder(observations)
        depth = self.depth_encoder(observations)
        rgb = self.rgb_encoder(observations)
        if cfg.ablate_instruction:
            instr = instr * 0
        if cfg.ablate_depth:
            depth = depth * 0
        if cfg.ablate_rgb:
            rgb = rgb * 0
        x = torch.cat([instr, depth, rgb], 1)
        if cfg.SEQ2SEQ.use_prev_action:
            x = torch.cat([x, self.prev_action_embedding(_prev_action_index(prev_actions, masks))], 1)
        x, h_out = self.state_encoder(x, rnn_states, masks)
        if cfg.PROGRESS_MONITOR.use and AuxLosses.is_active():
            AuxLosses.register_loss("progress_monitor",
                                    _progress_loss(self.progress_monitor, x, observations),
                                    cfg.PROGRESS_MONITOR.alpha)
        return x, h_out


# --------------------------------------------------------------------------- #
# models/policy.py:10-58
# --------------------------------------------------------------------------- #
class ILPolicy(nn.Module):
    def __init__(self, net, dim_actions):
        super().__init__()
        self.net = net
        self.dim_actions = dim_actions
        self.action_distribution = hb017.CategoricalNet(net.output_size, dim_actions)

    def act(self, observations, rnn_states, prev_actions, masks, deterministic=False):
        features, rnn_states = self.net(observations, rnn_states, prev_actions, masks)
        dist = self.action_distribution(features)
        return (dist.mode() if deterministic else dist.sample()), rnn_states

    def build_distribution(self, observations, rnn_states, prev_actions, masks):
        features, _ = self.net(observations, rnn_states, prev_actions, masks)
        return self.action_distribution(features)


class CMAPolicy(ILPolicy):
    def __init__(self, observation_space, action_space, model_config):
        super().__init__(CMANet(observation_space, model_config, action_space.n), action_space.n)

    @classmethod
    def from_config(cls, config, observation_space, action_space):
        return cls(observation_space, action_space, config.MODEL)


class Seq2SeqPolicy(ILPolicy):
    def __init__(self, observation_space, action_space, model_config):
        super().__init__(Seq2SeqNet(observation_space, model_config, action_space.n),
                         action_space.n)

    @classmethod
    def from_config(cls, config, observation_space, action_space):
        return cls(observation_space, action_space, config.MODEL)


# --------------------------------------------------------------------------- #
# common/base_il_trainer.py:134-180
# --------------------------------------------------------------------------- #
def update_agent(policy, optimizer, observations, prev_actions, not_done_masks,
                 corrected_actions, weights, hidden_size=512, step_grad=True,
                 loss_accumulation_scalar=1, aux=AuxLosses):
    t, n = corrected_actions.size()
    h0 = torch.zeros(n, policy.net.num_recurrent_layers, hidden_size)
    aux.clear()
    dist = policy.build_distribution(observations, h0, prev_actions, not_done_masks)
    logits = dist.logits.view(t, n, -1)
    action_loss = F.cross_entropy(logits.permute(0, 2, 1), corrected_actions, reduction="none")
    action_loss = ((weights * action_loss).sum(0) / weights.sum(0)).mean()
    aux_loss = aux.reduce((weights > 0).view(-1)) if aux.is_active() else 0.0
    loss = (action_loss + aux_loss) / loss_accumulation_scalar
    loss.backward()
    if step_grad:
        optimizer.step()
        optimizer.zero_grad()
    if isinstance(aux_loss, torch.Tensor):
        aux_loss = aux_loss.item()
    return loss.item(), action_loss.item(), aux_loss


# --------------------------------------------------------------------------- #
# dagger_trainer.py:39-114 (collate) and :199-211 (inflection weights)
# --------------------------------------------------------------------------- #
def inflection_weights(oracle_actions, coef):
    table = torch.tensor([1.0, coef])
    infl = torch.cat([torch.tensor([1], dtype=torch.long),
                      (oracle_actions[1:] != oracle_actions[:-1]).long()])
    return table[infl]


def collate_fn(batch):
    def pad(t, max_len, fill=0):
        d = max_len - t.size(0)
        if d == 0:
            return t
        return torch.cat([t, torch.full_like(t[0:1], fill).expand(d, *t.size()[1:])], 0)

    obs_l, prev_l, act_l, w_l = map(list, zip(*batch))
    tmax = max(p.size(0) for p in prev_l)
    obs = defaultdict(list)
    for o in obs_l:
        for k, v in o.items():
            obs[k].append(pad(v, tmax, 1.0))  # Q7
    out = {}
    for k, v in obs.items():
        s = torch.stack(v, 1)
        out[k] = s.view(-1, *s.size()[2:])
    prev = torch.stack([pad(p, tmax) for p in prev_l], 1)
    act = torch.stack([pad(a, tmax) for a in act_l], 1)
    w = torch.stack([pad(x, tmax) for x in w_l], 1)
    masks = torch.ones_like(act, dtype=torch.uint8)
    masks[0] = 0
    return out, prev.view(-1, 1), masks.view(-1, 1), act, w
