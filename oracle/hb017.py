"""Restatement of the habitat-lab / habitat_baselines **v0.1.7** symbols the
reference's policy hot path calls.  TEST INFRASTRUCTURE (see oracle/__init__).

habitat-lab v0.1.7 is pinned by the reference in ``README.md:25-41`` but is not
vendored in ``/root/reference`` and not installed in this image; the reference
holds no tests/golden vectors for it -> **parity unpinned** at this boundary.
What is restated (published algorithm, v0.1.7 module path -> reference call site):

* ``habitat_baselines/rl/ddppo/policy/resnet.py``       -> GroupNorm ResNet
  (``resnet_encoders.py:10,41``)
* ``habitat_baselines/rl/ddppo/policy/resnet_policy.py`` ``ResNetEncoder``
  (``resnet_encoders.py:11,31-43``)
* ``habitat_baselines/rl/models/rnn_state_encoder.py`` ``build_rnn_state_encoder``
  (``cma_policy.py:10-12,126,172``; ``seq2seq_policy.py:7-9,109``)
* ``habitat_baselines/utils/common.py`` ``CategoricalNet`` / ``CustomFixedCategorical``
  (``policy.py:5,19``)
* ``habitat_baselines/rl/ppo/policy.py`` ``Net`` / ``Policy`` / ``CriticHead``
  (``cma_policy.py:13``, ``policy.py:4``)
* ``habitat.Config`` (yacs CfgNode look-alike) and ``gym.spaces`` look-alikes.
"""
from __future__ import annotations

import abc

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# gym.spaces / habitat.Config look-alikes
# --------------------------------------------------------------------------- #
class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)


class Config(dict):
    """Attribute-access dict with no-op ``defrost/freeze`` (yacs CfgNode surface
    used by ``cma_policy.py:62-64``, ``seq2seq_policy.py:41-43``)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, Config):
                self[k] = Config(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def defrost(self):
        for v in self.values():
            if isinstance(v, Config):
                v.defrost()

    def freeze(self):
        for v in self.values():
            if isinstance(v, Config):
                v.freeze()

    def clone(self):
        import copy

        return copy.deepcopy(self)


# --------------------------------------------------------------------------- #
# habitat_baselines.common.baseline_registry
# --------------------------------------------------------------------------- #
class BaselineRegistry:
    _policies: dict = {}
    _trainers: dict = {}

    @classmethod
    def register_policy(cls, to_register=None, *, name=None):
        def wrap(c):
            cls._policies[name or c.__name__] = c
            return c

        return wrap if to_register is None else wrap(to_register)

    @classmethod
    def register_trainer(cls, to_register=None, *, name=None):
        def wrap(c):
            cls._trainers[name or c.__name__] = c
            return c

        return wrap if to_register is None else wrap(to_register)

    @classmethod
    def get_policy(cls, name):
        return cls._policies.get(name)

    @classmethod
    def get_trainer(cls, name):
        return cls._trainers.get(name)


baseline_registry = BaselineRegistry()


# --------------------------------------------------------------------------- #
# habitat_baselines/rl/ddppo/policy/resnet.py  (GroupNorm ResNet)
# --------------------------------------------------------------------------- #
def conv3x3(in_planes, out_planes, stride=1, groups=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride,
                     padding=1, bias=False, groups=groups)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


class BasicBlock(nn.Module):
    expansion = 1
    resneXt = False

    def __init__(self, inplanes, planes, ngroups, stride=1, downsample=None, cardinality=1):
        super().__init__()
        self.convs = nn.Sequential(
            conv3x3(inplanes, planes, stride, groups=cardinality),
            nn.GroupNorm(ngroups, planes),
            nn.ReLU(True),
            conv3x3(planes, planes, groups=cardinality),
            nn.GroupNorm(ngroups, planes),
        )
        self.downsample = downsample
        self.relu = nn.ReLU(True)

    def forward(self, x):
        residual = x
        out = self.convs(x)
        if self.downsample is not None:
            residual = self.downsample(x)
        return self.relu(out + residual)


class Bottleneck(nn.Module):
    expansion = 4
    resneXt = False

    def __init__(self, inplanes, planes, ngroups, stride=1, downsample=None, cardinality=1):
        super().__init__()
        self.convs = nn.Sequential(
            conv1x1(inplanes, planes),
            nn.GroupNorm(ngroups, planes),
            nn.ReLU(True),
            conv3x3(planes, planes, stride, groups=cardinality),
            nn.GroupNorm(ngroups, planes),
            nn.ReLU(True),
            conv1x1(planes, planes * self.expansion),
            nn.GroupNorm(ngroups, planes * self.expansion),
        )
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.convs(x)
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.relu(out + identity)


class GNResNet(nn.Module):
    def __init__(self, in_channels, base_planes, ngroups, block, layers, cardinality=1):
        super().__init__()
        self.conv1 = nn.Sequential(
            nn.Conv2d(in_channels, base_planes, kernel_size=7, stride=2, padding=3, bias=False),
            nn.GroupNorm(ngroups, base_planes),
            nn.ReLU(True),
        )
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.cardinality = cardinality
        self.inplanes = base_planes
        self.layer1 = self._make_layer(block, ngroups, base_planes, layers[0])
        self.layer2 = self._make_layer(block, ngroups, base_planes * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(block, ngroups, base_planes * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(block, ngroups, base_planes * 8, layers[3], stride=2)
        self.final_channels = self.inplanes
        self.final_spatial_compress = 1.0 / (2 ** 5)

    def _make_layer(self, block, ngroups, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                conv1x1(self.inplanes, planes * block.expansion, stride),
                nn.GroupNorm(ngroups, planes * block.expansion),
            )
        layers = [block(self.inplanes, planes, ngroups, stride, downsample,
                        cardinality=self.cardinality)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, ngroups))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.conv1(x)
        x = self.maxpool(x)
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        x = self.layer4(x)
        return x


def resnet18(in_channels, base_planes, ngroups):
    return GNResNet(in_channels, base_planes, ngroups, BasicBlock, [2, 2, 2, 2])


def resnet50(in_channels, base_planes, ngroups):
    return GNResNet(in_channels, base_planes, ngroups, Bottleneck, [3, 4, 6, 3])


class _ResnetModule:
    """Stands in for the module object ``habitat_baselines.rl.ddppo.policy.resnet``
    (selected with ``getattr(resnet, backbone)`` at ``resnet_encoders.py:41``)."""

    resnet18 = staticmethod(resnet18)
    resnet50 = staticmethod(resnet50)


resnet = _ResnetModule()


# --------------------------------------------------------------------------- #
# habitat_baselines/rl/ddppo/policy/resnet_policy.py  ResNetEncoder
# --------------------------------------------------------------------------- #
class ResNetEncoder(nn.Module):
    def __init__(self, observation_space, baseplanes=32, ngroups=32, spatial_size=128,
                 make_backbone=None, normalize_visual_inputs=False):
        super().__init__()
        if "rgb" in observation_space.spaces:
            self._n_input_rgb = observation_space.spaces["rgb"].shape[2]
            spatial_size = observation_space.spaces["rgb"].shape[0] // 2
        else:
            self._n_input_rgb = 0
        if "depth" in observation_space.spaces:
            self._n_input_depth = observation_space.spaces["depth"].shape[2]
            spatial_size = observation_space.spaces["depth"].shape[0] // 2
        else:
            self._n_input_depth = 0
        assert not normalize_visual_inputs, "RunningMeanAndVar unused by the reference path"
        self.running_mean_and_var = nn.Sequential()
        if not self.is_blind:
            input_channels = self._n_input_depth + self._n_input_rgb
            self.backbone = make_backbone(input_channels, baseplanes, ngroups)
            final_spatial = int(spatial_size * self.backbone.final_spatial_compress)
            after_compression_flat_size = 2048
            num_compression_channels = int(round(after_compression_flat_size / (final_spatial ** 2)))
            self.compression = nn.Sequential(
                nn.Conv2d(self.backbone.final_channels, num_compression_channels,
                          kernel_size=3, padding=1, bias=False),
                nn.GroupNorm(1, num_compression_channels),
                nn.ReLU(True),
            )
            self.output_shape = (num_compression_channels, final_spatial, final_spatial)

    @property
    def is_blind(self):
        return self._n_input_rgb + self._n_input_depth == 0

    def layer_init(self):
        for layer in self.modules():
            if isinstance(layer, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(layer.weight, nn.init.calculate_gain("relu"))
                if layer.bias is not None:
                    nn.init.constant_(layer.bias, val=0)

    def forward(self, observations):
        if self.is_blind:
            return None
        cnn_input = []
        if self._n_input_rgb > 0:
            cnn_input.append(observations["rgb"].permute(0, 3, 1, 2) / 255.0)
        if self._n_input_depth > 0:
            cnn_input.append(observations["depth"].permute(0, 3, 1, 2))
        x = torch.cat(cnn_input, dim=1)
        x = F.avg_pool2d(x, 2)
        x = self.running_mean_and_var(x)
        x = self.backbone(x)
        x = self.compression(x)
        return x


# --------------------------------------------------------------------------- #
# habitat_baselines/rl/models/rnn_state_encoder.py
# --------------------------------------------------------------------------- #
class RNNStateEncoder(nn.Module):
    """1-layer GRU/LSTM state encoder.  Hidden I/O is batch-first ``(N, Lr, H)``.
    ``x.size(0) == N`` -> single step; otherwise ``x`` is time-major ``(T*N, .)``
    and the hidden state is zeroed wherever ``masks == 0`` before that step
    (upstream does this with PackedSequences split at episode boundaries; a
    T-step loop is the same arithmetic)."""

    def __init__(self, input_size, hidden_size, num_layers=1, rnn_type="GRU"):
        super().__init__()
        self._rnn_type = rnn_type
        self.num_recurrent_layers = num_layers * (2 if rnn_type == "LSTM" else 1)
        self.rnn = getattr(nn, rnn_type)(input_size=input_size, hidden_size=hidden_size,
                                         num_layers=num_layers)
        for name, param in self.rnn.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(param)
            elif "bias" in name:
                nn.init.constant_(param, 0)

    def _unpack(self, h):
        if self._rnn_type == "LSTM":
            return tuple(torch.chunk(h.contiguous(), 2, 0))
        return h.contiguous()

    def _pack(self, h):
        if self._rnn_type == "LSTM":
            return torch.cat(h, 0)
        return h

    def forward(self, x, hidden_states, masks):
        hidden_states = hidden_states.permute(1, 0, 2)  # (Lr, N, H)
        n = hidden_states.size(1)
        m = masks.to(dtype=x.dtype)
        if x.size(0) == n:
            hidden_states = hidden_states * m.view(1, n, 1)
            x, h = self.rnn(x.unsqueeze(0), self._unpack(hidden_states))
            hidden_states = self._pack(h)
            x = x.squeeze(0)
        else:
            t = x.size(0) // n
            xs = x.view(t, n, -1)
            ms = m.view(t, n)
            outs = []
            for i in range(t):
                hidden_states = hidden_states * ms[i].view(1, n, 1)
                o, h = self.rnn(xs[i].unsqueeze(0), self._unpack(hidden_states))
                hidden_states = self._pack(h)
                outs.append(o.squeeze(0))
            x = torch.cat(outs, 0)
        return x, hidden_states.permute(1, 0, 2)


def build_rnn_state_encoder(input_size, hidden_size, rnn_type="GRU", num_layers=1):
    rnn_type = rnn_type.upper()
    assert rnn_type in ("GRU", "LSTM")
    return RNNStateEncoder(input_size, hidden_size, num_layers, rnn_type)


# --------------------------------------------------------------------------- #
# habitat_baselines/utils/common.py
# --------------------------------------------------------------------------- #
class CustomFixedCategorical(torch.distributions.Categorical):
    def sample(self, sample_shape=torch.Size()):  # noqa: B008
        return super().sample(sample_shape).unsqueeze(-1)

    def log_probs(self, actions):
        return (super().log_prob(actions.squeeze(-1)).view(actions.size(0), -1)
                .sum(-1).unsqueeze(-1))

    def mode(self):
        return self.probs.argmax(dim=-1, keepdim=True)


class CategoricalNet(nn.Module):
    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.linear = nn.Linear(num_inputs, num_outputs)
        nn.init.orthogonal_(self.linear.weight, gain=0.01)
        nn.init.constant_(self.linear.bias, 0)

    def forward(self, x):
        return CustomFixedCategorical(logits=self.linear(x))


# --------------------------------------------------------------------------- #
# habitat_baselines/rl/ppo/policy.py
# --------------------------------------------------------------------------- #
class Net(nn.Module, metaclass=abc.ABCMeta):
    pass


class CriticHead(nn.Module):
    def __init__(self, input_size):
        super().__init__()
        self.fc = nn.Linear(input_size, 1)
        nn.init.orthogonal_(self.fc.weight)
        nn.init.constant_(self.fc.bias, 0)

    def forward(self, x):
        return self.fc(x)


class Policy(nn.Module, metaclass=abc.ABCMeta):
    def __init__(self, net, dim_actions):
        super().__init__()
        self.net = net
        self.dim_actions = dim_actions
        self.action_distribution = CategoricalNet(self.net.output_size, self.dim_actions)
        self.critic = CriticHead(self.net.output_size)


# --------------------------------------------------------------------------- #
# habitat_baselines/rl/ppo/ppo.py + rl/ddppo/algo/ddppo.py (v0.1.7): only what
# ``common/ddppo_alg.py:9-149`` inherits — the constructor (Adam over the parameters
# that require grad), the four hooks around backward/step (``before_step`` clips the global
# gradient norm).  The distributed mixin's reducer priming (``before_backward``) is a
# no-op in a single process.  habitat_baselines/common/rollout_storage.py: the reference
# subclass overrides every method it uses, so the base is an empty shell.
# --------------------------------------------------------------------------- #
class PPO(nn.Module):
    def __init__(self, actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef,
                 entropy_coef, lr=None, eps=None, max_grad_norm=None, use_clipped_value_loss=True,
                 use_normalized_advantage=True):
        super().__init__()
        self.actor_critic = actor_critic
        self.clip_param, self.ppo_epoch, self.num_mini_batch = clip_param, ppo_epoch, num_mini_batch
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.use_normalized_advantage = use_normalized_advantage
        self.optimizer = torch.optim.Adam(
            [p for p in actor_critic.parameters() if p.requires_grad], lr=lr, eps=eps)
        self.device = next(actor_critic.parameters()).device

    def forward(self, *x):
        raise NotImplementedError

    def before_backward(self, loss):
        pass

    def after_backward(self, loss):
        pass

    def before_step(self):
        nn.utils.clip_grad_norm_(self.actor_critic.parameters(), self.max_grad_norm)

    def after_step(self):
        pass


class DDPPO(PPO):
    pass


class RolloutStorage:
    pass
