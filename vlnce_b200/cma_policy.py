"""``CMAPolicy`` / ``CMANet`` — the cross-modal-attention policy of the reference
(``vlnce_baselines/models/cma_policy.py:24-309``) as a drop-in: same registry name,
constructor, ``from_config``, attribute surface (``net.num_recurrent_layers``,
``net.rgb_encoder.cnn``, ``net.depth_encoder.visual_encoder`` …) and ``state_dict`` keys.
The torch sub-modules only hold parameters; ``forward`` sequences CUDA kernels
(``vlnce_b200.engine.CMAEngine``)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .aux_losses import AuxLosses
from .compat import Net, baseline_registry
from .encoders import resnet_encoders
from .encoders.instruction_encoder import InstructionEncoder
from .engine import CMAEngine
from .policy import ILPolicy


class _GRUStateEncoder(nn.Module):
    """Parameter container of habitat_baselines v0.1.7 ``build_rnn_state_encoder`` (GRU, or LSTM with the
    packed ``(h, c)`` hidden state): attribute ``rnn`` (orthogonal weights, zero biases), ``num_recurrent_layers``
    (= layers for a GRU, 2 x layers for an LSTM)."""

    def __init__(self, input_size, hidden_size, rnn_type="GRU", num_layers=1):
        super().__init__()
        rnn_type = rnn_type.upper()
        if rnn_type not in ("GRU", "LSTM"):
            raise ValueError(f"STATE_ENCODER.rnn_type must be GRU or LSTM, got {rnn_type!r}")
        if num_layers != 1:
            raise NotImplementedError("state encoders are single-layer in every net of the reference")
        self.num_recurrent_layers = num_layers * (2 if rnn_type == "LSTM" else 1)
        self.rnn = getattr(nn, rnn_type)(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers)
        for name, p in self.rnn.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(p)
            elif "bias" in name:
                nn.init.constant_(p, 0)


class CMANet(Net):
    def __init__(self, observation_space, model_config, num_actions):
        nn.Module.__init__(self)
        self.model_config = model_config
        model_config.defrost()
        model_config.INSTRUCTION_ENCODER.final_state_only = False
        model_config.freeze()
        self.instruction_encoder = InstructionEncoder(model_config.INSTRUCTION_ENCODER)
        assert model_config.DEPTH_ENCODER.cnn_type in ["VlnResnetDepthEncoder"]
        self.depth_encoder = getattr(resnet_encoders, model_config.DEPTH_ENCODER.cnn_type)(
            observation_space, output_size=model_config.DEPTH_ENCODER.output_size,
            checkpoint=model_config.DEPTH_ENCODER.ddppo_checkpoint,
            backbone=model_config.DEPTH_ENCODER.backbone,
            trainable=model_config.DEPTH_ENCODER.trainable, spatial_output=True)
        assert model_config.RGB_ENCODER.cnn_type in ["TorchVisionResNet18", "TorchVisionResNet50"]
        self.rgb_encoder = getattr(resnet_encoders, model_config.RGB_ENCODER.cnn_type)(
            model_config.RGB_ENCODER.output_size, normalize_visual_inputs=model_config.normalize_rgb,
            trainable=model_config.RGB_ENCODER.trainable, spatial_output=True)
        self.prev_action_embedding = nn.Embedding(num_actions + 1, 32)
        hidden_size = model_config.STATE_ENCODER.hidden_size
        self._hidden_size = hidden_size
        self.rgb_linear = nn.Sequential(
            nn.AdaptiveAvgPool1d(1), nn.Flatten(),
            nn.Linear(self.rgb_encoder.output_shape[0], model_config.RGB_ENCODER.output_size), nn.ReLU(True))
        self.depth_linear = nn.Sequential(
            nn.Flatten(),
            nn.Linear(int(np.prod(self.depth_encoder.output_shape)), model_config.DEPTH_ENCODER.output_size),
            nn.ReLU(True))
        rnn_input_size = (model_config.DEPTH_ENCODER.output_size + model_config.RGB_ENCODER.output_size
                          + self.prev_action_embedding.embedding_dim)
        self.state_encoder = _GRUStateEncoder(rnn_input_size, hidden_size,
                                              model_config.STATE_ENCODER.rnn_type, 1)
        cat_size = (hidden_size + model_config.RGB_ENCODER.output_size
                    + model_config.DEPTH_ENCODER.output_size + self.instruction_encoder.output_size)
        self.rgb_kv = nn.Conv1d(self.rgb_encoder.output_shape[0],
                                hidden_size // 2 + model_config.RGB_ENCODER.output_size, 1)
        self.depth_kv = nn.Conv1d(self.depth_encoder.output_shape[0],
                                  hidden_size // 2 + model_config.DEPTH_ENCODER.output_size, 1)
        self.state_q = nn.Linear(hidden_size, hidden_size // 2)
        self.text_k = nn.Conv1d(self.instruction_encoder.output_size, hidden_size // 2, 1)
        self.text_q = nn.Linear(self.instruction_encoder.output_size, hidden_size // 2)
        self.register_buffer("_scale", torch.tensor(1.0 / ((hidden_size // 2) ** 0.5)))
        self.second_state_compress = nn.Sequential(
            nn.Linear(cat_size + self.prev_action_embedding.embedding_dim, hidden_size), nn.ReLU(True))
        self.second_state_encoder = _GRUStateEncoder(hidden_size, hidden_size,
                                                     model_config.STATE_ENCODER.rnn_type, 1)
        self._output_size = hidden_size
        self.progress_monitor = nn.Linear(self.output_size, 1)
        if model_config.PROGRESS_MONITOR.use:
            nn.init.kaiming_normal_(self.progress_monitor.weight, nonlinearity="tanh")
            nn.init.constant_(self.progress_monitor.bias, 0)
        self.precision = "bf16"  # conv-trunk arithmetic: "bf16" (tcgen05) | "fp32" (CUDA cores)
        self.train()

    @property
    def output_size(self):
        return self._output_size

    @property
    def is_blind(self):
        return False

    @property
    def num_recurrent_layers(self):
        return self.state_encoder.num_recurrent_layers + self.second_state_encoder.num_recurrent_layers

    def engine(self) -> CMAEngine:
        eng = self.__dict__.get("_engine")
        if eng is None or eng.precision != self.precision:
            eng = CMAEngine(self, self.precision)
            self.__dict__["_engine"] = eng
        return eng

    # features + hidden, the reference's Net.forward contract (cma_policy.py:219-309)
    def forward(self, observations, rnn_states, prev_actions, masks):
        from .autograd import cma_features

        x, h_out, hat = cma_features(self, None, observations, rnn_states, prev_actions, masks)
        self._register_progress(hat, observations)
        return x, h_out

    def logits_and_states(self, head_linear, observations, rnn_states, prev_actions, masks):
        from .autograd import cma_features

        logits, h_out, hat = cma_features(self, head_linear, observations, rnn_states, prev_actions, masks)
        self._register_progress(hat, observations)
        return logits, h_out

    def _register_progress(self, hat, observations):
        if hat is not None and AuxLosses.is_active():  # cma_policy.py:296-307 (Q3 broadcast kept)
            loss = F.mse_loss(hat.squeeze(1), observations["progress"], reduction="none")
            AuxLosses.register_loss("progress_monitor", loss, self.model_config.PROGRESS_MONITOR.alpha)


@baseline_registry.register_policy
class CMAPolicy(ILPolicy):
    def __init__(self, observation_space, action_space, model_config):
        super().__init__(CMANet(observation_space=observation_space, model_config=model_config,
                                num_actions=action_space.n), action_space.n)

    @classmethod
    def from_config(cls, config, observation_space, action_space):
        return cls(observation_space=observation_space, action_space=action_space,
                   model_config=config.MODEL)
