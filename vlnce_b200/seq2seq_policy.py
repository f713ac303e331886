"""``Seq2SeqPolicy`` / ``Seq2SeqNet`` — the concat-and-GRU baseline of the reference
(``vlnce_baselines/models/seq2seq_policy.py:20-179``) behind the same registry name,
constructor, ``from_config`` and ``state_dict`` keys; forward on the CUDA library."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib
from .aux_losses import AuxLosses
from .cma_policy import _GRUStateEncoder
from .compat import Net, baseline_registry
from .encoders import resnet_encoders
from .encoders.instruction_encoder import InstructionEncoder
from .engine import HeadEngine, linear, off
from .lib import C, ptr
from .policy import ILPolicy


class Seq2SeqEngine(HeadEngine):
    """Launch plan of ``Seq2SeqNet.forward`` (seq2seq_policy.py:142-179)."""

    def __init__(self, net, precision: str = "bf16"):
        super().__init__(net, precision)
        ie = net.instruction_encoder
        self.Hs, self.dirs, self.E = ie.config.hidden_size, (2 if ie.config.bidirectional else 1), ie.config.embedding_size

    def run_backward(self, head_linear, g, d_raw, d_x, d_hat):
        """Hand-written backward of the last ``forward(save=True)`` (backward.seq2seq_backward)."""
        from .backward import seq2seq_backward

        return seq2seq_backward(self, head_linear, g, d_raw, d_x, d_hat)

    def forward(self, observations, rnn_states, prev_actions, masks, save: bool = False, visual=None):
        assert visual is None, "the pipelined trunk prefetch is built for the CMA engine only"
        net = self.net
        cfg = net.model_config
        rnn_states, masks_f, prev, R, N, T = self._prep_rows(rnn_states, prev_actions, masks)
        st = lib.stream()
        H = self.H
        ie = net.instruction_encoder
        if ie.config.bidirectional:
            raise NotImplementedError("Seq2Seq with a bidirectional final-state encoder is shape-inconsistent "
                                      "in the reference as well (instruction_encoder.py:89-90)")
        Hs = ie.config.hidden_size
        d_out, r_out = cfg.DEPTH_ENCODER.output_size, cfg.RGB_ENCODER.output_size
        width = Hs + d_out + r_out + (32 if cfg.SEQ2SEQ.use_prev_action else 0)
        xin = self.buf("s2s_in", (R, width))
        # instruction final hidden state -> xin[:, :Hs]
        _, hn, lengths, L = self._encode_instruction(observations, R, save)
        ones = self.buf("ones", (R,))
        ones.fill_(1.0)
        C.vlnce_mask_rows(ptr(hn), Hs, ptr(ones), ptr(xin), width, R, Hs, st)
        # depth: Flatten((R,128,4,4)) -> Linear -> ReLU   (resnet_encoders.py:64-72)
        c, h, w = net.depth_encoder.visual_encoder.output_shape
        dep = self.buf("s2s_depth", (R, h * w, c))
        net.depth_encoder.features_into(observations, dep, c, self.precision, save=save)
        dep_flat = self.buf("s2s_depth_flat", (R, c * h * w))
        C.vlnce_positions_to_nchw(ptr(dep), h * w * c, c, ptr(dep_flat), c * h * w, R, h * w, c, st)
        fc = net.depth_encoder.visual_fc[1]
        linear(ptr(dep_flat), c * h * w, ptr(fc.weight), off(xin, Hs), width, ptr(fc.bias), R, d_out, c * h * w, act=1)
        # rgb: global-pooled (R,2048) -> Linear -> ReLU   (resnet_encoders.py:147-152)
        cr = net.rgb_encoder.resnet_layer_size
        rgb = self.buf("s2s_rgb", (R, 1, cr))
        net.rgb_encoder.features_into(observations, rgb, cr, self.precision, save=save)
        fc = net.rgb_encoder.fc[1]
        linear(ptr(rgb), cr, ptr(fc.weight), off(xin, Hs + d_out), width, ptr(fc.bias), R, r_out, cr, act=1)
        if cfg.SEQ2SEQ.use_prev_action:  # seq2seq_policy.py:158-162
            pa = net.prev_action_embedding.weight
            C.vlnce_prev_action_embed(ptr(prev), ptr(masks_f), ptr(pa), off(xin, width - 32), width, None, 0, R, 32,
                                      pa.shape[0], st)
        # modality ablations (seq2seq_policy.py:147-152): the encoder OUTPUTS are multiplied by 0
        if cfg.get("ablate_instruction", False):
            xin[:, :Hs].zero_()
        if cfg.get("ablate_depth", False):
            xin[:, Hs:Hs + d_out].zero_()
        if cfg.get("ablate_rgb", False):
            xin[:, Hs + d_out:Hs + d_out + r_out].zero_()
        x_out = self.buf("x_out", (R, H))
        Lr = net.num_recurrent_layers  # 1 (GRU) or 2 (LSTM: h, c)
        h_out = self.buf("rnn_states_out", (N, Lr, H))
        self._gru_forward("gru", net.state_encoder.rnn, ptr(xin), width, width, ptr(rnn_states), Lr * H, masks_f,
                          ptr(x_out), H, ptr(h_out), Lr * H, T, N, save)
        hat = None
        if cfg.PROGRESS_MONITOR.use:
            hat = self.buf("progress_hat", (R, 1))
            pm = net.progress_monitor
            linear(ptr(x_out), H, ptr(pm.weight), ptr(hat), 1, ptr(pm.bias), R, 1, H, act=2)
        self.ctx.update(R=R, N=N, T=T, L=L, lengths=lengths, prev=prev, d_out=d_out, r_out=r_out, width=width,
                        saved=save)
        return x_out, h_out, hat


class Seq2SeqNet(Net):
    def __init__(self, observation_space, model_config, num_actions):
        nn.Module.__init__(self)
        self.model_config = model_config
        self.instruction_encoder = InstructionEncoder(model_config.INSTRUCTION_ENCODER)
        assert model_config.DEPTH_ENCODER.cnn_type in ["VlnResnetDepthEncoder"]
        self.depth_encoder = getattr(resnet_encoders, model_config.DEPTH_ENCODER.cnn_type)(
            observation_space, output_size=model_config.DEPTH_ENCODER.output_size,
            checkpoint=model_config.DEPTH_ENCODER.ddppo_checkpoint,
            backbone=model_config.DEPTH_ENCODER.backbone, trainable=model_config.DEPTH_ENCODER.trainable)
        assert model_config.RGB_ENCODER.cnn_type in ["TorchVisionResNet18", "TorchVisionResNet50"]
        self.rgb_encoder = getattr(resnet_encoders, model_config.RGB_ENCODER.cnn_type)(
            model_config.RGB_ENCODER.output_size, normalize_visual_inputs=model_config.normalize_rgb,
            trainable=model_config.RGB_ENCODER.trainable, spatial_output=False)
        if model_config.SEQ2SEQ.use_prev_action:
            self.prev_action_embedding = nn.Embedding(num_actions + 1, 32)
        rnn_input_size = (self.instruction_encoder.output_size + model_config.DEPTH_ENCODER.output_size
                          + model_config.RGB_ENCODER.output_size)
        if model_config.SEQ2SEQ.use_prev_action:
            rnn_input_size += self.prev_action_embedding.embedding_dim
        self.state_encoder = _GRUStateEncoder(rnn_input_size, model_config.STATE_ENCODER.hidden_size,
                                              model_config.STATE_ENCODER.rnn_type, 1)
        self.progress_monitor = nn.Linear(model_config.STATE_ENCODER.hidden_size, 1)
        nn.init.kaiming_normal_(self.progress_monitor.weight, nonlinearity="tanh")
        nn.init.constant_(self.progress_monitor.bias, 0)
        self.precision = "bf16"
        self.train()

    @property
    def output_size(self):
        return self.model_config.STATE_ENCODER.hidden_size

    @property
    def is_blind(self):
        return False

    @property
    def num_recurrent_layers(self):
        return self.state_encoder.num_recurrent_layers

    def engine(self) -> Seq2SeqEngine:
        eng = self.__dict__.get("_engine")
        if eng is None or eng.precision != self.precision:
            eng = Seq2SeqEngine(self, self.precision)
            self.__dict__["_engine"] = eng
        return eng

    def _run(self, head_linear, observations, rnn_states, prev_actions, masks):
        from .autograd import cma_features  # one autograd node around the engine (forward + hand-written backward)

        out, h_out, hat = cma_features(self, head_linear, observations, rnn_states, prev_actions, masks)
        if hat is not None and AuxLosses.is_active():  # seq2seq_policy.py:166-177
            loss = F.mse_loss(hat.squeeze(1), observations["progress"], reduction="none")
            AuxLosses.register_loss("progress_monitor", loss, self.model_config.PROGRESS_MONITOR.alpha)
        return out, h_out

    def forward(self, observations, rnn_states, prev_actions, masks):
        return self._run(None, observations, rnn_states, prev_actions, masks)

    def logits_and_states(self, head_linear, observations, rnn_states, prev_actions, masks):
        return self._run(head_linear, observations, rnn_states, prev_actions, masks)


@baseline_registry.register_policy
class Seq2SeqPolicy(ILPolicy):
    def __init__(self, observation_space, action_space, model_config):
        super().__init__(Seq2SeqNet(observation_space=observation_space, model_config=model_config,
                                    num_actions=action_space.n), action_space.n)

    @classmethod
    def from_config(cls, config, observation_space, action_space):
        config.defrost()
        config.MODEL.TORCH_GPU_ID = config.TORCH_GPU_ID
        config.freeze()
        return cls(observation_space=observation_space, action_space=action_space, model_config=config.MODEL)
