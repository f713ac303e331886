"""Visual encoders with the reference's module tree / state_dict keys
(``vlnce_baselines/models/encoders/resnet_encoders.py``) whose forward runs on the
hand-written conv trunks (``vlnce_b200.trunk``) instead of torchvision / habitat modules.

The torch modules below are *parameter containers*: they fix the key names
(``cnn.0.weight``, ``visual_encoder.backbone.layer1.0.convs.0.weight`` …) so reference
checkpoints load unchanged; none of their ``forward`` methods is ever used for compute.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .. import lib
from ..lib import C, ptr
from ..trunk import DepthTrunk, RGBTrunk

import os as _os

TF32_DEPTH_FP16 = _os.environ.get("VLNCE_TF32_DEPTH", "fp16") != "tf32"


# --------------------------------------------------------------------------------------
# parameter containers for the DD-PPO GroupNorm ResNet (habitat_baselines v0.1.7
# rl/ddppo/policy/resnet.py + resnet_policy.py ResNetEncoder) — structure only
# --------------------------------------------------------------------------------------
def _conv(cin, cout, k, stride=1, pad=0):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=False)


class _GNBottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, ngroups, stride=1, downsample=None):
        super().__init__()
        self.convs = nn.Sequential(
            _conv(inplanes, planes, 1), nn.GroupNorm(ngroups, planes), nn.ReLU(True),
            _conv(planes, planes, 3, stride, 1), nn.GroupNorm(ngroups, planes), nn.ReLU(True),
            _conv(planes, planes * 4, 1), nn.GroupNorm(ngroups, planes * 4))
        self.relu = nn.ReLU(True)
        self.downsample = downsample


class _GNBasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, ngroups, stride=1, downsample=None):
        super().__init__()
        self.convs = nn.Sequential(
            _conv(inplanes, planes, 3, stride, 1), nn.GroupNorm(ngroups, planes), nn.ReLU(True),
            _conv(planes, planes, 3, 1, 1), nn.GroupNorm(ngroups, planes))
        self.downsample = downsample
        self.relu = nn.ReLU(True)


class _GNResNet(nn.Module):
    def __init__(self, in_channels, base_planes, ngroups, block, layers):
        super().__init__()
        self.conv1 = nn.Sequential(_conv(in_channels, base_planes, 7, 2, 3),
                                   nn.GroupNorm(ngroups, base_planes), nn.ReLU(True))
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.inplanes = base_planes
        self.layer1 = self._make(block, ngroups, base_planes, layers[0], 1)
        self.layer2 = self._make(block, ngroups, base_planes * 2, layers[1], 2)
        self.layer3 = self._make(block, ngroups, base_planes * 4, layers[2], 2)
        self.layer4 = self._make(block, ngroups, base_planes * 8, layers[3], 2)
        self.final_channels = self.inplanes
        self.final_spatial_compress = 1.0 / 32

    def _make(self, block, ngroups, planes, blocks, stride):
        ds = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            ds = nn.Sequential(_conv(self.inplanes, planes * block.expansion, 1, stride),
                               nn.GroupNorm(ngroups, planes * block.expansion))
        out = [block(self.inplanes, planes, ngroups, stride, ds)]
        self.inplanes = planes * block.expansion
        out += [block(self.inplanes, planes, ngroups) for _ in range(1, blocks)]
        return nn.Sequential(*out)


_BACKBONES = {"resnet50": (_GNBottleneck, [3, 4, 6, 3]), "resnet18": (_GNBasicBlock, [2, 2, 2, 2])}


class _DepthVisualEncoder(nn.Module):
    """``visual_encoder`` of the reference: accepts forward hooks; calling it returns the
    (R,128,4,4) NCHW feature map the DAgger feature cache stores (dagger_trainer.py:300-314)."""

    def __init__(self, depth_hw, backbone, baseplanes, ngroups):
        super().__init__()
        block, layers = _BACKBONES[backbone]
        self.running_mean_and_var = nn.Sequential()
        self.backbone = _GNResNet(1, baseplanes, ngroups, block, layers)
        spatial = depth_hw // 2
        final_spatial = int(spatial * self.backbone.final_spatial_compress)
        nc = int(round(2048 / (final_spatial ** 2)))
        self.compression = nn.Sequential(_conv(self.backbone.final_channels, nc, 3, 1, 1),
                                         nn.GroupNorm(1, nc), nn.ReLU(True))
        self.output_shape = (nc, final_spatial, final_spatial)
        self._trunks = {}

    def trunk(self, precision) -> DepthTrunk:
        dev = self.compression[0].weight.device
        if precision == "tf32" and TF32_DEPTH_FP16:
            # fp16 keeps 10 explicit mantissa bits — TF32's — and GroupNorm renormalises every layer of this trunk to
            # O(1), so the range fp16 gives up is never used: in the tf32 mode the depth trunk stays on the
            # image-stationary fp16 kernel (1.07 ms per 64 images; the per-layer TF32 path is 2.97 ms, error 5.0e-3 vs
            # 4.5e-3).  VLNCE_TF32_DEPTH=tf32 selects the per-layer fp32-storage path.
            precision = "bf16"
        key = (precision, dev)
        if key not in self._trunks:
            self._trunks[key] = DepthTrunk(self, dev, precision)
        return self._trunks[key]

    def forward(self, observations):  # hook-compatible entry (NCHW out)
        depth = observations["depth"]
        lib.require_cuda(depth, "depth")
        c, h, w = self.output_shape
        out = torch.empty(depth.shape[0], h * w, c, device=depth.device, dtype=torch.float32)
        self.trunk(getattr(self, "precision", "bf16")).forward(depth.contiguous().float(), out, c, h)
        return out.view(-1, h, w, c).permute(0, 3, 1, 2).contiguous()


class VlnResnetDepthEncoder(nn.Module):
    """``resnet_encoders.py:17-115``."""

    def __init__(self, observation_space, output_size=128, checkpoint="NONE", backbone="resnet50",
                 resnet_baseplanes=32, normalize_visual_inputs=False, trainable=False,
                 spatial_output=False):
        super().__init__()
        if normalize_visual_inputs:
            raise NotImplementedError("RunningMeanAndVar input normalisation is not on the reference path")
        self.trainable = bool(trainable)
        shape = observation_space.spaces["depth"].shape
        depth_hw = shape[-3]
        self.visual_encoder = _DepthVisualEncoder(depth_hw, backbone, resnet_baseplanes,
                                                  resnet_baseplanes // 2)
        for p in self.visual_encoder.parameters():
            p.requires_grad_(self.trainable)  # resnet_encoders.py:44-45
        if checkpoint != "NONE":  # resnet_encoders.py:48-61
            w = torch.load(checkpoint, map_location="cpu")
            sd = {}
            for k, v in w["state_dict"].items():
                parts = k.split(".")[2:]
                if parts[0] == "visual_encoder":
                    sd[".".join(parts[1:])] = v
            self.visual_encoder.load_state_dict(sd, strict=True)
        self.spatial_output = spatial_output
        if not spatial_output:
            self.output_shape = (output_size,)
            self.visual_fc = nn.Sequential(
                nn.Flatten(), nn.Linear(int(np.prod(self.visual_encoder.output_shape)), output_size),
                nn.ReLU(True))
        else:
            c, h, w = self.visual_encoder.output_shape
            self.spatial_embeddings = nn.Embedding(h * w, 64)
            self.output_shape = (c + 64, h, w)

    @property
    def is_blind(self):
        return False

    def features_into(self, observations, out: torch.Tensor, ld: int, precision: str, save: bool = False):
        """Writes the (R, P, C) channels-last trunk features into ``out[:, :, :C]`` (row stride
        ``ld``), from the cached ``depth_features`` (resnet_encoders.py:92-93) or the trunk.  A trainable trunk
        (``DEPTH_ENCODER.trainable``) runs the fp32 path and, with ``save``, records the tape its backward walks."""
        c, h, w = self.visual_encoder.output_shape
        st = lib.stream()
        self._taped = None
        if "depth_features" in observations:
            f = observations["depth_features"]
            lib.require_cuda(f, "depth_features")
            f = f.contiguous().float()
            C.vlnce_nchw_to_positions(ptr(f), ptr(out), ld, f.shape[0], c, h * w, st)
            nchw = f
        else:
            depth = observations["depth"]
            lib.require_cuda(depth, "depth")
            tr = self.visual_encoder.trunk("fp32" if self.trainable else precision)
            tr.tape = None
            if self.trainable:  # the fused optimiser updates the flat parameter buffer in place, behind torch's version
                for cv in tr._all():  # counters: repack the weights on every call
                    cv._ver = None
            if self.trainable and save:
                tr.begin_tape()
                self._taped = tr
            tr.forward(depth.contiguous().float(), out, ld, h)
            nchw = None
        _run_hooks(self.visual_encoder, observations, out, ld, c, h, w, nchw)

    def backward_trunk(self, dout_ptr: int, ld: int, g, pname):
        """Trunk parameter gradients of the last ``features_into(..., save=True)`` from the gradient w.r.t. the
        (R, P, ld) positions it wrote (no-op when the trunk did not run: frozen, or cached features)."""
        if getattr(self, "_taped", None) is not None:
            self._taped.backward(dout_ptr, ld, g, pname)
            self._taped.tape = None
            self._taped = None


# --------------------------------------------------------------------------------------
class _SpatialAvgPool(nn.Module):
    pass


class _RGBSequential(nn.Sequential):
    """``cnn`` of the reference (children 0..7 of torchvision ResNet [+ pools]); calling it on a
    normalised NCHW batch (the reference's calling convention, resnet_encoders.py:198-199)
    runs the CUDA trunk and returns NCHW features, so forward hooks keep working."""

    out_hw = 4

    def trunk(self, precision, normalize: bool = False) -> RGBTrunk:
        dev = self[0].weight.device
        cache = self.__dict__.setdefault("_trunks", {})
        half = torch.float16 if self[1].training else torch.bfloat16
        key = (precision, dev, half, bool(normalize))
        if key not in cache:
            cache[key] = RGBTrunk(self, dev, precision, half, normalize=normalize)
        return cache[key]

    def forward(self, x):
        lib.require_cuda(x, "rgb")
        # inverse of the reference's normalize(): the trunk consumes 0..255 NHWC
        img = (x * 255.0).permute(0, 2, 3, 1).contiguous().float()
        tr = self.trunk(getattr(self, "precision", "bf16"))
        o = self.out_hw
        out = torch.empty(x.shape[0], o * o, tr.out_channels, device=x.device, dtype=torch.float32)
        tr.forward(img, out, tr.out_channels, o)
        return out.view(-1, o, o, tr.out_channels).permute(0, 3, 1, 2).contiguous()


def _run_hooks(module: nn.Module, inp, out, ld, c, h, w, nchw: Optional[torch.Tensor]):
    """Feed registered forward hooks the NCHW map they expect (dagger_trainer.py:300-314)."""
    if not module._forward_hooks:
        return
    if nchw is None:
        nchw = out.view(-1, h * w, ld)[:, :, :c].reshape(-1, h, w, c).permute(0, 3, 1, 2).contiguous()
    for hook in list(module._forward_hooks.values()):
        hook(module, (inp,), nchw)


class TorchVisionResNet(nn.Module):
    """``resnet_encoders.py:118-219``."""

    def __init__(self, output_size, resnet_version="resnet50", normalize_visual_inputs=False,
                 trainable=False, spatial_output=False, single_spatial_filter=True):
        super().__init__()
        import torchvision.models as tvm

        self.trainable = bool(trainable)
        self.normalize_visual_inputs = normalize_visual_inputs  # MODEL.normalize_rgb (config/default.py:216)
        self.spatial_output = spatial_output
        # reference: pretrained=True (resnet_encoders.py:136); no network here -> seeded init,
        # real weights arrive through load_state_dict.
        try:
            net = getattr(tvm, resnet_version)(weights=None)
        except TypeError:
            net = getattr(tvm, resnet_version)(pretrained=False)
        mods = list(net.children())
        self.resnet_layer_size = mods[-1].in_features
        self.cnn = _RGBSequential(*mods[:-1])
        for p in self.cnn.parameters():
            p.requires_grad_(self.trainable)  # resnet_encoders.py:141-143
        self.cnn.train(self.trainable)
        if not spatial_output:
            self.output_shape = (output_size,)
            self.cnn.out_hw = 1
            self.fc = nn.Sequential(nn.Flatten(), nn.Linear(self.resnet_layer_size, output_size), nn.ReLU())
        else:
            if single_spatial_filter:
                self.cnn = _RGBSequential(*list(self.cnn.children())[:-1])
                self.cnn.out_hw = 4
            else:
                self.cnn.out_hw = 1  # Q2: global pool stays, the 4x4 map is a broadcast
            self.cnn.avgpool = _SpatialAvgPool()
            self.spatial_embeddings = nn.Embedding(16, 64)
            self.output_shape = (self.resnet_layer_size + 64, 4, 4)

    @property
    def is_blind(self):
        return False

    def features_into(self, observations, out: torch.Tensor, ld: int, precision: str, save: bool = False):
        """(R, P, 2048|512) channels-last trunk features into ``out[:, :, :C]``.  A trainable trunk
        (``RGB_ENCODER.trainable``) runs the fp32 path and, with ``save``, records the tape its backward walks."""
        c = self.resnet_layer_size
        o = self.cnn.out_hw
        st = lib.stream()
        self._taped = None
        if "rgb_features" in observations:  # resnet_encoders.py:194-195
            f = observations["rgb_features"]
            lib.require_cuda(f, "rgb_features")
            f = f.contiguous().float()
            C.vlnce_nchw_to_positions(ptr(f), ptr(out), ld, f.shape[0], c, f.shape[2] * f.shape[3], st)
            nchw = f
        else:
            rgb = observations["rgb"]
            lib.require_cuda(rgb, "rgb")
            if rgb.dtype not in (torch.uint8, torch.float32):
                rgb = rgb.float()
            tr = self.cnn.trunk("fp32" if self.trainable else precision, self.normalize_visual_inputs)
            tr.tape = None
            if self.trainable:  # see VlnResnetDepthEncoder.features_into
                for cv in tr._all():
                    cv._ver = None
            if self.trainable and save:
                tr.begin_tape()
                self._taped = tr
            tr.forward(rgb.contiguous(), out, ld, o)
            nchw = None
        _run_hooks(self.cnn, observations, out, ld, c, o, o, nchw)

    backward_trunk = VlnResnetDepthEncoder.backward_trunk


class TorchVisionResNet50(TorchVisionResNet):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, resnet_version="resnet50", **kwargs)


class TorchVisionResNet18(TorchVisionResNet):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, resnet_version="resnet18", **kwargs)
