"""Observation transforms in front of the policy on the device: the transformer classes the
reference's configs enable (``RL.POLICY.OBS_TRANSFORMS.ENABLED_TRANSFORMS``: ``ResizeShortestEdge`` of
habitat_baselines v0.1.7, ``CenterCropperPerSensor`` and ``ObsStack`` of ``habitat_extensions/obs_transformers.py``)
with the same constructor arguments and ``forward(observations)`` contract, plus ``apply_obs_transforms_batch``.

``ResizeShortestEdge`` followed by ``CenterCropperPerSensor`` (the RxR recipes, ``rxr_cma_en.yaml:28-30``) is fused
into ONE kernel per sensor (``vlnce_resize_area_crop``) that computes only the cropped window of the resized image
and keeps uint8 frames uint8; cropping / stacking alone are views and copies (no arithmetic).
"""
from __future__ import annotations

import copy

import numbers
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import lib
from .compat import HAVE_HABITAT, Box, baseline_registry
from .lib import C, ptr


def _resized_hw(h: int, w: int, size: int) -> Tuple[int, int]:
    scale = size / min(h, w)  # hb v0.1.7 image_resize_shortest_edge
    return int(h * scale), int(w * scale)


def _crop_origin(h: int, w: int, crop: Tuple[int, int]) -> Tuple[int, int]:
    cy, cx = crop  # hb v0.1.7 center_crop: size is (height, width)
    return h // 2 - cy // 2, w // 2 - cx // 2


def resize_area_crop(img: torch.Tensor, size: Optional[int], crop: Optional[Tuple[int, int]] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N,H,W,C) u8|f32 NHWC -> area-resize so that the shortest edge is ``size`` (None: keep) -> center crop to
    ``crop`` (None: keep).  ``out``: optional (N, ...) destination whose frame n starts at ``out[n]``."""
    lib.require_cuda(img, "observation")
    squeeze = img.dim() == 3
    x = (img.unsqueeze(0) if squeeze else img).contiguous()
    if x.dtype not in (torch.uint8, torch.float32):
        x = x.float()
    n, h, w, c = x.shape
    rh, rw = _resized_hw(h, w, size) if size is not None else (h, w)
    ch, cw = crop if crop is not None else (rh, rw)
    y0, x0 = _crop_origin(rh, rw, (ch, cw))
    if out is None:
        out = torch.empty(n, ch, cw, c, dtype=x.dtype, device=x.device)
        stride = 0
    else:
        assert out.dtype == x.dtype and out.shape[0] == n and tuple(out.shape[-3:]) == (ch, cw, c)
        stride = out.stride(0)
    C.vlnce_resize_area_crop(ptr(x), ptr(out), lib.code(x), stride, n, h, w, c, rh, rw, y0, x0, ch, cw, lib.stream())
    return out.squeeze(0) if squeeze else out


class ResizeShortestEdge:
    """hb v0.1.7 ``ResizeShortestEdge(size, channels_last=True, trans_keys=("rgb","depth","semantic"))``."""

    def __init__(self, size: Optional[int], channels_last: bool = True,
                 trans_keys: Tuple[str, ...] = ("rgb", "depth", "semantic")):
        assert channels_last, "the simulator's layout (NHWC) is the only one on this path"
        self._size, self.trans_keys = size, trans_keys

    def transform_observation_space(self, observation_space):
        """hb v0.1.7 ``ResizeShortestEdge.transform_observation_space``: image boxes shrink to the resized (h, w)."""
        observation_space = copy.deepcopy(observation_space)
        if self._size is not None:
            for key, sp in observation_space.spaces.items():
                if key in self.trans_keys and len(sp.shape) >= 3:
                    h, w = sp.shape[-3], sp.shape[-2]
                    if self._size != min(h, w):
                        observation_space.spaces[key] = _box_with_hw(sp, _resized_hw(h, w, self._size))
        return observation_space

    @classmethod
    def from_config(cls, config):
        return cls(config.RL.POLICY.OBS_TRANSFORMS.RESIZE_SHORTEST_EDGE.SIZE)

    @torch.no_grad()
    def forward(self, observations: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self._size is not None:
            observations.update({k: resize_area_crop(observations[k], self._size) for k in self.trans_keys
                                 if k in observations})
        return observations

    __call__ = forward


class CenterCropperPerSensor:
    """``habitat_extensions/obs_transformers.py:21-87``."""

    def __init__(self, sensor_crops: List[Tuple[str, Union[int, Tuple[int, int]]]], channels_last: bool = True):
        assert channels_last
        self.sensor_crops = {k: ((int(v), int(v)) if isinstance(v, numbers.Number) else tuple(v))
                             for k, v in dict(sensor_crops).items()}

    def transform_observation_space(self, observation_space):
        """``obs_transformers.py:43-65``."""
        observation_space = copy.deepcopy(observation_space)
        for key, sp in observation_space.spaces.items():
            if key in self.sensor_crops and tuple(sp.shape[-3:-1]) != self.sensor_crops[key]:
                observation_space.spaces[key] = _box_with_hw(sp, self.sensor_crops[key])
        return observation_space

    @classmethod
    def from_config(cls, config):
        """``obs_transformers.py:83-87``."""
        return cls(config.RL.POLICY.OBS_TRANSFORMS.CENTER_CROPPER_PER_SENSOR.SENSOR_CROPS)

    @torch.no_grad()
    def forward(self, observations):
        for k, (cy, cx) in self.sensor_crops.items():
            if k in observations:
                img = observations[k]
                y0, x0 = _crop_origin(img.shape[-3], img.shape[-2], (cy, cx))
                observations[k] = img[..., y0:y0 + cy, x0:x0 + cx, :]
        return observations

    __call__ = forward


class ObsStack:
    """``habitat_extensions/obs_transformers.py:90-145``: stack several sensors into one (B, K, ...)."""

    def __init__(self, sensor_rewrites: List[Tuple[str, Sequence[str]]]):
        self.rewrite_dict = dict(sensor_rewrites)

    def transform_observation_space(self, observation_space):
        """``obs_transformers.py:102-128``: the stacked sensor is a Box of shape (K,) + sensor shape."""
        import numpy as np

        observation_space = copy.deepcopy(observation_space)
        for target, sensors in self.rewrite_dict.items():
            orig = observation_space.spaces[sensors[0]]
            for k in sensors:
                del observation_space.spaces[k]
            low = orig.low if np.isscalar(orig.low) else np.min(orig.low)
            high = orig.high if np.isscalar(orig.high) else np.max(orig.high)
            observation_space.spaces[target] = Box(low=low, high=high, shape=(len(sensors),) + tuple(orig.shape),
                                                   dtype=orig.dtype)
        return observation_space

    @classmethod
    def from_config(cls, config):
        """``obs_transformers.py:143-145``."""
        return cls(config.RL.POLICY.OBS_TRANSFORMS.OBS_STACK.SENSOR_REWRITES)

    @torch.no_grad()
    def forward(self, observations):
        for new_key, old_keys in self.rewrite_dict.items():
            stacked = torch.stack([observations[k] for k in old_keys], dim=1)
            for k in old_keys:
                del observations[k]
            observations[new_key] = stacked
        return observations

    __call__ = forward


def _box_with_hw(box, hw):
    """hb v0.1.7 ``overwrite_gym_box_shape``: the same Box with its (h, w) replaced (channels last)."""
    import numpy as np

    shape = tuple(box.shape[:-3]) + (int(hw[0]), int(hw[1])) + tuple(box.shape[-1:])
    low = box.low if np.isscalar(box.low) else np.min(box.low)
    high = box.high if np.isscalar(box.high) else np.max(box.high)
    return Box(low=low, high=high, shape=shape, dtype=box.dtype)


def apply_obs_transforms_obs_space(obs_space, obs_transforms):
    """hb v0.1.7 ``apply_obs_transforms_obs_space``."""
    for t in obs_transforms:
        obs_space = t.transform_observation_space(obs_space)
    return obs_space


# with habitat_baselines installed the classes register under the reference's names, so
# ``get_active_obs_transforms(config)`` (``ddppo_waypoint_trainer.py:92``, ``recollect_trainer.py:57``) builds these
if HAVE_HABITAT:  # pragma: no cover - not installable in the build image
    for _cls in (CenterCropperPerSensor, ObsStack):
        baseline_registry.register_obs_transformer(_cls, name=_cls.__name__)


def apply_obs_transforms_batch(batch: Dict[str, torch.Tensor], obs_transforms) -> Dict[str, torch.Tensor]:
    """hb v0.1.7 ``apply_obs_transforms_batch``; a ``ResizeShortestEdge`` directly followed by a
    ``CenterCropperPerSensor`` runs as one fused kernel per sensor."""
    i = 0
    while i < len(obs_transforms):
        t = obs_transforms[i]
        nxt = obs_transforms[i + 1] if i + 1 < len(obs_transforms) else None
        if isinstance(t, ResizeShortestEdge) and isinstance(nxt, CenterCropperPerSensor) and t._size is not None:
            for k in t.trans_keys:
                if k in batch:
                    batch[k] = resize_area_crop(batch[k], t._size, nxt.sensor_crops.get(k))
            for k, crop in nxt.sensor_crops.items():
                if k in batch and k not in t.trans_keys:
                    batch = CenterCropperPerSensor([(k, crop)])(batch)
            i += 2
            continue
        batch = t(batch)
        i += 1
    return batch
