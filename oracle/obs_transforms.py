"""CPU restatement of the observation transforms.  TEST INFRASTRUCTURE.

* ``image_resize_shortest_edge`` / ``center_crop`` / ``ResizeShortestEdge`` / ``apply_obs_transforms_batch`` —
  habitat_baselines **v0.1.7** (``common/obs_transformers.py``, ``utils/common.py``): un-vendored, restated from the
  published source, **parity unpinned** at that boundary; the arithmetic is torch's ``F.interpolate(mode="area")``.
* ``CenterCropperPerSensor`` / ``ObsStack`` — ``habitat_extensions/obs_transformers.py:21-145``; pinned by running the
  UNMODIFIED reference classes over these hb functions (``tests/test_obs_transforms.py``).
"""
from __future__ import annotations

import numbers

import torch


def image_resize_shortest_edge(img, size, channels_last=False):
    img = torch.as_tensor(img)
    no_batch = img.dim() == 3
    if img.dim() < 3 or img.dim() > 5:
        raise NotImplementedError()
    if no_batch:
        img = img.unsqueeze(0)
    if channels_last:
        h, w = img.shape[-3:-1]
        img = img.permute(0, 3, 1, 2) if img.dim() == 4 else img.permute(0, 1, 4, 2, 3)
    else:
        h, w = img.shape[-2:]
    scale = size / min(h, w)
    h, w = int(h * scale), int(w * scale)
    img = torch.nn.functional.interpolate(img.float(), size=(h, w), mode="area").to(dtype=img.dtype)
    if channels_last:
        img = img.permute(0, 2, 3, 1) if img.dim() == 4 else img.permute(0, 1, 3, 4, 2)
    return img.squeeze(0) if no_batch else img


def get_image_height_width(img, channels_last=False):
    return tuple(img.shape[-3:-1]) if channels_last else tuple(img.shape[-2:])


def center_crop(img, size, channels_last=False):
    h, w = get_image_height_width(img, channels_last)
    if isinstance(size, numbers.Number):
        size = (int(size), int(size))
    cropy, cropx = size
    startx, starty = w // 2 - cropx // 2, h // 2 - cropy // 2
    if channels_last:
        return img[..., starty:starty + cropy, startx:startx + cropx, :]
    return img[..., starty:starty + cropy, startx:startx + cropx]


def overwrite_gym_box_shape(box, shape):
    return box


class ObservationTransformer(torch.nn.Module):
    def transform_observation_space(self, observation_space, **kw):
        return observation_space


class ResizeShortestEdge(ObservationTransformer):
    def __init__(self, size, channels_last=True, trans_keys=("rgb", "depth", "semantic")):
        super().__init__()
        self._size, self.channels_last, self.trans_keys = size, channels_last, trans_keys

    @torch.no_grad()
    def forward(self, observations):
        if self._size is not None:
            observations.update({k: image_resize_shortest_edge(observations[k], self._size, self.channels_last)
                                 for k in self.trans_keys if k in observations})
        return observations


class CenterCropperPerSensor(ObservationTransformer):  # habitat_extensions/obs_transformers.py:21-87
    def __init__(self, sensor_crops, channels_last=True):
        super().__init__()
        self.sensor_crops = {k: ((int(v), int(v)) if isinstance(v, numbers.Number) else v)
                             for k, v in dict(sensor_crops).items()}
        self.channels_last = channels_last

    @torch.no_grad()
    def forward(self, observations):
        observations.update({k: center_crop(observations[k], self.sensor_crops[k], self.channels_last)
                             for k in self.sensor_crops if k in observations})
        return observations


class ObsStack(ObservationTransformer):  # :90-145
    def __init__(self, sensor_rewrites):
        super().__init__()
        self.rewrite_dict = dict(sensor_rewrites)

    @torch.no_grad()
    def forward(self, observations):
        for new_key, old_keys in self.rewrite_dict.items():
            new = torch.stack([observations[k] for k in old_keys], dim=1)
            for k in old_keys:
                del observations[k]
            observations[new_key] = new
        return observations


def apply_obs_transforms_batch(batch, obs_transforms):
    for t in obs_transforms:
        batch = t(batch)
    return batch
