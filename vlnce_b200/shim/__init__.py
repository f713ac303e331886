"""Minimal stand-ins for the habitat-lab / habitat_baselines v0.1.7 and gym symbols the
reference's policy files import (``cma_policy.py:7-14``, ``policy.py:4-5``,
``resnet_encoders.py:8-11``), used only when those packages are not installed (they are
not in this image).  With the real packages present ``vlnce_b200.compat`` binds to them
instead, so the policies register in the reference's own ``baseline_registry`` and drop
into ``run.py`` unchanged.  Only the *interface* lives here — every forward runs in
``libvlnce_b200.so``.
"""
from __future__ import annotations

import copy

import numpy as np
import torch
import torch.nn as nn


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
    """yacs ``CfgNode`` look-alike: attribute access + ``defrost/freeze/clone``."""

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
        pass

    def freeze(self):
        pass

    def clone(self):
        return copy.deepcopy(self)


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


class Net(nn.Module):
    pass


class Policy(nn.Module):
    pass


class CustomFixedCategorical(torch.distributions.Categorical):
    """``models/utils.py:269-289`` interface (sample keeps a trailing dim, ``mode``)."""

    def sample(self, sample_shape=torch.Size()):  # noqa: B008
        return super().sample(sample_shape).unsqueeze(-1)

    def log_prob(self, actions):
        return (super().log_prob(actions.squeeze(-1)).view(actions.size(0), -1)
                .sum(-1).unsqueeze(-1))

    def mode(self):
        return self.probs.argmax(dim=-1, keepdim=True)
