"""Import the UNMODIFIED reference model files from ``/root/reference`` with the
habitat / gym dependency stubbed by ``oracle/hb017.py``.  TEST INFRASTRUCTURE.

Only usable in the build container (``/root/reference`` does not exist on the GPU
box); used by ``oracle/make_golden.py`` and ``tests/test_oracle_vs_reference.py``
to pin ``oracle/vlnce.py`` and to freeze golden vectors.  Recipe = SURVEY.md
Appendix A.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

from . import hb017

REFERENCE_ROOT = os.environ.get("VLNCE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vlnce_baselines", "models"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


_loaded = None


def load():
    """Returns a namespace with the reference's classes (imported once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    assert available(), f"{REFERENCE_ROOT} not present"
    sys.dont_write_bytecode = True  # reference dir is read-only

    _mod("gym", Space=hb017.Space, spaces=None)
    sp = _mod("gym.spaces", Space=hb017.Space, Box=hb017.Box, Discrete=hb017.Discrete,
              Dict=hb017.Dict)
    sys.modules["gym"].spaces = sp
    _mod("habitat", Config=hb017.Config)
    _mod("habitat.config")
    _mod("habitat.config.default", Config=hb017.Config)
    _mod("habitat.core")
    _mod("habitat.core.simulator", Observations=dict)
    _mod("habitat_baselines")
    _mod("habitat_baselines.common")
    _mod("habitat_baselines.common.baseline_registry",
         baseline_registry=hb017.baseline_registry, BaselineRegistry=hb017.BaselineRegistry)
    _mod("habitat_baselines.utils")
    _mod("habitat_baselines.utils.common", CategoricalNet=hb017.CategoricalNet,
         CustomFixedCategorical=hb017.CustomFixedCategorical)
    _mod("habitat_baselines.rl")
    _mod("habitat_baselines.rl.ppo")
    _mod("habitat_baselines.rl.ppo.policy", Net=hb017.Net, Policy=hb017.Policy,
         CriticHead=hb017.CriticHead)
    _mod("habitat_baselines.rl.models")
    _mod("habitat_baselines.rl.models.rnn_state_encoder",
         build_rnn_state_encoder=hb017.build_rnn_state_encoder)
    _mod("habitat_baselines.rl.ddppo")
    _mod("habitat_baselines.rl.ddppo.algo")
    _mod("habitat_baselines.rl.ddppo.algo.ddppo", DDPPO=hb017.DDPPO)
    _mod("habitat_baselines.common.rollout_storage", RolloutStorage=hb017.RolloutStorage)
    _mod("habitat_baselines.rl.ddppo.policy", resnet=hb017.resnet)
    _mod("habitat_baselines.rl.ddppo.policy.resnet", resnet18=hb017.resnet18,
         resnet50=hb017.resnet50)
    sys.modules["habitat_baselines.rl.ddppo.policy"].resnet = hb017.resnet
    _mod("habitat_baselines.rl.ddppo.policy.resnet_policy", ResNetEncoder=hb017.ResNetEncoder)

    # bare packages so that vlnce_baselines/__init__.py (trainers -> lmdb/habitat) is bypassed
    base = os.path.join(REFERENCE_ROOT, "vlnce_baselines")
    for name, sub in (("vlnce_baselines", ""), ("vlnce_baselines.common", "common"),
                      ("vlnce_baselines.models", "models"),
                      ("vlnce_baselines.models.encoders", "models/encoders")):
        m = _mod(name)
        m.__path__ = [os.path.join(base, sub)]

    # torchvision: pretrained=True needs the network; build seeded random weights instead
    import torchvision.models as tvm

    if not getattr(tvm, "_vlnce_oracle_patched", False):
        for fn in ("resnet18", "resnet50"):
            orig = getattr(tvm, fn)

            def make(orig=orig):
                def wrapped(pretrained=False, **kw):
                    kw.pop("weights", None)
                    return orig(weights=None, **kw)

                return wrapped

            setattr(tvm, fn, make())
        tvm._vlnce_oracle_patched = True

    torch.distributions.Distribution.set_default_validate_args(False)

    ns = types.SimpleNamespace()
    ns.aux_losses = importlib.import_module("vlnce_baselines.common.aux_losses")
    ns.utils = importlib.import_module("vlnce_baselines.models.utils")
    ns.resnet_encoders = importlib.import_module("vlnce_baselines.models.encoders.resnet_encoders")
    ns.instruction_encoder = importlib.import_module(
        "vlnce_baselines.models.encoders.instruction_encoder")
    ns.policy = importlib.import_module("vlnce_baselines.models.policy")
    ns.cma_policy = importlib.import_module("vlnce_baselines.models.cma_policy")
    ns.seq2seq_policy = importlib.import_module("vlnce_baselines.models.seq2seq_policy")
    # waypoint path
    ns.waypoint_predictors = importlib.import_module("vlnce_baselines.models.waypoint_predictors")
    ns.waypoint_policy = importlib.import_module("vlnce_baselines.models.waypoint_policy")
    ns.ddppo_alg = importlib.import_module("vlnce_baselines.common.ddppo_alg")
    ns.rollout_storage = importlib.import_module("vlnce_baselines.common.rollout_storage")
    ns.WaypointPolicy = ns.waypoint_policy.WaypointPolicy
    ns.WDDPPO = ns.ddppo_alg.WDDPPO
    ns.ActionDictRolloutStorage = ns.rollout_storage.ActionDictRolloutStorage
    ns.TruncatedNormal = ns.utils.TruncatedNormal
    ns.AuxLosses = ns.aux_losses.AuxLosses
    ns.CMAPolicy = ns.cma_policy.CMAPolicy
    ns.Seq2SeqPolicy = ns.seq2seq_policy.Seq2SeqPolicy
    _loaded = ns
    return ns
