"""Bind to the real habitat_baselines / gym when installed (drop-in into the reference's
``run.py``), else to the in-package shim (this image has neither)."""
try:  # pragma: no cover - not installable in the build image
    from gym import Space, spaces  # type: ignore
    from habitat import Config  # type: ignore
    from habitat_baselines.common.baseline_registry import baseline_registry  # type: ignore
    from habitat_baselines.rl.ppo.policy import Net, Policy  # type: ignore

    Box, Discrete, Dict = spaces.Box, spaces.Discrete, spaces.Dict
    HAVE_HABITAT = True
except Exception:  # ImportError and friends
    from .shim import (Box, Config, Dict, Discrete, Net, Policy, Space,  # noqa: F401
                       baseline_registry)

    HAVE_HABITAT = False

from .shim import CustomFixedCategorical  # noqa: E402,F401  (same class in both worlds)
