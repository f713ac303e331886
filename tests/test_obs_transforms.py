"""observation transforms.  The reference's own ``CenterCropperPerSensor`` / ``ObsStack``
(``habitat_extensions/obs_transformers.py``) are imported UNMODIFIED over the restated habitat_baselines helpers
(build container) to pin the restatement; the product's fused resize+crop launch plan is executed through the numpy
model of the kernel and must be bit-exact for uint8 frames (fp32 depth within 1e-6)."""
import importlib
import os
import sys
import types
import warnings

import pytest
import torch

from oracle import hb017, obs_transforms as oracle_ot, reference_import
from vlnce_b200 import lib, obs_transforms as ot

from abi_emulator import Emulator

warnings.filterwarnings("ignore")
GOLD = os.path.join(os.path.dirname(__file__), "golden", "obs_transforms.pt")


def frames(seed=5, n=2, h=60, w=80):
    g = torch.Generator().manual_seed(seed)
    return {"rgb": torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8),
            "depth": torch.rand(n, h, w, 1, generator=g), "instruction": torch.randint(1, 9, (n, 5), generator=g)}


def pipeline(ns, size=32, crops=(("rgb", (28, 28)), ("depth", (32, 32)))):
    return [ns.ResizeShortestEdge(size), ns.CenterCropperPerSensor(list(crops))]


def reference_classes():
    """habitat_extensions/obs_transformers.py with its habitat imports bound to the restated v0.1.7 helpers."""
    reference_import.load()
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    import logging
    mod("habitat.core.logging", logger=logging.getLogger("habitat"))
    sys.modules["habitat.config"].Config = hb017.Config
    mod("habitat_baselines.common.obs_transformers", ObservationTransformer=oracle_ot.ObservationTransformer)
    common = sys.modules["habitat_baselines.utils.common"]
    common.center_crop, common.get_image_height_width = oracle_ot.center_crop, oracle_ot.get_image_height_width
    common.overwrite_gym_box_shape = oracle_ot.overwrite_gym_box_shape
    reg = sys.modules["habitat_baselines.common.baseline_registry"].baseline_registry
    if not hasattr(reg, "register_obs_transformer"):
        type(reg).register_obs_transformer = classmethod(lambda cls, *a, **k: (lambda c: c))
    sys.modules["gym"].spaces = sys.modules["gym.spaces"]
    pkg = types.ModuleType("habitat_extensions")
    pkg.__path__ = [os.path.join(reference_import.REFERENCE_ROOT, "habitat_extensions")]
    sys.modules["habitat_extensions"] = pkg
    return importlib.import_module("habitat_extensions.obs_transformers")


@pytest.mark.skipif(not reference_import.available(), reason="/root/reference not present")
def test_crop_and_stack_equal_reference_classes():
    ref = reference_classes()
    a = ref.CenterCropperPerSensor([("rgb", (28, 20)), ("depth", (16, 16))]).forward(frames())
    b = oracle_ot.CenterCropperPerSensor([("rgb", (28, 20)), ("depth", (16, 16))]).forward(frames())
    c = ot.CenterCropperPerSensor([("rgb", (28, 20)), ("depth", (16, 16))]).forward(frames())
    for k in a:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]), k
    def pano():
        f = frames()
        return {f"rgb_{i}": f["rgb"] + i for i in range(3)}
    keys = [f"rgb_{i}" for i in range(3)]
    sa = ref.ObsStack([("rgb", keys)]).forward(pano())
    sb = oracle_ot.ObsStack([("rgb", keys)]).forward(pano())
    sc = ot.ObsStack([("rgb", keys)]).forward(pano())
    assert list(sa) == list(sb) == list(sc) == ["rgb"] and torch.equal(sa["rgb"], sb["rgb"]) and torch.equal(sa["rgb"], sc["rgb"])


def run_product(fused=True):
    with lib.dry_run(executor=Emulator()) as rec:
        if fused:
            out = ot.apply_obs_transforms_batch(frames(), pipeline(ot))
        else:
            out = frames()
            for t in pipeline(ot):
                out = t(out)
    return out, rec.calls


def test_fused_resize_crop_plan_matches_oracle():
    want = oracle_ot.apply_obs_transforms_batch(frames(), pipeline(oracle_ot))
    for fused in (True, False):
        got, calls = run_product(fused)
        assert calls.count("vlnce_resize_area_crop") == 2
        assert got["rgb"].dtype == torch.uint8 and torch.equal(got["rgb"], want["rgb"])          # bit-exact
        torch.testing.assert_close(got["depth"], want["depth"], rtol=0, atol=1e-6)
        assert torch.equal(got["instruction"], want["instruction"])
    gold = torch.load(GOLD, weights_only=False)
    assert torch.equal(want["rgb"], gold["rgb"]) and torch.allclose(want["depth"], gold["depth"], atol=1e-6)


def test_resize_into_stack_slot():
    """writing the transformed frame straight into a pano-stack slot (y_frame_stride)"""
    f = frames()
    dst = torch.zeros(2, 3, 28, 28, 3, dtype=torch.uint8)
    with lib.dry_run(executor=Emulator()):
        ot.resize_area_crop(f["rgb"], 32, (28, 28), out=dst[:, 1])
    want = oracle_ot.center_crop(oracle_ot.image_resize_shortest_edge(f["rgb"], 32, True), (28, 28), True)
    assert torch.equal(dst[:, 1], want) and int(dst[:, 0].sum()) == 0 and int(dst[:, 2].sum()) == 0


if __name__ == "__main__":
    out = oracle_ot.apply_obs_transforms_batch(frames(), pipeline(oracle_ot))
    torch.save({k: v.clone() for k, v in out.items()}, GOLD)
    print("frozen", GOLD)


def test_observation_space_transforms_and_from_config():
    """``transform_observation_space`` / ``from_config`` of the three transformers (``obs_transformers.py:43-65,83-87,
    102-128,143-145``; hb v0.1.7 ``ResizeShortestEdge``): what ``get_active_obs_transforms`` +
    ``apply_obs_transforms_obs_space`` need so that the policy sees the post-transform sensor shapes."""
    from vlnce_b200 import obs_transforms as ot
    from vlnce_b200.compat import Box, Config, Dict as SpaceDict

    space = SpaceDict({"rgb": Box(0, 255, (480, 640, 3), "uint8"), "depth": Box(0.0, 1.0, (480, 640, 1), "float32"),
                       "rgb_1": Box(0, 255, (224, 224, 3), "uint8"), "rgb_2": Box(0, 255, (224, 224, 3), "uint8")})
    cfg = Config(RL=dict(POLICY=dict(OBS_TRANSFORMS=dict(
        RESIZE_SHORTEST_EDGE=dict(SIZE=256),
        CENTER_CROPPER_PER_SENSOR=dict(SENSOR_CROPS=[("rgb", (224, 224)), ("depth", (256, 256))]),
        OBS_STACK=dict(SENSOR_REWRITES=[("rgb_pano", ["rgb_1", "rgb_2"])])))))
    ts = [ot.ResizeShortestEdge.from_config(cfg), ot.CenterCropperPerSensor.from_config(cfg), ot.ObsStack.from_config(cfg)]
    out = ot.apply_obs_transforms_obs_space(space, ts)
    assert out.spaces["rgb"].shape == (224, 224, 3) and out.spaces["depth"].shape == (256, 256, 1)   # rxr_cma_en.yaml:28-30
    assert out.spaces["rgb_pano"].shape == (2, 224, 224, 3) and "rgb_1" not in out.spaces
    assert space.spaces["rgb"].shape == (480, 640, 3)            # the input space is not modified
    mid = ts[0].transform_observation_space(space)
    assert mid.spaces["rgb"].shape == (256, 341, 3)              # shortest edge 480 -> 256, 640 -> int(640 * 256 / 480)
