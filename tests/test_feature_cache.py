"""the device-resident DAgger feature cache against the reference pipeline
(IWTrajectoryDataset items -> collate_fn -> float32 staging, dagger_trainer.py:39-114,189-211,560-566) and the
msgpack_numpy record wire format.  The kernels run on a numpy model here (tests/abi_emulator.py); the CUDA kernels
themselves are compared with the same expectations in tests/test_zz_gpu_variants.py."""
import warnings

import numpy as np
import pytest
import torch

from abi_emulator import Emulator
from oracle import vlnce as oracle_vlnce
from vlnce_b200 import collate, feature_cache, lib

warnings.filterwarnings("ignore")


def trajectories(seed=7, n=6, big=False):
    """what ``_update_dataset`` stores: cached-feature trajectories of ragged length (dagger_trainer.py:334-352)"""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        t = int(torch.randint(1, 9, (1,), generator=g))
        c = 2048 if big else 8
        obs = {"rgb_features": torch.rand(t, c, 4, 4, generator=g) * 3, "depth_features": torch.rand(t, 5, 3, generator=g),
               "instruction": torch.randint(1, 2504, (1, 7), generator=g).expand(t, 7).contiguous(),
               "progress": torch.rand(t, 1, generator=g)}
        actions = torch.randint(0, 4, (t,), generator=g)
        prev = torch.cat([torch.zeros(1, dtype=torch.int64), actions[:-1]])
        out.append((obs, prev, actions))
    return out


def reference_pipeline(trajs, idx, coef, fp16):
    """IWTrajectoryDataset.__next__ + collate_fn (pinned bit-for-bit to the reference in tests/test_collate.py) +
    the trainer's ``.to(dtype=torch.float32)``."""
    items = []
    for i in idx:
        obs, prev, act = trajs[i]
        o = {k: (v.numpy().astype(np.float16) if fp16 else v.numpy()) for k, v in obs.items()}  # :342-346
        o = {k: torch.from_numpy(np.copy(v)) for k, v in o.items()}                             # :189-191
        items.append((o, prev, act, oracle_vlnce.inflection_weights(act, coef)))
    obs, prev, masks, actions, weights = collate.collate_fn(items)
    return {k: v.to(torch.float32) for k, v in obs.items()}, prev, masks, actions, weights


@pytest.mark.parametrize("fp16", [False, True])
def test_cache_batch_equals_reference_pipeline(fp16):
    trajs = trajectories()
    with lib.dry_run(executor=Emulator()) as rec:
        cache = feature_cache.TrajectoryCache("cpu", fp16=fp16)
        for obs, prev, act in trajs:
            cache.append(obs, prev, act)
        assert len(cache) == len(trajs)
        for idx in ([0, 1, 2, 3, 4, 5], [4, 2], [3], [5, 5, 0]):
            got = cache.batch(idx, inflection_weight_coef=3.2)
            want = reference_pipeline(trajs, idx, 3.2, fp16)
            assert list(got[0]) == list(want[0])
            for k in want[0]:
                assert got[0][k].dtype == torch.float32 and torch.equal(got[0][k], want[0][k]), k
            for a, b in zip(got[1:], want[1:]):
                assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    assert rec.calls.count("vlnce_collate_rows") == 4 * 4 and rec.calls.count("vlnce_collate_actions") == 4
    if fp16:  # the reference's lmdb_fp16 casts every key, token ids included: ids above 2048 lose their last bit
        ids = torch.cat([t[0]["instruction"][:1] for t in trajs])
        assert (ids > 2048).any() and cache.arenas["instruction"].dtype == torch.float16


def test_cache_without_inflection_weighting():
    trajs = trajectories(seed=9, n=3)
    with lib.dry_run(executor=Emulator()):
        cache = feature_cache.TrajectoryCache("cpu")
        for t in trajs:
            cache.append(*t)
        w = cache.batch([0, 1, 2], inflection_weight_coef=3.2, use_iw=False)[4]
    lens = torch.tensor([t[1].shape[0] for t in trajs])
    assert torch.equal(w, (torch.arange(w.shape[0]).view(-1, 1) < lens.view(1, -1)).float())


def test_cache_rejects_mismatched_trajectories():
    trajs = trajectories(n=2)
    cache = feature_cache.TrajectoryCache("cpu")
    cache.append(*trajs[0])
    bad = dict(trajs[1][0])
    bad.pop("progress")
    with pytest.raises(lib.VlnceError):
        cache.append(bad, trajs[1][1], trajs[1][2])
    bad = dict(trajs[1][0], progress=trajs[1][0]["progress"].expand(-1, 2))
    with pytest.raises(lib.VlnceError):
        cache.append(bad, trajs[1][1], trajs[1][2])


# a record exactly as msgpack_numpy.packb([obs, prev, oracle], use_bin_type=True) lays it out, assembled by hand from
# the msgpack spec: fixarray(3) | fixmap(1) {fixstr "p": ndarray-map} | ndarray-map | ndarray-map, where ndarray-map
# = fixmap(5) {bin8 "nd": true, bin8 "type": fixstr dtype.str, bin8 "kind": bin8 "", bin8 "shape": array, bin8 "data": bin8}
def _nd(dtype_str, shape, payload):
    def b8(x):
        return b"\xc4" + bytes([len(x)]) + x

    out = b"\x85" + b8(b"nd") + b"\xc3" + b8(b"type") + bytes([0xa0 | len(dtype_str)]) + dtype_str.encode()
    out += b8(b"kind") + b8(b"") + b8(b"shape") + bytes([0x90 | len(shape)]) + bytes(shape) + b8(b"data") + b8(payload)
    return out


def test_record_wire_format_known_answer():
    p = np.array([[0.25], [0.5]], dtype="<f4")
    prev, act = np.array([0, 2], dtype="<i8"), np.array([2, 2], dtype="<i8")
    blob = b"\x93" + b"\x81" + b"\xa1p" + _nd("<f4", (2, 1), p.tobytes()) + _nd("<i8", (2,), prev.tobytes()) + \
        _nd("<i8", (2,), act.tobytes())
    assert feature_cache.pack_record({"p": p}, prev, act) == blob
    obs, pr, ac = feature_cache.unpack_record(blob)
    assert obs["p"].dtype == np.float32 and np.array_equal(obs["p"], p)
    assert np.array_equal(pr, prev) and np.array_equal(ac, act) and ac.dtype == np.int64


def test_record_round_trip_into_cache():
    trajs = trajectories(seed=11, n=3)
    with lib.dry_run(executor=Emulator()):
        direct, via = feature_cache.TrajectoryCache("cpu", fp16=True), feature_cache.TrajectoryCache("cpu")
        for obs, prev, act in trajs:
            direct.append(obs, prev, act)
            blob = feature_cache.pack_record({k: v.numpy() for k, v in obs.items()}, prev.numpy(), act.numpy(), fp16=True)
            via.append_record(blob)   # float16 arrays inside the record, as lmdb_fp16 writes them
        a, b = direct.batch([2, 0, 1], 2.0), via.batch([2, 0, 1], 2.0)
    for k in a[0]:
        assert torch.equal(a[0][k], b[0][k])
    for x, y in zip(a[1:], b[1:]):
        assert torch.equal(x, y)
