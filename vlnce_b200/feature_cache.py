"""Device-resident DAgger feature cache.

The reference caches every collected trajectory in LMDB as a msgpack_numpy record
``[obs {key: (T, ...) array}, prev_actions (T,) int64, oracle_actions (T,) int64]`` — optionally cast to float16
(``IL.DAGGER.lmdb_fp16``) — (``dagger_trainer.py:342-362``), reads records back on DataLoader workers, builds the
inflection weights per item (``:189-211``), pads / stacks them in ``collate_fn`` (``:39-114``) and finally casts every
observation to float32 on the device (``:560-566``).  Here the trajectories stay in HBM in their record dtype and a
batch is assembled by one kernel launch per observation key (``vlnce_collate_rows``) plus one for the action /
weight / mask matrices (``vlnce_collate_actions``): no host round trip, no float32 copy of the cache.

* ``TrajectoryCache.append(obs, prev_actions, oracle_actions)`` / ``append_record(blob)`` — what ``txn.put`` stores;
* ``TrajectoryCache.batch(indices, coef)`` — the five outputs of ``collate_fn`` over ``IWTrajectoryDataset`` items,
  observations already float32 (bit-identical to the reference pipeline; with ``fp16=True`` the same float16 rounding
  as ``lmdb_fp16``, including its effect on token ids above 2048 — the reference casts *every* key);
* ``pack_record`` / ``unpack_record`` — the record wire format: msgpack with msgpack_numpy's ndarray extension map
  ``{b"nd": True, b"type": dtype.str, b"kind": b"", b"shape": shape, b"data": raw bytes}`` (msgpack_numpy is not
  installed in this image and unpinned in the reference's requirements.txt; the format below is its published one).
  The LMDB file itself (B+tree pages, transactions) is a storage engine and stays out of scope.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import lib
from .lib import C, ptr


# ------------------------------------------------------------------------------------------ record wire format
def _encode(obj):
    if isinstance(obj, np.ndarray):
        return {b"nd": True, b"type": obj.dtype.str, b"kind": b"", b"shape": list(obj.shape),
                b"data": np.ascontiguousarray(obj).tobytes()}
    if isinstance(obj, (np.bool_, np.number)):
        return {b"nd": False, b"type": obj.dtype.str, b"data": obj.tobytes()}
    return obj


def _decode(obj):
    def get(k):
        return obj[k] if k in obj else obj[k.decode()]

    if b"nd" in obj or "nd" in obj:
        dt = get(b"type")
        dt = np.dtype(dt.decode() if isinstance(dt, bytes) else dt)
        if get(b"nd") is True:
            return np.ndarray(buffer=get(b"data"), dtype=dt, shape=tuple(get(b"shape")))
        return np.frombuffer(get(b"data"), dtype=dt)[0]
    return obj


def pack_record(obs: Dict[str, np.ndarray], prev_actions, oracle_actions, fp16: bool = False) -> bytes:
    """``msgpack_numpy.packb([obs, prev_actions, oracle_actions], use_bin_type=True)`` (dagger_trainer.py:342-362)."""
    import msgpack

    o = {k: (np.asarray(v).astype(np.float16) if fp16 else np.asarray(v)) for k, v in obs.items()}
    rec = [o, np.asarray(prev_actions, dtype=np.int64), np.asarray(oracle_actions, dtype=np.int64)]
    return msgpack.packb(rec, default=_encode, use_bin_type=True)


def unpack_record(blob):
    """``msgpack_numpy.unpackb(blob, raw=False)`` (dagger_trainer.py:171-176) -> [obs, prev_actions, oracle_actions]."""
    import msgpack

    return msgpack.unpackb(bytes(blob), object_hook=_decode, raw=False, strict_map_key=False)


# ------------------------------------------------------------------------------------------ device cache
class _Arena:
    """Rows of one key, back to back, in a device tensor that grows geometrically (copies are device-side)."""

    def __init__(self, device, dtype, row_shape):
        self.device, self.dtype, self.row_shape = device, dtype, tuple(row_shape)
        self.D = int(np.prod(self.row_shape)) if self.row_shape else 1
        self.rows = 0
        self.data = torch.empty((0, self.D), dtype=dtype, device=device)

    def append(self, t: torch.Tensor):
        n = t.shape[0]
        if self.rows + n > self.data.shape[0]:
            grown = torch.empty((max(2 * self.data.shape[0], self.rows + n, 64), self.D), dtype=self.dtype,
                                device=self.device)
            grown[: self.rows].copy_(self.data[: self.rows])
            self.data = grown
        self.data[self.rows: self.rows + n].copy_(t.reshape(n, self.D), non_blocking=True)
        self.rows += n


class TrajectoryCache:
    """What the reference keeps in ``lmdb_features_dir``, resident in HBM."""

    def __init__(self, device, fp16: bool = False):
        self.device = torch.device(device)
        self.fp16 = bool(fp16)
        self.arenas: Dict[str, _Arena] = {}
        self.prev = _Arena(self.device, torch.int64, ())
        self.oracle = _Arena(self.device, torch.int64, ())
        self.row_start: List[int] = []
        self.lengths: List[int] = []

    def __len__(self):
        return len(self.lengths)

    @property
    def nbytes(self) -> int:
        return sum(a.rows * a.D * a.data.element_size() for a in (*self.arenas.values(), self.prev, self.oracle))

    # ---- writing (dagger_trainer.py:334-362)
    def append(self, obs: Dict[str, torch.Tensor], prev_actions, oracle_actions) -> int:
        prev_actions = torch.as_tensor(prev_actions, dtype=torch.int64)
        oracle_actions = torch.as_tensor(oracle_actions, dtype=torch.int64)
        T = prev_actions.shape[0]
        assert oracle_actions.shape[0] == T and T > 0
        if self.arenas and set(obs) != set(self.arenas):
            raise lib.VlnceError(f"trajectory keys {sorted(obs)} differ from the cache's {sorted(self.arenas)}")
        for k, v in obs.items():
            v = torch.as_tensor(v)
            assert v.shape[0] == T, (k, v.shape, T)
            if self.fp16:
                v = v.to(torch.float16)   # lmdb_fp16 casts every key (dagger_trainer.py:344-346)
            elif v.dtype == torch.float64:
                v = v.float()
            if k not in self.arenas:
                self.arenas[k] = _Arena(self.device, v.dtype, v.shape[1:])
            a = self.arenas[k]
            if v.dtype != a.dtype or tuple(v.shape[1:]) != a.row_shape:
                raise lib.VlnceError(f"{k}: {v.dtype}{tuple(v.shape[1:])} does not match the cache's "
                                     f"{a.dtype}{a.row_shape}")
            a.append(v)
        self.row_start.append(self.prev.rows)
        self.lengths.append(T)
        self.prev.append(prev_actions)
        self.oracle.append(oracle_actions)
        return len(self.lengths) - 1

    def append_record(self, blob) -> int:
        obs, prev, oracle = unpack_record(blob)
        return self.append({k: torch.from_numpy(np.copy(v)) for k, v in obs.items()}, np.copy(prev), np.copy(oracle))

    # ---- reading (dagger_trainer.py:189-211 + 39-114 + 560-566)
    def batch(self, indices: Sequence[int], inflection_weight_coef: float = 1.0, use_iw: bool = True):
        """-> (observations {k: (T*N, ...) f32}, prev_actions (T*N,1) i64, not_done_masks (T*N,1) u8,
        corrected_actions (T,N) i64, weights (T,N) f32), T = the longest selected trajectory."""
        lib.require_cuda(self.prev.data, "TrajectoryCache storage")  # no CPU path: the batch is assembled by kernels
        N = len(indices)
        lens = [self.lengths[i] for i in indices]
        T = max(lens)
        dev, st = self.device, lib.stream()
        starts = torch.tensor([self.row_start[i] for i in indices], dtype=torch.int64)
        lengths = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
        row_off = starts.to(dev, non_blocking=True)
        obs = {}
        for k, a in self.arenas.items():
            out = torch.empty((T * N, *a.row_shape), dtype=torch.float32, device=dev)
            offs = (starts * a.D).to(dev, non_blocking=True)
            C.vlnce_collate_rows(ptr(a.data), lib.code(a.dtype), ptr(offs), ptr(lengths), ptr(out), T, N, a.D, 1.0, st)
            obs[k] = out
        prev = torch.empty((T * N, 1), dtype=torch.int64, device=dev)
        actions = torch.empty((T, N), dtype=torch.int64, device=dev)
        weights = torch.empty((T, N), dtype=torch.float32, device=dev)
        masks = torch.empty((T * N, 1), dtype=torch.uint8, device=dev)
        C.vlnce_collate_actions(ptr(self.prev.data), ptr(self.oracle.data), ptr(row_off), ptr(lengths), ptr(prev),
                                ptr(actions), ptr(weights), ptr(masks), T, N,
                                float(inflection_weight_coef) if use_iw else 1.0, st)
        return obs, prev, masks, actions, weights
