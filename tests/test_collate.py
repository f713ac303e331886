"""collate_fn / inflection weights.

The reference's ``collate_fn`` lives in ``dagger_trainer.py`` whose imports (lmdb, habitat, tensorboard ...) are not
installable here, so the UNMODIFIED function is lifted out of the file with ``ast`` and executed on its own
(build container only); the restatement (``oracle.vlnce.collate_fn``) and the product (``vlnce_b200.collate``) must
reproduce it bit-for-bit.  A digest fixture carries the result to the GPU box."""
import ast
import os
import warnings
from collections import defaultdict

import pytest
import torch

from oracle import reference_import, vlnce as oracle_vlnce
from vlnce_b200 import collate, lib

warnings.filterwarnings("ignore")
GOLD = os.path.join(os.path.dirname(__file__), "golden", "collate.pt")


def trajectories(seed=4, n=5):
    g = torch.Generator().manual_seed(seed)
    batch = []
    for i in range(n):
        t = int(torch.randint(3, 9, (1,), generator=g))
        obs = {"rgb_features": torch.rand(t, 8, 4, 4, generator=g), "depth_features": torch.rand(t, 4, 4, 4, generator=g),
               "instruction": torch.randint(1, 100, (1, 7), generator=g).expand(t, 7).contiguous()}
        actions = torch.randint(0, 4, (t,), generator=g)
        prev = torch.cat([torch.zeros(1, dtype=torch.int64), actions[:-1]])
        batch.append((obs, prev, actions, oracle_vlnce.inflection_weights(actions, 3.2)))
    return batch


def digest(out):
    obs, prev, masks, actions, weights = out
    d = {"obs/" + k: v for k, v in obs.items()}
    d.update(prev=prev, masks=masks, actions=actions, weights=weights)
    return {k: torch.tensor([float(v.double().sum()), float((v.double() * torch.arange(1, v.numel() + 1).view(v.shape)).sum()),
                             float(v.numel())] + list(v.shape)) for k, v in d.items()}


def reference_collate():
    path = os.path.join(reference_import.REFERENCE_ROOT, "vlnce_baselines", "dagger_trainer.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "collate_fn")
    ns = {"torch": torch, "defaultdict": defaultdict, "ObservationsDict": dict}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["collate_fn"]


@pytest.mark.skipif(not reference_import.available(), reason="/root/reference not present")
def test_collate_equals_reference_source():
    ref = reference_collate()(trajectories())
    for impl in (oracle_vlnce.collate_fn, collate.collate_fn):
        got = impl(trajectories())
        assert list(got[0]) == list(ref[0])
        for k in ref[0]:
            assert torch.equal(got[0][k], ref[0][k]), k
        for a, b in zip(got[1:], ref[1:]):
            assert a.dtype == b.dtype and torch.equal(a, b)


def test_collate_golden():
    gold = torch.load(GOLD, weights_only=False)
    for impl in (oracle_vlnce.collate_fn, collate.collate_fn):
        d = digest(impl(trajectories()))
        assert set(d) == set(gold)
        for k in d:
            assert torch.equal(d[k], gold[k]), k


def test_inflection_weights_plan():
    """``vlnce_b200.collate.inflection_weights`` on a padded (T,N) matrix == the per-trajectory reference rule followed
    by collate_fn's zero padding (numpy model of the kernel; the kernel itself: tests/test_gpu_head_ops.py)."""
    from abi_emulator import Emulator

    batch = trajectories()
    _, _, _, actions, weights = collate.collate_fn(batch)
    lengths = torch.tensor([b[1].shape[0] for b in batch])
    with lib.dry_run(executor=Emulator()):
        got = collate.inflection_weights(actions, 3.2, lengths)
    assert torch.equal(got, weights)


if __name__ == "__main__":  # freeze the fixture from the reference's own function (build container)
    torch.save(digest(reference_collate()(trajectories())), GOLD)
    print("frozen", GOLD)
