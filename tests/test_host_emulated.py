"""Host launch plans executed on host memory through the numpy model of the head kernels
(``tests/abi_emulator.py``) and compared with the golden vectors frozen from the unmodified
reference: pointer arithmetic, leading dimensions, buffer reuse and the hand-written backward are
checked numerically without a GPU.  (The CUDA kernels themselves are checked by the -m gpu tests.)"""
import os
import warnings

import pytest
import torch

import vlnce_b200
from oracle import cases
from vlnce_b200 import lib

from abi_emulator import Emulator

warnings.filterwarnings("ignore")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def check(got, want, tol, skip=()):
    worst = ("", 0.0)
    for k, v in got.items():
        if k in skip:
            continue
        g = want[k]
        if v.dtype in (torch.int64, torch.uint8, torch.bool):
            assert torch.equal(v, g), k
            continue
        scale = float(g.abs().max()) + 1e-12
        err = float((v.double() - g.double()).abs().max()) / scale
        if err > worst[1]:
            worst = (k, err)
        assert err < tol or float((v.double() - g.double()).abs().max()) < 1e-7, (k, err)
    return worst


@pytest.mark.parametrize("name", ["cma_train_features", "cma_train_features_nopm", "seq2seq_train_features"])
def test_il_train_plan_matches_golden(name):
    with lib.dry_run(executor=Emulator()):
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses)
    worst = check(got, gold(name), 2e-4)
    print(name, "worst", worst)
