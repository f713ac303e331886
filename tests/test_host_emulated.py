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


# ---------------------------------------------------------------------------------------------
# waypoint path: the conv trunks are not emulated — their features come from the oracle's trunks
# ---------------------------------------------------------------------------------------------
def substitute_trunks(pol, case):
    from oracle import waypoint as oracle_wp

    opol = cases.build_policy(oracle_wp, case)
    cases.set_mode(opol, case)

    def rgb_into(observations, out, ld, precision):
        x = observations["rgb"].float().permute(0, 3, 1, 2) / 255.0
        with torch.no_grad():
            f = opol.net.rgb_encoder.cnn(x)[:, :, 0, 0]         # (F, 512): the global pool, broadcast 4x4 (Q2)
        out.view(out.shape[0], -1)[:, : f.shape[1]] = f          # position 0, channels [0, 512)

    def dep_into(observations, out, ld, precision):
        with torch.no_grad():
            f = opol.net.depth_encoder.visual_encoder({"depth": observations["depth"]})  # (F,128,4,4)
        out[:, :, : f.shape[1]] = f.flatten(2).permute(0, 2, 1)

    pol.net.rgb_encoder.features_into = rgb_into
    pol.net.depth_encoder.features_into = dep_into


class _Impl:
    """vlnce_b200 with the trunks of every built WaypointPolicy substituted (test only)."""

    AuxLosses = vlnce_b200.AuxLosses

    def __init__(self, case):
        self.case = case

    @property
    def WaypointPolicy(self):
        case = self.case

        class Wrapped:
            @staticmethod
            def from_config(*a, **kw):
                pol = vlnce_b200.WaypointPolicy.from_config(*a, **kw)
                substitute_trunks(pol, case)
                return pol

        return Wrapped


@pytest.mark.parametrize("name", ["wp_act", "wp_act_u8", "wp_eval"])
def test_waypoint_plan_matches_golden(name):
    case = cases.CASES[name]
    with lib.dry_run(executor=Emulator()):
        got = cases.run_case(_Impl(case), name, aux=vlnce_b200.AuxLosses)
    want = gold(name)
    assert set(got) == {k for k in want if not k.startswith("_")}
    worst = check(got, want, 3e-4)
    print(name, "worst", worst)
