"""Pin the oracle against the UNMODIFIED reference files (build container only —
/root/reference does not exist on the GPU box, where this module skips)."""
import warnings

import pytest
import torch

from oracle import cases, reference_import, vlnce

pytestmark = pytest.mark.skipif(not reference_import.available(),
                                reason="/root/reference not present")
warnings.filterwarnings("ignore")


@pytest.mark.parametrize("name", ["cma_act_bntrain", "cma_train_features", "seq2seq_act", "seq2seq_train_features",
                                  "wp_act", "wp_eval", "wp_act_hpn", "wp_eval_hpn", "wp_act_dd", "wp_eval_d_"])
def test_oracle_equals_reference_live(name):
    ref = reference_import.load()
    a = cases.run_case(ref, name, aux=ref.AuxLosses)
    b = cases.run_case(cases.oracle_impl(cases.CASES[name]), name, aux=vlnce.AuxLosses)
    assert set(a) == set(b)
    for k in a:
        tol = 2e-5 if name.startswith("wp_") else 1e-6  # waypoint: einsum vs bmm reduction order
        assert torch.equal(a[k], b[k]) or torch.allclose(a[k], b[k], rtol=tol, atol=tol * 0.1 * float(a[k].abs().max())
                                                         + 1e-7), k


def test_state_dict_keys_match_reference():
    ref = reference_import.load()
    for name in ("cma_act", "seq2seq_act", "cma_rxr_act", "wp_act", "wp_act_dc", "wp_act_dd"):
        case = cases.CASES[name]
        r = cases.build_policy(ref, case).state_dict()
        o = cases.build_policy(cases.oracle_impl(case), case).state_dict()
        assert list(r) == list(o)
        assert all(r[k].shape == o[k].shape for k in r)
