"""bench.py's ``roofline`` object: the per-launch model (vlnce_b200/roofline.py) evaluated on the real launch plan of
both trunks in dry-run mode (same ``ConvTrunk._conv`` hook as bench.py installs) — flops must equal bench.py's own
accounting, and the assembled object must keep the contract keys whichever resource bounds."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def work():
    import roofline_model

    return roofline_model.collect(64, 256)


def test_plan_flops_match_bench_accounting(work):
    # 107 conv launches per step (53 RGB ResNet-50 + 54 depth ResNet-50 incl. compression); bench.py's own flop count
    # of the same launches, measured on the B200 (profiles/r01_logs/bench_final.log): 728.27 GFLOP per step
    assert len(work) == 107
    assert abs(sum(w["flops"] for w in work) / 1e9 - 728.27) < 0.5
    assert sum(1 for w in work if w["trunk"] == "RGBTrunk") == 53


def test_assemble_contract_keys(work):
    from vlnce_b200 import roofline

    peaks = {"bf16_tflops_sustained": 1369.7, "hbm_gbs": 6534.8}
    flops = 3 * sum(w["flops"] for w in work)
    roof = roofline.assemble(3 * 3.0719, flops, work * 3, peaks, 79.9e6, 321, steps=3)
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"):
        assert k in roof
    # the conv trunks are tensor-core class -> the headline pair is the tensor one, the byte model sits beside
    assert roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s" and roof["peak"] == 1369.7
    assert roof["hbm"]["unit"] == "GB/s" and roof["hbm"]["peak"] == 6534.8
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert abs(roof["tensor"]["achieved"] - 237.0) < 1.0                      # the figure measured in round 1
    lm = roof["layer_model"]
    assert lm["launches"] == 107 and 0.9 < lm["ideal_ms_per_step"] < 1.05
    assert abs(lm["frac_of_ideal"] - lm["ideal_ms_per_step"] / 3.0719) < 1e-9
    json.dumps(roof)                                                          # serialisable
    # a launch list that does not match the timed launches falls back to the tensor-peak figures
    roof = roofline.assemble(3 * 3.0719, flops, work, peaks, None, 321, steps=3)
    assert roof["bound"] == "tensor" and "layer_model" not in roof and roof["unit"] == "TFLOP/s"
