"""bench.py host-side pieces (the timed legs need a GPU): metric / workload follow BASELINE.json, the nvidia-smi clock
parser, the host-core count used for the CPU baseline."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_metric_and_workload_follow_baseline():
    bench = load_bench()
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.METRIC.startswith("policy steps/sec (batch 64") and "policy steps/sec" in base["metric"]
    assert bench.ROWS_PER_GPU == 64 and bench.HW == 256 and bench.LENGTH == 80     # configs[1] (C2)
    case = bench.oracle_case()
    assert case["policy"] == "CMAPolicy" and case["n"] == 64 and case["t"] == 1 and case["pm"]
    # the product arm builds its policy from the package alone (no checker import on that path) ...
    src = open(os.path.join(ROOT, "bench.py")).read()
    main_src = src[src.index("def main():"):]
    assert "from oracle" not in main_src and "os._exit" not in src
    # ... and it is the same seeded policy the checker rebuilds on its side
    import torch
    import vlnce_b200
    from oracle import cases
    from vlnce_b200 import synth

    a = synth.build_cma_policy(wseed=bench.WSEED, hw=bench.HW).state_dict()
    b = cases.build_policy(vlnce_b200, case).state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert "C2" in bench.WORKLOAD and "DAgger" in bench.WORKLOAD


def test_clock_parser_and_host_cores():
    bench = load_bench()
    lines = ["1965, 1965, 712.3, Not Active, Not Active, Not Active, Not Active",
             "1950, 1965, 901.0, Not Active, Not Active, Not Active, Active",
             "210, 1965, 150.0, Not Active, Not Active, Not Active, Not Active", "garbage"]
    c = bench.summarize_clocks(lines)
    assert c["sm_max_mhz"] == 1965.0 and c["sm_mhz"] == 1957.5 and c["reasons"] == ["sw_power_cap"]
    assert bench.summarize_clocks([]) == {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    assert 1 <= bench.host_cores() <= (os.cpu_count() or 1)
