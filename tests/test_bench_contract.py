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
    case = bench.model_case()
    assert case["policy"] == "CMAPolicy" and case["n"] == 64 and case["t"] == 1 and case["pm"]
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
