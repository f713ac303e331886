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


def test_side_workloads_keep_the_row_count_and_the_layout():
    """``--workload c2t8`` (8 trajectories x 8 steps) and ``c5`` (RxR): rows per GPU, the time-major (T, N) layout of the
    generated batch and the checker's case description stay consistent."""
    bench = load_bench()
    bench.select_workload("c2t8")
    assert bench.ROWS_PER_GPU == 64 and bench.T_STEPS == 8 and "T=8 N=8/GPU" in bench.WORKLOAD
    case = bench.oracle_case()
    assert case["t"] == 8 and case["n"] == 8
    bench.HW = 32                                      # tiny frames: only the layout is checked here
    obs, prev, masks, actions, weights = bench.make_batch(3, 64)
    assert actions.shape == (8, 8) and weights.shape == (8, 8) and prev.shape == (64, 1) and obs["rgb"].shape[0] == 64
    assert int(masks.view(8, 8)[0].sum()) == 0 and int(masks.view(8, 8)[1:].min()) == 1     # hidden reset at t = 0 only
    bench.select_workload("c5")
    assert bench.ROWS_PER_GPU == 32 and bench.T_STEPS == 1 and bench.RXR_DIM == 768 and bench.instr_key() == "rxr_instruction"
    bench.select_workload("c2")
    assert bench.T_STEPS == 1 and bench.ROWS_PER_GPU == 64
