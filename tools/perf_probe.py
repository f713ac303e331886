"""Quick device-time probe of policy.act() phases (not the benchmark): CUDA-event timings of
the RGB trunk, depth trunk and the whole act() at the BASELINE batch (N=64, 256x256, L=80)."""
import argparse
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import vlnce_b200  # noqa: E402
from oracle import cases  # noqa: E402  (only for the seeded-policy builder)


def timed(fn, iters, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    case = dict(cases.CASES["cma_act"], n=args.n)
    pol = cases.build_policy(vlnce_b200, case).cuda().eval()
    pol.net.precision = args.precision
    obs, h, prev, masks = cases.inputs(case)
    obs = {k: v.cuda() for k, v in obs.items()}
    h, prev, masks = h.cuda(), prev.cuda(), masks.cuda()
    n = args.n
    out_rgb = torch.zeros(n, 16, 2112, device="cuda")
    out_dep = torch.zeros(n, 16, 192, device="cuda")
    res = {"n": n, "precision": args.precision}
    res["rgb_trunk_ms"] = timed(lambda: pol.net.rgb_encoder.features_into(obs, out_rgb, 2112, args.precision),
                                args.iters, args.warmup)
    res["depth_trunk_ms"] = timed(lambda: pol.net.depth_encoder.features_into(obs, out_dep, 192, args.precision),
                                  args.iters, args.warmup)
    res["act_ms"] = timed(lambda: pol.act(obs, h, prev, masks, deterministic=True), args.iters, args.warmup)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        pol.act(obs, h, prev, masks, deterministic=True)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            pol.act(obs, h, prev, masks, deterministic=True)
    res["act_graph_ms"] = timed(g.replay, args.iters, args.warmup)
    res["rgb_tflops"] = 10.68e9 * n / (res["rgb_trunk_ms"][0] * 1e-3) / 1e12
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/perf_probe.jsonl", "a") as f:
        f.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
