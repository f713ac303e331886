#!/usr/bin/env python
"""Act-only throughput ("report act-only (T=1, no_grad) and train separately") — a side measurement for
DESIGN.md; the headline benchmark is bench.py (train step, config C2).

    python tools/bench_act.py [--iters 30]

``policy.act(observations, rnn_states, prev_actions, masks, deterministic=True)`` under ``torch.no_grad()``, inputs
resident in HBM (6 rotating batches > L2), CUDA events after warm-up, eager launches (what an evaluation loop does):
  * C2 shapes: CMAPolicy, 64 rows, 256x256 RGB-D, L=80;
  * C1 shapes: Seq2SeqPolicy, 2 rows (the reference's own CPU-runnable case).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    import vlnce_b200
    from oracle import cases  # case table + seeded-weight builder only
    from vlnce_b200 import lib, synth

    dev = torch.device("cuda", 0)
    out = {}
    for label, base, n in (("C2 CMAPolicy act, 64 rows", "cma_act", 64), ("C1 Seq2SeqPolicy act, 2 rows", "seq2seq_act", 2)):
        case = dict(cases.CASES[base], n=n)
        pol = cases.build_policy(vlnce_b200, case).to(dev).eval()
        layers = cases.layers_of(case)
        sets = []
        for i in range(6):
            obs, h, prev, masks = synth.make_act_batch(500 + i, n, hw=256, length=80, layers=layers)
            sets.append(({k: v.to(dev) for k, v in obs.items()}, h.to(dev), prev.to(dev), masks.to(dev)))

        def act(i):
            with torch.no_grad():
                return pol.act(*sets[i % 6], deterministic=True)

        for i in range(5):
            act(i)
        torch.cuda.synchronize()
        c0 = lib.C.vlnce_launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(args.iters):
            act(i)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        out[label] = {"ms_per_call": ms, "policy_steps_per_s": n / (ms * 1e-3),
                      "launches_per_call": (lib.C.vlnce_launch_count() - c0) // args.iters}
        # the same calls with the forward replayed from a CUDA graph (ILPolicy.graph_act)
        pol.graph_act = True
        ref = act(0)
        for i in range(3):
            act(i)
        torch.cuda.synchronize()
        a.record()
        for i in range(args.iters):
            act(i)
        b.record()
        torch.cuda.synchronize()
        msg = a.elapsed_time(b) / args.iters
        pol.graph_act = False
        eager = act(0)
        torch.cuda.synchronize()
        out[label].update({"graph_ms_per_call": msg, "graph_policy_steps_per_s": n / (msg * 1e-3),
                           "graph_equals_eager": bool(torch.equal(ref[0], eager[0]) and torch.equal(ref[1], eager[1]))})
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_act.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
