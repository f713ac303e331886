#!/usr/bin/env python
"""Throughput of the waypoint path (BASELINE config C4 shapes) on one GPU — a
side measurement for DESIGN.md, not the headline benchmark (that is bench.py / config C2).

    python tools/bench_waypoint.py [--envs 4] [--steps 16] [--iters 10] [--cpu-seconds 10]

Measures, with CUDA events after warm-up, inputs resident in HBM:
  * ``act``      rollout step: WaypointPolicy.act on N envs (12 pano frames + history frame each);
  * ``update``   one WDDPPO update: GAE + ppo_epoch x num_mini_batch fused mini-batches over a (T, N) rollout
                 (evaluate_actions + PPO loss + backward + clip + Adam), rows = T*N env-steps per epoch;
and the CPU oracle (oracle/waypoint.py, torch-CPU fp32, the restated reference loop) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402


def fill(storage, T, N, synth, length=80):
    g = torch.Generator().manual_seed(3)
    first = synth.make_waypoint_rows(g, N, length=length)
    for k in storage.observations:
        storage.observations[k][0].copy_(first[k])
    for _ in range(T):
        storage.insert(synth.make_waypoint_rows(g, N, length=length), torch.randn(N, 2, 256, generator=g) * 0.3,
                       synth.make_waypoint_actions(g, N), -2 + 0.2 * torch.randn(N, 1, generator=g),
                       torch.randn(N, 1, generator=g) * 0.3, torch.randn(N, 1, generator=g),
                       (torch.rand(N, 1, generator=g) > 0.05).to(torch.uint8))
    return torch.randn(N, 1, generator=g) * 0.3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    import vlnce_b200
    from oracle import cases, waypoint as oracle_wp
    from vlnce_b200 import lib, synth
    from vlnce_b200.ppo_step import ActionDictRolloutStorage, WDDPPOStep

    N, T = args.envs, args.steps
    case = dict(cases.CASES["wp_eval"], n=N, t=T, length=80)
    dev = torch.device("cuda", 0)
    pol = cases.build_policy(vlnce_b200, case).to(dev)
    cases.set_mode(pol, case)
    space = oracle_wp.storage_obs_space(length=80)
    st = ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=2)
    nv = fill(st, T, N, synth)
    st.to(dev)
    agent = WDDPPOStep(pol, ppo_epoch=2, num_mini_batch=min(4, N))

    def ev_time(fn, iters):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    obs0 = {k: v[0] for k, v in st.observations.items()}
    prev0 = {k: v[0] for k, v in st.prev_actions.items()}

    def act():
        pol.act(obs0, st.recurrent_hidden_states[0], {k: v.clone() for k, v in prev0.items()}, st.masks[0])

    def update():
        st.compute_returns(nv.to(dev), True, 0.99, 0.95)
        agent.update(st)

    for _ in range(3):
        act()
    update()
    c0 = lib.C.vlnce_launch_count()
    act_ms = ev_time(act, args.iters)
    c1 = lib.C.vlnce_launch_count()
    upd_ms = ev_time(update, max(2, args.iters // 3))
    c2 = lib.C.vlnce_launch_count()
    out = {
        "workload": f"C4 shapes: WaypointPolicy (1-wpn-cc), {N} envs x {T} steps, 12x224x224x3 rgb f32 + 12x256x256 depth "
                    "per env-step, L=80, bf16/fp16 trunks, fp32 head",
        "act": {"ms": act_ms, "env_steps_per_s": N / (act_ms * 1e-3), "frames_per_call": N * 13,
                "launches_per_call": (c1 - c0) // args.iters},
        "update": {"ms": upd_ms, "env_steps_per_s": 2 * T * N / (upd_ms * 1e-3),
                   "note": "ppo_epoch=2 x num_mini_batch: every stored env-step is evaluated twice per update",
                   "launches_per_update": (c2 - c1) // max(2, args.iters // 3)},
    }
    # ---- CPU oracle, bounded sample (one mini-batch of the same rollout)
    if args.cpu_seconds > 0:
        opol = cases.build_policy(oracle_wp, case)
        cases.set_mode(opol, case)
        ost = oracle_wp.ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=2)
        nv2 = fill(ost, T, N, synth)
        ost.compute_returns(nv2, True, 0.99, 0.95)
        opt = torch.optim.Adam([p for p in opol.parameters() if p.requires_grad], lr=2e-4, eps=1e-5)
        adv = oracle_wp.advantages_of(ost, False)
        sample = next(iter(ost.recurrent_generator(adv, min(4, N))))
        rows = sample[6].shape[0]
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < args.cpu_seconds or n == 0:
            loss, _ = oracle_wp.wddppo_losses(opol, sample)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(opol.parameters(), 0.2)
            opt.step()
            n += 1
        dt = time.perf_counter() - t0
        out["cpu_oracle_update"] = {"env_steps_per_s": rows * n / dt, "cores": torch.get_num_threads(),
                                    "sample": f"{n} mini-batches of {rows} rows, oracle/waypoint.py torch-CPU fp32"}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_waypoint.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
