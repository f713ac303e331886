#!/usr/bin/env python
"""BASELINE config C4 on hardware: ``DDPPOWaypointTrainer``-shaped loop, data parallel over the GPUs of one box —
per rank N = 4 synthetic environments x T = 16 steps of ``WaypointPolicy.act`` (rollout collection into the
device-resident ``ActionDictRolloutStorage``), GAE, then ``WDDPPOStep.update`` (2 epochs x 4 recurrent mini-batches,
ONE NCCL all-reduce of the flat gradient per mini-batch, clip, Adam) — ``ddppo_waypoint_trainer.py:283-306,497-560``,
``ddppo_alg.py:37-149``.  The simulator is stubbed: random next observations, reward N(0,1), done with p = 0.05.

    python tools/bench_c4.py [--iters 5]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c4.py

Rank 0 prints one JSON line: env-steps/s over all ranks (collection + update), ms per phase, launches per phase.
A side measurement for DESIGN.md (the headline benchmark is bench.py / config C2)."""
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
    ap.add_argument("--envs", type=int, default=4)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-graph-act", action="store_true", help="launch the forward of every act() eagerly")
    ap.add_argument("--no-visual-cache", action="store_true",
                    help="re-run both ResNets on the stored frames in the PPO update, as the reference does (default: the "
                         "frozen trunks' features of every collected step are kept in the rollout storage)")
    args = ap.parse_args()
    import torch.distributed as dist

    import vlnce_b200
    from oracle import cases, waypoint as oracle_wp   # case table / observation-space description only (side tool)
    from vlnce_b200 import lib, synth
    from vlnce_b200.ppo_step import ActionDictRolloutStorage, WDDPPOStep, select_history_frames
    from vlnce_b200.waypoint_policy import VISUAL_CACHE_KEYS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, T, L = args.envs, args.steps, 80
    case = dict(cases.CASES["wp_eval"], n=N, t=T, length=L)
    torch.manual_seed(1 + rank * N)   # ddppo_waypoint_trainer.py:330-336
    pol = cases.build_policy(vlnce_b200, case).to(dev)
    pol.net.precision = args.precision
    cases.set_mode(pol, case)
    pol.graph_act = not args.no_graph_act
    agent = WDDPPOStep(pol)
    space = oracle_wp.storage_obs_space(length=L)
    storage = ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=2)
    storage.to(dev)
    if not args.no_visual_cache:
        storage.enable_visual_cache(pol)
    gen = torch.Generator().manual_seed(100 + rank)
    # a pool of pre-rendered observations (the simulator stub): new frames every step without host RNG in the loop
    pool = [{k: v.to(dev) for k, v in synth.make_waypoint_rows(gen, N, length=L).items()} for _ in range(6)]
    instr = pool[0]["instruction"]
    for p in pool:
        p["instruction"] = instr   # an episode keeps its instruction

    def collect():
        """One rollout: T act() steps + insert (``_collect_rollout_step``), then the bootstrap value + GAE."""
        pol.eval()
        for k in pool[0]:
            storage.observations[k][0].copy_(pool[0][k])
        for t in range(T):
            obs = {k: v[storage.step] for k, v in storage.observations.items() if k not in VISUAL_CACHE_KEYS}
            with torch.no_grad():
                (value, actions, elems, _modes, _vars, logp, h, _dist) = pol.act(
                    obs, storage.recurrent_hidden_states[storage.step],
                    {k: v[storage.step] for k, v in storage.prev_actions.items()}, storage.masks[storage.step])
            if storage.visual_cache:
                storage.store_visual_features(pol.visual_features())
            nxt = dict(pool[(t + 1) % len(pool)])
            hist = select_history_frames(obs, elems["pano"], 12)      # frames the agent moved towards (:190-199)
            nxt["rgb_history"], nxt["depth_history"] = hist["rgb"].float(), hist["depth"]
            rewards = torch.randn(N, 1, device=dev)
            masks = (torch.rand(N, 1, device=dev) > 0.05).to(torch.uint8)
            storage.insert(nxt, h, elems, logp, value, rewards, masks)
        with torch.no_grad():
            obs = {k: v[storage.step] for k, v in storage.observations.items() if k not in VISUAL_CACHE_KEYS}
            next_value = pol.get_value(obs, storage.recurrent_hidden_states[storage.step],
                                       {k: v[storage.step] for k, v in storage.prev_actions.items()},
                                       storage.masks[storage.step])
        storage.compute_returns(next_value, True, 0.99, 0.95)

    def update():
        pol.train()
        pol.net.rgb_encoder.eval()
        pol.net.depth_encoder.eval()
        out = agent.update(storage)
        storage.after_update()
        return out

    def timed(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = lib.C.vlnce_launch_count()
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b), int(lib.C.vlnce_launch_count() - c0), r

    for _ in range(args.warmup):
        collect()
        update()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms_c = ms_u = 0.0
    lc = lu = 0
    losses = None
    for _ in range(args.iters):
        t1, l1, _ = timed(collect)
        t2, l2, losses = timed(update)
        ms_c += t1
        ms_u += t2
        lc, lu = l1, l2
    total = torch.tensor([ms_c + ms_u], device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    if rank == 0:
        steps = N * T * world * args.iters
        line = {"metric": "env steps/sec (C4: DD-PPO waypoint trainer loop, 4 envs x 16 steps per rank, synthetic envs)",
                "value": steps / (total.item() * 1e-3), "unit": "env steps/s", "n_gpus": world,
                "collect_ms": ms_c / args.iters, "update_ms": ms_u / args.iters, "launches_collect": lc, "launches_update": lu,
                "precision": args.precision, "visual_cache": bool(storage.visual_cache), "graph_act": bool(pol.graph_act), "losses_last_update": [float(x) for x in losses],
                "config": {"envs_per_rank": N, "rollout_steps": T, "ppo_epoch": 2, "num_mini_batch": 4,
                           "frames_per_env_step": 13, "rgb": "12x224x224x3 f32", "depth": "12x256x256x1 f32", "L": L}}
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(line, open(os.path.join(ROOT, "gpurun_out", f"bench_c4_n{world}.json"), "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
