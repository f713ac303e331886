#!/usr/bin/env python
"""BASELINE config C3 on hardware: the ``RecollectTrainer`` teacher-forcing loop (``recollect_trainer.py:128-185``) on
a pre-rendered R2R-shaped rollout buffer resident in HBM — 128 trajectories of T ~ U{20..90} raw sensor frames
(480 x 640 RGB u8 + depth f32), processed as micro-batches of ``IL.batch_size`` trajectories: device collate
(pad / time-major stack / inflection weights) -> ``apply_obs_transforms_batch`` (ResizeShortestEdge 256 +
CenterCropperPerSensor: rgb 224 x 224, depth 256 x 256, one fused kernel per sensor) -> ``_update_agent`` with both
ResNets inside the forward, the GRUs unrolled over T, and gradient accumulation
(``loss_accumulation_scalar = effective_batch_size // batch_size``, optimiser step on the last micro-batch).

    python tools/bench_c3.py [--trajectories 128] [--micro-batch 8] [--parity]

Prints one JSON line: policy steps (valid rows) per second for one effective batch, padded rows, ms per phase.
``--parity`` first runs the SAME loop function at a reduced size against the CPU oracle (restated collate_fn,
obs transforms, update_agent with torch Adam).  A side measurement for DESIGN.md; the headline is bench.py (C2)."""
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

RAW_H, RAW_W, L = 480, 640, 80


def make_trajectories(n, tmin, tmax, dev, seed, raw_hw=(RAW_H, RAW_W)):
    """[(obs {rgb (T,H,W,3) u8, depth (T,H,W,1) f32, instruction (T,L) i64, progress (T,1)}, prev (T,), actions (T,))]"""
    g = torch.Generator().manual_seed(seed)
    gd = torch.Generator(device=dev).manual_seed(seed)
    out = []
    for _ in range(n):
        t = int(torch.randint(tmin, tmax + 1, (1,), generator=g))
        ids = torch.randint(1, 2504, (1, L), generator=g)
        ids[:, int(torch.randint(L // 2, L + 1, (1,), generator=g)):] = 0
        obs = {"rgb": torch.randint(0, 256, (t, raw_hw[0], raw_hw[1], 3), generator=gd, device=dev, dtype=torch.uint8),
               "depth": torch.rand(t, raw_hw[0], raw_hw[1], 1, generator=gd, device=dev),
               "instruction": ids.expand(t, L).contiguous().to(dev),
               "progress": torch.rand(t, 1, generator=g).to(dev)}
        actions = torch.randint(0, 4, (t,), generator=g)
        prev = torch.cat([torch.zeros(1, dtype=torch.int64), actions[:-1]])
        out.append((obs, prev.to(dev), actions.to(dev)))
    return out


def run_effective_batch(step, transforms, trajs, micro, coef=3.2):
    """One optimiser step of the Recollect loop over ``trajs`` in micro-batches of ``micro`` trajectories.
    Returns the per-micro-batch loss triples (device tensors) and (valid rows, padded rows)."""
    from vlnce_b200 import collate
    from vlnce_b200.obs_transforms import apply_obs_transforms_batch

    scalar = (len(trajs) + micro - 1) // micro
    losses, valid, padded = [], 0, 0
    for i in range(0, len(trajs), micro):
        part = trajs[i:i + micro]
        lengths = torch.tensor([a.shape[0] for _, _, a in part])
        obs, prev, masks, actions, _ = collate.collate_fn([(o, p, a, torch.zeros(a.shape[0], device=a.device)) for o, p, a in part])
        weights = collate.inflection_weights(actions, coef, lengths)
        obs = apply_obs_transforms_batch(obs, transforms)
        last = i + micro >= len(trajs)
        if step.skip_padded_rows:   # real rows from the lengths the collate step knows (no device read)
            tmax, n = int(lengths.max()), len(part)
            real = torch.arange(tmax).view(tmax, 1) < lengths.view(1, n)
            step.valid_rows_hint = real.reshape(-1).nonzero().reshape(-1).pin_memory()
        out = step.update_agent(obs, prev, masks, actions, weights, step_grad=last, loss_accumulation_scalar=scalar,
                                return_tensor=True)
        losses.append(out.clone())
        valid += int(lengths.sum())
        padded += int(lengths.max()) * len(part)
    return losses, (valid, padded)


def parity(dev, skip=True):
    """The same loop at 4 trajectories x T in [3, 6], micro-batch 2, fp32 trunks, against the CPU oracle."""
    import vlnce_b200
    from oracle import cases, obs_transforms as oracle_ot, vlnce as oracle_vlnce
    from vlnce_b200 import obs_transforms as ot
    from vlnce_b200.il_step import ILTrainerStep

    case = dict(cases.CASES["cma_train_pixels"], rgb_hw=224)
    trajs = make_trajectories(4, 3, 6, dev, seed=5)
    pol = cases.build_policy(vlnce_b200, case).to(dev).eval()
    pol.net.precision = "fp32"
    vlnce_b200.AuxLosses.activate()
    step = ILTrainerStep(pol, lr=2.5e-4)
    step.skip_padded_rows = skip
    tf = [ot.ResizeShortestEdge(256), ot.CenterCropperPerSensor([("rgb", (224, 224)), ("depth", (256, 256))])]
    got, _ = run_effective_batch(step, tf, trajs, 2)
    got = torch.stack(got).cpu()
    vlnce_b200.AuxLosses.deactivate()
    # oracle: restated transforms + collate_fn + update_agent (torch Adam), same accumulation pattern
    opol = cases.build_policy(oracle_vlnce, case).eval()
    opt = torch.optim.Adam([p for p in opol.parameters() if p.requires_grad], lr=2.5e-4)
    otf = [oracle_ot.ResizeShortestEdge(256), oracle_ot.CenterCropperPerSensor([("rgb", (224, 224)), ("depth", (256, 256))])]
    oracle_vlnce.AuxLosses.activate()
    want = []
    cpu = [({k: v.cpu() for k, v in o.items()}, p.cpu(), a.cpu()) for o, p, a in trajs]
    for i in range(0, 4, 2):
        b = oracle_vlnce.collate_fn([(o, p, a, oracle_vlnce.inflection_weights(a, 3.2)) for o, p, a in cpu[i:i + 2]])
        obs = oracle_ot.apply_obs_transforms_batch(dict(b[0]), otf)
        want.append(oracle_vlnce.update_agent(opol, opt, obs, *b[1:], step_grad=i + 2 >= 4, loss_accumulation_scalar=2,
                                              aux=oracle_vlnce.AuxLosses))
    oracle_vlnce.AuxLosses.deactivate()
    want = torch.tensor(want)
    ref = dict(opol.named_parameters())
    num = den = 0.0
    for k, p in pol.named_parameters():
        if p.requires_grad:
            d0 = (p.detach().cpu() - ref[k].detach()).double()
            num += float(d0.pow(2).sum())
            den += float(ref[k].detach().double().pow(2).sum())
    return {"losses_max_rel": float(((got - want).abs() / want.abs().clamp_min(1e-6)).max()),
            "params_after_step_rel_l2": (num / den) ** 0.5, "micro_batches": 2, "trajectories": 4,
            "oracle": "oracle/vlnce.py + oracle/obs_transforms.py, torch-CPU fp32, same seeded weights and frames"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trajectories", type=int, default=128)
    ap.add_argument("--micro-batch", type=int, default=8)
    ap.add_argument("--tmin", type=int, default=20)
    ap.add_argument("--tmax", type=int, default=90)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--no-skip-padded", action="store_true",
                    help="push the padded rows of every micro-batch through the conv trunks too, as the reference does "
                         "(default: rows with loss weight 0 bypass the frozen eval-mode trunks, ILTrainerStep.skip_padded_rows)")
    args = ap.parse_args()
    import vlnce_b200
    from vlnce_b200 import lib, obs_transforms as ot, synth
    from vlnce_b200.il_step import ILTrainerStep

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    par = parity(dev, skip=not args.no_skip_padded) if args.parity else None
    pol = synth.build_cma_policy(wseed=13, hw=256, rgb_hw=224).to(dev).eval()
    vlnce_b200.AuxLosses.activate()
    step = ILTrainerStep(pol, lr=2.5e-4)
    step.skip_padded_rows = not args.no_skip_padded
    tf = [ot.ResizeShortestEdge(256), ot.CenterCropperPerSensor([("rgb", (224, 224)), ("depth", (256, 256))])]
    trajs = make_trajectories(args.trajectories, args.tmin, args.tmax, dev, seed=9)
    hbm_gb = sum(v.numel() * v.element_size() for o, _, _ in trajs for v in o.values()) / 1e9
    run_effective_batch(step, tf, trajs[: 2 * args.micro_batch], args.micro_batch)   # warm-up: plans, tile selection
    torch.cuda.synchronize()
    ms = 0.0
    rows = (0, 0)
    c0 = lib.C.vlnce_launch_count()
    for _ in range(args.iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        losses, rows = run_effective_batch(step, tf, trajs, args.micro_batch)
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    launches = int(lib.C.vlnce_launch_count() - c0) // args.iters
    ms /= args.iters
    line = {"metric": "policy steps/sec (C3: Recollect teacher-forcing loop, 128 trajectories, raw 480x640 frames)",
            "value": rows[0] / (ms * 1e-3), "unit": "policy steps/s (valid rows)", "n_gpus": 1,
            "ms_per_effective_batch": ms, "valid_rows": rows[0], "padded_rows": rows[1],
            "padded_rows_per_s": rows[1] / (ms * 1e-3), "launches_per_effective_batch": launches,
            "last_losses": [float(x) for x in losses[-1]], "rollout_buffer_gb_in_hbm": hbm_gb,
            "config": {"trajectories": args.trajectories, "T": [args.tmin, args.tmax], "micro_batch": args.micro_batch,
                       "loss_accumulation_scalar": (args.trajectories + args.micro_batch - 1) // args.micro_batch,
                       "raw_frames": "480x640 rgb u8 + depth f32", "transforms": "ResizeShortestEdge(256) + CenterCrop rgb 224 / depth 256",
                       "precision": "rgb trunk bf16 / depth trunk fp16 tcgen05, head fp32",
                       "padded_rows_skip_the_trunks": bool(step.skip_padded_rows)}}
    if par is not None:
        line["parity"] = par
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", "bench_c3.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
