#!/usr/bin/env python
"""Library baseline on the B200 ("the reference on the same B200 with stock PyTorch CUDA eager as the library
baseline to beat"): the oracle's restatement of the reference model — plain torch modules, i.e. cuDNN convolutions / RNNs
and cuBLAS GEMMs — run in CUDA eager mode on the bench workload (config C2: CMAPolicy forward + DAgger loss + backward
+ Adam, T=1, 64 rows, 256x256 RGB-D, L=80, frozen trunks with eval-mode norms), in three arithmetic modes:

    fp32     TF32 off everywhere (the true-fp32 parity reference)
    tf32     torch defaults + TF32 matmuls allowed (cuDNN conv / RNN TF32, as stock PyTorch would run the reference)
    bf16     torch.autocast(bfloat16) around the forward
    bf16_cl  the same with the conv trunks' weights in channels_last (cuDNN's NHWC tensor-core kernels, no layout
             transposes; the simulator's NHWC images permuted to NCHW are already channels_last-strided)

This is TEST INFRASTRUCTURE (it executes oracle/, which only tests / smoke / the bench's CPU legs may do) and a side
measurement for DESIGN.md — not a bench arm.  CUDA events, warm-up, 6 rotating input batches (> L2).

    python tests/library_baseline.py [--rows 64] [--steps 20] [--warmup 5]
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
import torch.nn.functional as F  # noqa: E402


def step_fn(pol, opt, aux, batch, autocast):
    obs, prev, masks, actions, weights = batch
    t, n = actions.shape
    h0 = torch.zeros(n, pol.net.num_recurrent_layers, 512, device=actions.device)
    aux.clear()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        dist = pol.build_distribution(obs, h0, prev, masks)
    logits = dist.logits.float().view(t, n, -1)
    ce = F.cross_entropy(logits.permute(0, 2, 1), actions, reduction="none")
    loss = ((weights * ce).sum(0) / weights.sum(0)).mean() + aux.reduce((weights > 0).view(-1))
    loss.backward()
    opt.step()
    opt.zero_grad()
    return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--modes", default="fp32,tf32,bf16,bf16_cl")
    ap.add_argument("--no-act", action="store_true")
    ap.add_argument("--out", default="library_baseline.json")
    args = ap.parse_args()
    from oracle import cases, vlnce as oracle_vlnce
    from vlnce_b200 import synth

    assert torch.cuda.is_available()
    dev = torch.device("cuda", 0)
    case = dict(cases.CASES["cma_train_pixels"], t=1, n=args.rows)
    sets = []
    for i in range(6):
        b = synth.make_train_batch(1000 + i, 1, args.rows, hw=256, length=80)
        sets.append(tuple({k: v.to(dev) for k, v in x.items()} if isinstance(x, dict) else x.to(dev) for x in b))
    out = {"workload": f"C2: oracle CMAPolicy (torch modules: cuDNN / cuBLAS), CUDA eager, fwd+loss+bwd+Adam, T=1 N={args.rows}, "
                       "256x256 RGB-D, L=80, frozen trunks / eval-mode norms", "modes": {}}
    aux = oracle_vlnce.AuxLosses
    for mode in args.modes.split(","):
        try:
            tf32 = mode != "fp32"
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            pol = cases.build_policy(oracle_vlnce, case).to(dev).eval()
            # cuDNN's RNN backward needs train mode; the frozen trunks keep eval-mode norms (the benchmarked config)
            pol.train()
            pol.net.rgb_encoder.cnn.eval()
            pol.net.depth_encoder.visual_encoder.eval()
            if mode.endswith("_cl"):
                pol.net.rgb_encoder.to(memory_format=torch.channels_last)
                pol.net.depth_encoder.to(memory_format=torch.channels_last)
            opt = torch.optim.Adam([p for p in pol.parameters() if p.requires_grad], lr=2.5e-4)
            aux.activate()
            for i in range(args.warmup):
                step_fn(pol, opt, aux, sets[i % 6], mode.startswith("bf16"))
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(args.steps):
                loss = step_fn(pol, opt, aux, sets[i % 6], mode.startswith("bf16"))
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / args.steps
            out["modes"][mode] = {"ms_per_step": ms, "policy_steps_per_s": args.rows / (ms * 1e-3), "loss": float(loss)}
        except Exception as e:  # keep the other modes
            out["modes"][mode] = {"error": f"{type(e).__name__}: {e}"[:300]}
            continue
        finally:
            aux.deactivate()  # evaluation runs without the auxiliary-loss registry (base_il_trainer.py:278-280)
        try:  # the two frozen trunks alone (what RGBTrunk / DepthTrunk replace)
            for name, enc in (("rgb_trunk_ms", pol.net.rgb_encoder), ("depth_trunk_ms", pol.net.depth_encoder)):
                def run(i):
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode.startswith("bf16")):
                        return enc(sets[i % 6][0])
                for i in range(3):
                    run(i)
                torch.cuda.synchronize()
                a.record()
                for i in range(args.steps):
                    run(i)
                b.record()
                torch.cuda.synchronize()
                out["modes"][mode][name] = a.elapsed_time(b) / args.steps
        except Exception as e:
            out["modes"][mode]["trunk_error"] = f"{type(e).__name__}: {e}"[:300]
        if args.no_act:
            continue
        try:
            # act-only (no_grad), the same shapes: compare with tools/bench_act.py
            h0 = torch.zeros(args.rows, pol.net.num_recurrent_layers, 512, device=dev)

            def act(i):
                obs, prev, masks = sets[i % 6][0], sets[i % 6][1], sets[i % 6][2]
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode.startswith("bf16")):
                    return pol.act(obs, h0, prev, masks, deterministic=True)

            for i in range(3):
                act(i)
            torch.cuda.synchronize()
            a.record()
            for i in range(args.steps):
                act(i)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / args.steps
            out["modes"][mode].update(act_ms_per_call=ms, act_policy_steps_per_s=args.rows / (ms * 1e-3))
        except Exception as e:
            out["modes"][mode]["act_error"] = f"{type(e).__name__}: {e}"[:300]
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", args.out), "w"), indent=1)


if __name__ == "__main__":
    main()
