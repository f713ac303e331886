#!/usr/bin/env python
"""Benchmark of the VLN-CE policy hot path (BASELINE.json metric: policy steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one ``_update_agent`` call (``common/base_il_trainer.py:134-180``) on the
BASELINE config C2 batch: CMAPolicy (bi-LSTM instruction encoder, progress monitor), T=1,
64 trajectories per GPU, synthetic 256x256 RGB (u8) + depth (f32) + 80-token instructions:
obs -> both conv trunks -> CMA head -> inflection-weighted CE + progress loss -> head backward
-> (grad all-reduce) -> Adam.  ``value`` = policy steps (rows) per second over all GPUs.

Prints ONE JSON line (see the task contract): value (inputs resident in HBM), e2e (host
buffers through the public ILTrainerStep API, H2D/D2H inside the timed region), roofline of
the dominant kernel (tcgen05 conv tiles), cpu_baseline (the CPU oracle port on the host
cores), clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402

ROWS_PER_GPU = 64
HW = 256
LENGTH = 80
METRIC = "policy steps/sec (batch 64, 256x256 RGB-D, 80-tok instr)"
WORKLOAD = ("C2: CMAPolicy fwd+DAgger loss+bwd+Adam, T=1 N=64/GPU, rgb u8 256x256x3, depth f32 256x256x1, "
            "L=80, progress monitor on, conv trunks frozen/eval-mode norms")


WSEED = 13  # seeded "trained-like" weights (vlnce_b200.synth.seeded_state_dict); the parity case cma_train_pixels uses the same
RGB_HW, RXR_DIM, ACTIONS, COEF = 256, 0, 4, 3.2
T_STEPS = 1   # time steps per trajectory in a batch (rows = T_STEPS x trajectories, time-major as collate_fn stacks them)
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (default)
    "c2": dict(rows=64, hw=256, rgb_hw=256, length=80, rxr=0, actions=4, coef=3.2),
    # BASELINE.json configs[4]: RxR-shaped multilingual-BERT features (512 tokens x 768), 6 actions, frames after the
    # RxR observation transforms (rgb 224x224, depth 256x256: rxr_cma_en.yaml:28-30), 256 rows over 8 GPUs = 32 / GPU
    "c5": dict(rows=32, hw=256, rgb_hw=224, length=512, rxr=768, actions=6, coef=1.9),
    # SURVEY 8d's training variant of C2: the same 64 rows as 8 trajectories x 8 time steps (sequence GRUs with mask
    # resets, per-trajectory loss normalisation) — side measurement
    "c2t8": dict(rows=64, hw=256, rgb_hw=256, length=80, rxr=0, actions=4, coef=3.2, t=8),
}


def select_workload(name):
    global ROWS_PER_GPU, HW, RGB_HW, LENGTH, RXR_DIM, ACTIONS, COEF, METRIC, WORKLOAD, T_STEPS
    w = WORKLOADS[name]
    T_STEPS = w.get("t", 1)
    ROWS_PER_GPU, HW, RGB_HW, LENGTH = w["rows"], w["hw"], w["rgb_hw"], w["length"]
    RXR_DIM, ACTIONS, COEF = w["rxr"], w["actions"], w["coef"]
    if name == "c2t8":
        METRIC = "policy steps/sec (batch 64 = 8 trajectories x 8 steps, 256x256 RGB-D, 80-tok instr)"
        WORKLOAD = WORKLOAD.replace("T=1 N=64/GPU", "T=8 N=8/GPU")
    if name == "c5":
        METRIC = "policy steps/sec (RxR: batch 32/GPU, 224x224 RGB + 256x256 depth, 512x768 BERT instruction features)"
        WORKLOAD = ("C5: CMAPolicy fwd+DAgger loss+bwd+Adam, T=1 N=32/GPU, rgb u8 224x224x3, depth f32 256x256x1, "
                    "rxr_instruction f32 512x768, 6 actions, progress monitor on, conv trunks frozen/eval-mode norms")


def make_batch(seed, rows):
    from vlnce_b200 import synth

    return synth.make_train_batch(seed, T_STEPS, rows // T_STEPS, hw=HW, rgb_hw=RGB_HW, length=LENGTH, num_actions=ACTIONS, coef=COEF,
                                  rxr_dim=RXR_DIM)


def instr_key():
    return "rxr_instruction" if RXR_DIM else "instruction"


def oracle_case(rows=None):
    rows = ROWS_PER_GPU if rows is None else rows
    """CHECKER side only (reference arm, cpu_baseline leg, parity block): the oracle's description of the same policy."""
    from oracle import cases

    case = dict(cases.CASES["cma_train_pixels"], t=T_STEPS, n=rows // T_STEPS)
    assert case["wseed"] == WSEED
    if RXR_DIM:
        case.update(rgb_hw=RGB_HW, length=LENGTH, rxr=RXR_DIM, actions=ACTIONS, coef=COEF)
    return case


# ----------------------------------------------------------------------------------------
def clocks_sampler(stop, out):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    dev = os.environ.get("LOCAL_RANK", "0")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                              "-i", dev], stdout=subprocess.PIPE, text=True)
    except Exception:
        return
    t = threading.Thread(target=lambda: [out.append(line) for line in p.stdout], daemon=True)
    t.start()
    stop.wait()
    p.terminate()
    t.join(timeout=2)


def summarize_clocks(lines):
    sm, mx, reasons = [], [], set()
    for ln in lines:
        f = [x.strip() for x in ln.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0]))
            mx.append(float(f[1]))
        except ValueError:
            continue
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    busy = [x for x in sm if x > 0.5 * max(sm)] or sm
    return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def host_cores() -> int:
    """CPU threads this process may really use: affinity mask capped by the cgroup quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


# ----------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores: the oracle port
    (``kind: port`` — /root/reference does not exist on the GPU box), bounded sample."""
    if rank != 0:
        return
    from oracle import cases, vlnce as oracle_vlnce

    from vlnce_b200 import synth

    rows = ROWS_PER_GPU  # the full 64-row step: same config as the GPU arm
    torch.set_num_threads(host_cores())
    case = oracle_case(rows)
    pol = cases.build_policy(oracle_vlnce, case).eval()
    opt = torch.optim.Adam(pol.parameters(), lr=2.5e-4)
    obs, prev, masks, actions, weights = make_batch(1000, rows)
    oracle_vlnce.AuxLosses.activate()

    def step():
        return oracle_vlnce.update_agent(pol, opt, obs, prev, masks, actions, weights)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = rows * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "policy steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_step": rows},
        "cpu_baseline": {"value": val, "unit": "policy steps/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} full {rows}-row steps (+{args.warmup} warm-up), oracle/vlnce.py on torch {torch.__version__} CPU fp32"},
        "e2e": {"value": val, "unit": "policy steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
def cpu_baseline_leg(seconds=15.0):
    from oracle import cases, vlnce as oracle_vlnce
    from vlnce_b200 import synth

    rows = ROWS_PER_GPU
    torch.set_num_threads(host_cores())
    pol = cases.build_policy(oracle_vlnce, oracle_case(rows)).eval()
    opt = torch.optim.Adam(pol.parameters(), lr=2.5e-4)
    obs, prev, masks, actions, weights = make_batch(1000, rows)
    oracle_vlnce.AuxLosses.activate()
    oracle_vlnce.update_agent(pol, opt, obs, prev, masks, actions, weights)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        oracle_vlnce.update_agent(pol, opt, obs, prev, masks, actions, weights)
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds or n >= 12:
            break
    oracle_vlnce.AuxLosses.deactivate()
    out = {"value": rows * n / dt, "unit": "policy steps/s", "cores": torch.get_num_threads(), "kind": "port",
           "sample": f"{n} full {rows}-row steps after 1 warm-up (the benchmarked batch), oracle/vlnce.py on torch-CPU fp32; "
                     "oracle/hb017.py (habitat_baselines v0.1.7 restatement) is parity-unpinned at the dependency boundary"}
    try:  # run.py:66-67 forces one thread when CUDA is present
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        oracle_vlnce.AuxLosses.activate()
        oracle_vlnce.update_agent(pol, opt, obs, prev, masks, actions, weights)
        oracle_vlnce.AuxLosses.deactivate()
        out["one_thread_value"] = rows / (time.perf_counter() - t0)
    except Exception:
        pass
    return out


def parity_block(gpu_snapshot, host_batch):
    """ONE CPU-oracle step (checker) on the benchmark's own 64-row batch with the same seeded weights, compared with
    what the CUDA path produced for that batch before its first optimiser step."""
    import torch.nn.functional as F

    from oracle import cases, vlnce as oracle_vlnce

    torch.set_num_threads(host_cores())
    rows = ROWS_PER_GPU
    pol = cases.build_policy(oracle_vlnce, oracle_case(rows)).eval()
    obs = {k: host_batch[k] for k in (instr_key(), "rgb", "depth", "progress")}
    prev, masks, actions, weights = (host_batch[k] for k in ("prev_actions", "masks", "actions", "weights"))
    aux = oracle_vlnce.AuxLosses
    aux.activate()
    aux.clear()
    h0 = torch.zeros(rows // T_STEPS, pol.net.num_recurrent_layers, 512)
    dist = pol.build_distribution(obs, h0, prev, masks)
    logits = dist.logits.view(T_STEPS, rows // T_STEPS, -1)
    ce = F.cross_entropy(logits.permute(0, 2, 1), actions, reduction="none")
    action_loss = ((weights * ce).sum(0) / weights.sum(0)).mean()
    aux_loss = aux.reduce((weights > 0).view(-1))
    loss = action_loss + aux_loss
    loss.backward()
    aux.deactivate()
    ref_p = dist.logits.detach().exp()
    our_p = gpu_snapshot["logp"].exp()
    top2 = ref_p.topk(2, dim=-1).values
    margin = (top2[:, 0] - top2[:, 1])
    eq = (ref_p.argmax(-1) == our_p.argmax(-1))
    num = den = 0.0
    worst = ("", 0.0)
    named = dict(pol.named_parameters())
    for k, g in gpu_snapshot["grads"].items():
        r = named[k].grad if named[k].grad is not None else torch.zeros_like(named[k])
        num += float((g - r).double().pow(2).sum())
        den += float(r.double().pow(2).sum())
        rel = float((g - r).norm() / (r.norm() + 1e-12))
        if rel > worst[1] and float(r.norm()) > 1e-6:
            worst = (k, rel)
    ref_l = [float(loss), float(action_loss), float(aux_loss)]
    our_l = gpu_snapshot["losses"]
    return {"rows": rows, "probs_max_abs": float((ref_p - our_p).abs().max()),
            "probs_rel": float((ref_p - our_p).abs().max() / ref_p.max()),
            "loss_rel": abs(our_l[0] - ref_l[0]) / abs(ref_l[0]), "losses_ours": our_l, "losses_oracle": ref_l,
            "grad_rel_l2": (num / max(den, 1e-30)) ** 0.5, "grad_worst_tensor": worst[0], "grad_worst_rel_l2": worst[1],
            "argmax_equal": int(eq.sum()), "argmax_rows": rows,
            "argmax_mismatch_max_margin": float(margin[~eq].max()) if (~eq).any() else 0.0,
            "oracle": "oracle/vlnce.py torch-CPU fp32, same seeded weights, batch = input set 0"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="conv-trunk arithmetic: bf16/fp16 tcgen05 tiles (default) | tf32 = fp32 storage, kind::tf32 "
                         "products (what stock torch does for fp32 convolutions on a GPU) | fp32 = CUDA-core exact mode")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the workload's rows (64) are split over the ranks (64 / world rows per GPU) "
                         "instead of 64 rows per GPU; the line then says \"scaling\": \"strong\"")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = BASELINE metric config (default); "
                    "c5 = RxR-shaped features (side measurement)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--eager-allreduce", action="store_true",
                    help="world > 1: keep the NCCL all-reduce + Adam outside the CUDA graph (round-1 scheme)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the CPU-oracle parity block (one 64-row oracle step)")
    ap.add_argument("--profile", action="store_true", help="also write gpurun_out/step_profile.json")
    ap.add_argument("--pdl", type=int, default=None, choices=[0, 1],
                    help="programmatic dependent launch of the chained kernels (default: VLNCE_PDL / library default)")
    ap.add_argument("--ncu-step", action="store_true",
                    help="after warm-up run ONE eager, serial step between cudaProfilerStart/Stop and exit "
                         "(for `ncu --profile-from-start off`); prints no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    select_workload(args.workload)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.strong and world > 1:
        global ROWS_PER_GPU
        assert ROWS_PER_GPU % world == 0, "--strong: the workload's rows must divide over the ranks"
        ROWS_PER_GPU //= world
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    import vlnce_b200
    from vlnce_b200 import lib, synth
    from vlnce_b200.il_step import ILTrainerStep

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.pdl is not None:
        lib.C.vlnce_set_pdl(int(args.pdl))
    pol = synth.build_cma_policy(wseed=WSEED, hw=HW, rgb_hw=RGB_HW, num_actions=ACTIONS, rxr_dim=RXR_DIM).to(dev).eval()
    pol.net.precision = args.precision
    vlnce_b200.AuxLosses.activate()
    step = ILTrainerStep(pol, lr=2.5e-4)

    # ---- inputs: NSETS distinct batches, rotated so that consecutive steps never re-read the
    # same inputs from L2 (6 x 29.4 MB > 126 MB L2); pinned host copies for the e2e leg
    NSETS = 6
    host_sets, dev_sets = [], []
    for i in range(NSETS):
        b = make_batch(1000 + 17 * rank + i, ROWS_PER_GPU)
        obs, prev, masks, actions, weights = b
        flat = dict(obs, prev_actions=prev, masks=masks, actions=actions, weights=weights)
        host_sets.append({k: v.pin_memory() for k, v in flat.items()})
        dev_sets.append({k: v.to(dev) for k, v in flat.items()})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_sets[0].values())
    static = {k: torch.empty_like(v) for k, v in dev_sets[0].items() if k not in ("rgb", "depth")}
    static_next = {k: torch.empty_like(dev_sets[0][k]) for k in ("rgb", "depth")}

    def split(d):
        obs = {k: d[k] for k in (instr_key(), "rgb", "depth", "progress") if k in d}
        return obs, d["prev_actions"], d["masks"], d["actions"], d["weights"]

    VIS = ("rgb", "depth")

    def eager_step(d):
        """One step of the pipelined training loop: head fwd/loss/bwd/Adam of the CURRENT batch
        (non-visual tensors of d + the trunk features prefetched by the previous step) while the
        frozen trunks of the NEXT batch (static_next rgb/depth) run on a forked stream."""
        obs, prev, masks, actions, weights = split(d)
        obs_nv = {k: v for k, v in obs.items() if k not in VIS}
        return step.pipelined_step(obs_nv, prev, masks, actions, weights, static_next, apply=True)

    def grads_only(d):
        obs, prev, masks, actions, weights = split(d)
        obs_nv = {k: v for k, v in obs.items() if k not in VIS}
        step.pipelined_step(obs_nv, prev, masks, actions, weights, static_next, apply=False)

    def load_static(src, non_blocking=False, src_next=None):
        """current batch's non-visual tensors + the NEXT batch's images (pipelined loop)"""
        for k, v in static.items():
            v.copy_(src[k], non_blocking=non_blocking)
        nxt = src if src_next is None else src_next
        for k, v in static_next.items():
            v.copy_(nxt[k], non_blocking=non_blocking)

    def prime(first):
        """pipeline fill: trunk features of the first batch into the 'cur' slot"""
        step._prepare()
        step.prefetch_visual({k: first[k] for k in ("rgb", "depth")})
        step.rotate_visual(first["rgb"].shape[0])

    # ---- parity snapshot: forward + loss + backward of input set 0 at the seeded weights, BEFORE the first optimiser
    # step (this pass also autotunes every conv geometry); compared with one CPU-oracle step after the timed legs
    snapshot = None
    if rank == 0 and world == 1 and not args.no_parity:
        obs0, prev0, masks0, actions0, weights0 = split(dev_sets[0])
        step.compute_grads(obs0, prev0, masks0, actions0, weights0)
        eng = pol.net.engine()
        torch.cuda.synchronize()
        snapshot = {"logp": eng.buf("logits", (ROWS_PER_GPU, ACTIONS)).detach().cpu().clone(),
                    "losses": step._losses.tolist(),
                    "grads": {k: v.detach().cpu().clone() for k, v in step.fp.grad_views.items()}}

    # ---- warm-up (eager), count launches per step, then capture the step into a CUDA graph
    prime(dev_sets[0])
    for i in range(2):
        load_static(dev_sets[i % NSETS], src_next=dev_sets[(i + 1) % NSETS])
        eager_step(static)
    torch.cuda.synchronize()
    c0 = lib.C.vlnce_launch_count()
    load_static(dev_sets[2 % NSETS], src_next=dev_sets[3 % NSETS])
    eager_step(static)
    torch.cuda.synchronize()
    launches_per_step = int(lib.C.vlnce_launch_count() - c0)
    graph = None
    graph_has_allreduce = False
    losses = step._losses
    if args.ncu_step:
        from vlnce_b200 import engine as engine_mod

        engine_mod.set_overlap(False)
        load_static(dev_sets[3 % NSETS], src_next=dev_sets[4 % NSETS])
        eager_step(static)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        load_static(dev_sets[4 % NSETS], src_next=dev_sets[5 % NSETS])
        eager_step(static)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_step": True, "launches": launches_per_step}), flush=True)
        return
    if not args.no_graph:
        try:
            # The whole pipelined step is ONE graph, the NCCL gradient all-reduce included (world > 1): inside the graph
            # the all-reduce + Adam of batch i sit on the head branch and overlap the frozen trunks of batch i+1 on the
            # forked branch instead of trailing every replay.  If this NCCL build refuses capture, fall back to
            # forward+loss+backward in the graph and all-reduce + Adam launched eagerly after each replay.
            def capture(body):
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    body(static)
                    side.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        body(static)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                return g

            graph_has_allreduce = True
            if world > 1 and args.eager_allreduce:
                raise RuntimeError("eager all-reduce requested")
            graph = capture(eager_step)
        except Exception as e:  # graph capture is an optimisation, not a requirement
            graph = None
            torch.cuda.synchronize()
            if world > 1:
                try:
                    graph_has_allreduce = False
                    graph = capture(grads_only)
                except Exception as e2:
                    print(f"[bench] CUDA-graph capture failed ({type(e2).__name__}: {e2}); running eagerly", file=sys.stderr)
                    graph = None
                    torch.cuda.synchronize()
            else:
                print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)

    def run_step():
        if graph is not None:
            graph.replay()
            if world > 1 and not graph_has_allreduce:
                step.apply_grads(True)
        else:
            eager_step(static)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- leg 1: inputs resident in HBM ----------------------------------------------------
    for i in range(args.warmup):
        load_static(dev_sets[i % NSETS], src_next=dev_sets[(i + 1) % NSETS])
        run_step()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples), daemon=True)
    th.start()
    time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        # device->device rotation of the input sets (see config.l2): batch i's small tensors, batch i+1's images
        load_static(dev_sets[i % NSETS], src_next=dev_sets[(i + 1) % NSETS])
        run_step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    rows_total = ROWS_PER_GPU * world * args.steps
    value = rows_total / (ms_total * 1e-3)

    # ---- leg 2: end to end through the public API, host buffers ---------------------------
    # Every step's inputs travel pinned-host -> device INSIDE the timed region (one H2D per step)
    # and the step's losses travel back (D2H + host read).  Like a DataLoader prefetcher, the copy
    # of batch i+1 runs on a copy stream while batch i computes.
    host_loss = [torch.empty(3).pin_memory() for _ in range(2)]
    ev_loss = [torch.cuda.Event(), torch.cuda.Event()]
    stage_buf = {k: torch.empty_like(v) for k, v in {**static, **static_next}.items()}
    copy_stream = torch.cuda.Stream()
    ev_ready, ev_free = torch.cuda.Event(), torch.cuda.Event()

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free)
            for k, v in stage_buf.items():   # H2D from pinned memory: batch i's small tensors + batch i+1's images
                src = host_sets[(i + 1) % NSETS] if k in static_next else host_sets[i % NSETS]
                v.copy_(src[k], non_blocking=True)
            ev_ready.record(copy_stream)

    def e2e_loop(nsteps):
        last = None
        main = torch.cuda.current_stream()
        ev_free.record(main)
        prefetch(0)
        for i in range(nsteps):
            main.wait_event(ev_ready)
            load_static(stage_buf)                  # device-side hand-over into the step's buffers
            ev_free.record(main)
            if i + 1 < nsteps:
                prefetch(i + 1)
            run_step()
            # D2H of THIS step's losses into a pinned slot; the host reads the PREVIOUS step's slot (its copy finished
            # while this step was being enqueued), so every step's result is read inside the timed region without
            # draining the queue after each step — what a trainer that logs its loss one step late does
            host_loss[i & 1].copy_(losses, non_blocking=True)
            ev_loss[i & 1].record(main)
            if i > 0:
                ev_loss[(i - 1) & 1].synchronize()
                last = host_loss[(i - 1) & 1].tolist()
        ev_loss[(nsteps - 1) & 1].synchronize()
        last = host_loss[(nsteps - 1) & 1].tolist()
        return last

    e2e_loop(args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    last = e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_wall = time.perf_counter() - t_wall
    e2e_value = rows_total / (max(e2e_ms * 1e-3, e2e_wall if world == 1 else 0.0))
    stop.set()
    th.join(timeout=3)

    # ---- roofline of the dominant kernel (tcgen05 conv tiles), timed live with CUDA events
    # around every conv launch of 3 further eager steps on the launching stream
    from vlnce_b200 import trunk as trunk_mod

    from vlnce_b200 import roofline as roofline_mod

    rec, work = [], []
    orig = trunk_mod.ConvTrunk._conv

    def timed_conv(self, x, shape_in, cv, y, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(self, x, shape_in, cv, y, **kw)
        b.record()
        n, h, w, c = shape_in
        gemm_c = kw.get("gemm_c")
        if kw.get("stem_s2d"):  # x = s2d cells, (h, w) = the stem's output size; algorithmic 7x7x3 taps
            flops = 2.0 * n * h * w * cv.K * 147
        elif gemm_c is not None:
            flops = 2.0 * n * h * w * gemm_c * cv.K
        else:
            p, q = self._out_hw(h, w, cv)
            flops = 2.0 * n * p * q * cv.K * cv.R * cv.S * c
        rec.append((a, b, flops, type(self).__name__))
        try:  # per-launch algorithmic bytes / flops for the layer roofline model (never allowed to break the timing)
            work.append(roofline_mod.conv_launch_work(self, x, shape_in, cv, y, kw))
        except Exception:
            pass

    orig_fused = trunk_mod.DepthTrunk._forward_fused

    def timed_fused(self, depth, out, ld_out, out_hw):
        """the image-stationary depth trunk is ONE launch covering all of that trunk's convolutions"""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ok = orig_fused(self, depth, out, ld_out, out_hw)
        b.record()
        if ok:
            n, h, w, _ = depth.shape
            plan = next(p for p in self._plans.values() if p is not False)
            table = plan[2]
            flops = byts = 0.0
            for row in table.tolist():
                if row[0] != 0:
                    continue
                cin, cout, k, stride, hin, win = row[9], row[10], row[11], row[12], row[13], row[14]
                taps_k = 49 if row[17] else k * k * cin      # the stem's operand is the zero-padded 7x7 im2col
                flops += 2.0 * n * (hin // stride) * (win // stride) * cout * taps_k
                byts += 2.0 * cout * k * k * cin             # weights once
            byts += 4.0 * n * h * w + 4.0 * out.shape[0] * out.shape[1] * self.out_channels   # depth image in, features out
            rec.append((a, b, flops, "DepthTrunk"))
            work.append({"flops": flops, "bytes": byts, "trunk": "DepthTrunk", "K": 0, "C": 0, "R": 0, "pixels": n})
        return ok

    roof = None
    from vlnce_b200 import engine as engine_mod

    engine_mod.set_overlap(False)  # serial launches: clean per-kernel times
    if args.profile:
        with lib.profile() as prof:
            for i in range(3):
                load_static(dev_sets[i % NSETS], src_next=dev_sets[(i + 1) % NSETS])
                torch.cuda._sleep(30_000_000)  # enqueue the step behind a spin: device time, not launch latency
                eager_step(static)
                torch.cuda.synchronize()
        summ = prof.summary()
        for v in summ.values():
            v["us"] = round(v["us"] / 3, 1)
            v["calls"] //= 3
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"note": "CUDA-event time per C-ABI call, serial streams, warm caches, per step",
                   "total_us": round(sum(v["us"] for v in summ.values()), 1), "calls": summ},
                  open("gpurun_out/step_profile.json", "w"), indent=1)
    if args.precision == "bf16":
        trunk_mod.ConvTrunk._conv = timed_conv
        trunk_mod.DepthTrunk._forward_fused = timed_fused
        for i in range(3):
            load_static(dev_sets[i % NSETS], src_next=dev_sets[(i + 1) % NSETS])
            # park the stream behind a ~15 ms spin so the whole step is enqueued before its first kernel
            # runs: the event pairs then bracket back-to-back device time, not the host's launch latency
            torch.cuda._sleep(30_000_000)
            eager_step(static)
            torch.cuda.synchronize()
        trunk_mod.ConvTrunk._conv = orig
        trunk_mod.DepthTrunk._forward_fused = orig_fused
        conv_ms = sum(a.elapsed_time(b) for a, b, _, _ in rec)
        conv_flops = sum(f for _, _, f, _ in rec)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        roof = {"kernel": "conv_tma_kernel + conv_umma_kernel (tcgen05 implicit-GEMM conv tiles, all 107 conv launches/step)",
                "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s",
                "launches_timed": len(rec), "conv_ms_per_step": conv_ms / 3,
                "algorithmic_gflop_per_step": conv_flops / 3 / 1e9}
        try:  # quote the fraction against the resource that bounds the launches (per-launch model)
            roof = roofline_mod.assemble(conv_ms, conv_flops, work, peaks, traffic, len(rec), steps=3)
        except Exception as e:  # keep the tensor-peak figures above
            roof["layer_model_error"] = repr(e)

    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg()
    if snapshot is not None:
        try:
            parity = parity_block(snapshot, {k: v for k, v in host_sets[0].items()})
        except Exception as e:  # the checker must never take the measurement down
            parity = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "policy steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "tf32": "tf32"}.get(args.precision, "f32"), "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": ROWS_PER_GPU * world, "rows_per_gpu": ROWS_PER_GPU,
                       "parallelism": f"dp{world}", "precision": {"bf16": "rgb trunk bf16 / depth trunk fp16 tcgen05, head fp32",
                                     "tf32": "trunks fp32 storage / TF32 tcgen05 products, head fp32"}.get(
                                         args.precision, "fp32 CUDA cores"),
                       "l2": f"inputs rotated over {NSETS} distinct batches ({NSETS * h2d_bytes / 1e6:.0f} MB > L2)",
                       "cuda_graph": graph is not None, "allreduce_in_graph": bool(graph is not None and graph_has_allreduce and world > 1),
                       "pdl": bool(lib.C.vlnce_get_pdl()),
                       "schedule": "pipelined loop: frozen trunks of batch i+1 overlap head fwd/bwd/Adam of batch i "
                                   "(every timed step = one trunk pass + one head update; fill done in warm-up)"},
            "e2e": {"value": e2e_value, "unit": "policy steps/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 12, "ms_per_step": e2e_ms / args.steps, "last_losses": last,
                    "d2h_read": "every step's 3 losses are copied to pinned memory and read by the host inside the timed "
                                "region; the read of step i happens while step i+1 runs (one step of lag)"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": summarize_clocks(samples),
        }
        if roof is not None:
            line["roofline"] = roof
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    sys.stderr.flush()
    # Orderly teardown (graph before communicator, both before interpreter shutdown) so that the process exits
    # normally and exit hooks run.
    torch.cuda.synchronize()
    graph = None  # noqa: F841  (drops the captured graph and its private pool)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
