"""Device-time probe of the step's phases (not the benchmark): CUDA-graph replays of the RGB
trunk, the depth trunk, both trunks concurrently, the head update alone, at the BASELINE batch."""
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import vlnce_b200  # noqa: E402
from oracle import cases  # noqa: E402  (seeded-policy builder only)
from vlnce_b200 import synth  # noqa: E402
from vlnce_b200.il_step import ILTrainerStep  # noqa: E402


def graph_time(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    import argparse

    from vlnce_b200 import lib

    ap = argparse.ArgumentParser()
    ap.add_argument("--caps", default="0", help="comma-separated persistent-grid caps for conv_tma (0 = none)")
    ap.add_argument("--chunks", default="", help="comma-separated depth-trunk chunk counts to sweep")
    ap.add_argument("--trunks-only", default="", help="precision (bf16 | tf32 | fp32): time the two trunks in that mode and exit")
    args = ap.parse_args()
    if args.trunks_only:
        prec = args.trunks_only
        n = 64
        pol = synth.build_cma_policy().cuda().eval()
        pol.net.precision = prec
        obs = {"rgb": torch.randint(0, 256, (n, 256, 256, 3), dtype=torch.uint8, device="cuda"),
               "depth": torch.rand(n, 256, 256, 1, device="cuda")}
        o1, o2 = torch.zeros(n, 16, 2112, device="cuda"), torch.zeros(n, 16, 192, device="cuda")
        pol.net.rgb_encoder.features_into(obs, o1, 2112, prec)
        pol.net.depth_encoder.features_into(obs, o2, 192, prec)
        torch.cuda.synchronize()
        res = {"precision": prec,
               "rgb_trunk_ms": graph_time(lambda: pol.net.rgb_encoder.features_into(obs, o1, 2112, prec), iters=10),
               "depth_trunk_ms": graph_time(lambda: pol.net.depth_encoder.features_into(obs, o2, 192, prec), iters=10)}
        with lib.profile() as prof:
            pol.net.rgb_encoder.features_into(obs, o1, 2112, prec)
            pol.net.depth_encoder.features_into(obs, o2, 192, prec)
        res["calls"] = {k: {"calls": v["calls"], "us": round(v["us"], 1)} for k, v in list(prof.summary().items())[:6]}
        print(json.dumps(res))
        return
    n = 64
    case = dict(cases.CASES["cma_train_pixels"], t=1, n=n)
    pol = cases.build_policy(vlnce_b200, case).cuda().eval()
    vlnce_b200.AuxLosses.activate()
    step = ILTrainerStep(pol)
    obs, prev, masks, actions, weights = (cases._to(v, "cuda") for v in synth.make_train_batch(1, 1, n))
    eng = pol.net.engine()
    eng.ensure_pool(torch.device("cuda"))
    step._prepare()
    rgb_e, dep_e = eng.visual_buffers(n, "next")
    eng.encode_visual(obs, rgb_e, dep_e)  # autotune every layer once, uncapped
    torch.cuda.synchronize()
    sweep = {}
    for cap in [int(c) for c in args.caps.split(",")][1:]:
        lib.C.vlnce_set_tma_cta_cap(cap)
        sweep[cap] = {
            "rgb_trunk_ms": graph_time(lambda: pol.net.rgb_encoder.features_into(obs, rgb_e, eng.Crgb, "bf16")),
            "both_trunks_ms": graph_time(lambda: eng.encode_visual(obs, rgb_e, dep_e))}
        print("cap", cap, json.dumps(sweep[cap]), flush=True)
    lib.C.vlnce_set_tma_cta_cap(int(args.caps.split(",")[0]))
    from vlnce_b200 import trunk as trunk_mod

    csweep = {}
    for k in [int(c) for c in args.chunks.split(",") if c]:
        trunk_mod.DEPTH_CHUNKS = k
        pol.net.depth_encoder.features_into(obs, dep_e, eng.Cdep, "bf16")  # tune the chunk geometries
        torch.cuda.synchronize()
        csweep[k] = {
            "depth_trunk_ms": graph_time(lambda: pol.net.depth_encoder.features_into(obs, dep_e, eng.Cdep, "bf16")),
            "both_trunks_ms": graph_time(lambda: eng.encode_visual(obs, rgb_e, dep_e)),
            "pipelined_step_ms": None}
        print("chunks", k, json.dumps(csweep[k]), flush=True)
    trunk_mod.DEPTH_CHUNKS = int(os.environ.get("VLNCE_DEPTH_CHUNKS", "1"))
    res = {"cap": int(args.caps.split(",")[0]), "cap_sweep": sweep, "chunk_sweep": csweep}
    res["rgb_trunk_ms"] = graph_time(lambda: pol.net.rgb_encoder.features_into(obs, rgb_e, eng.Crgb, "bf16"))
    res["depth_trunk_ms"] = graph_time(lambda: pol.net.depth_encoder.features_into(obs, dep_e, eng.Cdep, "bf16"))
    res["both_trunks_ms"] = graph_time(lambda: eng.encode_visual(obs, rgb_e, dep_e))
    step.prefetch_visual(obs)
    step.rotate_visual(n)
    obs_nv = {k: v for k, v in obs.items() if k not in ("rgb", "depth")}

    def head():
        step.compute_grads(obs_nv, prev, masks, actions, weights, visual_slot="cur")
        step.apply_grads(True)

    res["head_update_ms"] = graph_time(head)
    res["full_step_ms"] = graph_time(lambda: step.update_agent(obs, prev, masks, actions, weights, return_tensor=True))
    vis_next = {k: obs[k] for k in ("rgb", "depth")}
    res["pipelined_step_ms"] = graph_time(lambda: step.pipelined_step(obs_nv, prev, masks, actions, weights, vis_next))
    for cap in [int(c) for c in args.caps.split(",")][1:]:   # persistent-grid cap of conv_tma: SMs left to the head
        lib.C.vlnce_set_tma_cta_cap(cap)
        res.setdefault("pipelined_by_cap", {})[cap] = graph_time(
            lambda: step.pipelined_step(obs_nv, prev, masks, actions, weights, vis_next))
    lib.C.vlnce_set_tma_cta_cap(int(args.caps.split(",")[0]))
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/phases.json", "w").write(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
