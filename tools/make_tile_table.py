#!/usr/bin/env python
"""Measure the conv tile-selection table shipped as ``vlnce_b200/tile_table_b200.json`` (run on a B200):
runs the policies of the BASELINE configs once per batch bucket with the autotuner on and an empty table, and dumps
every (geometry -> kernel variant | tile width | ring depth) decision.

    VLNCE_TILE_TABLE=0 python tools/make_tile_table.py gpurun_out/tile_table_b200.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VLNCE_TILE_TABLE", "0")
from vlnce_b200 import synth, trunk  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tile_table_b200.json"
    dev = "cuda"
    for rgb_hw in (256, 224):                      # C2 / C1 frames, C3 / C5 frames after the observation transforms
        pol = synth.build_cma_policy(rgb_hw=rgb_hw).to(dev).eval()
        for n in (64, 32, 16, 8, 4, 2, 1):
            obs = {"rgb": torch.randint(0, 256, (n, rgb_hw, rgb_hw, 3), dtype=torch.uint8, device=dev),
                   "depth": torch.rand(n, 256, 256, 1, device=dev)}
            o1 = torch.zeros(n, 16, 2112, device=dev)
            o2 = torch.zeros(n, 16, 192, device=dev)
            pol.net.rgb_encoder.features_into(obs, o1, 2112, "bf16")
            t = pol.net.depth_encoder.visual_encoder.trunk("bf16")
            t.fused = False                        # the per-layer depth path (fallback geometries / training tape sizes)
            pol.net.depth_encoder.features_into(obs, o2, 192, "bf16")
            t.fused = True
            torch.cuda.synchronize()
            print(f"rgb {rgb_hw} n {n}: {len(trunk._TUNE_CACHE)} entries", flush=True)
    # waypoint policy: ResNet-18 RGB trunk at 224, frames = 13 per env
    from oracle import cases
    import vlnce_b200

    pol = cases.build_policy(vlnce_b200, cases.CASES["wp_act"]).to(dev).eval()
    for n in (64, 32, 16, 8, 4, 2, 1):
        obs = {"rgb": torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev)}
        o1 = torch.zeros(n, 16, 576, device=dev)
        pol.net.rgb_encoder.features_into(obs, o1, 16 * 576, "bf16")
        torch.cuda.synchronize()
    print(f"resnet18: {len(trunk._TUNE_CACHE)} entries", flush=True)
    json.dump(dict(sorted(trunk._TUNE_CACHE.items())), open(out, "w"), indent=0)
    print("wrote", out)


if __name__ == "__main__":
    main()
