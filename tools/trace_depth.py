"""Per-step timeline of the image-stationary depth trunk (csrc/depth_fused.cu): CTA 0's %globaltimer stamps for its
first image.  Usage: python tools/trace_depth.py [n_images]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlnce_b200 import synth  # noqa: E402
from vlnce_b200.lib import C, ptr  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    pol = synth.build_cma_policy().cuda().eval()
    enc = pol.net.depth_encoder
    d = torch.rand(n, 256, 256, 1, device="cuda")
    out = torch.zeros(n, 16, 192, device="cuda")
    for _ in range(3):
        enc.features_into({"depth": d}, out, 192, "bf16")
    torch.cuda.synchronize()
    tr = enc.visual_encoder.trunk("bf16")
    plan = next(p for p in tr._plans.values() if p is not False)
    table = plan[2]
    n_steps = table.shape[0]
    trace = torch.zeros(n_steps * 2 * 8 + 128, dtype=torch.int64, device="cuda")
    C.vlnce_depth_fused_set_trace(ptr(trace))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    enc.features_into({"depth": d}, out, 192, "bf16")
    b.record()
    torch.cuda.synchronize()
    C.vlnce_depth_fused_set_trace(None)
    fine = trace.cpu()[n_steps * 2 * 8:]
    t = trace.cpu()[:n_steps * 2 * 8].view(n_steps, 2, 8)
    t0 = int(t[t > 0].min())
    print(f"n_images {n}: im2col + fused kernel {a.elapsed_time(b) * 1e3:.1f} us; columns: us since first stamp")
    print("step slab kind  cin->cout k s  hin | mma_start mma_done epi_start stats_done stored x_released | mma  wait  stats  apply")
    prev_end = 0.0
    for si in range(n_steps):
        row = table[si].tolist()
        for sl in range(2):
            r = t[si, sl]
            if int(r[:5].max()) == 0 and int(r[5]) == 0:
                continue
            us = [(int(v) - t0) / 1e3 if int(v) > 0 else float("nan") for v in r[:6]]
            mma = us[1] - us[0]
            print(f"{si:3d} {sl} k{row[0]} {row[9]:5d}->{row[10]:<5d} {row[11]} {row[12]} {row[13]:4d} | "
                  f"{us[0]:8.1f} {us[1]:8.1f} {us[2]:8.1f} {us[3]:8.1f} {us[4]:8.1f} {us[5]:8.1f} | "
                  f"{mma:6.1f} {us[2] - us[1]:5.1f} {us[3] - us[2]:6.1f} {us[4] - us[3]:6.1f}   step total {us[4] - prev_end:6.1f}"
                  f"  mma-phase {us[1] - prev_end:6.1f}")
            prev_end = us[4]


if __name__ == "__main__":
    main()
