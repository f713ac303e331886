"""Per-layer device timing of the tensor-core conv tiles on the BASELINE batch (N=64):
every distinct conv geometry of the RGB ResNet-50 trunk (and optionally the depth trunk),
each legal tile width, CUDA-event timed on warm clocks.  Writes gpurun_out/conv_layers.json."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlnce_b200 import lib  # noqa: E402
from vlnce_b200.lib import C, ptr  # noqa: E402

# (name, H(=W) in, Cin, Cout, R, stride, residual, count) — torchvision ResNet-50 v1.5 @256x256
RGB = [
    ("stem_gemm", 128, 192, 64, 1, 1, False, 1),
    ("l1.c1 1x1 64->64", 64, 64, 64, 1, 1, False, 1), ("l1.c1 1x1 256->64", 64, 256, 64, 1, 1, False, 2),
    ("l1.c2 3x3 64", 64, 64, 64, 3, 1, False, 3), ("l1.c3 1x1 64->256 +res", 64, 64, 256, 1, 1, True, 3),
    ("l1.ds 1x1 64->256", 64, 64, 256, 1, 1, False, 1),
    ("l2.c1 1x1 256->128", 64, 256, 128, 1, 1, False, 1), ("l2.c2 3x3/2 128", 64, 128, 128, 3, 2, False, 1),
    ("l2.ds 1x1/2 256->512", 64, 256, 512, 1, 2, False, 1), ("l2.c1 1x1 512->128", 32, 512, 128, 1, 1, False, 3),
    ("l2.c2 3x3 128", 32, 128, 128, 3, 1, False, 3), ("l2.c3 1x1 128->512 +res", 32, 128, 512, 1, 1, True, 4),
    ("l3.c1 1x1 512->256", 32, 512, 256, 1, 1, False, 1), ("l3.c2 3x3/2 256", 32, 256, 256, 3, 2, False, 1),
    ("l3.ds 1x1/2 512->1024", 32, 512, 1024, 1, 2, False, 1), ("l3.c1 1x1 1024->256", 16, 1024, 256, 1, 1, False, 5),
    ("l3.c2 3x3 256", 16, 256, 256, 3, 1, False, 5), ("l3.c3 1x1 256->1024 +res", 16, 256, 1024, 1, 1, True, 6),
    ("l4.c1 1x1 1024->512", 16, 1024, 512, 1, 1, False, 1), ("l4.c2 3x3/2 512", 16, 512, 512, 3, 2, False, 1),
    ("l4.ds 1x1/2 1024->2048", 16, 1024, 2048, 1, 2, False, 1), ("l4.c1 1x1 2048->512", 8, 2048, 512, 1, 1, False, 2),
    ("l4.c2 3x3 512", 8, 512, 512, 3, 1, False, 2), ("l4.c3 1x1 512->2048 +res", 8, 512, 2048, 1, 1, True, 3),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--bn", type=int, default=-1, help="-1: sweep all tile widths")
    args = ap.parse_args()
    dev = "cuda"
    out = []
    for _ in range(50):  # clocks up
        torch.randn(4096, 4096, device=dev) @ torch.randn(4096, 4096, device=dev)
    total_best = 0.0
    for name, hw, cin, cout, r, stride, res, count in RGB:
        if args.only and not any(o in name for o in args.only.split(",")):
            continue
        n = args.n
        pad = r // 2
        if name == "stem_gemm":
            x = torch.randn(n * hw * hw, 1, 1, cin, device=dev).bfloat16()
            geom = (n * hw * hw, 1, 1, cin, cout, 1, 1, 1, 0)
            p = q = 1
            m = n * hw * hw
        else:
            x = torch.randn(n, hw, hw, cin, device=dev).bfloat16()
            geom = (n, hw, hw, cin, cout, r, r, stride, pad)
            p = q = (hw + 2 * pad - r) // stride + 1
            m = n * p * q
        w = (torch.randn(cout, r, r, cin, device=dev) * 0.05).bfloat16()
        scale, shift = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        y = torch.empty(m, cout, device=dev, dtype=torch.bfloat16)
        resid = torch.randn(m, cout, device=dev).bfloat16() if res else None
        flops = 2.0 * m * cout * r * r * cin
        bytes_min = 2.0 * (x.numel() + w.numel() + y.numel() * (2 if res else 1))
        row = {"layer": name, "count": count, "M": m, "gflop": flops / 1e9, "min_MB": bytes_min / 1e6, "bn": {}}
        cands = [bn for bn in (32, 64, 128, 256) if cout % bn == 0] if args.bn < 0 else [args.bn]
        for bn in cands:
            def launch():
                C.vlnce_conv2d_bf16(ptr(x), ptr(w), lib.BF16, ptr(scale), ptr(shift), ptr(resid), ptr(y), lib.BF16,
                                    *geom, 1, bn, lib.stream())
            for _ in range(3):
                launch()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.reps):
                launch()
            b.record()
            b.synchronize()
            us = a.elapsed_time(b) * 1e3 / args.reps
            row["bn"][bn] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1),
                             "min_traffic_gbs": round(bytes_min / us / 1e3, 0)}
        if r == 3 and stride == 1 and hw % 16 == 0 and cin % 64 == 0 and cout % 64 == 0 and args.bn < 0:
            for bn in [b for b in (64, 128, 256) if cout % b == 0]:
                for ring in (3, 6):
                    def launch():
                        C.vlnce_conv3x3_halo(ptr(x), ptr(w), lib.BF16, ptr(scale), ptr(shift), ptr(resid), ptr(y),
                                             lib.BF16, n, hw, hw, cin, cout, 1, bn | (ring << 16), lib.stream())
                    for _ in range(3):
                        launch()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(args.reps):
                        launch()
                    b.record()
                    b.synchronize()
                    us = a.elapsed_time(b) * 1e3 / args.reps
                    row["bn"][f"halo{bn}r{ring}"] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1),
                                                    "min_traffic_gbs": round(bytes_min / us / 1e3, 0)}
        if stride == 1 and r in (1, 3) and cin % 64 == 0 and cout % 64 == 0 and args.bn < 0 and (r == 1 or hw % 16 == 0):
            MT2, NOSTAT = 1 << 21, 1 << 22
            for bn in [b for b in (64, 128, 256) if cout % b == 0]:
                for tag, fl in (("", NOSTAT), ("s", 0), ("m", MT2 | NOSTAT), ("ms", MT2)):
                    if "m" in tag and bn == 256:
                        continue
                    def launch():
                        C.vlnce_conv_tma(ptr(x), ptr(w), lib.BF16, ptr(scale), ptr(shift), ptr(resid), ptr(y), lib.BF16,
                                         geom[0], geom[1], geom[2], geom[3], cout, r, 1, bn | (4 << 16) | fl, lib.stream())
                    for _ in range(3):
                        launch()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(args.reps):
                        launch()
                    b.record()
                    b.synchronize()
                    us = a.elapsed_time(b) * 1e3 / args.reps
                    row["bn"][f"P{bn}{tag}"] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1),
                                               "min_traffic_gbs": round(bytes_min / us / 1e3, 0)}
        if stride == 2 and cin % 64 == 0 and cout % 64 == 0 and (hw // 2) % 8 == 0 and args.bn < 0:
            MT2, NOSTAT = 1 << 21, 1 << 22
            for bn in [b for b in (64, 128, 256) if cout % b == 0]:
                for tag, fl in (("", NOSTAT), ("s", 0), ("m", MT2 | NOSTAT)):
                    if "m" in tag and bn == 256:
                        continue
                    def launch():
                        C.vlnce_conv_s2_tma(ptr(x), ptr(w), lib.BF16, ptr(scale), ptr(shift), ptr(y), lib.BF16, n, hw, hw,
                                            cin, cout, r, 1, bn | (4 << 16) | fl, lib.stream())
                    for _ in range(3):
                        launch()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(args.reps):
                        launch()
                    b.record()
                    b.synchronize()
                    us = a.elapsed_time(b) * 1e3 / args.reps
                    row["bn"][f"S{bn}{tag}"] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1),
                                               "min_traffic_gbs": round(bytes_min / us / 1e3, 0)}
        best = min(row["bn"].items(), key=lambda kv: kv[1]["us"])
        row["best_bn"], row["best_us"] = best[0], best[1]["us"]
        total_best += best[1]["us"] * count
        print(f"{name:28s} x{count} M={m:7d} {flops/1e9:6.1f} GF  " +
              "  ".join(f"bn{bn}:{v['us']:7.1f}us/{v['tflops']:6.1f}TF" for bn, v in row["bn"].items()), flush=True)
        out.append(row)
    if not args.only or "stem" in args.only.split(","):
        # s2d stem: cell image + TMA conv, each variant; and the gather stem for reference
        n, hw = args.n, 256
        img = torch.randint(0, 256, (n, hw, hw, 3), device=dev, dtype=torch.uint8)
        cells = torch.empty(n, 131, 131, 16, device=dev, dtype=torch.bfloat16)
        ws = (torch.randn(64, 256, device=dev) * 0.05).bfloat16()
        wg = (torch.randn(64, 192, device=dev) * 0.05).bfloat16()
        scale, shift = torch.ones(64, device=dev), torch.zeros(64, device=dev)
        y = torch.empty(n, 128, 128, 64, device=dev, dtype=torch.bfloat16)
        MT2, NOSTAT = 1 << 21, 1 << 22

        def timeit(fn):
            for _ in range(3):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.reps):
                fn()
            b.record()
            b.synchronize()
            return a.elapsed_time(b) * 1e3 / args.reps

        res = {"gather_stem": timeit(lambda: C.vlnce_conv2d_stem_rgb(
            ptr(img), lib.U8, ptr(wg), lib.BF16, ptr(scale), ptr(shift), ptr(y), lib.BF16, n, hw, hw, 64, 1, 0,
            lib.stream()))}
        try:
            res["s2d_cells"] = timeit(lambda: C.vlnce_stem_s2d(ptr(img), lib.U8, ptr(cells), lib.BF16, n, hw, hw,
                                                               lib.stream()))
            for tag, fl in (("", NOSTAT), ("s", 0), ("m", MT2 | NOSTAT), ("ms", MT2)):
                res["s2d_conv64" + tag] = timeit(lambda: C.vlnce_conv_stem_tma(
                    ptr(cells), ptr(ws), lib.BF16, ptr(scale), ptr(shift), ptr(y), lib.BF16, n, 128, 128, 64, 1,
                    64 | (4 << 16) | fl, lib.stream()))
        except lib.VlnceError as e:
            res["s2d_error"] = str(e)
        print("stem (us):", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
        out.append({"layer": "rgb stem variants", "us": res})
    print(f"sum of best per-layer times x count: {total_best:.0f} us")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/conv_layers.json", "w"), indent=1)


if __name__ == "__main__":
    main()
