#!/usr/bin/env python
"""Per-layer roofline model of the conv tiles of one bench step (BASELINE config C2: batch 64, 256x256 RGB-D), from the
launch plan itself — no GPU needed: the two trunks are sequenced on the host in dry-run mode with a tiny batch, every
``ConvTrunk._conv`` call is intercepted, and its algorithmic flops / bytes are evaluated at the bench batch size with
``vlnce_b200.roofline``.  Peaks: MEASURED_PEAKS.json (sustained bf16 TFLOP/s, HBM copy GB/s).

    python tools/roofline_model.py [--batch 64] [--hw 256] [--measured-conv-ms 3.07] [--md profiles/r01e_layer_roofline.md]
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


def collect(batch: int, hw: int):
    import vlnce_b200
    from oracle import cases
    from vlnce_b200 import lib, roofline, synth, trunk as trunk_mod

    case = dict(cases.CASES["cma_act"], n=2, hw=hw, length=8)
    work = []
    orig = trunk_mod.ConvTrunk._conv

    def hook(self, x, shape_in, cv, y, **kw):
        orig(self, x, shape_in, cv, y, **kw)
        w = roofline.conv_launch_work(self, x, (batch,) + tuple(shape_in[1:]), cv, y, kw)
        w["name"] = f"{type(self).__name__[:-5].lower()} {cv.R}x{cv.S} {w['C']}->{cv.K}" + \
            ("/s2" if cv.stride == 2 else "") + (" +res" if kw.get("residual") is not None else "") + \
            (" (stem)" if (kw.get("stem_s2d") or kw.get("stem_rgb") or kw.get("gemm_c") is not None) else "")
        work.append(w)

    with lib.dry_run():
        pol = cases.build_policy(vlnce_b200, case)
        pol.eval()
        obs, h, prev, masks = synth.make_act_batch(1, 2, hw=hw, length=8)
        trunk_mod.ConvTrunk._conv = hook
        try:
            with torch.no_grad():
                pol.act(obs, h, prev, masks, deterministic=True)
        finally:
            trunk_mod.ConvTrunk._conv = orig
    return work


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--measured-conv-ms", type=float, default=None,
                    help="bench.py roofline.conv_ms_per_step of the same config (serial, CUDA-event timed)")
    ap.add_argument("--md", default=None)
    ap.add_argument("--launch-csv", default=None,
                    help="ncu launch list of one serial step (gpu__time_duration.sum); its conv launches are matched to "
                         "the plan in order and a per-layer measured-vs-roofline table is appended (--gap-md)")
    ap.add_argument("--gap-md", default=None)
    args = ap.parse_args()
    from vlnce_b200 import roofline

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # driver-written file absent: the round-1 measurements of this pool's B200s
        peaks = {"bf16_tflops_sustained": 1369.7, "hbm_gbs": 6534.8}
    pt, ph = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
    work = collect(args.batch, args.hw)
    summ = roofline.summarize(work, args.measured_conv_ms or 0.0, pt, ph)
    # group identical launches
    rows, order = {}, []
    for w in work:
        key = (w["trunk"], w["name"], w["pixels"])
        if key not in rows:
            rows[key] = dict(w, count=0)
            order.append(key)
        rows[key]["count"] += 1
    lines = [f"# per-layer roofline model of the conv tiles — batch {args.batch}, {args.hw}x{args.hw} RGB-D (config C2)", "",
             f"Peaks (MEASURED_PEAKS.json): {pt:.1f} TFLOP/s sustained bf16, {ph:.1f} GB/s HBM copy -> ridge "
             f"{pt * 1e3 / ph:.0f} flop/byte.  Algorithmic bytes: input tile + weights + output tile (+ residual) in the "
             "dtypes this path stores them in.  `t_min` = max(flops / tensor peak, bytes / HBM peak) per launch.", "",
             "| trunk | layer | launches | pixels | GFLOP | MB | flop/B | tensor us | hbm us | t_min us | bound |",
             "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
    for key in order:
        r = rows[key]
        tt, th = r["flops"] / (pt * 1e12) * 1e6, r["bytes"] / (ph * 1e9) * 1e6
        lines.append(f"| {r['trunk'][:-5]} | {r['name']} | {r['count']} | {r['pixels']} | {r['flops'] / 1e9:.2f} | "
                     f"{r['bytes'] / 1e6:.1f} | {r['flops'] / r['bytes']:.0f} | {tt:.1f} | {th:.1f} | {max(tt, th):.1f} | "
                     f"{'hbm' if th > tt else 'tensor'} |")
    for tr in ("RGBTrunk", "DepthTrunk"):
        sub = [w for w in work if w["trunk"] == tr]
        s = roofline.summarize(sub, 0.0, pt, ph)
        lines.append("")
        lines.append(f"**{tr}**: {s['launches']} conv launches, {s['algorithmic_gflop_per_step']:.1f} GFLOP, "
                     f"{s['algorithmic_gb_per_step']:.2f} GB; tensor-only {s['tensor_ms_per_step']:.3f} ms, HBM-only "
                     f"{s['hbm_ms_per_step']:.3f} ms, per-launch roofline {s['ideal_ms_per_step']:.3f} ms "
                     f"({s['hbm_bound_launches']} of {s['launches']} launches HBM-bound).")
    lines.append("")
    lines.append(f"**Both trunks**: {summ['launches']} launches, {summ['algorithmic_gflop_per_step']:.1f} GFLOP, "
                 f"{summ['algorithmic_gb_per_step']:.2f} GB per step; tensor-only {summ['tensor_ms_per_step']:.3f} ms, "
                 f"HBM-only {summ['hbm_ms_per_step']:.3f} ms, per-launch roofline **{summ['ideal_ms_per_step']:.3f} ms**; "
                 f"aggregate bound: {summ['bound']}.")
    if args.measured_conv_ms:
        lines.append(f"Measured conv device time (bench.py `roofline.conv_ms_per_step`, serial launches): "
                     f"{args.measured_conv_ms:.3f} ms -> **{100 * summ['frac_of_ideal']:.0f} % of the per-launch roofline** "
                     f"({summ['achieved_gbs']:.0f} GB/s of algorithmic traffic, "
                     f"{summ['algorithmic_gflop_per_step'] / args.measured_conv_ms:.0f} TFLOP/s).")
    text = "\n".join(lines) + "\n"
    print(text)
    if args.md:
        open(os.path.join(ROOT, args.md), "w").write(text)
    print(json.dumps(summ))
    if args.launch_csv:
        gap = gap_table(work, args.launch_csv, pt, ph)
        print(gap)
        if args.gap_md:
            open(os.path.join(ROOT, args.gap_md), "w").write(gap)


def gap_table(work, csv_path, pt, ph):
    """Measured (ncu, cold-cache, serialised) duration of every conv launch against its roofline bound, grouped by
    layer shape and ranked by the time lost."""
    import csv
    import io
    import re

    rows = list(csv.DictReader(io.StringIO("".join(l for l in open(csv_path) if l.startswith('"')))))
    conv = [r for r in rows if any(k in r["Kernel Name"] for k in ("conv_umma_kernel", "conv_tma_kernel", "conv3x3_halo_kernel"))]
    assert len(conv) == len(work), (len(conv), len(work))
    groups, order = {}, []
    for w, r in zip(work, conv):
        m = re.search(r"(conv\w*_kernel<[^>]*>)", r["Kernel Name"])
        kern = m.group(1).replace("_kernel", "") if m else r["Kernel Name"][:40]
        key = (w["trunk"], w["name"], w["pixels"])
        if key not in groups:
            groups[key] = dict(w, count=0, us=0.0, kernels=set())
            order.append(key)
        g = groups[key]
        g["count"] += 1
        g["us"] += float(r["Metric Value"]) / 1e3
        g["kernels"].add(kern)
    out = ["# measured conv launches vs their roofline bound (round-1 launch list, one serial eager step under ncu)", "",
           "Durations are ncu `gpu__time_duration.sum` of `profiles/r01d_launches_step.md`'s capture (cold caches, serialised "
           "launches: read them as shares); `t_min` as in `r01e_layer_roofline.md`.  Sorted by the time lost per step.", "",
           "| trunk | layer | launches | kernel picked | measured us (sum) | t_min us (sum) | % of roofline | lost us |",
           "|---|---|---:|---|---:|---:|---:|---:|"]
    tab = []
    for key in order:
        g = groups[key]
        tmin = max(g["flops"] / (pt * 1e12), g["bytes"] / (ph * 1e9)) * 1e6 * g["count"]
        tab.append((g["us"] - tmin, g, tmin))
    tab.sort(key=lambda t: -t[0])
    tot_m = tot_i = 0.0
    for lost, g, tmin in tab:
        tot_m += g["us"]
        tot_i += tmin
        out.append(f"| {g['trunk'][:-5]} | {g['name']} | {g['count']} | {', '.join(sorted(g['kernels']))} | {g['us']:.1f} | "
                   f"{tmin:.1f} | {100 * tmin / g['us']:.0f} | {lost:.1f} |")
    for tr in ("RGBTrunk", "DepthTrunk"):
        m = sum(g["us"] for _, g, _ in tab if g["trunk"] == tr)
        i = sum(t for _, g, t in tab if g["trunk"] == tr)
        out.append("")
        out.append(f"**{tr}** conv launches: measured {m:.0f} us, roofline {i:.0f} us ({100 * i / m:.0f} %).")
    out.append("")
    out.append(f"**All conv launches**: measured {tot_m:.0f} us, roofline {tot_i:.0f} us ({100 * tot_i / tot_m:.0f} %).")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    main()
