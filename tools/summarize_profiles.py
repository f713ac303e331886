#!/usr/bin/env python
"""Turn the raw ncu outputs of a GPU visit (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_step.csv profiles/r01c_launches_step.md
    python tools/summarize_profiles.py full gpurun_out/x.ncu-rep profiles/r01c_conv_tma_ncu_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
]


def short(k):
    k = re.sub(r"void |vlnce::|<unnamed>::|unnamed>::|\(anonymous namespace\)::", "", k)
    return re.sub(r"\(.*", "", k)


def launches(src, dst, note=""):
    txt = open(src).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        us = float(r["Metric Value"]) / 1e3
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
        n += 1
    conv = sum(us for k, (_, us) in agg.items() if k.startswith("conv") or k.startswith("depth_fused"))
    # DRAM traffic of the same launches (when the pass also collected dram__bytes_*): mean per conv-class launch ->
    # profiles/conv_traffic.json, the source of bench.py's roofline.traffic
    dram, ids = collections.defaultdict(float), {}
    for r in rows:
        if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            unit = r["Metric Unit"].lower()
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            dram[r["ID"]] += float(r["Metric Value"]) * scale
            ids[r["ID"]] = short(r["Kernel Name"])
    conv_ids = [i for i, k in ids.items() if k.startswith("conv") or k.startswith("depth_fused")]
    if conv_ids:
        import json
        import os

        mean = sum(dram[i] for i in conv_ids) / len(conv_ids)
        json.dump({"kernel": "tcgen05 conv-class launches of one eager training step (conv_tma / conv_umma / "
                             "conv3x3_halo / depth_fused), " + os.path.basename(dst),
                   "launches": len(conv_ids), "dram_bytes_per_launch": mean,
                   "dram_bytes_per_step": sum(dram[i] for i in conv_ids),
                   "note": "mean of dram__bytes_read.sum + dram__bytes_write.sum over exactly the launches whose "
                           "algorithmic bytes bench.py sums (same pass as the launch list; serial launches, caches warm "
                           "from the producing layer)"},
                  open(os.path.join(os.path.dirname(dst), "conv_traffic.json"), "w"), indent=1)
    with open(dst, "w") as f:
        f.write("# launch list of one training step (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
        f.write("Command: `ncu --profile-from-start off --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --clock-control none --csv "
                "python bench.py --ncu-step --no-cpu-baseline` (one eager, serial-stream step between "
                "cudaProfilerStart/Stop; B200, batch 64, bf16/fp16 trunks).  Cold-cache serialised launch times: "
                "compare SHARES, not absolutes.\n" + note + "\n")
        f.write(f"One step = {n} launches, {tot:.0f} us summed device time.\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |\n")
        f.write(f"\ntcgen05 conv tiles (`conv_tma_kernel` + `conv_umma_kernel` + `conv3x3_halo_kernel` + `depth_fused_kernel`): {conv:.0f} us = "
                f"{100 * conv / tot:.1f}% of the step (bench.py's live CUDA-event measurement of the same launches: "
                "`roofline.conv_ms_per_step`).\n")


def full(src, dst, note=""):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full captures\n\n" + note + "\n\n")
        for i, r in enumerate(rows[2:]):
            f.write(f"## launch {i}: `{short(r[hdr.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in hdr:
                    j = hdr.index(m)
                    f.write(f"| `{m}` | {r[j]} | {units[j]} |\n")
            f.write("\n")


KERNEL_METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,"
                  "sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,"
                  "lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed")


def kernels(src, dst, note=""):
    """Per-kernel evidence table of one eager training step: duration, DRAM bytes -> achieved HBM GB/s and its share
    of the measured copy bandwidth, tensor-pipe active %, SM / L2 / L1 throughput (time-weighted means over the
    kernel's launches).  ``src`` = csv of `ncu --metrics <KERNEL_METRICS> --csv`."""
    import json
    import os

    txt = open(src).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)   # -> us
        d[r["Metric Name"]] = v
    peak = 6534.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["kernel"], {"n": 0, "us": 0.0, "bytes": 0.0, "w": collections.defaultdict(float), "grids": set()})
        us = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["us"] += us
        a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["grids"].add(d["grid"].replace(" ", ""))
        for m in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                  "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"):
            a["w"][m] += d.get(m, 0.0) * us
    tot = sum(a["us"] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# every kernel of one training step under ncu: HBM GB/s, tensor-pipe active %, unit throughputs\n\n")
        f.write("Command: `ncu --profile-from-start off --clock-control none --csv --metrics " + KERNEL_METRICS +
                " python bench.py --ncu-step --no-cpu-baseline` (one eager, serial-stream step of C2: batch 64, bf16 / "
                "fp16 trunks, between cudaProfilerStart/Stop; B200).  Launches are serialised and replayed per metric "
                "pass, so durations are cold-ish and the table is for SHARES and per-kernel unit utilisation, not step "
                "time.  HBM GB/s = (dram read + write bytes) / duration, summed over the kernel's launches; % of the "
                f"measured copy bandwidth {peak:.0f} GB/s (MEASURED_PEAKS.json).  Percentages are time-weighted means.\n"
                + note + "\n\n")
        f.write(f"One step = {sum(a['n'] for a in agg.values())} launches, {tot:.0f} us summed.\n\n")
        f.write("| kernel | launches | us | share | DRAM MB | HBM GB/s | % HBM peak | tensor pipe % | SM thr % | warps active % | L2 thr % | L1 thr % |\n")
        f.write("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            us = max(a["us"], 1e-9)
            gbs = a["bytes"] / (us * 1e-6) / 1e9
            w = a["w"]
            f.write(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['bytes'] / 1e6:.1f} | {gbs:.0f} | "
                    f"{100 * gbs / peak:.1f} | {w['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'] / us:.1f} | "
                    f"{w['sm__throughput.avg.pct_of_peak_sustained_elapsed'] / us:.1f} | "
                    f"{w['sm__warps_active.avg.pct_of_peak_sustained_active'] / us:.1f} | "
                    f"{w['lts__throughput.avg.pct_of_peak_sustained_elapsed'] / us:.1f} | "
                    f"{w['l1tex__throughput.avg.pct_of_peak_sustained_elapsed'] / us:.1f} |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full, "kernels": kernels}[sys.argv[1]](*sys.argv[2:])
