"""Per-launch roofline model of the conv tiles (measurement helper for bench.py / tools; no device work).

Every conv launch of the two trunks has an algorithmic cost in both currencies:

* flops  = 2 * (output pixels) * K * R * S * C                      (the 7x7x3 = 147 taps for the RGB stem);
* bytes  = the operands the algorithm has to move once: the input activation tile (for the stem: the image /
  pooled depth the stem operand is built from), the weights, the output tile in the dtype this path stores it in
  (16-bit activations; fp32 for GroupNorm raw outputs), and the residual tile when one is fused.

A launch cannot finish faster than max(flops / tensor peak, bytes / HBM peak); the sum of those per-launch bounds is
the roofline of the whole conv share of a step, and which of the two sums is larger names the bound.  With 16-bit
activations at batch 64 most of ResNet-50's layers at 56x56 / 28x28 are HBM-bound (K, C <= 256: < 150 flop/byte
against the machine's ~210), so a pure tensor-peak fraction understates how close those launches are to their
limit; the per-launch model is what DESIGN.md §6 and ``roofline.layer_model`` in the bench line report.
"""
from __future__ import annotations

from typing import Dict, List


def _esize(t) -> int:
    return int(t.element_size())


def conv_launch_work(trunk, x, shape_in, cv, y, kw) -> Dict[str, float]:
    """Algorithmic flops / bytes of one ``ConvTrunk._conv`` call (same arguments as the call)."""
    n, h, w, c = shape_in
    gemm_c = kw.get("gemm_c")
    residual = kw.get("residual")
    if kw.get("stem_s2d"):
        # x = 2x2 space-to-depth cells of the image, (h, w) = the stem's OUTPUT size: the algorithm reads the
        # (2h x 2w x 3) image once (counted at the cell image's 16-bit element size / 4 pixels per 16-channel cell
        # would overcount the zero padding: 3 channels * 4 pixels of 16)
        flops = 2.0 * n * h * w * cv.K * 147
        in_bytes = n * (2 * h) * (2 * w) * 3 * _esize(x)
        p, q = h, w
    elif kw.get("stem_rgb"):
        p, q = trunk._out_hw(h, w, cv)
        flops = 2.0 * n * p * q * cv.K * 147
        in_bytes = n * h * w * 3 * _esize(x)
    elif gemm_c is not None:
        # explicit stem operand (n*p*q, gemm_c) built by an im2col pre-pass: (h, w) = output size.  The algorithm's
        # input is what the operand was built from — 49 (depth) / 147 (rgb) taps over a stride-2 window, i.e. 4 source
        # pixels per output pixel and channel
        taps = 49 if cv.Cin == 1 else 147
        flops = 2.0 * n * h * w * cv.K * taps
        in_bytes = n * (2 * h) * (2 * w) * cv.Cin * _esize(x)
        p, q = h, w
    else:
        p, q = trunk._out_hw(h, w, cv)
        flops = 2.0 * n * p * q * cv.K * cv.R * cv.S * c
        in_bytes = n * h * w * c * _esize(x)
    w_bytes = cv.K * cv.R * cv.S * cv.Cin * _esize(cv.w)
    out_bytes = n * p * q * cv.K * _esize(y)
    res_bytes = n * p * q * cv.K * _esize(residual) if residual is not None else 0
    return {"flops": float(flops), "bytes": float(in_bytes + w_bytes + out_bytes + res_bytes),
            "trunk": type(trunk).__name__, "K": cv.K, "C": c if gemm_c is None else gemm_c, "R": cv.R,
            "pixels": n * p * q}


def summarize(work: List[Dict[str, float]], conv_ms_total: float, tensor_tflops: float, hbm_gbs: float,
              steps: int = 1) -> Dict[str, float]:
    """``work``: one entry per timed launch (``steps`` steps' worth); ``conv_ms_total``: their summed device time."""
    t_tensor = sum(x["flops"] for x in work) / (tensor_tflops * 1e12) * 1e3
    t_hbm = sum(x["bytes"] for x in work) / (hbm_gbs * 1e9) * 1e3
    t_ideal = sum(max(x["flops"] / (tensor_tflops * 1e12), x["bytes"] / (hbm_gbs * 1e9)) for x in work) * 1e3
    n_hbm = sum(1 for x in work if x["bytes"] / (hbm_gbs * 1e9) > x["flops"] / (tensor_tflops * 1e12))
    out = {
        "tensor_ms_per_step": t_tensor / steps, "hbm_ms_per_step": t_hbm / steps, "ideal_ms_per_step": t_ideal / steps,
        "bound": "hbm" if t_hbm > t_tensor else "tensor",
        "hbm_bound_launches": n_hbm // steps, "launches": len(work) // steps,
        "algorithmic_gb_per_step": sum(x["bytes"] for x in work) / steps / 1e9,
        "algorithmic_gflop_per_step": sum(x["flops"] for x in work) / steps / 1e9,
    }
    if conv_ms_total > 0:
        out["frac_of_ideal"] = t_ideal / conv_ms_total
        out["achieved_gbs"] = sum(x["bytes"] for x in work) / (conv_ms_total * 1e-3) / 1e9
    return out


def assemble(conv_ms_total: float, conv_flops_total: float, work: List[Dict[str, float]], peaks: Dict[str, float],
             traffic, launches_timed: int, steps: int = 3) -> Dict[str, object]:
    """The ``roofline`` object of the bench line for the dominant kernel family (the tcgen05 conv tiles).

    ``achieved`` / ``peak`` / ``frac`` = algorithmic conv flops / summed conv launch time against the sustained bf16
    tensor throughput (the conv trunks are tensor-core class).  The unfused per-launch byte model is kept
    beside it under ``hbm`` and ``layer_model`` (``frac_of_ideal`` = sum of per-launch bounds / measured time)."""
    pt = float(peaks.get("bf16_tflops_sustained", 1400.0))
    ph = float(peaks.get("hbm_gbs", 6500.0))
    src = "MEASURED_PEAKS.json" if peaks else "fallback (1.4 PFLOP/s, 6.5 TB/s)"
    tflops = conv_flops_total / (conv_ms_total * 1e-3) / 1e12
    roof = {"kernel": f"tcgen05 conv tiles: conv_tma_kernel / conv_umma_kernel / conv3x3_halo_kernel (RGB trunk) + "
                      f"depth_fused_kernel (whole depth trunk, one launch): {launches_timed // steps} launches/step",
            "bound": "tensor", "achieved": tflops, "peak": pt, "unit": "TFLOP/s", "frac": tflops / pt, "traffic": traffic,
            "peak_source": src + " bf16_tflops_sustained", "launches_timed": launches_timed,
            "conv_ms_per_step": conv_ms_total / steps, "algorithmic_gflop_per_step": conv_flops_total / steps / 1e9}
    if work and len(work) == launches_timed:
        lm = summarize(work, conv_ms_total, pt, ph, steps=steps)
        roof["layer_model"] = lm
        roof["tensor"] = {"achieved": tflops, "peak": pt, "unit": "TFLOP/s", "frac": tflops / pt}
        roof["hbm"] = {"achieved": lm["achieved_gbs"], "peak": ph, "unit": "GB/s", "frac": lm["achieved_gbs"] / ph}
        # SURVEY 8d classes the conv trunks as tensor-core work: the headline pair stays the tensor one.  The unfused
        # per-launch byte model (every conv reads its input and writes its output once) is a work list, not an
        # algorithmic minimum — fusing 1x1 chains removes bytes — so it is reported beside, never instead.
        roof["algorithmic_bytes_per_launch"] = lm["algorithmic_gb_per_step"] * 1e9 / max(lm["launches"], 1)
    return roof
