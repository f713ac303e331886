"""Cycle-level trace of one persistent conv_tma launch (CTA 0): where each role waits.
Usage: python tools/trace_conv.py [layer] [code]   layer in {l3c1, l2c2, stem}"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlnce_b200 import lib  # noqa: E402
from vlnce_b200.lib import C, ptr  # noqa: E402

MT2, NOSTAT = 1 << 21, 1 << 22
EV = {1: "A:want_a", 2: "A:a_free", 3: "A:a_issued", 4: "B:b_free", 5: "B:b_issued", 10: "M:acc_free", 11: "M:a_full",
      12: "M:b_full", 13: "M:issued", 20: "E:want_acc", 21: "E:acc_full", 22: "E:staged", 23: "E:barrier", 24: "S:stored", 25: "E:stg_ready", 26: "E:tmem_ld"}


def main():
    layer = sys.argv[1] if len(sys.argv) > 1 else "l3c1"
    code = eval(sys.argv[2]) if len(sys.argv) > 2 else (128 | (4 << 16) | NOSTAT)
    dev = "cuda"
    n = 64
    geo = {"l3c1": (16, 1024, 256, 1), "l2c2": (32, 128, 128, 3), "l1c1": (64, 256, 64, 1), "l1c2": (64, 64, 64, 3),
           "l1c3": (64, 64, 256, 1, True), "l2c3": (32, 128, 512, 1, True), "l3c3": (16, 256, 1024, 1, True),
           "l3c2": (16, 256, 256, 3), "l4c3": (8, 512, 2048, 1, True)}
    trace = torch.zeros(5 * 768 * 2, dtype=torch.int64, device=dev)
    if layer == "stem":
        img = torch.randint(0, 256, (n, 256, 256, 3), device=dev, dtype=torch.uint8)
        cells = torch.empty(n, 131, 131, 16, device=dev, dtype=torch.bfloat16)
        C.vlnce_stem_s2d(ptr(img), lib.U8, ptr(cells), lib.BF16, n, 256, 256, lib.stream())
        w = (torch.randn(64, 256, device=dev) * 0.05).bfloat16()
        scale, shift = torch.ones(64, device=dev), torch.zeros(64, device=dev)
        y = torch.empty(n, 128, 128, 64, device=dev, dtype=torch.bfloat16)

        def launch():
            C.vlnce_conv_stem_tma(ptr(cells), ptr(w), lib.BF16, ptr(scale), ptr(shift), ptr(y), lib.BF16, n, 128, 128,
                                  64, 1, code, lib.stream())
    else:
        hw, cin, cout, r = geo[layer][:4]
        res = torch.randn(n, hw, hw, cout, device=dev).bfloat16() if len(geo[layer]) > 4 else None
        x = torch.randn(n, hw, hw, cin, device=dev).bfloat16()
        w = (torch.randn(cout, r, r, cin, device=dev) * 0.05).bfloat16()
        scale, shift = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        y = torch.empty(n, hw, hw, cout, device=dev, dtype=torch.bfloat16)

        def launch():
            C.vlnce_conv_tma(ptr(x), ptr(w), lib.BF16, ptr(scale), ptr(shift), ptr(res), ptr(y), lib.BF16, n, hw, hw, cin,
                             cout, r, 1, code, lib.stream())
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    C.vlnce_debug_set_trace(ptr(trace))
    launch()
    torch.cuda.synchronize()
    C.vlnce_debug_set_trace(None)
    t = trace.cpu().view(5, 768, 2)
    recs = []
    for role in range(5):
        for i in range(768):
            tag, clk = int(t[role, i, 0]), int(t[role, i, 1])
            if clk == 0:
                break
            recs.append((clk, tag >> 32, tag & 0xffffffff))
    recs.sort()
    t0 = recs[0][0]
    print(f"layer {layer} code {code:#x}: {len(recs)} records, span {recs[-1][0] - t0} cycles")
    last = {}
    for clk, ev, idx in recs[:int(os.environ.get("TRACE_LINES", "260"))]:
        role = EV[ev][0]
        d = clk - last.get(role, t0)
        last[role] = clk
        print(f"{clk - t0:8d} (+{d:6d}) {EV[ev]:12s} {idx}")


if __name__ == "__main__":
    main()
