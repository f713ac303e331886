"""B200: error of the TF32 tensor-core instruction projection (engine.linear_tc) against the golden of
cma_rxr_train_features_l512 (precomputed trunk features: every other op of the step is fp32)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
import vlnce_b200
from vlnce_b200 import engine
from oracle import cases

name = "cma_rxr_train_features_l512"
gold = torch.load(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", name + ".pt"), weights_only=False)
for tc_rows in (2048, 0):
    engine.HEAD_TC_MIN_ROWS = tc_rows
    got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, device="cuda", precision="bf16")
    rows = []
    for k, w in gold.items():
        if k.startswith("_") or k not in got:
            continue
        g, w = got[k].float(), w.float()
        if k.startswith("grad/") and w.numel() == 11 and g.numel() == 11:
            rel = abs(g[2] - w[2]).item() / max(w[2].item(), 1e-12)      # digest: norm
            kind = "norm-digest"
        else:
            rel = (g - w).norm().item() / max(w.norm().item(), 1e-12)
            kind = "rel-l2"
        rows.append((rel, k, kind))
    rows.sort(reverse=True)
    print("HEAD_TC_MIN_ROWS =", tc_rows)
    for rel, k, kind in rows[:8]:
        print(f"   {k:70s} {kind:12s} {rel:.3e}")
