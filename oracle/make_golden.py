"""Freeze golden vectors for the parity cases by running the UNMODIFIED reference
model files (``/root/reference``; build container only).  TEST INFRASTRUCTURE.

    python -m oracle.make_golden            # writes tests/golden/<case>.pt

Each fixture stores the outputs only; weights and inputs are regenerated from the
seeds in ``oracle/cases.py`` (``vlnce_b200.synth``), plus a checksum of both so a
drifting RNG is detected instead of silently mis-compared.
"""
from __future__ import annotations

import os
import sys

import torch

from . import cases, reference_import


def checksum(case):
    pol = cases.build_policy(cases.oracle_impl(case), case)
    w = sum(float(v.double().abs().sum()) for v in pol.state_dict().values())
    inp = cases.inputs(case)
    i = 0.0
    for part in inp:  # observations + every other tensor of the batch
        for v in (part.values() if isinstance(part, dict) else [part]):
            i += float(v.double().abs().sum())
    if case["kind"] in ("act", "train"):  # fixtures frozen before the waypoint cases: observations only
        i = sum(float(v.double().abs().sum()) for v in inp[0].values())
    return torch.tensor([w, i], dtype=torch.float64)


def main(out_dir="tests/golden", only=None):
    assert reference_import.available(), "needs /root/reference (build container)"
    ref = reference_import.load()
    os.makedirs(out_dir, exist_ok=True)
    for name, case in cases.CASES.items():
        if only and name not in only:
            continue
        res = cases.run_case(ref, name, aux=ref.AuxLosses)
        res["_checksum"] = checksum(case)
        res["_source"] = "reference@729d141b via oracle/reference_import.py, torch " + torch.__version__
        path = os.path.join(out_dir, name + ".pt")
        torch.save(res, path)
        print(name, {k: (tuple(v.shape) if hasattr(v, "shape") else v) for k, v in list(res.items())[:4]},
              os.path.getsize(path))


if __name__ == "__main__":
    main(only=set(sys.argv[1:]) or None)
