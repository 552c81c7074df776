#!/usr/bin/env python3
"""A few systems of N = 200, 400 steps, snapshots every step: clusters (8 / 4 / 2 CTAs per system) vs one CTA each."""
import os, sys, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics
eng = _cuda.get_engine()
for B in (1, 10, 18, 19, 37, 38, 74, 75, 148):
    x0, v0, m32 = ics.datagen_ensemble_ic(B, 200, seed=1)
    x, v = eng.to_device(x0), eng.to_device(v0)
    a = torch.zeros_like(x)
    m_d, f32 = eng._masses_dev(m32)
    ox = torch.empty((B, 401, 200, 3), dtype=torch.float64, device=eng.device)
    ov, oa = torch.empty_like(ox), torch.empty_like(ox)
    res = {}
    for mode in ("cluster", "one-CTA"):
        os.environ.pop("NB_ENSEMBLE_NO_CLUSTER", None)
        if mode == "one-CTA":
            os.environ["NB_ENSEMBLE_NO_CLUSTER"] = "1"
        best = 1e30
        for _ in range(4):
            x.copy_(eng.to_device(x0)); v.copy_(eng.to_device(v0))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.ensemble_device(x, v, a, m_d, f32, 0, B, 200, 1e-3, 1e-9, 400, 1, np.float64, True, True, ox, ov, oa, 401, 0)
            e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[mode] = round(best, 3)
    print(json.dumps({"B": B, "ms_per_400_steps": res}), flush=True)
