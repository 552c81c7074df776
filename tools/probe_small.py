#!/usr/bin/env python3
"""Single-system step time at small/mid N: K3 on a cluster of 8 CTAs, K3 on one CTA, per-step kernels (K2)."""
import os, sys, time, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics
eng = _cuda.get_engine()
sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 200, 256, 384, 512, 768, 1024, 1536, 2048]
for dtype in (np.float64, np.float32):
    for n in sizes:
        x, v, m = ics.plummer_ic(n, seed=7)
        a = eng.accelerations(x, m, 0.01, dtype)
        res = {}
        for path in ("K3-cluster", "K3-one-CTA", "K2"):
            if path != "K2" and n > 1024:
                continue
            os.environ.pop("NB_ENSEMBLE_NO_CLUSTER", None)
            if path == "K3-one-CTA":
                os.environ["NB_ENSEMBLE_NO_CLUSTER"] = "1"
            _cuda.SMALL_SYSTEM_MAX_BODIES = 0 if path == "K2" else 10**9
            steps = 400
            eng.run(x, v, a, m, 1e-3, 0.01, 20, 20, dtype=dtype, snapshots=False)
            best = 1e30
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                eng.run(x, v, a, m, 1e-3, 0.01, steps, steps, dtype=dtype, snapshots=False)
                torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
            res[path] = round(best / steps * 1e6, 2)
        print(json.dumps({"n": n, "dtype": np.dtype(dtype).name, "us_per_step_incl_call_overhead": res}), flush=True)
