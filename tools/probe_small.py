#!/usr/bin/env python3
"""Single-system step rate at small/mid N: one-launch ensemble kernel (K3, B=1) vs per-step kernels (K1/K2)."""
import sys, time, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics
eng = _cuda.get_engine()
for dtype in (np.float64, np.float32):
    for n in (64, 128, 200, 256, 384, 512, 768, 1024, 2048, 4096, 8192):
        x, v, m = ics.plummer_ic(n, seed=7)
        a = eng.accelerations(x, m, 0.01, dtype)
        res = {}
        for path, thr in (("K3", 10**9), ("K2", 0)):
            if path == "K3" and n > 1024:
                continue
            _cuda.SMALL_SYSTEM_MAX_BODIES = thr
            steps = 200
            eng.run(x, v, a, m, 1e-3, 0.01, 20, 20, dtype=dtype, snapshots=False)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            eng.run(x, v, a, m, 1e-3, 0.01, steps, steps, dtype=dtype, snapshots=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            res[path] = round(dt / steps * 1e6, 2)
        print(json.dumps({"n": n, "dtype": np.dtype(dtype).name, "us_per_step": res}), flush=True)
