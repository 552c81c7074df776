#!/usr/bin/env python3
"""Single-system leapfrog rate across N (K2, one launch per step): ms per step, interactions/s, fraction of the pipe peak."""
import sys, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics
from hpc.sharded import ShardedSystem
eng = _cuda.get_engine()
sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384, 32768, 65536, 131072]
for dtype, lanes in ((np.float32, 128), (np.float64, 64)):
    peak = eng.sm_count * lanes * 2 * 1.965e9 / 20.0
    for n in sizes:
        x, v, m = ics.plummer_ic(n, seed=7)
        s = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=0, world=1, rank=0)
        steps = max(5, min(400, int(2e12 / (n * float(n)))))
        s.advance(steps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            e0.record(); s.advance(steps); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
        rate = n * (n - 1.0) / (best * 1e-3)
        print(json.dumps({"n": n, "dtype": np.dtype(dtype).name, "ms_per_step": round(best, 4),
                          "Ginter_per_s": round(rate / 1e9, 1), "frac_pipe_peak": round(rate / peak, 3)}), flush=True)
