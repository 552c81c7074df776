#!/usr/bin/env python3
"""Mid-size single systems, leapfrog step time of the three step kernels: K2s "group" (nb_group.cu: one CTA per group
of bodies, shared-memory reduction; the default where the groups fit one wave), K2 "tiles" (nb_force.cu, NB_NO_GROUP=1)
and K2p "one_launch" (nb_persist.cu, NB_PERSIST=1); us per step, interactions/s, fraction of the 20-flop pipe peak."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics  # noqa: E402
from hpc.sharded import ShardedSystem  # noqa: E402

eng = _cuda.get_engine()
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 12288, 16384, 18000, 32768]
for dtype, lanes in ((np.float32, 128), (np.float64, 64)):
    peak = eng.sm_count * lanes * 2 * 1.965e9 / 20.0
    for n in sizes:
        x, v, m = ics.plummer_ic(n, seed=7)
        row = {"n": n, "dtype": np.dtype(dtype).name}
        for mode in ("group", "tiles", "one_launch"):
            for key in ("NB_PERSIST", "NB_NO_GROUP"):
                os.environ.pop(key, None)
            if mode != "group":
                os.environ["NB_NO_GROUP"] = "1"      # K2: (i-tile x segment) grid, one launch per step
            if mode == "one_launch":
                os.environ["NB_PERSIST"] = "1"       # K2p: all steps in one cooperative launch
            s = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=0)
            steps = max(10, min(400, int(2e12 / (n * float(n)))))
            s.advance(steps)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e30
            for _ in range(3):
                e0.record(); s.advance(steps); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / steps)
            eng.step_status(s.ws, n)
            rate = n * (n - 1.0) / (best * 1e-3)
            row[mode] = {"us_per_step": round(best * 1e3, 2), "Ginter_per_s": round(rate / 1e9, 1),
                         "frac_pipe_peak": round(rate / peak, 3)}
        print(json.dumps(row), flush=True)
