#!/usr/bin/env python3
"""K4 (energy) time and FP64-pipe fraction vs N: N^2 pair terms of 13 FP64 operations each (3 sub, 3 fma, 5 for the
refined reciprocal square root, 1 mul, 1 add)."""
import sys, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics
eng = _cuda.get_engine()
peak_ops = eng.sm_count * 64 * 1.965e9
for n in [int(a) for a in sys.argv[1:]] or [200, 2048, 16384, 65536, 262144]:
    x, v, m = ics.plummer_ic(n, seed=7)
    pos_d, vel_d = eng.to_device(x), eng.to_device(v)
    m_d, f32 = eng._masses_dev(m)
    eng.energy_slab(pos_d, vel_d, m_d, f32, n, 0, n, 0.01)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if n > 100000 else 10
    e0.record()
    for _ in range(reps):
        ku = eng.energy_slab(pos_d, vel_d, m_d, f32, n, 0, n, 0.01)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"n": n, "ms": round(ms, 4), "Gpairs_per_s": round(n * float(n) / ms / 1e6, 1),
                      "frac_fp64_pipe_13ops": round(n * float(n) * 13 / (ms * 1e-3) / peak_ops, 3)}), flush=True)
