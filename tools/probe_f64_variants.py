#!/usr/bin/env python3
"""Time the two float64 hot kernels of ONE build of the library (NBODY_B200_LIB selects it): the 300x200x400 ensemble
(K3, two lanes) and one N = 65,536 force evaluation (K1).  Used to compare experiment builds (build.py --out ... -D ...)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics  # noqa: E402

eng = _cuda.get_engine()
B, N, T = 300, 200, 400
x0, v0, m32 = ics.datagen_ensemble_ic(B, N, seed=42)
x0_d, v0_d = eng.to_device(x0), eng.to_device(v0)
x, v, a = x0_d.clone(), v0_d.clone(), torch.zeros_like(x0_d)
m_d, f32 = eng._masses_dev(m32)
out = tuple(torch.empty((B, T + 1, N, 3), dtype=torch.float64, device=eng.device) for _ in range(3))
ms = []
for rep in range(6):
    x.copy_(x0_d); v.copy_(v0_d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.ensemble_device(x, v, a, m_d, f32, 0, B, N, 1e-3, 1e-9, T, 1, np.float64, True, True, *out, T + 1, 0)
    e1.record(); e1.synchronize()
    ms.append(e0.elapsed_time(e1))
ens = min(ms[2:])
n = 65536
xs, _, m = ics.plummer_ic(n, seed=7)
pos_d = eng.to_device(xs)
md, mf = eng._masses_dev(m)
stream = eng.pack(pos_d, md, mf, n, np.float64)
ws = eng.workspace(n, n, np.float64)
for _ in range(3):
    eng.accel_slab(stream, n, 0, n, 0.01, ws)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.accel_slab(stream, n, 0, n, 0.01, ws)
e1.record(); e1.synchronize()
k1 = e0.elapsed_time(e1) / 10
print(f"{os.environ.get('NBODY_B200_LIB', 'default'):60s} ensemble 300x200x400: {ens:.3f} ms   force N=65536 f64: {k1:.3f} ms")
