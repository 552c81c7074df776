import sys, time; sys.path.insert(0, "nbody-gnn-hpc_b200")
import numpy as np, torch
from hpc import _cuda, ics
from hpc.ensemble import simulate_ensemble
eng = _cuda.get_engine()
for n, B, steps in ((2000, 40, 100), (1100, 60, 100), (5000, 10, 50)):
    x0 = np.empty((B, n, 3)); v0 = np.empty((B, n, 3))
    for b in range(B): x0[b], v0[b], m = ics.plummer_ic(n, seed=b)
    for dtype in ("float64", "float32"):
        simulate_ensemble(x0[:2], v0[:2], m, dt=1e-3, softening=0.01, n_steps=3, dtype=dtype)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = simulate_ensemble(x0, v0, m, dt=1e-3, softening=0.01, n_steps=steps, save_interval=steps, dtype=dtype)
        torch.cuda.synchronize(); tb = time.perf_counter() - t0
        t0 = time.perf_counter()
        for b in range(B):
            a0 = eng.accelerations(x0[b], m, 0.01, np.dtype(dtype))
            eng.run(x0[b], v0[b], a0, m, 1e-3, 0.01, steps, steps, dtype=np.dtype(dtype))
        torch.cuda.synchronize(); ts = time.perf_counter() - t0
        inter = B * steps * n * (n - 1.0)
        print(f"N={n} B={B} steps={steps} {dtype}: batched {tb*1e3:.1f} ms ({inter/tb/1e9:.0f} G int/s), one by one {ts*1e3:.1f} ms ({inter/ts/1e9:.0f} G int/s)")
