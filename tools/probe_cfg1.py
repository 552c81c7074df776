#!/usr/bin/env python3
"""configs[0]: the reference's README default -- one simulation, N = 200, 400 steps, float64 -- through the
NBodySimulator API (host state in, list of 401 host state dicts out), next to the CPU oracle port on one thread."""
import sys, time, json, cProfile, pstats, io
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc.nbody import NBodySimulator
import oracle

def once():
    sim = NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42)
    sim.masses = np.random.RandomState(42).uniform(1e10, 1e12, 200).astype(np.float32)
    sim.accelerations = sim._compute_accelerations()
    return sim, sim.run(400, save_interval=1, verbose=False)

once()
t = []
for _ in range(5):
    t0 = time.perf_counter(); sim, states = once(); t.append(time.perf_counter() - t0)
sim = NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42)
sim.run(400, save_interval=1, verbose=False)
r = []
for _ in range(5):
    t0 = time.perf_counter(); sim.run(400, save_interval=1, verbose=False); r.append(time.perf_counter() - t0)
x, v, m = sim.positions.copy(), sim.velocities.copy(), sim.masses.copy()
oracle.build()
import os
os.environ["OMP_NUM_THREADS"] = "1"
a = oracle.accel_direct(x, m, 1e-9)
t0 = time.perf_counter(); oracle.run(x, v, a, m, 1e-3, 1e-9, 400, 1); cpu = time.perf_counter() - t0
print(json.dumps({"construct_plus_run_ms": round(1e3 * min(t), 3), "run_400_steps_ms": round(1e3 * min(r), 3),
                  "states": len(states), "cpu_oracle_run_ms_all_threads_default": round(1e3 * cpu, 2)}))
pr = cProfile.Profile(); pr.enable(); sim.run(400, save_interval=1, verbose=False); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(12); print(s.getvalue()[:2500])
