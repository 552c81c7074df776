#!/usr/bin/env python3
"""Tiny run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import ics
from hpc.ensemble import simulate_ensemble
from hpc.nbody import NBodySimulator, compute_accelerations_direct, compute_total_energy
for n in (33, 700):
    x, v, m = ics.plummer_ic(n, seed=1)
    for dt in ("float64", "float32"):
        compute_accelerations_direct(x, m, 0.01, dtype=dt)
        compute_accelerations_direct(x, m, 0.0, dtype=dt)
    compute_total_energy(x, v, m, 0.01)
np.random.seed(0)
for dt in ("float64", "float32"):
    sim = NBodySimulator(n_particles=700, box_size=10.0, dt=1e-3, seed=3, dtype=dt)
    sim.run(3, save_interval=2, verbose=False)
rng = np.random.RandomState(2)
for B, n in ((5, 37), (3, 600), (310, 16)):
    x0, v0, m = rng.rand(B, n, 3), rng.rand(B, n, 3), rng.uniform(1e9, 1e10, n)
    for dt in ("float64", "float32"):
        simulate_ensemble(x0, v0, m, dt=1e-3, softening=0.05, n_steps=24 if B > 300 else 4, save_interval=2, dtype=dt)
print("sanitize target ok")
