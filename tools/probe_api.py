#!/usr/bin/env python3
"""Where the host time of NBodySimulator.run(400) / step() at N = 200 goes (configs[0]): cProfile of 20 calls plus
wall-clock of the pieces.  Development aid; run on the GPU box."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import ics  # noqa: E402
from hpc.nbody import NBodySimulator  # noqa: E402

m32 = ics.shared_masses(200, 42)


def make():
    sim = NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42)
    sim.masses = m32.copy()
    sim.accelerations = sim._compute_accelerations()
    return sim


make().run(400, verbose=False)
sims = [make() for _ in range(20)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in sims:
    s.run(400, verbose=False)
print(f"run(400): {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per call")

sims = [make() for _ in range(20)]
pr = cProfile.Profile()
pr.enable()
for s in sims:
    s.run(400, verbose=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

# pieces
sim = make()
rs = sim._device_state()
for n_steps in (400, 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        w = rs.advance_async(n_steps, 1, snapshots=True)
        t1 = time.perf_counter()
        w()
    torch.cuda.synchronize()
    print(f"advance_async({n_steps}) + wait: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    rs.advance_async(400, 1, snapshots=False)
e1.record()
torch.cuda.synchronize()
print(f"kernel only, 400 steps no snapshots: {e0.elapsed_time(e1) / 20:.3f} ms")
snaps = torch.empty((3, 401, 200, 3), dtype=torch.float64, device="cuda")
t0 = time.perf_counter()
for _ in range(20):
    sim._engine().to_host(snaps)
print(f"to_host of the 3 stacks: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(200):
    sim._host_touched = True
    sim._device_state()
print(f"upload (ResidentSystem build): {(time.perf_counter() - t0) / 200 * 1e6:.1f} us")
sim = make()
sim.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    sim.step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"step(): host {(t1 - t0) / 2000 * 1e6:.2f} us per call, with drain {(t2 - t0) / 2000 * 1e6:.2f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    sim.step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
