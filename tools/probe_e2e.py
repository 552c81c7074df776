#!/usr/bin/env python3
"""End-to-end ensemble time vs number of D2H chunks (development aid)."""
import os, sys, time, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import ics
from hpc.ensemble import simulate_ensemble
x0, v0, m32 = ics.datagen_ensemble_ic(300, 200, seed=42)
for chunks in (1, 2, 4, 8, 16, 32, 8):
    os.environ["NBODY_D2H_CHUNKS"] = str(chunks)
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = simulate_ensemble(x0, v0, m32, dt=1e-3, n_steps=400)
        ts.append((time.perf_counter() - t0) * 1e3); del out
    print(json.dumps({"chunks": chunks, "ms": [round(t, 2) for t in ts]}), flush=True)
# raw pinned D2H bandwidth for reference
d = torch.empty(1732320000 // 8, dtype=torch.float64, device="cuda"); h = torch.empty_like(d, device="cpu", pin_memory=True)
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    print(json.dumps({"raw_d2h_GBps": round(1.73232 / (time.perf_counter() - t0), 2)}), flush=True)
