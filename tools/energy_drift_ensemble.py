#!/usr/bin/env python3
"""Relative energy and momentum drift of EVERY trajectory of a data-generation-sized ensemble (300 x 200 bodies x 400
steps), float64 and float32 engines, evaluated on the device by K4b (hpc.metrics.snapshot_energies) from the snapshot
stacks K3 leaves in HBM -- the ensemble form of the deliverable "relative energy-drift curves overlaid".

    python tools/energy_drift_ensemble.py [--ics plummer|default] [--out profiles/r02_energy_drift_ensemble.json]

Per step: median and maximum over the 300 systems of |E_k - E_0| / |E_0|, and of |p_k - p_0| / sum m |v|."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import ics  # noqa: E402
from hpc.ensemble import simulate_ensemble  # noqa: E402
from hpc.metrics import snapshot_energies  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ics", default="plummer", choices=["plummer", "default"])
    ap.add_argument("--systems", type=int, default=300)
    ap.add_argument("--bodies", type=int, default=200)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    B, N, T = a.systems, a.bodies, a.steps
    if a.ics == "plummer":
        x0, v0 = np.empty((B, N, 3)), np.empty((B, N, 3))
        for b in range(B):
            x0[b], v0[b], m = ics.plummer_ic(N, seed=1000 + b)
        eps = 0.01
    else:
        x0, v0, m = ics.datagen_ensemble_ic(B, N, seed=42)
        eps = 1e-9
    res = {"N": N, "systems": B, "n_steps": T, "dt": 1e-3, "softening": eps, "gpus": 1,
           "ics": "Plummer (G*M = 1), seeds 1000.." if a.ics == "plummer" else "reference-default (seeds 42.., shared float32 masses)",
           "steps": list(range(T + 1)), "curves": {}}
    for tag, dtype in (("f64", "float64"), ("f32", "float32")):
        t0 = time.perf_counter()
        out = simulate_ensemble(x0, v0, m, dt=1e-3, softening=eps, n_steps=T, dtype=dtype, outputs="device")
        e = snapshot_energies(out["positions"], out["velocities"], m, softening=eps)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        E = e["total"]                                             # (B, T+1)
        drift = np.abs(E - E[:, :1]) / np.abs(E[:, :1])
        p = e["momentum"]                                          # (B, T+1, 3)
        scale = (np.asarray(m, dtype=np.float64)[None, :, None] * np.abs(v0)).sum(axis=(1, 2))     # sum m |v| at t = 0
        pdrift = np.linalg.norm(p - p[:, :1], axis=-1) / scale[:, None]
        res["curves"][tag] = {"rel_drift": np.median(drift, axis=0).tolist(), "rel_drift_max": drift.max(axis=0).tolist(),
                              "momentum_drift_median": np.median(pdrift, axis=0).tolist(),
                              "momentum_drift_max": pdrift.max(axis=0).tolist(), "wall_s": round(wall, 3)}
        print(f"{tag}: median |dE/E0| at step {T}: {np.median(drift[:, -1]):.3e}, max over systems and steps {drift.max():.3e}; "
              f"momentum drift max {pdrift.max():.3e}; {wall * 1e3:.1f} ms for run + energies of {B * (T + 1)} snapshots", flush=True)
    if a.out:
        Path(a.out).write_text(json.dumps(res) + "\n")


if __name__ == "__main__":
    main()
