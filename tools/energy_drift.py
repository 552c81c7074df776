#!/usr/bin/env python3
"""Relative energy-drift curves of the float32 and float64 engines (and, at small N, the CPU oracle) on
Plummer initial conditions -- the deliverable BASELINE.json north_star calls "energy-drift curves overlaid".

    python tools/energy_drift.py N n_steps save_interval [--oracle] [--out file.json]
    torchrun --nproc-per-node P tools/energy_drift.py ...        # sharded over P GPUs

Energies are evaluated by K4 (float64 on the device) from the synchronised snapshots (x_k, v_k)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import ics  # noqa: E402
from hpc.sharded import ShardedSystem  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n", type=int)
    ap.add_argument("n_steps", type=int)
    ap.add_argument("save_interval", type=int)
    ap.add_argument("--oracle", action="store_true", help="also run the CPU oracle (small N only)")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.n
    eps = 0.01 if n <= 16384 else 1e-3
    x, v, m = ics.plummer_ic(n, seed=7)
    res = {"N": n, "n_steps": a.n_steps, "save_interval": a.save_interval, "dt": 1e-3, "softening": eps,
           "ics": "Plummer (G*M = 1), seed 7", "gpus": world, "steps": list(range(0, a.n_steps + 1, a.save_interval)),
           "curves": {}}
    for tag in a.dtypes.split(","):
        dtype = np.float64 if tag == "f64" else np.float32
        sysm = ShardedSystem(x, v, m, dt=1e-3, softening=eps, dtype=dtype, device=local, world=world, rank=rank)
        e = [sysm.energy()]
        t0 = time.perf_counter()
        for _ in range(a.n_steps // a.save_interval):
            sysm.advance(a.save_interval)
            e.append(sysm.energy())
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        e = np.array(e)
        res["curves"][tag] = {"K": e[:, 0].tolist(), "U": e[:, 1].tolist(), "E": e[:, 2].tolist(),
                              "rel_drift": ((e[:, 2] - e[0, 2]) / abs(e[0, 2])).tolist(),
                              "wall_s_incl_energy": round(dt_s, 3), "exchange": sysm.exchange}
        if rank == 0:
            d = res["curves"][tag]["rel_drift"]
            print(f"N={n} {tag}: E0={e[0, 2]:.9e} max|dE/E0|={np.abs(d).max():.3e} final={d[-1]:+.3e} ({dt_s:.1f} s)", flush=True)
    if a.oracle and rank == 0:
        sys.path.insert(0, str(ROOT))
        import oracle
        out = oracle.run(x, v, oracle.accel_direct(x, m, eps), m, 1e-3, eps, a.n_steps, a.save_interval)
        e = np.array([oracle.total_energy(out["positions"][k], out["velocities"][k], m, eps, parallel=True)
                      for k in range(out["positions"].shape[0])])
        res["curves"]["cpu_oracle_f64"] = {"E": e[:, 2].tolist(), "rel_drift": ((e[:, 2] - e[0, 2]) / abs(e[0, 2])).tolist()}
        d = res["curves"]["cpu_oracle_f64"]["rel_drift"]
        print(f"N={n} cpu oracle: E0={e[0, 2]:.9e} max|dE/E0|={np.abs(d).max():.3e} final={d[-1]:+.3e}", flush=True)
        if "f64" in res["curves"]:
            g = np.array(res["curves"]["f64"]["E"])
            print(f"   GPU f64 vs oracle energy curve: max rel diff {np.abs(g - e[:, 2]).max() / abs(e[0, 2]):.3e}", flush=True)
    if rank == 0 and a.out:
        Path(a.out).write_text(json.dumps(res) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
