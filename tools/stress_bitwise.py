#!/usr/bin/env python3
"""Randomised bit-exactness soak (compute-sanitizer's racecheck is closed on the pool): for --seconds, random shapes of
  (a) the interval-scheduled ensemble kernel (systems handed between CTA lanes through flags) against the same systems
      run in small one-system-per-CTA batches,
  (b) the fused force + leapfrog + peer-store step kernel on 2..4 virtual ranks (wait_seq and PEER_SYNC ordering)
      against the one-GPU step kernels,
  (c) slab-wise K2 steps against full-system steps,
each compared BIT FOR BIT.  A race in the mbarrier / flag / arrival-word protocols shows up as a mismatch sooner or later.

    python tools/stress_bitwise.py --seconds 60
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
sys.path.insert(0, str(ROOT / "tests"))
from hpc import _cuda, ics  # noqa: E402
from hpc.ensemble import simulate_ensemble  # noqa: E402
from hpc.sharded import ShardedSystem  # noqa: E402
from test_gpu_sharded import VirtualRanks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=2024)
    a = ap.parse_args()
    rng = np.random.RandomState(a.seed)
    eng = _cuda.get_engine()
    t_end = time.time() + a.seconds
    counts = {"ensemble": 0, "peer": 0, "slab": 0}
    while time.time() < t_end:
        # (a) ensemble
        n = int(rng.choice([16, 33, 64, 100, 200, 256, 333]))
        B = int(rng.randint(2 * eng.sm_count + 1, 5 * eng.sm_count))
        steps = int(rng.randint(0, 40))
        every = int(rng.randint(1, 6))
        dtype = str(rng.choice(["float64", "float32"]))
        x0 = rng.rand(B, n, 3) * 4 - 2
        v0 = rng.rand(B, n, 3) - 0.5
        m = rng.uniform(1e9, 1e11, n)
        kw = dict(dt=1e-3, softening=0.05, n_steps=steps, save_interval=every, dtype=dtype)
        big = simulate_ensemble(x0, v0, m, **kw)
        lo = int(rng.randint(0, B - 40))
        small = simulate_ensemble(x0[lo:lo + 40], v0[lo:lo + 40], m, **kw)
        for key in ("positions", "velocities", "accelerations", "final_positions", "final_velocities"):
            if not np.array_equal(big[key][lo:lo + 40], small[key]):
                print(f"MISMATCH ensemble B={B} N={n} steps={steps} every={every} {dtype} key={key}", flush=True)
                return 1
        counts["ensemble"] += 1
        # (b) peer step kernel on virtual ranks
        n = int(rng.randint(700, 30000))
        world = int(rng.randint(2, 5))
        dtype = rng.choice([np.float64, np.float32])
        mode = str(rng.choice(["wait", "sync"]))
        x, v, mm = ics.plummer_ic(n, seed=int(rng.randint(1, 10 ** 6)))
        vr = VirtualRanks(eng, x, v, mm, dtype, world, 1e-3, 0.01)
        k1, k2 = int(rng.randint(1, 6)), int(rng.randint(1, 4))
        vr.advance(k1, mode)
        vr.advance(k2, mode)
        one = ShardedSystem(x, v, mm, dt=1e-3, softening=0.01, dtype=dtype, device=eng.device)
        one.advance(k1)
        one.advance(k2)
        npad4 = eng.padded_bodies(n) * 4
        for r in range(world):
            i0, i1 = vr.bounds[r]
            if not (torch.equal(vr.cur[r][:npad4], one.cur[:npad4]) and torch.equal(vr.vel[r], one.vel[i0:i1])
                    and torch.equal(vr.acc[r], one.acc[i0:i1])):
                print(f"MISMATCH peer N={n} world={world} {np.dtype(dtype).name} mode={mode} rank={r}", flush=True)
                return 1
        counts["peer"] += 1
        # (c) slab-wise K2 steps
        cut = (int(rng.randint(1, n // 32)) * 32)
        pos_d = eng.to_device(x)
        m_d, f32 = eng._masses_dev(mm)
        tdt = torch.float64 if dtype == np.float64 else torch.float32
        res = []
        for slabs in ([(0, n)], [(0, cut), (cut, n - cut)]):
            cur = eng.pack(pos_d, m_d, f32, n, dtype)
            nxt = cur.clone()
            vel = eng.to_device(v, tdt)
            acc = eng.accel_slab(cur, n, 0, n, 0.01)
            ws = eng.workspace(n, n, dtype)
            for i0, n_i in slabs:
                eng.kick_drift_slab(cur, nxt, vel[i0:i0 + n_i], acc[i0:i0 + n_i], n, i0, n_i, 1e-3)
            cur, nxt = nxt, cur
            for _ in range(3):
                for i0, n_i in slabs:
                    eng.step_slab(cur, nxt, vel[i0:i0 + n_i], acc[i0:i0 + n_i], n, i0, n_i, 1e-3, 0.01,
                                  _cuda.NB_STEP_CONTINUE, None, None, None, ws)
                cur, nxt = nxt, cur
            res.append((cur.clone(), vel.clone(), acc.clone()))
        if not all(torch.equal(p, q) for p, q in zip(*res)):
            print(f"MISMATCH slab N={n} cut={cut} {np.dtype(dtype).name}", flush=True)
            return 1
        counts["slab"] += 1
    print(f"stress ok: {counts} random cases, all bit-identical, in {a.seconds:.0f} s (seed {a.seed})", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
