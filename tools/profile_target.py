#!/usr/bin/env python3
"""Small fixed workload for ncu captures: one launch of each hot kernel (development aid)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics  # noqa: E402

which = sys.argv[1:] or ["f32", "f64", "ens"]
eng = _cuda.get_engine()
if "f32" in which or "f64" in which:
    n = 65536
    x, v, m = ics.plummer_ic(n, seed=7)
    pos_d = eng.to_device(x)
    m_d, f32 = eng._masses_dev(m)
    for tag, dtype in (("f32", np.float32), ("f64", np.float64)):
        if tag in which:
            stream = eng.pack(pos_d, m_d, f32, n, dtype)
            ws = eng.workspace(n, n, dtype)
            for _ in range(2):
                eng.accel_slab(stream, n, 0, n, 0.01, ws)
            torch.cuda.synchronize()
if "ens" in which or "ens400" in which:
    B, steps = (300, 400) if "ens400" in which else (296, 50)
    x0, v0, m32 = ics.datagen_ensemble_ic(B, 200, seed=42)
    x, v = eng.to_device(x0), eng.to_device(v0)
    a = torch.zeros_like(x)
    m_d, f32 = eng._masses_dev(m32)
    ox = torch.empty((B, steps + 1, 200, 3), dtype=torch.float64, device=eng.device)
    ov, oa = torch.empty_like(ox), torch.empty_like(ox)
    for _ in range(2):
        eng.ensemble_device(x, v, a, m_d, f32, 0, B, 200, 1e-3, 1e-9, steps, 1, np.float64, True, True, ox, ov, oa,
                            steps + 1, 0)
    torch.cuda.synchronize()
if "win" in which:
    B, rows, N = 300, 401, 200
    g = torch.Generator(device=eng.device).manual_seed(1)
    pos = torch.randn((B, rows, N, 3), dtype=torch.float64, device=eng.device, generator=g)
    vel = torch.randn((B, rows, N, 3), dtype=torch.float64, device=eng.device, generator=g)
    for _ in range(2):
        eng.window_gather(pos, vel, rows, 10, 1)
    torch.cuda.synchronize()
if "clu" in which:
    # one system on a cluster of 8 CTAs (configs[0]) and ten (evaluate.py), 400 steps, snapshots every step
    for B in (1, 10):
        x0, v0, m32 = ics.datagen_ensemble_ic(B, 200, seed=42)
        x, v = eng.to_device(x0), eng.to_device(v0)
        a = torch.zeros_like(x)
        m_d, f32 = eng._masses_dev(m32)
        ox = torch.empty((B, 401, 200, 3), dtype=torch.float64, device=eng.device)
        ov, oa = torch.empty_like(ox), torch.empty_like(ox)
        for _ in range(2):
            eng.ensemble_device(x, v, a, m_d, f32, 0, B, 200, 1e-3, 1e-9, 400, 1, np.float64, True, True, ox, ov, oa, 401, 0)
        torch.cuda.synchronize()
if "energy" in which:
    n = 65536
    x, v, m = ics.plummer_ic(n, seed=7)
    pos_d, vel_d = eng.to_device(x), eng.to_device(v)
    m_d, f32 = eng._masses_dev(m)
    for _ in range(2):
        eng.energy_slab(pos_d, vel_d, m_d, f32, n, 0, n, 0.01)
    torch.cuda.synchronize()
if "snap" in which:
    # K4b on the stacks of a 300 x 401 x 200 ensemble
    B, S, N = 300, 401, 200
    g = torch.Generator(device=eng.device).manual_seed(2)
    pos = torch.randn((B, S, N, 3), dtype=torch.float64, device=eng.device, generator=g)
    vel = torch.randn((B, S, N, 3), dtype=torch.float64, device=eng.device, generator=g)
    m_d, f32 = eng._masses_dev(ics.shared_masses(N, 42))
    for _ in range(2):
        eng.snapshot_energies(pos, vel, m_d, f32, 0, 1e-9)
    torch.cuda.synchronize()
if "peer" in which:
    # the fused force + leapfrog + peer-store step kernel, one rank that is its own peer, N = 65,536 float32
    n = 65536
    x, v, m = ics.plummer_ic(n, seed=7)
    pos_d = eng.to_device(x)
    m_d, f32 = eng._masses_dev(m)
    cur = eng.pack(pos_d, m_d, f32, n, np.float32)
    nxt = cur.clone()
    vel = eng.to_device(v, torch.float32)
    ws = eng.workspace(n, n, np.float32)
    acc = eng.accel_slab(cur, n, 0, n, 1e-3, ws)
    flags = torch.zeros(16, dtype=torch.int32, device=eng.device)
    eng.kick_drift_slab(cur, nxt, vel, acc, n, 0, n, 1e-3)
    cur, nxt = nxt, cur
    for k in range(1, 4):
        eng.step_peer_slab(cur, [nxt.data_ptr()], [flags.data_ptr()], 0, 0, k, vel, acc, n, 0, n, 1e-3, 1e-3,
                           _cuda.NB_STEP_CONTINUE | _cuda.NB_STEP_PEER_SYNC, None, None, None, ws)
        cur, nxt = nxt, cur
    torch.cuda.synchronize()
    eng.step_status(ws, n)
if "group" in which:
    # K2s: whole mid-size systems, one CTA per group of bodies (N = 4,096 and 8,192, both precisions, five steps each)
    from hpc.sharded import ShardedSystem
    for n in (4096, 8192):
        x, v, m = ics.plummer_ic(n, seed=7)
        for dtype in (np.float32, np.float64):
            s = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=0)
            for _ in range(2):
                s.advance(5)
            torch.cuda.synchronize()
if "persist" in which:
    # K2p (opt-in): ten leapfrog steps of N = 16,384 in one cooperative launch, both precisions
    import os
    from hpc.sharded import ShardedSystem
    os.environ["NB_PERSIST"] = "1"
    n = 16384
    x, v, m = ics.plummer_ic(n, seed=7)
    for dtype in (np.float32, np.float64):
        s = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=0)
        for _ in range(2):
            s.advance(10)
        torch.cuda.synchronize()
        eng.step_status(s.ws, n)
    os.environ.pop("NB_PERSIST", None)
print("ok")
