"""world_size-2 test of the i-slab sharded mode on CPU (gloo): slab partition, in-place all-gather of
the position stream, snapshot rows, gathered state and the energy all-reduce.  The arithmetic is the
oracle's (tests/fake_engine.py); the CUDA kernels' slab invariance is checked by the GPU tests."""
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, n_steps, save_interval, q):
    for p in (str(ROOT), str(ROOT / "nbody-gnn-hpc_b200"), str(ROOT / "tests")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    from fake_engine import FakeEngine
    from hpc import ics
    from hpc.sharded import ShardedSystem
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        x, v, m = ics.plummer_ic(n, seed=7)
        sysm = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=np.float64, world=world, rank=rank,
                             engine=FakeEngine())
        n_snap = 1 + n_steps // save_interval
        sp, sv, sa = (torch.zeros((n_snap, n, 3), dtype=torch.float64) for _ in range(3))
        sysm.advance(n_steps, sp, sv, sa, save_interval)
        # every rank filled only its own rows of the snapshots: sum them
        for t in (sp, sv, sa):
            dist.all_reduce(t)
        pos, vel, acc = sysm.positions(), sysm.velocities(), sysm.accelerations()
        e = sysm.energy()
        if rank == 0:
            q.put({"pos": pos, "vel": vel, "acc": acc, "snap_pos": sp.numpy(), "snap_vel": sv.numpy(), "energy": e,
                   "slab": sysm.slab, "n_i": sysm.n_i})
        else:
            q.put({"n_i": sysm.n_i, "i0": sysm.i0})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,n_steps,save_interval", [(200, 6, 2), (97, 5, 1)])
def test_two_rank_slab_run_matches_single(oracle_mod, n, n_steps, save_interval):
    from hpc import ics
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, n_steps, save_interval, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    main = next(r for r in results if "pos" in r)
    other = next(r for r in results if "pos" not in r)
    assert main["slab"] % 32 == 0 and main["n_i"] + other["n_i"] == n and other["i0"] == main["slab"]
    x, v, m = ics.plummer_ic(n, seed=7)
    chk = oracle_mod.run(x, v, oracle_mod.accel_direct(x, m, 0.01), m, 1e-3, 0.01, n_steps, save_interval)
    assert np.abs(main["pos"] - chk["final_positions"]).max() < 1e-13
    assert np.abs(main["vel"] - chk["final_velocities"]).max() < 1e-13
    assert np.abs(main["acc"] - chk["final_accelerations"]).max() < 1e-12 * np.abs(chk["final_accelerations"]).max()
    assert np.abs(main["snap_pos"][1:] - chk["positions"][1:]).max() < 1e-13
    assert np.abs(main["snap_vel"][1:] - chk["velocities"][1:]).max() < 1e-13
    e = oracle_mod.total_energy(chk["final_positions"], chk["final_velocities"], m, 0.01)
    assert np.allclose(main["energy"], e, rtol=1e-11)


def test_slab_bounds():
    from hpc.sharded import slab_bounds
    for n, w in ((262144, 8), (1048576, 8), (200, 2), (97, 2), (33, 4), (1000003, 8)):
        slab, b = slab_bounds(n, w)
        assert slab % 32 == 0 and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(hi - lo <= slab for lo, hi in b)
