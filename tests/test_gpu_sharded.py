"""The sharded i-slab mode on real kernels: nb_step_peer_* (force + leapfrog + peer stores + arrival words in ONE
kernel, csrc/nb_force.cu kEpiStepPeer) against the plain one-GPU step kernels, bit for bit.

 * one rank that is its own peer (n_ranks = 1): every code path of the fused epilogue on a 1-GPU box;
 * two VIRTUAL ranks on one GPU -- separate stream buffers, flag arrays, velocities and workspaces, exactly what two
   processes hold -- ordered either by wait_seq on one CUDA stream or by NB_STEP_PEER_SYNC on a stream per rank
   (two kernels of different ranks in flight on the same GPU, really waiting for each other's arrival words);
 * a lost peer is an ERROR the host sees (nb_step_status), and the launches after it do not compute;
 * two real ranks (spawned processes, NCCL + symmetric memory) when the box has >= 2 GPUs.

No counterpart in the reference (single host, src/hpc/nbody.py:202-218); the contract is SURVEY 8(e)2: the same bits
for every rank count."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent


def _tdt(torch, dtype):
    return torch.float64 if np.dtype(dtype) == np.float64 else torch.float32


class VirtualRanks:
    """`world` ranks of a sharded system held by ONE process on ONE GPU, driven through the engine's slab calls."""

    def __init__(self, eng, x, v, m, dtype, world, dt, eps):
        import torch
        from hpc.sharded import slab_bounds
        self.torch, self.eng, self.world, self.dt, self.eps = torch, eng, world, dt, eps
        self.n = n = len(x)
        self.slab, self.bounds = slab_bounds(n, world)
        total = max(self.slab * world, eng.padded_bodies(n))
        tdt = _tdt(torch, dtype)
        pos_d = eng.to_device(x)
        m_d, f32 = eng._masses_dev(m)
        self.cur, self.nxt, self.vel, self.acc, self.ws, self.flags = [], [], [], [], [], []
        for r in range(world):
            i0, i1 = self.bounds[r]
            cur = torch.zeros(total * 4, dtype=tdt, device=eng.device)
            eng.pack(pos_d, m_d, f32, n, dtype, out=cur)
            self.cur.append(cur)
            self.nxt.append(cur.clone())
            self.vel.append(eng.to_device(np.ascontiguousarray(v[i0:i1]), tdt))
            ws = eng.workspace(n, i1 - i0, dtype)
            self.ws.append(ws)
            self.acc.append(eng.accel_slab(cur, n, i0, i1 - i0, eps, ws))
            self.flags.append(torch.zeros(16, dtype=torch.int32, device=eng.device))
        self.seq = 0

    def _open(self):
        """First half step of an advance: kick + drift of every slab, slabs handed round by plain copies."""
        for r in range(self.world):
            i0, i1 = self.bounds[r]
            self.eng.kick_drift_slab(self.cur[r], self.nxt[r], self.vel[r], self.acc[r], self.n, i0, i1 - i0, self.dt)
        for r in range(self.world):
            lo, hi = r * self.slab * 4, (r + 1) * self.slab * 4
            for q in range(self.world):
                if q != r:
                    self.nxt[q][lo:hi].copy_(self.nxt[r][lo:hi])
        self.cur, self.nxt = self.nxt, self.cur

    def advance(self, n_steps, mode, snaps=None):
        from hpc import _cuda
        torch, eng = self.torch, self.eng
        self._open()
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in range(self.world)] if mode == "sync" else None
        for k in range(1, n_steps + 1):
            flags = _cuda.NB_STEP_CONTINUE if k < n_steps else 0
            if snaps is not None:
                flags |= _cuda.NB_STEP_SNAPSHOT
            first = k == 1
            self.seq += 1
            nxt_ptrs = [t.data_ptr() for t in self.nxt]
            flag_ptrs = [t.data_ptr() for t in self.flags]
            for r in range(self.world):
                i0, i1 = self.bounds[r]
                sp, sv, sa = (s[k] for s in snaps) if snaps is not None else (None, None, None)
                args = (self.cur[r], nxt_ptrs, flag_ptrs, r)
                tail = (self.vel[r], self.acc[r], self.n, i0, i1 - i0, self.dt, self.eps)
                if mode == "sync":
                    with torch.cuda.stream(streams[r]):
                        eng.step_peer_slab(*args, 0, self.seq, *tail, flags | _cuda.NB_STEP_PEER_SYNC, sp, sv, sa,
                                           self.ws[r])
                else:
                    eng.step_peer_slab(*args, 0 if first else self.seq - 1, self.seq, *tail, flags, sp, sv, sa,
                                       self.ws[r])
            if mode == "sync":
                # On ONE GPU the next step's kernels (programmatic dependent launch: resident early, waiting for
                # their predecessor) must not take the SM slots the other rank's kernel of THIS step still needs.
                torch.cuda.synchronize()
            if k < n_steps:
                self.cur, self.nxt = self.nxt, self.cur
        torch.cuda.synchronize()
        for r in range(self.world):
            eng.step_status(self.ws[r], self.n)


def _single(eng, x, v, m, dtype, dt, eps, chunks, snaps=None):
    from hpc.sharded import ShardedSystem
    one = ShardedSystem(x, v, m, dt=dt, softening=eps, dtype=dtype, engine=None, device=eng.device)
    for c in chunks:
        if snaps is not None:
            one.advance(c, *snaps)
        else:
            one.advance(c)
    return one


@pytest.mark.parametrize("n,dtype", [(5000, np.float64), (4099, np.float32), (40000, np.float32)])
def test_one_rank_is_its_own_peer_bitwise(engine, n, dtype):
    """n_ranks = 1: the fused epilogue stores the slab into its own next stream and exchanges arrival words with
    itself -- wait_seq path and NB_STEP_PEER_SYNC path -- and must give the bits of nb_step_*."""
    import torch
    from hpc import ics
    x, v, m = ics.plummer_ic(n, seed=7)
    for mode in ("wait", "sync"):
        vr = VirtualRanks(engine, x, v, m, dtype, 1, 1e-3, 0.01)
        snaps = tuple(torch.zeros((5, n, 3), dtype=torch.float64, device=engine.device) for _ in range(3))
        vr.advance(4, mode, snaps)
        vr.advance(2, mode)
        ref_snaps = tuple(torch.zeros_like(s) for s in snaps)
        one = _single(engine, x, v, m, dtype, 1e-3, 0.01, [], None)
        one.advance(4, *ref_snaps)
        one.advance(2)
        npad4 = engine.padded_bodies(n) * 4
        assert torch.equal(vr.cur[0][:npad4], one.cur[:npad4]), mode
        assert torch.equal(vr.vel[0], one.vel) and torch.equal(vr.acc[0], one.acc), mode
        for a, b in zip(snaps, ref_snaps):
            assert torch.equal(a[1:], b[1:]), mode
        assert int(vr.flags[0][0]) == vr.seq                      # the last arrival word this rank published to itself


@pytest.mark.parametrize("n,dtype,world", [(6000, np.float64, 2), (20000, np.float32, 2), (9001, np.float32, 3)])
@pytest.mark.parametrize("mode", ["wait", "sync"])
def test_virtual_ranks_on_one_gpu_bitwise(engine, n, dtype, world, mode):
    import torch
    from hpc import ics
    x, v, m = ics.plummer_ic(n, seed=11)
    vr = VirtualRanks(engine, x, v, m, dtype, world, 1e-3, 0.01)
    vr.advance(5, mode)
    vr.advance(3, mode)
    one = _single(engine, x, v, m, dtype, 1e-3, 0.01, [5, 3])
    for r in range(world):
        i0, i1 = vr.bounds[r]
        npad4 = engine.padded_bodies(n) * 4
        assert torch.equal(vr.cur[r][:npad4], one.cur[:npad4]), (mode, r)     # every rank holds the full new stream
        assert torch.equal(vr.vel[r], one.vel[i0:i1]) and torch.equal(vr.acc[r], one.acc[i0:i1]), (mode, r)
        assert [int(f) for f in vr.flags[r][:world]] == [vr.seq] * world


def test_lost_peer_is_an_error_not_a_stale_step(engine):
    """Rank 1 of 2 never runs.  NB_STEP_PEER_SYNC: rank 0's kernel times out waiting for rank 1's arrival word,
    records it, nb_step_status raises; the next launch on that workspace leaves without computing; a wait_seq that
    is never reached fails the same way before the force pass.  (NB_PEER_TIMEOUT_MS is set in conftest.py.)"""
    import torch
    from hpc import _cuda, ics
    n = 4096
    x, v, m = ics.plummer_ic(n, seed=3)
    for phase in ("after", "before"):
        vr = VirtualRanks(engine, x, v, m, np.float64, 2, 1e-3, 0.01)
        vr._open()
        i0, i1 = vr.bounds[0]
        nxt_ptrs = [t.data_ptr() for t in vr.nxt]
        flag_ptrs = [t.data_ptr() for t in vr.flags]
        vel_before = vr.vel[0].clone()
        if phase == "after":
            engine.step_peer_slab(vr.cur[0], nxt_ptrs, flag_ptrs, 0, 0, 1, vr.vel[0], vr.acc[0], n, i0, i1 - i0, 1e-3,
                                  0.01, _cuda.NB_STEP_CONTINUE | _cuda.NB_STEP_PEER_SYNC, None, None, None, vr.ws[0])
            with pytest.raises(RuntimeError, match="rank 1 never arrived.*did not finish the step"):
                engine.step_status(vr.ws[0], n)
            assert not torch.equal(vr.vel[0], vel_before)       # the step itself ran; its hand-over failed
            vel_before = vr.vel[0].clone()
        else:
            vr.flags[0][0] = 7                                  # this rank's own word is there; rank 1's never comes
            engine.step_peer_slab(vr.cur[0], nxt_ptrs, flag_ptrs, 0, 7, 8, vr.vel[0], vr.acc[0], n, i0, i1 - i0, 1e-3,
                                  0.01, _cuda.NB_STEP_CONTINUE, None, None, None, vr.ws[0])
            with pytest.raises(RuntimeError, match="rank 1 never arrived.*not published in time"):
                engine.step_status(vr.ws[0], n)
            assert torch.equal(vr.vel[0], vel_before)           # nothing was computed from stale positions
        # sticky: later launches on this workspace leave at once, without computing and without another timeout
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        engine.step_peer_slab(vr.cur[0], nxt_ptrs, flag_ptrs, 0, 0, 9, vr.vel[0], vr.acc[0], n, i0, i1 - i0, 1e-3, 0.01,
                              _cuda.NB_STEP_CONTINUE, None, None, None, vr.ws[0])
        t1.record()
        with pytest.raises(RuntimeError, match="never arrived"):
            engine.step_status(vr.ws[0], n)
        assert torch.equal(vr.vel[0], vel_before) and t0.elapsed_time(t1) < 100.0
        # a zeroed workspace is usable again
        vr.ws[0][0].zero_()
        engine.step_slab(vr.cur[0], vr.nxt[0], vr.vel[0], vr.acc[0], n, i0, i1 - i0, 1e-3, 0.01, 0, None, None, None,
                         vr.ws[0])
        engine.step_status(vr.ws[0], n)


def test_sharded_system_on_a_non_current_device_guard(engine):
    """ShardedSystem makes its engine's device current around every call: with one GPU the guard is a
    no-op, but the code path is the one multi-device callers take."""
    from hpc import ics
    from hpc.sharded import ShardedSystem
    x, v, m = ics.plummer_ic(2048, seed=5)
    s = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=np.float32, device=engine.device)
    s.advance(3)
    assert np.isfinite(s.positions()).all() and np.isfinite(s.energy()[2])


# ---- two real ranks -------------------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, q):
    for p in (str(ROOT), str(ROOT / "nbody-gnn-hpc_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from hpc import ics
    from hpc.sharded import ShardedSystem
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    out = {}
    try:
        for n, dtype in ((10000, np.float64), (40000, np.float32)):
            x, v, m = ics.plummer_ic(n, seed=7)
            for exchange in ("nccl", "peer"):
                sh = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=rank, world=world, rank=rank,
                                   exchange=exchange)
                sh.advance(5)
                sh.advance(2)
                pos, vel, acc, e = sh.positions(), sh.velocities(), sh.accelerations(), sh.energy()
                if rank == 0:
                    one = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=rank)
                    one.advance(5)
                    one.advance(2)
                    same = (np.array_equal(pos, one.positions()) and np.array_equal(vel, one.velocities())
                            and np.array_equal(acc, one.accelerations()))
                    out[(n, np.dtype(dtype).name, exchange)] = (sh.exchange, same, e[2], one.energy()[2])
                dist.barrier()
                del sh
    finally:
        dist.destroy_process_group()
    if rank == 0:
        q.put(out)


def test_two_real_ranks_bitwise_equal_to_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2); the virtual-rank tests cover the kernel on one GPU")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(out) == 4
    for key, (used, same, e, e1) in out.items():
        assert used == key[2], (key, used)            # exchange="peer" really ran the fused kernel
        assert same, key
        assert abs(e - e1) <= 1e-12 * abs(e1), key
