"""One large system split by i-slab over the GPUs of a box (one process per GPU).

Rank r owns the bodies of slab r: it keeps their velocities and accelerations, evaluates their
accelerations against ALL bodies (K1/K2 on rows [i0, i0+n_i)) and writes their new positions into
its slab of the next position stream; one all-gather of that stream per leapfrog step (NCCL over
NVLink/NVSwitch, in place: every rank's send buffer is its own slab of the receive buffer) gives
every rank the full new positions.  Nothing else is communicated.

Because the j-segmentation of the force kernels depends on N only (nb_segment_plan), a body's
acceleration is the same bits on any rank count: sharded runs are bit-identical to one-GPU runs.

The reference has no multi-device path; this replaces running NBodySimulator.step
(reference src/hpc/nbody.py:202-218) on a single host for systems too large for it.
"""
from __future__ import annotations

import numpy as np

import contextlib
import functools

from . import _cuda

CHUNK = 32


def _on_own_device(method):
    """Run a ShardedSystem method with the engine's device current: the C side launches on the current CUDA
    device while the stream comes from the engine's device, so the two must agree whatever the caller's
    current device is."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with self._device_guard():
            return method(self, *args, **kwargs)
    return wrapper  # NB_CHUNK_BODIES: slab boundaries are multiples of it


def slab_bounds(n: int, world: int):
    """Equal padded slabs of `slab` bodies (multiple of CHUNK); rank r owns [r*slab, min((r+1)*slab, n))."""
    slab = -(-n // world)
    slab = -(-slab // CHUNK) * CHUNK
    return slab, [(min(r * slab, n), min((r + 1) * slab, n)) for r in range(world)]


def _all_gather_inplace(dist, full, slab_elems: int, rank: int, group):
    """All-gather with each rank's contribution already in place at full[rank*slab_elems : ...]."""
    mine = full[rank * slab_elems:(rank + 1) * slab_elems]
    try:
        dist.all_gather_into_tensor(full, mine, group=group)
    except (RuntimeError, NotImplementedError):
        world = dist.get_world_size(group)
        parts = [full[r * slab_elems:(r + 1) * slab_elems] for r in range(world)]
        dist.all_gather(parts, mine.clone(), group=group)


class ShardedSystem:
    """Device-resident state of one system, advanced slab-wise.  world == 1 is the plain one-GPU case."""

    def __init__(self, positions, velocities, masses, dt: float, softening: float, dtype=np.float64, device=None,
                 world: int = 1, rank: int = 0, group=None, engine=None, accelerations=None, exchange: str = "auto",
                 check_peers: bool = True):
        import torch
        self.torch = torch
        self.eng = engine if engine is not None else _cuda.get_engine(device)
        self.check_peers = bool(check_peers)
        with self._device_guard():
            self._init(positions, velocities, masses, dt, softening, dtype, world, rank, group, engine,
                       accelerations, exchange)

    def _device_guard(self):
        dev = getattr(self.eng, "device", None)
        if dev is not None and getattr(dev, "type", "cpu") == "cuda":
            return self.torch.cuda.device(dev)
        return contextlib.nullcontext()

    def _init(self, positions, velocities, masses, dt, softening, dtype, world, rank, group, engine, accelerations,
              exchange):
        torch = self.torch
        self.n = int(np.asarray(positions).shape[0])
        self.dt, self.softening = float(dt), float(softening)
        self.dtype = np.dtype(dtype)
        self.world, self.rank, self.group = int(world), int(rank), group
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
        self.slab, bounds = slab_bounds(self.n, self.world)
        self.i0, i1 = bounds[self.rank]
        self.n_i = i1 - self.i0
        eng = self.eng
        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        total = max(self.slab * self.world, eng.padded_bodies(self.n))
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        self.masses_host = np.asarray(masses)
        pos_d = eng.to_device(pos)
        m_d, f32 = eng._masses_dev(self.masses_host)
        self._m_d, self._m_f32 = m_d, f32
        # exchange: "nccl" = one in-place all-gather per step; "peer" = the drift kernel stores its slab into
        # every rank's stream over NVLink and ranks order themselves with arrival words (no collective call);
        # "auto" = peer when torch symmetric memory can map the buffers, else nccl.
        self.exchange = "nccl"
        self._peer = None
        every_rank_has_bodies = all(hi > lo for lo, hi in bounds)
        if self.world > 1 and exchange in ("auto", "peer") and engine is None and every_rank_has_bodies:
            try:
                self._peer = self._setup_peer(total * 4, tdt)
                self.exchange = "peer"
            except Exception as e:  # symmetric memory unavailable on this build / topology
                if exchange == "peer":
                    raise
                self._peer_error = repr(e)
        if self._peer is not None:
            self.cur, self.nxt = self._peer["bufs"]
            self.cur.zero_()
        else:
            self.cur = torch.zeros(total * 4, dtype=tdt, device=eng.device)
        eng.pack(pos_d, m_d, f32, self.n, self.dtype, out=self.cur)
        if self._peer is not None:
            self.nxt.copy_(self.cur)
            self._seq = 0
            torch.cuda.synchronize(eng.device)
            self.dist.barrier(group=self.group)
        else:
            self.nxt = self.cur.clone()
        sl = slice(self.i0, self.i0 + self.n_i)
        vel = np.ascontiguousarray(velocities, dtype=np.float64)[sl]
        self.vel = eng.to_device(np.ascontiguousarray(vel), tdt) if self.n_i else torch.zeros((0, 3), dtype=tdt, device=eng.device)
        self.ws = eng.workspace(self.n, max(self.n_i, 1), self.dtype)
        if accelerations is None:
            self.acc = (eng.accel_slab(self.cur, self.n, self.i0, self.n_i, self.softening, self.ws)
                        if self.n_i else torch.zeros((0, 3), dtype=tdt, device=eng.device))
        else:
            acc = np.ascontiguousarray(accelerations, dtype=np.float64)[sl]
            self.acc = eng.to_device(np.ascontiguousarray(acc), tdt) if self.n_i else torch.zeros((0, 3), dtype=tdt, device=eng.device)
        self.steps_done = 0

    def _setup_peer(self, n_elems: int, tdt):
        """Both stream buffers in symmetric memory: peer pointers for the stores, signal pads for the arrival words."""
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else self.dist.group.WORLD
        bufs, handles = [], []
        for _ in range(2):
            t = symm.empty(n_elems, dtype=tdt, device=self.eng.device)
            handles.append(symm.rendezvous(t, group))
            bufs.append(t)
        ptrs = [[int(p) for p in h.buffer_ptrs] for h in handles]
        flags = [int(p) for p in handles[0].signal_pad_ptrs]
        if handles[0].signal_pad_size < 64 * 4:
            raise RuntimeError("signal pad too small")
        return {"bufs": bufs, "handles": handles, "ptrs": {bufs[0].data_ptr(): ptrs[0], bufs[1].data_ptr(): ptrs[1]},
                "flags": flags}

    # -- communication --------------------------------------------------------------------------
    def _exchange(self, stream):
        if self.world > 1:
            _all_gather_inplace(self.dist, stream, self.slab * 4, self.rank, self.group)

    # -- stepping -------------------------------------------------------------------------------
    @_on_own_device
    def advance(self, n_steps: int, snap_pos=None, snap_vel=None, snap_acc=None, save_interval: int = 1):
        """n_steps kick-drift-kick steps.  snap_*: optional device tensors (n_snap, N, 3) float64; each
        rank fills the rows of its own slab for the saved steps (row s = state after s*save_interval)."""
        if n_steps <= 0:
            return
        eng = self.eng
        if self.world == 1 and hasattr(eng, "run_device"):
            # one GPU: the whole loop is enqueued by one C call (no per-step Python between launches)
            in_a = eng.run_device(self.cur, self.nxt, self.vel, self.acc, self.n, self.dt, self.softening, n_steps,
                                  save_interval, snap_pos, snap_vel, snap_acc, self.ws)
            if not in_a:
                self.cur, self.nxt = self.nxt, self.cur
            self.steps_done += n_steps
            return
        if self.n_i:
            eng.kick_drift_slab(self.cur, self.nxt, self.vel, self.acc, self.n, self.i0, self.n_i, self.dt)
        self._exchange(self.nxt)
        self.cur, self.nxt = self.nxt, self.cur
        snap = 1
        for k in range(1, n_steps + 1):
            flags = _cuda.NB_STEP_CONTINUE if k < n_steps else 0
            save = snap_pos is not None and (k % save_interval) == 0
            if save:
                flags |= _cuda.NB_STEP_SNAPSHOT
            sp = snap_pos[snap] if save else None
            sv = snap_vel[snap] if save else None
            sa = snap_acc[snap] if save else None
            if self._peer is not None:
                # fused: force + leapfrog + store of the slab into every rank's next stream + arrival word.
                # The first step of an advance() reads a stream that NCCL completed (no wait needed).
                # The kernel's last block waits for every rank's arrival word (NB_STEP_PEER_SYNC), so the next
                # force pass needs no wait of its own.
                self._seq += 1
                eng.step_peer_slab(self.cur, self._peer["ptrs"][self.nxt.data_ptr()], self._peer["flags"], self.rank,
                                   0, self._seq, self.vel, self.acc, self.n, self.i0, self.n_i, self.dt,
                                   self.softening, flags | _cuda.NB_STEP_PEER_SYNC, sp, sv, sa, self.ws)
                if k < n_steps:
                    self.cur, self.nxt = self.nxt, self.cur
            else:
                if self.n_i:
                    eng.step_slab(self.cur, self.nxt, self.vel, self.acc, self.n, self.i0, self.n_i, self.dt,
                                  self.softening, flags, sp, sv, sa, self.ws)
                if k < n_steps:
                    self._exchange(self.nxt)
                    self.cur, self.nxt = self.nxt, self.cur
            if save:
                snap += 1
        self.steps_done += n_steps
        if self._peer is not None and self.check_peers:
            # a rank that never arrived is an error, not a step from stale positions (synchronises this stream)
            eng.step_status(self.ws, self.n)

    # -- results --------------------------------------------------------------------------------
    @_on_own_device
    def positions(self) -> np.ndarray:
        """Full (N,3) float64 positions (every rank holds them)."""
        return self.eng.unpack(self.cur, self.n).cpu().numpy()

    def _gather_rows(self, local):
        torch = self.torch
        t = local.to(torch.float64)
        if self.world == 1:
            return t.cpu().numpy()
        full = torch.zeros((self.slab * self.world, 3), dtype=torch.float64, device=t.device)
        full[self.rank * self.slab:self.rank * self.slab + self.n_i] = t
        _all_gather_inplace(self.dist, full.view(-1), self.slab * 3, self.rank, self.group)
        return full[:self.n].cpu().numpy()

    @_on_own_device
    def velocities(self) -> np.ndarray:
        return self._gather_rows(self.vel)

    @_on_own_device
    def accelerations(self) -> np.ndarray:
        return self._gather_rows(self.acc)

    @_on_own_device
    def energy(self):
        """(K, U, K+U): every rank sums its slab (K4), one all-reduce of two doubles."""
        torch = self.torch
        eng = self.eng
        pos_d = eng.unpack(self.cur, self.n)
        vel_full = torch.zeros((self.n, 3), dtype=torch.float64, device=pos_d.device)
        if self.n_i:
            vel_full[self.i0:self.i0 + self.n_i] = self.vel.to(torch.float64)
            ku = eng.energy_slab(pos_d, vel_full, self._m_d, self._m_f32, self.n, self.i0, self.n_i, self.softening)
        else:
            ku = torch.zeros(2, dtype=torch.float64, device=pos_d.device)
        if self.world > 1:
            self.dist.all_reduce(ku, group=self.group)
        k, u = (float(t) for t in ku.cpu())
        return k, u, k + u
