"""A stand-in for hpc._cuda.Engine that runs on CPU torch tensors and delegates the arithmetic to the
oracle.  TEST INFRASTRUCTURE: lets the host-side logic (state dicts, bookkeeping, slab partition,
all-gather plumbing) be exercised with `-m "not gpu"` in a container without a GPU.  It is never
importable from the product package."""
import numpy as np
import torch

import oracle

NB_STEP_CONTINUE, NB_STEP_SNAPSHOT = 1, 2
G = 6.67430e-11


class FakeResident:
    """Stand-in for hpc._cuda.ResidentSystem: the state lives here between calls, the oracle advances it."""

    def __init__(self, eng, positions, velocities, accelerations, masses, dt, softening):
        self.eng, self.dt, self.softening = eng, dt, softening
        self.x = np.array(positions, dtype=np.float64)
        self.v = np.array(velocities, dtype=np.float64)
        self.a = np.array(accelerations, dtype=np.float64)
        self.m = np.array(masses)

    def advance(self, n_steps, save_interval=1, snapshots=False):
        self.eng.calls.append(("run", n_steps, save_interval, snapshots))
        out = oracle.run(self.x, self.v, self.a, self.m, self.dt, self.softening, n_steps, save_interval)
        self.x, self.v, self.a = out["final_positions"], out["final_velocities"], out["final_accelerations"]
        if snapshots:
            return {k: out[k] for k in ("positions", "velocities", "accelerations")}
        return None

    def advance_async(self, n_steps, save_interval=1, snapshots=False):
        out = self.advance(n_steps, save_interval, snapshots)
        return (lambda: out) if snapshots else None

    def step(self):
        self.advance(1)

    def download(self):
        self.eng.calls.append(("download",))
        return self.x.copy(), self.v.copy(), self.a.copy()

    def energy(self):
        return oracle.total_energy(self.x, self.v, self.m, self.softening)


class FakeEngine:
    device = torch.device("cpu")
    sm_count = 148

    def __init__(self):
        self.launches = 0
        self.calls = []

    # ---- host-level (what NBodySimulator / simulate_ensemble call) --------------------------------
    def accelerations(self, positions, masses, softening, dtype=np.float64):
        self.calls.append(("accelerations", np.dtype(dtype).name))
        return oracle.accel_direct(positions, masses, softening)

    def energy(self, positions, velocities, masses, softening):
        return oracle.total_energy(positions, velocities, masses, softening)

    def run(self, positions, velocities, accelerations, masses, dt, softening, n_steps, save_interval=1,
            dtype=np.float64, snapshots=True):
        self.calls.append(("run", n_steps, save_interval, snapshots))
        out = oracle.run(positions, velocities, accelerations, masses, dt, softening, n_steps, save_interval)
        res = {k: out[k] for k in ("final_positions", "final_velocities", "final_accelerations")}
        if snapshots:
            res.update(positions=out["positions"], velocities=out["velocities"], accelerations=out["accelerations"])
        return res

    def resident(self, positions, velocities, accelerations, masses, dt, softening, dtype=np.float64):
        self.calls.append(("resident", np.asarray(positions).shape[0]))
        return FakeResident(self, positions, velocities, accelerations, masses, dt, softening)

    def ensemble(self, x0, v0, masses, dt, softening, n_steps, save_interval=1, dtype=np.float64, a0=None,
                 snapshots=True):
        x0, v0 = np.asarray(x0, dtype=np.float64), np.asarray(v0, dtype=np.float64)
        m = np.asarray(masses)
        outs = []
        for b in range(x0.shape[0]):
            mb = m if m.ndim == 1 else m[b]
            a = oracle.accel_direct(x0[b], mb, softening) if a0 is None else a0[b]
            outs.append(oracle.run(x0[b], v0[b], a, mb, dt, softening, n_steps, save_interval))
        res = {k: np.stack([o[k] for o in outs]) for k in ("final_positions", "final_velocities", "final_accelerations")}
        if snapshots:
            res.update({k: np.stack([o[k] for o in outs]) for k in ("positions", "velocities", "accelerations")})
        return res

    # ---- device-level (what ShardedSystem calls) -----------------------------------------------------
    def padded_bodies(self, n):
        return max(32, -(-n // 32) * 32)

    def to_device(self, arr, dtype=None, pinned=True):
        t = torch.from_numpy(np.ascontiguousarray(arr).copy())
        return t.to(dtype) if dtype is not None else t

    def _masses_dev(self, masses):
        m = np.asarray(masses)
        return torch.from_numpy(np.ascontiguousarray(m).copy()), int(m.dtype == np.float32)

    def workspace(self, n, n_i, dtype):
        return (torch.empty(1), 0)

    def pack(self, pos_dev, masses_dev, masses_f32, n, dtype, out=None):
        stream = out if out is not None else torch.zeros(self.padded_bodies(n) * 4, dtype=torch.float64)
        s = stream.view(-1, 4)
        s[:n, :3] = pos_dev.to(stream.dtype)
        s[:n, 3] = (G * masses_dev.to(torch.float64)).to(stream.dtype)
        s[n:self.padded_bodies(n)] = 0
        self._masses = masses_dev.clone()
        return stream

    def unpack(self, stream, n):
        return stream.view(-1, 4)[:n, :3].to(torch.float64).clone()

    def _pos(self, stream, n):
        return stream.view(-1, 4)[:n, :3].to(torch.float64).numpy()

    def accel_slab(self, stream, n, i0, n_i, softening, ws=None):
        acc = oracle.accel_direct_rows(self._pos(stream, n), self._masses.numpy(), i0, n_i, softening)
        self.launches += 2
        return torch.from_numpy(acc).to(stream.dtype)

    def kick_drift_slab(self, cur, nxt, vel, acc, n, i0, n_i, dt):
        half = 0.5 * dt
        vel += half * acc
        nxt.view(-1, 4)[i0:i0 + n_i, :3] = cur.view(-1, 4)[i0:i0 + n_i, :3] + dt * vel
        self.launches += 1

    def step_slab(self, cur, nxt, vel, acc, n, i0, n_i, dt, softening, flags, snap_pos, snap_vel, snap_acc, ws):
        half = 0.5 * dt
        a = self.accel_slab(cur, n, i0, n_i, softening)
        acc.copy_(a)
        vel += half * acc
        if flags & NB_STEP_SNAPSHOT and snap_pos is not None:
            snap_pos[i0:i0 + n_i] = cur.view(-1, 4)[i0:i0 + n_i, :3]
            snap_vel[i0:i0 + n_i] = vel
            snap_acc[i0:i0 + n_i] = acc
        if flags & NB_STEP_CONTINUE:
            vel += half * acc
            nxt.view(-1, 4)[i0:i0 + n_i, :3] = cur.view(-1, 4)[i0:i0 + n_i, :3] + dt * vel

    def energy_slab(self, pos_d, vel_d, m_d, masses_f32, n, i0, n_i, softening):
        pos, vel, m = pos_d.numpy(), vel_d.numpy(), m_d.numpy().astype(np.float64)
        k = 0.5 * (m[i0:i0 + n_i] * (vel[i0:i0 + n_i] ** 2).sum(axis=1)).sum()
        u = 0.0
        for i in range(i0, i0 + n_i):
            d = pos - pos[i]
            r = np.sqrt((d * d).sum(axis=1) + softening * softening)
            t = G * m * m[i] / r
            t[i] = 0.0
            u -= 0.5 * t.sum()
        return torch.tensor([k, u], dtype=torch.float64)
