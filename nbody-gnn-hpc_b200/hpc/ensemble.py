"""Ensembles of independent simulations in one kernel launch.

The reference generates its training data by mapping ``generate_single_simulation`` over an
``mp.Pool`` (reference scripts/generate_data.py:32-58,142-149) and its evaluation ground truth
with a Python loop of simulators (scripts/evaluate.py:81-93).  Here the whole batch is one launch
of the persistent ensemble kernel (K3, ``csrc/nb_ensemble.cu``); the host only draws the initial
conditions with NumPy -- in the reference's order -- and slices the stacked result.
"""
from __future__ import annotations

import numpy as np

from . import nbody
from .ics import reference_default_ic

__all__ = ["simulate_ensemble", "generate_simulations", "generate_single_simulation"]


def simulate_ensemble(positions, velocities, masses, dt: float = 1e-3, softening: float = nbody.SOFTENING,
                      n_steps: int = 400, save_interval: int = 1, *, dtype=None, accelerations=None,
                      snapshots: bool = True, device=None, outputs: str = "host", fields=None) -> dict:
    """Advance B independent systems of N bodies by n_steps kick-drift-kick steps.

    positions, velocities: (B,N,3); masses: (N,) shared by all systems or (B,N).
    Returns float64 arrays: 'positions', 'velocities', 'accelerations' of shape
    (B, 1 + n_steps//save_interval, N, 3) -- per system exactly what the reference's
    ``np.stack([s[key] for s in sim.run(...)])`` yields -- plus 'times' and the final state.
    outputs="device" leaves the stacks in HBM as torch tensors (72*N bytes per system-step never cross PCIe):
    the input of the window kernel (``hpc.checkpoint.sliding_windows_device``) and of anything else on the GPU.
    fields=("positions", "velocities") ships only those stacks with the run; the others stay in HBM and are copied
    when the result's entry is first read (the run is PCIe-bound: a stack nobody reads is a third of its time).
    """
    if outputs not in ("host", "device"):
        raise ValueError("outputs must be 'host' or 'device'")
    eng = nbody._cuda.get_engine(device)
    kw = {"outputs": outputs} if outputs != "host" else {}
    if fields is not None:
        kw["fields"] = fields
    out = eng.ensemble(positions, velocities, masses, float(dt), float(softening), int(n_steps), int(save_interval),
                       dtype=nbody._engine_dtype(dtype), a0=accelerations, snapshots=snapshots, **kw)
    t, times = 0.0, [0.0]
    for k in range(1, n_steps + 1):
        t += dt                                   # running sum, reference nbody.py:217
        if k % save_interval == 0:
            times.append(t)
    out["times"] = np.array(times)
    return out


def generate_simulations(sim_args: list, *, dtype=None, device=None) -> list:
    """Batch form of the reference's ``generate_single_simulation`` (generate_data.py:32-58).

    sim_args: list of tuples (sim_id, n_particles, n_steps, save_interval, box_size, seed,
    shared_masses) exactly as generate_data.py:131-134 builds them.  Simulations that share
    (n_particles, n_steps, save_interval) advance together in one launch.  Returns the list of
    per-simulation dicts the reference function returns, in input order.
    """
    results = [None] * len(sim_args)
    groups: dict = {}
    for idx, a in enumerate(sim_args):
        groups.setdefault((a[1], a[2], a[3]), []).append(idx)
    for (n, n_steps, save_interval), idxs in groups.items():
        x0 = np.empty((len(idxs), n, 3))
        v0 = np.empty((len(idxs), n, 3))
        m = []
        for r, idx in enumerate(idxs):
            _sid, _n, _ns, _si, box_size, seed, shared = sim_args[idx]
            x0[r], v0[r], own = reference_default_ic(n, seed, box_size)   # ctor draws, nbody.py:175-181
            m.append(own if shared is None else np.asarray(shared))       # generate_data.py:45-46
        same = all(mm is m[0] or (mm.dtype == m[0].dtype and np.array_equal(mm, m[0])) for mm in m)
        masses = m[0] if same else np.stack([np.asarray(mm, dtype=np.float64) for mm in m])
        out = simulate_ensemble(x0, v0, masses, dt=0.001, softening=nbody.SOFTENING, n_steps=n_steps,
                                save_interval=save_interval, dtype=dtype, device=device)
        for r, idx in enumerate(idxs):
            results[idx] = {
                'positions': out['positions'][r],
                'velocities': out['velocities'][r],
                'accelerations': out['accelerations'][r],
                'masses': m[r].copy(),
                'times': out['times'].copy(),
                'n_steps': out['positions'].shape[1],
            }
    return results


def generate_single_simulation(args):
    """Same signature and result as the reference's worker function (generate_data.py:32-58)."""
    return generate_simulations([args])[0]
