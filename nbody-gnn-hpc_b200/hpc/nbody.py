"""N-body gravitational simulation on NVIDIA B200 -- drop-in for the reference's ``hpc.nbody``.

Same names, arguments and results as ``/root/reference/src/hpc/nbody.py`` so that
``scripts/generate_data.py`` and ``scripts/evaluate.py`` of the reference run unchanged with this
package directory on ``sys.path`` in place of the reference's ``src``:

    NBodySimulator, compute_accelerations_direct, compute_total_energy, leapfrog_step,
    run_parallel_simulations, G, SOFTENING

The arithmetic runs in hand-written sm_100a CUDA kernels behind a C ABI (``include/nbody_b200.h``,
bound in ``hpc/_cuda.py``).  There is no CPU fallback: without the built library or without a
CUDA device every force evaluation raises ``EngineUnavailable``.

Where the state lives.  ``positions``, ``velocities``, ``accelerations`` and ``masses`` are ordinary NumPy
arrays whenever the caller looks at them or assigns them -- exactly the reference's attributes -- but between
``step()`` / ``run()`` calls the state stays in device memory: it is uploaded when the caller has assigned or may
have written the host arrays since the last upload, and downloaded when the caller next reads an attribute (or,
after every call, while the caller still holds a reference to the position / velocity arrays, so that such aliases
see the in-place updates of reference nbody.py:205-214).  ``for _ in range(k): sim.step()`` therefore costs k kernel
launches and no copies; ``run()`` returns a list whose state dictionaries are built when they are first accessed.
"""
from __future__ import annotations

import os
import sys
import warnings
from typing import Optional, Tuple

import numpy as np

from . import _cuda

# Physical constants (reference nbody.py:18-19)
G = 6.67430e-11
SOFTENING = 1e-9

_bh_warned = False


def _engine_dtype(dtype) -> np.dtype:
    if dtype is None:
        dtype = os.environ.get("NBODY_DTYPE", "float64")
    d = np.dtype(dtype)
    if d not in (np.dtype(np.float64), np.dtype(np.float32)):
        raise ValueError(f"dtype must be float64 or float32, got {d}")
    return d


def compute_accelerations_direct(positions: np.ndarray, masses: np.ndarray,
                                 softening: float = SOFTENING, *, dtype=None) -> np.ndarray:
    """Direct O(N^2) softened gravitational accelerations (reference nbody.py:22-66).

    positions (N,3), masses (N,) float64 or float32 -> new (N,3) float64 array.
    ``dtype`` selects the kernel precision (float64 default; float32 is the fast kernel).
    """
    positions = np.asarray(positions)
    if positions.ndim != 2 or positions.shape[1] != 3:
        raise ValueError(f"positions must have shape (N, 3), got {positions.shape}")
    masses = np.asarray(masses)
    if masses.shape != (positions.shape[0],):
        raise ValueError(f"masses must have shape ({positions.shape[0]},), got {masses.shape}")
    if positions.shape[0] == 0:
        return np.zeros((0, 3))
    return _cuda.get_engine().accelerations(positions, masses, float(softening), _engine_dtype(dtype))


def compute_total_energy(positions: np.ndarray, velocities: np.ndarray, masses: np.ndarray,
                         softening: float = SOFTENING) -> Tuple[float, float, float]:
    """(kinetic, potential, total) energy (reference nbody.py:101-130), float64 on the device."""
    return _cuda.get_engine().energy(np.asarray(positions), np.asarray(velocities), np.asarray(masses),
                                     float(softening))


def leapfrog_step(positions: np.ndarray, velocities: np.ndarray, accelerations: np.ndarray,
                  masses: np.ndarray, dt: float):
    """Opening half of a leapfrog step (reference nbody.py:69-98; never called by the reference).

    Returns (new_positions, velocities_half, accelerations) -- the accelerations are passed
    through, as in the reference.  O(N) host arithmetic, not part of the hot path.
    """
    velocities_half = velocities + 0.5 * dt * accelerations
    new_positions = positions + dt * velocities_half
    return new_positions, velocities_half, accelerations


def _refs(holder: dict, name: str) -> int:
    return sys.getrefcount(holder[name])


# what _refs reports for an array nobody but its holder refers to (calibrated, not assumed)
_UNSHARED_REFS = _refs({"x": np.zeros(1)}, "x")


class _Pending:
    """Snapshot stacks of one device segment: a callable that waits for them, then the stacks themselves."""
    __slots__ = ("_wait", "_value")

    def __init__(self, wait):
        self._wait, self._value = wait, None

    def get(self) -> dict:
        if self._wait is not None:
            self._value, self._wait = self._wait(), None
        return self._value


class StateList(list):
    """The list ``NBodySimulator.run`` returns (reference nbody.py:232-248): one state dictionary per saved step,
    built when it is first accessed.  The snapshots arrive from the device as three stacked arrays; a state's
    'positions' / 'velocities' / 'accelerations' are the rows of those stacks (disjoint, writable, owned by this
    list alone -- independent of each other and of the simulator, like the reference's copies).  Behaves as the
    plain list it subclasses in every other respect."""

    def __init__(self, first: dict, masses: np.ndarray):
        super().__init__([first])
        self._masses = masses
        self._lazy = []          # (index in the list, stacks, row, time, step) not yet materialised

    def _append_rows(self, stacks: dict, rows, times, steps) -> None:
        for r, t, k in zip(rows, times, steps):
            self._lazy.append((len(self), stacks, r, t, k))
            super().append(None)

    def _materialise(self) -> None:
        if self._lazy:
            for idx, pending, r, t, k in self._lazy:
                stacks = pending.get()
                super().__setitem__(idx, {
                    'positions': stacks['positions'][r],
                    'velocities': stacks['velocities'][r],
                    'accelerations': stacks['accelerations'][r],
                    'masses': self._masses.copy(),
                    'time': t,
                    'step': k,
                })
            self._lazy = []

    def __getitem__(self, i):
        self._materialise()
        return super().__getitem__(i)

    def __iter__(self):
        self._materialise()
        return super().__iter__()

    def __reversed__(self):
        self._materialise()
        return super().__reversed__()

    def __contains__(self, x):
        self._materialise()
        return super().__contains__(x)

    def __eq__(self, other):
        self._materialise()
        return super().__eq__(other)

    __hash__ = None

    def __repr__(self):
        self._materialise()
        return super().__repr__()

    def __add__(self, other):
        self._materialise()
        return list(super().__iter__()) + list(other)

    def __reduce__(self):
        self._materialise()
        return (list, (list(super().__iter__()),))

    def copy(self):
        self._materialise()
        return list(super().__iter__())

    def index(self, *a):
        self._materialise()
        return super().index(*a)

    def count(self, x):
        self._materialise()
        return super().count(x)

    def pop(self, *a):
        self._materialise()
        return super().pop(*a)

    def sort(self, **kw):
        self._materialise()
        return super().sort(**kw)

    def reverse(self):
        self._materialise()
        return super().reverse()

    def insert(self, *a):
        self._materialise()
        return super().insert(*a)

    def remove(self, x):
        self._materialise()
        return super().remove(x)

    def __setitem__(self, i, v):
        self._materialise()
        return super().__setitem__(i, v)

    def __delitem__(self, i):
        self._materialise()
        return super().__delitem__(i)


class NBodySimulator:
    """High-performance N-body gravitational simulator (reference nbody.py:133-337).

    Extra keyword-only options: ``dtype`` ('float64' default, or 'float32' for the fast kernels;
    env NBODY_DTYPE overrides the default) and ``device`` (CUDA device index).
    """

    _ARRAYS = ("positions", "velocities", "accelerations")

    def __init__(self,
                 n_particles: int = 1000,
                 box_size: float = 1.0,
                 mass_range: Tuple[float, float] = (1e10, 1e12),
                 dt: float = 1e-3,
                 softening: float = SOFTENING,
                 use_barnes_hut: bool = False,
                 theta: float = 0.5,
                 seed: Optional[int] = None,
                 *, dtype=None, device=None):
        global _bh_warned
        self.n_particles = n_particles
        self.box_size = box_size
        self.dt = dt
        self.softening = softening
        # The flag is kept for call compatibility (generate_data.py:41 sets it for N > 500; reference
        # nbody.py:193-198 then walks a theta = 0.5 tree).  This engine evaluates the exact direct sum the tree
        # approximates -- O(N^2), but 2.7e12 interactions/s -- and says so once per process.
        self.use_barnes_hut = use_barnes_hut
        if use_barnes_hut and not _bh_warned:
            _bh_warned = True
            warnings.warn(
                "use_barnes_hut=True: the B200 engine evaluates the exact direct sum at every N (theta is ignored), "
                "so trajectories are the direct-sum ones, not the reference's theta=0.5 tree approximation "
                "(which also ignores the simulator's softening, reference nbody.py:197-198)", stacklevel=2)
        self.theta = theta
        self.seed = seed
        self.dtype = _engine_dtype(dtype)
        self.device = device

        self._host = {}            # the host arrays the caller sees
        self._resident = None      # _cuda.ResidentSystem holding the state on the device, or None
        self._host_stale = False   # the device state is newer than the host arrays
        self._host_touched = True  # the caller assigned, or may have written, host arrays since the last upload

        # same draws, same order, same (global) generator as reference nbody.py:175-181
        if seed is not None:
            np.random.seed(seed)
        self.positions = (np.random.rand(n_particles, 3) - 0.5) * box_size
        self.velocities = (np.random.rand(n_particles, 3) - 0.5) * 0.1 * box_size
        self.masses = np.random.uniform(mass_range[0], mass_range[1], n_particles)

        self.accelerations = self._compute_accelerations()

        self.time = 0.0
        self.step_count = 0
        self.history = []

    # ---- state attributes: host arrays on demand, device memory in between -------------------------------------
    def _read(self, name: str) -> np.ndarray:
        """Attribute read by the CALLER: the host array, refreshed from the device if it is stale.  The caller may
        write through the returned reference (the reference's factories do, nbody.py:294-300), so the host copy
        counts as touched."""
        self._sync_host()
        self._host_touched = True
        return self._host[name]

    def _write(self, name: str, value) -> None:
        if self._host_stale:
            self._sync_host()              # the other attributes must be current before the host becomes the truth
        self._host[name] = value
        self._host_touched = True

    positions = property(lambda self: self._read("positions"), lambda self, v: self._write("positions", v))
    velocities = property(lambda self: self._read("velocities"), lambda self, v: self._write("velocities", v))
    accelerations = property(lambda self: self._read("accelerations"), lambda self, v: self._write("accelerations", v))
    masses = property(lambda self: self._read("masses"), lambda self, v: self._write("masses", v))

    def _sync_host(self) -> None:
        """Bring the host arrays up to date: positions and velocities IN PLACE (the reference's += updates,
        nbody.py:205-214, keep aliases valid), accelerations rebound (nbody.py:211)."""
        if not self._host_stale:
            return
        pos, vel, acc = self._resident.download()
        h = self._host
        for name, new in (("positions", pos), ("velocities", vel)):
            old = h.get(name)
            if isinstance(old, np.ndarray) and old.shape == new.shape and old.dtype == new.dtype and old.flags.writeable:
                old[...] = new
            else:
                h[name] = new
        h["accelerations"] = acc
        self._host_stale = False

    def _aliased(self) -> bool:
        """Does anybody but this object hold a reference to the position / velocity host arrays (or a view of
        them)?  Such an alias must see every step's in-place update, and may be written through at any time.
        CPython reference counts against a baseline measured at import; anything that holds extra references (a
        debugger, a tracer) only errs towards 'aliased', i.e. towards the reference's eager behaviour."""
        h = self._host
        return _refs(h, "positions") > _UNSHARED_REFS or _refs(h, "velocities") > _UNSHARED_REFS

    def _device_state(self):
        """The resident device state, (re)built from the host arrays when the caller has touched them."""
        if self._resident is None or self._host_touched or self._aliased():
            self._sync_host()
            h = self._host
            self._resident = self._engine().resident(h["positions"], h["velocities"], h["accelerations"], h["masses"],
                                                     float(self.dt), float(self.softening), self.dtype)
            self._host_touched = False
        else:
            self._resident.dt, self._resident.softening = float(self.dt), float(self.softening)
        return self._resident

    # ------------------------------------------------------------------------------------------
    def _engine(self):
        return _cuda.get_engine(self.device)

    def _compute_accelerations(self) -> np.ndarray:
        """Accelerations of the current host state (reference nbody.py:193-200)."""
        if self.n_particles == 0:
            return np.zeros((0, 3))
        self._sync_host()
        h = self._host
        return self._engine().accelerations(h["positions"], h["masses"], float(self.softening), self.dtype)

    def _advance(self, n_steps: int, save_interval: int, snapshots: bool):
        """n_steps on the device; the live state stays there (host arrays refreshed only while aliased).  With
        snapshots, returns a callable that waits for the device and hands out the host stacks."""
        rs = self._device_state()
        pending = rs.advance_async(int(n_steps), int(save_interval), snapshots=snapshots)
        self._host_stale = True
        t, dt = self.time, self.dt
        for _ in range(n_steps):
            t += dt                       # a running float sum, reference nbody.py:217
        self.time = t
        self.step_count += n_steps
        if self._aliased():
            self._sync_host()             # in place, as the reference's  +=  updates are
        return pending

    def step(self) -> None:
        """Advance the simulation by one kick-drift-kick step (reference nbody.py:202-218)."""
        self._device_state().step()
        self._host_stale = True
        self.time += self.dt              # reference nbody.py:217
        self.step_count += 1
        if self._aliased():
            self._sync_host()             # in place, as the reference's  +=  updates are

    def run(self, n_steps: int, save_interval: int = 1, verbose: bool = True) -> list:
        """Run n_steps, returning the list of saved states (reference nbody.py:220-248).

        State 0 is the state on entry; one more state per ``save_interval`` steps.  With
        ``verbose`` the energy is printed every max(1, n_steps // 10) steps, which splits the run
        into that many device segments.
        """
        s0 = self.step_count
        masses = self._host["masses"]
        states = StateList(self.get_state(), masses)
        if n_steps <= 0:
            self.history = states
            return states
        report = max(1, n_steps // 10)
        # times[k]: the float the reference holds after k additions of dt (nbody.py:217)
        times = [self.time]
        for _ in range(n_steps):
            times.append(times[-1] + self.dt)
        done = 0
        while done < n_steps:
            stop = min(n_steps, (done // report + 1) * report) if verbose else n_steps
            seg = stop - done
            phase = done % save_interval
            if phase:  # a report point fell between two save points: walk to the next save point first
                lead = min(save_interval - phase, seg)
                out = _Pending(self._advance(lead, lead, snapshots=True))
                done += lead
                seg -= lead
                if done % save_interval == 0:
                    states._append_rows(out, [1], [times[done]], [s0 + done])
                out.get()                 # the pinned staging block is reused by the next segment
            if seg > 0:
                out = _Pending(self._advance(seg, save_interval, snapshots=True))
                rows = range(1, seg // save_interval + 1)
                ks = [done + r * save_interval for r in rows]
                states._append_rows(out, rows, [times[k] for k in ks], [s0 + k for k in ks])   # overlaps the GPU
                done += seg
                out.get()                 # the segment is complete (and its snapshots on the host) before we go on
            if verbose and done % report == 0:
                energy = self.get_energy()
                print(f"Step {done}/{n_steps}, Time: {self.time:.4f}, Energy: {energy[2]:.6e}")
        self.history = states
        return states

    def get_state(self) -> dict:
        """Current simulation state as a dictionary of copies (reference nbody.py:250-259)."""
        self._sync_host()
        h = self._host
        return {
            'positions': h["positions"].copy(),
            'velocities': h["velocities"].copy(),
            'accelerations': h["accelerations"].copy(),
            'masses': h["masses"].copy(),
            'time': self.time,
            'step': self.step_count
        }

    def set_state(self, state: dict) -> None:
        """Restore the simulation from a state dictionary (reference nbody.py:261-268)."""
        self._host_stale = False           # whatever the device holds is superseded
        self.positions = state['positions'].copy()
        self.velocities = state['velocities'].copy()
        self.accelerations = state['accelerations'].copy()
        self.masses = state['masses'].copy()
        self.time = state['time']
        self.step_count = state['step']
        self.n_particles = self._host["positions"].shape[0]

    def get_energy(self) -> Tuple[float, float, float]:
        """Current (kinetic, potential, total) energy (reference nbody.py:270-273)."""
        if self._host_stale and not self._host_touched:
            return self._resident.energy()          # the state is on the device: evaluate it there
        h = self._host
        return self._engine().energy(h["positions"], h["velocities"], h["masses"], float(self.softening))

    # ------------------------------------------------------------------------------------------
    @classmethod
    def create_solar_system(cls, scale: float = 1.0) -> 'NBodySimulator':
        """Sun + 8 planets on circular-orbit speeds (reference nbody.py:275-303)."""
        sim = cls(n_particles=9, box_size=50.0, dt=0.01)
        # (name, mass [solar masses], distance [AU], orbital speed [km/s])
        bodies = [
            ('Sun', 1.0, 0.0, 0.0),
            ('Mercury', 1.66e-7, 0.39, 47.87),
            ('Venus', 2.45e-6, 0.72, 35.02),
            ('Earth', 3.00e-6, 1.0, 29.78),
            ('Mars', 3.23e-7, 1.52, 24.07),
            ('Jupiter', 9.55e-4, 5.2, 13.07),
            ('Saturn', 2.86e-4, 9.58, 9.69),
            ('Uranus', 4.37e-5, 19.22, 6.81),
            ('Neptune', 5.15e-5, 30.05, 5.43),
        ]
        sim.masses = np.array([b[1] for b in bodies]) * 1.989e30 * scale
        sim.positions = np.zeros((9, 3))
        sim.velocities = np.zeros((9, 3))
        for i, (_name, _mass, dist, vel) in enumerate(bodies):
            sim.positions[i, 0] = dist * 1.496e11 * scale
            sim.velocities[i, 1] = vel * 1000 * scale
        sim.accelerations = sim._compute_accelerations()
        return sim

    @classmethod
    def create_galaxy_collision(cls, n_per_galaxy: int = 500) -> 'NBodySimulator':
        """Two exponential discs on a collision course (reference nbody.py:305-337)."""
        n_total = 2 * n_per_galaxy
        sim = cls(n_particles=n_total, box_size=100.0, dt=0.01)
        for sl, centre, vx in ((slice(0, n_per_galaxy), -20.0, 2.0), (slice(n_per_galaxy, n_total), 20.0, -2.0)):
            theta = np.random.rand(n_per_galaxy) * 2 * np.pi
            r = np.random.exponential(5.0, n_per_galaxy)
            sim.positions[sl, 0] = centre + r * np.cos(theta)
            sim.positions[sl, 1] = r * np.sin(theta)
            sim.positions[sl, 2] = np.random.randn(n_per_galaxy) * 0.5
            sim.velocities[sl, 0] = vx
        # rotation is added to the first galaxy only, about the origin -- as the reference does (:330-334)
        for i in range(n_per_galaxy):
            r = np.sqrt(sim.positions[i, 0] ** 2 + sim.positions[i, 1] ** 2)
            if r > 0:
                sim.velocities[i, 0] += -sim.positions[i, 1] / r * 0.5
                sim.velocities[i, 1] += sim.positions[i, 0] / r * 0.5
        sim.accelerations = sim._compute_accelerations()
        return sim


def run_parallel_simulations(configs: list, n_workers: int = None) -> list:
    """Run several simulations and return their histories (reference nbody.py:340-362).

    The reference maps a local closure over an mp.Pool (which cannot pickle it).  One GPU already
    runs the simulations faster than a pool of CPU workers, so they are executed in order on the
    current device; ``n_workers`` is accepted and ignored.  For many equal-sized systems use
    ``hpc.ensemble.simulate_ensemble``, which advances all of them in one kernel launch.
    """
    results = []
    for config in configs:
        sim = NBodySimulator(**config.get('init', {}))
        results.append(sim.run(**config.get('run', {})))
    return results
