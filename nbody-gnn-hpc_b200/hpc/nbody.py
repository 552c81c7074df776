"""N-body gravitational simulation on NVIDIA B200 -- drop-in for the reference's ``hpc.nbody``.

Same names, arguments and results as ``/root/reference/src/hpc/nbody.py`` so that
``scripts/generate_data.py`` and ``scripts/evaluate.py`` of the reference run unchanged with this
package directory on ``sys.path`` in place of the reference's ``src``:

    NBodySimulator, compute_accelerations_direct, compute_total_energy, leapfrog_step,
    run_parallel_simulations, G, SOFTENING

The arithmetic runs in hand-written sm_100a CUDA kernels behind a C ABI (``include/nbody_b200.h``,
bound in ``hpc/_cuda.py``).  There is no CPU fallback: without the built library or without a
CUDA device every force evaluation raises ``EngineUnavailable``.

Host NumPy arrays stay the source of truth between calls, exactly as in the reference: callers
assign ``sim.masses`` / ``sim.positions`` (or write into slices of them) and then call
``_compute_accelerations()`` or ``run()``.  ``run()`` uploads the state once, advances all steps on
the device with snapshots written by the kernels, and downloads the stacked snapshots once.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from . import _cuda

# Physical constants (reference nbody.py:18-19)
G = 6.67430e-11
SOFTENING = 1e-9

_backend_override = None


def _backend():
    """The object that executes the hot path: the CUDA engine (tests may install a stand-in)."""
    if _backend_override is not None:
        return _backend_override
    return _cuda.get_engine()


def _set_backend_for_tests(backend) -> None:
    """Install a stand-in backend (tests of the host-side bookkeeping on machines without a GPU)."""
    global _backend_override
    _backend_override = backend


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
    return _backend().accelerations(positions, masses, float(softening), _engine_dtype(dtype))


def compute_total_energy(positions: np.ndarray, velocities: np.ndarray, masses: np.ndarray,
                         softening: float = SOFTENING) -> Tuple[float, float, float]:
    """(kinetic, potential, total) energy (reference nbody.py:101-130), float64 on the device."""
    return _backend().energy(np.asarray(positions), np.asarray(velocities), np.asarray(masses), float(softening))


def leapfrog_step(positions: np.ndarray, velocities: np.ndarray, accelerations: np.ndarray,
                  masses: np.ndarray, dt: float):
    """Opening half of a leapfrog step (reference nbody.py:69-98; never called by the reference).

    Returns (new_positions, velocities_half, accelerations) -- the accelerations are passed
    through, as in the reference.  O(N) host arithmetic, not part of the hot path.
    """
    velocities_half = velocities + 0.5 * dt * accelerations
    new_positions = positions + dt * velocities_half
    return new_positions, velocities_half, accelerations


class NBodySimulator:
    """High-performance N-body gravitational simulator (reference nbody.py:133-337).

    Extra keyword-only options: ``dtype`` ('float64' default, or 'float32' for the fast kernels;
    env NBODY_DTYPE overrides the default) and ``device`` (CUDA device index).
    """

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
        self.n_particles = n_particles
        self.box_size = box_size
        self.dt = dt
        self.softening = softening
        # The flag is kept for call compatibility (generate_data.py:41 sets it for N > 500); the
        # engine always evaluates the exact direct sum, which the tree code approximates.
        self.use_barnes_hut = use_barnes_hut
        self.theta = theta
        self.seed = seed
        self.dtype = _engine_dtype(dtype)
        self.device = device

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

    # ------------------------------------------------------------------------------------------
    def _engine(self):
        if _backend_override is not None:
            return _backend_override
        return _cuda.get_engine(self.device)

    def _compute_accelerations(self) -> np.ndarray:
        """Accelerations of the current host state (reference nbody.py:193-200)."""
        if self.n_particles == 0:
            return np.zeros((0, 3))
        return self._engine().accelerations(self.positions, self.masses, float(self.softening), self.dtype)

    def _advance(self, n_steps: int, save_interval: int, snapshots: bool) -> Optional[dict]:
        """Upload (x, v, a), advance n_steps on the device, download; updates the live state."""
        out = self._engine().run(self.positions, self.velocities, self.accelerations, self.masses, float(self.dt),
                                 float(self.softening), int(n_steps), int(save_interval), dtype=self.dtype,
                                 snapshots=snapshots)
        # in-place, as the reference's  +=  updates are (aliases of the arrays stay valid)
        self.positions[...] = out["final_positions"]
        self.velocities[...] = out["final_velocities"]
        self.accelerations = np.array(out["final_accelerations"])  # rebinding, reference nbody.py:211
        for _ in range(n_steps):
            self.time += self.dt          # a running float sum, reference nbody.py:217
        self.step_count += n_steps
        return out if snapshots else None

    def step(self) -> None:
        """Advance the simulation by one kick-drift-kick step (reference nbody.py:202-218)."""
        self._advance(1, 1, snapshots=False)

    def run(self, n_steps: int, save_interval: int = 1, verbose: bool = True) -> list:
        """Run n_steps, returning the list of saved states (reference nbody.py:220-248).

        State 0 is the state on entry; one more state per ``save_interval`` steps.  With
        ``verbose`` the energy is printed every max(1, n_steps // 10) steps, which splits the run
        into that many device segments.
        """
        s0 = self.step_count
        states = [self.get_state()]
        if n_steps <= 0:
            self.history = states
            return states
        report = max(1, n_steps // 10)
        masses = self.masses
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
                self._advance(lead, 1, snapshots=False)
                done += lead
                seg -= lead
                if done % save_interval == 0:
                    states.append(self._state_from(self.positions, self.velocities, self.accelerations, masses,
                                                   times[done], s0 + done))
            if seg > 0:
                out = self._advance(seg, save_interval, snapshots=True)
                for r in range(1, seg // save_interval + 1):
                    k = done + r * save_interval
                    states.append(self._state_from(out["positions"][r], out["velocities"][r],
                                                   out["accelerations"][r], masses, times[k], s0 + k))
                done += seg
            if verbose and done % report == 0:
                energy = self.get_energy()
                print(f"Step {done}/{n_steps}, Time: {self.time:.4f}, Energy: {energy[2]:.6e}")
        self.history = states
        return states

    @staticmethod
    def _state_from(pos, vel, acc, masses, time, step) -> dict:
        return {
            'positions': np.array(pos),
            'velocities': np.array(vel),
            'accelerations': np.array(acc),
            'masses': masses.copy(),
            'time': time,
            'step': step,
        }

    def get_state(self) -> dict:
        """Current simulation state as a dictionary of copies (reference nbody.py:250-259)."""
        return {
            'positions': self.positions.copy(),
            'velocities': self.velocities.copy(),
            'accelerations': self.accelerations.copy(),
            'masses': self.masses.copy(),
            'time': self.time,
            'step': self.step_count
        }

    def set_state(self, state: dict) -> None:
        """Restore the simulation from a state dictionary (reference nbody.py:261-268)."""
        self.positions = state['positions'].copy()
        self.velocities = state['velocities'].copy()
        self.accelerations = state['accelerations'].copy()
        self.masses = state['masses'].copy()
        self.time = state['time']
        self.step_count = state['step']
        self.n_particles = self.positions.shape[0]

    def get_energy(self) -> Tuple[float, float, float]:
        """Current (kinetic, potential, total) energy (reference nbody.py:270-273)."""
        return self._engine().energy(self.positions, self.velocities, self.masses, float(self.softening))

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
