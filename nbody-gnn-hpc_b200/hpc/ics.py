"""Seeded initial conditions for tests, smoke and bench (host NumPy only).

* ``reference_default_ic`` reproduces what the reference simulator draws in its
  constructor (/root/reference/src/hpc/nbody.py:175-181: positions, velocities,
  masses, in that order, from the seeded *global* NumPy generator) without
  touching the global generator, and ``shared_masses`` reproduces the float32
  mass vector the data-generation and evaluation scripts assign afterwards
  (/root/reference/scripts/generate_data.py:108-109, scripts/evaluate.py:76-77).
* ``plummer_ic`` and ``uniform_sphere_ic`` are well-conditioned systems in
  N-body units (G*M = 1), the only ones on which a 400-step position tolerance
  is meaningful.
"""
from __future__ import annotations

import numpy as np

G = 6.67430e-11  # nbody.py:18


def reference_default_ic(n_particles: int, seed: int, box_size: float = 10.0,
                         mass_range=(1e10, 1e12)):
    """(positions, velocities, masses) exactly as NBodySimulator.__init__ draws them.

    np.random.seed(seed) followed by rand/rand/uniform on the global generator is
    the same stream as RandomState(seed) (legacy MT19937), so this does not
    disturb global state.
    """
    rng = np.random.RandomState(seed)
    positions = (rng.rand(n_particles, 3) - 0.5) * box_size                 # nbody.py:179
    velocities = (rng.rand(n_particles, 3) - 0.5) * 0.1 * box_size          # nbody.py:180
    masses = rng.uniform(mass_range[0], mass_range[1], n_particles)         # nbody.py:181
    return positions, velocities, masses


def shared_masses(n_particles: int, seed: int = 42) -> np.ndarray:
    """float32 masses shared by every simulation of a data-generation run."""
    return np.random.RandomState(seed).uniform(1e10, 1e12, n_particles).astype(np.float32)


def datagen_ensemble_ic(n_sims: int, n_particles: int, seed: int = 42, box_size: float = 10.0,
                        first_sim: int = 0):
    """Stacked ICs of simulations first_sim .. first_sim+n_sims-1 of a generate_data.py run.

    Simulation i uses seed ``seed + i`` (generate_data.py:131-134) and the shared
    float32 masses.  Returns x0, v0 of shape (B, N, 3) float64 and masses (N,) float32.
    """
    x0 = np.empty((n_sims, n_particles, 3))
    v0 = np.empty((n_sims, n_particles, 3))
    for b in range(n_sims):
        x0[b], v0[b], _ = reference_default_ic(n_particles, seed + first_sim + b, box_size)
    return x0, v0, shared_masses(n_particles, seed)


def plummer_ic(n_particles: int, seed: int = 7, r_max: float = 20.0):
    """Plummer sphere (Aarseth, Henon & Wielen 1974), scale a = 1, G*M = 1.

    Equal masses m = 1/(G*N), centre of mass and mean velocity removed.
    """
    rng = np.random.RandomState(seed)
    n = n_particles
    # radii from the inverse cumulative mass profile, clipped at r_max
    u = rng.uniform(0.0, 1.0, n)
    u = np.clip(u, 1e-10, (r_max ** 3) / (1.0 + r_max * r_max) ** 1.5)
    r = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    cos_t = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    sin_t = np.sqrt(1.0 - cos_t * cos_t)
    pos = np.stack([r * sin_t * np.cos(phi), r * sin_t * np.sin(phi), r * cos_t], axis=1)
    # speeds: q = v / v_esc by von Neumann rejection on g(q) = q^2 (1 - q^2)^(7/2)
    q = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        x = rng.uniform(0.0, 1.0, todo.size)
        y = rng.uniform(0.0, 0.1, todo.size)
        ok = y < x * x * (1.0 - x * x) ** 3.5
        q[todo[ok]] = x[ok]
        todo = todo[~ok]
    v = q * np.sqrt(2.0) * (1.0 + r * r) ** (-0.25)
    cos_t = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    sin_t = np.sqrt(1.0 - cos_t * cos_t)
    vel = np.stack([v * sin_t * np.cos(phi), v * sin_t * np.sin(phi), v * cos_t], axis=1)
    pos -= pos.mean(axis=0)
    vel -= vel.mean(axis=0)
    masses = np.full(n, 1.0 / (G * n))
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel), masses


def uniform_sphere_ic(n_particles: int, seed: int = 11, virial: bool = True):
    """Uniform-density unit sphere, G*M = 1, equal masses; cold or virialised (2K = |U|)."""
    rng = np.random.RandomState(seed)
    n = n_particles
    r = rng.uniform(0.0, 1.0, n) ** (1.0 / 3.0)
    cos_t = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    sin_t = np.sqrt(1.0 - cos_t * cos_t)
    pos = np.stack([r * sin_t * np.cos(phi), r * sin_t * np.sin(phi), r * cos_t], axis=1)
    if virial:
        # |U| = 3/5 for a uniform unit sphere with G*M = 1, so sigma_1d^2 = |U| / 3
        vel = rng.normal(0.0, np.sqrt(0.6 / 3.0), (n, 3))
    else:
        vel = np.zeros((n, 3))
    pos -= pos.mean(axis=0)
    vel -= vel.mean(axis=0)
    masses = np.full(n, 1.0 / (G * n))
    return np.ascontiguousarray(pos), np.ascontiguousarray(vel), masses
