"""NumPy float64 restatement of the reference force loop and leapfrog step -- TEST INFRASTRUCTURE.

Used where the C oracle is not wanted: spot rows of very large systems
 and as an independent second opinion on the C port in
tests/test_oracle_golden.py.  Follows /root/reference/src/hpc/nbody.py:41-64
(force) and :202-218 (step) operation for operation; only the loop over j is
vectorised, so sums are taken in NumPy's pairwise order rather than ascending j.
"""
from __future__ import annotations

import numpy as np

G = 6.67430e-11        # nbody.py:18
SOFTENING = 1e-9       # nbody.py:19


def accel_rows_numpy(positions, masses, rows, softening: float = SOFTENING) -> np.ndarray:
    """Accelerations of the listed rows only: nbody.py:41-64 with the j loop as array ops."""
    pos = np.asarray(positions, dtype=np.float64)
    m = np.asarray(masses).astype(np.float64)          # float32 masses are promoted, never re-rounded
    rows = np.atleast_1d(np.asarray(rows, dtype=np.int64))
    out = np.zeros((rows.size, 3))
    eps2 = softening * softening
    for k, i in enumerate(rows):
        d = pos - pos[i]                               # :47-49
        r2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2] + eps2   # :52
        r = np.sqrt(r2)                                # :53
        r3 = r * r2                                    # :54
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = G * m / r3                        # :57
        factor[i] = 0.0                                # :46  (i != j)
        out[k] = (factor[:, None] * d).sum(axis=0)     # :58-60
    return out


def step_numpy(positions, velocities, accelerations, masses, dt: float, softening: float = SOFTENING):
    """One NBodySimulator.step(), nbody.py:202-218; returns new (x, v, a)."""
    x = np.array(positions, dtype=np.float64)
    v = np.array(velocities, dtype=np.float64)
    a = np.array(accelerations, dtype=np.float64)
    v += 0.5 * dt * a                                  # :205
    x += dt * v                                        # :208
    a = accel_rows_numpy(x, masses, np.arange(x.shape[0]), softening)   # :211
    v += 0.5 * dt * a                                  # :214
    return x, v, a
