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


def window_count(n_states: int, sequence_length: int, stride: int = 1) -> int:
    """Samples one trajectory yields: len(range(0, n_steps - L, stride)), checkpoint.py:365 (the pre-count at :333,
    (n_steps - L) // stride, is one short when stride does not divide n_steps - L; the loop is what writes)."""
    return len(range(0, n_states - sequence_length, stride))


def sliding_windows(positions, velocities, n_states: int, sequence_length: int, stride: int = 1):
    """The sample loop of create_training_dataset, /root/reference/src/hpc/checkpoint.py:362-384, for one trajectory:
    inputs (S, L, N, 6) float32, targets (S, N, 6) float32.  Plain loop, as in the reference."""
    positions = np.asarray(positions)
    velocities = np.asarray(velocities)
    n = positions.shape[1]
    ins, tgs = [], []
    for i in range(0, n_states - sequence_length, stride):                                  # :365
        ins.append(np.concatenate([positions[i:i + sequence_length],
                                   velocities[i:i + sequence_length]], axis=-1).astype(np.float32))   # :367-370
        tgs.append(np.concatenate([positions[i + sequence_length],
                                   velocities[i + sequence_length]], axis=-1).astype(np.float32))     # :373-376
    if not ins:
        return np.zeros((0, sequence_length, n, 6), np.float32), np.zeros((0, n, 6), np.float32)
    return np.stack(ins), np.stack(tgs)


def snapshot_energies(positions, velocities, masses, G: float = 6.67430e-11, softening: float = SOFTENING):
    """compute_energy_error / compute_momentum_error of the reference, src/utils/metrics.py:62-137, restated per
    snapshot without the N x N x 3 temporary: for every stored step t
        K_t = 1/2 sum_i m_i |v_i|^2                                   (:86)
        U_t = -1/2 G sum_{i != j} m_i m_j / sqrt(|x_i - x_j|^2 + eps^2)   (:90-102)
        p_t = sum_i m_i v_i                                           (:131)
    positions, velocities (S, N, 3); masses (N,).  Returns (K (S,), U (S,), p (S, 3))."""
    pos = np.asarray(positions, dtype=np.float64)
    vel = np.asarray(velocities, dtype=np.float64)
    m = np.asarray(masses)          # dtype kept: np.outer(masses, masses) (:82) rounds the products of float32 masses
    S, N = pos.shape[0], pos.shape[1]                                                     # to float32
    K, U, P = np.zeros(S), np.zeros(S), np.zeros((S, 3))
    for t in range(S):
        K[t] = 0.5 * np.sum(m * np.sum(vel[t] ** 2, axis=1))
        P[t] = np.sum(m[:, None] * vel[t], axis=0)
        u = 0.0
        for i in range(N):
            d = pos[t] - pos[t, i]
            inv_r = 1.0 / np.sqrt(np.sum(d * d, axis=1) + softening ** 2)
            inv_r[i] = 0.0                                            # np.fill_diagonal(inv_r, 0), :98
            u += np.sum((m[i] * m) * inv_r)                           # row i of m_matrix * inv_r, :102
        U[t] = -0.5 * G * u
    return K, U, P
