"""Energy and momentum of stored trajectories -- the conservation metrics of the reference's ``utils.metrics``
(reference src/utils/metrics.py:62-137), evaluated on the GPU for every snapshot of a trajectory or of a whole
ensemble in one launch (K4b, ``csrc/nb_energy.cu``).

    compute_energy_error(positions, velocities, masses, G, softening)   -> (energy_per_step, max relative error)
    compute_momentum_error(velocities, masses)                          -> (|p| per step, max relative error)

are the reference's functions, same arguments and results.  ``snapshot_energies`` is the batched form: stacks of shape
(S, N, 3) or (B, S, N, 3), host ndarrays or float64 device tensors -- e.g. the stacks
``simulate_ensemble(..., outputs="device")`` leaves in HBM, so the drift curves of 300 trajectories cost one kernel
launch and 5 doubles per snapshot of PCIe traffic.  The reference evaluates the same sums with a Python loop over the
steps and an N x N x 3 temporary per step.  No CPU fallback: the arithmetic is the CUDA kernel's.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from . import _cuda

G_DEFAULT = 6.67430e-11      # reference metrics.py:65
SOFTENING_DEFAULT = 1e-9     # reference metrics.py:66


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch")


def snapshot_energies(positions, velocities, masses, G: float = G_DEFAULT, softening: float = SOFTENING_DEFAULT,
                      *, device=None) -> dict:
    """Kinetic, potential and total energy and the total momentum vector of every snapshot.

    positions, velocities: (S, N, 3) or (B, S, N, 3); masses: (N,) shared or (B, N).  positions may be None
    (momentum and kinetic energy only).  Returns host float64 arrays 'kinetic', 'potential', 'total' of shape (S,) /
    (B, S) and 'momentum' of shape (S, 3) / (B, S, 3)."""
    import torch
    if _is_tensor(velocities):
        eng = _cuda.get_engine(velocities.device.index if device is None else device)
    else:
        eng = _cuda.get_engine(device)
    with torch.cuda.device(eng.device):
        def stack(a):
            if a is None:
                return None
            t = a if _is_tensor(a) else eng.to_device(np.ascontiguousarray(a, dtype=np.float64))
            t = t.to(device=eng.device, dtype=torch.float64)
            return (t[None] if t.dim() == 3 else t).contiguous()
        vel_d = stack(velocities)
        pos_d = stack(positions)
        if vel_d.dim() != 4 or vel_d.shape[-1] != 3:
            raise ValueError(f"stacks must have shape (S, N, 3) or (B, S, N, 3), got {tuple(vel_d.shape)}")
        B, S, N = (int(d) for d in vel_d.shape[:3])
        m = np.asarray(masses.detach().cpu().numpy() if _is_tensor(masses) else masses)
        if m.shape == (N,):
            stride = 0
        elif m.shape == (B, N):
            stride = N
        else:
            raise ValueError(f"masses must have shape ({N},) or ({B}, {N}); got {m.shape}")
        m_d, f32 = eng._masses_dev(np.ascontiguousarray(m))
        out = eng.snapshot_energies(pos_d, vel_d, m_d, f32, stride, float(softening), float(G)).cpu().numpy()
    batched = np.ndim(velocities) == 4 if not _is_tensor(velocities) else velocities.dim() == 4
    if not batched:
        out = out[0]
    return {"kinetic": out[..., 0].copy(), "potential": out[..., 1].copy(), "total": out[..., 0] + out[..., 1],
            "momentum": out[..., 2:5].copy()}


def compute_energy_error(positions, velocities, masses, G: float = G_DEFAULT,
                         softening: float = SOFTENING_DEFAULT) -> Tuple[np.ndarray, float]:
    """Total energy at each stored step and the largest relative deviation from the first
    (reference metrics.py:62-109).  positions, velocities (n_steps, N, 3); masses (N,)."""
    e = snapshot_energies(positions, velocities, masses, G, softening)["total"]
    relative_error = np.abs((e - e[0]) / e[0])                      # metrics.py:107
    return e, float(np.max(relative_error))


def compute_momentum_error(velocities, masses) -> Tuple[np.ndarray, float]:
    """|total momentum| at each stored step and its largest relative change (reference metrics.py:112-137)."""
    p = snapshot_energies(None, velocities, masses)["momentum"]
    momentum_mag = np.linalg.norm(p, axis=-1)                       # metrics.py:133
    initial_mag = max(momentum_mag[0], 1e-10)                       # metrics.py:136
    relative_error = np.abs((momentum_mag - momentum_mag[0]) / initial_mag)
    return momentum_mag, float(np.max(relative_error))
