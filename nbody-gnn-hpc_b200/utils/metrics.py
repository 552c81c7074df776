"""``utils.metrics`` import path of the reference (src/utils/metrics.py:62-137) -> the GPU implementations."""
from hpc.metrics import compute_energy_error, compute_momentum_error, snapshot_energies  # noqa: F401

__all__ = ["compute_energy_error", "compute_momentum_error", "snapshot_energies"]
