"""HPC module -- B200-native drop-in for the reference's ``hpc`` package (reference src/hpc/__init__.py).

Put this package's parent directory (``nbody-gnn-hpc_b200/``) on ``sys.path`` where the reference
puts its ``src`` (scripts/generate_data.py:26) and ``from hpc.nbody import NBodySimulator`` resolves
to the CUDA engine.  ``CheckpointManager`` is imported lazily because it needs h5py.
"""
from .nbody import NBodySimulator  # noqa: F401

__all__ = ["NBodySimulator", "CheckpointManager"]


def __getattr__(name):
    if name == "CheckpointManager":
        from .checkpoint import CheckpointManager
        return CheckpointManager
    if name == "BarnesHutTree":
        raise AttributeError(
            "BarnesHutTree is outside this engine's scope: the B200 kernels evaluate the exact direct sum "
            "at every N (NBodySimulator accepts use_barnes_hut, warns once and runs the direct sum)")
    raise AttributeError(name)
