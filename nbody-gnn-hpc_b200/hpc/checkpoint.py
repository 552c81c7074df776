"""Checkpoint management: the sink the simulator's snapshots are handed to.

Same public surface and on-disk layout as the reference's ``hpc.checkpoint``
(/root/reference/src/hpc/checkpoint.py:19-299 ``CheckpointManager``, :302-398
``create_training_dataset``), so files written here are read by the reference's GNN loader and vice
versa:

* per-state files ``{name}.h5`` / ``{name}.npz`` (arrays as datasets, int/float entries as attributes,
  optional ``metadata`` group, ``created_at`` attribute);
* per-trajectory files ``{name}_trajectory.h5`` with float64 gzip datasets ``positions``,
  ``velocities``, ``accelerations`` of shape (n_steps, N, 3), ``times``, ``steps``, ``masses``, attribute
  ``n_steps``, optional ``metadata`` group;
* training sets with float32 gzip-4 datasets ``inputs`` (S, L, N, 6) and ``targets`` (S, N, 6), attributes
  ``sequence_length``, ``n_samples``, optional ``masses``.

What is different is how the data gets there.  The engine returns whole stacked trajectories
((T+1, N, 3) arrays), so ``save_trajectory_arrays`` writes each dataset with ONE assignment instead of
one per step, and ``create_training_dataset`` builds every trajectory's sliding windows as a strided
view and writes them with one assignment per trajectory instead of one per sample.
``save_trajectory(states, ...)`` keeps the reference's list-of-dicts signature
(scripts/generate_data.py:154-165 calls it that way) and forwards to the array path.

h5py is imported lazily: the engine itself never needs it.
"""
from __future__ import annotations

import json
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np


def _h5py():
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("hpc.checkpoint needs h5py for HDF5 files (use format='npz' for single states)") from e
    return h5py


def _write_metadata(f, metadata: Optional[Dict]) -> None:
    if metadata:
        grp = f.create_group('metadata')
        for key, value in metadata.items():
            grp.attrs[key] = value if isinstance(value, (int, float, str)) else json.dumps(value)


def _read_metadata(f) -> Dict:
    out = {}
    for key in f['metadata'].attrs.keys():
        value = f['metadata'].attrs[key]
        try:
            out[key] = json.loads(value)
        except (json.JSONDecodeError, TypeError):
            out[key] = value
    return out


def _split_state(state: Dict):
    """A state dict's entries by how the file formats store them: arrays (datasets) and plain numbers (attributes).
    Anything else (strings, nested objects) has no place in either format and is left out, as in the reference."""
    arrays = {k: v for k, v in state.items() if isinstance(v, np.ndarray)}
    numbers = {k: v for k, v in state.items() if isinstance(v, (int, float)) and not isinstance(v, bool)}
    return arrays, numbers


class _Hdf5States:
    """``{name}.h5``: one gzip dataset per array, one root attribute per number, attribute ``created_at``,
    optional group ``metadata`` (format of reference checkpoint.py:44-70,118-142)."""
    suffix = ".h5"

    @staticmethod
    def write(path: Path, state: Dict, metadata: Optional[Dict]) -> None:
        arrays, numbers = _split_state(state)
        with _h5py().File(path, 'w') as f:
            for key, arr in arrays.items():
                f.create_dataset(key, data=arr, compression='gzip')
            for key, value in numbers.items():
                f.attrs[key] = value
            _write_metadata(f, metadata)
            f.attrs['created_at'] = datetime.now().isoformat()

    @staticmethod
    def read(path: Path) -> Dict:
        with _h5py().File(path, 'r') as f:
            state = {key: f[key][:] for key in f.keys() if key != 'metadata'}
            state.update({key: f.attrs[key] for key in f.attrs.keys() if key != 'created_at'})
            if 'metadata' in f:
                state['metadata'] = _read_metadata(f)
        return state


class _NpzStates:
    """``{name}.npz``: arrays under their own names, numbers as 0-d arrays named ``scalar_<key>``, the metadata dict
    as one JSON string ``metadata_json`` (format of reference checkpoint.py:72-106,144-170)."""
    suffix = ".npz"
    _NUMBER, _META = "scalar_", "metadata_json"

    @classmethod
    def write(cls, path: Path, state: Dict, metadata: Optional[Dict]) -> None:
        arrays, numbers = _split_state(state)
        payload = dict(arrays)
        payload.update({cls._NUMBER + k: np.array(v) for k, v in numbers.items()})
        if metadata:
            payload[cls._META] = np.array(json.dumps(metadata))
        np.savez_compressed(path, **payload)

    @classmethod
    def read(cls, path: Path) -> Dict:
        state = {}
        with np.load(path, allow_pickle=True) as data:
            for key in data.files:
                if key == cls._META:
                    state['metadata'] = json.loads(str(data[key]))
                elif key.startswith(cls._NUMBER):
                    state[key[len(cls._NUMBER):]] = data[key].item()
                else:
                    state[key] = data[key]
        return state


class CheckpointManager:
    """Saves and loads simulation states and trajectories (public surface of reference checkpoint.py:19-299)."""

    _STATE_FORMATS = {"hdf5": _Hdf5States, "npz": _NpzStates}
    _TRAJECTORY_SUFFIX = "_trajectory.h5"

    def __init__(self, checkpoint_dir: str = "./data/checkpoints", format: str = "hdf5"):
        self.checkpoint_dir = Path(checkpoint_dir)
        self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self.format = format

    # ---- single states (not on the engine's path; kept because the class is public API) ----------
    def save_state(self, state: Dict, name: str, metadata: Optional[Dict] = None) -> str:
        """Save one state dict in the manager's format; returns the file path."""
        codec = self._STATE_FORMATS.get(self.format, _NpzStates)       # anything but "hdf5" means npz, as upstream
        path = self.checkpoint_dir / (name + codec.suffix)
        codec.write(path, state, metadata)
        return str(path)

    def load_state(self, name: str) -> Dict:
        """Load a state saved by save_state, whichever format it is in (HDF5 looked for first)."""
        for codec in (_Hdf5States, _NpzStates):
            path = self.checkpoint_dir / (name + codec.suffix)
            if path.exists():
                return codec.read(path)
        raise FileNotFoundError(f"Checkpoint '{name}' not found")

    # ---- trajectories ----------------------------------------------------------------------------
    def save_trajectory_arrays(self, name: str, positions: np.ndarray, velocities: np.ndarray,
                               accelerations: np.ndarray, masses: np.ndarray, times=None, steps=None,
                               metadata: Optional[Dict] = None) -> str:
        """Write a whole stacked trajectory ((n_steps, N, 3) arrays) -- one assignment per dataset.

        Produces exactly the file save_trajectory produces for the equivalent list of states.
        """
        filepath = self.checkpoint_dir / f"{name}_trajectory.h5"
        n_steps = positions.shape[0]
        with _h5py().File(filepath, 'w') as f:
            f.attrs['n_steps'] = n_steps
            for key, arr in (('positions', positions), ('velocities', velocities), ('accelerations', accelerations)):
                ds = f.create_dataset(key, shape=tuple(arr.shape), dtype='float64', compression='gzip')
                ds[...] = arr
            f.create_dataset('times', data=np.arange(n_steps) if times is None else np.asarray(times))
            f.create_dataset('steps', data=np.arange(n_steps) if steps is None else np.asarray(steps))
            f.create_dataset('masses', data=masses)
            _write_metadata(f, metadata)
            f.attrs['created_at'] = datetime.now().isoformat()
        return str(filepath)

    def save_trajectory(self, states: List[Dict], name: str, metadata: Optional[Dict] = None) -> str:
        """Save a list of state dicts as one trajectory file (reference :172-236)."""
        return self.save_trajectory_arrays(
            name,
            np.stack([s['positions'] for s in states]),
            np.stack([s['velocities'] for s in states]),
            np.stack([s['accelerations'] for s in states]),
            states[0]['masses'],
            times=np.array([s.get('time', i) for i, s in enumerate(states)]),
            steps=np.array([s.get('step', i) for i, s in enumerate(states)]),
            metadata=metadata)

    def load_trajectory(self, name: str) -> Dict:
        """Load a trajectory file (reference :238-273)."""
        filepath = self.checkpoint_dir / f"{name}_trajectory.h5"
        if not filepath.exists():
            raise FileNotFoundError(f"Trajectory '{name}' not found")
        with _h5py().File(filepath, 'r') as f:
            trajectory = {key: f[key][:] for key in ('positions', 'velocities', 'accelerations', 'times', 'steps',
                                                     'masses')}
            trajectory['n_steps'] = f.attrs['n_steps']
            if 'metadata' in f:
                trajectory['metadata'] = _read_metadata(f)
        return trajectory

    def list_checkpoints(self) -> List[str]:
        """Sorted names of everything stored here; trajectory files are listed as ``<name> (trajectory)``."""
        tag = self._TRAJECTORY_SUFFIX[:-len(".h5")]
        stored = [p.stem for p in self.checkpoint_dir.iterdir() if p.suffix in (".h5", ".npz")]
        return sorted(stem.replace(tag, " (trajectory)") for stem in stored)

    def trajectory_exists(self, name: str) -> bool:
        """Resume check used by generate_data.py:128 (reference :286-289)."""
        return (self.checkpoint_dir / (name + self._TRAJECTORY_SUFFIX)).exists()

    def delete_checkpoint(self, name: str) -> bool:
        """Remove ONE file stored under this name -- a state (.h5, then .npz) before a trajectory -- and say whether
        there was one."""
        candidates = [self.checkpoint_dir / (name + sfx) for sfx in (".h5", ".npz", self._TRAJECTORY_SUFFIX)]
        victim = next((p for p in candidates if p.exists()), None)
        if victim is None:
            return False
        victim.unlink()
        return True


def sliding_windows(positions: np.ndarray, velocities: np.ndarray, n_steps: int, sequence_length: int,
                    stride: int = 1):
    """(inputs, targets) of one trajectory as float32 arrays: inputs (S, L, N, 6), targets (S, N, 6).

    Sample s starts at step i = s*stride: inputs[s] = states i .. i+L-1, targets[s] = state i+L, state =
    [x, y, z, vx, vy, vz] per particle (reference checkpoint.py:365-384).  The windows are taken from one
    (T, N, 6) float32 array through a strided view, so nothing is copied until the caller writes them out.
    """
    L = sequence_length
    starts = np.arange(0, n_steps - L, stride)
    state = np.concatenate([positions[:n_steps], velocities[:n_steps]], axis=-1).astype(np.float32)
    if starts.size == 0:
        n = state.shape[1]
        return np.zeros((0, L, n, 6), np.float32), np.zeros((0, n, 6), np.float32)
    windows = np.lib.stride_tricks.sliding_window_view(state, L, axis=0)       # (T-L+1, N, 6, L)
    inputs = np.moveaxis(windows, -1, 1)[starts]                                # (S, L, N, 6) view + gather
    targets = state[starts + L]
    return inputs, targets


def sliding_windows_device(positions, velocities, n_steps: int, sequence_length: int, stride: int = 1, device=None):
    """The sample loop of the reference (checkpoint.py:362-384) for B trajectories whose float64 stacks
    (B, rows, N, 3) are torch CUDA tensors -- e.g. ``simulate_ensemble(..., outputs="device")``: returns
    (inputs (B*S, L, N, 6), targets (B*S, N, 6)) float32 CUDA tensors, trajectory-major, i.e. the 'inputs' and
    'targets' datasets of the reference's dataset file.  One launch of the window kernel (K5, csrc/nb_windows.cu)."""
    from . import _cuda
    eng = _cuda.get_engine(device if device is not None else positions.device.index)
    return eng.window_gather(positions, velocities, int(n_steps), int(sequence_length), int(stride))


def create_training_dataset(trajectories: List[Dict], output_path: str, sequence_length: int = 10, stride: int = 1,
                            masses: Optional[np.ndarray] = None) -> str:
    """(input sequence, next state) pairs of all trajectories in one HDF5 file (reference :302-398)."""
    counts = []
    sample_n = None
    for traj in trajectories:
        n = len(range(0, traj['n_steps'] - sequence_length, stride))
        counts.append(n)
        if sample_n is None and n > 0:
            sample_n = traj['positions'].shape[1]
    total = int(sum(counts))
    if total == 0:
        raise ValueError("No samples could be created from trajectories")
    shape_in = (sequence_length, sample_n, 6)
    shape_tg = (sample_n, 6)
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    with _h5py().File(output_path, 'w') as f:
        inputs_ds = f.create_dataset('inputs', shape=(total,) + shape_in, dtype='float32', compression='gzip',
                                     compression_opts=4, chunks=(min(100, total),) + shape_in)
        targets_ds = f.create_dataset('targets', shape=(total,) + shape_tg, dtype='float32', compression='gzip',
                                      compression_opts=4, chunks=(min(100, total),) + shape_tg)
        at = 0
        for traj, n in zip(trajectories, counts):
            if n == 0:
                continue
            inputs, targets = sliding_windows(traj['positions'], traj['velocities'], traj['n_steps'],
                                              sequence_length, stride)
            inputs_ds[at:at + n] = inputs
            targets_ds[at:at + n] = targets
            at += n
        f.attrs['sequence_length'] = sequence_length
        f.attrs['n_samples'] = total
        f.attrs['created_at'] = datetime.now().isoformat()
        if masses is not None:
            f.create_dataset('masses', data=np.asarray(masses).astype(np.float32))
    print(f"Created dataset with {total} samples at {output_path}")
    return str(output_path)
