#!/usr/bin/env python3
"""Generate tests/golden/windows_*.npz by EXECUTING the reference's own create_training_dataset.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden_windows.py

The reference module src/hpc/checkpoint.py imports h5py, which this image lacks; tests/fake_h5py.py (an in-memory
stand-in that records every dataset write) is injected as `h5py` so that the reference function runs unmodified and
the arrays it writes can be read back.  Nothing is copied from the reference: the fixture holds inputs we generate
(seeded random trajectories) and the `inputs` / `targets` datasets its function produced.
"""
from __future__ import annotations

import importlib.util
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import fake_h5py  # noqa: E402

REF = Path("/root/reference/src/hpc/checkpoint.py")


def load_reference():
    sys.modules["h5py"] = fake_h5py
    spec = importlib.util.spec_from_file_location("ref_checkpoint", str(REF))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    rng = np.random.RandomState(2024)
    out = {}
    # (n_trajectories, stored states, bodies, sequence_length, stride); in "b" stride 3 does not divide 23 - 4 and
    # the reference raises IndexError (its pre-count is one sample per trajectory short of its loop)
    cases = {"a": (3, 23, 6, 4, 1), "b": (2, 23, 5, 4, 3), "c": (2, 41, 8, 10, 1), "d": (2, 24, 5, 4, 5)}
    for tag, (B, T, N, L, stride) in cases.items():
        pos = rng.standard_normal((B, T, N, 3)) * 1e3
        vel = rng.standard_normal((B, T, N, 3)) * 1e-2
        trajs = [{"positions": pos[b], "velocities": vel[b], "n_steps": T} for b in range(B)]
        with tempfile.TemporaryDirectory() as d:
            path = Path(d) / "ds.h5"
            try:
                ref.create_training_dataset(trajs, str(path), sequence_length=L, stride=stride)
                with fake_h5py.File(path, "r") as f:
                    inputs, targets = f["inputs"][:], f["targets"][:]
                    n_samples = int(f.attrs["n_samples"])
                err = ""
            except Exception as e:  # the reference's own pre-count (:333) can be short of its loop (:365)
                inputs = targets = np.zeros(0, np.float32)
                n_samples, err = -1, repr(e)
        out.update({f"{tag}_pos": pos, f"{tag}_vel": vel, f"{tag}_params": np.array([B, T, N, L, stride]),
                    f"{tag}_inputs": inputs, f"{tag}_targets": targets, f"{tag}_n_samples": np.array(n_samples),
                    f"{tag}_error": np.array(err)})
        print(tag, inputs.shape, targets.shape, n_samples, err)
    np.savez_compressed(HERE / "windows_reference.npz", **out)


if __name__ == "__main__":
    main()
