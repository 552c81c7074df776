"""The reference's own data-generation script, UNCHANGED, driven against this repo's `hpc` package (authoring
container only: /root/reference is not on the GPU box).  CUDA is replaced by the oracle-backed stand-in and
h5py by tests/fake_h5py.py, so this checks the drop-in boundary -- imports, constructor/attribute protocol,
state dicts, checkpoint files, dataset builder -- not the arithmetic."""
import os
import runpy
import sys
from pathlib import Path

import numpy as np
import pytest

import fake_h5py
from fake_engine import FakeEngine

SCRIPT = Path("/root/reference/scripts/generate_data.py")


@pytest.mark.skipif(not SCRIPT.exists(), reason="reference tree not mounted")
def test_generate_data_script_runs_unchanged(monkeypatch, tmp_path, oracle_mod, capsys):
    pytest.importorskip("tqdm")
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    import hpc                      # this repo's package is now the cached `hpc`; the script's own
    import hpc.checkpoint           # sys.path.insert(0, <reference>/src) cannot shadow it
    from hpc import ics, nbody
    assert "nbody-gnn-hpc_b200" in hpc.__file__
    fake = FakeEngine()
    monkeypatch.setattr(hpc._cuda, "get_engine", lambda device=None: fake)    # the seam lives in the tests
    # the script prepends <reference>/src to sys.path and pins thread-count env vars: undo both afterwards
    monkeypatch.setattr(sys, "path", list(sys.path))
    for var in ("OMP_NUM_THREADS", "NUMBA_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        if var in os.environ:
            monkeypatch.setenv(var, os.environ[var])
        else:
            monkeypatch.setenv(var, "1")
            monkeypatch.delenv(var)
    out = tmp_path / "data"
    monkeypatch.setattr(sys, "argv", ["generate_data.py", "--particles", "24", "--simulations", "5", "--steps", "12",
                                      "--workers", "1", "--output-dir", str(out), "--sequence-length", "5",
                                      "--batch-size", "2", "--seed", "42"])
    runpy.run_path(str(SCRIPT), run_name="__main__")
    text = capsys.readouterr().out
    assert "DATA GENERATION COMPLETE" in text and "Generated 5 trajectories" in text
    mgr = hpc.checkpoint.CheckpointManager(str(out / "checkpoints"))
    m32 = ics.shared_masses(24, 42)
    for i in range(5):
        tr = mgr.load_trajectory(f"sim_{i:04d}")
        assert tr["positions"].shape == (13, 24, 3) and tr["n_steps"] == 13
        assert np.array_equal(tr["masses"], m32) and tr["masses"].dtype == np.float32
        x0, v0, _ = ics.reference_default_ic(24, 42 + i)
        assert np.array_equal(tr["positions"][0], x0) and np.array_equal(tr["velocities"][0], v0)
        chk = oracle_mod.run(x0, v0, oracle_mod.accel_direct(x0, m32), m32, 1e-3, 1e-9, 12, 1)
        assert np.abs(tr["positions"] - chk["positions"]).max() < 1e-9
        assert np.array_equal(tr["times"], chk["times"]) and tr["metadata"]["seed"] == 42 + i
    with fake_h5py.File(out / "train_dataset.h5", "r") as f:      # 4 of 5 trajectories, 13 - 5 = 8 samples each
        assert f["inputs"].shape == (32, 5, 24, 6) and f["targets"].shape == (32, 24, 6)
    with fake_h5py.File(out / "val_dataset.h5", "r") as f:
        assert f["inputs"].shape == (8, 5, 24, 6)
