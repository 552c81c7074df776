"""hpc.checkpoint (the snapshot sink): same API and file layout as the reference's
checkpoint.py, written from stacked trajectories.  h5py is absent from this image, so HDF5 paths run
against tests/fake_h5py.py; the npz path runs for real.  The sliding-window dataset is compared with a
per-sample restatement of the reference's loop (checkpoint.py:365-384)."""
import sys

import numpy as np
import pytest

import fake_h5py


@pytest.fixture()
def ckpt(monkeypatch, tmp_path):
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    from hpc import checkpoint
    return checkpoint, tmp_path


def _states(T=9, N=5, seed=0):
    rng = np.random.RandomState(seed)
    masses = rng.uniform(1e10, 1e12, N).astype(np.float32)
    t, out = 0.0, []
    for k in range(T):
        out.append({"positions": rng.rand(N, 3), "velocities": rng.rand(N, 3), "accelerations": rng.rand(N, 3),
                    "masses": masses, "time": t, "step": k})
        t += 0.001
    return out


def test_trajectory_roundtrip_and_layout(ckpt):
    checkpoint, tmp = ckpt
    mgr = checkpoint.CheckpointManager(str(tmp / "ck"))
    states = _states()
    assert not mgr.trajectory_exists("sim_0003")
    path = mgr.save_trajectory(states, "sim_0003", metadata={"n_particles": 5, "seed": 45, "tags": [1, 2]})
    assert path.endswith("sim_0003_trajectory.h5") and mgr.trajectory_exists("sim_0003")
    with fake_h5py.File(path, "r") as f:
        assert sorted(f.keys()) == ["accelerations", "masses", "metadata", "positions", "steps", "times", "velocities"]
        assert f["positions"].shape == (9, 5, 3) and f["positions"].dtype == np.float64
        assert f["positions"].kw.get("compression") == "gzip" and f["positions"].writes == 1   # one assignment
        assert f["masses"].dtype == np.float32 and f.attrs["n_steps"] == 9 and "created_at" in f.attrs
    tr = mgr.load_trajectory("sim_0003")
    assert np.array_equal(tr["positions"], np.stack([s["positions"] for s in states]))
    assert np.array_equal(tr["times"], np.array([s["time"] for s in states])) and list(tr["steps"]) == list(range(9))
    assert tr["metadata"] == {"n_particles": 5, "seed": 45, "tags": [1, 2]} and tr["n_steps"] == 9
    # the array entry point writes the same file
    mgr.save_trajectory_arrays("b", tr["positions"], tr["velocities"], tr["accelerations"], tr["masses"],
                               times=tr["times"], steps=tr["steps"])
    tb = mgr.load_trajectory("b")
    assert all(np.array_equal(tb[k], tr[k]) for k in ("positions", "velocities", "accelerations", "times", "masses"))
    assert mgr.list_checkpoints() == ["b (trajectory)", "sim_0003 (trajectory)"]
    assert mgr.delete_checkpoint("b") and not mgr.trajectory_exists("b") and not mgr.delete_checkpoint("b")
    with pytest.raises(FileNotFoundError):
        mgr.load_trajectory("nope")


@pytest.mark.parametrize("fmt", ["npz", "hdf5"])
def test_state_roundtrip(ckpt, fmt):
    checkpoint, tmp = ckpt
    mgr = checkpoint.CheckpointManager(str(tmp / fmt), format=fmt)
    st = _states(1)[0]
    st["time"], st["step"] = 0.125, 7
    mgr.save_state(st, "s", metadata={"note": "x", "cfg": {"a": 1}})
    back = mgr.load_state("s")
    assert np.array_equal(back["positions"], st["positions"]) and back["masses"].dtype == np.float32
    assert back["time"] == 0.125 and back["step"] == 7 and back["metadata"]["cfg"] == {"a": 1}
    with pytest.raises(FileNotFoundError):
        mgr.load_state("missing")


def test_training_dataset_equals_per_sample_loop(ckpt):
    checkpoint, tmp = ckpt
    L = 4
    trajs = []
    for seed, T in ((1, 12), (2, 4), (3, 7)):        # the second is too short: contributes no sample
        st = _states(T, 6, seed)
        trajs.append({"positions": np.stack([s["positions"] for s in st]),
                      "velocities": np.stack([s["velocities"] for s in st]), "n_steps": T})
    masses = np.linspace(1e10, 2e10, 6)
    out = checkpoint.create_training_dataset(trajs, str(tmp / "d" / "train.h5"), sequence_length=L, stride=1,
                                             masses=masses)
    exp_in, exp_tg = [], []
    for tr in trajs:                                   # reference checkpoint.py:365-384, sample by sample
        for i in range(0, tr["n_steps"] - L, 1):
            exp_in.append(np.concatenate([tr["positions"][i:i + L], tr["velocities"][i:i + L]], axis=-1).astype(np.float32))
            exp_tg.append(np.concatenate([tr["positions"][i + L], tr["velocities"][i + L]], axis=-1).astype(np.float32))
    with fake_h5py.File(out, "r") as f:
        assert f["inputs"].shape == (11, L, 6, 6) and f["inputs"].dtype == np.float32
        assert f["targets"].shape == (11, 6, 6)
        assert np.array_equal(f["inputs"][:], np.stack(exp_in)) and np.array_equal(f["targets"][:], np.stack(exp_tg))
        assert f["inputs"].writes == 2                 # one assignment per contributing trajectory, not per sample
        assert f["inputs"].kw["compression_opts"] == 4 and f["inputs"].kw["chunks"] == (11, L, 6, 6)
        assert f.attrs["sequence_length"] == L and f.attrs["n_samples"] == 11
        assert f["masses"].dtype == np.float32
    with pytest.raises(ValueError):
        checkpoint.create_training_dataset(trajs[1:2], str(tmp / "e.h5"), sequence_length=L)


def test_package_exports_checkpoint_manager_lazily(ckpt):
    import hpc
    assert hpc.CheckpointManager is ckpt[0].CheckpointManager
    with pytest.raises(AttributeError):
        hpc.BarnesHutTree
