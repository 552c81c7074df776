"""K5, sliding-window training samples (reference src/hpc/checkpoint.py:362-384).

CPU: the oracle restatement against arrays produced by EXECUTING the reference's create_training_dataset
(tests/golden/windows_reference.npz, made by tests/golden/make_golden_windows.py), and the host-side
hpc.checkpoint.sliding_windows against the oracle.  GPU: the CUDA kernel against the oracle, bit for bit
(float64 -> float32 rounding is the only arithmetic), bulk-copy path and element-wise path, at the bench shape too.
"""
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden" / "windows_reference.npz"


def _oracle_all(oracle_np, pos, vel, T, L, stride):
    ins, tgs = zip(*(oracle_np.sliding_windows(pos[b], vel[b], T, L, stride) for b in range(pos.shape[0])))
    return np.concatenate(ins), np.concatenate(tgs)


@pytest.fixture(scope="module")
def oracle_np():
    from oracle import numpy_oracle
    return numpy_oracle


@pytest.mark.parametrize("tag", ["a", "c", "d"])
def test_oracle_windows_equal_reference_output(oracle_np, tag):
    g = np.load(GOLDEN)
    B, T, N, L, stride = (int(v) for v in g[f"{tag}_params"])
    ins, tgs = _oracle_all(oracle_np, g[f"{tag}_pos"], g[f"{tag}_vel"], T, L, stride)
    assert int(g[f"{tag}_n_samples"]) == ins.shape[0] == B * oracle_np.window_count(T, L, stride)
    assert ins.dtype == np.float32 and np.array_equal(ins, g[f"{tag}_inputs"])
    assert np.array_equal(tgs, g[f"{tag}_targets"])


def test_reference_rejects_strides_that_do_not_divide(oracle_np):
    """Recorded behaviour of the reference: its pre-count (n_steps - L) // stride is one per trajectory short of its
    loop when stride does not divide n_steps - L, and the write raises IndexError.  The oracle and the product count
    what the loop produces."""
    g = np.load(GOLDEN)
    assert "IndexError" in str(g["b_error"])
    B, T, N, L, stride = (int(v) for v in g["b_params"])
    assert oracle_np.window_count(T, L, stride) == (T - L) // stride + 1


@pytest.mark.parametrize("shape", [(2, 23, 6, 4, 1), (1, 30, 5, 10, 3), (2, 12, 4, 11, 1), (1, 8, 3, 8, 1)])
def test_host_sliding_windows_equal_oracle(oracle_np, shape):
    from hpc.checkpoint import sliding_windows
    B, T, N, L, stride = shape
    rng = np.random.RandomState(sum(shape))
    pos, vel = rng.standard_normal((B, T, N, 3)) * 50, rng.standard_normal((B, T, N, 3))
    for b in range(B):
        ins, tgs = sliding_windows(pos[b], vel[b], T, L, stride)
        o_in, o_tg = oracle_np.sliding_windows(pos[b], vel[b], T, L, stride)
        assert np.array_equal(np.asarray(ins), o_in) and np.array_equal(np.asarray(tgs), o_tg)


def test_window_count_entry_point_without_gpu(oracle_np):
    from hpc import _cuda
    lib = _cuda.load_library()
    for T, L, stride in ((401, 10, 1), (23, 4, 3), (24, 4, 5), (10, 10, 1), (5, 10, 1), (11, 10, 7)):
        assert lib.nb_window_count(T, L, stride) == oracle_np.window_count(T, L, stride)


# ---- GPU ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [
    (3, 23, 6, 4, 1),       # golden-sized, even N: bulk-copy path
    (2, 30, 5, 10, 3),      # odd N: element-wise path, stride that does not divide
    (4, 41, 200, 10, 1),    # the data-generation body count, several tiles per trajectory
    (2, 64, 1024, 10, 2),   # large states: few fit in shared memory
    (1, 12, 4000, 10, 1),   # a state larger than the tile budget: element-wise path
    (2, 11, 8, 10, 1),      # exactly one sample per trajectory
    (2, 10, 8, 10, 1),      # no sample at all
])
def test_window_gather_equals_oracle(engine, oracle_np, shape):
    import torch
    B, T, N, L, stride = shape
    rng = np.random.RandomState(sum(shape))
    rows = T + 2                                              # the stacks may hold more rows than n_states
    pos, vel = rng.standard_normal((B, rows, N, 3)) * 1e3, rng.standard_normal((B, rows, N, 3))
    ins, tgs = engine.window_gather(engine.to_device(pos), engine.to_device(vel), T, L, stride)
    torch.cuda.synchronize()
    o_in, o_tg = _oracle_all(oracle_np, pos, vel, T, L, stride)
    assert tuple(ins.shape) == o_in.shape and tuple(tgs.shape) == o_tg.shape
    assert np.array_equal(ins.cpu().numpy(), o_in)
    assert np.array_equal(tgs.cpu().numpy(), o_tg)


@pytest.mark.gpu
def test_device_resident_datagen_to_windows(engine, oracle_np):
    """simulate_ensemble(outputs='device') -> window_gather: the trajectories never leave HBM; equal to the host path."""
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    x0, v0, m32 = ics.datagen_ensemble_ic(6, 200, seed=5)
    host = simulate_ensemble(x0, v0, m32, dt=1e-3, n_steps=30, save_interval=1)
    dev = simulate_ensemble(x0, v0, m32, dt=1e-3, n_steps=30, save_interval=1, outputs="device")
    assert dev["positions"].is_cuda and np.array_equal(dev["positions"].cpu().numpy(), host["positions"])
    assert np.array_equal(dev["velocities"].cpu().numpy(), host["velocities"])
    ins, tgs = engine.window_gather(dev["positions"], dev["velocities"], 31, 10, 1)
    o_in, o_tg = _oracle_all(oracle_np, host["positions"], host["velocities"], 31, 10, 1)
    assert np.array_equal(ins.cpu().numpy(), o_in) and np.array_equal(tgs.cpu().numpy(), o_tg)
