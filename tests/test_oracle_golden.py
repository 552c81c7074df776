"""The CPU oracle against vectors produced by the reference itself (tests/golden/make_golden.py ran
the reference's Numba functions).  This is what pins the oracle; the GPU tests then hold the CUDA
path to the oracle.  Also: first-principles known answers and, when /root/reference is mounted
(authoring container only), a live comparison."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from conftest import rel_rows
from hpc import ics

REF = Path("/root/reference/src/hpc/nbody.py")


@pytest.mark.parametrize("mode,tol", [("fast", 1e-14), ("fast_serial", 1e-14), ("strict", 1e-14),
                                      ("strict_reversed", 1e-14)])
def test_accel_modes_vs_reference_vectors(oracle_mod, golden, mode, tol):
    g = golden("accel_default_n200.npz")
    m32 = ics.shared_masses(200, 42)
    for seed in (42, 43, 9999):
        x, _, m64 = ics.reference_default_ic(200, seed)
        assert rel_rows(oracle_mod.accel_direct(x, m32, mode=mode), g[f"acc_f32mass_seed{seed}"]).max() < tol
        assert rel_rows(oracle_mod.accel_direct(x, m64, mode=mode), g[f"acc_ctor_f64mass_seed{seed}"]).max() < tol


def test_float32_masses_are_promoted_not_rerounded(oracle_mod):
    x, _, m64 = ics.reference_default_ic(64, 1)
    m32 = m64.astype(np.float32)
    a32 = oracle_mod.accel_direct(x, m32)
    assert np.array_equal(a32, oracle_mod.accel_direct(x, m32.astype(np.float64)))
    assert not np.array_equal(a32, oracle_mod.accel_direct(x, m64))


def test_numpy_restatement_agrees_with_c(oracle_mod, golden):
    g = golden("accel_plummer_n2048_rows.npz")
    x, v, m = ics.plummer_ic(2048, seed=7)
    rows = g["rows"]
    a_np = oracle_mod.accel_rows_numpy(x, m, rows, float(g["softening"]))
    a_c = oracle_mod.accel_direct(x, m, float(g["softening"]))
    assert rel_rows(a_np, g["acc_rows"]).max() < 1e-13
    assert rel_rows(a_c[rows], g["acc_rows"]).max() < 1e-13
    assert np.array_equal(oracle_mod.accel_direct_rows(x, m, 512, 100, 0.01), a_c[512:612])
    e = oracle_mod.total_energy(x, v, m, 0.01)
    assert np.allclose(e, g["energy"], rtol=1e-12)
    assert np.allclose(oracle_mod.total_energy(x, v, m, 0.01, parallel=True), g["energy"], rtol=1e-12)


def test_energy_default_ics(oracle_mod, golden):
    g = golden("accel_default_n200.npz")
    for seed in (42, 43, 9999):
        x, v, _ = ics.reference_default_ic(200, seed)
        e = oracle_mod.total_energy(x, v, ics.shared_masses(200, 42))
        assert np.allclose(e, g[f"energy_f32mass_seed{seed}"], rtol=1e-13)


@pytest.mark.parametrize("name", ["plummer_n200", "sphere_n256", "plummer_n1024"])
def test_run_well_conditioned_vs_reference(oracle_mod, golden, name):
    """400 steps on well-conditioned systems: the oracle tracks the reference to ~1e-15."""
    g = golden(f"traj_{name}.npz")
    x, v, m = g["x0"], g["v0"], g["masses"]
    a0 = oracle_mod.accel_direct(x, m, float(g["softening"]))
    out = oracle_mod.run(x, v, a0, m, float(g["dt"]), float(g["softening"]), 400, 1)
    keep = g["steps_kept"]
    assert np.abs(out["positions"][keep] - g["positions"]).max() < 1e-12
    assert np.abs(out["velocities"][keep] - g["velocities"]).max() < 1e-12
    assert rel_rows(out["accelerations"][keep], g["accelerations"]).max() < 1e-11
    assert np.array_equal(out["times"], g["times"]) and np.array_equal(out["steps"], g["steps"])
    e1 = oracle_mod.total_energy(out["final_positions"], out["final_velocities"], m, float(g["softening"]))
    assert np.allclose(e1, g["energy1"], rtol=1e-11)


def test_run_default_ics_chaos_envelope(oracle_mod, golden):
    """Reference-default ICs: identical for the first steps, then diverging at the rate two summation
    orders of the reference itself diverge -- the oracle is inside that envelope."""
    g = golden("traj_default_n200_seed42.npz")
    x, v, _ = ics.reference_default_ic(200, 42)
    m32 = ics.shared_masses(200, 42)
    out = oracle_mod.run(x, v, oracle_mod.accel_direct(x, m32), m32, 1e-3, 1e-9, 400, 1)
    d = {int(k): np.abs(out["positions"][k] - g["positions"][r]).max() for r, k in enumerate(g["steps_kept"])}
    assert d[0] == 0 and d[1] < 1e-15 and d[10] < 1e-13 and d[50] < 1e-10 and d[100] < 1e-6
    rev = oracle_mod.run(x, v, oracle_mod.accel_direct(x, m32, mode="strict_reversed"), m32, 1e-3, 1e-9, 400, 1,
                         mode="strict_reversed")
    env = np.abs(rev["positions"][400] - g["positions"][-1]).max()
    assert d[400] < 1e3 * max(env, 1e-3)        # same order of magnitude as the reorder noise
    assert np.array_equal(out["times"], g["times"])
    assert out["final_time"] == float(g["final_time"]) and out["final_step"] == int(g["final_step"])


def test_bookkeeping_save_interval(oracle_mod, golden):
    g = golden("bookkeeping_n16.npz")
    out = oracle_mod.run(g["positions0"], g["velocities0"], g["accelerations0"], g["masses"], 1e-3, 1e-9, 50, 7)
    assert np.array_equal(out["steps"], g["steps"]) and np.array_equal(out["times"], g["times"])
    assert np.abs(out["positions"] - g["positions"]).max() < 1e-10
    assert out["final_step"] == 50 and out["final_time"] == float(g["final_time"])


def test_ensemble_vs_reference(oracle_mod, golden):
    g = golden("ensemble_default_b4_n200_t20.npz")
    x0, v0, m32 = ics.datagen_ensemble_ic(4, 200, seed=42)
    out = oracle_mod.ensemble_run(x0, v0, m32, 1e-3, 1e-9, 20, 1)
    assert np.abs(out["positions"] - g["positions"]).max() < 1e-10
    assert np.abs(out["velocities"] - g["velocities"]).max() < 1e-9
    assert np.abs(out["final_positions"] - g["positions"][:, -1]).max() < 1e-10


def test_known_answers(oracle_mod, golden):
    G = oracle_mod.G
    x = np.array([[0.0, 0.0, 0.0], [3.0, 4.0, 0.0]])
    m = np.array([2.0e10, 5.0e10])
    a = oracle_mod.accel_direct(x, m, 0.5)
    r3 = (25.0 + 0.25) ** 1.5
    assert np.allclose(a[0], G * m[1] * x[1] / r3, rtol=1e-14)
    assert np.allclose(a[1], -G * m[0] * x[1] / r3, rtol=1e-14)
    xs, _, ms = ics.plummer_ic(512, seed=1)
    acc = oracle_mod.accel_direct(xs, ms, 0.01)
    assert np.abs((ms[:, None] * acc).sum(axis=0)).max() < 1e-12 * np.abs(ms[:, None] * acc).sum()
    sol = golden("solar_system.npz")
    a = oracle_mod.accel_direct(sol["positions"], sol["masses"], float(sol["softening"]))
    assert rel_rows(a, sol["accelerations"]).max() < 1e-13
    assert 5.8e-3 < np.linalg.norm(a[3]) < 6.0e-3         # Earth's acceleration towards the Sun, m/s^2


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted (GPU box)")
def test_live_against_reference_numba(oracle_mod):
    pytest.importorskip("numba")
    spec = importlib.util.spec_from_file_location("ref_nbody", str(REF))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for n, seed in ((37, 3), (200, 7), (777, 11)):
        x, v, m = ics.reference_default_ic(n, seed)
        assert rel_rows(oracle_mod.accel_direct(x, m), ref.compute_accelerations_direct(x, m, 1e-9)).max() < 1e-13
        e = ref.compute_total_energy(x, v, m, 1e-9)
        assert np.allclose(oracle_mod.total_energy(x, v, m), e, rtol=1e-12)
    sim = ref.NBodySimulator(n_particles=50, box_size=10.0, dt=1e-3, seed=9)
    x, v, m, a = sim.positions.copy(), sim.velocities.copy(), sim.masses.copy(), sim.accelerations.copy()
    states = sim.run(15, save_interval=4, verbose=False)
    out = oracle_mod.run(x, v, a, m, 1e-3, 1e-9, 15, 4)
    assert np.abs(out["positions"] - np.stack([s["positions"] for s in states])).max() < 1e-12
    assert [s["step"] for s in states] == list(out["steps"])
