"""GPU parity, K2/K3: trajectories from the fused leapfrog kernels against the oracle and the
golden vectors produced by the reference.

Tolerances (BASELINE.json north_star): float64 positions <= 1e-8 after 400 steps -- asserted on
well-conditioned systems (Plummer / uniform sphere); the reference-default ICs are chaotic
(e-folding ~11 steps), so there the bar is applied over the first 50 steps
and the later divergence is bounded by the reference's own reorder envelope.  float32: <= 1e-5
relative per step.
"""
import numpy as np
import pytest

from conftest import rel_rows

pytestmark = pytest.mark.gpu

POS_TOL = 1e-8


def _sim_from(x, v, m, dt, eps, dtype=None):
    from hpc.nbody import NBodySimulator
    sim = NBodySimulator(n_particles=x.shape[0], box_size=10.0, dt=dt, softening=eps, seed=0, dtype=dtype)
    sim.positions, sim.velocities, sim.masses = x.copy(), v.copy(), m.copy()
    sim.accelerations = sim._compute_accelerations()
    return sim


@pytest.mark.parametrize("name", ["plummer_n200", "sphere_n256", "plummer_n1024"])
def test_run_f64_well_conditioned_vs_golden(golden, name):
    """400 steps, float64: positions within 1e-8 of the reference's own run (observed ~1e-15)."""
    g = golden(f"traj_{name}.npz")
    sim = _sim_from(g["x0"], g["v0"], g["masses"], float(g["dt"]), float(g["softening"]))
    states = sim.run(400, save_interval=1, verbose=False)
    assert len(states) == 401
    for row, k in enumerate(g["steps_kept"]):
        assert np.abs(states[k]["positions"] - g["positions"][row]).max() < POS_TOL
        assert np.abs(states[k]["velocities"] - g["velocities"][row]).max() < POS_TOL
        assert rel_rows(states[k]["accelerations"], g["accelerations"][row]).max() < 1e-10
    assert np.array_equal(np.array([s["time"] for s in states]), g["times"])
    assert [s["step"] for s in states] == list(g["steps"])
    e1 = sim.get_energy()
    assert np.allclose(e1, g["energy1"], rtol=1e-10)
    assert abs(e1[2] - g["energy0"][2]) / abs(g["energy0"][2]) < 1e-6


def test_run_f64_default_ics_datagen_sequence(golden):
    """The exact call sequence of generate_data.py:36-49 / evaluate.py:85-93 on the default ICs."""
    from hpc import ics
    from hpc.nbody import NBodySimulator
    g = golden("traj_default_n200_seed42.npz")
    sim = NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42)
    sim.masses = ics.shared_masses(200, 42).copy()
    sim.accelerations = sim._compute_accelerations()
    states = sim.run(400, save_interval=1, verbose=False)
    assert len(states) == 401 and sim.step_count == 400 and sim.time == float(g["final_time"])
    assert states[0]["masses"].dtype == np.float32          # ends up in HDF5 'masses', checkpoint.py:223
    assert states[5]["positions"].dtype == np.float64
    scale = np.abs(g["positions"][0]).max()
    for row, k in enumerate(g["steps_kept"]):
        d = np.abs(states[k]["positions"] - g["positions"][row]).max()
        if k <= 50:
            assert d < POS_TOL, (k, d)
        elif k <= 100:
            assert d < 1e-6 * scale, (k, d)      # reorder envelope at step 100: 1e-11 .. 2e-8
    # chaotic tail: statistically the same system (ranges within the envelope of RESULTS_ANALYSIS.md:33-34)
    assert np.abs(states[400]["positions"]).max() < 1e5 and np.abs(states[400]["velocities"]).max() < 1e6
    assert np.array_equal(np.array([s["time"] for s in states]), g["times"])


def test_ensemble_vs_golden_and_oracle(golden, oracle_mod):
    """K3 against 4 data-generation simulations run by the reference (first 20 steps) and against the
    oracle on a ragged batch."""
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    g = golden("ensemble_default_b4_n200_t20.npz")
    x0, v0, m32 = ics.datagen_ensemble_ic(4, 200, seed=42)
    out = simulate_ensemble(x0, v0, m32, dt=1e-3, n_steps=20, save_interval=1)
    assert out["positions"].shape == (4, 21, 200, 3)
    assert np.abs(out["positions"] - g["positions"]).max() < POS_TOL
    assert np.abs(out["velocities"] - g["velocities"]).max() < 1e-7
    assert rel_rows(out["accelerations"], g["accelerations"]).max() < 1e-9
    # per-system masses, N not a multiple of anything, save_interval 3
    rng = np.random.RandomState(3)
    B, N = 7, 77
    x0 = rng.rand(B, N, 3) * 4 - 2
    v0 = rng.rand(B, N, 3) - 0.5
    m = rng.uniform(1e9, 1e11, (B, N))
    out = simulate_ensemble(x0, v0, m, dt=2e-3, softening=0.05, n_steps=30, save_interval=3)
    assert out["positions"].shape == (B, 11, N, 3)
    for b in range(B):
        chk = oracle_mod.run(x0[b], v0[b], oracle_mod.accel_direct(x0[b], m[b], 0.05), m[b], 2e-3, 0.05, 30, 3)
        assert np.abs(out["positions"][b] - chk["positions"]).max() < POS_TOL
        assert np.abs(out["final_velocities"][b] - chk["final_velocities"]).max() < POS_TOL
    assert np.array_equal(out["times"], chk["times"])


def test_ensemble_interval_schedule_matches_small_batches(engine):
    """More systems than workers: K3 cuts the systems x steps line into one interval per worker, so systems are
    handed from one CTA lane to the next in mid-run; the result must be bit-identical to running the same
    systems in small batches (one system per CTA, no hand-over)."""
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    B = 3 * engine.sm_count * 2 + 5
    x0, v0, m32 = ics.datagen_ensemble_ic(B, 64, seed=100)
    big = simulate_ensemble(x0, v0, m32[:64], dt=1e-3, n_steps=40, save_interval=4)
    for lo in range(0, B, 97):
        small = simulate_ensemble(x0[lo:lo + 97], v0[lo:lo + 97], m32[:64], dt=1e-3, n_steps=40, save_interval=4)
        for key in ("positions", "velocities", "accelerations", "final_positions", "final_velocities"):
            assert np.array_equal(big[key][lo:lo + 97], small[key]), key


def test_k2_path_matches_k3_path(monkeypatch, golden):
    """The per-step kernels (K1/K2) and the one-launch kernel (K3) integrate the same system to the same
    trajectory within rounding (different summation order only)."""
    from hpc import _cuda
    g = golden("traj_plummer_n200.npz")
    sim = _sim_from(g["x0"], g["v0"], g["masses"], 1e-3, 0.01)
    a = sim.run(400, save_interval=50, verbose=False)
    monkeypatch.setattr(_cuda, "SMALL_SYSTEM_MAX_BODIES", 0)
    sim = _sim_from(g["x0"], g["v0"], g["masses"], 1e-3, 0.01)
    b = sim.run(400, save_interval=50, verbose=False)
    assert len(a) == len(b) == 9
    for sa, sb in zip(a, b):
        assert np.abs(sa["positions"] - sb["positions"]).max() < 1e-12
        assert np.abs(sa["velocities"] - sb["velocities"]).max() < 1e-12
    assert np.abs(b[-1]["positions"] - g["positions"][-1]).max() < POS_TOL


def test_run_f64_n4096_vs_oracle(oracle_mod):
    """K2 multi-CTA path, 100 steps at N = 4096 against the oracle's run."""
    from hpc import ics
    x, v, m = ics.plummer_ic(4096, seed=7)
    sim = _sim_from(x, v, m, 1e-3, 0.01)
    states = sim.run(100, save_interval=25, verbose=False)
    chk = oracle_mod.run(x, v, oracle_mod.accel_direct(x, m, 0.01), m, 1e-3, 0.01, 100, 25)
    assert len(states) == 5
    for r in range(5):
        assert np.abs(states[r]["positions"] - chk["positions"][r]).max() < POS_TOL
        assert np.abs(states[r]["velocities"] - chk["velocities"][r]).max() < POS_TOL


def test_run_f32_per_step_tolerance(oracle_mod):
    """float32 kernels: one step from an identical state differs from the float64 oracle by < 1e-5
    relative (positions, velocities: max-norm; accelerations: global max-norm), N = 200 and 4096."""
    from hpc import ics
    for n, (x, v, m), eps in ((200, ics.plummer_ic(200, seed=7), 0.01), (4096, ics.plummer_ic(4096, seed=7), 0.01)):
        a0 = oracle_mod.accel_direct(x, m, eps)
        chk = oracle_mod.run(x, v, a0, m, 1e-3, eps, 1, 1)
        sim = _sim_from(x, v, m, 1e-3, eps, dtype="float32")
        sim.accelerations = a0.copy()
        sim.step()
        assert np.abs(sim.positions - chk["final_positions"]).max() / np.abs(chk["final_positions"]).max() < 1e-5
        assert np.abs(sim.velocities - chk["final_velocities"]).max() / np.abs(chk["final_velocities"]).max() < 1e-5
        assert (np.abs(sim.accelerations - chk["final_accelerations"]).max()
                / np.abs(chk["final_accelerations"]).max()) < 1e-5


def test_energy_drift_f32_vs_f64():
    """Relative energy drift over 200 steps, Plummer N = 2048: both precisions stay small and close."""
    from hpc import ics
    x, v, m = ics.plummer_ic(2048, seed=7)
    drift = {}
    for dtype in ("float64", "float32"):
        sim = _sim_from(x, v, m, 1e-3, 0.01, dtype=dtype)
        e0 = sim.get_energy()[2]
        sim.run(200, save_interval=200, verbose=False)
        drift[dtype] = abs(sim.get_energy()[2] - e0) / abs(e0)
    assert drift["float64"] < 1e-6
    assert drift["float32"] < 1e-4


def test_step_and_bookkeeping(golden):
    """step(), save_interval > 1 and verbose segmentation reproduce the reference's bookkeeping."""
    from hpc.nbody import NBodySimulator
    g = golden("bookkeeping_n16.npz")
    sim = NBodySimulator(n_particles=16, box_size=10.0, dt=0.001, seed=5)
    assert np.array_equal(sim.positions, g["positions0"]) and np.array_equal(sim.masses, g["masses"])
    assert rel_rows(sim.accelerations, g["accelerations0"]).max() < 1e-10
    states = sim.run(50, save_interval=7, verbose=True)     # verbose: report every 5 steps -> segmented run
    assert [s["step"] for s in states] == list(g["steps"])
    assert np.array_equal(np.array([s["time"] for s in states]), g["times"])
    assert sim.step_count == int(g["final_step"]) and sim.time == float(g["final_time"])
    assert np.abs(np.stack([s["positions"] for s in states]) - g["positions"]).max() < POS_TOL
    assert np.abs(sim.positions - g["final_positions"]).max() < POS_TOL
    sim2 = NBodySimulator(n_particles=16, box_size=10.0, dt=0.001, seed=5)
    for _ in range(50):
        sim2.step()
    assert np.abs(sim2.positions - g["final_positions"]).max() < POS_TOL and sim2.time == sim.time
    sim2.set_state(states[2])
    assert sim2.step_count == 14 and np.array_equal(sim2.positions, states[2]["positions"])


def test_host_buffer_abi_run_and_ensemble(oracle_mod):
    """nbh_run / nbh_ensemble_run / nbh_total_energy: host pointers through the C ABI, no torch."""
    from hpc import _cuda, ics
    lib = _cuda.load_library()
    x, v, m = ics.plummer_ic(1500, seed=7)
    a0 = oracle_mod.accel_direct(x, m, 0.01)
    chk = oracle_mod.run(x, v, a0, m, 1e-3, 0.01, 20, 10)
    px, pv, pa = x.copy(), v.copy(), a0.copy()
    sp, sv, sa = (np.zeros((3, 1500, 3)) for _ in range(3))
    rc = lib.nbh_run(px.ctypes.data, pv.ctypes.data, pa.ctypes.data, m.ctypes.data, 0, 1500, 1e-3, 0.01, 20, 10, 0,
                     sp.ctypes.data, sv.ctypes.data, sa.ctypes.data)
    assert rc == 0, lib.nb_last_error()
    assert np.abs(sp - chk["positions"]).max() < POS_TOL and np.abs(px - chk["final_positions"]).max() < POS_TOL
    assert np.abs(pv - chk["final_velocities"]).max() < POS_TOL
    x0, v0, m32 = ics.datagen_ensemble_ic(3, 200, seed=42)
    ex, ev, ea = x0.copy(), v0.copy(), np.zeros_like(x0)
    ox, ov, oa = (np.zeros((3, 11, 200, 3)) for _ in range(3))
    rc = lib.nbh_ensemble_run(ex.ctypes.data, ev.ctypes.data, ea.ctypes.data, m32.ctypes.data, 1, 0, 3, 200, 1e-3,
                              1e-9, 10, 1, 0, ox.ctypes.data, ov.ctypes.data, oa.ctypes.data)
    assert rc == 0, lib.nb_last_error()
    chk = oracle_mod.ensemble_run(x0, v0, m32, 1e-3, 1e-9, 10)
    assert np.abs(ox - chk["positions"]).max() < POS_TOL
    kut = np.zeros(3)
    rc = lib.nbh_total_energy(x.ctypes.data, v.ctypes.data, m.ctypes.data, 0, 1500, 0.01, kut.ctypes.data)
    assert rc == 0 and np.allclose(kut, oracle_mod.total_energy(x, v, m, 0.01), rtol=1e-12)


def test_generate_simulations_matches_simulator_loop():
    """hpc.ensemble.generate_simulations (batched) == looping the reference's worker recipe."""
    from hpc import ics
    from hpc.ensemble import generate_simulations
    from hpc.nbody import NBodySimulator
    m32 = ics.shared_masses(200, 42)
    args = [(i, 200, 12, 1, 10.0, 42 + i, m32) for i in range(3)]
    batch = generate_simulations(args)
    for i, traj in enumerate(batch):
        sim = NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42 + i)
        sim.masses = m32.copy()
        sim.accelerations = sim._compute_accelerations()
        states = sim.run(12, save_interval=1, verbose=False)
        assert traj["n_steps"] == 13 and traj["masses"].dtype == np.float32
        # a_0 comes from K1 in the loop and from K3 itself in the batch: same values to rounding only
        assert np.abs(traj["positions"] - np.stack([s["positions"] for s in states])).max() < POS_TOL
        assert np.array_equal(traj["times"], np.array([s["time"] for s in states]))


def test_slabwise_steps_bitwise_equal_full_step(engine):
    """Two i-slabs stepped one after the other into the same position stream give exactly the bits of
    one full-system step (what makes sharded runs bit-identical to one-GPU runs), both precisions."""
    import torch
    from hpc import _cuda, ics
    n = 3000
    x, v, m = ics.plummer_ic(n, seed=7)
    for dtype, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        pos_d = engine.to_device(x)
        m_d, f32 = engine._masses_dev(m)
        results = []
        for slabs in ([(0, n)], [(0, 1504), (1504, n - 1504)]):
            cur = engine.pack(pos_d, m_d, f32, n, dtype)
            nxt = cur.clone()
            vel = engine.to_device(v, tdt)
            acc = engine.accel_slab(cur, n, 0, n, 0.01)
            ws = engine.workspace(n, n, dtype)
            for i0, n_i in slabs:
                engine.kick_drift_slab(cur, nxt, vel[i0:i0 + n_i], acc[i0:i0 + n_i], n, i0, n_i, 1e-3)
            cur, nxt = nxt, cur
            for k in range(3):
                for i0, n_i in slabs:
                    engine.step_slab(cur, nxt, vel[i0:i0 + n_i], acc[i0:i0 + n_i], n, i0, n_i, 1e-3, 0.01,
                                     _cuda.NB_STEP_CONTINUE, None, None, None, ws)
                cur, nxt = nxt, cur
            results.append((cur.clone(), vel.clone(), acc.clone()))
        for a, b in zip(*results):
            assert torch.equal(a, b)


def test_sharded_system_single_rank_matches_simulator(oracle_mod):
    from hpc import ics
    from hpc.sharded import ShardedSystem
    x, v, m = ics.plummer_ic(2500, seed=7)
    sysm = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=np.float64)
    sysm.advance(8)
    chk = oracle_mod.run(x, v, oracle_mod.accel_direct(x, m, 0.01), m, 1e-3, 0.01, 8, 8)
    assert np.abs(sysm.positions() - chk["final_positions"]).max() < POS_TOL
    assert np.abs(sysm.velocities() - chk["final_velocities"]).max() < POS_TOL
    e = oracle_mod.total_energy(chk["final_positions"], chk["final_velocities"], m, 0.01, parallel=True)
    assert np.allclose(sysm.energy(), e, rtol=1e-10)


def test_config3_n16384_both_precisions_vs_oracle(oracle_mod):
    """BASELINE config 3 (single system N = 16,384 Plummer), shortened to 40 steps so the CPU oracle stays at
    a few seconds: float64 within 1e-8 on positions, float32 within 1e-5 relative (max-norm)."""
    from hpc import ics
    x, v, m = ics.plummer_ic(16384, seed=7)
    a0 = oracle_mod.accel_direct(x, m, 0.01)
    chk = oracle_mod.run(x, v, a0, m, 1e-3, 0.01, 40, 40)
    for dtype, tol in (("float64", POS_TOL), ("float32", 1e-5 * np.abs(chk["final_positions"]).max())):
        sim = _sim_from(x, v, m, 1e-3, 0.01, dtype=dtype)
        states = sim.run(40, save_interval=40, verbose=False)
        assert len(states) == 2
        assert np.abs(states[-1]["positions"] - chk["final_positions"]).max() < tol
        assert np.abs(states[-1]["velocities"] - chk["final_velocities"]).max() < max(tol, 1e-5 * np.abs(chk["final_velocities"]).max() if dtype == "float32" else tol)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 257, 600, 1024, 1100])
def test_ensemble_body_counts(oracle_mod, n):
    """K3 thread layouts at the edges: one body, odd counts (dummy second body), one part only, the
    one-CTA-per-SM variant (N > ~500), the shared-memory limit (1024) and the per-system fallback above it."""
    from hpc.ensemble import simulate_ensemble
    rng = np.random.RandomState(n)
    B = 3
    x0 = rng.rand(B, n, 3) * 4 - 2
    v0 = rng.rand(B, n, 3) - 0.5
    m = rng.uniform(1e9, 1e10, n)
    out = simulate_ensemble(x0, v0, m, dt=1e-3, softening=0.05, n_steps=6, save_interval=2)
    assert out["positions"].shape == (B, 4, n, 3)
    for b in range(B):
        chk = oracle_mod.run(x0[b], v0[b], oracle_mod.accel_direct(x0[b], m, 0.05), m, 1e-3, 0.05, 6, 2)
        assert np.abs(out["positions"][b] - chk["positions"]).max() < POS_TOL
        assert np.abs(out["velocities"][b] - chk["velocities"]).max() < POS_TOL
        assert np.abs(out["final_accelerations"][b] - chk["final_accelerations"]).max() <= 1e-10 * max(
            np.abs(chk["final_accelerations"]).max(), 1e-300)


def test_ensemble_f32_and_leftover_scheduler(engine):
    """float32 ensemble within 1e-5 of float64; B slightly above the worker count (two lanes per SM) and far
    above it give the same bits as small batches, with per-system masses."""
    from hpc.ensemble import simulate_ensemble
    rng = np.random.RandomState(5)
    grid = 2 * engine.sm_count
    for B in (grid + 3, 2 * grid + grid // 2):
        x0 = rng.rand(B, 48, 3) * 4 - 2
        v0 = rng.rand(B, 48, 3) - 0.5
        m = rng.uniform(1e9, 1e10, (B, 48))
        big = simulate_ensemble(x0, v0, m, dt=1e-3, softening=0.05, n_steps=60, save_interval=5)
        for lo in (0, B - 50):
            small = simulate_ensemble(x0[lo:lo + 50], v0[lo:lo + 50], m[lo:lo + 50], dt=1e-3, softening=0.05,
                                      n_steps=60, save_interval=5)
            for key in ("positions", "velocities", "accelerations", "final_positions", "final_accelerations"):
                assert np.array_equal(big[key][lo:lo + 50], small[key]), (B, lo, key)
    f32 = simulate_ensemble(x0[:8], v0[:8], m[:8], dt=1e-3, softening=0.05, n_steps=60, save_interval=5, dtype="float32")
    f64 = simulate_ensemble(x0[:8], v0[:8], m[:8], dt=1e-3, softening=0.05, n_steps=60, save_interval=5)
    assert np.abs(f32["positions"] - f64["positions"]).max() / np.abs(f64["positions"]).max() < 1e-5


@pytest.mark.parametrize("n_steps", [0, 1, 2, 3, 7])
def test_ensemble_short_runs_hand_over(engine, n_steps):
    """Runs of a few steps with systems shared between workers: heads of one or two steps, a tail that waits for
    a flag raised by the other lane of the same CTA, the n_steps == 0 launch (a_0 and the first snapshot only).
    B is chosen so that the systems x steps line does not divide by the worker count."""
    from hpc.ensemble import simulate_ensemble
    rng = np.random.RandomState(11 + n_steps)
    workers = 2 * engine.sm_count
    for B in (workers + 1, workers + workers // 2 + 1, engine.sm_count + 3):
        x0 = rng.rand(B, 24, 3) * 4 - 2
        v0 = rng.rand(B, 24, 3) - 0.5
        m = rng.uniform(1e9, 1e10, 24)
        big = simulate_ensemble(x0, v0, m, dt=1e-3, softening=0.05, n_steps=n_steps, save_interval=1)
        assert big["positions"].shape == (B, n_steps + 1, 24, 3)
        for lo in (0, B // 2, B - 40):
            small = simulate_ensemble(x0[lo:lo + 40], v0[lo:lo + 40], m, dt=1e-3, softening=0.05, n_steps=n_steps,
                                      save_interval=1)
            for key in ("positions", "velocities", "accelerations", "final_positions", "final_velocities",
                        "final_accelerations"):
                assert np.array_equal(big[key][lo:lo + 40], small[key]), (B, lo, key)


def test_ensemble_two_lanes_float32_and_n200(engine, oracle_mod):
    """The two-lane build (integrator warps; B >= 2 x SMs) in float32 and at the compile-time N = 200 shape:
    float64 rows equal the oracle, float32 within 1e-5, and both equal their own one-lane runs bit for bit."""
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    B = 2 * engine.sm_count + 7
    x0, v0, m32 = ics.datagen_ensemble_ic(B, 200, seed=321)
    kw = dict(dt=1e-3, softening=0.05, n_steps=12, save_interval=3)
    out64 = simulate_ensemble(x0, v0, m32, **kw)
    out32 = simulate_ensemble(x0, v0, m32, dtype="float32", **kw)
    for b in (0, B // 2, B - 1):
        chk = oracle_mod.run(x0[b], v0[b], oracle_mod.accel_direct(x0[b], m32, 0.05), m32, 1e-3, 0.05, 12, 3)
        assert np.abs(out64["positions"][b] - chk["positions"]).max() < POS_TOL
        assert np.abs(out64["velocities"][b] - chk["velocities"]).max() < 1e-8 * max(1.0, np.abs(chk["velocities"]).max())
        scale = np.abs(chk["positions"]).max()
        assert np.abs(out32["positions"][b] - chk["positions"]).max() / scale < 1e-5
    for out, dtype in ((out64, "float64"), (out32, "float32")):
        for lo in (0, B - 30):
            small = simulate_ensemble(x0[lo:lo + 30], v0[lo:lo + 30], m32, dtype=dtype, **kw)   # one lane, 30 CTAs
            for key in ("positions", "velocities", "accelerations", "final_positions"):
                assert np.array_equal(out[key][lo:lo + 30], small[key]), (dtype, lo, key)


@pytest.mark.parametrize("save_interval", [1, 4])
def test_ensemble_chunked_drain_equals_one_launch(engine, oracle_mod, monkeypatch, save_interval):
    """The bench's end-to-end path: outputs of 32 MB and more are produced by several launches (step chunks whose
    snapshot rows drain to pinned host memory while the next chunk runs; every launch re-enters the systems from
    their parked state with its own snap_offset).  Must equal the single-launch result bit for bit, and the oracle."""
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    B, N, T = 300, 200, 64
    x0, v0, m32 = ics.datagen_ensemble_ic(B, N, seed=77)
    kw = dict(dt=1e-3, n_steps=T, save_interval=save_interval)
    monkeypatch.setenv("NBODY_D2H_CHUNKS", "8")
    chunked = simulate_ensemble(x0, v0, m32, **kw)
    monkeypatch.setenv("NBODY_D2H_CHUNKS", "1")
    single = simulate_ensemble(x0, v0, m32, **kw)
    assert chunked["positions"].shape == (B, 1 + T // save_interval, N, 3)
    for key in ("positions", "velocities", "accelerations", "final_positions", "final_velocities", "final_accelerations"):
        assert np.array_equal(chunked[key], single[key]), key
    for b in (0, 151, B - 1):
        chk = oracle_mod.run(x0[b], v0[b], oracle_mod.accel_direct(x0[b], m32, 1e-9), m32, 1e-3, 1e-9, T, save_interval)
        k = min(chunked["positions"].shape[1], 1 + 48 // save_interval)  # the default ICs are chaotic: first 48 steps
        assert np.abs(chunked["positions"][b, :k] - chk["positions"][:k]).max() < POS_TOL


@pytest.mark.parametrize("n", [33, 600, 1024])
def test_ensemble_two_lanes_other_body_counts(engine, n):
    """Two lanes at the runtime-shape build: an odd small N, one j-part (N > 512) and the shared-memory limit, where
    two lanes only just fit; equal to one-lane batches bit for bit."""
    from hpc.ensemble import simulate_ensemble
    rng = np.random.RandomState(n)
    B = 2 * engine.sm_count + 5
    x0 = rng.rand(B, n, 3) * 4 - 2
    v0 = rng.rand(B, n, 3) - 0.5
    m = rng.uniform(1e9, 1e10, n)
    kw = dict(dt=1e-3, softening=0.05, n_steps=5, save_interval=2)
    big = simulate_ensemble(x0, v0, m, **kw)
    for lo in (0, B - 20):
        small = simulate_ensemble(x0[lo:lo + 20], v0[lo:lo + 20], m, **kw)
        for key in ("positions", "velocities", "accelerations", "final_positions", "final_accelerations"):
            assert np.array_equal(big[key][lo:lo + 20], small[key]), (n, lo, key)


@pytest.mark.parametrize("n,dtype", [(16, "float64"), (17, "float64"), (200, "float64"), (200, "float32"),
                                     (257, "float64"), (600, "float32"), (1024, "float64")])
def test_cluster_kernel_bitwise_equals_one_cta_kernel(engine, monkeypatch, n, dtype):
    """Few systems run one per cluster of 8 CTAs (positions exchanged by st.async through distributed shared memory); the
    j-parts and every sum are those of the one-CTA kernel, so the results are the same bits -- also with slabs that
    do not divide (N = 17: two CTAs of the cluster own nothing), per-system masses, n_steps = 0 and given a_0."""
    from hpc.ensemble import simulate_ensemble
    rng = np.random.RandomState(n)
    B = 3
    x0 = rng.rand(B, n, 3) * 4 - 2
    v0 = rng.rand(B, n, 3) - 0.5
    m = rng.uniform(1e9, 1e10, (B, n))
    a0 = rng.rand(B, n, 3)
    for kw in (dict(n_steps=7, save_interval=2), dict(n_steps=0, save_interval=1), dict(n_steps=5, save_interval=1, accelerations=a0)):
        monkeypatch.delenv("NB_ENSEMBLE_NO_CLUSTER", raising=False)
        clu = simulate_ensemble(x0, v0, m, dt=1e-3, softening=0.05, dtype=dtype, **kw)
        monkeypatch.setenv("NB_ENSEMBLE_NO_CLUSTER", "1")
        one = simulate_ensemble(x0, v0, m, dt=1e-3, softening=0.05, dtype=dtype, **kw)
        for key in ("positions", "velocities", "accelerations", "final_positions", "final_velocities", "final_accelerations"):
            assert np.array_equal(clu[key], one[key]), (n, dtype, kw.get("n_steps"), key)


@pytest.mark.parametrize("B", [10, 18, 19, 37, 38, 74, 75])
def test_cluster_sizes_bitwise_equal_one_cta_kernel(engine, monkeypatch, B):
    """The launch gives every system the largest cluster (8, 4 or 2 CTAs) that C x B <= SMs allows -- 10 systems
    (evaluate.py's ground truth) run on clusters of 8, 19..37 on 4, 38..74 on 2, 75 on single CTAs: same bits."""
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    x0, v0, m32 = ics.datagen_ensemble_ic(B, 200, seed=B)
    kw = dict(dt=1e-3, n_steps=9, save_interval=3)
    monkeypatch.delenv("NB_ENSEMBLE_NO_CLUSTER", raising=False)
    clu = simulate_ensemble(x0, v0, m32, **kw)
    monkeypatch.setenv("NB_ENSEMBLE_NO_CLUSTER", "1")
    one = simulate_ensemble(x0, v0, m32, **kw)
    for key in ("positions", "velocities", "accelerations", "final_positions", "final_velocities", "final_accelerations"):
        assert np.array_equal(clu[key], one[key]), (B, key)


def test_interval_schedule_with_another_tenant_on_the_gpu(engine):
    """The persistent ensemble kernel's workers hand systems to each other through flags and spin on them: every CTA
    must be resident at once.  With half the SMs held by another stream's kernel for 300 ms (nb_probe_occupy: one
    CTA per SM holding 200 KB of shared memory) the launch -- cooperative whenever systems are shared -- still
    completes, with the bits of the undisturbed run, instead of tail workers spinning for CTAs that are not running."""
    import ctypes
    import torch
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    B = 2 * engine.sm_count + 9                     # more systems than workers: hand-over flags in use
    x0, v0, m32 = ics.datagen_ensemble_ic(B, 200, seed=77)
    kw = dict(dt=1e-3, softening=1e-9, n_steps=30, save_interval=5)
    quiet = simulate_ensemble(x0, v0, m32, **kw)
    other = torch.cuda.Stream()
    rc = engine.lib.nb_probe_occupy(engine.sm_count // 2, 200 * 1024, 300.0, ctypes.c_void_p(other.cuda_stream))
    assert rc == 0, engine.lib.nb_last_error()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    busy = simulate_ensemble(x0, v0, m32, **kw)
    t1.record()
    torch.cuda.synchronize()
    for key in ("positions", "velocities", "accelerations", "final_positions"):
        assert np.array_equal(busy[key], quiet[key]), key
    assert t0.elapsed_time(t1) < 3000.0


@pytest.mark.parametrize("n,dtype,steps,every", [(700, "float64", 5, 1), (1000, "float32", 4, 2), (4099, "float32", 7, 3),
                                                 (4096, "float64", 3, 1), (16384, "float32", 6, 6), (9000, "float64", 2, 1),
                                                 (18000, "float32", 3, 1), (20000, "float32", 3, 1), (32768, "float32", 2, 2),
                                                 (33, "float64", 4, 1), (641, "float32", 5, 5)])
def test_step_kernel_variants_are_bit_identical(engine, monkeypatch, n, dtype, steps, every):
    """The three ways a whole system is stepped on one GPU give the same BITS -- same segment plan, same inner loops,
    same epilogue arithmetic: K2s (csrc/nb_group.cu, the default up to ~18,900 bodies: one CTA per group of bodies,
    segment partials reduced in shared memory), K2 (nb_force.cu: (i-tile x segment) grid, partials reduced by the
    last CTA of a tile; NB_NO_GROUP=1 -- what i-slabs and larger systems always use) and K2p (nb_persist.cu, opt-in
    NB_PERSIST=1: all steps in one cooperative launch).  Final state, every snapshot row, and the accelerations of a
    plain force evaluation."""
    import torch
    from hpc import ics
    from hpc.sharded import ShardedSystem
    x, v, m = ics.plummer_ic(n, seed=13)
    res = {}
    for mode, env in (("group", {}), ("tiles", {"NB_NO_GROUP": "1"}), ("persist", {"NB_NO_GROUP": "1", "NB_PERSIST": "1"})):
        for key in ("NB_NO_GROUP", "NB_PERSIST"):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        s = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=np.dtype(dtype), device=engine.device)
        a0 = s.acc.clone()
        n_snap = 1 + steps // every
        snaps = tuple(torch.zeros((n_snap, n, 3), dtype=torch.float64, device=engine.device) for _ in range(3))
        s.advance(steps, *snaps, save_interval=every)
        s.advance(2)                                   # a second call continues from the state the first left
        engine.step_status(s.ws, n)
        res[mode] = (a0, s.cur[:engine.padded_bodies(n) * 4].clone(), s.vel.clone(), s.acc.clone()) + tuple(t.clone() for t in snaps)
    for mode in ("tiles", "persist"):
        for a, b in zip(res["group"], res[mode]):
            assert torch.equal(a, b), mode
    assert torch.isfinite(res["group"][1]).all() and (res["group"][4][1:] != 0).any()


@pytest.mark.parametrize("n,B,dtype", [(1100, 5, "float64"), (1500, 7, "float32"), (3000, 3, "float64"), (2049, 4, "float32")])
def test_batched_mid_size_ensemble_equals_per_system_runs(engine, oracle_mod, n, B, dtype):
    """Ensembles of systems too large for one CTA's shared memory (what the unchanged generate_data.py asks for with
    --particles in the thousands): all B systems side by side in every launch (K2s, grid = groups x B) must give the
    BITS of running each system on its own, snapshots included, and match the oracle."""
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    x0 = np.empty((B, n, 3))
    v0 = np.empty((B, n, 3))
    for b in range(B):
        x0[b], v0[b], m = ics.plummer_ic(n, seed=50 + b)
    masses = np.stack([m * (1.0 + 0.1 * b) for b in range(B)]) if B % 2 else m          # per-system and shared masses
    kw = dict(dt=1e-3, softening=0.01, n_steps=9, save_interval=4, dtype=dtype)
    out = simulate_ensemble(x0, v0, masses, **kw)
    assert out["positions"].shape == (B, 3, n, 3) and np.isfinite(out["positions"]).all()
    for b in range(B):
        mb = masses[b] if masses.ndim == 2 else masses
        a0 = engine.accelerations(x0[b], mb, 0.01, np.dtype(dtype))
        one = engine.run(x0[b], v0[b], a0, mb, 1e-3, 0.01, 9, 4, dtype=np.dtype(dtype))
        for key in ("positions", "velocities", "accelerations", "final_positions", "final_velocities",
                    "final_accelerations"):
            assert np.array_equal(out[key][b], one[key]), (b, key)
    if dtype == "float64":
        mb = masses[0] if masses.ndim == 2 else masses
        chk = oracle_mod.run(x0[0], v0[0], oracle_mod.accel_direct(x0[0], mb, 0.01), mb, 1e-3, 0.01, 9, 4)
        assert np.abs(out["positions"][0] - chk["positions"]).max() < POS_TOL
    dev = simulate_ensemble(x0, v0, masses, outputs="device", **kw)
    assert np.array_equal(dev["positions"].cpu().numpy(), out["positions"])
