"""Host-side behaviour of the hpc.nbody mirror, with the CUDA engine replaced by the oracle-backed
stand-in (tests/fake_engine.py): what the reference's callers rely on -- RNG order, state dicts,
dtype preservation, time accumulation, save_interval / verbose bookkeeping, the batched datagen
driver.  Arithmetic parity of the real kernels is the GPU tests' job."""
import numpy as np
import pytest

from fake_engine import FakeEngine
from hpc import ics, nbody
from hpc.ensemble import generate_simulations, generate_single_simulation, simulate_ensemble


@pytest.fixture()
def fake(oracle_mod, monkeypatch):
    """The seam is in the TESTS: hpc._cuda.get_engine is patched to hand out the stand-in (the product has no
    backend switch)."""
    from hpc import _cuda
    eng = FakeEngine()
    monkeypatch.setattr(_cuda, "get_engine", lambda device=None: eng)
    return eng


def test_constructor_draw_order_and_global_rng(fake, golden):
    g = golden("accel_default_n200.npz")
    sim = nbody.NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42)
    x, v, m = ics.reference_default_ic(200, 42)
    assert np.array_equal(sim.positions, x) and np.array_equal(sim.velocities, v) and np.array_equal(sim.masses, m)
    assert np.abs(sim.accelerations - g["acc_ctor_f64mass_seed42"]).max() < 1e-9 * np.abs(sim.accelerations).max()
    # the constructor consumed the GLOBAL generator exactly as the reference does (nbody.py:175-181)
    nxt = np.random.rand()
    np.random.seed(42)
    np.random.rand(200, 3), np.random.rand(200, 3), np.random.uniform(1e10, 1e12, 200)
    assert nxt == np.random.rand()
    assert (sim.time, sim.step_count, sim.history) == (0.0, 0, [])
    assert sim.use_barnes_hut is False and sim.theta == 0.5 and sim.softening == nbody.SOFTENING


def test_datagen_call_sequence_matches_reference_run(fake, golden):
    """generate_data.py:36-49: construct, assign float32 masses, recompute a, run(400)."""
    g = golden("traj_default_n200_seed42.npz")
    sim = nbody.NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42, use_barnes_hut=False)
    sim.masses = ics.shared_masses(200, 42).copy()
    sim.accelerations = sim._compute_accelerations()
    states = sim.run(400, save_interval=1, verbose=False)
    assert len(states) == 401 and sim.history is states
    assert sorted(states[0]) == ["accelerations", "masses", "positions", "step", "time", "velocities"]
    assert states[7]["masses"].dtype == np.float32 and states[7]["positions"].dtype == np.float64
    assert isinstance(states[7]["time"], float) and isinstance(states[7]["step"], int)
    assert np.array_equal(np.array([s["time"] for s in states]), g["times"])      # running float sum of dt
    assert [s["step"] for s in states] == list(g["steps"])
    for r, k in enumerate(g["steps_kept"]):
        if k <= 50:
            assert np.abs(states[k]["positions"] - g["positions"][r]).max() < 1e-9
    assert sim.time == float(g["final_time"]) and sim.step_count == 400
    # states are independent copies
    states[3]["positions"][0, 0] = 123.0
    assert states[4]["positions"][0, 0] != 123.0 and sim.positions[0, 0] != 123.0
    assert ("run", 400, 1, True) in fake.calls            # ONE device run, not 400 round trips


def test_save_interval_and_verbose_segmentation(fake, golden, capsys):
    g = golden("bookkeeping_n16.npz")
    for verbose in (False, True):
        sim = nbody.NBodySimulator(n_particles=16, box_size=10.0, dt=0.001, seed=5)
        states = sim.run(50, save_interval=7, verbose=verbose)
        assert [s["step"] for s in states] == list(g["steps"])
        assert np.array_equal(np.array([s["time"] for s in states]), g["times"])
        assert np.abs(np.stack([s["positions"] for s in states]) - g["positions"]).max() < 1e-10
        assert sim.step_count == 50 and sim.time == float(g["final_time"])
        assert np.abs(sim.positions - g["final_positions"]).max() < 1e-10
    out = capsys.readouterr().out
    assert out.count("Step ") == 10 and "Step 50/50" in out and "Energy:" in out


def test_step_in_place_and_set_state(fake):
    sim = nbody.NBodySimulator(n_particles=16, box_size=10.0, dt=0.001, seed=5)
    pos_alias, acc_before = sim.positions, sim.accelerations
    sim.step()
    assert sim.positions is pos_alias            # += semantics: updated in place (nbody.py:208)
    assert sim.accelerations is not acc_before   # rebinding (nbody.py:211)
    assert sim.step_count == 1 and sim.time == 0.001
    st = sim.get_state()
    sim.step()
    sim.set_state(st)
    assert sim.step_count == 1 and np.array_equal(sim.positions, st["positions"])
    assert sim.positions is not st["positions"]
    second = sim.run(3, verbose=False)
    assert [s["step"] for s in second] == [1, 2, 3, 4]
    assert second[-1]["time"] == ((((0.0 + 0.001) + 0.001) + 0.001) + 0.001)


def test_run_zero_steps_and_empty_system(fake):
    sim = nbody.NBodySimulator(n_particles=5, seed=1)
    states = sim.run(0, verbose=False)
    assert len(states) == 1 and states[0]["step"] == 0
    assert nbody.compute_accelerations_direct(np.zeros((0, 3)), np.zeros(0)).shape == (0, 3)
    with pytest.raises(ValueError):
        nbody.compute_accelerations_direct(np.zeros((4, 2)), np.ones(4))
    with pytest.raises(ValueError):
        nbody.compute_accelerations_direct(np.zeros((4, 3)), np.ones(5))
    with pytest.raises(ValueError):
        nbody.NBodySimulator(n_particles=4, seed=1, dtype="float16")


def test_factories_and_module_functions(fake, golden):
    np.random.seed(0)
    sol = nbody.NBodySimulator.create_solar_system()
    g = golden("solar_system.npz")
    assert np.array_equal(sol.positions, g["positions"]) and np.array_equal(sol.velocities, g["velocities"])
    assert np.array_equal(sol.masses, g["masses"]) and sol.dt == float(g["dt"])
    gal = nbody.NBodySimulator.create_galaxy_collision(n_per_galaxy=40)
    assert gal.positions.shape == (80, 3) and gal.positions[:40, 0].mean() < 0 < gal.positions[40:, 0].mean()
    x, v, m = ics.plummer_ic(64, seed=2)
    a = nbody.compute_accelerations_direct(x, m, 0.01)
    nx, vh, a2 = nbody.leapfrog_step(x, v, a, m, 1e-3)
    assert np.array_equal(vh, v + 0.5 * 1e-3 * a) and np.array_equal(nx, x + 1e-3 * vh) and a2 is a
    k, u, e = nbody.compute_total_energy(x, v, m, 0.01)
    assert e == k + u and u < 0 < k
    res = nbody.run_parallel_simulations([{"init": {"n_particles": 8, "seed": 1}, "run": {"n_steps": 3, "verbose": False}}] * 2)
    assert len(res) == 2 and len(res[0]) == 4
    assert (nbody.G, nbody.SOFTENING) == (6.67430e-11, 1e-9)


def test_dtype_option_and_env(fake, monkeypatch):
    sim = nbody.NBodySimulator(n_particles=8, seed=1, dtype="float32")
    assert sim.dtype == np.float32 and ("accelerations", "float32") in fake.calls
    monkeypatch.setenv("NBODY_DTYPE", "float32")
    assert nbody.NBodySimulator(n_particles=8, seed=1).dtype == np.float32


def test_batched_datagen_driver(fake, golden):
    g = golden("ensemble_default_b4_n200_t20.npz")
    m32 = ics.shared_masses(200, 42)
    args = [(i, 200, 20, 1, 10.0, 42 + i, m32) for i in range(4)]
    trajs = generate_simulations(args)
    for b, t in enumerate(trajs):
        assert sorted(t) == ["accelerations", "masses", "n_steps", "positions", "times", "velocities"]
        assert t["n_steps"] == 21 and t["masses"].dtype == np.float32 and t["positions"].shape == (21, 200, 3)
        assert np.abs(t["positions"] - g["positions"][b]).max() < 1e-10
    one = generate_single_simulation(args[2])
    assert np.array_equal(one["positions"], trajs[2]["positions"])
    # own (constructor) masses when shared_masses is None, mixed shapes grouped separately
    mixed = generate_simulations([(0, 30, 5, 1, 10.0, 7, None), (1, 20, 4, 2, 10.0, 8, None)])
    assert mixed[0]["positions"].shape == (6, 30, 3) and mixed[1]["positions"].shape == (3, 20, 3)
    assert np.array_equal(mixed[0]["masses"], ics.reference_default_ic(30, 7)[2])
    out = simulate_ensemble(np.zeros((2, 3, 3)) + np.arange(3)[None, :, None], np.zeros((2, 3, 3)), np.ones(3) * 1e10,
                            n_steps=6, save_interval=3)
    assert out["positions"].shape == (2, 3, 3, 3) and np.array_equal(out["times"], [0.0, 0.003, 0.006])


# ---- where the state lives between calls -----------------------------------------------------
def _count(fake, what):
    return sum(1 for c in fake.calls if c[0] == what)


def test_step_loop_keeps_the_state_on_the_device(fake, oracle_mod):
    """`for _ in range(k): sim.step()` (reference scripts/benchmark_bh_temp.py:24,32): one upload, no download until
    the caller looks at the state."""
    sim = nbody.NBodySimulator(n_particles=16, box_size=10.0, dt=0.001, seed=5)
    x, v, a, m = (sim.positions.copy(), sim.velocities.copy(), sim.accelerations.copy(), sim.masses.copy())
    fake.calls.clear()
    for _ in range(25):
        sim.step()
    assert _count(fake, "resident") == 1 and _count(fake, "download") == 0 and _count(fake, "run") == 25
    assert sim.step_count == 25
    e = sim.get_energy()                                  # evaluated where the state is: still no download
    assert _count(fake, "download") == 0
    chk = oracle_mod.run(x, v, a, m, 0.001, 1e-9, 25, 25)
    assert np.abs(sim.positions - chk["final_positions"]).max() < 1e-12      # first look: one download
    assert _count(fake, "download") == 1
    assert np.allclose(e, oracle_mod.total_energy(chk["final_positions"], chk["final_velocities"], m, 1e-9), rtol=1e-12)
    sim.velocities, sim.accelerations                     # already current
    assert _count(fake, "download") == 1
    sim.run(5, verbose=False)                             # the caller has looked: the host copy may have been written
    assert _count(fake, "resident") == 2


def test_aliases_see_in_place_updates_and_writes_are_honoured(fake, oracle_mod):
    sim = nbody.NBodySimulator(n_particles=12, box_size=10.0, dt=0.001, seed=9)
    m, a0 = sim.masses.copy(), sim.accelerations.copy()
    p = sim.positions                                     # an alias the caller keeps (a view counts too)
    vview = sim.velocities[:4]
    x0, v0 = p.copy(), sim.velocities.copy()
    sim.step()
    chk = oracle_mod.run(x0, v0, a0, m, 0.001, 1e-9, 1, 1)
    assert np.array_equal(p, chk["final_positions"]) and np.array_equal(vview, chk["final_velocities"][:4])
    assert sim.positions is p
    # writing through the alias between steps changes the simulation, as it does in the reference
    p[0, 0] += 0.5
    x1 = p.copy()
    v1 = sim.velocities.copy()
    a1 = sim.accelerations.copy()
    sim.step()
    chk2 = oracle_mod.run(x1, v1, a1, m, 0.001, 1e-9, 1, 1)
    assert np.array_equal(p, chk2["final_positions"])
    del p, vview
    # writing through the attribute without keeping a reference (the factories' pattern, nbody.py:294-300)
    sim.positions[3, 1] = 7.25
    x2, v2, a2 = sim.positions.copy(), sim.velocities.copy(), sim.accelerations.copy()
    sim.step()
    chk3 = oracle_mod.run(x2, v2, a2, m, 0.001, 1e-9, 1, 1)
    assert np.array_equal(sim.positions, chk3["final_positions"])
    # assigning masses (generate_data.py:46) re-uploads, keeps the dtype
    n_up = _count(fake, "resident")
    sim.masses = (m * 2).astype(np.float32)
    sim.step()
    assert _count(fake, "resident") == n_up + 1 and sim.masses.dtype == np.float32
    # dt and softening are plain attributes read at every call
    sim.dt = 0.002
    t = sim.time
    sim.step()
    assert sim.time == t + 0.002


def test_state_list_is_lazy_and_list_like(fake):
    import pickle
    sim = nbody.NBodySimulator(n_particles=10, box_size=10.0, dt=0.001, seed=3)
    states = sim.run(30, save_interval=3, verbose=False)
    assert isinstance(states, list) and len(states) == 11 and len(states._lazy) == 10     # nothing built yet
    assert states[4]["step"] == 12 and not states._lazy                                    # built on first access
    assert [s["step"] for s in states] == list(range(0, 31, 3))
    assert all(s["positions"].flags.writeable and s["positions"].shape == (10, 3) for s in states)
    states[2]["positions"][:] = -1.0                       # rows are disjoint: nothing else changes
    assert (states[3]["positions"] != -1.0).all() and (states[1]["positions"] != -1.0).all()
    assert (sim.positions != -1.0).all()
    assert np.array_equal(states[-1]["positions"], sim.positions)
    plain = pickle.loads(pickle.dumps(states))
    assert type(plain) is list and len(plain) == 11 and np.array_equal(plain[5]["velocities"], states[5]["velocities"])
    fresh = sim.run(4, verbose=False)
    assert len(fresh + [1]) == 6 and len(fresh.copy()) == 5 and fresh[-1]["step"] == 34
    stacked = np.stack([s["positions"] for s in sim.run(6, verbose=False)])      # generate_data.py:51
    assert stacked.shape == (7, 10, 3)
