#!/usr/bin/env python3
"""Generate tests/golden/*.npz by EXECUTING the reference's own Numba code.

Run in the authoring container only (needs /root/reference, numba, numpy):

    python tests/golden/make_golden.py

The reference module is loaded by file path because ``hpc/__init__.py`` there
imports h5py, which this image lacks.  Nothing is copied
from the reference: the files written hold inputs we generate and the outputs
its functions return.  The fixtures are small (a few hundred kB in total) and
are what pins the CPU oracle (tests/test_oracle_golden.py) and, through it, the
CUDA path.  /root/reference does not exist on the GPU box, which is why the
vectors are committed.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import ics  # noqa: E402  (host-only IC generators of this repo)

REF = Path("/root/reference/src/hpc/nbody.py")
KEEP_STEPS = [0, 1, 2, 5, 10, 20, 50, 100, 200, 400]


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_nbody", str(REF))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_run(ref, pos, vel, masses, dt, softening, n_steps):
    """Drive the reference simulator the way generate_data.py:36-49 does."""
    n = pos.shape[0]
    sim = ref.NBodySimulator(n_particles=n, box_size=10.0, dt=dt, softening=softening, seed=0)
    sim.positions = pos.copy()
    sim.velocities = vel.copy()
    sim.masses = masses.copy()
    sim.accelerations = sim._compute_accelerations()
    e0 = sim.get_energy()
    states = sim.run(n_steps, save_interval=1, verbose=False)
    e1 = sim.get_energy()
    return states, e0, e1


def pack_states(states, keep):
    return {
        "steps_kept": np.array(keep),
        "positions": np.stack([states[k]["positions"] for k in keep]),
        "velocities": np.stack([states[k]["velocities"] for k in keep]),
        "accelerations": np.stack([states[k]["accelerations"] for k in keep]),
        "times": np.array([s["time"] for s in states]),
        "steps": np.array([s["step"] for s in states]),
    }


def main():
    import numba

    ref = load_reference()
    meta = {
        "numba": numba.__version__,
        "numpy": np.__version__,
        "threading_layer_threads": numba.get_num_threads(),
        "reference_file": str(REF),
        "G": ref.G,
        "SOFTENING": ref.SOFTENING,
    }

    # 1. single-shot accelerations on the reference's default ICs (float32 shared masses)
    out = {}
    for seed in (42, 43, 9999):
        sim = ref.NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=seed)
        x, v, m64 = ics.reference_default_ic(200, seed)
        assert np.array_equal(sim.positions, x) and np.array_equal(sim.velocities, v)
        assert np.array_equal(sim.masses, m64), "RNG order of nbody.py:179-181 not reproduced"
        out[f"acc_ctor_f64mass_seed{seed}"] = sim.accelerations.copy()
        m32 = ics.shared_masses(200, 42)
        sim.masses = m32.copy()
        out[f"acc_f32mass_seed{seed}"] = sim._compute_accelerations()
        out[f"energy_f32mass_seed{seed}"] = np.array(sim.get_energy())
    np.savez_compressed(HERE / "accel_default_n200.npz", **out)

    # 2. default-IC trajectory, exactly the evaluate.py / generate_data.py call sequence
    x, v, _ = ics.reference_default_ic(200, 42)
    m32 = ics.shared_masses(200, 42)
    sim = ref.NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42)
    sim.masses = m32.copy()
    sim.accelerations = sim._compute_accelerations()
    states = sim.run(400, save_interval=1, verbose=False)
    d = pack_states(states, KEEP_STEPS)
    d["masses_dtype"] = np.array(str(states[0]["masses"].dtype))
    d["final_time"] = np.array(sim.time)
    d["final_step"] = np.array(sim.step_count)
    np.savez_compressed(HERE / "traj_default_n200_seed42.npz", **d)
    meta["state_keys"] = sorted(states[0].keys())
    meta["state_dtypes"] = {k: str(getattr(states[0][k], "dtype", type(states[0][k]).__name__))
                            for k in states[0]}

    # 2b. save_interval = 7 bookkeeping (snapshot count, steps, times)
    sim = ref.NBodySimulator(n_particles=16, box_size=10.0, dt=0.001, seed=5)
    st7 = sim.run(50, save_interval=7, verbose=False)
    np.savez_compressed(HERE / "bookkeeping_n16.npz",
                        positions0=st7[0]["positions"], velocities0=st7[0]["velocities"],
                        masses=st7[0]["masses"], accelerations0=st7[0]["accelerations"],
                        steps=np.array([s["step"] for s in st7]),
                        times=np.array([s["time"] for s in st7]),
                        positions=np.stack([s["positions"] for s in st7]),
                        final_positions=sim.positions, final_velocities=sim.velocities,
                        final_time=np.array(sim.time), final_step=np.array(sim.step_count))

    # 3. well-conditioned systems in N-body units: Plummer and uniform sphere
    for name, (x, v, m), n_steps in (
        ("plummer_n200", ics.plummer_ic(200, seed=7), 400),
        ("plummer_n1024", ics.plummer_ic(1024, seed=7), 400),
        ("sphere_n256", ics.uniform_sphere_ic(256, seed=11), 400),
    ):
        states, e0, e1 = ref_run(ref, x, v, m, dt=1e-3, softening=0.01, n_steps=n_steps)
        keep = [0, 1, 10, 100, 400]
        d = pack_states(states, keep)
        d.update(x0=x, v0=v, masses=m, energy0=np.array(e0), energy1=np.array(e1),
                 dt=np.array(1e-3), softening=np.array(0.01))
        np.savez_compressed(HERE / f"traj_{name}.npz", **d)

    # 4. a larger single evaluation (N=2048 Plummer), rows subsampled to keep the file small
    x, v, m = ics.plummer_ic(2048, seed=7)
    acc = ref.compute_accelerations_direct(x, m, 0.01)
    rows = np.arange(0, 2048, 16)
    np.savez_compressed(HERE / "accel_plummer_n2048_rows.npz", rows=rows, acc_rows=acc[rows],
                        softening=np.array(0.01), energy=np.array(ref.compute_total_energy(x, v, m, 0.01)))

    # 5. known-answer: the reference's own solar-system factory
    np.random.seed(0)
    sol = ref.NBodySimulator.create_solar_system()
    np.savez_compressed(HERE / "solar_system.npz", positions=sol.positions, velocities=sol.velocities,
                        masses=sol.masses, accelerations=sol.accelerations,
                        softening=np.array(sol.softening), dt=np.array(sol.dt))

    # 6. small ensemble, first 20 steps of 4 data-generation simulations (seeds 42..45)
    x0, v0, m32 = ics.datagen_ensemble_ic(4, 200, seed=42)
    ens_p, ens_v, ens_a = [], [], []
    for b in range(4):
        sim = ref.NBodySimulator(n_particles=200, box_size=10.0, dt=0.001, seed=42 + b)
        assert np.array_equal(sim.positions, x0[b])
        sim.masses = m32.copy()
        sim.accelerations = sim._compute_accelerations()
        st = sim.run(20, save_interval=1, verbose=False)
        ens_p.append(np.stack([s["positions"] for s in st]))
        ens_v.append(np.stack([s["velocities"] for s in st]))
        ens_a.append(np.stack([s["accelerations"] for s in st]))
    np.savez_compressed(HERE / "ensemble_default_b4_n200_t20.npz", positions=np.stack(ens_p),
                        velocities=np.stack(ens_v), accelerations=np.stack(ens_a))

    (HERE / "golden_meta.json").write_text(json.dumps(meta, indent=1, sort_keys=True) + "\n")
    total = sum(p.stat().st_size for p in HERE.glob("*.npz"))
    print(f"wrote {len(list(HERE.glob('*.npz')))} fixtures, {total / 1024:.0f} kB")


if __name__ == "__main__":
    os.environ.setdefault("NUMBA_NUM_THREADS", "8")
    main()
