"""energy / momentum of every snapshot of a trajectory stack (K4b, hpc.metrics).

CPU part: the NumPy restatement (oracle/numpy_oracle.snapshot_energies) against tests/golden/metrics_reference.npz,
which holds what the reference's OWN compute_energy_error / compute_momentum_error (src/utils/metrics.py:62-137)
returned for these inputs (tests/golden/make_golden_metrics.py), and live against the reference where it is mounted.
GPU part: the CUDA kernel, through hpc.metrics / utils.metrics, against both."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference/src/utils/metrics.py")
RTOL = 1e-12        # float64 sums of the same terms in another order


def _cases(golden):
    g = golden("metrics_reference.npz")
    p = golden("traj_plummer_n200.npz")
    d = golden("ensemble_default_b4_n200_t20.npz")
    m32 = np.random.RandomState(42).uniform(1e10, 1e12, 200).astype(np.float32)
    cases = [("plummer", p["positions"], p["velocities"], p["masses"], 6.67430e-11, float(p["softening"]))]
    cases += [(f"default{b}", d["positions"][b], d["velocities"][b], m32, 6.67430e-11, 1e-9) for b in range(4)]
    cases += [(t, g[f"{t}_pos"], g[f"{t}_vel"], g[f"{t}_masses"], 2.5e-11, 0.05) for t in ("odd", "two", "one", "n33")]
    return g, cases


def _check(g, tag, energies, momentum_vec, scale):
    ref_e = g[f"{tag}_energies"]
    assert np.allclose(energies, ref_e, rtol=RTOL, atol=RTOL * np.abs(ref_e).max()), tag
    mag = np.linalg.norm(momentum_vec, axis=-1)
    ref_p = g[f"{tag}_momentum"]
    assert np.allclose(mag, ref_p, rtol=1e-9, atol=1e-13 * scale), tag


def test_numpy_restatement_matches_reference_outputs(golden):
    from oracle import numpy_oracle
    g, cases = _cases(golden)
    for tag, pos, vel, m, G, eps in cases:
        K, U, P = numpy_oracle.snapshot_energies(pos, vel, m, G, eps)
        scale = float((np.asarray(m, dtype=np.float64)[None, :, None] * np.abs(vel)).sum(axis=(1, 2)).max())
        _check(g, tag, K + U, P, scale)


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")
def test_numpy_restatement_matches_reference_live():
    from oracle import numpy_oracle
    spec = importlib.util.spec_from_file_location("ref_metrics", str(REF))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.RandomState(5)
    pos, vel, m = rng.standard_normal((5, 40, 3)), rng.standard_normal((5, 40, 3)), rng.uniform(1e10, 1e12, 40)
    e, err = ref.compute_energy_error(pos, vel, m)
    pm, perr = ref.compute_momentum_error(vel, m)
    K, U, P = numpy_oracle.snapshot_energies(pos, vel, m)
    assert np.allclose(K + U, e, rtol=RTOL) and np.allclose(np.linalg.norm(P, axis=1), pm, rtol=1e-10)


@pytest.mark.gpu
def test_gpu_metrics_match_reference_outputs(golden):
    from utils.metrics import compute_energy_error, compute_momentum_error       # the reference's import path
    g, cases = _cases(golden)
    for tag, pos, vel, m, G, eps in cases:
        e, err = compute_energy_error(pos, vel, m, G=G, softening=eps)
        pm, perr = compute_momentum_error(vel, m)
        assert e.shape == (pos.shape[0],) and isinstance(err, float) and isinstance(perr, float)
        ref_e = g[f"{tag}_energies"]
        assert np.allclose(e, ref_e, rtol=RTOL, atol=RTOL * np.abs(ref_e).max()), tag
        assert np.isclose(err, float(g[f"{tag}_energy_error"]), rtol=1e-6, atol=1e-13), tag
        # |sum m v| of a system at rest (Plummer: zero net momentum by construction) is rounding noise of the sum:
        # the scale of the comparison is sum m |v|, not the result
        scale = float((np.asarray(m, dtype=np.float64)[None, :, None] * np.abs(vel)).sum(axis=(1, 2)).max())
        assert np.allclose(pm, g[f"{tag}_momentum"], rtol=1e-9, atol=1e-13 * scale), tag
        # the relative momentum error divides by |p_0|: compare only where that is not noise
        if float(g[f"{tag}_momentum"][0]) > 1e-6 * scale:
            assert np.isclose(perr, float(g[f"{tag}_momentum_error"]), rtol=1e-6, atol=1e-12), tag


@pytest.mark.gpu
def test_gpu_batched_device_stacks_vs_restatement(engine):
    """The batched form on the stacks an ensemble launch leaves in HBM: every (system, snapshot) against the NumPy
    restatement; shared float32 masses and per-system float64 masses; body counts around the pairing scheme's edges."""
    import torch
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    from hpc.metrics import snapshot_energies
    from oracle import numpy_oracle
    x0, v0, m32 = ics.datagen_ensemble_ic(6, 200, seed=42)
    dev = simulate_ensemble(x0, v0, m32, dt=1e-3, n_steps=12, save_interval=3, outputs="device")
    res = snapshot_energies(dev["positions"], dev["velocities"], m32)
    assert res["total"].shape == (6, 5) and res["momentum"].shape == (6, 5, 3)
    pos_h, vel_h = dev["positions"].cpu().numpy(), dev["velocities"].cpu().numpy()
    for b in range(6):
        K, U, P = numpy_oracle.snapshot_energies(pos_h[b], vel_h[b], m32)
        assert np.allclose(res["kinetic"][b], K, rtol=RTOL) and np.allclose(res["potential"][b], U, rtol=RTOL)
        assert np.allclose(res["momentum"][b], P, rtol=1e-10, atol=1e-12 * np.abs(P).max())
    rng = np.random.RandomState(8)
    for N in (1, 2, 3, 31, 32, 64, 257, 1000):
        pos, vel = rng.standard_normal((3, 2, N, 3)), rng.standard_normal((3, 2, N, 3))
        m = rng.uniform(1e9, 1e11, (3, N))
        res = snapshot_energies(pos, vel, m, G=1.0e-10, softening=0.02)
        for b in range(3):
            K, U, P = numpy_oracle.snapshot_energies(pos[b], vel[b], m[b], 1.0e-10, 0.02)
            assert np.allclose(res["total"][b], K + U, rtol=RTOL, atol=RTOL * np.abs(U).max()), N
            assert np.allclose(res["momentum"][b], P, rtol=1e-10, atol=1e-12 * np.abs(P).max()), N
    # one large system per call goes through K4 state by state
    x, v, m = ics.plummer_ic(5000, seed=7)
    big = snapshot_energies(np.stack([x, x * 1.01]), np.stack([v, v]), m, softening=0.01)
    k0, u0, _ = engine.energy(x, v, m, 0.01)
    assert np.isclose(big["kinetic"][0], k0, rtol=1e-12) and np.isclose(big["potential"][0], u0, rtol=1e-12)
    assert np.allclose(big["momentum"][0], (m[:, None] * v).sum(axis=0), rtol=1e-9, atol=1e-12 * np.abs(m[:, None] * v).sum())
