"""GPU parity, K1: accelerations from the CUDA kernels (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): float64 kernel <= 1e-10 relative on accelerations
(per-particle vector norm); float32 kernel <= 1e-5 relative (global max-norm, the norm SURVEY.md 8c
shows is the meaningful one for single precision).
"""
import ctypes

import numpy as np
import pytest

from conftest import rel_rows

pytestmark = pytest.mark.gpu

F64_TOL = 1e-10
F32_TOL = 1e-5


def _ics():
    from hpc import ics
    return ics


@pytest.mark.parametrize("seed", [42, 43, 9999])
def test_accel_f64_default_ics_vs_golden(golden, seed):
    """Reference-default ICs, float32 shared masses: against vectors produced by the reference itself."""
    from hpc.nbody import compute_accelerations_direct
    ics = _ics()
    g = golden("accel_default_n200.npz")
    x, _, m64 = ics.reference_default_ic(200, seed)
    m32 = ics.shared_masses(200, 42)
    a = compute_accelerations_direct(x, m32)
    assert a.dtype == np.float64 and a.shape == (200, 3)
    assert rel_rows(a, g[f"acc_f32mass_seed{seed}"]).max() < F64_TOL
    a = compute_accelerations_direct(x, m64)
    assert rel_rows(a, g[f"acc_ctor_f64mass_seed{seed}"]).max() < F64_TOL


@pytest.mark.parametrize("n,kind", [(1, "default"), (2, "default"), (3, "default"), (31, "default"), (33, "default"),
                                    (200, "default"), (257, "plummer"), (1000, "sphere"), (2048, "plummer"),
                                    (4099, "plummer"), (16384, "plummer")])
def test_accel_f64_vs_oracle(oracle_mod, n, kind):
    from hpc.nbody import compute_accelerations_direct
    ics = _ics()
    if kind == "default":
        x, _, m = ics.reference_default_ic(n, 5)
        eps = 1e-9
    elif kind == "plummer":
        x, _, m = ics.plummer_ic(n, seed=7)
        eps = 0.01
    else:
        x, _, m = ics.uniform_sphere_ic(n, seed=11)
        eps = 0.01
    a = compute_accelerations_direct(x, m, eps)
    ref = oracle_mod.accel_direct(x, m, eps)
    if n == 1:
        assert np.array_equal(a, np.zeros((1, 3)))
        return
    assert rel_rows(a, ref).max() < F64_TOL


@pytest.mark.parametrize("n,kind", [(2, "default"), (200, "default"), (1000, "sphere"), (4099, "plummer"),
                                    (16384, "plummer"), (65536, "plummer")])
def test_accel_f32_vs_oracle(oracle_mod, n, kind):
    from hpc.nbody import compute_accelerations_direct
    ics = _ics()
    if kind == "default":
        x, _, m = ics.reference_default_ic(n, 5)
        m = ics.shared_masses(n, 42)
        eps = 1e-9
    elif kind == "plummer":
        x, _, m = ics.plummer_ic(n, seed=7)
        eps = 0.01
    else:
        x, _, m = ics.uniform_sphere_ic(n, seed=11)
        eps = 0.01
    a = compute_accelerations_direct(x, m, eps, dtype="float32")
    assert a.dtype == np.float64
    if n > 20000:   # the full oracle would take a minute: check 512 spot rows with the NumPy restatement
        rows = np.arange(0, n, n // 512)
        ref = oracle_mod.accel_rows_numpy(x, m, rows, eps)
        a = a[rows]
    else:
        ref = oracle_mod.accel_direct(x, m, eps)
    assert np.abs(a - ref).max() / np.abs(ref).max() < F32_TOL
    assert np.median(rel_rows(a, ref)) < 2e-6


def test_accel_large_f64_spot_rows(oracle_mod):
    """N = 65,536 float64 (multi-segment, big-tile variant) against spot rows of the NumPy oracle."""
    from hpc.nbody import compute_accelerations_direct
    x, _, m = _ics().plummer_ic(65536, seed=7)
    a = compute_accelerations_direct(x, m, 0.01)
    rows = np.arange(7, 65536, 257)
    ref = oracle_mod.accel_rows_numpy(x, m, rows, 0.01)
    assert rel_rows(a[rows], ref).max() < F64_TOL


def test_two_body_known_answer():
    """a = G m / (r^2 + eps^2)^(3/2) * d, and the zero-softening variant skips i == j."""
    from hpc.nbody import G, compute_accelerations_direct
    x = np.array([[0.0, 0.0, 0.0], [3.0, 4.0, 0.0]])
    m = np.array([2.0e10, 5.0e10])
    for eps in (0.0, 1e-9, 0.5):
        a = compute_accelerations_direct(x, m, eps)
        r3 = (25.0 + eps * eps) ** 1.5
        exp0 = G * m[1] * np.array([3.0, 4.0, 0.0]) / r3
        exp1 = -G * m[0] * np.array([3.0, 4.0, 0.0]) / r3
        assert np.allclose(a[0], exp0, rtol=1e-13, atol=0) and np.allclose(a[1], exp1, rtol=1e-13, atol=0)
        a32 = compute_accelerations_direct(x, m, eps, dtype="float32")
        assert np.allclose(a32[0], exp0, rtol=2e-6) and np.allclose(a32[1], exp1, rtol=2e-6)


def test_momentum_and_symmetry():
    """sum_i m_i a_i = 0 to rounding; a cube of equal masses exerts no force on its centre."""
    from hpc.nbody import compute_accelerations_direct
    x, _, m = _ics().plummer_ic(4096, seed=3)
    a = compute_accelerations_direct(x, m, 0.01)
    p = (m[:, None] * a).sum(axis=0)
    assert np.abs(p).max() < 1e-12 * np.abs(m[:, None] * a).sum()
    cube = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)] + [[0, 0, 0]], dtype=float)
    a = compute_accelerations_direct(cube, np.full(9, 1e11), 1e-9)
    assert np.abs(a[8]).max() < 1e-14 * np.abs(a[:8]).max()


def test_solar_system_factory(golden):
    from hpc.nbody import NBodySimulator
    g = golden("solar_system.npz")
    np.random.seed(0)
    sim = NBodySimulator.create_solar_system()
    assert np.array_equal(sim.positions, g["positions"]) and np.array_equal(sim.masses, g["masses"])
    assert rel_rows(sim.accelerations, g["accelerations"]).max() < F64_TOL
    earth = np.linalg.norm(sim.accelerations[3])
    assert 5.8e-3 < earth < 6.0e-3


def test_slab_rows_bitwise_equal_full(engine):
    """An i-slab evaluated alone gives the same bits as the same rows of a full evaluation (the
    invariant the sharded mode rests on), in both precisions."""
    import torch
    x, _, m = _ics().plummer_ic(5000, seed=7)
    for dtype in (np.float64, np.float32):
        pos_d = engine.to_device(x)
        m_d, f32 = engine._masses_dev(m)
        stream = engine.pack(pos_d, m_d, f32, 5000, dtype)
        full = engine.accel_slab(stream, 5000, 0, 5000, 0.01)
        for i0, n_i in ((0, 1250), (1250, 1250), (3750, 1250), (4999, 1), (100, 333)):
            part = engine.accel_slab(stream, 5000, i0, n_i, 0.01)
            assert torch.equal(part, full[i0:i0 + n_i])


def test_pack_unpack_roundtrip(engine):
    x, _, m = _ics().reference_default_ic(333, 1)
    pos_d = engine.to_device(x)
    m_d, f32 = engine._masses_dev(m)
    s64 = engine.pack(pos_d, m_d, f32, 333, np.float64)
    assert np.array_equal(engine.unpack(s64, 333).cpu().numpy(), x)
    s32 = engine.pack(pos_d, m_d, f32, 333, np.float32)
    assert np.array_equal(engine.unpack(s32, 333).cpu().numpy(), x.astype(np.float32).astype(np.float64))


def test_host_buffer_abi_accel(oracle_mod):
    """nbh_accel_direct: the C ABI with host pointers, no torch involved in the call."""
    from hpc import _cuda
    lib = _cuda.load_library()
    x, _, m = _ics().reference_default_ic(200, 42)
    m32 = _ics().shared_masses(200)
    out = np.zeros((200, 3))
    for use_f32, tol in ((0, F64_TOL), (1, None)):
        rc = lib.nbh_accel_direct(x.ctypes.data, m32.ctypes.data, 1, 200, 1e-9, use_f32, out.ctypes.data)
        assert rc == 0, lib.nb_last_error()
        ref = oracle_mod.accel_direct(x, m32)
        if tol:
            assert rel_rows(out, ref).max() < tol
        else:
            assert np.abs(out - ref).max() / np.abs(ref).max() < F32_TOL
    rc = lib.nbh_accel_direct(x.ctypes.data, m32.ctypes.data, 1, 0, 1e-9, 0, out.ctypes.data)
    assert rc == 1 and b"bad argument" in lib.nb_last_error()


def test_energy_vs_oracle(oracle_mod, golden):
    from hpc.nbody import compute_total_energy
    ics = _ics()
    x, v, _ = ics.reference_default_ic(200, 42)
    m32 = ics.shared_masses(200)
    e = compute_total_energy(x, v, m32)
    ref = golden("accel_default_n200.npz")["energy_f32mass_seed42"]
    assert np.allclose(e, ref, rtol=1e-12)
    x, v, m = ics.plummer_ic(4096, seed=7)
    e = compute_total_energy(x, v, m, 0.01)
    ref = oracle_mod.total_energy(x, v, m, 0.01, parallel=True)
    assert np.allclose(e, ref, rtol=1e-12)
    # virial sanity of the Plummer sampler itself: 2K/|U| close to 1
    assert 0.9 < 2 * e[0] / abs(e[1]) < 1.1
