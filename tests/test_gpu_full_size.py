"""BASELINE.json's full sizes through size-independent properties (the oracle would need minutes to hours there):
configs[3] N = 262,144 and configs[4] N = 1,048,576 float32 -- total momentum, spot rows against the NumPy oracle,
an i-slab equal to the same rows of the full evaluation bit for bit, one leapfrog step equal to its definition --
and configs[1], the 300 x 200 x 400 float64 ensemble -- first row = input, last row = final state, per-system
momentum, the time axis, three systems against the oracle over the first 48 steps."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel_rows(a, ref):
    return np.linalg.norm(a - ref, axis=1) / np.linalg.norm(ref, axis=1)


@pytest.mark.parametrize("n,eps", [(262144, 1e-3), (1048576, 1e-3)])
def test_large_system_f32_properties(engine, oracle_mod, n, eps):
    import torch
    from hpc import ics
    x, v, m = ics.plummer_ic(n, seed=7)
    pos_d = engine.to_device(x)
    m_d, f32 = engine._masses_dev(m)
    stream = engine.pack(pos_d, m_d, f32, n, np.float32)
    ws = engine.workspace(n, n, np.float32)
    acc = engine.accel_slab(stream, n, 0, n, eps, ws)
    a = acc.double().cpu().numpy()
    # sum_i m_i a_i = 0 (every pair cancels): float32 rounding of ~n terms per body
    ma = m[:, None] * a
    assert np.abs(ma.sum(axis=0)).max() < 2e-6 * np.abs(ma).sum()
    # spot rows against the float64 NumPy oracle (reference nbody.py:41-64)
    rows = np.arange(11, n, n // 16)
    ref = oracle_mod.accel_rows_numpy(x, m, rows, eps)
    assert (np.abs(a[rows] - ref).max() / np.abs(ref).max()) < 1e-5
    assert np.median(_rel_rows(a[rows], ref)) < 5e-6
    # the slab a rank of 8 would own, evaluated alone: same bits (the invariant the sharded mode rests on)
    i0, n_i = 5 * (n // 8), n // 8
    part = engine.accel_slab(stream, n, i0, n_i, eps)
    assert torch.equal(part, acc[i0:i0 + n_i])
    del part
    # one fused leapfrog step == its definition applied to the accelerations above (nbody.py:202-218), in float32
    vel = engine.to_device(v, torch.float32)
    nxt = stream.clone()
    a0 = acc.clone()
    dt = np.float32(1e-3)
    half = np.float32(0.5 * 1e-3)
    engine.kick_drift_slab(stream, nxt, vel, a0, n, 0, n, 1e-3)
    v_half = (v.astype(np.float32) + half * a.astype(np.float32)).astype(np.float32)
    x_new = (x.astype(np.float32) + dt * v_half).astype(np.float32)
    got = engine.unpack(nxt, n).cpu().numpy()
    assert np.array_equal(got.astype(np.float32), x_new)
    assert np.array_equal(vel.cpu().numpy(), v_half)


def test_datagen_ensemble_full_size_properties(engine, oracle_mod):
    from hpc import ics
    from hpc.ensemble import simulate_ensemble
    B, N, T = 300, 200, 400
    x0, v0, m32 = ics.datagen_ensemble_ic(B, N, seed=42)
    out = simulate_ensemble(x0, v0, m32, dt=1e-3, n_steps=T, save_interval=1)
    pos, vel, acc = out["positions"], out["velocities"], out["accelerations"]
    assert pos.shape == vel.shape == acc.shape == (B, T + 1, N, 3) and pos.dtype == np.float64
    assert np.array_equal(pos[:, 0], x0) and np.array_equal(vel[:, 0], v0)          # get_state() before the loop
    assert np.array_equal(pos[:, -1], out["final_positions"]) and np.array_equal(vel[:, -1], out["final_velocities"])
    assert np.array_equal(acc[:, -1], out["final_accelerations"])
    t, times = 0.0, [0.0]
    for _ in range(T):
        t += 1e-3
        times.append(t)
    assert np.array_equal(out["times"], np.array(times))                              # running sum, nbody.py:217
    # every snapshot of every system: sum_i m_i a_i = 0 to rounding
    m = m32.astype(np.float64)
    ma = m[None, None, :, None] * acc
    assert (np.abs(ma.sum(axis=2)).max(axis=-1) < 1e-11 * np.abs(ma).sum(axis=(2, 3))).all()
    # total momentum of a system is conserved by the leapfrog up to the same rounding, step after step
    p = (m[None, None, :, None] * vel).sum(axis=2)
    scale = np.abs(m[None, None, :, None] * vel).sum(axis=(2, 3))
    assert (np.abs(p - p[:, :1]).max(axis=-1) < 1e-10 * scale).all()
    for b in (0, 149, 299):                                                           # chaotic ICs: first 48 steps
        chk = oracle_mod.run(x0[b], v0[b], oracle_mod.accel_direct(x0[b], m32, 1e-9), m32, 1e-3, 1e-9, 48, 1)
        assert np.abs(pos[b, :49] - chk["positions"]).max() < 1e-8
