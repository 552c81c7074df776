"""CPU oracle for the direct-sum gravity + leapfrog path -- TEST INFRASTRUCTURE.

This package is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under
``nbody-gnn-hpc_b200/`` imports it, and the product path raises if its CUDA
library is missing instead of falling back to anything here.

Two restatements of the reference algorithm live here:

* ``oracle_force.c`` / ``oracle_strict.c`` -- plain C, loaded through ctypes
  (functions below without a suffix).  Follows
  ``/root/reference/src/hpc/nbody.py:22-66,101-130,202-259`` and the ensemble
  driver ``/root/reference/scripts/generate_data.py:32-58``.
* ``numpy_oracle.py`` -- NumPy float64, row-vectorised, for spot rows at large N.

Parity pin: the reference holds no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``tests/golden/*.npz`` were produced by ``tests/golden/make_golden.py``
executing the reference's Numba functions, and ``tests/test_oracle_golden.py``
holds the oracle to them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

from .numpy_oracle import G, SOFTENING, accel_rows_numpy, step_numpy  # noqa: F401  (re-exported for tests)

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"
_LIB = None
_LEVEL = None

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_long_p = ctypes.POINTER(ctypes.c_long)


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def _pick_level() -> str:
    forced = os.environ.get("NBODY_ORACLE_LEVEL")
    if forced:
        return forced
    flags = _cpu_flags()
    if {"avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"} <= flags:
        return "v4"
    if {"avx2", "fma", "bmi2"} <= flags:
        return "v3"
    return "generic"


def build(force: bool = False) -> None:
    """Compile the C oracle (gcc + make; no network, no reference sources)."""
    targets = [_BUILD / f"liboracle_{lv}.so" for lv in ("generic", "v3", "v4")]
    if not force and all(t.exists() for t in targets):
        newest_src = max((_HERE / s).stat().st_mtime for s in ("oracle_force.c", "oracle_strict.c", "Makefile"))
        if all(t.stat().st_mtime >= newest_src for t in targets):
            return
    subprocess.run(["make", "-C", str(_HERE), "-B", "all"], check=True, capture_output=True)


def lib() -> ctypes.CDLL:
    global _LIB, _LEVEL
    if _LIB is not None:
        return _LIB
    level = _pick_level()
    path = _BUILD / f"liboracle_{level}.so"
    if not path.exists():
        build()
    L = ctypes.CDLL(str(path))
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    L.oracle_accel_direct.argtypes = [vp, vp, ci, ci, cd, vp]
    L.oracle_accel_direct_rows.argtypes = [vp, vp, ci, ci, cd, ci, ci, vp]
    L.oracle_accel_direct_serial.argtypes = [vp, vp, ci, ci, cd, vp]
    L.oracle_accel_direct_strict.argtypes = [vp, vp, ci, ci, cd, ci, vp]
    L.oracle_total_energy.argtypes = [vp, vp, vp, ci, ci, cd, vp]
    L.oracle_total_energy_parallel.argtypes = [vp, vp, vp, ci, ci, cd, vp]
    L.oracle_step.argtypes = [vp, vp, vp, vp, ci, ci, cd, cd, ci]
    L.oracle_run.argtypes = [vp, vp, vp, vp, ci, ci, cd, cd, ci, ci, ci, cd, ctypes.c_long,
                             vp, vp, vp, vp, vp, vp, vp]
    L.oracle_ensemble_run.argtypes = [vp, vp, vp, ci, ci, ci, cd, cd, ci, ci, vp, vp, vp, vp]
    L.oracle_num_threads.restype = ci
    L.oracle_set_num_threads.argtypes = [ci]
    L.oracle_set_num_threads.restype = None
    for name in ("oracle_accel_direct", "oracle_accel_direct_rows", "oracle_accel_direct_serial",
                 "oracle_accel_direct_strict", "oracle_total_energy", "oracle_total_energy_parallel",
                 "oracle_step", "oracle_run", "oracle_ensemble_run"):
        getattr(L, name).restype = None
    _LIB, _LEVEL = L, level
    return L


def isa_level() -> str:
    lib()
    return _LEVEL


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def use_all_cores() -> int:
    """Use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().oracle_set_num_threads(n)
    return num_threads()


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _masses(m):
    m = np.asarray(m)
    if m.dtype == np.float32:
        return np.ascontiguousarray(m), 1
    return np.ascontiguousarray(m, dtype=np.float64), 0


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


FORCE_MODES = {"fast": 0, "fast_serial": 1, "strict": 2, "strict_reversed": 3}


def accel_direct(positions, masses, softening: float = SOFTENING, mode: str = "fast") -> np.ndarray:
    """compute_accelerations_direct, nbody.py:22-66.  mode selects the rounding model."""
    pos = _f64(positions)
    m, f32 = _masses(masses)
    n = pos.shape[0]
    acc = np.zeros_like(pos)
    L = lib()
    if mode == "fast":
        L.oracle_accel_direct(_ptr(pos), _ptr(m), f32, n, float(softening), _ptr(acc))
    elif mode == "fast_serial":
        L.oracle_accel_direct_serial(_ptr(pos), _ptr(m), f32, n, float(softening), _ptr(acc))
    elif mode in ("strict", "strict_reversed"):
        L.oracle_accel_direct_strict(_ptr(pos), _ptr(m), f32, n, float(softening),
                                     int(mode == "strict_reversed"), _ptr(acc))
    else:
        raise ValueError(mode)
    return acc


def accel_direct_rows(positions, masses, i0: int, n_i: int, softening: float = SOFTENING) -> np.ndarray:
    """Rows [i0, i0+n_i) of the acceleration array (the i-slab one rank owns)."""
    pos = _f64(positions)
    m, f32 = _masses(masses)
    acc = np.zeros_like(pos)
    lib().oracle_accel_direct_rows(_ptr(pos), _ptr(m), f32, pos.shape[0], float(softening), i0, n_i, _ptr(acc))
    return acc[i0:i0 + n_i].copy()


def total_energy(positions, velocities, masses, softening: float = SOFTENING, parallel: bool = False):
    """compute_total_energy, nbody.py:101-130 -> (kinetic, potential, total)."""
    pos, vel = _f64(positions), _f64(velocities)
    m, f32 = _masses(masses)
    out = np.zeros(3)
    fn = lib().oracle_total_energy_parallel if parallel else lib().oracle_total_energy
    fn(_ptr(pos), _ptr(vel), _ptr(m), f32, pos.shape[0], float(softening), _ptr(out))
    return float(out[0]), float(out[1]), float(out[2])


def run(positions, velocities, accelerations, masses, dt: float, softening: float, n_steps: int,
        save_interval: int = 1, mode: str = "fast", time0: float = 0.0, step0: int = 0) -> dict:
    """NBodySimulator.run, nbody.py:220-248, from an explicit (x, v, a) state.

    Returns the stacked snapshots (what generate_data.py:51-58 builds from the
    state list) plus the final live state.
    """
    pos, vel, acc = _f64(positions).copy(), _f64(velocities).copy(), _f64(accelerations).copy()
    m, f32 = _masses(masses)
    n = pos.shape[0]
    n_snap = 1 + n_steps // save_interval
    out_p = np.empty((n_snap, n, 3))
    out_v = np.empty((n_snap, n, 3))
    out_a = np.empty((n_snap, n, 3))
    out_t = np.empty(n_snap)
    out_s = np.empty(n_snap, dtype=np.int64)
    t_fin = ctypes.c_double(0.0)
    s_fin = ctypes.c_long(0)
    lib().oracle_run(_ptr(pos), _ptr(vel), _ptr(acc), _ptr(m), f32, n, float(dt), float(softening),
                     int(n_steps), int(save_interval), FORCE_MODES[mode], float(time0), int(step0),
                     _ptr(out_p), _ptr(out_v), _ptr(out_a), _ptr(out_t), _ptr(out_s),
                     ctypes.addressof(t_fin), ctypes.addressof(s_fin))
    return {"positions": out_p, "velocities": out_v, "accelerations": out_a, "times": out_t,
            "steps": out_s, "final_positions": pos, "final_velocities": vel,
            "final_accelerations": acc, "final_time": t_fin.value, "final_step": s_fin.value}


def ensemble_run(x0, v0, masses, dt: float, softening: float, n_steps: int, save_interval: int = 1,
                 outputs: bool = True) -> dict:
    """B independent simulations with shared masses, generate_data.py:32-58,142-149."""
    x0, v0 = _f64(x0), _f64(v0)
    m, f32 = _masses(masses)
    B, n = x0.shape[0], x0.shape[1]
    n_snap = 1 + n_steps // save_interval
    scratch = np.empty((B, 9 * n))
    if outputs:
        out_p = np.empty((B, n_snap, n, 3))
        out_v = np.empty((B, n_snap, n, 3))
        out_a = np.empty((B, n_snap, n, 3))
    else:
        out_p = out_v = out_a = None
    lib().oracle_ensemble_run(_ptr(x0), _ptr(v0), _ptr(m), f32, B, n, float(dt), float(softening),
                              int(n_steps), int(save_interval), _ptr(out_p), _ptr(out_v), _ptr(out_a),
                              _ptr(scratch))
    final = scratch.reshape(B, 3, n, 3)
    return {"positions": out_p, "velocities": out_v, "accelerations": out_a,
            "final_positions": final[:, 0].copy(), "final_velocities": final[:, 1].copy(),
            "final_accelerations": final[:, 2].copy()}
