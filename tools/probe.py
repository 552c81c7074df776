#!/usr/bin/env python3
"""Quick device-side timings of every kernel family (development aid; bench.py is the contract)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import _cuda, ics  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    eng = _cuda.get_engine()
    out = {"sm_count": eng.sm_count, "clock_khz": eng.sm_clock_khz}
    for mode in ("ffma", "ffma2", "dfma"):
        out[f"peak_{mode}_tflops"] = round(eng.fma_peak_tflops(mode), 2)
    print(json.dumps(out), flush=True)
    sizes = [int(a) for a in sys.argv[1:]] or [16384, 65536, 262144]
    for n in sizes:
        x, v, m = ics.plummer_ic(n, seed=7)
        pos_d = eng.to_device(x)
        m_d, f32 = eng._masses_dev(m)
        for dtype in (np.float32, np.float64):
            if dtype == np.float64 and n > 300000:
                continue
            stream = eng.pack(pos_d, m_d, f32, n, dtype)
            ws = eng.workspace(n, n, dtype)
            best, med = timed(lambda: eng.accel_slab(stream, n, 0, n, 0.01, ws), reps=5 if n < 500000 else 2,
                              warm=2 if n < 500000 else 1)
            inter = n * (n - 1.0)
            print(json.dumps({"kernel": "accel", "n": n, "dtype": np.dtype(dtype).name, "ms_best": round(best, 4),
                              "ms_median": round(med, 4), "Ginter_per_s": round(inter / best / 1e6, 1),
                              "seg_plan": eng.segment_plan(n)}), flush=True)
    # ensemble, datagen configuration
    for B in (296, 300, 1200, 300, 450):
        x0, v0, m32 = ics.datagen_ensemble_ic(B, 200, seed=42)
        x, v = eng.to_device(x0), eng.to_device(v0)
        a = torch.zeros_like(x)
        m_d, f32 = eng._masses_dev(m32)
        ox = torch.empty((B, 401, 200, 3), dtype=torch.float64, device=eng.device)
        ov, oa = torch.empty_like(ox), torch.empty_like(ox)
        for dtype in (np.float64, np.float32):
            def go():
                x.copy_(torch.from_numpy(x0)); v.copy_(torch.from_numpy(v0))
            def run():
                eng.ensemble_device(x, v, a, m_d, f32, 0, B, 200, 1e-3, 1e-9, 400, 1, dtype, True, True, ox, ov, oa, 401, 0)
            go()
            best, med = timed(run, reps=3, warm=1)
            inter = B * 400 * 200 * 199.0
            print(json.dumps({"kernel": "ensemble", "B": B, "dtype": np.dtype(dtype).name, "ms_best": round(best, 3),
                              "ms_median": round(med, 3), "Ginter_per_s": round(inter / best / 1e6, 1),
                              "sim_steps_per_s": round(B * 400 / best * 1e3), "GBps_snap": round(B * 401 * 200 * 72 / best / 1e6, 1)}), flush=True)
        del ox, ov, oa
    # e2e ensemble through the public API
    from hpc.ensemble import simulate_ensemble
    x0, v0, m32 = ics.datagen_ensemble_ic(300, 200, seed=42)
    for _ in range(3):
        t0 = time.perf_counter()
        out = simulate_ensemble(x0, v0, m32, dt=1e-3, n_steps=400)
        t1 = time.perf_counter()
        print(json.dumps({"e2e_ensemble_ms": round((t1 - t0) * 1e3, 2)}), flush=True)
        del out


if __name__ == "__main__":
    main()
