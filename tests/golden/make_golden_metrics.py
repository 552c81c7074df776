#!/usr/bin/env python3
"""Generate tests/golden/metrics_reference.npz by EXECUTING the reference's own compute_energy_error and
compute_momentum_error (src/utils/metrics.py:62-137; pure NumPy, imported by file path).

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden_metrics.py

Inputs: trajectories already stored as fixtures (produced by the reference's simulator, make_golden.py) and two small
seeded random stacks with odd / tiny body counts.  Nothing is copied from the reference: the fixture holds our inputs
and the arrays its functions returned for them.
"""
from __future__ import annotations

import importlib.util
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/src/utils/metrics.py")


def main():
    spec = importlib.util.spec_from_file_location("ref_metrics", str(REF))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    g = np.load(HERE / "traj_plummer_n200.npz")
    e, err = ref.compute_energy_error(g["positions"], g["velocities"], g["masses"], softening=float(g["softening"]))
    p, perr = ref.compute_momentum_error(g["velocities"], g["masses"])
    out.update(plummer_energies=e, plummer_energy_error=err, plummer_momentum=p, plummer_momentum_error=perr)
    g = np.load(HERE / "ensemble_default_b4_n200_t20.npz")
    m32 = np.random.RandomState(42).uniform(1e10, 1e12, 200).astype(np.float32)      # generate_data.py:108-109
    for b in range(4):
        e, err = ref.compute_energy_error(g["positions"][b], g["velocities"][b], m32)   # default G and softening
        p, perr = ref.compute_momentum_error(g["velocities"][b], m32)
        out.update({f"default{b}_energies": e, f"default{b}_energy_error": err, f"default{b}_momentum": p,
                    f"default{b}_momentum_error": perr})
    rng = np.random.RandomState(99)
    for tag, (S, N) in {"odd": (6, 7), "two": (3, 2), "one": (2, 1), "n33": (4, 33)}.items():
        pos = rng.standard_normal((S, N, 3)) * 3.0
        vel = rng.standard_normal((S, N, 3))
        m = rng.uniform(1e9, 1e11, N)
        e, err = ref.compute_energy_error(pos, vel, m, G=2.5e-11, softening=0.05)
        p, perr = ref.compute_momentum_error(vel, m)
        out.update({f"{tag}_pos": pos, f"{tag}_vel": vel, f"{tag}_masses": m, f"{tag}_energies": e,
                    f"{tag}_energy_error": err, f"{tag}_momentum": p, f"{tag}_momentum_error": perr})
    np.savez_compressed(HERE / "metrics_reference.npz", **out)
    print({k: np.shape(v) for k, v in out.items() if "energies" in k})


if __name__ == "__main__":
    main()
