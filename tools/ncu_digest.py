#!/usr/bin/env python3
"""Digest an .ncu-rep (ncu --set full) into profiles/<tag>_ncu_full_summary.csv and profiles/summary.json.

    python tools/ncu_digest.py gpurun_out/prof_r1c.ncu-rep r01c "tools/profile_target.py ens400 f32 f64"
    python tools/ncu_digest.py gpurun_out/prof_r02_raw.csv r02 "..."        # a raw-page CSV exported on the GPU box

Runs here (no GPU needed): `ncu -i <rep> --page raw --csv` is parsed, one row per captured launch with the columns
the design document quotes, and a per-kernel digest (the last captured launch of each kernel) is written to
profiles/summary.json, which bench.py reads for `roofline.traffic`.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
COLS = [
    "gpu__time_duration.sum", "smsp__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    rep, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    if rep.endswith(".csv"):   # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): reports can exceed
        raw = Path(rep).read_text()                                          # what gpurun copies back
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    raw = raw[raw.index('"ID"'):] if '"ID"' in raw else raw
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    out = io.StringIO()
    w = csv.writer(out)
    cols = [c for c in COLS if c in hdr]
    w.writerow(["Kernel Name", "Block Size", "Grid Size"] + cols)
    w.writerow(["", "", ""] + [units[c] for c in cols])
    digest = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        w.writerow([d["Kernel Name"], d["Block Size"], d["Grid Size"]] + [d[c] for c in cols])
        key = d["Kernel Name"].split("<")[0].split("(")[0].replace("void ", "").strip()

        def val(name, scale=True):
            x = float(d[name])
            return x * SCALE.get(units[name], 1.0) if scale else x
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        digest[key] = {
            "kernel": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"],
            "duration_ms": val("gpu__time_duration.sum"),
            "dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "pipe_fma_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
            "pipe_fp64_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False),
            "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
            "registers": int(float(d["launch__registers_per_thread"])),
            "sm_clock_ghz": val("smsp__cycles_elapsed.avg.per_second", False),
            "source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full --clock-control none --import-source on, {cmd})",
        }
    (ROOT / "profiles" / f"{tag}_ncu_full_summary.csv").write_text(out.getvalue())
    summary = ROOT / "profiles" / "summary.json"
    merged = json.loads(summary.read_text()) if summary.exists() else {}
    merged.update(digest)  # kernels of other captures are kept
    # the sources these captures were taken on: bench.py compares it with the sources it runs (roofline.traffic_stale)
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "nbody-gnn-hpc_b200" / "csrc").glob("*.cu*")) + [ROOT / "include" / "nbody_b200.h"]:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    merged["_source_hash"] = h.hexdigest()[:16]
    try:
        merged["_git_head"] = subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short", "HEAD"],
                                             stdout=subprocess.PIPE, text=True).stdout.strip()
    except Exception:
        pass
    summary.write_text(json.dumps(merged, indent=1) + "\n")
    print(json.dumps({k: (round(v["duration_ms"], 3), round(v["pipe_fp64_pct"] or v["pipe_fma_pct"], 1)) for k, v in digest.items()}))
    print("source hash", merged["_source_hash"], "(re-digest ALL kernels' captures after any change to csrc/)")


if __name__ == "__main__":
    main()
