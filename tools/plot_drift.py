#!/usr/bin/env python3
"""Overlay the relative energy-drift curves of profiles/r01_energy_drift_n*.json (float32, float64, CPU oracle where
it was run) in one SVG, one panel per system size.  No plotting library in the image: the SVG is written by hand.

    python tools/plot_drift.py            # -> profiles/r01_energy_drift.svg
    python tools/plot_drift.py --files profiles/r02_energy_drift_ensemble*.json --out profiles/r02_energy_drift_ensemble.svg

Files written by tools/energy_drift_ensemble.py carry, per precision, the median and the maximum over the systems of the
ensemble: both are drawn (the maximum dashed).
"""
import argparse
import glob
import json
import math
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
COLORS = {"f64": "#1f77b4", "f32": "#d62728", "cpu_oracle_f64": "#2ca02c", "f64_max": "#1f77b4", "f32_max": "#d62728"}
LABELS = {"f64": "GPU float64", "f32": "GPU float32", "cpu_oracle_f64": "CPU oracle float64",
          "f64_max": "GPU float64, worst system", "f32_max": "GPU float32, worst system"}
W, H, PAD_L, PAD_R, PAD_T, PAD_B = 520, 300, 70, 20, 40, 45


def panel(d, ox, oy):
    steps = d["steps"]
    curves = {k: [abs(v) for v in c["rel_drift"]] for k, c in d["curves"].items()}
    if "systems" in d:   # an ensemble: the curve is the median over its systems; add the worst system
        LABELS.update({"f64": f"GPU float64, median of {d['systems']} systems", "f32": f"GPU float32, median of {d['systems']} systems"})
        curves.update({k + "_max": [abs(v) for v in c["rel_drift_max"]] for k, c in d["curves"].items() if "rel_drift_max" in c})
    vmax = max(max(v) for v in curves.values()) or 1e-16
    lo, hi = math.floor(math.log10(max(vmax * 1e-4, 1e-17))), math.ceil(math.log10(vmax))
    x0, x1, y0, y1 = ox + PAD_L, ox + W - PAD_R, oy + PAD_T, oy + H - PAD_B

    def X(s):
        return x0 + (x1 - x0) * (s - steps[0]) / max(steps[-1] - steps[0], 1)

    def Y(v):
        lv = math.log10(max(v, 10.0 ** lo))
        return y1 - (y1 - y0) * (lv - lo) / max(hi - lo, 1)
    out = [f'<rect x="{x0}" y="{y0}" width="{x1 - x0}" height="{y1 - y0}" fill="none" stroke="#444"/>']
    gpus = d.get("gpus", 1)
    what = f'{d["systems"]} x N = {d["N"]:,}' if "systems" in d else f'N = {d["N"]:,}'
    out.append(f'<text x="{ox + W / 2}" y="{oy + 22}" text-anchor="middle" font-size="14">{what}, {d["n_steps"]} steps, '
               f'dt = {d["dt"]}, eps = {d["softening"]}, {gpus} GPU{"s" if gpus != 1 else ""}</text>')
    for e in range(lo, hi + 1):
        y = Y(10.0 ** e)
        out.append(f'<line x1="{x0}" y1="{y:.1f}" x2="{x1}" y2="{y:.1f}" stroke="#ddd"/>')
        out.append(f'<text x="{x0 - 6}" y="{y + 4:.1f}" text-anchor="end" font-size="11">1e{e}</text>')
    for s in (steps[0], steps[len(steps) // 2], steps[-1]):
        out.append(f'<text x="{X(s):.1f}" y="{y1 + 16}" text-anchor="middle" font-size="11">{s}</text>')
    out.append(f'<text x="{(x0 + x1) / 2}" y="{y1 + 34}" text-anchor="middle" font-size="12">step</text>')
    out.append(f'<text x="{ox + 14}" y="{(y0 + y1) / 2}" text-anchor="middle" font-size="12" '
               f'transform="rotate(-90 {ox + 14} {(y0 + y1) / 2})">|E - E0| / |E0|</text>')
    for i, (k, v) in enumerate(curves.items()):
        pts = " ".join(f"{X(s):.1f},{Y(val):.1f}" for s, val in zip(steps, v) if s > steps[0])
        dash = ' stroke-dasharray="5,3"' if (k == "cpu_oracle_f64" or k.endswith("_max")) else ""
        out.append(f'<polyline points="{pts}" fill="none" stroke="{COLORS.get(k, "#000")}" stroke-width="1.6"{dash}/>')
        out.append(f'<text x="{x0 + 8}" y="{y0 + 16 + 14 * i}" font-size="11" fill="{COLORS.get(k, "#000")}">'
                   f'{LABELS.get(k, k)} (max {max(v):.2e})</text>')
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", nargs="*", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    files = a.files or sorted(glob.glob(str(ROOT / "profiles" / "r01_energy_drift_n*.json")), key=lambda p: json.load(open(p))["N"])
    cols = 2
    rows = (len(files) + cols - 1) // cols
    parts = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{cols * W}" height="{rows * H}" font-family="sans-serif">',
             f'<rect width="{cols * W}" height="{rows * H}" fill="white"/>']
    for i, f in enumerate(files):
        parts.append(panel(json.load(open(f)), (i % cols) * W, (i // cols) * H))
    parts.append("</svg>")
    out = Path(a.out) if a.out else ROOT / "profiles" / "r01_energy_drift.svg"
    out.write_text("\n".join(parts) + "\n")
    print(out)


if __name__ == "__main__":
    main()
