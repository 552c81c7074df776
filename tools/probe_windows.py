"""K5 tuning aid: window kernel time at the data-generation shape for 1..3 CTAs per SM (0 = the library default)."""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
import torch
from hpc import _cuda
import bench
eng = _cuda.get_engine()
g = torch.Generator(device=eng.device).manual_seed(1)
pos = torch.randn((300, 401, 200, 3), dtype=torch.float64, device=eng.device, generator=g)
vel = torch.randn((300, 401, 200, 3), dtype=torch.float64, device=eng.device, generator=g)
for per_sm in (0, 1, 2, 3):
    if per_sm:
        os.environ["NB_WINDOW_CTAS_PER_SM"] = str(per_sm)   # 0: the library's own choice
    r = bench.window_extras(eng, pos, vel)
    print(per_sm, r["ms"], r["achieved_GBps"], r["frac_of_hbm_peak"], r["device_fill_of_the_outputs_GBps"], flush=True)
