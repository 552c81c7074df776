#!/usr/bin/env python3
"""Run under torchrun on P GPUs: the i-slab sharded run must be bit-identical to the one-GPU run."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nbody-gnn-hpc_b200"))
from hpc import ics  # noqa: E402
from hpc.sharded import ShardedSystem  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for n, dtype, steps in ((10000, np.float64, 6), (40000, np.float32, 6), (262144, np.float32, 3)):
    x, v, m = ics.plummer_ic(n, seed=7)
    sh = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=local, world=world, rank=rank)
    sh.advance(steps)
    pos, vel, acc, e = sh.positions(), sh.velocities(), sh.accelerations(), sh.energy()
    if rank == 0:
        one = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=local)
        one.advance(steps)
        same = (np.array_equal(pos, one.positions()) and np.array_equal(vel, one.velocities())
                and np.array_equal(acc, one.accelerations()))
        e1 = one.energy()
        print(f"N={n} {np.dtype(dtype).name} P={world}: bitwise identical to 1 GPU: {same}; "
              f"energy {e[2]:.12e} vs {e1[2]:.12e}", flush=True)
        ok &= same and abs(e[2] - e1[2]) <= 1e-12 * abs(e1[2])
    dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)
