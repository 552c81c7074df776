#!/usr/bin/env python3
"""Run under torchrun on P GPUs: the i-slab sharded run (both exchange modes: NCCL all-gather and the fused
peer-store step) must be bit-identical to the one-GPU run; prints per-step times of both modes."""
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
cases = ((10000, np.float64, 6), (40000, np.float32, 6), (65536, np.float32, 8), (262144, np.float32, 3))
for n, dtype, steps in cases:
    x, v, m = ics.plummer_ic(n, seed=7)
    one = None
    for exchange in ("nccl", "peer"):
        try:
            sh = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=local, world=world, rank=rank,
                               exchange=exchange)
        except Exception as e:
            if rank == 0:
                print(f"N={n} exchange={exchange}: unavailable: {e!r}", flush=True)
            ok = False
            continue
        sh.advance(steps)
        sh.advance(2)                      # a second call exercises the hand-over between advance() calls
        pos, vel, acc, e = sh.positions(), sh.velocities(), sh.accelerations(), sh.energy()
        # timing: 10 more steps
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sh.advance(10); e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            if one is None:
                one = ShardedSystem(x, v, m, dt=1e-3, softening=0.01, dtype=dtype, device=local)
                one.advance(steps); one.advance(2)
                ref = (one.positions(), one.velocities(), one.accelerations(), one.energy())
            same = np.array_equal(pos, ref[0]) and np.array_equal(vel, ref[1]) and np.array_equal(acc, ref[2])
            print(f"N={n} {np.dtype(dtype).name} P={world} exchange={sh.exchange}: bitwise identical to 1 GPU: {same}; "
                  f"energy {e[2]:.12e} vs {ref[3][2]:.12e}; {ms.item():.4f} ms/step", flush=True)
            ok &= same and abs(e[2] - ref[3][2]) <= 1e-12 * abs(ref[3][2]) and sh.exchange == exchange
        dist.barrier()
        del sh
dist.destroy_process_group()
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)
