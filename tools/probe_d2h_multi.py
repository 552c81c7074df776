#!/usr/bin/env python3
"""Under torchrun on P GPUs: raw device->pinned-host copy bandwidth per rank when all ranks copy at once, with the
default placement and with the process bound to the CPUs of the GPU's NUMA node before the pinned allocation."""
import os, sys, time, glob
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def numa_of_gpu(idx):
    bdf = torch.cuda.get_device_properties(idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(idx), "pci_bus_id") else None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        if isinstance(bdf, bytes):
            bdf = bdf.decode()
    except Exception as e:
        return None, None, repr(e)
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    p = f"/sys/bus/pci/devices/{bdf}"
    try:
        node = int(open(p + "/numa_node").read())
        cpus = open(p + "/local_cpulist").read().strip()
    except Exception as e:
        return None, None, repr(e)
    return node, cpus, bdf


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def measure(tag):
    n = 1 << 30
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(dev); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([4 * n / dt / 1e9], device="cuda")
    allv = [torch.zeros_like(gbs) for _ in range(world)]
    dist.all_gather(allv, gbs)
    if rank == 0:
        v = [round(x.item(), 1) for x in allv]
        print(f"{tag}: per-rank D2H GB/s {v}, aggregate {sum(v):.1f}", flush=True)
    del host, dev
    dist.barrier()


node, cpus, info = numa_of_gpu(local)
allinfo = [None] * world
dist.all_gather_object(allinfo, (local, node, cpus, info, sorted(os.sched_getaffinity(0))[:4], len(os.sched_getaffinity(0))))
if rank == 0:
    print("numa nodes:", sorted(glob.glob("/sys/devices/system/node/node*")), flush=True)
    for a in allinfo:
        print("gpu", a, flush=True)
measure("default placement, all ranks at once")
if cpus:
    try:
        os.sched_setaffinity(0, set(parse_cpulist(cpus)) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
    except Exception as e:
        print("affinity failed", e)
    measure("bound to the GPU's local CPUs before the pinned allocation")
# one rank at a time, for reference
for r in range(min(world, 2)):
    if rank == r:
        n = 1 << 30
        dev = torch.empty(n, dtype=torch.uint8, device="cuda"); host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        host.copy_(dev); torch.cuda.synchronize()
        t0 = time.perf_counter(); host.copy_(dev, non_blocking=True); torch.cuda.synchronize()
        print(f"rank {r} alone: {n / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
    dist.barrier()
dist.destroy_process_group()
