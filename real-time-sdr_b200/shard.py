"""Sharding of independent stations over the ranks (GPUs) of one box, and the host gather of their results.

Stations never exchange data: rank r owns a contiguous range of global station indices and all
of their carried state; the only inter-rank traffic is the final gather of small per-station results (and, in bench.py,
the timing barrier).  No collective touches the data path, so the same code runs over gloo (tests) and nccl (rows are
staged through the rank's GPU there).
"""
from __future__ import annotations

import numpy as np


def station_range(rank: int, world: int, total: int) -> range:
    """Contiguous, balanced split of `total` stations: the first total % world ranks get one more."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(station: int, world: int, total: int) -> int:
    base, extra = divmod(total, world)
    edge = extra * (base + 1)
    return station // (base + 1) if station < edge else extra + (station - edge) // max(base, 1)


def gather_rows(local: np.ndarray, rank: int, world: int, total: int, dist=None) -> np.ndarray | None:
    """Host gather of per-station rows ([n_local, ...]) to rank 0 in global station order; None on other ranks."""
    if world == 1 or dist is None:
        return local
    import torch

    counts = [len(station_range(r, world, total)) for r in range(world)]
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    pad = np.zeros((max(counts), width), local.dtype)
    pad[: local.shape[0]] = local.reshape(local.shape[0], width)
    t = torch.from_numpy(pad.view(np.uint8).reshape(-1).copy())
    if dist.get_backend() == "nccl":  # NCCL moves device tensors only: stage the (small) rows through this rank's GPU
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, out, dst=0)
    if rank != 0:
        return None
    rows = [o.cpu().numpy().view(local.dtype).reshape(max(counts), width)[: counts[r]] for r, o in enumerate(out)]
    return np.concatenate(rows).reshape((total,) + local.shape[1:])


def bind_process_to_gpu_node(pci_bus_id: str) -> dict:
    """Best effort: pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers it
    allocates next (first touch) are local to that GPU's PCIe root.  With 8 ranks streaming 50 GB/s each, remote-node
    buffers halve the end-to-end rate.  Returns what was done (for the bench line); never raises."""
    import os

    info = {"numa_node": None, "cpus": None}
    try:
        dev = pci_bus_id.lower()
        if len(dev.split(":")[0]) == 8:  # 00000000:1B:00.0 -> 0000:1b:00.0
            dev = dev[4:]
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        if node < 0:
            return info
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        pass
    return info
