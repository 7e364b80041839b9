"""Right-hand sides are independent linear systems (columns of B, ShiftedLaplacianMultigridSolver.jl:46-47):
they are sharded over ranks / devices as contiguous column ranges with no data-path collective."""
from __future__ import annotations


def column_range(nrhs: int, nparts: int, part: int):
    """Contiguous range [c0, c1) of columns owned by `part` of `nparts` (same rule as the C library)."""
    if nparts < 1 or not (0 <= part < nparts):
        raise ValueError("bad partition")
    base, rem = divmod(int(nrhs), int(nparts))
    c0 = part * base + min(part, rem)
    c1 = c0 + base + (1 if part < rem else 0)
    return c0, c1


def all_ranges(nrhs: int, nparts: int):
    return [column_range(nrhs, nparts, p) for p in range(nparts)]


# ---- slab decomposition of ONE grid over the ranks (one process per GPU; include/helmholtz_b200.h hh_create_slab_nccl)
def broadcast_unique_id(make_id, src=0):
    """Rank `src` creates the 128-byte NCCL communicator id (make_id = api.slabUniqueId), every rank receives it.
    Uses the process group the launcher set up (torch.distributed, gloo or nccl)."""
    import torch.distributed as dist

    box = [make_id() if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def nccl_slabs():
    """`slabs` entry of a solver for the calling rank of the default process group."""
    import torch.distributed as dist

    from . import api

    return {"mode": "nccl", "rank": dist.get_rank(), "nranks": dist.get_world_size(),
            "unique_id": broadcast_unique_id(api.slabUniqueId)}


def gather_planes(X_local, planes, nodes, dst=0):
    """Assemble the whole-grid block from the per-rank blocks of owned planes (verification / output only, not part of
    the solve).  X_local: (n1*n2*(k1-k0)) x nrhs column-major numpy block of this rank, planes = (k0, k1).
    Returns the N x nrhs block on rank `dst`, None elsewhere."""
    import numpy as np
    import torch.distributed as dist

    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object((tuple(planes), np.asarray(X_local)), parts, dst=dst)
    if parts is None:
        return None
    plane = int(nodes[0]) * int(nodes[1])
    nrhs = parts[0][1].shape[1] if parts[0][1].ndim == 2 else 1
    X = np.zeros((plane * int(nodes[2]), nrhs), dtype=parts[0][1].dtype, order="F")
    seen = np.zeros(int(nodes[2]), dtype=np.int64)
    for (k0, k1), blk in parts:
        X[plane * k0:plane * k1, :] = np.asarray(blk).reshape((plane * (k1 - k0), nrhs), order="F")
        seen[k0:k1] += 1
    if not np.all(seen == 1):
        raise ValueError("the ranks' plane ranges do not tile the grid")
    return X


# ---- host placement: keep a rank's pinned staging buffers on the NUMA node its GPU hangs off -----------------------------
def bind_to_gpu_numa_node(device_index: int):
    """Restrict the calling process to the CPUs of the NUMA node the GPU `device_index` is attached to (Linux sysfs), so
    that pinned host buffers allocated afterwards are first-touched on that node and PCIe copies do not cross the
    inter-socket link.  Call before allocating host buffers.  Returns {"gpu_numa_node", "cpus"} or None when the
    topology is not exposed (no NUMA information, non-Linux, restricted sysfs) -- placement is then left to the OS."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"gpu_numa_node": node, "cpus": len(cpus)}
    except Exception:
        return None
