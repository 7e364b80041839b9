"""Multi-GPU path = independent right-hand sides sharded over ranks as contiguous column ranges with no
data-path collective.  world_size-2 gloo run on CPU of the host-side partition logic bench.py and the
library (column_range in hh_api.cu) share."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nrhs, out):
    import sys

    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft

    pkg = graft.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c0, c1 = pkg.sharding.column_range(nrhs, world, rank)
    # every rank "solves" only its own columns: here the stand-in solve is X[:, c] = (c + 1) * B[:, c]
    B = torch.arange(1, 5 * nrhs + 1, dtype=torch.float64).reshape(nrhs, 5)
    mine = torch.zeros_like(B)
    for c in range(c0, c1):
        mine[c] = (c + 1) * B[c]
    owned = torch.zeros(nrhs, dtype=torch.int64)
    owned[c0:c1] = 1
    # verification only (not part of the data path): disjoint cover and the assembled result
    dist.all_reduce(owned)
    dist.all_reduce(mine)
    if rank == 0:
        out.put((owned.tolist(), mine.numpy().copy(), B.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nrhs", [7, 16])
def test_rhs_sharding_world2_gloo(nrhs):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nrhs, out)) for r in range(2)]
    for p in procs:
        p.start()
    owned, mine, B = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert owned == [1] * nrhs  # every column owned by exactly one rank
    assert np.array_equal(mine, (np.arange(1, nrhs + 1)[:, None]) * B)


def test_column_range_matches_library_rule(pkg):
    for nrhs in (1, 2, 7, 16, 255, 256):
        for parts in (1, 2, 3, 4, 8):
            r = pkg.sharding.all_ranges(nrhs, parts)
            assert r[0][0] == 0 and r[-1][1] == nrhs
            assert all(r[i][1] == r[i + 1][0] for i in range(parts - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pkg.sharding.column_range(4, 2, 2)


# ---- slab decomposition of one grid over the ranks: host-side plumbing (partition, communicator id, gathering) ----
def _slab_worker(rank, world, port, out):
    import sys

    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft

    pkg = graft.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slabs = pkg.sharding.nccl_slabs()  # rank 0's id reaches every rank
    nodes = (5, 3, 33)
    geo = pkg.slabPartition(nodes[2], 3, world, rank)[0]
    plane = nodes[0] * nodes[1]
    # stand-in for this rank's solution block: the whole-grid node index, two right-hand sides
    k0, k1 = geo["own0"], geo["own1"]
    idx = np.arange(plane * k0, plane * k1, dtype=np.float64)
    Xl = np.asfortranarray(np.stack([idx, -idx], axis=1))
    X = pkg.sharding.gather_planes(Xl, (k0, k1), nodes)
    ids = [None] * world
    dist.all_gather_object(ids, (slabs["rank"], slabs["nranks"], slabs["unique_id"]))
    if rank == 0:
        out.put((ids, X))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_plumbing_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ids, X = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [i[0] for i in ids] == [0, 1] and all(i[1] == 2 for i in ids)
    assert len(ids[0][2]) == 128 and ids[0][2] == ids[1][2]
    N = 5 * 3 * 33
    assert X.shape == (N, 2)
    assert np.array_equal(X[:, 0], np.arange(N)) and np.array_equal(X[:, 1], -np.arange(N))
