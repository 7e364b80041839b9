"""The slab protocol of the multi-GPU solve of ONE grid (csrc/hh_slab.cuh), restated in numpy and
run by 2 and 3 processes over gloo on CPU: every rank keeps the planes hh_slab_partition gives it (halo planes
included), exchanges exactly the planes the library exchanges (the lower halo before a restriction, the upper halo
before an interpolation, both before a stencil apply) and works on its owned rows only.  The results are compared with
the oracle's whole-grid operators (Kronecker-assembled H, P, R = 2^-3 P^T, A_c = R H P).  This pins the host-side design
-- partition geometry, halo sufficiency, who sends what to whom, all-reduced dots -- without a GPU; the CUDA kernels
that follow the same protocol are checked against the whole-grid solve in tests/test_gpu_slab.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

NODES = (7, 5, 17)
LEVELS = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(v, geo, plane, rank, world, lower=True, upper=True):
    """v: local block (plane*nloc, k).  Fill plane zb-1 from the slab below and plane ze from the slab above
    (SlabTransport::exchange: `lower` moves data up, `upper` moves data down)."""
    zb, ze = geo["zb"], geo["ze"]
    ops, recv = [], []
    def pl(z):
        return slice(plane * z, plane * (z + 1))
    if lower:
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(v[pl(ze - 1)])), rank + 1))
        if rank > 0:
            buf = torch.empty((plane, v.shape[1]), dtype=torch.complex128)
            ops.append(dist.P2POp(dist.irecv, buf, rank - 1))
            recv.append((zb - 1, buf))
    if upper:
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.ascontiguousarray(v[pl(zb)])), rank - 1))
        if rank < world - 1:
            buf = torch.empty((plane, v.shape[1]), dtype=torch.complex128)
            ops.append(dist.P2POp(dist.irecv, buf, rank + 1))
            recv.append((ze, buf))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for z, buf in recv:
        v[pl(z)] = buf.numpy()


def _worker(rank, world, port, out):
    import sys

    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft

    pkg = graft.load_package()
    ho = graft.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the same whole-grid problem on every rank (the oracle's assembled operators are the truth)
    nodes = np.array(NODES)
    rng = np.random.default_rng(17)
    mesh = ho.getRegularMesh([0, 0.5, 0, 0.4, 0, 1.6], list(nodes - 1))
    m = 1.0 / (1.5 + rng.random(tuple(nodes))) ** 2
    w = 0.9 * ho.getMaximalFrequency(m, mesh)
    gamma = 0.05 * w * (1 + rng.random(tuple(nodes)))
    H = (ho.GetHelmholtzOperator(mesh, m, w, gamma, True, True) + ho.GetHelmholtzShiftOP(m, w, 0.2)).tocsr()
    P, nc = ho.getFWInterp(nodes)
    R = (P.T * 0.125).tocsr()
    Ac = (R @ H @ P).tocsr()
    N, Nc = int(np.prod(nodes)), int(np.prod(nc))
    k = 2
    x = rng.standard_normal((N, k)) + 1j * rng.standard_normal((N, k))
    geo = pkg.slabPartition(int(nodes[2]), LEVELS, world, rank)
    gf, gc = geo[0], geo[1]
    pf, pc = int(nodes[0] * nodes[1]), int(nc[0] * nc[1])

    def local_rows(g, plane):      # global node indices of the local planes / of the owned planes
        return np.arange(plane * g["koff"], plane * (g["koff"] + g["nloc"])), np.arange(plane * g["own0"], plane * g["own1"])

    locf, ownf = local_rows(gf, pf)
    locc, ownc = local_rows(gc, pc)

    def owned_op(A, own_rows, loc_cols, ncols):
        """rows of the owned nodes; assert they only touch columns this slab holds (halo sufficiency)"""
        sub = A[own_rows]
        touched = np.unique(sub.indices)
        assert np.all(np.isin(touched, loc_cols)), "an owned row reads a plane outside the slab's halo"
        return sub[:, loc_cols]

    def owned_slice(g, plane):
        return slice(plane * g["zb"], plane * g["ze"])

    # ---- stencil apply: both halos
    xl = np.zeros((len(locf), k), dtype=complex)
    xl[owned_slice(gf, pf)] = x[ownf]            # a rank starts with its owned planes only
    _exchange(xl, gf, pf, rank, world)
    y_own = owned_op(H, ownf, locf, N) @ xl
    err_apply = np.abs(y_own - (H @ x)[ownf]).max()
    # ---- restriction of the residual-like vector y: lower halo only
    yl = np.zeros((len(locf), k), dtype=complex)
    yl[owned_slice(gf, pf)] = y_own
    _exchange(yl, gf, pf, rank, world, lower=True, upper=False)
    Rl = owned_op(R, ownc, locf, N)
    if rank < world - 1:  # the upper halo was NOT exchanged: restriction must not read it
        upper = np.arange(pf * gf["ze"], pf * (gf["ze"] + 1))
        assert Rl[:, upper].nnz == 0
    bc_own = Rl @ yl
    err_restrict = np.abs(bc_own - (R @ (H @ x))[ownc]).max()
    # ---- coarse stencil apply (Galerkin operator): both halos
    bl = np.zeros((len(locc), k), dtype=complex)
    bl[owned_slice(gc, pc)] = bc_own
    _exchange(bl, gc, pc, rank, world)
    zc_own = owned_op(Ac, ownc, locc, Nc) @ bl
    err_coarse = np.abs(zc_own - (Ac @ (R @ (H @ x)))[ownc]).max()
    # ---- interpolation: upper halo only
    zl = np.zeros((len(locc), k), dtype=complex)
    zl[owned_slice(gc, pc)] = zc_own
    _exchange(zl, gc, pc, rank, world, lower=False, upper=True)
    Pl = owned_op(P, ownf, locc, Nc)
    if rank > 0:
        lower = np.arange(pc * (gc["zb"] - 1), pc * gc["zb"])
        assert Pl[:, lower].nnz == 0
    xf_own = Pl @ zl
    ref = P @ (Ac @ (R @ (H @ x)))
    err_prolong = np.abs(xf_own - ref[ownf]).max()
    # ---- all-reduced dot products over the owned planes
    d = torch.from_numpy(np.array([np.vdot(x[ownf, c], ref[ownf, c]) for c in range(k)]))
    dist.all_reduce(d)
    err_dot = np.abs(d.numpy() - np.array([np.vdot(x[:, c], ref[:, c]) for c in range(k)])).max()
    errs = torch.tensor([err_apply, err_restrict, err_coarse, err_prolong, err_dot], dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put(errs.tolist())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_protocol_matches_whole_grid_operators(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    errs = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scale = 1e4  # entries of H are O(1/h^2)
    assert max(errs[:4]) < 1e-9 * scale and errs[4] < 1e-6, errs
