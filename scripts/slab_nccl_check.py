"""One process per GPU (torchrun): slab-decomposed solve over NCCL against the whole-grid solve of the same library.
Every rank also solves the whole (small) grid on its own GPU and compares its planes.  Prints SLAB_NCCL_OK on rank 0.
Usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/slab_nccl_check.py [n]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = graft.load_package()
    cfg = pkg.workloads.config4(n=n, sigma=max(2.0, n / 32.0), seed=7, pad=max(4, n // 16))
    mesh = pkg.getRegularMesh(cfg["domain"], cfg["n_cells"])
    m = cfg["m"]
    w = pkg.getMaximalFrequency(m, mesh)
    nodes = mesh.n + 1
    gamma = cfg["gamma0_frac"] * w * np.ones(m.shape) + pkg.getABL(nodes, True, cfg["pad"], w)
    srcs = pkg.workloads.point_sources_top_grid(nodes, 2, 2)
    N = int(np.prod(nodes))
    B = np.zeros((N, len(srcs)), dtype=np.complex128, order="F")
    for c, s in enumerate(srcs):
        B[pkg.loc2cs(nodes, s) - 1, c] = 1.0 / mesh.h[0] ** 2
    ok = True
    for krylov, inner, prec, tol, cmp_tol in (("GMRES", 5, np.complex128, 1e-8, 1e-9), ("BiCGSTAB", 0, np.complex128, 1e-8, 1e-9),
                                                ("GMRES", 5, np.complex64, 1e-5, 2e-4)):
        def solver(slabs):
            MG = pkg.getMGparam(prec, pkg.Int64, 3, 1, 60, tol, "Jac", 0.8, 1, 2, "W", "GMRES")
            MG.coarseIters = 10
            hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
            A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, krylov, inner)
            A.slabs = slabs
            A.devices = [local]
            return A
        ref = solver(None)
        Xr, ref = pkg.solveLinearSystem(None, B.astype(prec), ref)
        pkg.clear(ref.MG)
        A = solver(pkg.sharding.nccl_slabs())
        k0, k1 = pkg.slabPlanes(A)
        plane = int(nodes[0] * nodes[1])
        Bl = np.asfortranarray(B[plane * k0:plane * k1, :].astype(prec))
        Xl, A = pkg.solveLinearSystem(None, Bl, A)
        err = np.linalg.norm(Xl - Xr[plane * k0:plane * k1, :]) / np.linalg.norm(Xr)
        same_iters = np.array_equal(A.iterations, ref.iterations) if prec == np.complex128 else \
            np.abs(A.iterations.astype(int) - ref.iterations.astype(int)).max() <= 1
        # point sources through the library's own scatter, and the assembled solution on rank 0
        Xp, A = pkg.solvePointSources(A, srcs, amplitudes=np.full(len(srcs), 1.0 / mesh.h[0] ** 2))
        errp = np.linalg.norm(Xp - Xr[plane * k0:plane * k1, :]) / np.linalg.norm(Xr)
        Xall = pkg.sharding.gather_planes(Xl, (k0, k1), nodes)
        if rank == 0:
            errg = np.linalg.norm(Xall - Xr) / np.linalg.norm(Xr)
            ok = ok and errg < cmp_tol
        good = bool(err < cmp_tol and errp < cmp_tol and same_iters)
        print(f"[rank {rank}/{world}] {krylov} {np.dtype(prec).name} planes [{k0},{k1}) iterations {A.iterations.tolist()} "
              f"ref {ref.iterations.tolist()} err {err:.2e} err_point {errp:.2e} {'ok' if good else 'FAIL'}", flush=True)
        ok = ok and good
        pkg.clear(A.MG)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SLAB_NCCL_OK" if int(flag.item()) == 1 else "SLAB_NCCL_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
