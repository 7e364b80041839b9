"""A/B of the round-2 Krylov savings at the headline size (257^3, one config-4 source): iteration counts at 1e-6 / 1e-9 and
the solutions with HH_SKIP_LAST_UPDATE / HH_SMALL_FUSED on (default) and off (the path that tests/test_gpu_headline_parity.py
verified against the CPU port: 29 / 51 iterations).  Also the reference's production multigrid (5 levels, K-cycle,
Jac-GMRES) in both precisions.  Prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 257
    pkg = graft.load_package()
    cfg = pkg.workloads.config4(n=n, sigma=8.0 * (n - 1) / 256, seed=1234, pad=max(4, 16 * (n - 1) // 256))
    mesh = pkg.getRegularMesh(cfg["domain"], cfg["n_cells"])
    m = cfg["m"]
    w = pkg.getMaximalFrequency(m, mesh)
    nodes = np.asarray(mesh.n) + 1
    gamma = cfg["gamma0_frac"] * w * np.ones(m.shape) + pkg.getABL(nodes, True, cfg["pad"], w)
    src = pkg.workloads.point_sources_top_grid(nodes, 16, 16)[100]
    N = int(np.prod(nodes))
    b = np.zeros(N, dtype=np.complex128)
    b[pkg.loc2cs(nodes, src) - 1] = 1.0 / mesh.h[0] ** 2

    def solver(prec, tol, levels=3, relax="Jac", pre=1, post=2, cycle="W", maxit=30):
        MG = pkg.getMGparam(prec, pkg.Int64, levels, 1, maxit, tol, relax, 0.8, pre, post, cycle, "GMRES", coarseIters=10)
        hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
        return pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)

    cases = [("bench_c128", pkg.ComplexF64, dict(), (1e-6, 1e-9)),
             ("bench_c64", pkg.ComplexF32, dict(), (2e-6,)),
             ("production_c128", pkg.ComplexF64, dict(levels=5, relax="Jac-GMRES", pre=lambda l: l + 1, post=lambda l: l + 1,
                                                      cycle="K", maxit=50), (1e-9,)),
             ("production_c64", pkg.ComplexF32, dict(levels=5, relax="Jac-GMRES", pre=lambda l: l + 1, post=lambda l: l + 1,
                                                     cycle="K", maxit=50), (1.8e-6,))]
    for name, prec, kw, tols in cases:
        res = {}
        for flag in ("1", "0"):
            os.environ["HH_SKIP_LAST_UPDATE"] = flag
            os.environ["HH_SMALL_FUSED"] = flag
            if os.environ.get("HH_CHECK_ALL"):  # also the other default-on paths of this round against their plain forms
                for k in ("HH_SCALAR_FAST", "HH_FUSE_RECOMPUTE", "HH_PRO_CACHE", "HH_FIRST_CONVERT", "HH_LEVEL_SWAP"):
                    os.environ[k] = flag
            A = solver(prec, tols[0], **kw)
            pkg.api._ensure_hierarchy(A, 0)
            del os.environ["HH_SKIP_LAST_UPDATE"], os.environ["HH_SMALL_FUSED"]
            if os.environ.get("HH_CHECK_ALL"):
                for k in ("HH_SCALAR_FAST", "HH_FUSE_RECOMPUTE", "HH_PRO_CACHE", "HH_FIRST_CONVERT", "HH_LEVEL_SWAP"):
                    os.environ.pop(k, None)
            out = []
            for tol in tols:
                A.MG.relativeTol = tol
                t0 = time.perf_counter()
                x, A = pkg.solveLinearSystem(None, b.astype(prec), A)
                out.append((int(A.iterations[0]), float(A.relres[0]), np.asarray(x).astype(np.complex128).copy(),
                            time.perf_counter() - t0))
            res[flag] = out
            pkg.clear(A.MG)
        for k, tol in enumerate(tols):
            a, o = res["1"][k], res["0"][k]
            print(json.dumps({"case": name, "n": n, "tol": tol, "iterations_new": a[0], "iterations_plain": o[0],
                              "relres_new": a[1], "relres_plain": o[1],
                              "rel_diff_solutions": float(np.linalg.norm(a[2] - o[2]) / np.linalg.norm(o[2])),
                              "seconds_new": round(a[3], 3), "seconds_plain": round(o[3], 3)}), flush=True)


if __name__ == "__main__":
    main()
