"""Small target for compute-sanitizer (memcheck / racecheck): the kernels that tests/__graft_entry__.smoke() does not reach
-- the bench's cycle (one pre-smoothing sweep: k_fine3d_tma_first writing only the residual, k_fine3d_tma_prob turning the
staged b tile into the x' tile in place), the fixed-length GMRES of the coarsest level (k_gmres_small_step_mw, k_combine,
the skipped last update pass), the mapped-memory state publication, a K-cycle with the Jac-GMRES smoother, and the same
solve in two in-process slabs (halo pack / unpack, all-reduced dots).

  compute-sanitizer --tool memcheck  python scripts/sanitize_case.py
  compute-sanitizer --tool racecheck python scripts/sanitize_case.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    pkg = graft.load_package()
    rng = np.random.default_rng(3)
    n = np.array((33, 41, 25))
    dom = sum([[0.0, 0.1 * (v - 1)] for v in n], [])
    mesh = pkg.getRegularMesh(dom, list(n - 1))
    m = 1.0 / (1.5 + 2.0 * rng.random(tuple(n))) ** 2
    w = 0.8 * pkg.getMaximalFrequency(m, mesh)
    gamma = 0.02 * w * (1.0 + rng.random(tuple(n))) + pkg.getABL(n, True, [3, 3, 4], w)
    N = int(np.prod(n))
    B = np.asfortranarray(rng.standard_normal((N, 3)) + 1j * rng.standard_normal((N, 3)))
    ref = None
    for name, kw, slabs, prec in (("W(1,2) Jacobi, coarsest GMRES(6)", dict(relax="Jac", cyc="W", pre=1, post=2), 0, np.complex128),
                                  ("same, ComplexF32", dict(relax="Jac", cyc="W", pre=1, post=2), 0, np.complex64),
                                  ("same in 2 slabs", dict(relax="Jac", cyc="W", pre=1, post=2), 2, np.complex128),
                                  ("K-cycle, Jac-GMRES", dict(relax="Jac-GMRES", cyc="K", pre=2, post=2), 0, np.complex128)):
        MG = pkg.getMGparam(prec, pkg.Int64, 3, 1, 12, 1e-6 if prec == np.complex128 else 1e-4, kw["relax"], 0.8, kw["pre"], kw["post"],
                            kw["cyc"], "GMRES", coarseIters=6)
        hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
        A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
        if slabs:
            A.slabs = {"mode": "local", "devices": [0] * slabs}
        X, A = pkg.solveLinearSystem(None, B.astype(prec), A)
        X = np.reshape(X, (N, 3)).astype(np.complex128)
        if ref is None:
            ref = X
        print(f"{name}: iterations {A.iterations.tolist()}, relres max {A.relres.max():.2e}, "
              f"difference to the first case {np.linalg.norm(X - ref) / np.linalg.norm(ref):.2e}", flush=True)
        pkg.clear(A.MG)
    print("SANITIZE_CASE_OK")


if __name__ == "__main__":
    main()
