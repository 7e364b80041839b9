"""Round-2 sweep on one B200 (config 4, 257^3): solver settings (incl. the reference's production multigrid: 4-5 levels,
K-cycle, Jac-GMRES smoother with nu(l) = l+1, inexact GMRES coarsest solve), batch sizes, and A/B switches of the kernels.
One process, one model; every row is a full solve to the stated tolerance, timed with CUDA events after one warm-up solve.
Writes gpurun_out/sweep_r02.jsonl (one JSON object per row) and prints a markdown table."""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=257)
ap.add_argument("--rows", default="all")
ap.add_argument("--out", default="gpurun_out/sweep_r02.jsonl")
a = ap.parse_args()
pkg = g.load_package()
n = a.n
cfg = pkg.workloads.config4(n=n, sigma=8.0 * (n - 1) / 256, seed=1234, pad=max(4, 16 * (n - 1) // 256))
mesh = pkg.getRegularMesh(cfg["domain"], cfg["n_cells"])
m = cfg["m"]
w = pkg.getMaximalFrequency(m, mesh)
gamma = 0.01 * w * np.ones(m.shape) + pkg.getABL(mesh.n + 1, True, cfg["pad"], w)
N = n**3
all_src = pkg.workloads.point_sources_top_grid(mesh.n + 1, 16, 16)
lp1 = lambda l: l + 1

ROWS = [
    # name, dict(settings)
    ("W(1,2) 3 levels, Jacobi, coarsest GMRES(10) [bench]", dict(nrhs=16)),
    ("same, batch 8", dict(nrhs=8)),
    ("same, batch 4", dict(nrhs=4)),
    ("same, batch 2", dict(nrhs=2)),
    ("same, batch 1", dict(nrhs=1)),
    ("bench, HH_COARSE_TILE=16x8", dict(nrhs=16, env={"HH_COARSE_TILE": "16x8"})),
    ("bench, HH_COARSE_TILE=alt", dict(nrhs=16, env={"HH_COARSE_TILE": "alt"})),
    ("bench, HH_SCALED_GMRES=0", dict(nrhs=16, env={"HH_SCALED_GMRES": "0"})),
    ("bench, coarsest GMRES(8)", dict(nrhs=16, coarse_iters=8)),
    ("bench, coarsest GMRES(12)", dict(nrhs=16, coarse_iters=12)),
    ("K-cycle 3 levels, Jac-GMRES nu=l+1, coarsest GMRES(10)", dict(nrhs=8, cycle="K", relax="Jac-GMRES", pre=lp1, post=lp1)),
    ("K-cycle 4 levels, Jac-GMRES nu=l+1, coarsest GMRES(10)", dict(nrhs=8, levels=4, cycle="K", relax="Jac-GMRES", pre=lp1, post=lp1)),
    ("K-cycle 5 levels, Jac-GMRES nu=l+1, coarsest GMRES(10) [reference production]",
     dict(nrhs=8, levels=5, cycle="K", relax="Jac-GMRES", pre=lp1, post=lp1, maxit=50)),
    ("K-cycle 5 levels, Jac-GMRES nu=l+1, coarsest GMRES(5)", dict(nrhs=8, levels=5, cycle="K", relax="Jac-GMRES", pre=lp1, post=lp1, coarse_iters=5, maxit=50)),
    ("W-cycle 4 levels, Jac-GMRES nu=2", dict(nrhs=8, levels=4, cycle="W", relax="Jac-GMRES", pre=2, post=2, maxit=50)),
    ("V-cycle 5 levels, Jac-GMRES nu=l+1", dict(nrhs=8, levels=5, cycle="V", relax="Jac-GMRES", pre=lp1, post=lp1, maxit=50)),
    ("W-cycle 4 levels, Jacobi(1,2), coarsest GMRES(10)", dict(nrhs=8, levels=4, maxit=20)),
    ("K-cycle 5 levels, Jac-GMRES nu=l+1, ComplexF32 tol 1e-5 [paper runs]",
     dict(nrhs=16, prec="c64", tol=1e-5, levels=5, cycle="K", relax="Jac-GMRES", pre=lp1, post=lp1, maxit=50)),
    ("bench settings, ComplexF32 tol 1e-5", dict(nrhs=16, prec="c64", tol=1e-5)),
    ("bench settings, mixed (ComplexF32 cycle in ComplexF64 FGMRES)", dict(nrhs=16, prec="mixed")),
]
if a.rows != "all":
    want = [int(v) for v in a.rows.split(",")]
    ROWS = [ROWS[i] for i in want]

os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
fout = open(a.out, "a")
print("| settings | precision | batch | iterations | ms / batch | RHS/s | true relres max |")
print("|---|---|---:|---:|---:|---:|---:|")
for name, s in ROWS:
    env = s.get("env", {})
    for k, v in env.items():
        os.environ[k] = v
    prec_s = s.get("prec", "c128")
    prec = np.complex64 if prec_s == "c64" else np.complex128
    tdt = torch.complex64 if prec_s == "c64" else torch.complex128
    tol = s.get("tol", 1e-6)
    nrhs = s["nrhs"]
    row = dict(name=name, prec=prec_s, nrhs=nrhs, tol=tol)
    try:
        MG = pkg.getMGparam(prec, pkg.Int64, s.get("levels", 3), 1, s.get("maxit", 30), tol, s.get("relax", "Jac"), 0.8,
                            s.get("pre", 1), s.get("post", 2), s.get("cycle", "W"), "GMRES", coarseIters=s.get("coarse_iters", 10))
        if prec_s == "mixed":
            MG.cyclePrecision = pkg.ComplexF32
        hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
        A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
        t0 = time.time()
        hd = pkg.api._ensure_hierarchy(A, 0)
        row["setup_s"] = time.time() - t0
        B = torch.zeros((nrhs, N), dtype=tdt, device="cuda")
        for c in range(nrhs):
            B[c, pkg.loc2cs(mesh.n + 1, all_src[(17 * c) % 256]) - 1] = 1.0 / mesh.h[0] ** 2
        X = torch.empty_like(B)
        pkg.solveLinearSystem_(None, B, X, A)  # warm-up (allocations)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pkg.solveLinearSystem_(None, B, X, A)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        Hop = pkg.HelmholtzOperator(hd)
        R = Hop.matvec(X) - B
        tr = float((torch.linalg.vector_norm(R, dim=1) / torch.linalg.vector_norm(B, dim=1)).max())
        its = A.iterations
        row.update(ms=ms, rhs_per_s=nrhs / (ms / 1e3), iters_min=int(its.min()), iters_max=int(its.max()), iters_mean=float(its.mean()),
                   true_relres_max=tr, converged=bool((A.relres <= tol).all()))
        print("| %s | %s | %d | %d-%d | %.1f | %.2f | %.2e |%s" % (name, prec_s, nrhs, its.min(), its.max(), ms, row["rhs_per_s"], tr,
                                                                  "" if row["converged"] else " NOT CONVERGED"), flush=True)
        del B, X, R
        pkg.clear(MG)
    except Exception as e:  # keep sweeping
        row["error"] = str(e)[:300]
        print("| %s | %s | %d | error: %s |" % (name, prec_s, nrhs, str(e)[:200]), flush=True)
    torch.cuda.empty_cache()
    fout.write(json.dumps(row) + "\n")
    fout.flush()
    for k in env:
        del os.environ[k]
