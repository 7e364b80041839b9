"""Exploration on the GPU box: iteration counts, time per solve and per-kernel device time."""
import argparse, ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=129)
ap.add_argument("--nrhs", type=int, default=4)
ap.add_argument("--prec", default="c128")
ap.add_argument("--levels", type=int, default=3)
ap.add_argument("--cycle", default="V")
ap.add_argument("--relax", default="Jac")
ap.add_argument("--coarse", default="GMRES")
ap.add_argument("--coarse-iters", type=int, default=10)
ap.add_argument("--inner", type=int, default=5)
ap.add_argument("--maxit", type=int, default=40)
ap.add_argument("--tol", type=float, default=1e-6)
ap.add_argument("--shift", type=float, default=0.2)
ap.add_argument("--omega-relax", type=float, default=0.8)
ap.add_argument("--pre", type=int, default=2)
ap.add_argument("--post", type=int, default=2)
ap.add_argument("--krylov", default="GMRES")
ap.add_argument("--sigma", type=float, default=None)
ap.add_argument("--entries", action="store_true")
a = ap.parse_args()
pkg = g.load_package()
n = a.n
t0 = time.time()
cfg = pkg.workloads.config4(n=n, sigma=a.sigma or 8.0 * (n - 1) / 256, seed=1234, pad=max(4, 16 * (n - 1) // 256))
print("model %.1fs" % (time.time() - t0), flush=True)
mesh = pkg.getRegularMesh(cfg["domain"], cfg["n_cells"])
m = cfg["m"]
w = pkg.getMaximalFrequency(m, mesh)
prec = np.complex64 if a.prec == "c64" else np.complex128
gamma = 0.01 * w * np.ones(m.shape) + pkg.getABL(mesh.n + 1, True, cfg["pad"], w)
MG = pkg.getMGparam(prec, pkg.Int64, a.levels, 1, a.maxit, a.tol, a.relax, a.omega_relax, a.pre, a.post, a.cycle, a.coarse,
                    coarseIters=a.coarse_iters)
hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
if a.prec == "mixed":
    MG.cyclePrecision = pkg.ComplexF32
Ainv = pkg.getShiftedLaplacianMultigridSolver(hp, MG, a.shift, a.krylov, a.inner)
t0 = time.time()
hd = pkg.api._ensure_hierarchy(Ainv, 0)
print("setup %.2fs" % (time.time() - t0), flush=True)
N = n**3
g1 = int(np.ceil(np.sqrt(a.nrhs)))
srcs = pkg.workloads.point_sources_top_grid(mesh.n + 1, g1, g1)[: a.nrhs]
tdt = torch.complex64 if a.prec == "c64" else torch.complex128
B = torch.zeros((a.nrhs, N), dtype=tdt, device="cuda")
for c, s in enumerate(srcs):
    B[c, pkg.loc2cs(mesh.n + 1, s) - 1] = 1.0 / mesh.h[0] ** 2
lib = hd.lib
for rep in range(2):
    lib.hh_profile_enable(hd.h, 1 if rep == 1 else 0)
    lib.hh_profile_reset(hd.h)
    torch.cuda.synchronize()
    t0 = time.time()
    X, Ainv = pkg.solveLinearSystem(None, B, Ainv)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("solve rep%d: %.3fs  iters %s  relres max %.2e  -> %.3f RHS/s" % (rep, dt, Ainv.iterations.tolist(), Ainv.relres.max(), a.nrhs / dt), flush=True)
# true residual
Hop = pkg.HelmholtzOperator(hd)
R = Hop.matvec(X) - B
print("true relres", (torch.linalg.vector_norm(R, dim=1) / torch.linalg.vector_norm(B, dim=1)).tolist())
tot = 0.0
rows = []
for t in range(lib.hh_profile_num_tags()):
    cnt, ms, by = C.c_int64(), C.c_double(), C.c_double()
    lib.hh_profile_get(hd.h, t, C.byref(cnt), C.byref(ms), C.byref(by))
    if cnt.value:
        rows.append((lib.hh_profile_tag_name(t).decode(), cnt.value, ms.value, by.value))
        tot += ms.value
for name, cnt, ms, by in sorted(rows, key=lambda r: -r[2]):
    print("%-16s n=%6d  %9.2f ms (%5.1f%%)  avg %8.3f ms  %8.1f GB/s" % (name, cnt, ms, 100 * ms / tot, ms / cnt, by / ms / 1e6 if ms else 0))
print("sum kernels %.1f ms" % tot)
if a.entries:
    ents = []
    for e in range(lib.hh_profile_num_entries(hd.h)):
        tag, cnt, ms, by = C.c_int(), C.c_int64(), C.c_double(), C.c_double()
        lib.hh_profile_entry(hd.h, e, C.byref(tag), C.byref(cnt), C.byref(ms), C.byref(by))
        ents.append((lib.hh_profile_tag_name(tag.value).decode(), cnt.value, ms.value, by.value))
    print("per (tag, bytes-per-launch) entries:")
    for name, cnt, ms, by in sorted(ents, key=lambda r: -r[2])[:40]:
        print("%-20s n=%6d  %9.2f ms (%5.1f%%)  avg %8.4f ms  %7.1f MB/launch  %8.1f GB/s" % (
            name, cnt, ms, 100 * ms / tot, ms / cnt, by / 1e6, by * cnt / ms / 1e6 if ms else 0))
