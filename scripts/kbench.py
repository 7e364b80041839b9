"""Kernel micro-benchmark / ncu target: a few operator applies and multigrid cycles on resident data."""
import argparse, ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=257)
ap.add_argument("--nrhs", type=int, default=8)
ap.add_argument("--prec", default="c128")
ap.add_argument("--applies", type=int, default=5)
ap.add_argument("--cycles", type=int, default=2)
ap.add_argument("--coarse-iters", type=int, default=10)
a = ap.parse_args()
pkg = g.load_package()
n = a.n
cfg = pkg.workloads.config4(n=n, sigma=8.0 * (n - 1) / 256, seed=1234, pad=max(4, 16 * (n - 1) // 256))
mesh = pkg.getRegularMesh(cfg["domain"], cfg["n_cells"])
m = cfg["m"]
w = pkg.getMaximalFrequency(m, mesh)
prec = np.complex128 if a.prec == "c128" else np.complex64
gamma = 0.01 * w * np.ones(m.shape) + pkg.getABL(mesh.n + 1, True, cfg["pad"], w)
MG = pkg.getMGparam(prec, pkg.Int64, 3, 1, 40, 1e-6, "Jac", 0.8, 2, 2, "V", "GMRES", coarseIters=a.coarse_iters)
hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
Ainv = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
hd = pkg.api._ensure_hierarchy(Ainv, 0)
lib = hd.lib
N = n**3
tdt = torch.complex128 if a.prec == "c128" else torch.complex64
torch.manual_seed(0)
X = torch.randn((a.nrhs, N), dtype=tdt, device="cuda")
Y = torch.empty_like(X)
lib.hh_profile_enable(hd.h, 1)
lib.hh_profile_reset(hd.h)
for _ in range(a.applies):
    pkg._lib.check(lib.hh_apply_device(hd.h, X.data_ptr(), Y.data_ptr(), a.nrhs, 0, 0.0, 0), hd.h)
for _ in range(a.cycles):
    pkg._lib.check(lib.hh_cycle_device(hd.h, X.data_ptr(), Y.data_ptr(), a.nrhs), hd.h)
torch.cuda.synchronize()
for t in range(lib.hh_profile_num_tags()):
    cnt, ms, by = C.c_int64(), C.c_double(), C.c_double()
    lib.hh_profile_get(hd.h, t, C.byref(cnt), C.byref(ms), C.byref(by))
    if cnt.value:
        print("%-16s n=%5d avg %8.3f ms  %8.1f GB/s" % (lib.hh_profile_tag_name(t).decode(), cnt.value, ms.value / cnt.value, by.value / ms.value / 1e6))
