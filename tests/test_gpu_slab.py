"""Slab decomposition of one grid (BASELINE config 5) against the whole-grid solve of the same
library on the same inputs: operator apply, Galerkin stencils, FGMRES / BiCGSTAB solves.  The slabs of these tests
share one GPU (hh_create_slab_local with a repeated device ordinal: one host thread per slab, halos by device copies),
which runs every slab kernel path -- restricted plane ranges, halo planes, global boundary rows, all-reduced dots --
on a single-GPU box; tests/test_gpu_slab_nccl.py covers the one-process-per-GPU NCCL transport on >= 2 GPUs."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _problem(pkg, nodes=(17, 13, 33), seed=3, neumann=True):
    rng = np.random.default_rng(seed)
    n = np.array(nodes)
    dom = [0.0, 0.1 * (n[0] - 1), 0.0, 0.12 * (n[1] - 1), 0.0, 0.09 * (n[2] - 1)]
    mesh = pkg.getRegularMesh(dom, list(n - 1))
    v = 1.5 + 2.0 * rng.random(tuple(n))
    m = 1.0 / v**2
    w = 0.8 * pkg.getMaximalFrequency(m, mesh)
    gamma = 0.02 * w * (1.0 + rng.random(tuple(n))) + pkg.getABL(n, neumann, [3, 3, 4], w)
    return mesh, m, w, gamma


def _solver(pkg, mesh, m, w, gamma, prec, slabs, levels=3, cycle="W", relax="Jac", krylov="GMRES", inner=5, tol=1e-8,
            neumann=True, pre=1, post=2, cyc_prec=None):
    MG = pkg.getMGparam(prec, pkg.Int64, levels, 1, 40, tol, relax, 0.8, pre, post, cycle, "GMRES")
    MG.coarseIters = 10
    if cyc_prec is not None:
        MG.cyclePrecision = cyc_prec
    hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, neumann, True)
    A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, krylov, inner)
    if slabs:
        A.slabs = {"mode": "local", "devices": [0] * slabs}
    return A


def _apply(pkg, A, X, shifted, shift, transpose):
    hd = pkg.api._ensure_hierarchy(A, 0)
    Y = np.empty_like(X, order="F")
    pkg._lib.check(hd.lib.hh_apply(hd.h, X.ctypes.data, Y.ctypes.data, X.shape[1], shifted, shift, transpose), hd.h)
    return Y


@pytest.mark.parametrize("prec,tol", [(np.complex128, 1e-13), (np.complex64, 3e-5)])
@pytest.mark.parametrize("neumann", [True, False])
def test_slab_apply_matches_whole_grid(gpu_pkg, prec, tol, neumann):
    pkg = gpu_pkg
    mesh, m, w, gamma = _problem(pkg, neumann=neumann)
    N = int(np.prod(mesh.n + 1))
    rng = np.random.default_rng(11)
    X = np.asfortranarray((rng.standard_normal((N, 3)) + 1j * rng.standard_normal((N, 3))).astype(prec))
    ref = _solver(pkg, mesh, m, w, gamma, prec, 0, neumann=neumann)
    for nslab in (2, 3, 8):
        A = _solver(pkg, mesh, m, w, gamma, prec, nslab, neumann=neumann)
        for shifted, shift, tr in ((0, 0.0, 0), (1, 0.2, 0), (1, 0.35, 1)):
            assert rel_err(_apply(pkg, A, X, shifted, shift, tr), _apply(pkg, ref, X, shifted, shift, tr)) < tol
        pkg.clear(A.MG)
    pkg.clear(ref.MG)


@pytest.mark.parametrize("prec,tol", [(np.complex128, 1e-13), (np.complex64, 2e-5)])
def test_slab_galerkin_stencils_match_whole_grid(gpu_pkg, prec, tol):
    """every slab's rows of A_c = R A P (levels 1 and 2) equal the whole-grid hierarchy's rows of the same planes;
    level 2 is built from exchanged halo rows of level 1"""
    pkg = gpu_pkg
    mesh, m, w, gamma = _problem(pkg)
    ref = _solver(pkg, mesh, m, w, gamma, prec, 0)
    hr = pkg.api._ensure_hierarchy(ref, 0)
    n3 = int(mesh.n[2] + 1)
    for nslab in (2, 3):
        A = _solver(pkg, mesh, m, w, gamma, prec, nslab)
        hd = pkg.api._ensure_hierarchy(A, 0)
        for level in (1, 2):
            nl = np.zeros(3, dtype=np.int64)
            pkg._lib.check(hr.lib.hh_level_nodes(hr.h, level, nl.ctypes.data_as(C.POINTER(C.c_int64))), hr.h)
            full = np.empty(27 * int(np.prod(nl)), dtype=prec)
            pkg._lib.check(hr.lib.hh_get_level_stencil(hr.h, level, full.ctypes.data), hr.h)
            full = full.reshape((27, nl[2], nl[1], nl[0]))
            for q in range(nslab):
                geo = pkg.slabPartition(n3, 3, nslab, q)[level]
                ns = np.zeros(3, dtype=np.int64)
                pkg._lib.check(hd.lib.hh_slab_level_stencil(hd.h, q, level, ns.ctypes.data_as(C.POINTER(C.c_int64)), None), hd.h)
                assert tuple(ns) == (nl[0], nl[1], geo["nloc"])
                loc = np.empty(27 * int(np.prod(ns)), dtype=prec)
                pkg._lib.check(hd.lib.hh_slab_level_stencil(hd.h, q, level, None, loc.ctypes.data), hd.h)
                loc = loc.reshape((27, ns[2], ns[1], ns[0]))
                own = loc[:, geo["zb"]:geo["ze"]]
                assert rel_err(own, full[:, geo["own0"]:geo["own1"]]) < tol, (nslab, level, q)
        pkg.clear(A.MG)
    pkg.clear(ref.MG)


def _rhs(pkg, mesh, prec, nrand=2):
    nodes = mesh.n + 1
    N = int(np.prod(nodes))
    srcs = pkg.workloads.point_sources_top_grid(nodes, 2, 1) + [[int(nodes[0] // 2), int(nodes[1] // 2), int(nodes[2] - 2)]]
    rng = np.random.default_rng(5)
    B = np.zeros((N, len(srcs) + nrand), dtype=prec, order="F")
    for c, s in enumerate(srcs):
        B[pkg.loc2cs(nodes, s) - 1, c] = 1.0 / mesh.h[0] ** 2
    B[:, len(srcs):] = rng.standard_normal((N, nrand)) + 1j * rng.standard_normal((N, nrand))
    return B, srcs


@pytest.mark.parametrize("krylov,inner", [("GMRES", 5), ("BiCGSTAB", 0)])
def test_slab_solve_matches_whole_grid(gpu_pkg, krylov, inner):
    """same iteration counts and the same solution (to reduction-order round-off) as the whole-grid solve"""
    pkg = gpu_pkg
    mesh, m, w, gamma = _problem(pkg, nodes=(33, 25, 33))
    B, _ = _rhs(pkg, mesh, np.complex128)
    ref = _solver(pkg, mesh, m, w, gamma, np.complex128, 0, krylov=krylov, inner=inner)
    Xr, ref = pkg.solveLinearSystem(None, B, ref)
    assert ref.relres.max() < 1e-8
    for nslab in (2, 3, 8):
        A = _solver(pkg, mesh, m, w, gamma, np.complex128, nslab, krylov=krylov, inner=inner)
        X, A = pkg.solveLinearSystem(None, B, A)
        assert np.array_equal(A.iterations, ref.iterations), (nslab, A.iterations, ref.iterations)
        assert rel_err(X, Xr) < 1e-9, nslab
        pkg.clear(A.MG)
    pkg.clear(ref.MG)


def test_slab_point_sources_and_true_residual(gpu_pkg, ho):
    """hh_solve_point_sources on slabs (sources land in the slab that owns their plane); the true residual is
    checked with the oracle's assembled operator"""
    pkg = gpu_pkg
    mesh, m, w, gamma = _problem(pkg, nodes=(33, 25, 33))
    B, srcs = _rhs(pkg, mesh, np.complex128, nrand=0)
    A = _solver(pkg, mesh, m, w, gamma, np.complex128, 3, tol=1e-7)
    X, A = pkg.solvePointSources(A, srcs, amplitudes=np.full(len(srcs), 1.0 / mesh.h[0] ** 2))
    omesh = ho.getRegularMesh(list(mesh.domain), list(mesh.n))
    H = ho.GetHelmholtzOperator(omesh, m, w, gamma, True, True)
    for c in range(B.shape[1]):
        assert np.linalg.norm(H @ X[:, c] - B[:, c]) / np.linalg.norm(B[:, c]) < 2e-7
    pkg.clear(A.MG)


@pytest.mark.parametrize("variant", ["c32", "mixed", "kcycle_jacgmres", "v22_adjoint"])
def test_slab_solve_variants(gpu_pkg, variant):
    pkg = gpu_pkg
    mesh, m, w, gamma = _problem(pkg, nodes=(33, 25, 33))
    kw = dict(krylov="GMRES", inner=5)
    prec, tol, cmp_tol, tr = np.complex128, 1e-8, 1e-9, 0
    if variant == "c32":
        prec, tol, cmp_tol = np.complex64, 1e-5, 2e-4
    elif variant == "mixed":
        kw["cyc_prec"] = np.complex64
        cmp_tol = 1e-7
    elif variant == "kcycle_jacgmres":
        kw.update(cycle="K", relax="Jac-GMRES", pre=2, post=2)
        cmp_tol = 1e-7
    elif variant == "v22_adjoint":
        kw.update(cycle="V", pre=2, post=2)
        tr = 1
    B, _ = _rhs(pkg, mesh, prec)
    ref = _solver(pkg, mesh, m, w, gamma, prec, 0, tol=tol, **kw)
    Xr, ref = pkg.solveLinearSystem(None, B, ref, tr)
    A = _solver(pkg, mesh, m, w, gamma, prec, 3, tol=tol, **kw)
    X, A = pkg.solveLinearSystem(None, B, A, tr)
    assert ref.relres.max() < 1.01 * tol and A.relres.max() < 1.01 * tol
    if variant in ("c32", "mixed", "kcycle_jacgmres"):
        assert np.abs(A.iterations.astype(int) - ref.iterations.astype(int)).max() <= 1
    else:
        assert np.array_equal(A.iterations, ref.iterations)
    assert rel_err(X, Xr) < cmp_tol
    pkg.clear(A.MG)
    pkg.clear(ref.MG)


def test_slab_error_paths(gpu_pkg):
    pkg = gpu_pkg
    mesh, m, w, gamma = _problem(pkg)
    B, _ = _rhs(pkg, mesh, np.complex128, nrand=0)
    # the exact coarsest solve is not distributed
    A = _solver(pkg, mesh, m, w, gamma, np.complex128, 2)
    A.MG.coarseSolveType = "NoMUMPS"
    with pytest.raises(pkg._lib.HelmholtzB200Error) as e:
        pkg.solveLinearSystem(None, B, A)
    assert e.value.code == pkg._lib.HH_ERR_UNSUPPORTED
    pkg.clear(A.MG)
    # more slabs than cells of the coarsest level (33 planes, 3 levels -> 8 cells)
    A = _solver(pkg, mesh, m, w, gamma, np.complex128, 9)
    with pytest.raises(pkg._lib.HelmholtzB200Error) as e:
        pkg.solveLinearSystem(None, B, A)
    assert e.value.code == pkg._lib.HH_ERR_ARG


def test_slab_split_launches_match(gpu_pkg, monkeypatch):
    """HH_HALO_SPLIT=1: interior planes and the one or two boundary planes of every stencil-type kernel are separate
    launches, as in the path that overlaps the NCCL halo exchange with the interior (same results)"""
    pkg = gpu_pkg
    mesh, m, w, gamma = _problem(pkg, nodes=(33, 25, 33))
    B, _ = _rhs(pkg, mesh, np.complex128)
    ref = _solver(pkg, mesh, m, w, gamma, np.complex128, 0)
    Xr, ref = pkg.solveLinearSystem(None, B, ref)
    monkeypatch.setenv("HH_HALO_SPLIT", "1")
    for nslab, prec, tol in ((2, np.complex128, 1e-9), (3, np.complex128, 1e-9), (2, np.complex64, 2e-4)):
        A = _solver(pkg, mesh, m, w, gamma, prec, nslab, tol=1e-8 if prec == np.complex128 else 1e-5)
        X, A = pkg.solveLinearSystem(None, B.astype(prec), A)
        if prec == np.complex128:
            assert np.array_equal(A.iterations, ref.iterations)
        assert rel_err(X, Xr) < tol
        pkg.clear(A.MG)
    pkg.clear(ref.MG)


def test_slab_diagonal_matches_whole_grid(gpu_pkg):
    """hh_get_diagonal on a slab handle: mass + Sommerfeld faces follow the global plane index"""
    pkg = gpu_pkg
    for neumann in (True, False):
        mesh, m, w, gamma = _problem(pkg, neumann=neumann)
        ref = _solver(pkg, mesh, m, w, gamma, np.complex128, 0, neumann=neumann)
        A = _solver(pkg, mesh, m, w, gamma, np.complex128, 3, neumann=neumann)
        out = []
        for S in (ref, A):
            hd = pkg.api._ensure_hierarchy(S, 0)
            d = np.empty(hd.N, dtype=np.complex128)
            pkg._lib.check(hd.lib.hh_get_diagonal(hd.h, 1, 0.2, d.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))), hd.h)
            out.append(d)
            pkg.clear(S.MG)
        assert rel_err(out[1], out[0]) < 1e-14
