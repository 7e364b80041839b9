"""Regression tests for defects found in review of round 1: Krylov arena sizing across calls with different
Krylov options, layout of CUDA tensors handed to solveLinearSystem, the un-pivoted exact coarsest solve on an
indefinite / singular coarse operator, the Neumann order of the operator the solver is handed."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import rel_err

pytestmark = pytest.mark.gpu


class _env:
    """Set environment switches for the creation of a handle (the library reads them in the solver's constructor) and put
    the previous values back afterwards."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        import os

        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *exc):
        import os

        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _problem(pkg, n=17, seed=9):
    cfg = pkg.workloads.config4(n=n, sigma=2.0, seed=seed, pad=3)
    mesh = pkg.getRegularMesh(cfg["domain"], cfg["n_cells"])
    m = cfg["m"]
    w = pkg.getMaximalFrequency(m, mesh)
    gamma = 0.01 * w * np.ones(m.shape) + pkg.getABL(mesh.n + 1, True, cfg["pad"], w)
    return cfg, mesh, m, w, gamma


def test_krylov_arena_follows_per_call_options(gpu_pkg, ho):
    """Krylov method and restart length are per-call options on a live hierarchy: BiCGSTAB with 2 RHS (7 vectors)
    followed by GMRES(20) with 1 RHS (41 vectors, stride N*capacity), GMRES(5) with 16 RHS followed by GMRES(20) with
    4 -- sequences that used to address the arena past its end."""
    pkg = gpu_pkg
    cfg, mesh, m, w, gamma = _problem(pkg)
    omesh = ho.getRegularMesh(cfg["domain"], cfg["n_cells"])
    H = ho.GetHelmholtzOperator(omesh, m, w, gamma, True, True)
    lu = spla.splu(H.tocsc())
    rng = np.random.default_rng(1)
    N = 17**3
    B = np.asfortranarray(rng.standard_normal((N, 16)) + 1j * rng.standard_normal((N, 16)))
    Xt = lu.solve(B)
    hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
    MG = pkg.getMGparam(pkg.ComplexF64, pkg.Int64, 2, 1, 80, 1e-10, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
    A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "BiCGSTAB", 0)
    for kry, inner, cols in (("BiCGSTAB", 0, 2), ("GMRES", 20, 1), ("GMRES", 5, 16), ("GMRES", 20, 4), ("BiCGSTAB", 0, 16),
                             ("GMRES", 30, 3)):
        A.Krylov, A.inner = kry, inner
        X, A = pkg.solveLinearSystem(None, B[:, :cols].copy(order="F"), A)
        assert rel_err(np.reshape(X, (N, -1)), Xt[:, :cols]) < 1e-7, (kry, inner, cols)


def test_torch_tensor_layouts(gpu_pkg, ho):
    """(N,), (N, nrhs) as the reference and the numpy path, or the zero-copy (nrhs, N); anything else is an error."""
    import torch

    pkg = gpu_pkg
    cfg, mesh, m, w, gamma = _problem(pkg)
    rng = np.random.default_rng(2)
    N = 17**3
    B = np.asfortranarray(rng.standard_normal((N, 3)) + 1j * rng.standard_normal((N, 3)))
    hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
    MG = pkg.getMGparam(pkg.ComplexF64, pkg.Int64, 2, 1, 30, 1e-8, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
    A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
    Xh, A = pkg.solveLinearSystem(None, B, A)
    Bc = torch.as_tensor(B, device="cuda")  # (N, nrhs), the reference's layout
    assert Bc.shape == (N, 3)
    Xc, A = pkg.solveLinearSystem(None, Bc, A)
    assert Xc.shape == (N, 3) and rel_err(Xc.cpu().numpy(), Xh) < 1e-13
    Xr, A = pkg.solveLinearSystem(None, Bc.t().contiguous(), A)  # (nrhs, N) zero-copy
    assert Xr.shape == (3, N) and rel_err(Xr.cpu().numpy().T, Xh) < 1e-13
    Xv, A = pkg.solveLinearSystem(None, Bc[:, 1].contiguous(), A)
    assert Xv.shape == (N,) and rel_err(Xv.cpu().numpy(), Xh[:, 1]) < 1e-12
    Hop = pkg.HelmholtzOperator(A.MG._hd)
    assert rel_err((Hop @ Bc).cpu().numpy(), Hop @ B) < 1e-14
    with pytest.raises(ValueError):
        pkg.solveLinearSystem(None, Bc[: N - 1], A)
    with pytest.raises(ValueError):
        pkg.solveLinearSystem_(None, Bc, torch.empty((3, N), dtype=torch.complex128, device="cuda"), A)


def test_exact_coarse_solve_reports_vanishing_pivot(gpu_pkg):
    """The exact coarsest solve factorises without pivoting.  A singular coarse operator (pure Neumann Laplacian: m = 0,
    no shift, no attenuation, no Sommerfeld) must be reported, not returned as Inf/NaN; shift = 0 with attenuation
    (the reference's getAfun branch) still works."""
    pkg = gpu_pkg
    n = 9
    mesh = pkg.getRegularMesh([0.0, 0.8, 0.0, 0.8, 0.0, 0.8], [n - 1] * 3)
    zero = np.zeros((n, n, n))
    MG = pkg.getMGparam(pkg.ComplexF64, pkg.Int64, 2, 1, 5, 1e-6, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
    hp = pkg.HelmholtzParam(mesh, zero, zero, 1.0, True, False)
    A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.0, "GMRES", 5)
    b = np.zeros(n**3, dtype=np.complex128)
    b[0] = 1.0
    with pytest.raises(pkg._lib.HelmholtzB200Error) as e:
        pkg.solveLinearSystem(None, b, A)
    assert "pivot" in str(e.value)
    cfg, mesh, m, w, gamma = _problem(pkg)
    MG = pkg.getMGparam(pkg.ComplexF64, pkg.Int64, 2, 1, 60, 1e-8, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
    hp = pkg.HelmholtzParam(mesh, gamma + 0.3 * w, m.ravel(order="F"), w, True, True)
    A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.0, "GMRES", 10)
    q = np.zeros(17**3, dtype=np.complex128)
    q[17 * 17 * 3 + 100] = 1.0
    x, A = pkg.solveLinearSystem(None, q, A)
    assert A.relres[0] <= 1e-8


def test_solver_follows_neumann_order_of_the_operator(gpu_pkg, ho):
    """solveLinearSystem builds its hierarchy from the operator it is handed (ShiftedLaplacianMultigridSolver.jl:65):
    GetHelmholtzOperator(Hparam, 1) + shift solves the first-order-Neumann system."""
    pkg = gpu_pkg
    cfg, mesh, m, w, gamma = _problem(pkg)
    omesh = ho.getRegularMesh(cfg["domain"], cfg["n_cells"])
    hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
    q = np.zeros(17**3, dtype=np.complex128)
    q[17 * 17 * 2 + 40] = 1.0
    for order in (1, 2):
        H1 = pkg.GetHelmholtzOperator(hp, order)
        MG = pkg.getMGparam(pkg.ComplexF64, pkg.Int64, 2, 1, 60, 1e-9, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
        A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 10)
        x, A = pkg.solveLinearSystem((H1 + pkg.GetHelmholtzShiftOP(m, w, 0.2)).H, q, A)
        Ho = ho.GetHelmholtzOperator(omesh, m, w, gamma, True, True, order)
        assert np.linalg.norm(Ho @ x - q) / np.linalg.norm(q) < 1e-8, order


@pytest.mark.parametrize("dim,neumann,prec,slabs", [(2, True, np.complex128, 0), (2, False, np.complex128, 0),
                                                    (3, True, np.complex128, 0), (3, False, np.complex64, 0),
                                                    (3, True, np.complex128, 3)])
def test_device_side_abl_and_frequency_sweep(gpu_pkg, ho, dim, neumann, prec, slabs):
    """hh_set_frequency_abl evaluates gamma0 + getABL on the device for a new frequency; the result equals
    the oracle's getABL (src/GetHelmholtz.jl:97-220) and the solve equals that of a fresh handle built from host arrays."""
    pkg = gpu_pkg
    rng = np.random.default_rng(5)
    nodes = [33, 25] if dim == 2 else [17, 13, 33]
    dom = sum([[0.0, 0.1 * (n - 1)] for n in nodes], [])
    mesh = pkg.getRegularMesh(dom, np.array(nodes) - 1)
    m = 1.0 / (1.5 + 2.0 * rng.random(nodes)) ** 2
    w0 = pkg.getMaximalFrequency(m, mesh)
    pad = [4, 3] if dim == 2 else [3, 3, 4]
    g0 = 0.01 * w0 * np.ones(nodes) + ho.getABL(nodes, neumann, pad, w0)
    lv = 2 if dim == 2 else 3
    MG = pkg.getMGparam(prec, pkg.Int64, lv, 1, 60, 1e-8 if prec == np.complex128 else 1e-5, "Jac", 0.8, 2, 2, "V", "GMRES", coarseIters=10)
    hp = pkg.HelmholtzParam(mesh, g0, m.ravel(order="F"), w0, neumann, True)
    A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
    if slabs:
        A.slabs = {"mode": "local", "devices": [0] * slabs}
    q = np.zeros(int(np.prod(nodes)), dtype=prec)
    q[int(np.prod(nodes)) // 2 + 3] = 1.0
    x0, A = pkg.solveLinearSystem(None, q, A)
    assert abs(pkg.getMaximalFrequencyDevice(A) - w0) <= (1e-12 if prec == np.complex128 else 1e-6) * w0
    w1 = 0.8 * w0
    A = pkg.setFrequencyABL(A, w1, 0.02 * w1, pad, w1)
    g1 = 0.02 * w1 + ho.getABL(nodes, neumann, pad, w1)
    assert np.abs(A.helmParam.gamma.reshape(nodes, order="F") - g1).max() <= (1e-13 if prec == np.complex128 else 1e-6) * g1.max()
    assert not pkg.hierarchyExists(A.MG)
    x1, A = pkg.solveLinearSystem(None, q, A)
    MG2 = pkg.getMGparam(prec, pkg.Int64, lv, 1, 60, MG.relativeTol, "Jac", 0.8, 2, 2, "V", "GMRES", coarseIters=10)
    hp2 = pkg.HelmholtzParam(mesh, g1, m.ravel(order="F"), w1, neumann, True)
    A2 = pkg.getShiftedLaplacianMultigridSolver(hp2, MG2, 0.2, "GMRES", 5)
    x2, A2 = pkg.solveLinearSystem(None, q, A2)
    assert rel_err(x1, x2) < (1e-9 if prec == np.complex128 else 2e-4)
    assert rel_err(x1, x0) > 1e-3  # really the new frequency


@pytest.mark.parametrize("prec,tol", [(np.complex128, 1e-13), (np.complex64, 2e-5)])
@pytest.mark.parametrize("nodes,pre,post,cyc,nrhs", [((33, 33, 33), 2, 2, "V", 2), ((65, 49, 37), 1, 2, "W", 3), ((41, 25, 33), 1, 4, "V", 1),
                                                      ((97, 17, 21), 2, 2, "W", 5)])
def test_fused_two_sweep_post_smoothing_equals_separate_sweeps(gpu_pkg, prec, tol, nodes, pre, post, cyc, nrhs):
    """k_fine3d_tma_pro2 (correction + two post-sweeps in one pass, x' and x1 never in HBM) against the same cycle with the
    second sweep as its own kernel (HH_FUSE_POST2=0), on grids with partial tiles in every direction."""
    import os

    pkg = gpu_pkg
    rng = np.random.default_rng(17)
    n = np.array(nodes)
    dom = sum([[0.0, 0.1 * (v - 1)] for v in n], [])
    mesh = pkg.getRegularMesh(dom, list(n - 1))
    m = 1.0 / (1.5 + 2.0 * rng.random(tuple(n))) ** 2
    w = 0.8 * pkg.getMaximalFrequency(m, mesh)
    gamma = 0.02 * w * (1.0 + rng.random(tuple(n))) + pkg.getABL(n, True, [3, 3, 4], w)
    N = int(np.prod(n))
    B = np.asfortranarray((rng.standard_normal((N, nrhs)) + 1j * rng.standard_normal((N, nrhs))).astype(prec))
    out = {}
    for flag in ("1", "0"):
        os.environ["HH_FUSE_POST2"] = flag
        try:
            MG = pkg.getMGparam(prec, pkg.Int64, 3, 1, 30, 1e-6, "Jac", 0.8, pre, post, cyc, "GMRES", coarseIters=6)
            hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
            A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
            hd = pkg.api._ensure_hierarchy(A, 0)
        finally:
            del os.environ["HH_FUSE_POST2"]
        Z = np.empty_like(B, order="F")
        pkg._lib.check(hd.lib.hh_cycle(hd.h, B.ctypes.data, Z.ctypes.data, nrhs), hd.h)
        X, A = pkg.solveLinearSystem(None, B, A)
        out[flag] = (Z, X, A.iterations.copy())
        pkg.clear(MG)
    assert rel_err(out["1"][0], out["0"][0]) < tol
    assert np.array_equal(out["1"][2], out["0"][2])
    assert rel_err(out["1"][1], out["0"][1]) < 100 * tol


@pytest.mark.parametrize("prec,tol", [(np.complex128, 1e-9), (np.complex64, 2e-4)])
@pytest.mark.parametrize("levels,relax,cyc,coarse,citers,pre,post", [
    (3, "Jac", "W", "GMRES", 10, 1, 2),          # the bench's cycle: coarsest Jacobi-GMRES(10)
    (3, "Jac-GMRES", "K", "GMRES", 5, 2, 2),     # Jac-GMRES smoother on every level + K-cycle FGMRES(2)
    (2, "Jac", "V", "GMRES", 13, 2, 2),          # more steps than one vector group: the un-fused path with the skipped last pass
])
def test_skipped_last_update_and_fused_level_gmres_match_plain_path(gpu_pkg, prec, tol, levels, relax, cyc, coarse, citers, pre, post):
    """Round-2 Krylov savings against the plain path (HH_SKIP_LAST_UPDATE=0 HH_SMALL_FUSED=0): the last column of every
    GMRES cycle takes h_{j+1,j} from the dot pass instead of a pass that writes a vector nobody reads, and the
    fixed-length GMRES of a level runs one scalar kernel per step with the Givens update deferred, plus a one-pass
    x (+)= D^-1 V y.  One cycle, the full solve and the iteration counts must agree."""
    import os

    pkg = gpu_pkg
    rng = np.random.default_rng(23)
    n = np.array((33, 41, 25))
    dom = sum([[0.0, 0.1 * (v - 1)] for v in n], [])
    mesh = pkg.getRegularMesh(dom, list(n - 1))
    m = 1.0 / (1.5 + 2.0 * rng.random(tuple(n))) ** 2
    w = 0.8 * pkg.getMaximalFrequency(m, mesh)
    gamma = 0.02 * w * (1.0 + rng.random(tuple(n))) + pkg.getABL(n, True, [3, 3, 4], w)
    N = int(np.prod(n))
    nrhs = 3
    B = np.asfortranarray((rng.standard_normal((N, nrhs)) + 1j * rng.standard_normal((N, nrhs))).astype(prec))
    B[:, 1] = 0  # a zero right-hand side rides along (frozen column)
    out = {}
    for flag in ("1", "0"):
        with _env(HH_SKIP_LAST_UPDATE=flag, HH_SMALL_FUSED=flag):
            MG = pkg.getMGparam(prec, pkg.Int64, levels, 1, 40, 1e-6 if prec == np.complex128 else 1e-4, relax, 0.8, pre, post, cyc,
                                coarse, coarseIters=citers)
            hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
            A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
            hd = pkg.api._ensure_hierarchy(A, 0)
        Z = np.empty_like(B, order="F")
        pkg._lib.check(hd.lib.hh_cycle(hd.h, B.ctypes.data, Z.ctypes.data, nrhs), hd.h)
        X, A = pkg.solveLinearSystem(None, B, A)
        out[flag] = (Z, np.reshape(X, (N, nrhs)).copy(), A.iterations.copy())
        pkg.clear(MG)
    assert np.all(out["1"][0][:, 1] == 0) and np.all(out["1"][1][:, 1] == 0)
    assert rel_err(out["1"][0], out["0"][0]) < tol
    if prec == np.complex128:
        assert np.array_equal(out["1"][2], out["0"][2])
    else:
        assert np.abs(out["1"][2].astype(int) - out["0"][2].astype(int)).max() <= 1
    # both solves stop at the same residual tolerance: they agree to a small multiple of it
    assert rel_err(out["1"][1], out["0"][1]) < (2e-5 if prec == np.complex128 else 5e-3)


@pytest.mark.parametrize("prec,tol", [(np.complex128, 1e-13), (np.complex64, 2e-5)])
@pytest.mark.parametrize("nodes,post,cyc,nrhs", [((33, 33, 33), 2, "W", 2), ((65, 49, 37), 1, "V", 3), ((41, 25, 33), 3, "V", 1),
                                                  ((97, 17, 21), 2, "W", 5)])
def test_recomputed_first_sweep_equals_stored_first_sweep(gpu_pkg, prec, tol, nodes, post, cyc, nrhs):
    """k_fine3d_tma_prob (cycles with ONE pre-smoothing sweep: the cycle start writes only the residual, and the
    correction + first post-sweep pass recomputes x1 = dinv .* b from b staged with halo) against the path that stores x1
    and reads it back (HH_FUSE_RECOMPUTE=0), on grids with partial tiles in every direction; also in 3 slabs."""
    import os

    pkg = gpu_pkg
    rng = np.random.default_rng(29)
    n = np.array(nodes)
    dom = sum([[0.0, 0.1 * (v - 1)] for v in n], [])
    mesh = pkg.getRegularMesh(dom, list(n - 1))
    m = 1.0 / (1.5 + 2.0 * rng.random(tuple(n))) ** 2
    w = 0.8 * pkg.getMaximalFrequency(m, mesh)
    gamma = 0.02 * w * (1.0 + rng.random(tuple(n))) + pkg.getABL(n, True, [3, 3, 4], w)
    N = int(np.prod(n))
    B = np.asfortranarray((rng.standard_normal((N, nrhs)) + 1j * rng.standard_normal((N, nrhs))).astype(prec))
    out = {}
    for flag, slabs in (("1", 0), ("0", 0), ("1", 3)):
        with _env(HH_FUSE_RECOMPUTE=flag):
            MG = pkg.getMGparam(prec, pkg.Int64, 3, 1, 30, 1e-6 if prec == np.complex128 else 1e-4, "Jac", 0.8, 1, post, cyc, "GMRES",
                                coarseIters=6)
            hp = pkg.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
            A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
            if slabs:
                A.slabs = {"mode": "local", "devices": [0] * slabs}
            hd = pkg.api._ensure_hierarchy(A, 0)
        Z = None
        if not slabs:
            Z = np.empty_like(B, order="F")
            pkg._lib.check(hd.lib.hh_cycle(hd.h, B.ctypes.data, Z.ctypes.data, nrhs), hd.h)
        X, A = pkg.solveLinearSystem(None, B, A)
        out[(flag, slabs)] = (Z, np.reshape(X, (N, nrhs)).copy(), A.iterations.copy())
        pkg.clear(MG)
    assert rel_err(out[("1", 0)][0], out[("0", 0)][0]) < tol
    assert np.array_equal(out[("1", 0)][2], out[("0", 0)][2])
    assert rel_err(out[("1", 0)][1], out[("0", 0)][1]) < 100 * tol
    if prec == np.complex128:
        assert np.array_equal(out[("1", 3)][2], out[("0", 0)][2])
    assert rel_err(out[("1", 3)][1], out[("0", 0)][1]) < (1e-9 if prec == np.complex128 else 5e-3)
