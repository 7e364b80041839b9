"""The solver on GetHelmholtzOperatorHO (src/GetHelmholtz.jl:54-72; the reference builds its hierarchy from whatever
shifted matrix solveLinearSystem is handed, src/ShiftedLaplacianMultigridSolver.jl:65): device copies of the stored
stencil, Galerkin coarse stencil, operator apply and FGMRES / BiCGSTAB solves against the oracle's Kronecker-assembled
matrices (direct sparse solve and the oracle's CPU run of the same algorithm)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import rel_err

pytestmark = pytest.mark.gpu

CASES = {"2d": ((33, 17), 2.0 / 3.0), "3d": ((17, 13, 9), [0.7, 0.9])}


def _case(pkg, ho, name, prec=np.complex128, tol=1e-9, krylov="GMRES", inner=5, cyc_prec=None):
    nodes, beta = CASES[name]
    rng = np.random.default_rng(4)
    nodes = np.array(nodes)
    dom = []
    for d, nd in enumerate(nodes):
        dom += [0.0, (0.1 + 0.01 * d) * (nd - 1)]
    om = ho.getRegularMesh(dom, list(nodes - 1))
    pm = pkg.getRegularMesh(dom, list(nodes - 1))
    m = 1.0 / (1.5 + rng.random(tuple(nodes))) ** 2
    w = 0.8 * ho.getMaximalFrequency(m, om)
    gamma = 0.05 * w * (1.0 + rng.random(tuple(nodes)))
    shift = 0.2
    H = ho.GetHelmholtzOperatorHO(om, m, w, gamma, True, True, beta).tocsr()
    SH = (H + ho.GetHelmholtzShiftOP(m, w, shift)).tocsr()
    Hp = pkg.GetHelmholtzOperatorHO(pm, m, w, gamma, True, True, beta)
    SHp = Hp + pkg.GetHelmholtzShiftOP(m, w, shift)
    MG = pkg.getMGparam(prec, pkg.Int64, 2, 1, 40, tol, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
    if cyc_prec is not None:
        MG.cyclePrecision = cyc_prec
    hp = pkg.HelmholtzParam(pm, gamma, m.ravel(order="F"), w, True, True)
    A = pkg.getShiftedLaplacianMultigridSolver(hp, MG, shift, krylov, inner)
    N = int(np.prod(nodes))
    B = np.zeros((N, 3), dtype=np.complex128, order="F")
    B[pkg.loc2cs(nodes, [int(v) // 2 + 1 for v in nodes[:-1]] + [1]) - 1, 0] = 1.0 / pm.h[0] ** 2
    B[:, 1:] = rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))
    return dict(om=om, pm=pm, m=m, w=w, gamma=gamma, shift=shift, H=H, SH=SH, Hp=Hp, SHp=SHp, A=A, MG=MG, B=B, nodes=nodes,
                beta=beta)


@pytest.mark.parametrize("name", ["2d", "3d"])
def test_ho_device_stencils_and_apply(gpu_pkg, ho, name):
    pkg = gpu_pkg
    c = _case(pkg, ho, name)
    c["A"].operatorHO = list(c["Hp"].beta)
    hd = pkg.api._ensure_hierarchy(c["A"], 0)
    ns = 3 ** len(c["nodes"])
    N = int(np.prod(c["nodes"]))
    # level 0: the shifted fine stencil on the device == the host construction == the oracle's assembly
    lvl0 = np.empty(ns * N, dtype=np.complex128)
    pkg._lib.check(hd.lib.hh_get_level_stencil(hd.h, 0, lvl0.ctypes.data), hd.h)
    want0 = ho.csr_to_stencil(c["SH"], c["nodes"])
    assert rel_err(lvl0.reshape(ns, N), want0) < 1e-14
    # level 1: Galerkin R SH P
    P, nc = ho.getFWInterp(c["nodes"])
    R = P.T * (0.5 ** len(c["nodes"]))
    want1 = ho.csr_to_stencil((R @ c["SH"] @ P).tocsr(), nc)
    lvl1 = np.empty(want1.size, dtype=np.complex128)
    pkg._lib.check(hd.lib.hh_get_level_stencil(hd.h, 1, lvl1.ctypes.data), hd.h)
    assert rel_err(lvl1.reshape(want1.shape), want1) < 1e-13
    # operator apply: un-shifted (the Krylov operator) and with the hierarchy's shift
    rng = np.random.default_rng(1)
    X = np.asfortranarray(rng.standard_normal((N, 3)) + 1j * rng.standard_normal((N, 3)))
    Y = np.empty_like(X, order="F")
    pkg._lib.check(hd.lib.hh_apply(hd.h, X.ctypes.data, Y.ctypes.data, 3, 0, 0.0, 0), hd.h)
    assert rel_err(Y, c["H"] @ X) < 1e-13
    pkg._lib.check(hd.lib.hh_apply(hd.h, X.ctypes.data, Y.ctypes.data, 3, 1, c["shift"], 0), hd.h)
    assert rel_err(Y, c["SH"] @ X) < 1e-13
    assert rel_err(c["SHp"] @ X, c["SH"] @ X) < 1e-13   # the host-side operator object
    pkg.clear(c["MG"])


@pytest.mark.parametrize("name", ["2d", "3d"])
@pytest.mark.parametrize("krylov,inner", [("GMRES", 5), ("BiCGSTAB", 0)])
def test_ho_solve_matches_direct_solve_and_oracle_run(gpu_pkg, ho, name, krylov, inner):
    pkg = gpu_pkg
    c = _case(pkg, ho, name, krylov=krylov, inner=inner)
    X, A = pkg.solveLinearSystem(c["SHp"].H, c["B"], c["A"])   # spelled like test/ShiftedLaplacianTest.jl:83
    assert A.operatorHO == list(c["Hp"].beta)
    lu = spla.splu(c["H"].tocsc())
    for col in range(c["B"].shape[1]):
        assert rel_err(X[:, col], lu.solve(c["B"][:, col])) < 1e-6
    # the oracle's CPU run of the same algorithm on the same matrix: same iteration counts
    MGo = ho.getMGparam(2, 1, 40, 1e-9, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
    hpo = ho.HelmholtzParam(c["om"], c["gamma"], c["m"].ravel(order="F"), c["w"], True, True)
    Ao = ho.getShiftedLaplacianMultigridSolver(hpo, MGo, c["shift"], krylov, inner)
    Xo, Ao = ho.solveLinearSystem(c["SH"].conj().T, c["B"], Ao)
    assert np.abs(np.asarray(A.iterations, dtype=int) - np.asarray(Ao.iters, dtype=int)).max() <= (0 if krylov == "GMRES" else 1)
    assert rel_err(X, Xo) < 1e-7
    # back to the plain operator on the same solver object: the handle switches and the plain solve still agrees
    A.operatorHO = None
    Hplain = ho.GetHelmholtzOperator(c["om"], c["m"], c["w"], c["gamma"], True, True)
    Xp, A = pkg.solveLinearSystem(None, c["B"][:, 0], A)
    assert rel_err(Xp, spla.splu(Hplain.tocsc()).solve(c["B"][:, 0])) < 1e-6
    pkg.clear(c["MG"])


@pytest.mark.parametrize("variant", ["c32", "mixed"])
def test_ho_solve_reduced_precision_cycle(gpu_pkg, ho, variant):
    """ComplexF32 on an odd 3-D grid runs on the pitched internal layout; mixed = ComplexF64 Krylov on the stored H with
    the ComplexF32 companion's cycle"""
    pkg = gpu_pkg
    if variant == "c32":
        c = _case(pkg, ho, "3d", prec=np.complex64, tol=1e-5)
        tol = 1e-4
    else:
        c = _case(pkg, ho, "3d", tol=1e-9, cyc_prec=np.complex64)
        tol = 1e-6
    X, A = pkg.solveLinearSystem(c["SHp"].H, c["B"].astype(c["MG"].VAL), c["A"])
    lu = spla.splu(c["H"].tocsc())
    for col in range(c["B"].shape[1]):
        assert rel_err(X[:, col], lu.solve(c["B"][:, col])) < tol
    pkg.clear(c["MG"])
