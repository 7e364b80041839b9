"""The oracle's solver against the implicit contract of the reference's tests and against direct solves."""
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import GOLDEN, rel_err


@pytest.fixture(scope="module")
def config1(ho, pkg):
    cfg = pkg.workloads.config1()
    m = cfg["m"]
    mesh = ho.getRegularMesh(cfg["domain"], cfg["n_cells"])
    w = 2 * np.pi * cfg["f"]
    H, gamma = ho.GetHelmholtzOperatorABL(mesh, m, w, w * np.ones(m.shape) * 0.01, True, cfg["pad"],
                                          ho.getMaximalFrequency(m, mesh), True)
    SH = H + ho.GetHelmholtzShiftOP(m, w, 0.02)
    n = mesh.nodes
    q = np.zeros(int(np.prod(n)), dtype=complex)
    q[ho.loc2cs(n, [n[0] // 2, 1]) - 1] = 1.0 / mesh.h[0] ** 2
    hp = ho.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
    return dict(H=H, SH=SH, q=q, hp=hp, xt=spla.splu(H.tocsc()).solve(q))


def test_config1_converges_within_30_outer_iterations(ho, config1):
    """test/ShiftedLaplacianTest.jl:51-61,78-89,126-135: GMRES(20) and BiCGSTAB, 1 and 2 RHS, tol 1e-6, maxIter 30."""
    c = config1
    MG = ho.getMGparam(2, 2, 30, 1e-6, "Jac", 0.75, 2, 2, "W", "NoMUMPS")
    A = ho.getShiftedLaplacianMultigridSolver(c["hp"], MG, 0.02, "GMRES", 20)
    A = ho.copySolver(A)
    x, A = ho.solveLinearSystem(c["SH"].conj().T, c["q"], A)
    assert np.linalg.norm(c["H"] @ x - c["q"]) / np.linalg.norm(c["q"]) < 1e-6
    assert A.iters[0] == 15  # SURVEY.md section 6 [PROBE]: 15 preconditioner applications
    assert 1e-6 < rel_err(x, c["xt"]) < 1e-5  # 1e-6 residual <-> 2e-6..8e-6 solution error
    A2 = ho.getShiftedLaplacianMultigridSolver(c["hp"], MG, 0.02, "BiCGSTAB", 0)
    x2, A2 = ho.solveLinearSystem(c["SH"].conj().T, c["q"], A2)
    assert np.linalg.norm(c["H"] @ x2 - c["q"]) / np.linalg.norm(c["q"]) < 1e-6
    rng = np.random.default_rng(0)
    b = rng.random((c["q"].size, 2)) + 1j * rng.random((c["q"].size, 2))
    ho.clear(A2)
    x3, A2 = ho.solveLinearSystem(c["SH"].conj().T, b, A2)
    assert x3.shape == b.shape and np.linalg.norm(c["H"] @ x3 - b) / np.linalg.norm(b) < 1e-6
    assert max(A2.iters) < 2 * 30


def test_config1_kcycle_jac_gmres(ho, config1):
    """test/ShiftedLaplacianTest.jl:137-142"""
    c = config1
    MG = ho.getMGparam(2, 2, 30, 1e-6, "Jac-GMRES", 0.75, 2, 2, "K", "Julia")
    A = ho.getShiftedLaplacianMultigridSolver(c["hp"], MG, 0.02, "GMRES", 5)
    x, A = ho.solveLinearSystem(c["SH"].conj().T, c["q"], A)
    assert np.linalg.norm(c["H"] @ x - c["q"]) / np.linalg.norm(c["q"]) < 1e-6


def test_zero_rhs_and_transpose(ho, config1):
    c = config1
    MG = ho.getMGparam(2, 2, 30, 1e-6, "Jac", 0.75, 2, 2, "W", "NoMUMPS")
    A = ho.getShiftedLaplacianMultigridSolver(c["hp"], MG, 0.02, "GMRES", 20)
    x, _ = ho.solveLinearSystem(c["SH"].conj().T, np.zeros_like(c["q"]), A)
    assert not np.any(x)
    y, A = ho.solveLinearSystem(c["SH"].conj().T, c["q"], A, 1)
    Ht = c["H"].conj().T
    assert np.linalg.norm(Ht @ y - c["q"]) / np.linalg.norm(c["q"]) < 1e-6


def test_solver_goldens_regression(ho):
    G = np.load(os.path.join(GOLDEN, "oracle_goldens.npz"))
    for name in ("2d", "3d"):
        nodes = [int(v) for v in G[f"{name}_nodes"]]
        domain = sum([[0.0, 0.1 * (n - 1)] for n in nodes], [])
        mesh = ho.getRegularMesh(domain, np.array(nodes) - 1)
        w = float(G[f"{name}_w"])
        H = ho.GetHelmholtzOperator(mesh, G[f"{name}_m"], w, G[f"{name}_gamma"], True, True)
        SH = H + ho.GetHelmholtzShiftOP(G[f"{name}_m"], w, 0.2)
        MG = ho.getMGparam(3, 1, 30, 1e-8, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
        ho.MGsetup(SH, nodes, MG)
        assert rel_err(ho.csr_to_stencil(MG.As[2], MG.nodes[2]), G[f"{name}_stencil_l2"]) < 1e-13
        assert rel_err(ho.MGcycle(MG, G[f"{name}_x"]), G[f"{name}_cycle"]) < 1e-12
        hp = ho.HelmholtzParam(mesh, G[f"{name}_gamma"], G[f"{name}_m"].ravel(order="F"), w, True, True)
        A = ho.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
        xs, A = ho.solveLinearSystem(SH.conj().T, G[f"{name}_q"], A)
        assert A.iters == list(G[f"{name}_solve_iters"])
        assert rel_err(xs, G[f"{name}_solve_fgmres5"]) < 1e-10
        assert rel_err(xs, G[f"{name}_direct"]) < 1e-7  # tol 1e-8 solve vs the direct solve
