"""Host-side mirror of the reference interface: names, argument orders and error behaviour
(src/ShiftedLaplacianMultigridSolver.jl, src/Helmholtz.jl) -- everything that needs no device."""
import numpy as np
import pytest


def _solver(pkg, krylov="GMRES", omega=3.0):
    mesh = pkg.getRegularMesh([0.0, 13.5, 0.0, 4.2], [16, 8])
    m = np.ones((17, 9))
    MG = pkg.getMGparam(pkg.ComplexF64, pkg.Int64, 2, 2, 30, 1e-6, "Jac", 0.75, 2, 2, "W", "NoMUMPS", 0.5, 0.0)
    hp = pkg.HelmholtzParam(mesh, np.zeros((17, 9)), m.ravel(order="F"), omega, True, True)
    return pkg.getShiftedLaplacianMultigridSolver(hp, MG, 0.02, krylov, 20, True), MG, hp


def test_constructor_forms_and_fields(pkg):
    A, MG, hp = _solver(pkg)
    assert list(A.shift) == [0.02, 0.02]            # scalar shift broadcast to ones(levels)*shift (:28-30)
    assert (A.Krylov, A.inner, A.doClear, A.verbose, A.setupTime, A.nPrec, A.solveTime) == ("GMRES", 20, 0, True, 0.0, 0, 0.0)
    A2 = pkg.getShiftedLaplacianMultigridSolver(hp, MG, np.array([0.2, 0.1]))
    assert A2.Krylov == "BiCGSTAB" and A2.inner == 5  # defaults of :24
    MG2 = pkg.getMGparam(5, 4, 50, 1e-5, "Jac-GMRES", 0.8, lambda l: l + 1, lambda l: l + 1, "K", "GMRES", 0.5, 0.0,
                         "FullWeighting")            # examples/PointSourceADR/runExperiments.jl:169-170 form
    o = MG2._options([0.2] * 5, 0)
    assert (o.levels, o.relax_type, o.cycle_type, o.coarse_type) == (5, 1, 2, 1)
    assert list(o.relax_pre)[:5] == [2, 3, 4, 5, 6] and o.relax_param == 0.8
    assert MG2.VAL == np.complex128
    MG3 = pkg.getMGparam(pkg.ComplexF32, pkg.Int64, 3, 1, 10, 1e-4, "Jac", 0.8, 2, 2, "V", "Julia")
    assert MG3.VAL == np.complex64 and MG3._options([0.1], 1).do_transpose == 1
    assert not pkg.hierarchyExists(MG)
    c = pkg.copySolver(A)
    assert c.MG is not A.MG and c.MG.levels == 2 and c.helmParam is A.helmParam and not pkg.hierarchyExists(c.MG)
    s = pkg.getShiftedHelmholtzParam(hp, 0.1)
    assert np.allclose(s.gamma, hp.gamma + 0.1 * 3.0)


def test_zero_rhs_returns_zeros_without_touching_the_device(pkg):
    """src/ShiftedLaplacianMultigridSolver.jl:40-43"""
    A, MG, hp = _solver(pkg)
    x, A = pkg.solveLinearSystem(None, np.zeros(17 * 9, dtype=complex), A)
    assert x.shape == (153,) and not np.any(x) and not pkg.hierarchyExists(MG)
    X, A = pkg.solveLinearSystem(None, np.zeros((153, 3), dtype=complex), A)
    assert X.shape == (153, 3)
    X1, A = pkg.solveLinearSystem(None, np.zeros((153, 1), dtype=complex), A)
    assert X1.shape == (153,)                        # N x 1 is flattened (:34-36)


def test_error_behaviour(pkg):
    A, MG, hp = _solver(pkg, omega=3.0 - 0.1j)
    with pytest.raises(TypeError):                   # GetHelmholtzShiftOP(m, omega::Float64, shift), :77
        pkg.solveLinearSystem(None, np.ones(153, dtype=complex), A)
    with pytest.raises(TypeError):
        pkg.GetHelmholtzShiftOP(np.ones(3), 1.0 + 1j, 0.1)
    A, MG, hp = _solver(pkg, krylov="CG")
    with pytest.raises(ValueError):
        pkg.solveLinearSystem(None, np.ones(153, dtype=complex), A)
    MG.relaxType = "VankaFaces"                      # elastic-only smoother: out of scope
    with pytest.raises(ValueError):
        MG._options([0.1], 0)
    with pytest.raises(TypeError):
        pkg.clear(object())


def test_workloads_are_deterministic(pkg):
    a = pkg.workloads.config4(n=17, sigma=2.0, seed=1234, pad=2)["m"]
    b = pkg.workloads.config4(n=17, sigma=2.0, seed=1234, pad=2)["m"]
    assert np.array_equal(a, b) and a.shape == (17, 17, 17)
    v = 1 / np.sqrt(a)
    assert abs(v.min() - 1.5) < 1e-12 and abs(v.max() - 4.5) < 1e-12
    s = pkg.workloads.point_sources_top_grid([257, 257, 257], 16, 16)
    assert len(s) == 256 and len({tuple(x) for x in s}) == 256 and all(x[2] == 1 for x in s)
    c3 = pkg.workloads.config3(33)
    assert c3["m"].shape == (33, 33, 33) and np.all(np.diff(1 / np.sqrt(c3["m"][0, 0, :])) >= 0)
    import os
    from conftest import GOLDEN
    vp = np.load(os.path.join(GOLDEN, "seg_salt_vp.npz"))["vp_ms"]
    c2 = pkg.workloads.config2(vp)
    assert c2["m"].shape == (257, 129) and abs(1 / np.sqrt(c2["m"].max()) - 1.5) < 1e-9


def test_remaining_reference_exports_on_the_host(pkg, ho):
    """getSommerfeldBC, getNodalLaplacianMatrix, dxxMat, getHelmholtzFun (src/GetHelmholtz.jl:2, PlainNodalLaplacian.jl:1):
    host-side mirrors against the oracle's restatements; no device needed (the closure is exercised on the high-order
    operator object, whose product is evaluated on the host)."""
    rng = np.random.default_rng(9)
    for nodes in ((9, 7), (6, 5, 4)):
        nodes = np.array(nodes)
        dom = sum([[0.0, (0.1 + 0.03 * d) * (nd - 1)] for d, nd in enumerate(nodes)], [])
        om, pm = ho.getRegularMesh(dom, list(nodes - 1)), pkg.getRegularMesh(dom, list(nodes - 1))
        m = 1.0 / (1.5 + rng.random(tuple(nodes))) ** 2
        w = 3.0
        for neumann in (True, False):
            for order in (1, 2):
                assert np.abs(pkg.getSommerfeldBC(pm, m, w, neumann, order) - ho.getSommerfeldBC(om, m, w, neumann, order)).max() < 1e-13
        for order in (1, 2):
            Lp, Lo = pkg.getNodalLaplacianMatrix(pm, order), ho.getNodalLaplacianMatrix(om, order)
            assert Lp.nnz == Lo.nnz and abs(Lp - Lo).max() < 1e-12 * abs(Lo).max()
            assert abs(pkg.dxxMat(7, 0.3, order) - ho.dxxMat(7, 0.3, order)).max() < 1e-13
        # getHelmholtzFun as the solver uses it: Afun(x) = SH' ' x - ShiftOP x = H x
        gamma = 0.05 * w * (1.0 + rng.random(tuple(nodes)))
        beta = 2.0 / 3.0 if len(nodes) == 2 else [0.7, 0.9]
        Hp = pkg.GetHelmholtzOperatorHO(pm, m, w, gamma, True, True, beta)
        shiftop = pkg.GetHelmholtzShiftOP(m, w, 0.2)
        SHT = (Hp + shiftop).H
        x = rng.standard_normal((Hp.shape[0], 2)) + 1j * rng.standard_normal((Hp.shape[0], 2))
        Afun = pkg.getHelmholtzFun(SHT, -shiftop)
        Ho_ = ho.GetHelmholtzOperatorHO(om, m, w, gamma, True, True, beta)
        assert np.abs(Afun(x) - Ho_ @ x).max() < 1e-12 * np.abs(Ho_ @ x).max()
        y = np.zeros_like(x, order="F")
        assert pkg.getHelmholtzFun(SHT, -1.0 * shiftop, y, 4)(x) is y and np.abs(y - Ho_ @ x).max() < 1e-12 * np.abs(y).max()


def test_multOpNeumann_is_the_first_order_neumann_laplacian(pkg, ho):
    """src/PlainNodalLaplacian.jl:150-188 == getNodalLaplacianMatrix(M, 1) * x (the reference's matrix-free precedent)"""
    rng = np.random.default_rng(1)
    om, pm = ho.getRegularMesh([0, 1.2, 0, 0.7], [11, 8]), pkg.getRegularMesh([0, 1.2, 0, 0.7], [11, 8])
    x = rng.standard_normal(12 * 9)
    y = np.zeros_like(x)
    assert pkg.multOpNeumann_(pm, x, y, pkg.Lap2DStencil) is y
    assert np.abs(y - ho.getNodalLaplacianMatrix(om, 1) @ x).max() < 1e-11
    m3 = pkg.getRegularMesh([0, 1, 0, 1, 0, 1], [2, 2, 2])
    z = np.ones(27)
    assert np.array_equal(pkg.multOpNeumann_(m3, np.ones(27), z), np.ones(27))  # the reference's 3-D branch is empty
