"""K1 parity: the matrix-free stencil (hh_apply through the C ABI) against the oracle's Kronecker
assembly of the reference operator (src/GetHelmholtz.jl:33-50, src/PlainNodalLaplacian.jl:32-46)."""
import itertools

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

# tolerance: ComplexF64 1e-13, ComplexF32 2e-5 relative
TOL = {np.complex128: 1e-13, np.complex64: 2e-5}


def _problem(ho, nodes, seed):
    rng = np.random.default_rng(seed)
    dim = len(nodes)
    domain = sum([[0.0, 1.0 + 0.37 * d] for d in range(dim)], [])
    mesh = ho.getRegularMesh(domain, np.array(nodes) - 1)
    m = rng.uniform(0.2, 1.0, size=nodes)
    g = rng.uniform(0.0, 1.0, size=nodes)
    return mesh, m, g, rng


@pytest.mark.parametrize("nodes", [[37, 21], [70, 33], [19, 13, 11], [40, 9, 35]])
@pytest.mark.parametrize("prec", [np.complex128, np.complex64])
def test_apply_matches_kronecker_assembly(gpu_pkg, ho, nodes, prec):
    pkg = gpu_pkg
    mesh, m, g, rng = _problem(ho, nodes, 11)
    pmesh = pkg.getRegularMesh(mesh.domain, mesh.n)
    N = int(np.prod(nodes))
    for neu, somm, order, omega in itertools.product((True, False), (True, False), (1, 2), (3.1, 3.1 - 0.4j)):
        H = ho.GetHelmholtzOperator(mesh, m, omega, g, neu, somm, order)
        Hp = pkg.GetHelmholtzOperator(pmesh, m, omega, g, neu, somm, order, precision=prec)
        for nrhs in (1, 3):
            x = (rng.standard_normal((N, nrhs)) + 1j * rng.standard_normal((N, nrhs))).astype(prec)
            y = Hp @ x
            assert y.shape == x.shape and y.dtype == np.dtype(prec)
            assert rel_err(y, H @ x.astype(np.complex128)) < TOL[prec]
        # shifted operator H + i*s*w^2*diag(m)  (GetHelmholtzShiftOP, src/GetHelmholtz.jl:81-83)
        if not np.iscomplexobj(omega):
            SH = H + ho.GetHelmholtzShiftOP(m, omega, 0.2)
            SHp = Hp + pkg.GetHelmholtzShiftOP(m, omega, 0.2)
            x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(prec)
            assert rel_err(SHp @ x, SH @ x.astype(np.complex128)) < TOL[prec]
            # adjoint (doTranspose = 1 operator)
            assert rel_err(SHp.H @ x, SH.conj().T @ x.astype(np.complex128)) < TOL[prec]
            # operator identity of getHelmholtzFun (GetHelmholtz.jl:85-95): SH x - i s w^2 m x = H x
            mv = m.ravel(order="F")
            lhs = (SHp @ x).astype(np.complex128) - 1j * 0.2 * omega**2 * mv * x.astype(np.complex128)
            assert rel_err(lhs, H @ x.astype(np.complex128)) < 10 * TOL[prec]


def test_apply_many_rhs_and_diagonal(gpu_pkg, ho):
    pkg = gpu_pkg
    nodes = [33, 17, 9]
    mesh, m, g, rng = _problem(ho, nodes, 5)
    pmesh = pkg.getRegularMesh(mesh.domain, mesh.n)
    N = int(np.prod(nodes))
    w = 2.7
    H = ho.GetHelmholtzOperator(mesh, m, w, g, True, True)
    Hp = pkg.GetHelmholtzOperator(pmesh, m, w, g, True, True)
    x = rng.standard_normal((N, 16)) + 1j * rng.standard_normal((N, 16))
    assert rel_err(Hp @ x, H @ x) < 1e-13
    # the diagonal (mass + Sommerfeld) the kernels compute on the fly equals GetHelmholtz.jl:41-47
    d = ho.helmholtz_diagonal(mesh, m, w, g, True, True)
    assert rel_err(Hp.diagonal_mass(), d) < 1e-14


def test_apply_error_paths(gpu_pkg):
    pkg = gpu_pkg
    mesh = pkg.getRegularMesh([0, 1, 0, 1], [8, 8])
    m = np.ones((9, 9))
    with pytest.raises(pkg._lib.HelmholtzB200Error):
        pkg.GetHelmholtzOperator(mesh, m, 1.0, np.zeros((9, 9)), True, True, 3)  # BC order not supported
    with pytest.raises(ValueError):
        pkg.GetHelmholtzOperator(mesh, np.ones(5), 1.0, np.zeros((9, 9)), True, True)
