"""doTranspose = 1 with GetHelmholtzOperatorHO (src/ShiftedLaplacianMultigridSolver.jl:68-70,78-80): the hierarchy of
the host-side adjoint stencils (checked on CPU in tests/test_oracle_ho_operator.py), the Krylov method on H^H."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from conftest import rel_err
from test_gpu_ho import _case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["2d", "3d"])
def test_ho_transposed_solve(gpu_pkg, ho, name):
    """doTranspose = 1 (src/ShiftedLaplacianMultigridSolver.jl:68-70,78-80): the hierarchy of the adjoint stencils, the
    Krylov method on H^H; against the direct solve with H^H and the oracle's transposed run"""
    pkg = gpu_pkg
    c = _case(pkg, ho, name)
    X, A = pkg.solveLinearSystem(c["SHp"].H, c["B"], c["A"], 1)
    lu = spla.splu(c["H"].conj().T.tocsc())
    for col in range(c["B"].shape[1]):
        assert rel_err(X[:, col], lu.solve(c["B"][:, col])) < 1e-6
    MGo = ho.getMGparam(2, 1, 40, 1e-9, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
    hpo = ho.HelmholtzParam(c["om"], c["gamma"], c["m"].ravel(order="F"), c["w"], True, True)
    Ao = ho.getShiftedLaplacianMultigridSolver(hpo, MGo, c["shift"], "GMRES", 5)
    Xo, Ao = ho.solveLinearSystem(c["SH"].conj().T, c["B"], Ao, 1)
    assert np.abs(np.asarray(A.iterations, dtype=int) - np.asarray(Ao.iters, dtype=int)).max() <= 1
    assert rel_err(X, Xo) < 1e-6
    # and back: the forward solve on the same solver object rebuilds the forward hierarchy
    X2, A = pkg.solveLinearSystem(c["SHp"].H, c["B"][:, 0], A, 0)
    assert rel_err(X2, spla.splu(c["H"].tocsc()).solve(c["B"][:, 0])) < 1e-6
    pkg.clear(c["MG"])
