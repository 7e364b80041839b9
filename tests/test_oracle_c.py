"""The C/OpenMP port of the oracle (the timed CPU arm of bench.py) is the same algorithm as the Python
oracle: identical operator, cycle and iteration counts."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err

sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.mark.parametrize("nodes,cycle", [([33, 17], "V"), ([17, 25, 17], "V"), ([17, 17, 17], "W")])
def test_c_port_matches_python_oracle(ho, nodes, cycle):
    import oracle_c

    rng = np.random.default_rng(3)
    domain = sum([[0.0, 0.1 * (n - 1)] for n in nodes], [])
    mesh = ho.getRegularMesh(domain, np.array(nodes) - 1)
    v = rng.uniform(1.5, 3.0, size=nodes)
    m = 1 / v**2
    w = ho.getMaximalFrequency(m, mesh)
    pad = [max(2, n // 8) for n in nodes]
    H, gamma = ho.GetHelmholtzOperatorABL(mesh, m, w, 0.01 * w * np.ones(nodes), True, pad, w, True)
    SH = H + ho.GetHelmholtzShiftOP(m, w, 0.2)
    MGo = ho.getMGparam(3, 1, 30, 1e-6, "Jac", 0.8, 2, 2, cycle, "GMRES", 10)
    ho.MGsetup(SH, nodes, MGo)
    oc = oracle_c.OracleC(nodes, mesh.h, m, gamma, w, True, True, 0.2, 3, 0.8, 2, 2, cycle, 10)
    N = int(np.prod(nodes))
    B = rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))
    assert rel_err(oc.apply(B), H @ B) < 1e-14
    assert rel_err(oc.apply(B, True), SH @ B) < 1e-14
    assert rel_err(oc.cycle(B), ho.MGcycle(MGo, B)) < 1e-12
    q, _ = ho.getAcousticPointSource(mesh)
    hp = ho.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
    A = ho.getShiftedLaplacianMultigridSolver(hp, MGo, 0.2, "GMRES", 5)
    xo, A = ho.solveLinearSystem(SH.conj().T, q, A)
    X, it, rr, secs = oc.solve(q)
    assert list(it) == A.iters
    assert rel_err(X[:, 0], xo) < 1e-10
    # bounded sample: stops after max_prec preconditioner applications
    X2, it2, rr2, _ = oc.solve(np.stack([q, 2 * q], axis=1), max_prec=3)
    assert list(it2) == [3, 3]
