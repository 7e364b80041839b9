"""GetHelmholtzOperatorHO (src/GetHelmholtz.jl:54-72, src/PlainNodalLaplacian.jl:49-141): the oracle's Kronecker
restatement, its known-answer anchors, and the library's host-side stencil form of the same operator (hh_ho_stencil).
The reference holds no test of this operator; what pins the restatement:
  * beta = 1 collapses to the reference's plain operator with first-order Neumann rows (Lap = G'G, M = I), i.e.
    GetHelmholtzOperator(..., orderNeumannBC=1) without the Sommerfeld BC scaling -- an identity between two separately
    restated code paths;
  * structure: Lap annihilates constants and is symmetric, M has unit row sums.
CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp


def _problem(ho, nodes, seed=2):
    rng = np.random.default_rng(seed)
    nodes = np.array(nodes)
    dom = []
    for d, nd in enumerate(nodes):
        dom += [0.0, (0.1 + 0.02 * d) * (nd - 1)]
    mesh = ho.getRegularMesh(dom, list(nodes - 1))
    m = 1.0 / (1.5 + rng.random(tuple(nodes))) ** 2
    w = 0.8 * ho.getMaximalFrequency(m, mesh)
    gamma = 0.05 * w * (1.0 + rng.random(tuple(nodes)))
    return mesh, m, w, gamma


@pytest.mark.parametrize("nodes", [(9, 7), (6, 5, 4)])
def test_ho_beta_one_is_the_plain_first_order_neumann_operator(ho, nodes):
    mesh, m, w, gamma = _problem(ho, nodes)
    Lap, M = ho.getSpreadNodalLaplacianAndMass(mesh, 1.0)
    assert abs(M - sp.identity(M.shape[0])).max() == 0.0
    assert abs(Lap - ho.getNodalLaplacianMatrix(mesh, 1)).max() < 1e-12
    H1 = ho.GetHelmholtzOperatorHO(mesh, m, w, gamma, True, False, 1.0)
    H0 = ho.GetHelmholtzOperator(mesh, m, w, gamma, True, False, 1)
    assert abs(H1 - H0).max() < 1e-12


@pytest.mark.parametrize("nodes,beta", [((9, 7), 2.0 / 3.0), ((9, 7), 5.0 / 6.0), ((6, 5, 4), [0.7, 0.9]), ((5, 4, 6), [1.0, 0.5])])
def test_ho_structure(ho, nodes, beta):
    mesh, m, w, gamma = _problem(ho, nodes)
    Lap, M = ho.getSpreadNodalLaplacianAndMass(mesh, beta)
    one = np.ones(Lap.shape[0])
    assert np.abs(Lap @ one).max() < 1e-9          # constants are in the null space (pure Neumann)
    assert abs(Lap - Lap.T).max() < 1e-12          # G' Gs with the symmetric spreading
    assert np.abs(M @ one - 1.0).max() < 1e-13     # averaging: unit row sums
    # footprint: 9 points in 2-D; in 3-D the Laplacian has no corner couplings (19 points), the mass 7
    dim = len(nodes)
    nnz_row = np.diff(Lap.tocsr().indptr).max()
    spread = (beta if np.isscalar(beta) else beta[0]) != 1.0
    assert nnz_row == ((9 if dim == 2 else 19) if spread else (5 if dim == 2 else 7))


@pytest.mark.parametrize("nodes,beta,neumann,somm,complex_w",
                         [((9, 7), 2.0 / 3.0, True, True, False), ((9, 7), 1.0, False, True, False), ((8, 6), 0.6, True, False, True),
                          ((6, 5, 4), [0.7, 0.9], True, True, False), ((5, 4, 6), 1.0, False, True, False),
                          ((4, 6, 5), [0.5, 0.8], True, False, True)])
def test_library_ho_stencil_matches_the_oracle(pkg, ho, nodes, beta, neumann, somm, complex_w):
    """hh_ho_stencil (tensor-product form, host C++) == the Kronecker assembly, entry by entry"""
    mesh, m, w, gamma = _problem(ho, nodes)
    if complex_w:
        w = w * (1.0 - 0.03j)
    H = ho.GetHelmholtzOperatorHO(mesh, m, w, gamma, neumann, somm, beta)
    want = ho.csr_to_stencil(H, mesh.nodes)
    pmesh = pkg.getRegularMesh(list(mesh.domain), list(mesh.n))
    got = pkg.GetHelmholtzOperatorHOStencil(pmesh, m, w, gamma, neumann, somm, beta)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_library_ho_stencil_rejects_bad_input(pkg):
    mesh = pkg.getRegularMesh([0, 1, 0, 1, 0, 1], [4, 4, 4])
    m = np.ones((5, 5, 5))
    with pytest.raises(ValueError):
        pkg.GetHelmholtzOperatorHOStencil(mesh, m, 1.0, 0 * m, True, True, 0.7)   # 3-D needs (beta_Lap, beta_mass)
    with pytest.raises(ValueError):
        pkg.GetHelmholtzOperatorHOStencil(mesh, m[:4], 1.0, 0 * m, True, True, 1.0)


@pytest.mark.parametrize("nodes,beta", [((9, 7), 2.0 / 3.0), ((6, 5, 4), [0.7, 0.9])])
def test_host_operator_object_matches_the_oracle_matrix(pkg, ho, nodes, beta):
    """GetHelmholtzOperatorHO(...) of the host mirror: H @ x, (H + shift) @ x and the adjoint view from the stored stencil"""
    mesh, m, w, gamma = _problem(ho, nodes)
    pmesh = pkg.getRegularMesh(list(mesh.domain), list(mesh.n))
    H = ho.GetHelmholtzOperatorHO(mesh, m, w, gamma, True, True, beta)
    SH = H + ho.GetHelmholtzShiftOP(m, w, 0.2)
    Hp = pkg.GetHelmholtzOperatorHO(pmesh, m, w, gamma, True, True, beta)
    SHp = Hp + pkg.GetHelmholtzShiftOP(m, w, 0.2)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((H.shape[0], 2)) + 1j * rng.standard_normal((H.shape[0], 2))
    scale = np.abs(H @ x).max()
    assert np.abs(Hp @ x - H @ x).max() < 1e-13 * scale
    assert np.abs(SHp @ x - SH @ x).max() < 1e-13 * scale
    assert np.abs(SHp.H @ x - SH.conj().T @ x).max() < 1e-13 * scale
    assert np.abs(Hp @ x[:, 0] - H @ x[:, 0]).max() < 1e-13 * scale
    hp = pkg.HelmholtzParam(pmesh, gamma, m.ravel(order="F"), w, True, True)
    assert np.abs(pkg.GetHelmholtzOperatorHO(hp, beta) @ x - H @ x).max() < 1e-13 * scale


@pytest.mark.parametrize("nodes,beta", [((9, 7), 2.0 / 3.0), ((6, 5, 4), [0.7, 0.9])])
def test_stencil_adjoint_matches_the_oracle_matrix(pkg, ho, nodes, beta):
    """hh_stencil_adjoint: how the transposed hierarchy of the high-order operator is formed (doTranspose = 1)"""
    mesh, m, w, gamma = _problem(ho, nodes)
    SH = ho.GetHelmholtzOperatorHO(mesh, m, w, gamma, True, True, beta) + ho.GetHelmholtzShiftOP(m, w, 0.2)
    want = ho.csr_to_stencil(SH.conj().T.tocsr(), mesh.nodes)
    got = pkg.stencilAdjoint(mesh.nodes, ho.csr_to_stencil(SH.tocsr(), mesh.nodes))
    assert np.abs(got - want).max() <= 1e-15 * np.abs(want).max()
    # an involution
    assert np.array_equal(pkg.stencilAdjoint(mesh.nodes, got), ho.csr_to_stencil(SH.tocsr(), mesh.nodes))


@pytest.mark.parametrize("nodes", [(9, 7), (6, 5, 4)])
def test_assembled_sparse_matrix_matches_the_oracle(pkg, ho, nodes):
    """hh_assemble_csc: the SparseMatrixCSC the reference's GetHelmholtzOperator / GetHelmholtzOperatorHO return (for
    callers that use the matrix itself, e.g. H \\ q in test/HelmholtzTest.jl:42): same pattern, same entries"""
    import scipy.sparse.linalg as spla

    mesh, m, w, gamma = _problem(ho, nodes)
    pmesh = pkg.getRegularMesh(list(mesh.domain), list(mesh.n))
    for neumann in (True, False):
        for order in (1, 2):
            for somm in (True, False):
                H = (ho.GetHelmholtzOperator(mesh, m, w, gamma, neumann, somm, order) + ho.GetHelmholtzShiftOP(m, w, 0.2)).tocsc()
                Hp = pkg.GetHelmholtzMatrix(pmesh, m, w, gamma, neumann, somm, order, 0.2)
                assert Hp.nnz == H.nnz and Hp.has_sorted_indices
                assert abs(Hp - H).max() <= 1e-15 * abs(H).max()
    beta = 2.0 / 3.0 if len(nodes) == 2 else [0.7, 0.9]
    wc = w * (1.0 - 0.02j)
    H = ho.GetHelmholtzOperatorHO(mesh, m, wc, gamma, True, True, beta).tocsc()
    Hp = pkg.GetHelmholtzMatrix(pmesh, m, wc, gamma, True, True, 2, 0.0, beta)
    assert Hp.nnz == H.nnz and abs(Hp - H).max() <= 1e-15 * abs(H).max()
    # the direct solve a caller would do with it
    q = np.zeros(H.shape[0], dtype=complex)
    q[H.shape[0] // 2] = 1.0
    assert np.allclose(spla.spsolve(Hp, q), spla.spsolve(H, q), rtol=1e-10, atol=0)
