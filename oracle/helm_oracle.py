"""CPU oracle for the shifted-Laplacian multigrid Helmholtz solve path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`helmholtz.jl_b200/`, the
C-ABI library) may import, call or link this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs use it, and only as the checker / the timed CPU arm.

It is a numpy/scipy restatement of the reference algorithm
(JuliaInv/Helmholtz.jl, paths relative to /root/reference):

* operator assembly           src/GetHelmholtz.jl:22-50, 81-83, 97-247
                              src/PlainNodalLaplacian.jl:4-46
* solver control flow         src/ShiftedLaplacianMultigridSolver.jl:33-102
* point sources / indexing    src/getPointSource.jl:63-112
* settings                    test/ShiftedLaplacianTest.jl:15-78, 126-141

PARITY STATUS: **parity unpinned** at the Multigrid/Krylov boundary.  The
multigrid cycle and the Krylov methods live in un-vendored Julia packages
(Multigrid.jl v0.8.0 tree 70508e00..., KrylovMethods.jl v0.6.0 tree ceb12d55...,
ParSpMatVec.jl v0.1.1, jInv.jl v1.0.0 -- Manifest.toml:46-52,91-97,106-110,
168-172) that are absent from /root/reference, no Julia runtime exists in this
image, and the reference's tests hold no assertion or golden vector
(SURVEY.md section 4).  What *is* pinned here:
  - the operator: the matrix-free formula is checked against the reference's
    own Kronecker assembly (restated line by line below) and against the
    known-answer tests the reference ships (manufactured solution
    test/testFictitiousSource2D.jl:13-57, attenuation equivalence
    test/AttenuationTest.jl:31-47, operator identity GetHelmholtz.jl:85-95);
  - the solve: against sparse direct solves `H\\q` (the reference's own notion
    of truth, test/HelmholtzTest.jl:42,52) and against the implicit contract of
    test/ShiftedLaplacianTest.jl (converges below 1e-6 within 30 outer
    iterations for GMRES(20) and BiCGSTAB, 1 and 2 RHS).
The MG/Krylov part restates the published algorithms those packages implement
(geometric Galerkin multigrid with full weighting / linear interpolation and
damped Jacobi; right-preconditioned flexible GMRES with modified Gram-Schmidt;
preconditioned BiCGSTAB).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------
# Mesh (jInv.Mesh.RegularMesh: only n (cells), h, dim, domain are used)
# --------------------------------------------------------------------------


@dataclass
class RegularMesh:
    domain: np.ndarray  # [x0,x1,y0,y1(,z0,z1)]
    n: np.ndarray  # cells per dimension

    @property
    def dim(self):
        return len(self.n)

    @property
    def h(self):
        d = np.asarray(self.domain, dtype=np.float64)
        return (d[1::2] - d[0::2]) / np.asarray(self.n, dtype=np.float64)

    @property
    def nodes(self):
        return np.asarray(self.n, dtype=np.int64) + 1


def getRegularMesh(domain, n):
    return RegularMesh(np.asarray(domain, dtype=np.float64).ravel(), np.asarray(n, dtype=np.int64).ravel())


# --------------------------------------------------------------------------
# Nodal Laplacian (src/PlainNodalLaplacian.jl:4-46)
# --------------------------------------------------------------------------


def getBC(orderNeumannBC=2):
    """PlainNodalLaplacian.jl:4-15"""
    if orderNeumannBC == 2:
        return 2.0
    if orderNeumannBC == 1:
        return 1.0
    raise ValueError("getNodalLaplacianMatrix: BC not supported")


def dxxMat(n, h, orderNeumannBC=2):
    """PlainNodalLaplacian.jl:18-30 -- 1-D -d^2/dx^2 with Neumann ghost elimination."""
    BC = getBC(orderNeumannBC)
    O1 = -np.ones(n - 1)
    O1[n - 2] = -BC
    O2 = 2.0 * np.ones(n)
    O2[0] = BC
    O2[n - 1] = BC
    O3 = -np.ones(n - 1)
    O3[0] = -BC
    return sp.diags([O1 / h**2, O2 / h**2, O3 / h**2], [-1, 0, 1], format="csc")


def getNodalLaplacianMatrix(mesh: RegularMesh, orderNeumannBC=2):
    """PlainNodalLaplacian.jl:32-46 (Kronecker assembly, column-major node order)."""
    nodes = mesh.nodes
    h = mesh.h
    I1 = sp.identity(nodes[0], format="csc")
    D1 = dxxMat(nodes[0], h[0], orderNeumannBC)
    I2 = sp.identity(nodes[1], format="csc")
    D2 = dxxMat(nodes[1], h[1], orderNeumannBC)
    if mesh.dim == 2:
        L = sp.kron(I2, D1) + sp.kron(D2, I1)
    else:
        I3 = sp.identity(nodes[2], format="csc")
        D3 = dxxMat(nodes[2], h[2], orderNeumannBC)
        L = sp.kron(I3, sp.kron(I2, D1) + sp.kron(D2, I1)) + sp.kron(D3, sp.kron(I2, I1))
    return L.tocsc()


# --------------------------------------------------------------------------
# Absorbing layer, Sommerfeld, operator (src/GetHelmholtz.jl)
# --------------------------------------------------------------------------


def getABL(n, NeumannAtFirstDim, ABLpad, ABLamp, code=None):
    """GetHelmholtz.jl:97-220.  `n` = node counts.  Returns an array of shape tuple(n)
    (Fortran/column-major semantics: index [i1,i2(,i3)])."""
    n = [int(v) for v in np.asarray(n).ravel()]
    pad = [int(v) for v in np.asarray(ABLpad).ravel()]
    dim = len(n)
    if code is None:
        code = np.ones((dim, 2), dtype=bool)
    else:
        code = np.array(code, dtype=bool).copy()
    if dim == 2:
        # live branch impl == 1, GetHelmholtz.jl:141-163
        gamma = np.zeros((n[0], n[1]))
        b_bwd1 = (np.arange(pad[0], 0, -1, dtype=np.float64) ** 2) / pad[0] ** 2
        b_bwd2 = (np.arange(pad[1], 0, -1, dtype=np.float64) ** 2) / pad[1] ** 2
        b_fwd1 = (np.arange(1, pad[0] + 1, dtype=np.float64) ** 2) / pad[0] ** 2
        b_fwd2 = (np.arange(1, pad[1] + 1, dtype=np.float64) ** 2) / pad[1] ** 2
        I1 = slice(n[0] - pad[0], n[0])
        I2 = slice(n[1] - pad[1], n[1])
        if not NeumannAtFirstDim:
            gamma[:, : pad[1]] += np.outer(np.ones(n[0]), b_bwd2)
            gamma[: pad[0], : pad[1]] -= np.outer(b_bwd1, b_bwd2)
            gamma[I1, : pad[1]] -= np.outer(b_fwd1, b_bwd2)
        gamma[:, I2] += np.outer(np.ones(n[0]), b_fwd2)
        gamma[: pad[0], :] += np.outer(b_bwd1, np.ones(n[1]))
        gamma[I1, :] += np.outer(b_fwd1, np.ones(n[1]))
        gamma[: pad[0], I2] -= np.outer(b_bwd1, b_fwd2)
        gamma[I1, I2] -= np.outer(b_fwd1, b_fwd2)
        gamma *= ABLamp
        return gamma
    # 3-D branch, GetHelmholtz.jl:164-218
    x1 = np.linspace(-1.0, 1.0, n[0])
    x2 = np.linspace(-1.0, 1.0, n[1])
    x3 = np.linspace(0.0, 1.0, n[2])
    if NeumannAtFirstDim:
        code[2, 0] = False

    def prof(x, p, cL, cR):
        g = np.zeros_like(x)
        if cL:
            gl = (x - x[p - 1]) ** 2
            gl[p:] = 0.0
            g += gl
        if cR:
            gr = (x - x[len(x) - p]) ** 2
            gr[: len(x) - p] = 0.0
            g += gr
        return g / (g.max() + 1e-5)

    g1 = prof(x1, pad[0], code[0, 0], code[0, 1])
    g2 = prof(x2, pad[1], code[1, 0], code[1, 1])
    g3 = prof(x3, pad[2], code[2, 0], code[2, 1])
    gamma = g1[:, None, None] + g2[None, :, None] + g3[None, None, :]
    gamma = gamma * ABLamp
    gamma[gamma >= ABLamp] = ABLamp
    return gamma


def getSommerfeldBC(mesh: RegularMesh, mNodal, omega, NeumannOnTop, orderNeumannBC=2):
    """GetHelmholtz.jl:222-247.  Note the reference's caller (GetHelmholtz.jl:45) never
    forwards orderNeumannBC, so BC = 2 always on this path."""
    BC = getBC(orderNeumannBC)
    ntup = tuple(int(v) for v in mesh.nodes)
    Somm = np.zeros(ntup, dtype=np.complex128)
    m = np.asarray(mNodal, dtype=np.float64).reshape(ntup, order="F")
    h = mesh.h
    if mesh.dim == 2:
        if not NeumannOnTop:
            Somm[:, 0] += -1j * omega * (BC / h[1]) * np.sqrt(m[:, 0])
        Somm[:, -1] += (-1j * omega * (BC / h[1])) * np.sqrt(m[:, -1])
        Somm[-1, :] += (-1j * omega * (BC / h[0])) * np.sqrt(m[-1, :])
        Somm[0, :] += (-1j * omega * (BC / h[0])) * np.sqrt(m[0, :])
    else:
        if not NeumannOnTop:
            Somm[:, :, 0] += -1j * omega * (BC / h[2]) * np.sqrt(m[:, :, 0])
        Somm[:, :, -1] += -1j * omega * (BC / h[2]) * np.sqrt(m[:, :, -1])
        Somm[:, 0, :] += -1j * omega * (BC / h[1]) * np.sqrt(m[:, 0, :])
        Somm[:, -1, :] += -1j * omega * (BC / h[1]) * np.sqrt(m[:, -1, :])
        Somm[0, :, :] += -1j * omega * (BC / h[0]) * np.sqrt(m[0, :, :])
        Somm[-1, :, :] += -1j * omega * (BC / h[0]) * np.sqrt(m[-1, :, :])
    return Somm


def helmholtz_diagonal(mesh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld):
    """The diagonal `mass` of GetHelmholtz.jl:41-47 (complex, length N, column-major)."""
    m = np.asarray(mNodal, dtype=np.float64).ravel(order="F")
    g = np.asarray(gamma, dtype=np.float64).ravel(order="F")
    mass = -(omega**2) * m * (1.0 - 1j * g / np.real(omega))
    if Sommerfeld:
        somm = getSommerfeldBC(mesh, mNodal, float(np.real(omega)), NeumannAtFirstDim)
        mass = mass - somm.ravel(order="F")
    return mass


def GetHelmholtzOperator(mesh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld, orderNeumannBC=2):
    """GetHelmholtz.jl:33-50 -> complex CSC matrix H."""
    Lap = getNodalLaplacianMatrix(mesh, orderNeumannBC)
    mass = helmholtz_diagonal(mesh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld)
    return (Lap + sp.diags(mass, 0, format="csc")).tocsc()


def GetHelmholtzOperatorABL(mesh, mNodal, omega, gamma, NeumannAtFirstDim, ABLpad, ABLamp, Sommerfeld, orderNeumannBC=2):
    """GetHelmholtz.jl:22-31 -> (H, gamma_with_ABL)."""
    abl = getABL(mesh.nodes, NeumannAtFirstDim, ABLpad, ABLamp)
    if gamma is None or (hasattr(gamma, "__len__") and len(gamma) == 0):
        gamma = abl
    else:
        gamma = np.asarray(gamma, dtype=np.float64).reshape(abl.shape, order="F") + abl
    H = GetHelmholtzOperator(mesh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld, orderNeumannBC)
    return H, gamma


# --------------------------------------------------------------------------
# High-order ("spread") operator: GetHelmholtzOperatorHO (src/GetHelmholtz.jl:54-72) on the spread nodal
# Laplacian and mass of src/PlainNodalLaplacian.jl:49-141.  No test of the reference exercises it; the
# restatement follows the Kronecker construction line by line.
# --------------------------------------------------------------------------


def ddxCN(n, h):
    """PlainNodalLaplacian.jl:56-60 -- 1-D cell-centred derivative of nodal values, n x (n+1)."""
    return sp.diags([-np.ones(n) / h, np.ones(n) / h], [0, 1], shape=(n, n + 1), format="csc")


def av3term(n, alpha=5.0 / 6.0):
    """PlainNodalLaplacian.jl:62-68 -- three-term average, first and last diagonal entry 1/2 + alpha/2."""
    t = (1.0 - alpha) / 2.0
    T = sp.diags([t * np.ones(n - 1), alpha * np.ones(n), t * np.ones(n - 1)], [-1, 0, 1], format="lil")
    T[0, 0] = 0.5 + alpha / 2.0
    T[n - 1, n - 1] = 0.5 + alpha / 2.0
    return T.tocsc()


def getNodalSpreadGradients(mesh: RegularMesh, avFunc):
    """PlainNodalLaplacian.jl:71-104 -> (G, Gs)."""
    n = mesh.n
    h = mesh.h
    eye = lambda k: sp.identity(int(k), format="csc")
    if mesh.dim == 2:
        t = ddxCN(int(n[0]), h[0])
        D1 = sp.kron(eye(n[1] + 1), t)
        D1s = sp.kron(avFunc(int(n[1] + 1)), t)
        t = ddxCN(int(n[1]), h[1])
        D2 = sp.kron(t, eye(n[0] + 1))
        D2s = sp.kron(t, avFunc(int(n[0] + 1)))
        return sp.vstack([D1, D2]).tocsc(), sp.vstack([D1s, D2s]).tocsc()
    a1, a2, a3 = avFunc(int(n[0] + 1)), avFunc(int(n[1] + 1)), avFunc(int(n[2] + 1))
    I1, I2, I3 = eye(n[0] + 1), eye(n[1] + 1), eye(n[2] + 1)
    t = ddxCN(int(n[0]), h[0])
    D1 = sp.kron(I3, sp.kron(I2, t))
    D1s = 0.5 * (sp.kron(I3, sp.kron(a2, t)) + sp.kron(a3, sp.kron(I2, t)))
    t = ddxCN(int(n[1]), h[1])
    D2 = sp.kron(I3, sp.kron(t, I1))
    D2s = 0.5 * (sp.kron(I3, sp.kron(t, a1)) + sp.kron(a3, sp.kron(t, I1)))
    t = ddxCN(int(n[2]), h[2])
    D3 = sp.kron(t, sp.kron(I2, I1))
    D3s = 0.5 * (sp.kron(t, sp.kron(I2, a1)) + sp.kron(t, sp.kron(a2, I1)))
    return sp.vstack([D1, D2, D3]).tocsc(), sp.vstack([D1s, D2s, D3s]).tocsc()


def getSpreadNodalLaplacianAndMass(mesh: RegularMesh, beta):
    """PlainNodalLaplacian.jl:106-141 -> (Lap, M).  2-D: beta scalar; 3-D: beta[0] Laplacian, beta[1] mass
    (a scalar 1 means [1, 1])."""
    n = mesh.n
    eye = lambda k: sp.identity(int(k), format="csc")
    avFunc = lambda k: av3term(k, 0.5)
    G, Gs = getNodalSpreadGradients(mesh, avFunc)
    if mesh.dim == 2:
        b = float(beta)
        Gs = (1.0 - b) * Gs + b * G
        Lap = G.T @ Gs
        M = 0.5 * sp.kron(av3term(int(n[1] + 1), b), eye(n[0] + 1)) + 0.5 * sp.kron(eye(n[1] + 1), av3term(int(n[0] + 1), b))
        return Lap.tocsc(), M.tocsc()
    if np.isscalar(beta):
        if beta != 1:
            raise ValueError("getSpreadNodalLaplacianAndMass: in 3-D beta is a pair (Laplacian, mass)")
        beta = [1.0, 1.0]
    Gs = (1.0 - beta[0]) * Gs + beta[0] * G
    Lap = G.T @ Gs
    I1, I2, I3 = eye(n[0] + 1), eye(n[1] + 1), eye(n[2] + 1)
    M = (1.0 / 3.0) * (sp.kron(I3, sp.kron(av3term(int(n[1] + 1), beta[1]), I1)) +
                       sp.kron(I3, sp.kron(I2, av3term(int(n[0] + 1), beta[1]))) +
                       sp.kron(av3term(int(n[2] + 1), beta[1]), sp.kron(I2, I1)))
    return Lap.tocsc(), M.tocsc()


def GetHelmholtzOperatorHO(mesh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld, beta=1.0):
    """GetHelmholtz.jl:54-72: H = Lap + M * Diagonal(mass), mass as in GetHelmholtzOperator."""
    Lap, M = getSpreadNodalLaplacianAndMass(mesh, beta)
    mass = helmholtz_diagonal(mesh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld)
    return (Lap + M @ sp.diags(mass, 0, format="csc")).tocsc()


def getMaximalFrequency(m, mesh):
    """GetHelmholtz.jl:75-79 (m is slowness squared)."""
    return (0.1 * 2 * math.pi) / (np.max(mesh.h) * math.sqrt(np.max(m)))


def GetHelmholtzShiftOP(mNodal, omega, shift):
    """GetHelmholtz.jl:81-83"""
    m = np.asarray(mNodal, dtype=np.float64).ravel(order="F")
    return sp.diags(m * (1j * shift * omega**2), 0, format="csc")


def getShiftedHelmholtzParamGamma(gamma, omega, s):
    """Helmholtz.jl:32-34"""
    return np.asarray(gamma) + s * np.real(omega)


# --------------------------------------------------------------------------
# Matrix-free statement of the same operator.
# This is what the CUDA kernels implement; tests prove it equals the Kronecker
# assembly above.
# --------------------------------------------------------------------------


def helmholtz_apply_matfree(x, nodes, h, mNodal, gamma, omega, NeumannOnTop, Sommerfeld, orderNeumannBC=2, shift=0.0,
                            transpose=False):
    """y = (H + i*shift*Re(w)^2 diag(m)) x on an N x k block (column-major nodes).
    transpose=True applies the conjugate transpose (what doTranspose=1 solves with)."""
    nodes = [int(v) for v in nodes]
    dim = len(nodes)
    BC = getBC(orderNeumannBC)
    x = np.asarray(x)
    squeeze = x.ndim == 1
    X = x.reshape((-1, 1)) if squeeze else x
    k = X.shape[1]
    shp = tuple(nodes) + (k,)
    U = X.reshape(shp, order="F")
    mesh = RegularMesh(np.array([0.0, 1.0] * dim), np.array(nodes) - 1)
    mesh_h = np.asarray(h, dtype=np.float64)

    class _M:  # light mesh with explicit h
        pass

    mm = _M()
    mm.nodes = np.array(nodes)
    mm.h = mesh_h
    mm.dim = dim
    c = helmholtz_diagonal(mm, mNodal, omega, gamma, NeumannOnTop, Sommerfeld)
    c = c + 1j * shift * (np.real(omega) ** 2) * np.asarray(mNodal, dtype=np.float64).ravel(order="F")
    if transpose:
        c = np.conj(c)
    Y = (c.reshape(tuple(nodes), order="F")[..., None]) * U
    for d in range(dim):
        n = nodes[d]
        ih2 = 1.0 / mesh_h[d] ** 2
        Ud = np.moveaxis(U, d, 0)
        Yd = np.moveaxis(Y, d, 0)
        diag = np.full(n, 2.0)
        diag[0] = BC
        diag[-1] = BC
        Yd += (diag * ih2).reshape((n,) + (1,) * (Ud.ndim - 1)) * Ud
        if not transpose:
            # row p couples to p+1 with -1 (or -BC if p is the first row), to p-1 with -1 (or -BC if last row)
            up = np.full(n - 1, -1.0)
            up[0] = -BC  # row 0 -> col 1
            lo = np.full(n - 1, -1.0)
            lo[-1] = -BC  # row n-1 -> col n-2
        else:
            # transposed: row p, col p+1 takes the (p+1 -> p) coefficient
            up = np.full(n - 1, -1.0)
            up[-1] = -BC  # (row n-2, col n-1) = L[n-1, n-2]
            lo = np.full(n - 1, -1.0)
            lo[0] = -BC  # (row 1, col 0) = L[0, 1]
        sh = (n - 1,) + (1,) * (Ud.ndim - 1)
        Yd[:-1] += (up * ih2).reshape(sh) * Ud[1:]
        Yd[1:] += (lo * ih2).reshape(sh) * Ud[:-1]
    out = Y.reshape((-1, k), order="F")
    return out[:, 0] if squeeze else out


# --------------------------------------------------------------------------
# Point sources (src/getPointSource.jl:63-112)
# --------------------------------------------------------------------------


def loc2cs(n, sub):
    """1-based subscripts -> 1-based column-major linear index (getPointSource.jl:82-102)."""
    n = [int(v) for v in n]
    sub = [int(v) for v in sub]
    if len(sub) == 2:
        return sub[0] + (sub[1] - 1) * n[0]
    return sub[0] + (sub[1] - 1) * n[0] + (sub[2] - 1) * n[0] * n[1]


def getTopPointSrc(mesh):
    n = mesh.n
    if mesh.dim == 3:
        return [(int(n[0]) + 1) // 2, (int(n[1]) + 1) // 2, 1]
    return [(int(n[0]) + 1) // 2, 1]


def getAcousticPointSource(mesh, src=None, dtype=np.complex128):
    """getPointSource.jl:105-112: q[src] = 1/||h||^2."""
    if src is None:
        src = getTopPointSrc(mesh)
    nodes = mesh.nodes
    q = np.zeros(int(np.prod(nodes)), dtype=dtype)
    q[loc2cs(nodes, src) - 1] = 1.0 / (np.linalg.norm(mesh.h) ** 2)
    return q, src


# --------------------------------------------------------------------------
# Geometric multigrid (Multigrid.jl -- un-vendored; standard algorithm,
# SURVEY.md section 3.3 / appendix A.3)
# --------------------------------------------------------------------------


def get1DFWInterp(n_nodes):
    """Linear interpolation on an odd node count: coarse j <-> fine 2j (0-based)."""
    if n_nodes <= 2:
        return sp.identity(n_nodes, format="csc")
    if n_nodes % 2 != 1:
        raise ValueError("getFWInterp(): geometric mode expects an odd number of nodes")
    half = 0.5 * np.ones(n_nodes - 1)
    P = sp.diags([half, np.ones(n_nodes), half], [-1, 0, 1], format="csc")
    return P[:, 0::2].tocsc()


def getFWInterp(nodes):
    Ps = [get1DFWInterp(int(n)) for n in nodes]
    P = Ps[0]
    for Pd in Ps[1:]:
        P = sp.kron(Pd, P)
    nc = [p.shape[1] for p in Ps]
    return P.tocsr(), nc


@dataclass
class MGparam:
    """Mirror of the Multigrid.MGparam fields the reference touches
    (test/ShiftedLaplacianTest.jl:63-64, ShiftedLaplacianMultigridSolver.jl:29,68,74-75,97)."""

    levels: int = 2
    numCores: int = 1
    maxOuterIter: int = 30
    relativeTol: float = 1e-6
    relaxType: str = "Jac"
    relaxParam: float = 0.75
    relaxPre: object = 2
    relaxPost: object = 2
    cycleType: str = "V"
    coarseSolveType: str = "NoMUMPS"  # "NoMUMPS"/"Julia" -> LU ; "GMRES" -> inexact
    coarseIters: int = 10  # Jacobi-preconditioned GMRES steps when coarseSolveType == "GMRES"
    doTranspose: int = 0
    As: list = field(default_factory=list)
    Ps: list = field(default_factory=list)
    Rs: list = field(default_factory=list)
    dinv: list = field(default_factory=list)
    nodes: list = field(default_factory=list)
    LU: object = None
    ncycles: int = 0

    def nsweeps(self, which, level):
        v = self.relaxPre if which == "pre" else self.relaxPost
        return int(v(level)) if callable(v) else int(v)


def getMGparam(levels, numCores, maxIter, relativeTol, relaxType, relaxParam, relaxPre, relaxPost, cycleType,
               coarseSolveType="NoMUMPS", coarseIters=10):
    return MGparam(levels, numCores, maxIter, relativeTol, relaxType, relaxParam, relaxPre, relaxPost, cycleType,
                   coarseSolveType, coarseIters)


def hierarchyExists(MG):
    return len(MG.As) > 0


def clearMG(MG):
    MG.As, MG.Ps, MG.Rs, MG.dinv, MG.nodes, MG.LU = [], [], [], [], [], None


def MGsetup(A, nodes, MG: MGparam, dtype=np.complex128):
    """Galerkin hierarchy A_{l+1} = R A_l P, R = 2^-dim P^T (ShiftedLaplacianMultigridSolver.jl:64-65)."""
    clearMG(MG)
    A = sp.csr_matrix(A, dtype=dtype)
    nodes = [int(v) for v in nodes]
    dim = len(nodes)
    for l in range(MG.levels):
        MG.As.append(A)
        MG.nodes.append(list(nodes))
        MG.dinv.append((MG.relaxParam / A.diagonal()).astype(dtype))
        if l == MG.levels - 1:
            break
        for n in nodes:
            if n % 2 != 1 or n < 3:
                raise ValueError(f"cannot coarsen node counts {nodes} (level {l + 1})")
        P, nc = getFWInterp(nodes)
        P = P.astype(dtype)
        R = (P.T * (0.5**dim)).tocsr()
        MG.Ps.append(P)
        MG.Rs.append(R)
        A = (R @ (A @ P)).tocsr()
        A.sort_indices()
        nodes = nc
    if MG.coarseSolveType in ("NoMUMPS", "Julia", "LU"):
        MG.LU = spla.splu(sp.csc_matrix(MG.As[-1], dtype=np.complex128 if dtype == np.complex128 else np.complex64))
    return MG


def _jacobi(A, dinv, x, b, nsweeps, x_is_zero=False):
    for s in range(nsweeps):
        if x_is_zero and s == 0:
            x = dinv[:, None] * b
        else:
            x = x + dinv[:, None] * (b - A @ x)
    return x


def _gmres_fixed(A, b, x, dinv, nsteps, x_is_zero=True):
    """`nsteps` steps of right-preconditioned (Jacobi) GMRES on each column: the Jac-GMRES
    smoother and the inexact "GMRES" coarsest solve.  One cycle, restart = nsteps, MGS."""
    X = x.copy()
    for c in range(b.shape[1]):
        r = b[:, c] if x_is_zero else b[:, c] - A @ X[:, c]
        beta = np.linalg.norm(r)
        if beta == 0.0:
            continue
        n = len(r)
        V = np.zeros((n, nsteps + 1), dtype=b.dtype)
        Z = np.zeros((n, nsteps), dtype=b.dtype)
        Hm = np.zeros((nsteps + 1, nsteps), dtype=b.dtype)
        V[:, 0] = r / beta
        j_done = 0
        for j in range(nsteps):
            Z[:, j] = dinv * V[:, j]
            w = A @ Z[:, j]
            for i in range(j + 1):
                Hm[i, j] = np.vdot(V[:, i], w)
                w = w - Hm[i, j] * V[:, i]
            Hm[j + 1, j] = np.linalg.norm(w)
            j_done = j + 1
            if abs(Hm[j + 1, j]) < 1e-300:
                break
            V[:, j + 1] = w / Hm[j + 1, j]
        e1 = np.zeros(j_done + 1, dtype=b.dtype)
        e1[0] = beta
        y = np.linalg.lstsq(Hm[: j_done + 1, :j_done], e1, rcond=None)[0]
        X[:, c] = X[:, c] + Z[:, :j_done] @ y
    return X


def recursiveCycle(MG: MGparam, b, x, level, x_is_zero=True):
    """One multigrid cycle on level `level` (0-based), b,x: N_l x k.  SURVEY.md section 3.3."""
    A = MG.As[level]
    dinv = MG.dinv[level]
    npre = MG.nsweeps("pre", level + 1)
    npost = MG.nsweeps("post", level + 1)
    if MG.relaxType == "Jac":
        x = _jacobi(A, dinv, x, b, npre, x_is_zero)
    elif MG.relaxType == "Jac-GMRES":
        x = _gmres_fixed(A, b, x, dinv, npre, x_is_zero)
    else:
        raise ValueError(MG.relaxType)
    r = b - A @ x
    bc = MG.Rs[level] @ r
    xc = np.zeros_like(bc)
    if level + 1 == MG.levels - 1:
        xc = coarsestSolve(MG, bc)
    else:
        if MG.cycleType == "V":
            xc = recursiveCycle(MG, bc, xc, level + 1, True)
        elif MG.cycleType == "W":
            xc = recursiveCycle(MG, bc, xc, level + 1, True)
            xc = recursiveCycle(MG, bc, xc, level + 1, False)
        elif MG.cycleType == "K":
            # two steps of FGMRES on the coarse system preconditioned by the recursive cycle
            Ac = MG.As[level + 1]
            for c in range(bc.shape[1]):
                xc[:, c] = _kcycle_fgmres(MG, Ac, bc[:, c], level + 1)
        else:
            raise ValueError(MG.cycleType)
    x = x + MG.Ps[level] @ xc
    if MG.relaxType == "Jac":
        x = _jacobi(A, dinv, x, b, npost, False)
    else:
        x = _gmres_fixed(A, b, x, dinv, npost, False)
    return x


def _kcycle_fgmres(MG, Ac, b, level):
    n = len(b)
    beta = np.linalg.norm(b)
    if beta == 0:
        return np.zeros_like(b)
    V = [b / beta]
    Z = []
    Hm = np.zeros((3, 2), dtype=b.dtype)
    for j in range(2):
        z = recursiveCycle(MG, V[j][:, None], np.zeros((n, 1), dtype=b.dtype), level, True)[:, 0]
        Z.append(z)
        w = Ac @ z
        for i in range(j + 1):
            Hm[i, j] = np.vdot(V[i], w)
            w = w - Hm[i, j] * V[i]
        Hm[j + 1, j] = np.linalg.norm(w)
        V.append(w / Hm[j + 1, j])
    e1 = np.zeros(3, dtype=b.dtype)
    e1[0] = beta
    y = np.linalg.lstsq(Hm, e1, rcond=None)[0]
    return Z[0] * y[0] + Z[1] * y[1]


def coarsestSolve(MG, bc):
    if MG.LU is not None:
        out = np.empty_like(bc)
        for c in range(bc.shape[1]):
            out[:, c] = MG.LU.solve(np.ascontiguousarray(bc[:, c]).astype(np.complex128)).astype(bc.dtype)
        return out
    if MG.coarseSolveType == "GMRES":
        return _gmres_fixed(MG.As[-1], bc, np.zeros_like(bc), MG.dinv[-1], MG.coarseIters, True)
    raise ValueError(MG.coarseSolveType)


def MGcycle(MG, b):
    """Preconditioner application z = M(b): one cycle from a zero initial guess."""
    squeeze = b.ndim == 1
    B = b.reshape(-1, 1) if squeeze else b
    if MG.levels == 1:
        Z = coarsestSolve(MG, B)
    else:
        Z = recursiveCycle(MG, B, np.zeros_like(B), 0, True)
    MG.ncycles += 1
    return Z[:, 0] if squeeze else Z


# --------------------------------------------------------------------------
# Krylov methods (KrylovMethods.jl -- un-vendored; textbook algorithms)
# --------------------------------------------------------------------------


def fgmres(Afun, b, restrt, tol=1e-6, maxIter=30, M=None, x=None):
    """Right-preconditioned flexible GMRES(restrt), modified Gram-Schmidt, Givens residual
    estimate, stop on ||r||/||b|| <= tol.  maxIter counts restart cycles.
    Returns (x, flag, err, n_prec, resvec)."""
    n = len(b)
    bnrm2 = np.linalg.norm(b)
    if bnrm2 == 0.0:
        return np.zeros_like(b), 0, 0.0, 0, []
    if M is None:
        M = lambda v: v
    if x is None:
        x = np.zeros_like(b)
        r = b.copy()
    else:
        r = b - Afun(x)
    err = np.linalg.norm(r) / bnrm2
    resvec = []
    nprec = 0
    if err <= tol:
        return x, 0, err, 0, resvec
    restrt = min(restrt, n - 1)
    dt = b.dtype
    for _ in range(maxIter):
        V = np.zeros((n, restrt + 1), dtype=dt)
        Z = np.zeros((n, restrt), dtype=dt)
        Hm = np.zeros((restrt + 1, restrt), dtype=dt)
        cs = np.zeros(restrt, dtype=dt)
        sn = np.zeros(restrt, dtype=dt)
        s = np.zeros(restrt + 1, dtype=dt)
        beta = np.linalg.norm(r)
        V[:, 0] = r / beta
        s[0] = beta
        jdone = 0
        for i in range(restrt):
            z = M(V[:, i])
            nprec += 1
            Z[:, i] = z
            w = Afun(z)
            for k in range(i + 1):
                Hm[k, i] = np.vdot(V[:, k], w)
                w = w - Hm[k, i] * V[:, k]
            Hm[i + 1, i] = np.linalg.norm(w)
            if abs(Hm[i + 1, i]) > 0:
                V[:, i + 1] = w / Hm[i + 1, i]
            for k in range(i):
                t = cs[k] * Hm[k, i] + sn[k] * Hm[k + 1, i]
                Hm[k + 1, i] = -np.conj(sn[k]) * Hm[k, i] + cs[k] * Hm[k + 1, i]
                Hm[k, i] = t
            a, bb = Hm[i, i], Hm[i + 1, i]
            den = math.sqrt(abs(a) ** 2 + abs(bb) ** 2)
            # complex Givens: c real, s complex, [c s; -conj(s) c] [a; b] = [rho; 0]
            if abs(a) == 0:
                cs[i], sn[i] = 0.0, 1.0
            else:
                cs[i] = abs(a) / den
                sn[i] = (a / abs(a)) * np.conj(bb) / den
            Hm[i, i] = cs[i] * a + sn[i] * bb
            Hm[i + 1, i] = 0.0
            s[i + 1] = -np.conj(sn[i]) * s[i]
            s[i] = cs[i] * s[i]
            err = abs(s[i + 1]) / bnrm2
            resvec.append(err)
            jdone = i + 1
            if err <= tol:
                break
        y = np.linalg.solve(np.triu(Hm[:jdone, :jdone]), s[:jdone])
        x = x + Z[:, :jdone] @ y
        if err <= tol:
            return x, 0, err, nprec, resvec
        r = b - Afun(x)
        err = np.linalg.norm(r) / bnrm2
        if err <= tol:
            return x, 0, err, nprec, resvec
    return x, -1, err, nprec, resvec


def bicgstab(Afun, b, tol=1e-6, maxIter=30, M=None, x=None):
    """Preconditioned BiCGSTAB (two preconditioner applications per iteration).
    Returns (x, flag, err, iters, nprec, resvec)."""
    bnrm2 = np.linalg.norm(b)
    if bnrm2 == 0.0:
        return np.zeros_like(b), 0, 0.0, 0, 0, []
    if M is None:
        M = lambda v: v
    if x is None:
        x = np.zeros_like(b)
        r = b.copy()
    else:
        r = b - Afun(x)
    err = np.linalg.norm(r) / bnrm2
    resvec = []
    if err <= tol:
        return x, 0, err, 0, 0, resvec
    rt = r.copy()
    rho = alpha = omega = 1.0
    p = v = None
    nprec = 0
    for it in range(1, maxIter + 1):
        rho1 = np.vdot(rt, r)
        if rho1 == 0:
            return x, -2, err, it, nprec, resvec
        if it == 1:
            p = r.copy()
        else:
            beta = (rho1 / rho) * (alpha / omega)
            p = r + beta * (p - omega * v)
        ph = M(p)
        nprec += 1
        v = Afun(ph)
        alpha = rho1 / np.vdot(rt, v)
        s = r - alpha * v
        err = np.linalg.norm(s) / bnrm2
        if err <= tol:
            x = x + alpha * ph
            resvec.append(err)
            return x, 0, err, it, nprec, resvec
        sh = M(s)
        nprec += 1
        t = Afun(sh)
        omega = np.vdot(t, s) / np.vdot(t, t)
        x = x + alpha * ph + omega * sh
        r = s - omega * t
        err = np.linalg.norm(r) / bnrm2
        resvec.append(err)
        if err <= tol:
            return x, 0, err, it, nprec, resvec
        rho = rho1
    return x, -1, err, maxIter, nprec, resvec


# --------------------------------------------------------------------------
# The solver plugin (src/ShiftedLaplacianMultigridSolver.jl)
# --------------------------------------------------------------------------


@dataclass
class HelmholtzParam:
    """src/Helmholtz.jl:13-20"""

    Mesh: RegularMesh
    gamma: np.ndarray
    m: np.ndarray
    omega: complex
    NeumannOnTop: bool
    Sommerfeld: bool


@dataclass
class ShiftedLaplacianMultigridSolver:
    """src/ShiftedLaplacianMultigridSolver.jl:4-15"""

    helmParam: HelmholtzParam
    MG: MGparam
    shift: np.ndarray
    Krylov: str = "BiCGSTAB"
    inner: int = 5
    doClear: int = 0
    verbose: bool = False
    setupTime: float = 0.0
    nPrec: int = 0
    solveTime: float = 0.0
    iters: list = field(default_factory=list)
    resvecs: list = field(default_factory=list)


def getShiftedLaplacianMultigridSolver(helmParam, MG, shift, Krylov="BiCGSTAB", inner=5, verbose=False):
    """ShiftedLaplacianMultigridSolver.jl:24-30"""
    if np.isscalar(shift):
        shift = np.ones(MG.levels) * shift
    return ShiftedLaplacianMultigridSolver(helmParam, MG, np.asarray(shift, dtype=np.float64), Krylov, inner, 0, verbose)


def solveLinearSystem(ShiftedHT, B, param: ShiftedLaplacianMultigridSolver, doTranspose=0, dtype=np.complex128):
    """ShiftedLaplacianMultigridSolver.jl:33-102.  `ShiftedHT` is the conjugate transpose of the
    shifted matrix (as the reference's caller passes, test/ShiftedLaplacianTest.jl:83); it is used only
    to build the hierarchy on the first call.  Right-hand sides are solved column by column
    (batched, not block, Krylov -- see DESIGN.md)."""
    import time

    B = np.asarray(B)
    vec_in = B.ndim == 1
    Bm = (B.reshape(-1, 1) if vec_in else B).astype(dtype)
    if param.doClear == 1:
        clearMG(param.MG)
    if np.linalg.norm(Bm) == 0.0:
        return np.zeros_like(B), param
    t0 = time.perf_counter()
    hp = param.helmParam
    MG = param.MG
    if not hierarchyExists(MG):
        SH = sp.csr_matrix(ShiftedHT).conj().T.tocsr()  # undo the adjoint the caller applied
        if doTranspose == 1:
            SH = SH.conj().T.tocsr()
        MGsetup(SH, hp.Mesh.nodes, MG, dtype)
        MG.doTranspose = doTranspose
    elif doTranspose != MG.doTranspose:
        SH = MG.As[0].conj().T.tocsr()
        MGsetup(SH, hp.Mesh.nodes, MG, dtype)
        MG.doTranspose = doTranspose
    SH = MG.As[0]
    mvec = np.asarray(hp.m, dtype=np.float64).ravel(order="F")
    sgn = -1.0 if doTranspose == 1 else 1.0
    shiftdiag = (sgn * 1j * param.shift[0] * (np.real(hp.omega) ** 2) * mvec).astype(dtype)

    def Afun(x):  # GetHelmholtz.jl:85-95: H x = SH x - i s w^2 m x
        return SH @ x - shiftdiag * x

    param.setupTime += time.perf_counter() - t0
    t0 = time.perf_counter()
    X = np.zeros_like(Bm)
    param.iters, param.resvecs = [], []
    worst = 0
    for c in range(Bm.shape[1]):
        b = Bm[:, c]
        if param.Krylov == "GMRES":
            x, flag, err, nprec, resvec = fgmres(Afun, b, param.inner, MG.relativeTol, MG.maxOuterIter,
                                                  lambda v: MGcycle(MG, v))
            param.nPrec += nprec
            it = int(math.ceil(nprec / max(param.inner, 1)))
        elif param.Krylov == "BiCGSTAB":
            x, flag, err, it, nprec, resvec = bicgstab(Afun, b, MG.relativeTol, MG.maxOuterIter,
                                                        lambda v: MGcycle(MG, v))
            param.nPrec += nprec
        else:
            raise ValueError(param.Krylov)
        X[:, c] = x
        param.iters.append(nprec)
        param.resvecs.append(resvec)
        worst = max(worst, it if flag == 0 else MG.maxOuterIter)
    param.solveTime += time.perf_counter() - t0
    if worst >= MG.maxOuterIter and param.verbose:
        print("WARNING: MG solver reached maximum iterations without convergence")
    return (X[:, 0] if vec_in else X), param


def copySolver(s: ShiftedLaplacianMultigridSolver):
    """ShiftedLaplacianMultigridSolver.jl:18-22 -- clone without hierarchy."""
    MG = s.MG
    MG2 = MGparam(MG.levels, MG.numCores, MG.maxOuterIter, MG.relativeTol, MG.relaxType, MG.relaxParam, MG.relaxPre,
                  MG.relaxPost, MG.cycleType, MG.coarseSolveType, MG.coarseIters)
    return getShiftedLaplacianMultigridSolver(s.helmParam, MG2, s.shift, s.Krylov, s.inner, s.verbose)


def clear(s: ShiftedLaplacianMultigridSolver):
    """ShiftedLaplacianMultigridSolver.jl:105-109"""
    clearMG(s.MG)
    s.doClear = 0


# --------------------------------------------------------------------------
# Galerkin coarse operator as stencil-coefficient arrays (the layout the CUDA
# library stores): used by the tests to compare hh_get_level_stencil output.
# --------------------------------------------------------------------------


def csr_to_stencil(A, nodes):
    """Return coef[s, node] (s = 0..3^dim-1, offset index (d1+1) + 3*(d2+1) (+ 9*(d3+1))) for a
    matrix with at most a 3^dim-point stencil on a column-major node grid."""
    nodes = [int(v) for v in nodes]
    dim = len(nodes)
    N = int(np.prod(nodes))
    A = sp.coo_matrix(A)
    ns = 3**dim
    coef = np.zeros((ns, N), dtype=A.dtype)
    strides = [1, nodes[0], nodes[0] * nodes[1]][:dim]
    row = A.row.astype(np.int64)
    col = A.col.astype(np.int64)
    s = np.zeros(len(row), dtype=np.int64)
    r = row.copy()
    c = col.copy()
    for d in reversed(range(dim)):
        rd = r // strides[d]
        cd = c // strides[d]
        r = r - rd * strides[d]
        c = c - cd * strides[d]
        dd = cd - rd
        assert np.all(np.abs(dd) <= 1)
        s += (dd + 1) * (3**d)
    np.add.at(coef, (s, row), A.data)
    return coef
