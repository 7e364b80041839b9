"""Hierarchy and cycle parity against the oracle's CSR Galerkin multigrid."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _setup(pkg, ho, nodes, levels, cycle, coarse, prec=np.complex128, relax="Jac", shift=0.2, pre=2, post=2, seed=3,
           coarse_iters=10, neu=True):
    rng = np.random.default_rng(seed)
    dim = len(nodes)
    domain = sum([[0.0, 0.1 * (n - 1)] for n in nodes], [])
    mesh = ho.getRegularMesh(domain, np.array(nodes) - 1)
    v = rng.uniform(1.5, 3.0, size=nodes)
    m = 1.0 / v**2
    w = ho.getMaximalFrequency(m, mesh)
    pad = [max(2, n // 8) for n in nodes]
    H, gamma = ho.GetHelmholtzOperatorABL(mesh, m, w, 0.01 * w * np.ones(nodes), neu, pad, w, True)
    SH = H + ho.GetHelmholtzShiftOP(m, w, shift)
    MGo = ho.getMGparam(levels, 1, 30, 1e-6, relax, 0.8, pre, post, cycle, coarse, coarse_iters)
    ho.MGsetup(SH, nodes, MGo)
    pmesh = pkg.getRegularMesh(domain, np.array(nodes) - 1)
    MG = pkg.getMGparam(prec, pkg.Int64, levels, 1, 30, 1e-6, relax, 0.8, pre, post, cycle, coarse,
                        coarseIters=coarse_iters)
    hp = pkg.HelmholtzParam(pmesh, gamma, m.ravel(order="F"), w, neu, True)
    Ainv = pkg.getShiftedLaplacianMultigridSolver(hp, MG, shift, "GMRES", 5)
    hd = pkg.api._ensure_hierarchy(Ainv, 0)
    return mesh, H, SH, MGo, Ainv, hd, rng


@pytest.mark.parametrize("nodes", [[33, 17], [17, 9, 13]])
def test_galerkin_stencils_match_RAP(gpu_pkg, ho, nodes):
    pkg = gpu_pkg
    mesh, H, SH, MGo, Ainv, hd, rng = _setup(pkg, ho, nodes, 3, "V", "NoMUMPS")
    for lvl in (1, 2):
        nl = np.zeros(len(nodes), dtype=np.int64)
        pkg._lib.check(hd.lib.hh_level_nodes(hd.h, lvl, nl.ctypes.data_as(C.POINTER(C.c_int64))), hd.h)
        assert list(nl) == MGo.nodes[lvl]
        Nl = int(np.prod(nl))
        ns = 3 ** len(nodes)
        coef = np.empty((ns, Nl), dtype=np.complex128)
        pkg._lib.check(hd.lib.hh_get_level_stencil(hd.h, lvl, coef.ctypes.data), hd.h)
        ref = ho.csr_to_stencil(MGo.As[lvl], MGo.nodes[lvl])
        assert rel_err(coef, ref) < 1e-13


@pytest.mark.parametrize("nodes,levels,cycle,coarse", [
    ([65, 33], 2, "W", "NoMUMPS"),
    ([65, 33], 3, "V", "Julia"),
    ([65, 33], 3, "W", "NoMUMPS"),
    ([33, 33], 3, "V", "GMRES"),
    ([17, 17, 17], 2, "V", "NoMUMPS"),
    ([17, 25, 17], 3, "V", "NoMUMPS"),
    ([17, 17, 33], 3, "W", "GMRES"),
])
def test_cycle_matches_oracle(gpu_pkg, ho, nodes, levels, cycle, coarse):
    pkg = gpu_pkg
    mesh, H, SH, MGo, Ainv, hd, rng = _setup(pkg, ho, nodes, levels, cycle, coarse)
    N = int(np.prod(nodes))
    for nrhs in (1, 3):
        B = np.asfortranarray(rng.standard_normal((N, nrhs)) + 1j * rng.standard_normal((N, nrhs)))
        Z = np.empty_like(B, order="F")
        pkg._lib.check(hd.lib.hh_cycle(hd.h, B.ctypes.data, Z.ctypes.data, nrhs), hd.h)
        Zo = ho.MGcycle(MGo, B)
        # classical vs modified Gram-Schmidt in the inexact coarsest solve perturbs at round-off level only
        assert rel_err(Z, Zo) < 1e-10


def test_cycle_complex64(gpu_pkg, ho):
    pkg = gpu_pkg
    mesh, H, SH, MGo, Ainv, hd, rng = _setup(pkg, ho, [17, 17, 17], 3, "V", "NoMUMPS", prec=np.complex64)
    N = 17**3
    B = np.asfortranarray((rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))).astype(np.complex64))
    Z = np.empty_like(B, order="F")
    pkg._lib.check(hd.lib.hh_cycle(hd.h, B.ctypes.data, Z.ctypes.data, 2), hd.h)
    assert rel_err(Z, ho.MGcycle(MGo, B.astype(np.complex128))) < 2e-4


@pytest.mark.parametrize("nodes,levels,cycle,relax,coarse", [
    ([65, 33], 3, "K", "Jac", "NoMUMPS"),
    ([65, 33], 3, "V", "Jac-GMRES", "NoMUMPS"),
    ([33, 33, 17], 3, "K", "Jac-GMRES", "GMRES"),
    ([65, 65], 4, "K", "Jac-GMRES", "GMRES"),
])
def test_kcycle_and_jac_gmres_match_oracle(gpu_pkg, ho, nodes, levels, cycle, relax, coarse):
    """The reference's production setting (examples/PointSourceADR/runExperiments.jl:114-122):
    K-cycle with the Jac-GMRES smoother and an inexact GMRES coarsest solve."""
    pkg = gpu_pkg
    mesh, H, SH, MGo, Ainv, hd, rng = _setup(pkg, ho, nodes, levels, cycle, coarse, relax=relax,
                                             pre=lambda l: l + 1, post=lambda l: l + 1)
    N = int(np.prod(nodes))
    B = np.asfortranarray(rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2)))
    Z = np.empty_like(B, order="F")
    pkg._lib.check(hd.lib.hh_cycle(hd.h, B.ctypes.data, Z.ctypes.data, 2), hd.h)
    Zo = ho.MGcycle(MGo, B)
    assert rel_err(Z, Zo) < 1e-8


@pytest.mark.parametrize("nodes,cycle,prec,tol", [
    ([41, 25, 49], "W", np.complex128, 1e-10),   # partial tiles in x and y, 2 z-chunks
    ([33, 33, 65], "V", np.complex128, 1e-10),
    ([41, 25, 49], "W", np.complex64, 3e-4),     # padded (pitched) ComplexF32 layout on the TMA kernels
])
def test_fused_cycle_kernels_match_oracle(gpu_pkg, ho, nodes, cycle, prec, tol):
    """Pre-smoothing count 1 / post 2 exercises every fused fine-level kernel: first sweep + residual + restriction
    in one pass (partial sums combined in a fixed order) and prolongation + first post-smoothing sweep."""
    pkg = gpu_pkg
    mesh, H, SH, MGo, Ainv, hd, rng = _setup(pkg, ho, nodes, 3, cycle, "GMRES", prec=prec, pre=1, post=2)
    N = int(np.prod(nodes))
    for nrhs in (1, 3, 4):
        B = np.asfortranarray((rng.standard_normal((N, nrhs)) + 1j * rng.standard_normal((N, nrhs))).astype(prec))
        Z = np.empty_like(B, order="F")
        pkg._lib.check(hd.lib.hh_cycle(hd.h, B.ctypes.data, Z.ctypes.data, nrhs), hd.h)
        Zo = ho.MGcycle(MGo, B.astype(np.complex128))
        assert rel_err(Z, Zo) < tol
        # run-to-run determinism of the fixed-order partial-sum combination
        Z2 = np.empty_like(B, order="F")
        pkg._lib.check(hd.lib.hh_cycle(hd.h, B.ctypes.data, Z2.ctypes.data, nrhs), hd.h)
        assert np.array_equal(Z, Z2)
