"""Generates the committed fixtures under tests/golden/ (run in the build container, where /root/reference
exists; the GPU box only reads the .npz files).

  seg_salt_vp.npz      the SEG/EAGE salt P-velocity model the reference ships as data
                       (examples/SEGmodel2Dsalt.dat, 128 rows x 256 columns, m/s) -- BASELINE config 2
  oracle_goldens.npz   seeded inputs / outputs of the oracle (oracle/helm_oracle.py) for small cases of every
                       stage of the path: absorbing layer, operator apply, Galerkin stencils, one cycle, solves.
                       The reference holds no golden vectors for this path, so these pin
                       the oracle itself; the oracle is pinned to the reference by the known-answer tests in
                       tests/test_oracle_operator.py.
"""
import os
import sys

import numpy as np
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

ho = graft.load_oracle()


def seg():
    src = "/root/reference/examples/SEGmodel2Dsalt.dat"
    vp = np.loadtxt(src)
    assert vp.shape == (128, 256)
    np.savez_compressed(os.path.join(HERE, "seg_salt_vp.npz"), vp_ms=vp.astype(np.int16))


def problem(nodes, seed, neu=True, somm=True):
    rng = np.random.default_rng(seed)
    domain = sum([[0.0, 0.1 * (n - 1)] for n in nodes], [])
    mesh = ho.getRegularMesh(domain, np.array(nodes) - 1)
    v = rng.uniform(1.5, 3.0, size=nodes)
    m = 1.0 / v**2
    w = ho.getMaximalFrequency(m, mesh)
    pad = [max(2, n // 8) for n in nodes]
    H, gamma = ho.GetHelmholtzOperatorABL(mesh, m, w, 0.01 * w * np.ones(nodes), neu, pad, w, somm)
    return mesh, m, gamma, w, H, rng


def goldens():
    out = {}
    out["abl2d_neu"] = ho.getABL([19, 13], True, [4, 3], 2.5)
    out["abl2d_noneu"] = ho.getABL([19, 13], False, [4, 3], 2.5)
    out["abl3d_neu"] = ho.getABL([11, 9, 10], True, [3, 2, 4], 1.7)
    out["abl3d_noneu"] = ho.getABL([11, 9, 10], False, [3, 2, 4], 1.7)
    for name, nodes in (("2d", [33, 17]), ("3d", [17, 9, 13])):
        mesh, m, gamma, w, H, rng = problem(nodes, 42)
        N = int(np.prod(nodes))
        x = rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))
        SH = H + ho.GetHelmholtzShiftOP(m, w, 0.2)
        out[f"{name}_nodes"] = np.array(nodes)
        out[f"{name}_m"] = m
        out[f"{name}_gamma"] = gamma
        out[f"{name}_w"] = np.array(w)
        out[f"{name}_x"] = x
        out[f"{name}_Hx"] = H @ x
        out[f"{name}_SHx"] = SH @ x
        out[f"{name}_SHtx"] = SH.conj().T @ x
        MG = ho.getMGparam(3, 1, 30, 1e-8, "Jac", 0.8, 2, 2, "V", "NoMUMPS")
        ho.MGsetup(SH, nodes, MG)
        out[f"{name}_stencil_l1"] = ho.csr_to_stencil(MG.As[1], MG.nodes[1])
        out[f"{name}_stencil_l2"] = ho.csr_to_stencil(MG.As[2], MG.nodes[2])
        out[f"{name}_cycle"] = ho.MGcycle(MG, x)
        q, _ = ho.getAcousticPointSource(mesh)
        hp = ho.HelmholtzParam(mesh, gamma, m.ravel(order="F"), w, True, True)
        A = ho.getShiftedLaplacianMultigridSolver(hp, MG, 0.2, "GMRES", 5)
        xs, A = ho.solveLinearSystem(SH.conj().T, q, A)
        out[f"{name}_q"] = q
        out[f"{name}_solve_fgmres5"] = xs
        out[f"{name}_solve_iters"] = np.array(A.iters)
        out[f"{name}_direct"] = spla.splu(H.tocsc()).solve(q)
    np.savez_compressed(os.path.join(HERE, "oracle_goldens.npz"), **out)


if __name__ == "__main__":
    seg()
    goldens()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
