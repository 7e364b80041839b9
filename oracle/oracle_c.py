"""ctypes wrapper of oracle/helm_oracle_c.c (C/OpenMP port of the oracle; TEST INFRASTRUCTURE / timed CPU arm
only -- see the header of helm_oracle_c.c).  Nothing in the product imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libhelm_oracle_c.so")
_lib = None


def load(build_if_missing=True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        if not build_if_missing:
            raise ImportError(SO + " is not built (make -C oracle)")
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(SO)
    i64p, dp = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    lib.horc_create.restype = C.c_void_p
    lib.horc_create.argtypes = [C.c_int, i64p, dp, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double,
                                C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.horc_destroy.argtypes = [C.c_void_p]
    lib.horc_setup_seconds.restype = C.c_double
    lib.horc_setup_seconds.argtypes = [C.c_void_p]
    lib.horc_level_nnz.restype = C.c_int64
    lib.horc_level_nnz.argtypes = [C.c_void_p, C.c_int]
    lib.horc_level_size.restype = C.c_int64
    lib.horc_level_size.argtypes = [C.c_void_p, C.c_int]
    lib.horc_cycle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.horc_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.horc_solve_fgmres.restype = C.c_double
    lib.horc_solve_fgmres.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                      C.POINTER(C.c_int32), dp]
    lib.horc_num_threads.restype = C.c_int
    lib.horc_set_threads.argtypes = [C.c_int]
    _lib = lib
    return lib


class OracleC:
    """Assembled-matrix Galerkin MG + FGMRES on the CPU (V or W cycle, Jacobi, inexact GMRES coarsest solve)."""

    def __init__(self, n_nodes, h, m, gamma, omega, neumann_top, sommerfeld, shift, levels, relax_param=0.8, npre=2,
                 npost=2, cycle="V", coarse_iters=10, order_bc=2):
        lib = load()
        self.lib = lib
        # all host cores unless HH_CPU_THREADS says otherwise (torchrun forces OMP_NUM_THREADS=1 on every rank)
        lib.horc_set_threads(int(os.environ.get("HH_CPU_THREADS", os.cpu_count() or 1)))
        nn = np.ascontiguousarray(np.asarray(n_nodes, dtype=np.int64))
        hh = np.ascontiguousarray(np.asarray(h, dtype=np.float64))
        mm = np.ascontiguousarray(np.asarray(m, dtype=np.float64).ravel(order="F"))
        gg = np.ascontiguousarray(np.asarray(gamma, dtype=np.float64).ravel(order="F"))
        w = complex(omega)
        self.N = int(np.prod(nn))
        self.levels = levels
        self.h = lib.horc_create(len(nn), nn.ctypes.data_as(C.POINTER(C.c_int64)), hh.ctypes.data_as(C.POINTER(C.c_double)),
                                 mm.ctypes.data_as(C.POINTER(C.c_double)), gg.ctypes.data_as(C.POINTER(C.c_double)), w.real,
                                 w.imag, int(neumann_top), int(sommerfeld), order_bc, float(shift), levels, relax_param, npre,
                                 npost, {"V": 0, "W": 1}[cycle], coarse_iters)
        if not self.h:
            raise RuntimeError("horc_create failed")

    @property
    def setup_seconds(self):
        return self.lib.horc_setup_seconds(self.h)

    @property
    def threads(self):
        return self.lib.horc_num_threads()

    def apply(self, X, shifted=False):
        X = np.asfortranarray(np.asarray(X, dtype=np.complex128).reshape(self.N, -1))
        Y = np.empty_like(X, order="F")
        self.lib.horc_apply(self.h, X.ctypes.data, Y.ctypes.data, X.shape[1], int(shifted))
        return Y

    def cycle(self, B):
        B = np.asfortranarray(np.asarray(B, dtype=np.complex128).reshape(self.N, -1))
        Z = np.empty_like(B, order="F")
        self.lib.horc_cycle(self.h, B.ctypes.data, Z.ctypes.data, B.shape[1])
        return Z

    def solve(self, B, inner=5, max_cycles=30, tol=1e-6, max_prec=0):
        B = np.asfortranarray(np.asarray(B, dtype=np.complex128).reshape(self.N, -1))
        X = np.empty_like(B, order="F")
        k = B.shape[1]
        iters = np.zeros(k, dtype=np.int32)
        relres = np.zeros(k, dtype=np.float64)
        secs = self.lib.horc_solve_fgmres(self.h, B.ctypes.data, X.ctypes.data, k, inner, max_cycles, tol, max_prec,
                                          iters.ctypes.data_as(C.POINTER(C.c_int32)),
                                          relres.ctypes.data_as(C.POINTER(C.c_double)))
        return X, iters, relres, secs

    def close(self):
        if self.h:
            self.lib.horc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
