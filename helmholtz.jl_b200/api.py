"""Host-side mirror of the reference's Julia interface for the solve path, on top of the C ABI.

Same names, argument order and error behaviour as JuliaInv/Helmholtz.jl (no Julia runtime exists in
this image, so the mirror is Python; the Julia shim with identical structure is in
`julia/HelmholtzB200.jl`):

  HelmholtzParam                      src/Helmholtz.jl:13-20
  getShiftedHelmholtzParam            src/Helmholtz.jl:32-34
  GetHelmholtzOperator (3 methods)    src/GetHelmholtz.jl:14-16, 22-31, 33-50
  GetHelmholtzShiftOP                 src/GetHelmholtz.jl:81-83
  getABL / getMaximalFrequency        src/GetHelmholtz.jl:97-220 / 75-79
  getAcousticPointSource, loc2cs, ... src/getPointSource.jl:63-112
  getMGparam (Multigrid.jl)           call sites test/ShiftedLaplacianTest.jl:63-64,127-128,137
  getShiftedLaplacianMultigridSolver  src/ShiftedLaplacianMultigridSolver.jl:24-30
  solveLinearSystem / copySolver / clear!   src/ShiftedLaplacianMultigridSolver.jl:33-102, 18-22, 105-109

All arithmetic happens in libhelmholtz_b200.so on the GPU; nothing here computes a solve on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import time

import numpy as np

from . import _lib as L

ComplexF64 = np.complex128
ComplexF32 = np.complex64
Int64 = np.int64


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


# ------------------------------------------------------------------ jInv.Mesh.RegularMesh (boundary type)
class RegularMesh:
    """Only what crosses the ABI: domain, n (cells), h, dim (jInv.Mesh.RegularMesh)."""

    def __init__(self, domain, n):
        self.domain = np.asarray(domain, dtype=np.float64).ravel()
        self.n = np.asarray(n, dtype=np.int64).ravel()
        self.dim = len(self.n)
        self.h = (self.domain[1::2] - self.domain[0::2]) / self.n.astype(np.float64)


def getRegularMesh(domain, n):
    return RegularMesh(domain, n)


# ------------------------------------------------------------------ HelmholtzParam
class HelmholtzParam:
    """src/Helmholtz.jl:13-20.  gamma already contains the absorbing layer."""

    def __init__(self, Mesh, gamma, m, omega, NeumannOnTop, Sommerfeld):
        self.Mesh = Mesh
        self.gamma = np.asarray(gamma, dtype=np.float64)
        self.m = np.asarray(m, dtype=np.float64)
        self.omega = complex(omega) if np.iscomplexobj(omega) else float(omega)
        self.NeumannOnTop = bool(NeumannOnTop)
        self.Sommerfeld = bool(Sommerfeld)


def getShiftedHelmholtzParam(p, s):
    """src/Helmholtz.jl:32-34"""
    return HelmholtzParam(p.Mesh, p.gamma + s * np.real(p.omega), p.m, p.omega, p.NeumannOnTop, p.Sommerfeld)


# ------------------------------------------------------------------ device handle
class _Handle:
    """Owns an hh_handle_t (HelmholtzParam + device state)."""

    def __init__(self, Mesh, m, omega, gamma, NeumannOnTop, Sommerfeld, orderNeumannBC=2, precision=ComplexF64,
                 devices=None, cycle_precision=None, slabs=None, levels=None):
        lib = L.load()
        nodes = (np.asarray(Mesh.n, dtype=np.int64) + 1).copy()
        N = int(np.prod(nodes))
        mm = np.ascontiguousarray(np.asarray(m, dtype=np.float64).ravel(order="F"))
        gg = np.ascontiguousarray(np.asarray(gamma, dtype=np.float64).ravel(order="F"))
        # NCCL slabs: m / gamma may hold only the planes slabs["model_planes"] = (k0, k1) of the last dimension
        mp0, mp1 = (slabs or {}).get("model_planes", (0, int(nodes[-1])))
        Nm = int(np.prod(nodes[:-1])) * (mp1 - mp0)
        if mm.size != Nm or gg.size != Nm:
            raise ValueError(f"m and gamma must have {Nm} entries (got {mm.size}, {gg.size})")
        h = np.ascontiguousarray(np.asarray(Mesh.h, dtype=np.float64))
        self.dim = int(Mesh.dim)
        self.nodes = nodes
        self.N = N
        self.dtype = np.dtype(precision)
        prec = L.HH_C64 if self.dtype == np.complex128 else L.HH_C32
        if self.dtype not in (np.dtype(np.complex128), np.dtype(np.complex64)):
            raise ValueError("precision must be ComplexF64 or ComplexF32")
        if cycle_precision is not None and np.dtype(cycle_precision) != self.dtype:
            # opt-in extension: ComplexF64 Krylov with the multigrid cycle evaluated in ComplexF32
            if not (self.dtype == np.complex128 and np.dtype(cycle_precision) == np.complex64):
                raise ValueError("cyclePrecision must be ComplexF32 on a ComplexF64 solver")
            prec = L.HH_C64_MIXED
        w = complex(omega)
        out = C.c_void_p()
        if devices is None:
            devices = [_current_device()]
        devs = np.asarray(devices, dtype=np.int32)
        common = (self.dim, _ptr(nodes, C.c_int64), _ptr(h, C.c_double), _ptr(mm, C.c_double), _ptr(gg, C.c_double),
                  w.real, w.imag, int(bool(NeumannOnTop)), int(bool(Sommerfeld)), int(orderNeumannBC), prec)
        self.slabs = slabs
        self.order = int(orderNeumannBC)
        self.planes = (0, int(nodes[-1]))  # planes of the last dimension the caller's B / X hold
        if slabs is None:
            rc = lib.hh_create_multi(*common, _ptr(devs, C.c_int), len(devs), C.byref(out))
            L.check(rc, None)
        else:
            # one grid split into slabs along the last dimension (include/helmholtz_b200.h, hh_create_slab_*)
            if levels is None:
                raise ValueError("a slab handle needs the number of multigrid levels")
            if slabs["mode"] == "local":
                sd = np.asarray(slabs.get("devices", devs), dtype=np.int32)
                rc = lib.hh_create_slab_local(*common, _ptr(sd, C.c_int), len(sd), int(levels), C.byref(out))
                L.check(rc, None)
                devs = sd
            elif slabs["mode"] == "nccl":
                uid = (C.c_char * 128).from_buffer_copy(bytes(slabs["unique_id"]))
                rc = lib.hh_create_slab_nccl(*common, int(devs[0]), int(levels), int(slabs["rank"]), int(slabs["nranks"]),
                                             C.cast(uid, C.c_void_p), int(mp0), int(mp1 - mp0), C.byref(out))
                L.check(rc, None)
                o0, o1 = C.c_int64(), C.c_int64()
                lib.hh_slab_info(out, None, None, None, C.byref(o0), C.byref(o1))
                self.planes = (int(o0.value), int(o1.value))
                self.N = int(np.prod(nodes[:-1])) * (self.planes[1] - self.planes[0])  # owned nodes: size of B / X here
            else:
                raise ValueError("slabs['mode'] must be 'local' or 'nccl'")
        self.h = out
        self.devices = [int(d) for d in devs]
        self.lib = lib

    def close(self):
        if getattr(self, "h", None):
            self.lib.hh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _current_device():
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def _as_block(x, N, dtype):
    """View x as an N x k column-major block of `dtype` without copying when it already is one
    (numpy's default reshape of a Fortran-ordered array would silently copy)."""
    a = np.asarray(x)
    if a.ndim == 1:
        a = a.reshape(N, 1) if a.flags.c_contiguous else np.ascontiguousarray(a).reshape(N, 1)
    elif a.ndim != 2 or a.shape[0] != N:
        a = a.reshape((N, -1), order="F")
    if a.dtype != dtype or not a.flags.f_contiguous:
        a = np.asfortranarray(a, dtype=dtype)
    return a


def _torch_rows(x, hd, what):
    """CUDA tensor -> (rows, layout): `rows` is the (nrhs, N) contiguous block the library works on (every right-hand
    side contiguous = Julia's column-major N x nrhs).  Accepted shapes, as the numpy path and the reference:
    (N,), (N, nrhs) [layout "cols": copied transposed], and the zero-copy (nrhs, N) with rows = right-hand sides
    [layout "rows"].  An N x N block is read as (N, nrhs) like the reference's."""
    import torch

    if x.device.index != hd.devices[0]:
        raise ValueError(f"{what} lives on cuda:{x.device.index} but the solver handle was created on cuda:{hd.devices[0]} "
                         "(set solver.devices = [B.device.index] before the first solve)")
    want = torch.complex128 if hd.dtype == np.complex128 else torch.complex64
    N = hd.N
    if x.dim() == 1 and x.shape[0] == N:
        return x.to(want).reshape(1, N).contiguous(), "vec"
    if x.dim() == 2 and x.shape[0] == N:
        return x.to(want).t().contiguous(), "cols"
    if x.dim() == 2 and x.shape[1] == N:
        return x.to(want).contiguous(), "rows"
    raise ValueError(f"{what} has shape {tuple(x.shape)}; expected ({N},), ({N}, nrhs) or the zero-copy (nrhs, {N})")


def _is_torch_cuda(x):
    try:
        import torch

        return isinstance(x, torch.Tensor) and x.is_cuda
    except Exception:
        return False


# ------------------------------------------------------------------ operator objects
class HelmholtzShiftOP:
    """What GetHelmholtzShiftOP returns: i*shift*omega^2*diag(m) (src/GetHelmholtz.jl:81-83), kept symbolic."""

    def __init__(self, shift, omega):
        self.shift = float(shift)
        self.omega = float(omega)

    def __neg__(self):
        return HelmholtzShiftOP(-self.shift, self.omega)

    def __mul__(self, a):
        if np.isscalar(a) and np.isreal(a):
            return HelmholtzShiftOP(float(a) * self.shift, self.omega)
        return NotImplemented

    __rmul__ = __mul__


def GetHelmholtzShiftOP(mNodal, omega, shift):
    if np.iscomplexobj(omega):
        # the reference's method is typed omega::Float64 (GetHelmholtz.jl:81)
        raise TypeError("GetHelmholtzShiftOP: omega must be real (Float64)")
    return HelmholtzShiftOP(shift, omega)


class HelmholtzOperator:
    """Matrix-free counterpart of the sparse matrix GetHelmholtzOperator returns: `H @ x`, `H * x`,
    `H + GetHelmholtzShiftOP(...)`, `H.H` / `H.T.conj()` (adjoint view)."""

    def __init__(self, handle, shift=0.0, adjoint=False):
        self._hd = handle
        self.shift = float(shift)
        self.adjoint = bool(adjoint)
        self.shape = (handle.N, handle.N)
        self.dtype = handle.dtype

    def __add__(self, other):
        if isinstance(other, HelmholtzShiftOP):
            return HelmholtzOperator(self._hd, self.shift + other.shift, self.adjoint)
        return NotImplemented

    __radd__ = __add__

    @property
    def H(self):
        return HelmholtzOperator(self._hd, self.shift, not self.adjoint)

    def adjoint_view(self):
        return self.H

    def matvec(self, x):
        hd = self._hd
        if _is_torch_cuda(x):
            import torch

            xx, layout = _torch_rows(x, hd, "x")
            y = torch.empty_like(xx)
            L.check(hd.lib.hh_apply_device(hd.h, xx.data_ptr(), y.data_ptr(), xx.shape[0], int(self.shift != 0.0),
                                           self.shift, int(self.adjoint)), hd.h)
            return y.t() if layout == "cols" else y.reshape(x.shape)
        vec = np.ndim(x) == 1
        X = _as_block(x, hd.N, hd.dtype)
        Y = np.empty_like(X, order="F")
        L.check(hd.lib.hh_apply(hd.h, X.ctypes.data, Y.ctypes.data, X.shape[1], int(self.shift != 0.0), self.shift,
                                int(self.adjoint)), hd.h)
        return Y[:, 0] if vec else Y

    __matmul__ = matvec
    __mul__ = matvec

    def diagonal_mass(self):
        """The complex diagonal added to the Laplacian (mass + Sommerfeld [+ shift]), ComplexF64."""
        hd = self._hd
        out = np.empty(hd.N, dtype=np.complex128)
        L.check(hd.lib.hh_get_diagonal(hd.h, int(self.shift != 0.0), self.shift, _ptr(out.view(np.float64), C.c_double)),
                hd.h)
        return out


def getABL(n, NeumannAtFirstDim, ABLpad, ABLamp):
    """src/GetHelmholtz.jl:97-220 (host-side Float64 set-up, computed by the library)."""
    lib = L.load()
    n = np.ascontiguousarray(np.asarray(n, dtype=np.int64).ravel())
    pad = np.ascontiguousarray(np.asarray(ABLpad, dtype=np.int64).ravel())
    if pad.size == 1:
        pad = np.repeat(pad, n.size)
    out = np.empty(int(np.prod(n)), dtype=np.float64)
    L.check(lib.hh_get_abl(len(n), _ptr(n, C.c_int64), int(bool(NeumannAtFirstDim)), _ptr(pad, C.c_int64), float(ABLamp),
                           _ptr(out, C.c_double)))
    return out.reshape(tuple(int(v) for v in n), order="F")


def getMaximalFrequency(m, M):
    """src/GetHelmholtz.jl:75-79"""
    lib = L.load()
    mm = np.ascontiguousarray(np.atleast_1d(np.asarray(m, dtype=np.float64)).ravel())
    h = np.ascontiguousarray(np.asarray(M.h, dtype=np.float64))
    out = C.c_double()
    L.check(lib.hh_get_maximal_frequency(_ptr(mm, C.c_double), mm.size, int(M.dim), _ptr(h, C.c_double), C.byref(out)))
    return out.value


def GetHelmholtzOperator(*args, precision=ComplexF64, devices=None):
    """The three methods of src/GetHelmholtz.jl:14-16, 22-31, 33-50, dispatched on the argument list:

      GetHelmholtzOperator(Hparam[, orderNeumannBC])                                             -> H
      GetHelmholtzOperator(Msh, m, omega, gamma, NeumannAtFirstDim, ABLpad, ABLamp, Sommerfeld[, order]) -> (H, gamma)
      GetHelmholtzOperator(Msh, m, omega, gamma, NeumannAtFirstDim, Sommerfeld[, order])        -> H
    """
    if isinstance(args[0], HelmholtzParam):
        hp = args[0]
        order = args[1] if len(args) > 1 else 2
        hd = _Handle(hp.Mesh, hp.m, hp.omega, hp.gamma, hp.NeumannOnTop, hp.Sommerfeld, order, precision, devices)
        return HelmholtzOperator(hd)
    Msh, m, omega, gamma, neumann = args[:5]
    if len(args) >= 8:
        pad, amp, somm = args[5:8]
        order = args[8] if len(args) > 8 else 2
        abl = getABL(np.asarray(Msh.n) + 1, neumann, pad, amp)
        if gamma is None or np.size(gamma) == 0:
            gamma = abl
        else:
            gamma = np.asarray(gamma, dtype=np.float64).reshape(abl.shape, order="F") + abl
        hd = _Handle(Msh, m, omega, gamma, neumann, somm, order, precision, devices)
        return HelmholtzOperator(hd), gamma
    somm = args[5]
    order = args[6] if len(args) > 6 else 2
    hd = _Handle(Msh, m, omega, gamma, neumann, somm, order, precision, devices)
    return HelmholtzOperator(hd)


# ------------------------------------------------------------------ point sources
def loc2cs(n, sub):
    """src/getPointSource.jl:82-102 (1-based in and out)."""
    lib = L.load()
    n = np.ascontiguousarray(np.asarray(n, dtype=np.int64).ravel())
    sub = np.ascontiguousarray(np.asarray(sub, dtype=np.int64).ravel())
    return int(lib.hh_point_source_index(len(sub), _ptr(n, C.c_int64), _ptr(sub, C.c_int64)))


def getTopPointSrc(Minv):
    n = Minv.n
    if Minv.dim == 3:
        return [(int(n[0]) + 1) // 2, (int(n[1]) + 1) // 2, 1]
    return [(int(n[0]) + 1) // 2, 1]


def getMidPointSrc(Minv):
    n = Minv.n
    return [(int(v) + 1) // 2 for v in n]


def getAcousticPointSource(Minv, TYPE=ComplexF64, src=None):
    """src/getPointSource.jl:105-112"""
    if src is None:
        src = getTopPointSrc(Minv)
    nodes = np.asarray(Minv.n, dtype=np.int64) + 1
    q = np.zeros(tuple(int(v) for v in nodes), dtype=TYPE, order="F")
    q.reshape(-1, order="F")[loc2cs(nodes, src) - 1] = 1.0 / (np.linalg.norm(Minv.h) ** 2)
    return q, src


# ------------------------------------------------------------------ MGparam (Multigrid.jl boundary type)
class MGparam:
    """The Multigrid.MGparam fields the reference reads or mutates.  The hierarchy itself lives on the
    device behind `_hd`."""

    def __init__(self, VAL, levels, numCores, maxOuterIter, relativeTol, relaxType, relaxParam, relaxPre, relaxPost,
                 cycleType, coarseSolveType, strongConnParam=0.5, FilteringParam=0.0, transferOperatorType="FullWeighting",
                 coarseIters=10):
        self.VAL = np.dtype(VAL)
        self.levels = int(levels)
        self.numCores = int(numCores)
        self.maxOuterIter = int(maxOuterIter)
        self.relativeTol = float(relativeTol)
        self.relaxType = relaxType
        self.relaxParam = float(relaxParam)
        self.relaxPre = relaxPre
        self.relaxPost = relaxPost
        self.cycleType = cycleType
        self.coarseSolveType = coarseSolveType
        self.strongConnParam = strongConnParam
        self.FilteringParam = FilteringParam
        self.transferOperatorType = transferOperatorType
        self.coarseIters = int(coarseIters)
        self.cyclePrecision = None  # extension: ComplexF32 evaluates the cycle in single precision inside a ComplexF64 Krylov
        self.doTranspose = 0
        self._hd = None
        self._built_for = None

    def _options(self, shift, doTranspose):
        o = L.hh_mg_options()
        o.levels = self.levels
        try:
            o.relax_type = {"Jac": L.HH_RELAX_JAC, "Jac-GMRES": L.HH_RELAX_JAC_GMRES}[self.relaxType]
        except KeyError:
            raise ValueError(f"relaxType {self.relaxType!r} is not supported on the acoustic path (Jac, Jac-GMRES)")
        try:
            o.cycle_type = {"V": L.HH_CYCLE_V, "W": L.HH_CYCLE_W, "K": L.HH_CYCLE_K}[str(self.cycleType)]
        except KeyError:
            raise ValueError(f"cycleType {self.cycleType!r} is not supported (V, W, K)")
        try:
            o.coarse_type = {"NoMUMPS": L.HH_COARSE_LU, "Julia": L.HH_COARSE_LU, "GMRES": L.HH_COARSE_GMRES}[
                self.coarseSolveType]
        except KeyError:
            raise ValueError(f"coarseSolveType {self.coarseSolveType!r} is not supported (NoMUMPS, Julia, GMRES)")
        o.coarse_iters = self.coarseIters
        o.do_transpose = int(doTranspose)
        o.relax_param = self.relaxParam
        for l in range(L.HH_MAX_LEVELS):
            o.relax_pre[l] = int(self.relaxPre(l + 1)) if callable(self.relaxPre) else int(self.relaxPre)
            o.relax_post[l] = int(self.relaxPost(l + 1)) if callable(self.relaxPost) else int(self.relaxPost)
            o.shift[l] = float(shift[min(l, len(shift) - 1)])
        return o

    def _signature(self, shift, doTranspose):
        pre = tuple(int(self.relaxPre(l + 1)) if callable(self.relaxPre) else int(self.relaxPre) for l in range(self.levels))
        post = tuple(int(self.relaxPost(l + 1)) if callable(self.relaxPost) else int(self.relaxPost) for l in range(self.levels))
        return (self.levels, self.relaxType, self.relaxParam, pre, post, str(self.cycleType), self.coarseSolveType,
                self.coarseIters, float(shift[0]), int(doTranspose), str(self.cyclePrecision))


def getMGparam(*args, **kw):
    """Multigrid.getMGparam.  Accepts both call forms seen in the reference:
      getMGparam(VAL, IND, levels, numCores, maxIter, relativeTol, relaxType, relaxParam, relaxPre, relaxPost,
                 cycleType, coarseSolveType[, strongConnParam, FilteringParam, transferOperatorType])   (tests)
      getMGparam(levels, numCores, ...)                                                                  (examples)
    """
    args = list(args)
    VAL = ComplexF64
    if isinstance(args[0], type) or isinstance(args[0], np.dtype):
        VAL = args[0]
        args = args[2:]
    return MGparam(VAL, *args, **kw)


def hierarchyExists(MG):
    return MG._hd is not None and bool(MG._hd.lib.hh_hierarchy_exists(MG._hd.h))


def clear(obj):
    """clear!(MG) / clear!(solver) / clear!(HelmholtzParam)."""
    if isinstance(obj, MGparam):
        if obj._hd is not None:
            L.check(obj._hd.lib.hh_clear(obj._hd.h), obj._hd.h)
            obj._hd.close()
        obj._hd = None
        obj._built_for = None
    elif isinstance(obj, ShiftedLaplacianMultigridSolver):
        clear(obj.MG)  # src/ShiftedLaplacianMultigridSolver.jl:105-109
        clear(obj.helmParam)
        obj.doClear = 0
    elif isinstance(obj, HelmholtzParam):
        pass  # src/Helmholtz.jl:25-30 only clears the mesh's cached operators
    else:
        raise TypeError("clear!: unsupported object")


# ------------------------------------------------------------------ the solver plugin
class ShiftedLaplacianMultigridSolver:
    """src/ShiftedLaplacianMultigridSolver.jl:4-15"""

    def __init__(self, helmParam, MG, shift, Krylov="BiCGSTAB", inner=5, doClear=0, verbose=False, setupTime=0.0,
                 nPrec=0, solveTime=0.0):
        self.helmParam = helmParam
        self.MG = MG
        self.shift = np.asarray(shift, dtype=np.float64)
        self.Krylov = Krylov
        self.inner = int(inner)
        self.doClear = doClear
        self.verbose = verbose
        self.setupTime = setupTime
        self.nPrec = nPrec
        self.solveTime = solveTime
        # extras the reference only prints
        self.iterations = None
        self.relres = None
        self.devices = None
        # extension: one grid split into slabs over several GPUs, {"mode": "local", "devices": [...]} (this process
        # drives all slabs; B, X stay whole-grid arrays) or {"mode": "nccl", "rank", "nranks", "unique_id"} (one process
        # per GPU; B, X hold this rank's planes, see slabPlanes)
        self.slabs = None
        # set when solveLinearSystem is handed a GetHelmholtzOperatorHO operator: [beta_Laplacian, beta_mass]
        self.operatorHO = None


def getShiftedLaplacianMultigridSolver(helmParam, MG, shift, Krylov="BiCGSTAB", inner=5, verbose=False):
    """src/ShiftedLaplacianMultigridSolver.jl:24-30"""
    if np.isscalar(shift):
        shift = np.ones(MG.levels) * float(shift)
    return ShiftedLaplacianMultigridSolver(helmParam, MG, shift, Krylov, inner, 0, verbose, 0.0, 0, 0.0)


def copySolver(s):
    """src/ShiftedLaplacianMultigridSolver.jl:18-22: clone the settings, not the hierarchy."""
    MG = s.MG
    MG2 = MGparam(MG.VAL, MG.levels, MG.numCores, MG.maxOuterIter, MG.relativeTol, MG.relaxType, MG.relaxParam,
                  MG.relaxPre, MG.relaxPost, MG.cycleType, MG.coarseSolveType, MG.strongConnParam, MG.FilteringParam,
                  MG.transferOperatorType, MG.coarseIters)
    MG2.cyclePrecision = MG.cyclePrecision
    s2 = getShiftedLaplacianMultigridSolver(s.helmParam, MG2, s.shift, s.Krylov, s.inner, s.verbose)
    s2.devices = s.devices
    s2.slabs = s.slabs
    s2.operatorHO = s.operatorHO
    if hasattr(s, "orderNeumannBC"):
        s2.orderNeumannBC = s.orderNeumannBC
    return s2


def _ensure_hierarchy(param, doTranspose):
    MG = param.MG
    hp = param.helmParam
    sig = MG._signature(param.shift, doTranspose)
    if MG._hd is not None and getattr(MG._hd, "cycle_precision", None) != MG.cyclePrecision:
        clear(MG)  # the precision of the cycle is a property of the handle
    if MG._hd is not None and (MG._hd.slabs is not None or getattr(param, "slabs", None) is not None) and \
            (MG._hd.slabs is not getattr(param, "slabs", None) or MG._hd.levels != MG.levels):
        clear(MG)  # the slab partition depends on the number of levels
    if MG._hd is None:
        MG._hd = _Handle(hp.Mesh, hp.m, hp.omega, hp.gamma, hp.NeumannOnTop, hp.Sommerfeld,
                         getattr(param, "orderNeumannBC", 2), MG.VAL, param.devices,
                         MG.cyclePrecision, getattr(param, "slabs", None), MG.levels)
        MG._hd.cycle_precision = MG.cyclePrecision
        MG._hd.levels = MG.levels
    hd = MG._hd
    want_ho = getattr(param, "operatorHO", None)
    if getattr(hd, "ho_beta", None) != want_ho:
        # the fine-level operator of the handle: plain 5/7-point operator or GetHelmholtzOperatorHO (hh_set_operator_ho)
        mm = np.ascontiguousarray(np.asarray(hp.m, dtype=np.float64).ravel(order="F"))
        gg = np.ascontiguousarray(np.asarray(hp.gamma, dtype=np.float64).ravel(order="F"))
        bb = np.ascontiguousarray(np.asarray(want_ho if want_ho is not None else [1.0, 1.0], dtype=np.float64))
        L.check(hd.lib.hh_set_operator_ho(hd.h, int(want_ho is not None), _ptr(mm, C.c_double), _ptr(gg, C.c_double),
                                          _ptr(bb, C.c_double)), hd.h)
        hd.ho_beta = want_ho
        MG._built_for = None
    if (not hd.lib.hh_hierarchy_exists(hd.h)) or MG._built_for != sig:
        # first call (hierarchyExists == false, :50-66), a flipped doTranspose (transposeHierarchy, :68-70) or
        # settings mutated since the last solve (the tests mutate MG.relaxType / MG.cycleType, test :86-87,138-139)
        o = MG._options(param.shift, doTranspose)
        L.check(hd.lib.hh_setup(hd.h, C.byref(o)), hd.h)
        MG._built_for = sig
        MG.doTranspose = int(doTranspose)
    return hd


def solveLinearSystem_(ShiftedHT, B, X, param, doTranspose=0):
    """In-place variant (jInv.LinearSolvers.solveLinearSystem!): X is overwritten.  Returns (X, param)."""
    if np.iscomplexobj(param.helmParam.omega) and complex(param.helmParam.omega).imag != 0.0:
        # GetHelmholtzShiftOP(m, omega::Float64, shift) has no method for a complex omega
        # (src/ShiftedLaplacianMultigridSolver.jl:77)
        raise TypeError("solveLinearSystem: complex omega is not supported by the shifted-Laplacian solver")
    if param.Krylov not in ("GMRES", "BiCGSTAB"):
        raise ValueError(f"Krylov {param.Krylov!r} is not supported (GMRES, BiCGSTAB)")
    if isinstance(ShiftedHT, (HelmholtzOperator, HelmholtzOperatorHO)) and abs(ShiftedHT.shift - float(param.shift[0])) > 0:
        raise ValueError("the shifted operator passed in does not carry solver.shift[1]")
    if isinstance(ShiftedHT, HelmholtzOperatorHO):
        # the reference builds its hierarchy from the matrix it is handed (:65): an HO matrix gives an HO hierarchy
        param.operatorHO = list(ShiftedHT.beta)
    MG = param.MG
    if isinstance(ShiftedHT, HelmholtzOperator):
        # ... and so does the Neumann order of the operator it is handed (GetHelmholtzOperator(Hparam, orderNeumannBC))
        if ShiftedHT._hd.N != int(np.prod(np.asarray(param.helmParam.Mesh.n) + 1)):
            raise ValueError("the operator passed in lives on another mesh than solver.helmParam")
        if getattr(param, "orderNeumannBC", 2) != ShiftedHT._hd.order:
            param.orderNeumannBC = ShiftedHT._hd.order
            clear(MG)
    elif ShiftedHT is not None and hasattr(ShiftedHT, "shape") and not isinstance(ShiftedHT, HelmholtzOperatorHO):
        Nn = int(np.prod(np.asarray(param.helmParam.Mesh.n) + 1))
        if tuple(ShiftedHT.shape) != (Nn, Nn):
            raise ValueError(f"the matrix passed in is {tuple(ShiftedHT.shape)}, the solver's mesh has {Nn} nodes")
    if param.doClear == 1:
        clear(MG)
    torch_in = _is_torch_cuda(B)
    if torch_in and param.devices is None and MG._hd is None:
        param.devices = [B.device.index]  # the handle is created where the right-hand sides live
    # one process per slab: this process sees its planes of B only, the library detects zero columns after the all-reduce
    local_view = (getattr(param, "slabs", None) or {}).get("mode") == "nccl"
    if local_view:
        pass
    elif torch_in:
        import torch

        if float(torch.linalg.vector_norm(B)) == 0.0:  # :40-43
            X.zero_()
            return X, param
    else:
        # norm(B) == 0.0 -> zeros, no solve (:40-43).  A host pass over a multi-GB block costs more than the
        # PCIe copies, so large blocks skip it: the library treats zero columns the same way on the device
        # (X = 0, 0 iterations).
        if np.size(B) <= (1 << 22) and not np.any(B):
            X[...] = 0
            return X, param
    t0 = time.perf_counter()
    hd = _ensure_hierarchy(param, doTranspose)
    param.setupTime += time.perf_counter() - t0
    so = L.hh_solve_options()
    so.krylov = L.HH_KRYLOV_GMRES if param.Krylov == "GMRES" else L.HH_KRYLOV_BICGSTAB
    so.inner = max(int(param.inner), 1)
    so.max_iter = MG.maxOuterIter
    so.do_transpose = int(doTranspose)
    so.rel_tol = MG.relativeTol
    t0 = time.perf_counter()
    if torch_in:
        import torch

        # (N,), (N, nrhs) as the reference, or the zero-copy (nrhs, N) block with rows = right-hand sides
        Bt, layout = _torch_rows(B, hd, "B")
        if not (isinstance(X, torch.Tensor) and X.is_cuda and X.shape == B.shape and X.dtype == Bt.dtype):
            raise ValueError("solveLinearSystem!: X must be a CUDA tensor with the shape of B and the solver's precision")
        if X.device != B.device:
            raise ValueError("solveLinearSystem!: X and B live on different devices")
        direct = layout != "cols" and X.is_contiguous()
        Xt = X.reshape(-1, hd.N) if direct else torch.empty_like(Bt)
        nrhs = Bt.shape[0]
        iters = np.zeros(nrhs, dtype=np.int32)
        relres = np.zeros(nrhs, dtype=np.float64)
        rc = L.check(hd.lib.hh_solve_device(hd.h, Bt.data_ptr(), Xt.data_ptr(), nrhs, C.byref(so),
                                            _ptr(iters, C.c_int32), _ptr(relres, C.c_double)), hd.h)
        if not direct:
            X.copy_(Xt.t() if layout == "cols" else Xt.reshape(X.shape))
    else:
        Bm = _as_block(B, hd.N, hd.dtype)
        nrhs = Bm.shape[1]
        # write straight into the caller's X when it already is an N x nrhs column-major block of the right type
        direct = isinstance(X, np.ndarray) and X.dtype == hd.dtype and X.size == Bm.size and X.flags.f_contiguous
        Xm = X if direct else np.empty_like(Bm, order="F")
        iters = np.zeros(nrhs, dtype=np.int32)
        relres = np.zeros(nrhs, dtype=np.float64)
        rc = L.check(hd.lib.hh_solve(hd.h, Bm.ctypes.data, Xm.ctypes.data, nrhs, C.byref(so), _ptr(iters, C.c_int32),
                                     _ptr(relres, C.c_double)), hd.h)
        if not direct:
            X[...] = Xm.reshape(X.shape, order="F")
    param.solveTime += time.perf_counter() - t0
    param.nPrec += int(iters.sum())
    param.iterations = iters
    param.relres = relres
    if rc == L.HH_NOT_CONVERGED:
        print("WARNING: MG solver reached maximum iterations without convergence")  # :97-99
    return X, param


def solveLinearSystem(ShiftedHT, B, param, doTranspose=0):
    """src/ShiftedLaplacianMultigridSolver.jl:33-102.  The first argument (the adjoint of the shifted
    matrix in the reference) is accepted for signature compatibility; the hierarchy is built matrix-free
    from param.helmParam and param.shift[1].  Returns (X, param); X has the shape of B."""
    if _is_torch_cuda(B):
        import torch

        dt = torch.complex128 if np.dtype(param.MG.VAL) == np.complex128 else torch.complex64
        X = torch.empty(B.shape, dtype=dt, device=B.device)
    else:
        B = np.asarray(B)
        if B.ndim == 2 and B.shape[1] == 1:
            B = B[:, 0]  # :34-36
        X = np.empty(B.shape, dtype=np.dtype(param.MG.VAL), order="F")
    return solveLinearSystem_(ShiftedHT, B, X, param, doTranspose)


def solvePointSources(param, srcs, amplitudes=None, doTranspose=0):
    """Solve for point sources without materialising a dense B on the host (hh_solve_point_sources).
    srcs: list of 1-based subscripts; amplitude default 1/||h||^2 as getAcousticPointSource."""
    return solvePointSources_(param, srcs, None, amplitudes, doTranspose)


def solvePointSources_(param, srcs, X, amplitudes=None, doTranspose=0):
    """In-place form: X is an N x nrhs column-major host block of the solver's precision (e.g. a pinned buffer) that
    receives the solutions; None allocates one."""
    hd = _ensure_hierarchy(param, doTranspose)
    MG = param.MG
    Mesh = param.helmParam.Mesh
    nodes = np.asarray(Mesh.n, dtype=np.int64) + 1
    idx = np.array([loc2cs(nodes, s) for s in srcs], dtype=np.int64)
    nrhs = len(idx)
    if amplitudes is None:
        amplitudes = np.full(nrhs, 1.0 / (np.linalg.norm(Mesh.h) ** 2))
    val = np.ascontiguousarray(np.asarray(amplitudes, dtype=np.complex128))
    so = L.hh_solve_options()
    so.krylov = L.HH_KRYLOV_GMRES if param.Krylov == "GMRES" else L.HH_KRYLOV_BICGSTAB
    so.inner = max(int(param.inner), 1)
    so.max_iter = MG.maxOuterIter
    so.do_transpose = int(doTranspose)
    so.rel_tol = MG.relativeTol
    if X is None:
        X = np.empty((hd.N, nrhs), dtype=hd.dtype, order="F")
    elif not (isinstance(X, np.ndarray) and X.dtype == hd.dtype and X.shape == (hd.N, nrhs) and X.flags.f_contiguous):
        raise ValueError(f"X must be a column-major {hd.N} x {nrhs} numpy block of dtype {hd.dtype}")
    iters = np.zeros(nrhs, dtype=np.int32)
    relres = np.zeros(nrhs, dtype=np.float64)
    t0 = time.perf_counter()
    rc = L.check(hd.lib.hh_solve_point_sources(hd.h, _ptr(idx, C.c_int64), _ptr(val.view(np.float64), C.c_double), nrhs,
                                               X.ctypes.data, C.byref(so), _ptr(iters, C.c_int32),
                                               _ptr(relres, C.c_double)), hd.h)
    param.solveTime += time.perf_counter() - t0
    param.nPrec += int(iters.sum())
    param.iterations = iters
    param.relres = relres
    if rc == L.HH_NOT_CONVERGED:
        print("WARNING: MG solver reached maximum iterations without convergence")
    return X, param


# ------------------------------------------------------------------ device-side set-up (frequency sweeps)
def setFrequencyABL(param, omega, gamma_const, ABLpad, ABLamp, fetch_gamma=True):
    """New frequency on the model the solver's handle already holds: omega is replaced and
    gamma <- gamma_const + getABL(n+1, NeumannOnTop, ABLpad, ABLamp) is evaluated ON THE DEVICE (hh_set_frequency_abl;
    the 9-argument GetHelmholtzOperator of src/GetHelmholtz.jl:22-31 without a host pass over the grid or an upload).
    The hierarchy is invalidated; the next solve rebuilds it.  param.helmParam follows (gamma is read back unless
    fetch_gamma is False, in which case param.helmParam.gamma is left stale and only the live handle is valid)."""
    if np.iscomplexobj(omega) and complex(omega).imag != 0.0:
        raise TypeError("setFrequencyABL: the shifted-Laplacian solver needs a real omega")
    MG = param.MG
    if MG._hd is None:
        _ensure_hierarchy(param, MG.doTranspose)
    hd = MG._hd
    nodes = hd.nodes
    pad = np.ascontiguousarray(np.asarray(ABLpad, dtype=np.int64).ravel())
    if pad.size == 1:
        pad = np.repeat(pad, nodes.size)
    L.check(hd.lib.hh_set_frequency_abl(hd.h, float(np.real(omega)), 0.0, float(gamma_const), _ptr(pad, C.c_int64), float(ABLamp)), hd.h)
    MG._built_for = None
    hp = param.helmParam
    gamma = hp.gamma
    if fetch_gamma:
        g = np.empty(hd.N, dtype=np.float64)
        L.check(hd.lib.hh_get_gamma(hd.h, _ptr(g, C.c_double)), hd.h)
        gamma = g.reshape(np.shape(hp.gamma), order="F") if np.size(hp.gamma) == hd.N else g
    param.helmParam = HelmholtzParam(hp.Mesh, gamma, hp.m, float(np.real(omega)), hp.NeumannOnTop, hp.Sommerfeld)
    return param


def getMaximalFrequencyDevice(param):
    """getMaximalFrequency (src/GetHelmholtz.jl:75-79) from the model resident on the device."""
    MG = param.MG
    if MG._hd is None:
        _ensure_hierarchy(param, MG.doTranspose)
    out = C.c_double()
    L.check(MG._hd.lib.hh_get_maximal_frequency_device(MG._hd.h, C.byref(out)), MG._hd.h)
    return out.value


# ------------------------------------------------------------------ slab decomposition helpers
def slabUniqueId():
    """128-byte NCCL communicator id (call on one rank, broadcast to the others)."""
    buf = (C.c_char * 128)()
    L.check(L.load().hh_nccl_unique_id(C.cast(buf, C.c_void_p)), None)
    return bytes(buf)


def slabPartition(n3_nodes, levels, nranks, rank):
    """Plane geometry of slab `rank`: one dict per level (0 = fine) with own0, own1, koff, nloc, zb, ze, n2g."""
    out = np.zeros(7 * int(levels), dtype=np.int64)
    rc = L.load().hh_slab_partition(int(n3_nodes), int(levels), int(nranks), int(rank), _ptr(out, C.c_int64))
    if rc != 0:
        raise ValueError("no slab partition: cells of the last dimension must be divisible by 2^(levels-1) and the "
                         "coarsest level needs at least one cell per slab")
    keys = ("own0", "own1", "koff", "nloc", "zb", "ze", "n2g")
    return [dict(zip(keys, (int(v) for v in out[7 * l:7 * l + 7]))) for l in range(int(levels))]


def slabPlanes(param):
    """Planes [k0, k1) of the last dimension that B / X of this process hold (whole grid unless NCCL slabs)."""
    hd = _ensure_hierarchy(param, param.MG.doTranspose)
    return hd.planes


def GetHelmholtzOperatorHOStencil(Msh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld, beta=1.0):
    """GetHelmholtzOperatorHO (src/GetHelmholtz.jl:54-72) as a stored stencil: coef[s, node], s the offset index
    (d1+1) + 3(d2+1) (+ 9(d3+1)), ComplexF64, column-major nodes (hh_ho_stencil; host-side set-up)."""
    nodes = (np.asarray(Msh.n, dtype=np.int64) + 1).copy()
    dim = int(Msh.dim)
    N = int(np.prod(nodes))
    mm = np.ascontiguousarray(np.asarray(mNodal, dtype=np.float64).ravel(order="F"))
    gg = np.ascontiguousarray(np.asarray(gamma, dtype=np.float64).ravel(order="F"))
    if mm.size != N or gg.size != N:
        raise ValueError(f"m and gamma must have prod(n+1) = {N} entries")
    if np.isscalar(beta):
        if dim == 3 and beta != 1:
            raise ValueError("getSpreadNodalLaplacianAndMass: in 3-D beta is a pair (Laplacian, mass)")
        beta = [float(beta), float(beta)]
    bb = np.ascontiguousarray(np.asarray(beta, dtype=np.float64))
    h = np.ascontiguousarray(np.asarray(Msh.h, dtype=np.float64))
    w = complex(omega)
    coef = np.empty((3 ** dim, N), dtype=np.complex128)
    L.check(L.load().hh_ho_stencil(dim, _ptr(nodes, C.c_int64), _ptr(h, C.c_double), _ptr(mm, C.c_double), _ptr(gg, C.c_double),
                                   w.real, w.imag, int(bool(NeumannAtFirstDim)), int(bool(Sommerfeld)), _ptr(bb, C.c_double),
                                   _ptr(coef.view(np.float64), C.c_double)), None)
    return coef


class HelmholtzOperatorHO:
    """What GetHelmholtzOperatorHO returns (src/GetHelmholtz.jl:54-72), kept as the stored stencil: `H @ x` on the
    host (numpy, for checks), `H + GetHelmholtzShiftOP(...)`, `H.H`.  Passing it (or its adjoint view, as the reference's
    callers do) to solveLinearSystem makes the solver run on this operator: hierarchy = Galerkin hierarchy of the
    shifted operator, Krylov operator = H (hh_set_operator_ho)."""

    def __init__(self, Msh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld, beta, shift=0.0, adjoint=False, coef=None):
        self.Mesh, self.m, self.omega, self.gamma = Msh, np.asarray(mNodal, dtype=np.float64), omega, np.asarray(gamma, dtype=np.float64)
        self.NeumannOnTop, self.Sommerfeld = bool(NeumannAtFirstDim), bool(Sommerfeld)
        dim = int(Msh.dim)
        if np.isscalar(beta):
            if dim == 3 and beta != 1:
                raise ValueError("getSpreadNodalLaplacianAndMass: in 3-D beta is a pair (Laplacian, mass)")
            beta = [float(beta), float(beta)]
        self.beta = [float(beta[0]), float(beta[1] if dim == 3 else beta[0])]
        self.shift = float(shift)
        self.adjoint = bool(adjoint)
        self.nodes = np.asarray(Msh.n, dtype=np.int64) + 1
        N = int(np.prod(self.nodes))
        self.shape = (N, N)
        self.dtype = np.dtype(np.complex128)
        self._coef = coef

    def __add__(self, other):
        if isinstance(other, HelmholtzShiftOP):
            return HelmholtzOperatorHO(self.Mesh, self.m, self.omega, self.gamma, self.NeumannOnTop, self.Sommerfeld, self.beta,
                                       self.shift + other.shift, self.adjoint, self._coef)
        return NotImplemented

    __radd__ = __add__

    @property
    def H(self):
        return HelmholtzOperatorHO(self.Mesh, self.m, self.omega, self.gamma, self.NeumannOnTop, self.Sommerfeld, self.beta,
                                   self.shift, not self.adjoint, self._coef)

    def stencil(self):
        """coef[s, node] of the (shifted) operator, ComplexF64."""
        if self._coef is None:
            self._coef = GetHelmholtzOperatorHOStencil(self.Mesh, self.m, self.omega, self.gamma, self.NeumannOnTop,
                                                       self.Sommerfeld, self.beta if int(self.Mesh.dim) == 3 else self.beta[0])
        coef = self._coef
        if self.shift != 0.0:
            coef = coef.copy()
            coef[coef.shape[0] // 2] += 1j * self.shift * float(np.real(self.omega)) ** 2 * self.m.ravel(order="F")
        return coef

    def matvec(self, x):
        """y = H x (or H^H x for the adjoint view) on the host from the stored stencil."""
        coef = self.stencil()
        nd = [int(v) for v in self.nodes]
        dim = len(nd)
        X = np.asarray(x, dtype=np.complex128)
        vec = X.ndim == 1
        X = X.reshape((int(np.prod(nd)), -1), order="F")
        k = X.shape[1]
        Xg = X.reshape(nd + [k], order="F")
        Y = np.zeros_like(Xg)
        for s_ in range(coef.shape[0]):
            off = [s_ % 3 - 1, (s_ // 3) % 3 - 1] + ([s_ // 9 - 1] if dim == 3 else [])
            cg = coef[s_].reshape(nd, order="F")[..., None]
            src = tuple(slice(max(0, o), nd[d] + min(0, o)) for d, o in enumerate(off))    # nodes p + off
            dst = tuple(slice(max(0, -o), nd[d] + min(0, -o)) for d, o in enumerate(off))  # nodes p
            if self.adjoint:   # (H^H x)_q = sum_p conj(H[p, q]) x_p with q = p + off
                Y[src] += np.conj(cg[dst]) * Xg[dst]
            else:              # (H x)_p = sum_off coef[off][p] x_{p+off}
                Y[dst] += cg[dst] * Xg[src]
        Y = Y.reshape(X.shape, order="F")
        return Y[:, 0] if vec else Y

    __matmul__ = matvec
    __mul__ = matvec


def GetHelmholtzOperatorHO(*args):
    """src/GetHelmholtz.jl:18-20 and 54-72:
      GetHelmholtzOperatorHO(Hparam[, beta])                                                   -> H
      GetHelmholtzOperatorHO(Msh, m, omega, gamma, NeumannAtFirstDim, Sommerfeld[, beta])      -> H"""
    if isinstance(args[0], HelmholtzParam):
        hp = args[0]
        beta = args[1] if len(args) > 1 else 1.0
        return HelmholtzOperatorHO(hp.Mesh, np.asarray(hp.m).reshape(tuple(np.asarray(hp.Mesh.n) + 1), order="F"), hp.omega,
                                   np.asarray(hp.gamma).reshape(tuple(np.asarray(hp.Mesh.n) + 1), order="F"), hp.NeumannOnTop,
                                   hp.Sommerfeld, beta)
    Msh, m, omega, gamma, neumann, somm = args[:6]
    beta = args[6] if len(args) > 6 else 1.0
    return HelmholtzOperatorHO(Msh, m, omega, gamma, neumann, somm, beta)


def stencilAdjoint(nodes, coef):
    """Conjugate transpose of an operator stored as coef[s, node] (hh_stencil_adjoint; host-side)."""
    nodes = np.ascontiguousarray(np.asarray(nodes, dtype=np.int64))
    cin = np.ascontiguousarray(np.asarray(coef, dtype=np.complex128))
    out = np.empty_like(cin)
    L.check(L.load().hh_stencil_adjoint(len(nodes), _ptr(nodes, C.c_int64), _ptr(cin.view(np.float64), C.c_double),
                                        _ptr(out.view(np.float64), C.c_double)), None)
    return out


def GetHelmholtzMatrix(Msh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld, orderNeumannBC=2, shift=0.0, betaHO=None):
    """The sparse matrix the reference's GetHelmholtzOperator (or, with betaHO, GetHelmholtzOperatorHO) returns, plus
    GetHelmholtzShiftOP(m, omega, shift), as scipy.sparse.csc_matrix (hh_assemble_csc; host-side).  For code that uses
    the matrix itself, e.g. the direct solves of test/HelmholtzTest.jl."""
    import scipy.sparse as sp

    nodes = (np.asarray(Msh.n, dtype=np.int64) + 1).copy()
    dim = int(Msh.dim)
    N = int(np.prod(nodes))
    mm = np.ascontiguousarray(np.asarray(mNodal, dtype=np.float64).ravel(order="F"))
    gg = np.ascontiguousarray(np.asarray(gamma, dtype=np.float64).ravel(order="F"))
    if mm.size != N or gg.size != N:
        raise ValueError(f"m and gamma must have prod(n+1) = {N} entries")
    h = np.ascontiguousarray(np.asarray(Msh.h, dtype=np.float64))
    w = complex(omega)
    bb = None
    if betaHO is not None:
        if np.isscalar(betaHO):
            if dim == 3 and betaHO != 1:
                raise ValueError("getSpreadNodalLaplacianAndMass: in 3-D beta is a pair (Laplacian, mass)")
            betaHO = [float(betaHO), float(betaHO)]
        bb = np.ascontiguousarray(np.asarray(betaHO, dtype=np.float64))
    colptr = np.zeros(N + 1, dtype=np.int64)
    lib = L.load()

    def call(rowval, nzval):
        L.check(lib.hh_assemble_csc(dim, _ptr(nodes, C.c_int64), _ptr(h, C.c_double), _ptr(mm, C.c_double), _ptr(gg, C.c_double),
                                    w.real, w.imag, int(bool(NeumannAtFirstDim)), int(bool(Sommerfeld)), int(orderNeumannBC),
                                    float(shift), _ptr(bb, C.c_double) if bb is not None else None, _ptr(colptr, C.c_int64),
                                    _ptr(rowval, C.c_int64) if rowval is not None else None,
                                    _ptr(nzval.view(np.float64), C.c_double) if nzval is not None else None), None)

    call(None, None)
    nnz = int(colptr[N])
    rowval = np.empty(nnz, dtype=np.int64)
    nzval = np.empty(nnz, dtype=np.complex128)
    call(rowval, nzval)
    return sp.csc_matrix((nzval, rowval, colptr), shape=(N, N))


# ------------------------------------------------------------------ remaining exports of the reference on this path
def getSommerfeldBC(Msh, mNodal, omega, NeumannOnTop, orderNeumannBC=2):
    """src/GetHelmholtz.jl:222-247: -i*omega*(BC/h_d)*sqrt(m) accumulated on every boundary face (edges and corners
    add up), the first face of the last dimension skipped when NeumannOnTop.  Host-side set-up; the kernels evaluate the
    same term in place (hh_get_diagonal returns it folded into the diagonal)."""
    if orderNeumannBC not in (1, 2):
        raise ValueError("getNodalLaplacianMatrix: BC not supported")
    BC = 2.0 if orderNeumannBC == 2 else 1.0
    nodes = tuple(int(v) + 1 for v in Msh.n)
    m = np.asarray(mNodal, dtype=np.float64).reshape(nodes, order="F")
    h = np.asarray(Msh.h, dtype=np.float64)
    somm = np.zeros(nodes, dtype=np.complex128)
    dim = len(nodes)
    for d in range(dim):
        for side in (0, -1):
            if d == dim - 1 and side == 0 and NeumannOnTop:
                continue
            sl = [slice(None)] * dim
            sl[d] = side
            sl = tuple(sl)
            somm[sl] += -1j * float(omega) * (BC / h[d]) * np.sqrt(m[sl])
    return somm


def getHelmholtzFun(ShiftedHelmholtzT, ShiftMat, y=None, numCores=1):
    """src/GetHelmholtz.jl:85-95: the closure x -> ShiftedHelmholtzT' * x + ShiftMat * x.  The solver calls it with
    ShiftMat = -GetHelmholtzShiftOP(m, omega, shift), which makes it x -> H x (ShiftedLaplacianMultigridSolver.jl:77-83).
    ShiftedHelmholtzT is an operator object (its adjoint view is applied, as the reference's SpMatMul does), ShiftMat a
    GetHelmholtzShiftOP object; y, when given, receives the result (the reference's preallocated work block)."""
    op = ShiftedHelmholtzT.H + ShiftMat

    def Hfun(x):
        out = op.matvec(x)
        if y is not None:
            y[...] = np.asarray(out).reshape(y.shape, order="F") if not _is_torch_cuda(out) else out.reshape(y.shape)
            return y
        return out

    return Hfun


def getNodalLaplacianMatrix(Msh, orderNeumannBC=2):
    """src/PlainNodalLaplacian.jl:32-46: -laplacian on the nodal grid with ghost-eliminated Neumann rows, as the sparse
    matrix the reference returns (real CSC).  Host-side (hh_assemble_csc with a zero mass term)."""
    nodes = tuple(int(v) + 1 for v in Msh.n)
    z = np.zeros(nodes)
    return GetHelmholtzMatrix(Msh, z, 1.0, z, True, False, orderNeumannBC).real.tocsc()


def dxxMat(n, h, orderNeumannBC=2):
    """src/PlainNodalLaplacian.jl:18-30: the 1-D factor of getNodalLaplacianMatrix (n nodes, spacing h)."""
    import scipy.sparse as sp

    if orderNeumannBC not in (1, 2):
        raise ValueError("getNodalLaplacianMatrix: BC not supported")
    BC = 2.0 if orderNeumannBC == 2 else 1.0
    lo = -np.ones(n - 1)
    lo[-1] = -BC
    di = 2.0 * np.ones(n)
    di[0] = di[-1] = BC
    up = -np.ones(n - 1)
    up[0] = -BC
    return sp.diags([lo / h**2, di / h**2, up / h**2], [-1, 0, 1], format="csc")


def Lap2DStencil(x1, x2, x3, x4, x5, h1invsq, h2invsq):
    """src/PlainNodalLaplacian.jl:150-152: centre x1, dimension-1 neighbours x2/x3, dimension-2 neighbours x4/x5."""
    return (2 * h1invsq + 2 * h2invsq) * x1 - h1invsq * (x2 + x3) - h2invsq * (x4 + x5)


def multOpNeumann_(M, x, y, op=Lap2DStencil):
    """src/PlainNodalLaplacian.jl:155-188 (`multOpNeumann!`): the reference's own matrix-free 2-D stencil apply, ghost
    value = centre value (first-order Neumann); a no-op in 3-D, as in the reference.  Host-side, vectorised: `op` is
    called once on whole arrays.  The device path of the same operator is hh_apply with orderNeumannBC = 1."""
    if int(M.dim) != 2:
        return y
    n1, n2 = int(M.n[0]) + 1, int(M.n[1]) + 1
    X = np.asarray(x).reshape((n1, n2), order="F")
    P = np.pad(X, 1, mode="edge")  # ghost = centre of the boundary node
    out = op(X, P[:-2, 1:-1], P[2:, 1:-1], P[1:-1, :-2], P[1:-1, 2:], 1.0 / float(M.h[0]) ** 2, 1.0 / float(M.h[1]) ** 2)
    y[...] = np.asarray(out).reshape(np.shape(y), order="F")
    return y
