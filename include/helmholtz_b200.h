/*
 * helmholtz_b200.h -- C ABI of libhelmholtz_b200.so
 *
 * B200 (sm_100a) implementation of the acoustic shifted-Laplacian multigrid Helmholtz
 * solve path of JuliaInv/Helmholtz.jl.  This header is the drop-in boundary: a Julia
 * `ccall`, a Python `ctypes` or any other FFI binds exactly these symbols.  Only plain
 * pointers, sizes and POD structs cross it; no CUDA / torch types.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   HelmholtzParam                       src/Helmholtz.jl:13-20            -> hh_create
 *   getShiftedHelmholtzParam             src/Helmholtz.jl:32-34            -> `shifted` flag of hh_apply
 *   GetHelmholtzOperator (H*x)           src/GetHelmholtz.jl:33-50         -> hh_apply / hh_apply_device
 *   GetHelmholtzShiftOP                  src/GetHelmholtz.jl:81-83         -> folded into the shifted stencil
 *   getHelmholtzFun (Afun closure)       src/GetHelmholtz.jl:85-95         -> hh_apply(shifted=0)
 *   getABL                               src/GetHelmholtz.jl:97-220        -> hh_get_abl
 *   getSommerfeldBC                      src/GetHelmholtz.jl:222-247       -> computed in-kernel; hh_get_diagonal exposes it
 *   getMaximalFrequency                  src/GetHelmholtz.jl:75-79         -> hh_get_maximal_frequency[_device]
 *   getABL inside GetHelmholtzOperator   src/GetHelmholtz.jl:22-31          -> hh_set_frequency_abl (device-side)
 *   getAcousticPointSource / loc2cs      src/getPointSource.jl:82-112      -> hh_point_source_index / hh_solve_point_sources
 *   Multigrid.getMGparam / MGsetup       (un-vendored; call sites test/ShiftedLaplacianTest.jl:63-64,
 *                                         src/ShiftedLaplacianMultigridSolver.jl:65)   -> hh_mg_options / hh_setup
 *   solveLinearSystem                    src/ShiftedLaplacianMultigridSolver.jl:33-102 -> hh_solve / hh_solve_device
 *   clear!                               src/ShiftedLaplacianMultigridSolver.jl:105-109 -> hh_clear
 *   copySolver                           src/ShiftedLaplacianMultigridSolver.jl:18-22  -> hh_create on the same model (no hierarchy)
 *   GetHelmholtzOperatorHO               src/GetHelmholtz.jl:54-72                     -> hh_ho_stencil / hh_set_operator_ho
 *   the returned SparseMatrixCSC itself  src/GetHelmholtz.jl:49,71                     -> hh_assemble_csc (host; for H \ q callers)
 *   (no counterpart: one grid over several GPUs)                                       -> hh_create_slab_local / hh_create_slab_nccl
 *
 * Memory layout: all arrays are Julia `Array`s: column-major, node (i,j,k) (0-based here)
 * at i + j*n1 + k*n1*n2; B and X are N x nrhs column-major (each right-hand side
 * contiguous); complex numbers are interleaved (re,im) = ComplexF64 / ComplexF32.
 * The caller owns every host array; the library copies m and gamma at hh_create and owns
 * all device memory (lifetime = handle).
 *
 * Error handling: every function returns an int status.  0 = ok; >0 = soft condition
 * (HH_NOT_CONVERGED mirrors the reference's printed WARNING,
 * ShiftedLaplacianMultigridSolver.jl:97-99); <0 = hard error, message via hh_last_error.
 * No exception or exit() crosses the ABI.  There is no CPU fallback: without a CUDA device
 * every compute entry point fails with HH_ERR_CUDA.
 */
#ifndef HELMHOLTZ_B200_H
#define HELMHOLTZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HH_VERSION 100 /* major*10000 + minor*100 + patch */

/* status codes */
#define HH_OK 0
#define HH_NOT_CONVERGED 1
#define HH_ERR_ARG (-1)
#define HH_ERR_CUDA (-2)
#define HH_ERR_STATE (-3)
#define HH_ERR_NAN (-4)
#define HH_ERR_UNSUPPORTED (-5)
#define HH_ERR_ALLOC (-6)

/* precision of the solve (Multigrid.MGparam{VAL}: ComplexF64 in the tests, ComplexF32 in the
 * paper runs, examples/PointSourceADR/runExperiments.jl:77) */
#define HH_C64 0 /* ComplexF64: B, X are double[2] per entry */
#define HH_C32 1 /* ComplexF32: B, X are float[2]  per entry */
/* Opt-in extension (no counterpart in the reference): ComplexF64 API, Krylov vectors, operator applies and residuals,
 * with the multigrid cycle (the flexible preconditioner) evaluated in ComplexF32.  FGMRES / BiCGSTAB converge to the
 * same ComplexF64 tolerance; the cycle moves half the bytes. */
#define HH_C64_MIXED 2

/* MGparam.relaxType (test/ShiftedLaplacianTest.jl:55-56) */
#define HH_RELAX_JAC 0
#define HH_RELAX_JAC_GMRES 1
/* MGparam.cycleType (test/ShiftedLaplacianTest.jl:60,139) */
#define HH_CYCLE_V 0
#define HH_CYCLE_W 1
#define HH_CYCLE_K 2
/* MGparam.coarseSolveType: "NoMUMPS"/"Julia" -> LU, "GMRES" -> inexact (runExperiments.jl:396) */
#define HH_COARSE_LU 0
#define HH_COARSE_GMRES 1
/* ShiftedLaplacianMultigridSolver.Krylov (ShiftedLaplacianMultigridSolver.jl:88-94) */
#define HH_KRYLOV_GMRES 0
#define HH_KRYLOV_BICGSTAB 1

#define HH_MAX_LEVELS 12

typedef struct hh_handle_s* hh_handle_t;

/* Mirror of the Multigrid.MGparam fields the reference sets (getMGparam positional arguments,
 * test/ShiftedLaplacianTest.jl:63-64) plus the solver's shift vector
 * (ShiftedLaplacianMultigridSolver.jl:6,28-30; only shift[0] is live on the Galerkin path, :64-65). */
typedef struct hh_mg_options {
    int32_t levels;                   /* MGparam.levels                                  */
    int32_t relax_type;               /* HH_RELAX_*                                      */
    int32_t cycle_type;               /* HH_CYCLE_*                                      */
    int32_t coarse_type;              /* HH_COARSE_*                                     */
    int32_t coarse_iters;             /* GMRES steps of the inexact coarsest solve       */
    int32_t do_transpose;             /* 1: hierarchy of the adjoint operator            */
    int32_t relax_pre[HH_MAX_LEVELS];  /* MGparam.relaxPre  (per level; Int or l->f(l))   */
    int32_t relax_post[HH_MAX_LEVELS]; /* MGparam.relaxPost                               */
    double relax_param;               /* MGparam.relaxParam: Jacobi damping              */
    double shift[HH_MAX_LEVELS];      /* solver.shift (fraction of omega^2 m)            */
} hh_mg_options;

/* Krylov controls: solver.Krylov, solver.inner, MGparam.relativeTol, MGparam.maxOuterIter */
typedef struct hh_solve_options {
    int32_t krylov;      /* HH_KRYLOV_*                                                        */
    int32_t inner;       /* GMRES restart length (ignored by BiCGSTAB)                         */
    int32_t max_iter;    /* GMRES: restart cycles; BiCGSTAB: iterations (MGparam.maxOuterIter) */
    int32_t do_transpose; /* solveLinearSystem's doTranspose                                    */
    double rel_tol;      /* MGparam.relativeTol, on ||r||/||b||                                */
} hh_solve_options;

/* -------- library / error -------- */
int hh_version(void);
/* message of the last failing call on this thread (h may be NULL for hh_create failures) */
const char* hh_last_error(hh_handle_t h);
/* number of visible CUDA devices (0 and HH_OK when none) */
int hh_device_count(int* count);

/* -------- host-side set-up helpers (Float64, as in the reference) -------- */
/* getABL(n,NeumannAtFirstDim,ABLpad,ABLamp): gamma_out has prod(n_nodes) entries. */
int hh_get_abl(int dim, const int64_t* n_nodes, int neumann_on_top, const int64_t* pad, double amp, double* gamma_out);
/* getMaximalFrequency(m,Mesh) */
int hh_get_maximal_frequency(const double* m, int64_t n, int dim, const double* h, double* omega_max);
/* GetHelmholtzOperatorHO(Mesh, m, omega, gamma, NeumannAtFirstDim, Sommerfeld, beta) (src/GetHelmholtz.jl:54-72 with
 * getSpreadNodalLaplacianAndMass, src/PlainNodalLaplacian.jl:106-141) as a stored stencil: coef_out[s*N + node] complex
 * (re,im) Float64, s = (d1+1)+3(d2+1)(+9(d3+1)) as in hh_get_level_stencil, 9 (2-D) or 27 (3-D) entries per node.
 * beta: 2-D beta[0] (Laplacian and mass); 3-D beta[0] Laplacian, beta[1] mass (the reference's beta == 1 is {1,1}).
 * Host-side set-up (no device needed); hh_set_operator_ho runs the solver on this operator. */
int hh_ho_stencil(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma, double omega_re,
                  double omega_im, int neumann_on_top, int sommerfeld, const double* beta, double* coef_out);
/* The explicit matrix the reference's GetHelmholtzOperator (src/GetHelmholtz.jl:33-50; ho_beta == NULL, order_neumann_bc
 * in {1,2}) or GetHelmholtzOperatorHO (:54-72; ho_beta as in hh_ho_stencil) returns, plus i*shift*Re(w)^2*diag(m)
 * (GetHelmholtzShiftOP), in compressed sparse column form with 0-based indices (SparseMatrixCSC minus one): for callers
 * that use the matrix itself (H \ q in test/HelmholtzTest.jl:42).  Two calls: rowval = nzval = NULL fills colptr[N+1]
 * (nnz = colptr[N]); the second call fills rowval[nnz] and nzval[2*nnz] (re,im).  Host-side, no device needed. */
int hh_assemble_csc(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma, double omega_re,
                    double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc, double shift,
                    const double* ho_beta, int64_t* colptr, int64_t* rowval, double* nzval);
/* conjugate transpose of a stored stencil (layout of hh_ho_stencil): (A^H)[p,p+off] = conj(A[p+off,p]).  Host-side;
 * this is how the transposed hierarchy of the high-order operator is formed (doTranspose = 1,
 * src/ShiftedLaplacianMultigridSolver.jl:68-70). */
int hh_stencil_adjoint(int dim, const int64_t* n_nodes, const double* coef_in, double* coef_out);
/* loc2cs: 1-based subscripts -> 1-based linear index */
int64_t hh_point_source_index(int dim, const int64_t* n_nodes, const int64_t* sub);

/* -------- problem handle = HelmholtzParam + device state -------- */
/* m, gamma: Float64[N] host arrays (gamma already contains the absorbing layer, as
 * HelmholtzParam.gamma does).  omega may be complex.  order_neumann_bc in {1,2}.
 * device: CUDA ordinal.  precision: HH_C64 / HH_C32. */
int hh_create(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma, double omega_re,
              double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc, int precision, int device,
              hh_handle_t* out);
/* same, sharding right-hand sides over several devices of one box (one host thread and one
 * replica of the hierarchy per device; no data-path collective) */
int hh_create_multi(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma,
                    double omega_re, double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc,
                    int precision, const int* devices, int n_devices, hh_handle_t* out);
int hh_destroy(hh_handle_t h);

/* -------- slab decomposition of ONE grid over several GPUs (the grid that exceeds one GPU; no counterpart in the
 * reference, whose solve is single-node shared-memory: src/ShiftedLaplacianMultigridSolver.jl:33-102) --------
 * The planes of the last dimension are split into contiguous slabs, one per GPU; every stencil-type kernel is preceded
 * by a neighbour exchange of one halo plane per side and every dot / norm is all-reduced over the slabs, so each slab
 * runs the same batched Krylov iteration in lockstep.  `levels` fixes the partition (slab cuts fall on planes of the
 * coarsest level) and must equal hh_mg_options.levels at hh_setup; the coarsest solve must be HH_COARSE_GMRES.
 *
 * hh_create_slab_local: all slabs inside this process, one host thread per slab, halos by peer copies
 *   (devices may repeat: several slabs on one GPU).  The handle is used like any other: m, gamma, B, X and the
 *   arrays of hh_apply are whole-grid host arrays.
 * hh_create_slab_nccl: this process holds slab `rank` of `nranks` (one process per GPU, e.g. under torchrun or Julia
 *   Distributed); halos by ncclSend/ncclRecv, reductions by ncclAllReduce (libnccl.so.2 is bound at run time;
 *   HH_NCCL_LIB overrides the path).  `unique_id`: 128 bytes from hh_nccl_unique_id on rank 0, broadcast by the caller.
 *   m and gamma hold the planes model_plane0 <= k < model_plane0 + model_planes of the last dimension (0, 0 = the
 *   whole grid; a process that cannot hold the whole model passes the slab's planes koff .. koff + nloc - 1 of
 *   hh_slab_partition, halo planes included; hh_update_model uses the same convention);
 *   B, X (host or device) hold the planes own0 <= k < own1 of hh_slab_info only,
 *   i.e. n1*n2*(own1-own0) entries per right-hand side; point-source indices stay whole-grid indices.
 *   Every rank must make the same sequence of calls; a rank that fails (e.g. out of memory) leaves the others waiting
 *   in the next collective, as with any NCCL program: size the batch with the memory of the fullest GPU in mind. */
int hh_create_slab_local(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma,
                         double omega_re, double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc,
                         int precision, const int* devices, int n_slabs, int levels, hh_handle_t* out);
int hh_nccl_unique_id(void* id128);
int hh_create_slab_nccl(int dim, const int64_t* n_nodes, const double* h, const double* m, const double* gamma,
                        double omega_re, double omega_im, int neumann_on_top, int sommerfeld, int order_neumann_bc,
                        int precision, int device, int levels, int rank, int nranks, const void* unique_id,
                        int64_t model_plane0, int64_t model_planes, hh_handle_t* out);
/* mode: 0 no slabs, 1 local, 2 NCCL; planes own0 <= k < own1 of the last dimension are the ones the caller's B / X hold */
int hh_slab_info(hh_handle_t h, int* mode, int* n_slabs, int* rank, int64_t* own0, int64_t* own1);
/* host-only: the partition itself.  out[7*l + 0..6] = own0, own1, koff (global index of local plane 0), nloc (planes
 * held, halo planes included), zb, ze (owned planes in local numbering), n2g (planes of level l), for l < levels. */
int hh_slab_partition(int64_t n3_nodes, int levels, int nranks, int rank, int64_t* out);
/* Galerkin stencil of level >= 1 held by slab `slab` of this process (0 for an NCCL handle), in the layout of
 * hh_get_level_stencil on the slab's LOCAL grid n_local_out[3] (halo planes included; hh_slab_partition gives the
 * plane geometry).  coef_out may be NULL to query the node counts only.  Parity hook for MGsetup under slabs. */
int hh_slab_level_stencil(hh_handle_t h, int slab, int level, int64_t* n_local_out, void* coef_out);
/* run all work of this handle on `cuda_stream` (a cudaStream_t; NULL = legacy default stream) */
int hh_set_stream(hh_handle_t h, void* cuda_stream);
/* new model / frequency on the same grid: invalidates the hierarchy (clear! + new HelmholtzParam) */
int hh_update_model(hh_handle_t h, const double* m, const double* gamma, double omega_re, double omega_im);

/* Device-side set-up for frequency sweeps on a resident model: replaces omega and sets
 * gamma <- gamma_const + getABL(n, NeumannOnTop, pad, amp) on the device (the 9-argument GetHelmholtzOperator,
 * src/GetHelmholtz.jl:22-31, with getABL :97-220 evaluated from its separable 1-D ramps); m is kept.  Invalidates the
 * hierarchy like hh_update_model.  Not available with the high-order operator. */
int hh_set_frequency_abl(hh_handle_t h, double omega_re, double omega_im, double gamma_const, const int64_t* pad, double amp);
/* gamma as the device holds it, Float64[N] (NCCL slab handle: the planes the caller's B / X hold) */
int hh_get_gamma(hh_handle_t h, double* gamma_out);
/* getMaximalFrequency (src/GetHelmholtz.jl:75-79) of the model the handle holds on the device */
int hh_get_maximal_frequency_device(hh_handle_t h, double* omega_max);

/* GetHelmholtzOperatorHO (src/GetHelmholtz.jl:54-72) as the operator of this handle: the fine level becomes the stored
 * stencil of hh_ho_stencil (built at hh_setup from Float64 copies of m and gamma, the arrays hh_create was given), the
 * hierarchy is the Galerkin hierarchy of  H_HO + i*shift*w^2*diag(m)  (the matrix a reference caller passes to
 * solveLinearSystem, :33,65) and the Krylov operator is H_HO.  enable = 0 returns to the plain operator.  Invalidates
 * the hierarchy.  hh_apply then needs hh_setup first, supports shift 0 and the hierarchy's shift, and applies the
 * operator the hierarchy was built for (do_transpose of hh_setup); slab handles are not supported with this operator.  hh_get_level_stencil(level 0) returns the shifted
 * fine stencil. */
int hh_set_operator_ho(hh_handle_t h, int enable, const double* m, const double* gamma, const double* beta);

/* -------- multigrid hierarchy (MGsetup / clear!) -------- */
int hh_setup(hh_handle_t h, const hh_mg_options* opts);
int hh_clear(hh_handle_t h);
int hh_hierarchy_exists(hh_handle_t h);
/* node counts of level `level` (0 = fine) */
int hh_level_nodes(hh_handle_t h, int level, int64_t* n_nodes_out);
/* Galerkin stencil of level >= 1 as coef[s*N_l + node], s = (d1+1)+3(d2+1)(+9(d3+1)); complex,
 * in the handle's precision, copied to host.  Parity hook for MGsetup. */
int hh_get_level_stencil(hh_handle_t h, int level, void* coef_out);
/* complex diagonal c_p of the fine operator (mass + Sommerfeld [+ shift]) in ComplexF64 */
int hh_get_diagonal(hh_handle_t h, int shifted, double shift, double* diag_out);

/* -------- operator apply: Y = H X (shifted=0) or (H + i*shift*w^2*diag(m)) X -------- */
int hh_apply(hh_handle_t h, const void* X, void* Y, int64_t nrhs, int shifted, double shift, int transpose);
int hh_apply_device(hh_handle_t h, const void* dX, void* dY, int64_t nrhs, int shifted, double shift, int transpose);

/* -------- one multigrid cycle Z = M(B) from a zero guess (the preconditioner) -------- */
int hh_cycle(hh_handle_t h, const void* B, void* Z, int64_t nrhs);
int hh_cycle_device(hh_handle_t h, const void* dB, void* dZ, int64_t nrhs);

/* -------- solveLinearSystem -------- */
/* B, X: N x nrhs complex in the handle's precision.  iters_out[nrhs]: preconditioner
 * applications per RHS; relres_out[nrhs]: final relative residual estimate.  Either may be NULL.
 * Zero right-hand sides return X = 0 (ShiftedLaplacianMultigridSolver.jl:40-43). */
int hh_solve(hh_handle_t h, const void* B, void* X, int64_t nrhs, const hh_solve_options* opts, int32_t* iters_out,
             double* relres_out);
int hh_solve_device(hh_handle_t h, const void* dB, void* dX, int64_t nrhs, const hh_solve_options* opts,
                    int32_t* iters_out, double* relres_out);
/* point sources: RHS r has one non-zero val[r] (re,im as double[2]) at 1-based linear index idx[r];
 * avoids shipping a dense N x nrhs B.  X is a host array. */
int hh_solve_point_sources(hh_handle_t h, const int64_t* idx, const double* val, int64_t nrhs, void* X,
                           const hh_solve_options* opts, int32_t* iters_out, double* relres_out);

/* -------- counters (solver.setupTime / solveTime / nPrec, ShiftedLaplacianMultigridSolver.jl:12-14) -------- */
int hh_get_counters(hh_handle_t h, double* setup_seconds, double* solve_seconds, int64_t* n_prec,
                    int64_t* kernel_launches);

/* -------- per-kernel device timing (CUDA events on the launching stream) -------- */
int hh_profile_enable(hh_handle_t h, int on);
int hh_profile_reset(hh_handle_t h);
int hh_profile_num_tags(void);
const char* hh_profile_tag_name(int tag);
/* launches, summed device milliseconds and summed algorithmic bytes of kernel class `tag` */
int hh_profile_get(hh_handle_t h, int tag, int64_t* launches, double* milliseconds, double* algorithmic_bytes);
/* finer view: one entry per (kernel class, algorithmic bytes per launch), i.e. launches doing identical work */
int hh_profile_num_entries(hh_handle_t h);
int hh_profile_entry(hh_handle_t h, int index, int* tag, int64_t* launches, double* milliseconds,
                     double* algorithmic_bytes_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* HELMHOLTZ_B200_H */
