// hh_solver.cuh -- host-side orchestration of the device-resident solve: hierarchy set-up
// (MGsetup), multigrid cycle (recursiveCycle), batched FGMRES / BiCGSTAB (solveGMRES_MG /
// solveBiCGSTAB_MG).  Mirrors the control flow of src/ShiftedLaplacianMultigridSolver.jl:33-102;
// the arithmetic of the un-vendored Multigrid.jl / KrylovMethods.jl is restated from the published
// algorithms.
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda.h>  // CUtensorMap types only; the encode function is fetched through the runtime

#include "../../include/helmholtz_b200.h"
#include "hh_kernels.cuh"

namespace hh {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define HH_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            throw hh::Error(HH_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" +    \
                                             __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
    } while (0)

#define HH_REQUIRE(cond, code, msg)                      \
    do {                                                 \
        if (!(cond)) throw hh::Error((code), (msg));     \
    } while (0)

// GetHelmholtzOperatorHO (src/GetHelmholtz.jl:54-72 with getSpreadNodalLaplacianAndMass, src/PlainNodalLaplacian.jl:106-141)
// as a stored stencil.  The Kronecker construction collapses to sums of tensor products of 1-D tridiagonals, so ONE
// entry of the operator is a closed form of the row node, the offset and the model at the column node: ho_coef below is
// that closed form, compiled for the host (hh_ho_stencil, hh_assemble_csc: pinned against the oracle's Kronecker
// assembly by the CPU tests) and for the device (k_ho_stencil: the solver's own set-up).
struct HoGeom {
    int dim;
    int n[3];
    double ih2[3];   // 1/h_d^2
    double hs[3];    // h_d
    double bl, bm;   // beta of the Laplacian spread and of the mass average
    double wre, w2r, w2i;
    int neumann_top, sommerfeld;
};
inline HoGeom ho_geom(int dim, const int64_t* n_nodes, const double* hsp, double wre, double wim, int neumann_on_top, int sommerfeld,
                      const double* beta) {
    HoGeom g{};
    g.dim = dim;
    for (int d = 0; d < 3; ++d) {
        g.n[d] = d < dim ? (int)n_nodes[d] : 1;
        g.ih2[d] = d < dim ? 1.0 / (hsp[d] * hsp[d]) : 0.0;
        g.hs[d] = d < dim ? hsp[d] : 1.0;
    }
    g.bl = beta[0];
    g.bm = dim == 3 ? beta[1] : beta[0];
    g.wre = wre;
    g.w2r = wre * wre - wim * wim;
    g.w2i = 2.0 * wre * wim;
    g.neumann_top = neumann_on_top;
    g.sommerfeld = sommerfeld;
    return g;
}
// entry of a 1-D tridiagonal table that couples node i with node i+o (o in {-1,0,1}); zero when i+o is outside
__host__ __device__ __forceinline__ double ho_tri(int i, int n, int o, double diag_in, double diag_end, double off) {
    if (o == 0) return (i == 0 || i == n - 1) ? diag_end : diag_in;
    if (o < 0) return i > 0 ? off : 0.0;
    return i < n - 1 ? off : 0.0;
}
// mass = -w^2 m (1 - i gamma / Re w) - Sommerfeld at node (i,j,k)   (GetHelmholtz.jl:62-68; getSommerfeldBC :222-247, BC = 2)
__host__ __device__ __forceinline__ void ho_mass(const HoGeom& g, double mp, double gp, int i, int j, int k, double& re, double& im) {
    const double gg = gp / g.wre;
    re = -mp * (g.w2r + g.w2i * gg);
    im = -mp * (g.w2i - g.w2r * gg);
    if (g.sommerfeld) {
        double sf = 0.0;
        const int idx[3] = {i, j, k};
        for (int d = 0; d < g.dim; ++d) {
            const bool first = idx[d] == 0, last = idx[d] == g.n[d] - 1;
            const bool top = (d == g.dim - 1) && g.neumann_top;
            if ((first && !top) || last) sf += 2.0 / g.hs[d];
        }
        im += g.wre * sf * sqrt(mp);
    }
}
// H_HO[(i,j,k), (i+di,j+dj,k+dk)] (+ i*shift_w2*m on the diagonal); m / gamma are dense column-major arrays.
// Returns false (entry zero) when the column node is outside the grid.
__host__ __device__ __forceinline__ bool ho_coef(const HoGeom& g, const double* __restrict__ m, const double* __restrict__ gam, int i,
                                                 int j, int k, int di, int dj, int dk, double shift_w2, double& re, double& im) {
    re = im = 0.0;
    const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    if (i + di < 0 || i + di >= n0 || j + dj < 0 || j + dj >= n1 || k + dk < 0 || k + dk >= n2) return false;
    const double I0 = di == 0, J0 = dj == 0, K0 = dk == 0;  // Kronecker deltas
    const double bl = g.bl, bm = g.bm;
    // T = ddxCN' * ddxCN, A = av3term(n, 1/2), B = av3term(n, beta_mass)
    const double t0 = ho_tri(i, n0, di, 2.0 * g.ih2[0], g.ih2[0], -g.ih2[0]);
    const double t1 = ho_tri(j, n1, dj, 2.0 * g.ih2[1], g.ih2[1], -g.ih2[1]);
    const double a0 = ho_tri(i, n0, di, 0.5, 0.75, 0.25), a1 = ho_tri(j, n1, dj, 0.5, 0.75, 0.25);
    const double b0 = ho_tri(i, n0, di, bm, 0.5 + 0.5 * bm, 0.5 * (1.0 - bm));
    const double b1 = ho_tri(j, n1, dj, bm, 0.5 + 0.5 * bm, 0.5 * (1.0 - bm));
    double lap, mm;
    if (g.dim == 2) {
        lap = t0 * ((1.0 - bl) * a1 + bl * J0) + t1 * ((1.0 - bl) * a0 + bl * I0);
        mm = 0.5 * (b1 * I0 + b0 * J0);
    } else {
        const double t2 = ho_tri(k, n2, dk, 2.0 * g.ih2[2], g.ih2[2], -g.ih2[2]);
        const double a2 = ho_tri(k, n2, dk, 0.5, 0.75, 0.25);
        const double b2 = ho_tri(k, n2, dk, bm, 0.5 + 0.5 * bm, 0.5 * (1.0 - bm));
        lap = t0 * (bl * J0 * K0 + 0.5 * (1.0 - bl) * (a1 * K0 + J0 * a2)) +
              t1 * (bl * I0 * K0 + 0.5 * (1.0 - bl) * (a0 * K0 + I0 * a2)) +
              t2 * (bl * I0 * J0 + 0.5 * (1.0 - bl) * (a0 * J0 + I0 * a1));
        mm = (1.0 / 3.0) * (b1 * I0 * K0 + b0 * J0 * K0 + b2 * I0 * J0);
    }
    const int64_t q = (int64_t)(i + di) + (int64_t)n0 * ((j + dj) + (int64_t)n1 * (k + dk));  // column node: M * Diagonal(mass)
    double mr, mi;
    ho_mass(g, m[q], gam[q], i + di, j + dj, k + dk, mr, mi);
    re = lap + mm * mr;
    im = mm * mi;
    if (di == 0 && dj == 0 && dk == 0) im += shift_w2 * m[q];
    return true;
}
// host side (hh_ho_stencil in include/helmholtz_b200.h): coef_out[2*(s*N + node) + {0,1}], Float64
inline void build_ho_stencil(int dim, const int64_t* n_nodes, const double* hsp, const double* m, const double* gamma,
                             double wre, double wim, int neumann_on_top, int sommerfeld, const double* beta,
                             double* coef_out) {
    HH_REQUIRE(dim == 2 || dim == 3, HH_ERR_ARG, "hh_ho_stencil: dim must be 2 or 3");
    HH_REQUIRE(wre != 0.0, HH_ERR_ARG, "hh_ho_stencil: Re(omega) must be non-zero");
    for (int d = 0; d < dim; ++d)
        HH_REQUIRE(n_nodes[d] >= 2 && hsp[d] > 0.0, HH_ERR_ARG, "hh_ho_stencil: node counts must be >= 2, spacings positive");
    const HoGeom g = ho_geom(dim, n_nodes, hsp, wre, wim, neumann_on_top, sommerfeld, beta);
    const int64_t N = (int64_t)g.n[0] * g.n[1] * g.n[2];
    const int NS = dim == 3 ? 27 : 9;
    for (int s = 0; s < NS; ++s) {
        const int di = s % 3 - 1, dj = (s / 3) % 3 - 1, dk = dim == 3 ? s / 9 - 1 : 0;
        for (int k = 0; k < g.n[2]; ++k)
            for (int j = 0; j < g.n[1]; ++j)
                for (int i = 0; i < g.n[0]; ++i) {
                    const int64_t p = i + (int64_t)g.n[0] * (j + (int64_t)g.n[1] * k);
                    double re, im;
                    ho_coef(g, m, gamma, i, j, k, di, dj, dk, 0.0, re, im);
                    coef_out[2 * ((int64_t)s * N + p)] = re;
                    coef_out[2 * ((int64_t)s * N + p) + 1] = im;
                }
    }
}

// device side (the solver's own set-up): one thread per node writes its 9 / 27 entries into the level's arrays (row
// pitch p0, leading dimension ldN per coefficient array, precision T).  adjoint != 0 writes the conjugate transpose,
// (A^H)[p, p+off] = conj(A[p+off, p]): the entry of row p+off at offset -off (doTranspose = 1).
template <typename T>
__global__ void __launch_bounds__(256) k_ho_stencil(HoGeom g, const double* __restrict__ m, const double* __restrict__ gam,
                                                    double shift_w2, int adjoint, int p0, int64_t ldN, cx<T>* __restrict__ coef) {
    const int64_t Nd = (int64_t)g.n[0] * g.n[1] * g.n[2];
    const int NS = g.dim == 3 ? 27 : 9;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < Nd; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p % g.n[0]);
        const int64_t t = p / g.n[0];
        const int j = (int)(t % g.n[1]), k = (int)(t / g.n[1]);
        const int64_t dst = i + (int64_t)p0 * (j + (int64_t)g.n[1] * k);
        for (int s = 0; s < NS; ++s) {
            const int di = s % 3 - 1, dj = (s / 3) % 3 - 1, dk = g.dim == 3 ? s / 9 - 1 : 0;
            double re = 0.0, im = 0.0;
            if (!adjoint) {
                ho_coef(g, m, gam, i, j, k, di, dj, dk, shift_w2, re, im);
            } else if (i + di >= 0 && i + di < g.n[0] && j + dj >= 0 && j + dj < g.n[1] && k + dk >= 0 && k + dk < g.n[2]) {
                ho_coef(g, m, gam, i + di, j + dj, k + dk, -di, -dj, -dk, shift_w2, re, im);
                im = -im;
            }
            coef[(int64_t)s * ldN + dst] = mk<T>((T)re, (T)im);
        }
    }
}

// Conjugate transpose of an operator stored as a 3^dim-point stencil on a dense column-major grid:
// (A^H)[p, p+off] = conj(A[p+off, p]) = conj(coef[-off][p+off]).  Host side, Float64 (re,im) pairs.
inline void adjoint_stencil(int dim, const int64_t* n_nodes, const double* in, double* out) {
    HH_REQUIRE((dim == 2 || dim == 3) && n_nodes && in && out && in != out, HH_ERR_ARG, "hh_stencil_adjoint: bad arguments");
    int64_t n[3] = {n_nodes[0], n_nodes[1], dim == 3 ? n_nodes[2] : 1};
    const int64_t N = n[0] * n[1] * n[2];
    const int NS = dim == 3 ? 27 : 9;
    for (int s = 0; s < NS; ++s) {
        const int di = s % 3 - 1, dj = (s / 3) % 3 - 1, dk = dim == 3 ? s / 9 - 1 : 0;
        const int sm = NS - 1 - s;  // index of the opposite offset
        for (int64_t k = 0; k < n[2]; ++k)
            for (int64_t j = 0; j < n[1]; ++j)
                for (int64_t i = 0; i < n[0]; ++i) {
                    const int64_t p = i + n[0] * (j + n[1] * k);
                    const bool inside = i + di >= 0 && i + di < n[0] && j + dj >= 0 && j + dj < n[1] && k + dk >= 0 && k + dk < n[2];
                    if (!inside) {
                        out[2 * ((int64_t)s * N + p)] = 0.0;
                        out[2 * ((int64_t)s * N + p) + 1] = 0.0;
                        continue;
                    }
                    const int64_t q = p + di + n[0] * (dj + n[1] * (int64_t)dk);
                    out[2 * ((int64_t)s * N + p)] = in[2 * ((int64_t)sm * N + q)];
                    out[2 * ((int64_t)s * N + p) + 1] = -in[2 * ((int64_t)sm * N + q) + 1];
                }
    }
}

// GetHelmholtzOperator (src/GetHelmholtz.jl:33-50 on the Laplacian of src/PlainNodalLaplacian.jl:18-46) as a stored
// stencil on the host, Float64: the same formulas the matrix-free kernels evaluate (fine_center / fine_w), for callers
// that need the explicit matrix (hh_assemble_csc).  shift adds i*shift*Re(w)^2*m to the diagonal (GetHelmholtzShiftOP).
inline void build_plain_stencil(int dim, const int64_t* n_nodes, const double* hsp, const double* m, const double* gamma,
                                double wre, double wim, int neumann_on_top, int sommerfeld, int order_bc, double shift,
                                double* coef_out) {
    HH_REQUIRE(dim == 2 || dim == 3, HH_ERR_ARG, "dim must be 2 or 3");
    HH_REQUIRE(order_bc == 1 || order_bc == 2, HH_ERR_ARG, "getNodalLaplacianMatrix: BC not supported");
    HH_REQUIRE(wre != 0.0, HH_ERR_ARG, "Re(omega) must be non-zero");
    int64_t n[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) {
        HH_REQUIRE(n_nodes[d] >= 2 && hsp[d] > 0.0, HH_ERR_ARG, "node counts must be >= 2, spacings positive");
        n[d] = n_nodes[d];
    }
    const int64_t N = n[0] * n[1] * n[2];
    const int NS = dim == 3 ? 27 : 9;
    const double BC = order_bc == 2 ? 2.0 : 1.0;
    const double w2r = wre * wre - wim * wim, w2i = 2.0 * wre * wim;
    std::fill(coef_out, coef_out + (size_t)2 * NS * N, 0.0);
    const int center = dim == 3 ? 13 : 4;
    const int64_t stride[3] = {1, 3, 9};
    for (int64_t k = 0; k < n[2]; ++k)
        for (int64_t j = 0; j < n[1]; ++j)
            for (int64_t i = 0; i < n[0]; ++i) {
                const int64_t p = i + n[0] * (j + n[1] * k);
                const int64_t idx[3] = {i, j, k};
                const double g = gamma[p] / wre;
                double re = -m[p] * (w2r + w2i * g), im = -m[p] * (w2i - w2r * g) + shift * wre * wre * m[p];
                double sf = 0.0;
                for (int d = 0; d < dim; ++d) {
                    const bool first = idx[d] == 0, last = idx[d] == n[d] - 1;
                    const double ih2 = 1.0 / (hsp[d] * hsp[d]);
                    re += ((first || last) ? BC : 2.0) * ih2;                         // dxxMat diagonal
                    if (!first) coef_out[2 * ((center - stride[d]) * N + p)] = -(last ? BC : 1.0) * ih2;   // sub-diagonal
                    if (!last) coef_out[2 * ((center + stride[d]) * N + p)] = -(first ? BC : 1.0) * ih2;   // super-diagonal
                    const bool top = (d == dim - 1) && neumann_on_top;
                    if (sommerfeld && ((first && !top) || last)) sf += 2.0 / hsp[d];  // getSommerfeldBC is called with BC = 2
                }
                im += wre * sf * std::sqrt(m[p]);
                coef_out[2 * (center * N + p)] = re;
                coef_out[2 * (center * N + p) + 1] = im;
            }
}

// stored stencil -> compressed sparse column arrays (0-based), entries with a zero coefficient dropped.  Two passes:
// rowval == nullptr only counts (colptr[N] = nnz).
inline int64_t stencil_to_csc(int dim, const int64_t* n_nodes, const double* coef, int64_t* colptr, int64_t* rowval, double* nzval) {
    int64_t n[3] = {n_nodes[0], n_nodes[1], dim == 3 ? n_nodes[2] : 1};
    const int64_t N = n[0] * n[1] * n[2];
    const int NS = dim == 3 ? 27 : 9;
    int64_t nnz = 0;
    for (int64_t kq = 0; kq < n[2]; ++kq)
        for (int64_t jq = 0; jq < n[1]; ++jq)
            for (int64_t iq = 0; iq < n[0]; ++iq) {
                const int64_t q = iq + n[0] * (jq + n[1] * kq);
                colptr[q] = nnz;
                // rows p = q - off in ascending order: offsets from (+1,+1,+1) down to (-1,-1,-1)
                for (int s = NS - 1; s >= 0; --s) {
                    const int di = s % 3 - 1, dj = (s / 3) % 3 - 1, dk = dim == 3 ? s / 9 - 1 : 0;
                    const int64_t ip = iq - di, jp = jq - dj, kp = kq - dk;
                    if (ip < 0 || ip >= n[0] || jp < 0 || jp >= n[1] || kp < 0 || kp >= n[2]) continue;
                    const int64_t p = ip + n[0] * (jp + n[1] * kp);
                    const double re = coef[2 * ((int64_t)s * N + p)], im = coef[2 * ((int64_t)s * N + p) + 1];
                    if (re == 0.0 && im == 0.0) continue;
                    if (rowval) {
                        rowval[nnz] = p;
                        nzval[2 * nnz] = re;
                        nzval[2 * nnz + 1] = im;
                    }
                    ++nnz;
                }
            }
    colptr[N] = nnz;
    return nnz;
}

}  // namespace hh
#include "hh_slab.cuh"
namespace hh {

// kernel classes for the per-kernel device timing (hh_profile_*)
enum Tag {
    T_FINE_APPLY = 0,
    T_FINE_RESID,
    T_FINE_JACOBI,
    T_FINE_JACOBI0,
    T_FINE_FIRST_RESID,
    T_FINE_FIRST_JACOBI,
    T_FINE_PROLONG_JACOBI,
    T_FINE_PROLONG_JACOBI2,
    T_COARSE_APPLY,
    T_COARSE_RESID,
    T_COARSE_JACOBI,
    T_COARSE_JACOBI0,
    T_RESTRICT,
    T_PROLONG,
    T_COARSEST_DENSE,
    T_DOT,
    T_AXPY,
    T_COPY,
    T_SCALAR,
    T_SETUP,
    T_HALO,
    T_ALLREDUCE,
    T_NTAGS
};
static const char* const kTagNames[T_NTAGS] = {
    "fine_apply",   "fine_resid",   "fine_jacobi",    "fine_jacobi0", "fine_first_resid", "fine_first_jacobi", "fine_prolong_jacobi", "fine_prolong_jacobi2",
    "coarse_apply", "coarse_resid",
    "coarse_jacobi", "coarse_jacobi0", "restrict",    "prolong",      "coarsest_dense", "krylov_dot",
    "krylov_axpy",  "copy",           "scalar",       "setup",        "halo_exchange", "allreduce"};

struct Profiler {
    struct Rec {
        int tag;
        cudaEvent_t a, b;
        double bytes;
        int64_t sub;
    };
    // launches with identical work: (kernel class, algorithmic bytes per launch)
    struct Entry {
        int tag;
        int64_t sub;
        int64_t count;
        double ms, bytes;
    };
    std::vector<Entry> entries;
    bool on = false;
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    int64_t count[T_NTAGS] = {0};
    double ms[T_NTAGS] = {0};
    double bytes[T_NTAGS] = {0};
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        HH_CUDA(cudaEventCreate(&e));
        return e;
    }
    cudaStream_t aux = nullptr;  // second stream events may have been recorded on (overlapped halo exchanges)
    void flush(cudaStream_t st) {
        if (recs.empty()) return;
        HH_CUDA(cudaStreamSynchronize(st));
        if (aux) HH_CUDA(cudaStreamSynchronize(aux));
        for (auto& r : recs) {
            float t = 0.f;
            HH_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
            count[r.tag] += 1;
            ms[r.tag] += t;
            bytes[r.tag] += r.bytes;
            bool found = false;
            for (auto& e : entries)
                if (e.tag == r.tag && e.sub == r.sub) {
                    e.count += 1;
                    e.ms += t;
                    e.bytes += r.bytes;
                    found = true;
                    break;
                }
            if (!found) entries.push_back({r.tag, r.sub, 1, (double)t, r.bytes});
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
    }
    void reset() {
        recs.clear();
        entries.clear();
        for (int i = 0; i < T_NTAGS; ++i) count[i] = 0, ms[i] = 0, bytes[i] = 0;
    }
    ~Profiler() {
        for (auto& r : recs) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        for (auto e : pool) cudaEventDestroy(e);
    }
};

template <typename U>
struct DevBuf {
    U* p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        if (count == 0) return;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(U));
        if (e != cudaSuccess) {
            p = nullptr;
            throw Error(HH_ERR_ALLOC, std::string("cudaMalloc of ") + std::to_string(count * sizeof(U)) +
                                          " bytes failed: " + cudaGetErrorString(e));
        }
        n = count;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) {
        o.p = nullptr;
        o.n = 0;
    }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
};

// Problem description shared by both precisions (HelmholtzParam, src/Helmholtz.jl:13-20)
struct Problem {
    int dim = 0;
    int n[3] = {1, 1, 1};
    double h[3] = {1, 1, 1};
    double w_re = 0, w_im = 0;
    int neumann_top = 0, sommerfeld = 0, order_bc = 2;
    int64_t N() const { return (int64_t)n[0] * n[1] * n[2]; }
};

// 1-D profiles of getABL (src/GetHelmholtz.jl:141-218), Float64 on the host: tab[0..3] as k_gamma_abl reads them
// (n = GLOBAL node counts).  The whole-array form hh_get_abl and the device-side hh_set_frequency_abl share them.
inline void abl_tables(int dim, const int64_t* n, int neumann_on_top, const int64_t* pad, std::vector<double> tab[4]) {
    HH_REQUIRE((dim == 2 || dim == 3) && n && pad, HH_ERR_ARG, "getABL: bad arguments");
    for (int d = 0; d < dim; ++d) HH_REQUIRE(n[d] >= 2 && pad[d] >= 1 && pad[d] <= n[d], HH_ERR_ARG, "getABL: pad must be in 1..n");
    if (dim == 2) {
        const int64_t n1 = n[0], n2 = n[1], p1 = pad[0], p2 = pad[1];
        tab[0].assign(n1, 0.0);  // left ramp of dim 1: ((p1..1)/p1)^2
        tab[1].assign(n1, 0.0);  // right ramp: ((1..p1)/p1)^2
        tab[2].assign(n2, 0.0);  // top ramp of dim 2 (absent under NeumannAtFirstDim)
        tab[3].assign(n2, 0.0);  // bottom ramp
        for (int64_t t = 0; t < p1; ++t) {
            tab[0][t] = (double)((p1 - t) * (p1 - t)) / (double)(p1 * p1);
            tab[1][n1 - p1 + t] = (double)((t + 1) * (t + 1)) / (double)(p1 * p1);
        }
        for (int64_t t = 0; t < p2; ++t) {
            if (!neumann_on_top) tab[2][t] = (double)((p2 - t) * (p2 - t)) / (double)(p2 * p2);
            tab[3][n2 - p2 + t] = (double)((t + 1) * (t + 1)) / (double)(p2 * p2);
        }
        return;
    }
    for (int d = 0; d < 3; ++d) {
        const int64_t nd = n[d], p = pad[d];
        const double x0 = d < 2 ? -1.0 : 0.0, x1 = 1.0;
        std::vector<double> x(nd);
        // Julia range(a, stop=b, length=n): a + i*(b-a)/(n-1) (evaluated like LinRange: lerp)
        for (int64_t i = 0; i < nd; ++i) {
            const double t = nd > 1 ? (double)i / (double)(nd - 1) : 0.0;
            x[i] = (1.0 - t) * x0 + t * x1;
        }
        tab[d].assign(nd, 0.0);
        const bool left = !(d == 2 && neumann_on_top);
        if (left)
            for (int64_t i = 0; i < p; ++i) tab[d][i] += (x[i] - x[p - 1]) * (x[i] - x[p - 1]);
        for (int64_t i = nd - p; i < nd; ++i) tab[d][i] += (x[i] - x[nd - p]) * (x[i] - x[nd - p]);
        double mx = 0.0;
        for (int64_t i = 0; i < nd; ++i) mx = std::max(mx, tab[d][i]);
        for (int64_t i = 0; i < nd; ++i) tab[d][i] /= (mx + 1e-5);
    }
    tab[3].assign(1, 0.0);
}

struct SolverBase {
    virtual ~SolverBase() {}
    virtual void set_model(const double* m, const double* gamma, double wre, double wim) = 0;
    virtual void set_frequency_abl(double wre, double wim, double gamma0, const int64_t* n_global, const int64_t* pad, double amp) = 0;
    virtual void get_gamma(double* out) = 0;
    virtual double max_m() = 0;
    virtual void setup(const hh_mg_options& o) = 0;
    virtual void clear() = 0;
    virtual bool hierarchy_exists() const = 0;
    virtual void level_nodes(int level, int64_t* out) const = 0;
    virtual void get_level_stencil(int level, void* out) = 0;
    virtual void get_diagonal(int shifted, double shift, double* out) = 0;
    virtual void apply_device(const void* dX, void* dY, int64_t nrhs, int shifted, double shift, int transpose) = 0;
    virtual void cycle_device(const void* dB, void* dZ, int64_t nrhs) = 0;
    virtual int solve_device(const void* dB, void* dX, int64_t nrhs, const hh_solve_options& o, int32_t* iters,
                             double* relres) = 0;
    virtual int64_t max_rhs_per_batch(const hh_solve_options& o, double reusable_bytes = 0.0) = 0;
    virtual size_t elem_size() const = 0;
    virtual void scatter_point_sources(void* dB, const int64_t* idx0, const double* val, int64_t nrhs) = 0;
    // ---- mixed precision (HH_C64_MIXED): a ComplexF64 Krylov solver whose preconditioner is the multigrid cycle of a
    // ComplexF32 solver.  The outer solver is "krylov_only" (no hierarchy of its own) and calls prec_hook; the inner
    // one exposes its cycle on blocks in its internal (possibly padded) layout.
    virtual void precondition_internal(const void* b, void* z, int nrhs) = 0;
    virtual void ensure_cycle_memory(int nrhs) = 0;
    virtual int64_t internal_ld() const = 0;
    virtual int internal_pitch() const = 0;
    virtual double per_rhs_bytes(const hh_solve_options& o) = 0;
    virtual double held_bytes() const = 0;
    virtual double cycle_bytes_per_rhs() const = 0;
    std::function<void(const void*, void*, int)> prec_hook;
    bool krylov_only = false;
    // Slab decomposition (hh_slab.cuh): `pb` then describes the LOCAL grid of this slab (pb.n[2] = sgeo[0].nloc planes,
    // halo planes included), sgeo[l] the plane geometry of level l, and `slab` the collectives.  Caller-facing blocks
    // (B, X of hh_solve_device) hold the owned planes only.
    std::shared_ptr<SlabTransport> slab;
    std::vector<SlabLevel> sgeo;
    // High-order / spread operator (GetHelmholtzOperatorHO, src/GetHelmholtz.jl:54-72): when set, the fine level is a
    // stored 3^dim-point stencil built on the host (build_ho_stencil) from these Float64 copies of m and gamma, and
    // every fine-level operation runs through the stored-stencil kernels of the coarse levels.
    bool ho = false;
    double ho_beta[2] = {1.0, 1.0};
    std::vector<double> ho_m, ho_g;
    int64_t caller_N() const {
        return slab ? (int64_t)pb.n[0] * pb.n[1] * (sgeo[0].own1 - sgeo[0].own0) : pb.N();
    }
    Problem pb;
    int device = 0;
    cudaStream_t stream = 0;
    Profiler prof;
    int64_t launches = 0;
    double setup_seconds = 0, solve_seconds = 0;
    int64_t n_prec = 0;
};

template <typename T>
class Solver : public SolverBase {
   public:
    typedef cx<T> C;
    static constexpr double S = sizeof(C);  // bytes per complex entry
    static constexpr double CR = sizeof(T); // bytes per real coefficient
    static constexpr int FINE_MINB = 4;     // resident CTAs per SM the z-marching kernels are compiled for
    static constexpr int COARSE_MINB = 2;

    struct GmresMem {
        DevBuf<zc> H, cs, sn, s, hcol, y, scale;
        DevBuf<double> bnorm, err, d, acc;
        DevBuf<int> done, jdone, nprec;
        GmresState st{};
    };
    // workspace of a fixed-length GMRES nested inside the cycle
    struct SmallWs {
        DevBuf<C> v, z;  // (steps+1) basis vectors; `steps` preconditioned vectors when flexible
        int steps = 0;
        GmresMem g;
        DevBuf<zc> np;   // ||v~_{j+1}||^2 partials of the last update pass, consumed by the next step's scalar kernel
    };
    struct Level {
        int n[3] = {1, 1, 1};
        int p0 = 1;                 // row pitch of every array of this level (n[0], or n[0]+1: see pitch_of)
        int64_t N = 0;              // padded node count p0*n[1]*n[2] = leading dimension of the level's vectors
        int64_t Nlog = 0;           // logical node count
        int zb = 0, ze = 1;         // planes computed by the kernels (slab: the owned planes; else 0, n[2])
        int koff = 0, n2g = 1;      // slab: global index of local plane 0, global plane count (else 0, n[2])
        DevBuf<C> coef, dinv;       // l >= 1 (dinv also on l = 0 when Jac-GMRES is used)
        DevBuf<C> scoef;            // coef * diag(dinv) on the levels that run a Jacobi-preconditioned GMRES
        DevBuf<C> x, b, t;          // work vectors N x kcap (l >= 1); l = 0 owns only t
        SmallWs gs;                 // Jac-GMRES smoother / inexact coarsest solve (Jacobi-preconditioned)
        SmallWs ks;                 // K-cycle: 2 steps of FGMRES preconditioned by the recursive cycle
        C *px = nullptr, *pb = nullptr, *pt = nullptr;
    };

    Solver(const Problem& p, int dev) {
        pb = p;
        device = dev;
        // A/B switch for the 3-D stencil kernels: tma (TMA-staged, default) | zmarch (register z-marching) |
        // simple (baseline one thread per node)
        const char* e = getenv("HH_FINE_KERNEL");
        fine_kernel = FK_TMA;
        if (e && !strcmp(e, "zmarch")) fine_kernel = FK_ZMARCH;
        if (e && !strcmp(e, "simple")) fine_kernel = FK_SIMPLE;
        const char* f = getenv("HH_FUSE_FIRST");  // A/B switch of the fused cycle start (default on)
        fuse_first = !(f && f[0] == '0');
        // opt-in: measured slower than exchange-then-compute on 8 B200 (profiles/bench_r01_config5_slab_n8_513_ab.jsonl)
        const char* ho = getenv("HH_HALO_OVERLAP");
        halo_overlap = ho && ho[0] == '1';
        const char* ct = getenv("HH_COARSE_TILE");
        if (ct && !strcmp(ct, "16x8")) force_tile = 0;
        if (ct && !strcmp(ct, "alt")) force_tile = 1;
        const char* sc = getenv("HH_SCALED_GMRES");  // A/B switch of the one-pass (A D^-1) apply (default on)
        scaled_gmres = !(sc && sc[0] == '0');
        // opt-in: measured SLOWER than the two kernels it replaces (7.4 ms against 3.4 + 2.3 ms per cycle at 257^3 x 16 RHS,
        // profiles/r02_rejected_experiments.md): half the HBM bytes, but twice the shared-memory traffic behind two CTA
        // barriers per plane at one CTA per SM
        const char* fp = getenv("HH_FUSE_POST2");
        fuse_post2 = fp && fp[0] == '1';
        const char* tr = getenv("HH_TMA_RESTRICT");
        tma_restrict = !(tr && tr[0] == '0');
        const char* hs = getenv("HH_HALO_SPLIT");
        split_always = hs && hs[0] == '1';
        // A/B switches of the round-2 Krylov savings (default on): the last column of a GMRES cycle skips its update pass
        // (HH_SKIP_LAST_UPDATE), the fixed-length GMRES of a level runs one scalar kernel / one all-reduce per step and a
        // one-pass solution update (HH_SMALL_FUSED)
        const char* sl = getenv("HH_SKIP_LAST_UPDATE");
        skip_last_update = !(sl && sl[0] == '0');
        const char* sf = getenv("HH_SMALL_FUSED");
        small_fused = !(sf && sf[0] == '0');
        // cycles with one pre-smoothing sweep: x1 = dinv .* b is recomputed by the correction pass instead of being stored
        // (k_fine3d_tma_prob; default on, HH_FUSE_RECOMPUTE=0 is the A/B)
        const char* fr = getenv("HH_FUSE_RECOMPUTE");
        fuse_recompute = !(fr && fr[0] == '0');
        // k_fine3d_tma_prob keeps the in-plane interpolation of a coarse plane in registers (HH_PRO_CACHE=0: recompute it
        // for every fine plane); levels >= 1 swap their iterate / scratch buffers instead of a copy (HH_LEVEL_SWAP=0)
        const char* pc = getenv("HH_PRO_CACHE");
        pro_cache = !(pc && pc[0] == '0');
        const char* ls = getenv("HH_LEVEL_SWAP");
        level_swap = !(ls && ls[0] == '0');
        const char* ms = getenv("HH_MAPPED_STATE");
        mapped_state = !(ms && ms[0] == '0');
        const char* sq = getenv("HH_SCALAR_FAST");  // one warp per reduced quantity + lane-parallel Givens (default on)
        scalar_fast = !(sq && sq[0] == '0');
        use_pitch = sizeof(T) == 4 && pb.dim == 3 && fine_kernel == FK_TMA;
    }
    ~Solver() override {
        if (h_state) {
            cudaSetDevice(device);
            cudaFreeHost(h_state);
        }
        if (comm_stream) {
            cudaSetDevice(device);
            cudaStreamDestroy(comm_stream);
            cudaEventDestroy(ev_ready);
            cudaEventDestroy(ev_landed);
        }
    }

    size_t elem_size() const override { return sizeof(C); }

    // ------------------------------------------------------------------ model
    void set_model(const double* m, const double* gamma, double wre, double wim) override {
        HH_CUDA(cudaSetDevice(device));
        const int64_t N = pb.N();
        pb.w_re = wre;
        pb.w_im = wim;
        std::vector<T> hm(N), hg(N);
        for (int64_t i = 0; i < N; ++i) {
            hm[i] = (T)m[i];
            hg[i] = (T)gamma[i];
        }
        if (d_m.n != (size_t)N) {
            d_m.alloc(N);
            d_g.alloc(N);
        }
        HH_CUDA(cudaMemcpy(d_m.p, hm.data(), N * sizeof(T), cudaMemcpyHostToDevice));
        HH_CUDA(cudaMemcpy(d_g.p, hg.data(), N * sizeof(T), cudaMemcpyHostToDevice));
        refresh_pitched();  // pitched copies for the kernels that work on the internal (padded) vectors
        h_op_valid[0] = h_op_valid[1] = false;
        clear();
    }

    // copies of m / gamma in the row pitch of the internal vectors (only when that differs from the dense one)
    void refresh_pitched() {
        if (fine_sy() == pb.n[0]) return;
        const int64_t Np = fineN();
        if (d_mp.n != (size_t)Np) {
            d_mp.alloc(Np);
            d_gp.alloc(Np);
        }
        HH_CUDA(cudaMemsetAsync(d_mp.p, 0, Np * sizeof(T), stream));
        HH_CUDA(cudaMemsetAsync(d_gp.p, 0, Np * sizeof(T), stream));
        const int64_t rows = (int64_t)pb.n[1] * pb.n[2];
        k_repitch<T><<<dim3(592, 1), 256, 0, stream>>>(d_m.p, d_mp.p, pb.n[0], rows, pb.n[0], fine_sy(), 0, 0);
        k_repitch<T><<<dim3(592, 1), 256, 0, stream>>>(d_g.p, d_gp.p, pb.n[0], rows, pb.n[0], fine_sy(), 0, 0);
        HH_CUDA(cudaStreamSynchronize(stream));
    }
    // Frequency sweep on a resident model: omega replaced, gamma <- gamma0 + getABL(...) evaluated on the
    // device (k_gamma_abl); m stays where it is.  Invalidates the hierarchy like set_model.
    void set_frequency_abl(double wre, double wim, double gamma0, const int64_t* n_global, const int64_t* pad, double amp) override {
        HH_CUDA(cudaSetDevice(device));
        HH_REQUIRE(d_m.p != nullptr, HH_ERR_STATE, "no model set");
        HH_REQUIRE(!ho, HH_ERR_UNSUPPORTED, "hh_set_frequency_abl: not available with the high-order operator (it keeps host copies of gamma)");
        std::vector<double> tab[4];
        abl_tables(pb.dim, n_global, pb.neumann_top, pad, tab);
        DevBuf<double> dt[4];
        for (int q = 0; q < 4; ++q) {
            dt[q].alloc(std::max<size_t>(tab[q].size(), 1));
            if (!tab[q].empty()) HH_CUDA(cudaMemcpyAsync(dt[q].p, tab[q].data(), tab[q].size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        }
        dim3 g, blk;
        grid3(pb.n, pb.dim, g, blk);
        const int koff = slab ? sgeo[0].koff : 0;
        launch(T_SETUP, 0, [&] {
            if (pb.dim == 3) k_gamma_abl<T, 3><<<g, blk, 0, stream>>>(d_g.p, pb.n[0], pb.n[1], pb.n[2], koff, gamma0, amp, dt[0].p, dt[1].p, dt[2].p, dt[3].p);
            else k_gamma_abl<T, 2><<<g, blk, 0, stream>>>(d_g.p, pb.n[0], pb.n[1], 1, 0, gamma0, amp, dt[0].p, dt[1].p, dt[2].p, dt[3].p);
        });
        HH_CUDA(cudaStreamSynchronize(stream));
        pb.w_re = wre;
        pb.w_im = wim;
        refresh_pitched();
        h_op_valid[0] = h_op_valid[1] = false;
        clear();
    }
    void get_gamma(double* out) override {
        HH_CUDA(cudaSetDevice(device));
        const int64_t N = pb.N();
        std::vector<T> hg(N);
        HH_CUDA(cudaMemcpy(hg.data(), d_g.p, N * sizeof(T), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < N; ++i) out[i] = (double)hg[i];
    }
    double max_m() override {  // over the planes this solver holds
        HH_CUDA(cudaSetDevice(device));
        const int nb = 296;
        DevBuf<double> part;
        part.alloc(nb);
        launch(T_SETUP, 0, [&] { k_max_partial<T><<<nb, 256, 0, stream>>>(d_m.p, pb.N(), part.p); });
        std::vector<double> h(nb);
        HH_CUDA(cudaMemcpyAsync(h.data(), part.p, nb * sizeof(double), cudaMemcpyDeviceToHost, stream));
        HH_CUDA(cudaStreamSynchronize(stream));
        return *std::max_element(h.begin(), h.end());
    }

    // Row pitch of the internal fine-level arrays.  TMA needs 16-byte multiples for every global stride: always
    // true for ComplexF64; ComplexF32 rows of an odd node count are padded by one (never touched) ghost node.
    int pitch_of(int n0) const { return (use_pitch && (n0 & 1)) ? n0 + 1 : n0; }
    int fine_sy() const { return pitch_of(pb.n[0]); }
    int64_t fineN() const { return (int64_t)fine_sy() * pb.n[1] * pb.n[2]; }

    // pitched = true: the operator on the solver's internal vectors; false: on dense caller arrays (hh_apply)
    FineOp<T> fine_op(double shift, int adj, bool pitched = false) const {
        FineOp<T> op;
        const bool pp = pitched && fine_sy() != pb.n[0];
        op.m = pp ? d_mp.p : d_m.p;
        op.g = pp ? d_gp.p : d_g.p;
        op.sy = pp ? fine_sy() : pb.n[0];
        const double wr = pb.w_re, wi = pb.w_im;
        op.a = (T)(wr * wr - wi * wi);
        op.b = (T)(2.0 * wr * wi);
        op.inv_wr = (T)(1.0 / wr);
        op.shift_w2 = (T)(shift * wr * wr);
        for (int d = 0; d < 3; ++d) {
            op.ih2[d] = d < pb.dim ? (T)(1.0 / (pb.h[d] * pb.h[d])) : T(0);
            // getSommerfeldBC is always called with its default orderNeumannBC = 2 (GetHelmholtz.jl:45)
            op.somm[d] = (pb.sommerfeld && d < pb.dim) ? (T)(wr * 2.0 / pb.h[d]) : T(0);
            op.n[d] = pb.n[d];
        }
        op.BC = (T)(pb.order_bc == 2 ? 2.0 : 1.0);
        op.neumann_top = pb.neumann_top;
        op.adj = adj;
        op.cdiag = nullptr;
        op.dinv = nullptr;
        op.koff = slab ? sgeo[0].koff : 0;
        op.n2g = slab ? sgeo[0].n2g : pb.n[2];
        op.zb = slab ? sgeo[0].zb : 0;
        op.ze = slab ? sgeo[0].ze : pb.n[2];
        return op;
    }
    // grid of the one-thread-per-node kernels that compute planes zb <= k < ze only
    static void grid3z(const int n[3], int dim, int zb, int ze, dim3& g, dim3& b) {
        int nn[3] = {n[0], n[1], dim == 3 ? ze - zb : 1};
        grid3(nn, dim, g, b);
    }

    // ------------------------------------------------------------------ slab collectives
    // fill the halo planes (zb-1 and ze) of every vector of the block `v` on level l from the neighbouring slabs
    // (sides: HALO_LOWER = plane zb-1 only, as restriction reads it; HALO_UPPER = plane ze only, as interpolation does)
    enum { HALO_LOWER = 1, HALO_UPPER = 2, HALO_BOTH = 3 };
    void halo_exchange(int l, const C* v, int nvec, int64_t ld = 0, int sides = HALO_BOTH, cudaStream_t on = nullptr) {
        if (!slab || slab->nranks == 1) return;
        const Level& L = levels[l];
        const size_t plane = (size_t)L.p0 * L.n[1] * sizeof(C);
        const int nmsg = ((sides & HALO_LOWER) ? (slab->rank > 0) + (slab->rank < slab->nranks - 1) : 0) +
                         ((sides & HALO_UPPER) ? (slab->rank > 0) + (slab->rank < slab->nranks - 1) : 0);
        launch(T_HALO, (double)plane * nvec * nmsg, [&] {
            slab->exchange(on ? on : stream, device, (char*)const_cast<C*>(v), (size_t)(ld ? ld : L.N) * sizeof(C), nvec, plane,
                           (size_t)L.zb * plane, (size_t)(L.zb - 1) * plane, (size_t)(L.ze - 1) * plane, (size_t)L.ze * plane,
                           (sides & HALO_LOWER) != 0, (sides & HALO_UPPER) != 0);
        }, on);
    }
    // ---- halo exchange overlapped with the planes that do not need it ------------------------------------------
    // A stencil-type kernel on planes [zb,ze) reads halo planes only for its first and last plane.  With a transport
    // that is pure stream work (NCCL) the exchange can run on a second stream while the interior planes are computed;
    // the one or two boundary planes follow once it has landed (HH_HALO_OVERLAP=1; off by default: on 8 B200 the NCCL
    // copy kernels competing with the stencil kernels plus the two extra launches cost more than the hidden latency).  NCCL operations still never run concurrently with each
    // other: the next one is issued after an event that follows everything enqueued so far.
    struct HaloReq {
        int level;
        const C* v;
        int64_t ld;
    };
    bool overlap_ok(int zb, int ze) const {
        return slab && slab->nranks > 1 && slab->stream_ordered() && halo_overlap && (ze - zb) >= 4;
    }
    // calls run(z0, z1) for sub-ranges that together cover [zb,ze) exactly once
    template <class F>
    void with_halos(std::initializer_list<HaloReq> reqs, int nvec, int zb, int ze, F&& run) {
        if (!overlap_ok(zb, ze)) {
            for (const HaloReq& r : reqs) halo_exchange(r.level, r.v, nvec, r.ld);
            if (split_always && slab && slab->nranks > 1 && (ze - zb) >= 4) {
                // HH_HALO_SPLIT=1: the launch pattern of the overlapped path with any transport (single-GPU tests)
                const int lo = slab->rank > 0 ? 1 : 0, hi = slab->rank < slab->nranks - 1 ? 1 : 0;
                run(zb + lo, ze - hi);
                if (lo) run(zb, zb + 1);
                if (hi) run(ze - 1, ze);
                return;
            }
            run(zb, ze);
            return;
        }
        if (!comm_stream) {
            HH_CUDA(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
            HH_CUDA(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
            HH_CUDA(cudaEventCreateWithFlags(&ev_landed, cudaEventDisableTiming));
            prof.aux = comm_stream;
        }
        const int lo = slab->rank > 0 ? 1 : 0, hi = slab->rank < slab->nranks - 1 ? 1 : 0;
        HH_CUDA(cudaEventRecord(ev_ready, stream));  // the producers of the vectors (and every earlier NCCL call)
        HH_CUDA(cudaStreamWaitEvent(comm_stream, ev_ready, 0));
        for (const HaloReq& r : reqs) halo_exchange(r.level, r.v, nvec, r.ld, HALO_BOTH, comm_stream);
        HH_CUDA(cudaEventRecord(ev_landed, comm_stream));
        run(zb + lo, ze - hi);
        HH_CUDA(cudaStreamWaitEvent(stream, ev_landed, 0));
        if (lo) run(zb, zb + 1);
        if (hi) run(ze - 1, ze);
    }
    int zbeg(const Level& L) const { return rz_b >= 0 ? rz_b : L.zb; }
    int zend(const Level& L) const { return rz_e >= 0 ? rz_e : L.ze; }
    // Sum the nblk partials of each of the nq quantities, all-reduce over the slabs and leave the result in `partial`
    // in the layout of nblk = 1 (which is returned).  No-op without slabs.
    // The scalar kernels that consume the sums read them from `red_out` (the partial buffer itself, or the all-reduced
    // block: no copy back).
    int slab_reduce(zc* partial, int nq, int nblk) {
        red_out = partial;
        if (!slab || slab->nranks == 1) return nblk;
        if (d_red.n < (size_t)nq) d_red.alloc((size_t)std::max(nq, 4096));
        launch(T_SCALAR, 0, [&] { k_sum_partials<<<(nq + 7) / 8, 256, 0, stream>>>(partial, nq, nblk, d_red.p); });
        launch(T_ALLREDUCE, 16.0 * nq, [&] { slab->allreduce(stream, (double*)d_red.p, 2 * nq, false); });
        red_out = d_red.p;
        return 1;
    }
    // range of a level's vectors that the Krylov / reduction kernels work on (slab: the owned planes)
    struct Span {
        int64_t off, len, ld;
    };
    Span span(int l) const {
        const Level& L = levels[l];
        if (!slab) return Span{0, L.N, L.N};
        const int64_t plane = (int64_t)L.p0 * L.n[1];
        return Span{(int64_t)L.zb * plane, (int64_t)(L.ze - L.zb) * plane, L.N};
    }
    // diagonal arrays for the TMA-staged kernels (3-D only), in the layout of `op`
    void precompute_diag(FineOp<T>& op, DevBuf<C>& cd, DevBuf<C>* dv, T damp) {
        const int64_t Np = (int64_t)op.sy * pb.n[1] * pb.n[2];
        if (!tma_ok(op, Np)) return;
        cd.alloc(Np);
        HH_CUDA(cudaMemsetAsync(cd.p, 0, Np * sizeof(C), stream));
        if (dv) {
            dv->alloc(Np);
            HH_CUDA(cudaMemsetAsync(dv->p, 0, Np * sizeof(C), stream));
        }
        dim3 g, blk;
        grid3(pb.n, pb.dim, g, blk);
        FineOp<T> o = op;
        launch(T_SETUP, 0, [&] { k_fine_precompute<T, 3><<<g, blk, 0, stream>>>(o, cd.p, dv ? dv->p : nullptr, damp); });
        op.cdiag = cd.p;
        op.dinv = dv ? dv->p : nullptr;
    }
    // the un-shifted operator H of the outer Krylov method (GetHelmholtz.jl:85-95), with cached diagonal
    const FineOp<T>& krylov_op(int transpose) {
        const int t = transpose ? 1 : 0;
        if (!h_op_valid[t]) {
            h_op[t] = fine_op(0.0, t, true);
            precompute_diag(h_op[t], h_cdiag[t], nullptr, T(0));
            h_op_valid[t] = true;
        }
        return h_op[t];
    }

    // ------------------------------------------------------------------ launch plumbing
    template <class F>
    void launch(int tag, double bytes, F&& f, cudaStream_t on = nullptr) {
        const int64_t sub = (int64_t)bytes;
        ++launches;
        if (prof.on) {
            cudaEvent_t a = prof.get(), b = prof.get();
            HH_CUDA(cudaEventRecord(a, on ? on : stream));
            f();
            HH_CUDA(cudaEventRecord(b, on ? on : stream));
            prof.recs.push_back({tag, a, b, bytes, sub});
            if (prof.recs.size() >= 8192) prof.flush(stream);
        } else {
            f();
        }
#ifdef HH_DEBUG_SYNC
        HH_CUDA(cudaStreamSynchronize(stream));
#endif
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            throw Error(HH_ERR_CUDA, std::string("kernel launch (") + kTagNames[tag] + "): " + cudaGetErrorString(e));
    }

    static void grid3(const int n[3], int dim, dim3& g, dim3& b) {
        if (dim == 3) {
            b = dim3(32, 4, 2);
            g = dim3((n[0] + 31) / 32, (n[1] + 3) / 4, (n[2] + 1) / 2);
        } else {
            b = dim3(32, 8, 1);
            g = dim3((n[0] + 31) / 32, (n[1] + 7) / 8, 1);
        }
    }
    // 2-D grids: how many slices of the RHS block the z dimension of a one-thread-per-node grid strides over
    static unsigned rhs_slices(const dim3& g, int nrhs) {
        const int64_t ctas = (int64_t)g.x * g.y;
        return (unsigned)std::max<int64_t>(1, std::min<int64_t>(nrhs, (4 * 148 + ctas - 1) / ctas));
    }
    static int vec_blocks(int64_t N, int nrhs) {
        int64_t nb = (N + 2047) / 2048;  // 8 elements per thread
        int64_t cap = std::max<int64_t>(1, (148 * 8 + nrhs - 1) / nrhs);
        return (int)std::max<int64_t>(1, std::min<int64_t>(nb, cap));
    }

    // ------------------------------------------------------------------ operator kernels
    // out = op(x) with MODE epilogue on the fine level
    void fine_stencil(int mode, const FineOp<T>& op0, const C* x, const C* b, C* out, int64_t ld, int nrhs, T damp) {
        with_halos({{0, x, ld}}, nrhs, op0.zb, op0.ze, [&](int z0, int z1) {
            FineOp<T> o = op0;
            o.zb = z0;
            o.ze = z1;
            fine_stencil_range(mode, o, x, b, out, ld, nrhs, damp);
        });
    }
    void fine_stencil_range(int mode, const FineOp<T>& op, const C* x, const C* b, C* out, int64_t ld, int nrhs, T damp) {
        dim3 g, blk;
        grid3z(pb.n, pb.dim, op.zb, op.ze, g, blk);
        const double N = pb.dim == 3 ? (double)pb.n[0] * pb.n[1] * (op.ze - op.zb) : (double)pb.N();
        const double coefb = 2.0 * CR * N;
        double bytes;
        int tag;
        if (mode == MODE_APPLY) bytes = 2 * S * N * nrhs + coefb, tag = T_FINE_APPLY;
        else if (mode == MODE_RESID) bytes = 3 * S * N * nrhs + coefb, tag = T_FINE_RESID;
        else bytes = 3 * S * N * nrhs + coefb, tag = T_FINE_JACOBI;
        if (op.cdiag != nullptr && (mode != MODE_JACOBI || op.dinv != nullptr) && tma_ok(op, ld) &&
            ((uintptr_t)x % 16 == 0) && (mode == MODE_APPLY || (uintptr_t)b % 16 == 0)) {
            launch(tag, bytes + S * N * (mode == MODE_JACOBI ? 2.0 : 1.0) - coefb,
                   [&] { tma3d_dispatch(mode, op, x, b, out, ld, nrhs); });
            return;
        }
        if (pb.dim == 3 && fine_kernel != FK_SIMPLE) {
            launch(tag, bytes, [&] { fine3d_dispatch(mode, op, x, b, out, ld, nrhs, damp); });
            return;
        }
        if (tma2d_ok(op.sy, op.n[1], ld, x, b, mode)) {
            launch(tag, bytes, [&] { stencil2d_dispatch<false>(mode, op, CoarseOp<T>{}, op.n, op.sy, x, b, out, ld, nrhs, damp); });
            return;
        }
        launch(tag, bytes, [&] {
            if (pb.dim == 3) {
                if (mode == MODE_APPLY) k_fine_stencil<T, 3, MODE_APPLY, 2><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs, damp);
                else if (mode == MODE_RESID) k_fine_stencil<T, 3, MODE_RESID, 2><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs, damp);
                else k_fine_stencil<T, 3, MODE_JACOBI, 2><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs, damp);
            } else {
                if (mode == MODE_APPLY) k_fine_stencil<T, 2, MODE_APPLY, 2><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs, damp);
                else if (mode == MODE_RESID) k_fine_stencil<T, 2, MODE_RESID, 2><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs, damp);
                else k_fine_stencil<T, 2, MODE_JACOBI, 2><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs, damp);
            }
        });
    }
    // choose the z-chunking so that the grid has a few CTAs per SM
    static void zchunks(int n2, int tiles, int groups, int pref, int& zchunk, int& nzc) {
        int want = (592 + tiles * groups - 1) / (tiles * groups);
        nzc = std::max((n2 + pref - 1) / pref, want);
        nzc = std::max(1, std::min(nzc, n2));
        zchunk = (n2 + nzc - 1) / nzc;
        nzc = (n2 + zchunk - 1) / zchunk;
    }
    template <int MODE, int KB>
    void fine3d_launch(const FineOp<T>& op, const C* x, const C* b, C* out, int64_t ld, int nrhs, T damp) {
        constexpr int TY = 8;
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (pb.n[0] + 31) / 32, ty = (pb.n[1] + TY - 1) / TY;
        int zchunk, nzc;
        zchunks(op.ze - op.zb, tx * ty, groups, 32, zchunk, nzc);
        dim3 g(tx * groups, ty, nzc), blk(32, TY, 1);
        k_fine3d_zmarch<T, MODE, KB, TY, FINE_MINB><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs, damp, zchunk, groups);
    }
    template <int MODE>
    void fine3d_mode(const FineOp<T>& op, const C* x, const C* b, C* out, int64_t ld, int nrhs, T damp) {
        const int pref = sizeof(T) == 8 ? 2 : 4;
        const int kb = std::min(pref, nrhs >= 4 ? 4 : (nrhs >= 2 ? 2 : 1));
        if (kb == 4) fine3d_launch<MODE, 4>(op, x, b, out, ld, nrhs, damp);
        else if (kb == 2) fine3d_launch<MODE, 2>(op, x, b, out, ld, nrhs, damp);
        else fine3d_launch<MODE, 1>(op, x, b, out, ld, nrhs, damp);
    }
    void fine3d_dispatch(int mode, const FineOp<T>& op, const C* x, const C* b, C* out, int64_t ld, int nrhs, T damp) {
        if (mode == MODE_APPLY) fine3d_mode<MODE_APPLY>(op, x, b, out, ld, nrhs, damp);
        else if (mode == MODE_RESID) fine3d_mode<MODE_RESID>(op, x, b, out, ld, nrhs, damp);
        else fine3d_mode<MODE_JACOBI>(op, x, b, out, ld, nrhs, damp);
    }
    // ---- TMA tensor maps (driver entry point fetched through the runtime; libcuda is not linked) ----
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn encode_fn() {
        static EncodeTiledFn fn = nullptr;
        if (!fn) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            HH_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
            HH_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, HH_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
            fn = (EncodeTiledFn)p;
        }
        return fn;
    }
    // TMA needs 16-byte aligned global strides: always true for ComplexF64, for ComplexF32 only on even grids
    bool tma_ok(const FineOp<T>& op, int64_t ld) const {
        if (pb.dim != 3 || fine_kernel != FK_TMA) return false;
        const int64_t es = 2 * sizeof(T);
        return (es * op.sy) % 16 == 0 && (es * op.sy * pb.n[1]) % 16 == 0 && (es * ld) % 16 == 0;
    }
    template <int MODE, int KB>
    void tma3d_launch(const FineOp<T>& op, const C* x, const C* b, C* out, int64_t ld, int nrhs) {
        typedef FineTmaCfg<T, MODE, KB> Cfg;
        constexpr int NS = 4;
        constexpr size_t smem = (size_t)NS * Cfg::STAGE_BYTES + NS * sizeof(uint64_t);
        static bool attr_set = false;  // per instantiation
        if (!attr_set) {
            HH_CUDA(cudaFuncSetAttribute(k_fine3d_tma<T, MODE, KB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (pb.n[0] + Cfg::TX - 1) / Cfg::TX, ty = (pb.n[1] + Cfg::TY - 1) / Cfg::TY;
        int zchunk, nzc;
        zchunks(op.ze - op.zb, tx * ty, groups, 64, zchunk, nzc);
        dim3 g(tx * groups, ty, nzc);
        // the RHS extent of the x / b maps is the true nrhs so that surplus slots of the last group are zero-filled
        TmaDesc tx_ = make_tmap_n(x, op.sy, ld, Cfg::PX, Cfg::TY + 2, KB, nrhs);
        TmaDesc tb_ = (MODE != MODE_APPLY) ? make_tmap_n(b, op.sy, ld, Cfg::TX, Cfg::TY, KB, nrhs) : tx_;
        TmaDesc tc_ = make_tmap_n(op.cdiag, op.sy, ld, Cfg::TX, Cfg::TY, 1, 0);
        TmaDesc td_ = (MODE == MODE_JACOBI) ? make_tmap_n(op.dinv, op.sy, ld, Cfg::TX, Cfg::TY, 1, 0) : tc_;
        k_fine3d_tma<T, MODE, KB, NS><<<g, 256, smem, stream>>>(op, tx_, tb_, tc_, td_, x, out, ld, nrhs, zchunk, groups);
    }
    TmaDesc make_tmap_n(const void* base, int sy, int64_t ld, int bx, int by, int kb, int nrhs) const {
        return make_tmap_g(base, pb.n, sy, ld, bx, by, kb, nrhs);
    }
    // tensor over a complex block on an n[0] x n[1] x n[2] grid viewed as reals: dims (2*n0, n1, n2[, nvec]),
    // box (2*bx, by, 1[, kb]); nvec == 0 -> rank 3
    static TmaDesc make_tmap_g(const void* base, const int* n, int sy, int64_t ld, int bx, int by, int kb, int nrhs) {
        static_assert(sizeof(TmaDesc) == sizeof(CUtensorMap), "CUtensorMap is 128 bytes");
        TmaDesc d;
        const cuuint64_t es = sizeof(T);
        const int rank = nrhs > 0 ? 4 : 3;
        cuuint64_t dims[4] = {(cuuint64_t)2 * n[0], (cuuint64_t)n[1], (cuuint64_t)n[2], (cuuint64_t)std::max(nrhs, 1)};
        cuuint64_t strides[3] = {2 * es * sy, 2 * es * (cuuint64_t)sy * n[1], 2 * es * (cuuint64_t)ld};
        cuuint32_t box[4] = {(cuuint32_t)(2 * bx), (cuuint32_t)by, 1, (cuuint32_t)kb};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode_fn()((CUtensorMap*)&d, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                 (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        HH_REQUIRE(r == CUDA_SUCCESS, HH_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return d;
    }
    // fused first sweep from zero + (residual | second sweep); see k_fine3d_tma_first
    template <int SECOND, int KB>
    void tma3d_first_launch(const FineOp<T>& op, const C* b, C* out, C* out2, int64_t ld, int nrhs) {
        typedef FineFirstCfg<T, KB> Cfg;
        constexpr int NS = 4;
        constexpr size_t smem = (size_t)NS * Cfg::STAGE_BYTES + NS * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            HH_CUDA(cudaFuncSetAttribute(k_fine3d_tma_first<T, SECOND, KB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (pb.n[0] + Cfg::TX - 1) / Cfg::TX, ty = (pb.n[1] + Cfg::TY - 1) / Cfg::TY;
        int zchunk, nzc;
        zchunks(op.ze - op.zb, tx * ty, groups, 64, zchunk, nzc);
        dim3 g(tx * groups, ty, nzc);
        TmaDesc mb = make_tmap_n(b, op.sy, ld, Cfg::PX, Cfg::TY + 2, KB, nrhs);
        TmaDesc md = make_tmap_n(op.dinv, op.sy, ld, Cfg::PX, Cfg::TY + 2, 1, 0);
        TmaDesc mc = make_tmap_n(op.cdiag, op.sy, ld, Cfg::TX, Cfg::TY, 1, 0);
        k_fine3d_tma_first<T, SECOND, KB, NS><<<g, 256, smem, stream>>>(op, mb, md, mc, b, out, out2, ld, nrhs, zchunk, groups);
    }
    bool can_fuse_first(const FineOp<T>& op, const C* b, int64_t ld) const {
        return !ho && op.cdiag != nullptr && op.dinv != nullptr && tma_ok(op, ld) && ((uintptr_t)b % 16 == 0) && fuse_first;
    }
    // second == 0: out = x1, out2 = b - A x1;  second == 1: out = x2
    void fine_first(int second, const FineOp<T>& op0, const C* b, C* out, C* out2, int64_t ld, int nrhs) {
        with_halos({{0, b, ld}}, nrhs, op0.zb, op0.ze, [&](int z0, int z1) {
            FineOp<T> o = op0;
            o.zb = z0;
            o.ze = z1;
            fine_first_range(second, o, b, out, out2, ld, nrhs);
        });
    }
    void fine_first_range(int second, const FineOp<T>& op, const C* b, C* out, C* out2, int64_t ld, int nrhs) {
        const double N = (double)pb.n[0] * pb.n[1] * (op.ze - op.zb);
        const double bytes = ((second == 0 && out != nullptr) ? 3.0 : 2.0) * S * N * nrhs + 2.0 * S * N;
        launch(second == 0 ? T_FINE_FIRST_RESID : T_FINE_FIRST_JACOBI, bytes, [&] {
            if (second == 0) {
                if (nrhs >= 2) tma3d_first_launch<0, 2>(op, b, out, out2, ld, nrhs);
                else tma3d_first_launch<0, 1>(op, b, out, out2, ld, nrhs);
            } else {
                if (nrhs >= 2) tma3d_first_launch<1, 2>(op, b, out, out2, ld, nrhs);
                else tma3d_first_launch<1, 1>(op, b, out, out2, ld, nrhs);
            }
        });
    }
    // fused coarse-grid correction + first post-smoothing sweep; see k_fine3d_tma_pro
    template <int KB>
    void tma3d_pro_launch(const FineOp<T>& op, const C* x, const C* b, const Level& Cc, const C* xc, C* out, int64_t ld, int nrhs) {
        typedef FineProCfg<T, KB> Cfg;
        constexpr int NS = 3;
        constexpr size_t smem = (size_t)NS * Cfg::STAGE_BYTES + NS * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            HH_CUDA(cudaFuncSetAttribute(k_fine3d_tma_pro<T, KB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (pb.n[0] + Cfg::TX - 1) / Cfg::TX, ty = (pb.n[1] + Cfg::TY - 1) / Cfg::TY;
        int zchunk, nzc;
        zchunks(op.ze - op.zb, tx * ty, groups, 64, zchunk, nzc);
        dim3 g(tx * groups, ty, nzc);
        TmaDesc mx = make_tmap_n(x, op.sy, ld, Cfg::PX, Cfg::TY + 2, KB, nrhs);
        TmaDesc mb = make_tmap_n(b, op.sy, ld, Cfg::TX, Cfg::TY, KB, nrhs);
        TmaDesc mc = make_tmap_n(op.cdiag, op.sy, ld, Cfg::TX, Cfg::TY, 1, 0);
        TmaDesc md = make_tmap_n(op.dinv, op.sy, ld, Cfg::TX, Cfg::TY, 1, 0);
        TmaDesc mxc = make_tmap_g(xc, Cc.n, Cc.p0, Cc.N, Cfg::CTX, Cfg::CTY, KB, nrhs);
        k_fine3d_tma_pro<T, KB, NS><<<g, 256, smem, stream>>>(op, mx, mb, mc, md, mxc, x, xc, out, ld, Cc.N, Cc.p0, Cc.n[1], nrhs, zchunk, groups);
    }
    bool can_fuse_prolong(const FineOp<T>& op, const C* x, const C* b, const Level& Cc, const C* xc, int64_t ld) const {
        return !ho && fuse_first && op.cdiag != nullptr && op.dinv != nullptr && tma_ok(op, ld) && tma_ok_level(Cc) &&
               ((uintptr_t)x % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)xc % 16 == 0);
    }
    void fine_prolong_jacobi(const FineOp<T>& op0, const C* x, const C* b, const Level& Cc, const C* xc, C* out, int64_t ld, int nrhs) {
        with_halos({{0, x, ld}, {1, xc, 0}}, nrhs, op0.zb, op0.ze, [&](int z0, int z1) {
            FineOp<T> o = op0;
            o.zb = z0;
            o.ze = z1;
            fine_prolong_jacobi_range(o, x, b, Cc, xc, out, ld, nrhs);
        });
    }
    void fine_prolong_jacobi_range(const FineOp<T>& op, const C* x, const C* b, const Level& Cc, const C* xc, C* out, int64_t ld, int nrhs) {
        const double N = (double)pb.n[0] * pb.n[1] * (op.ze - op.zb);
        launch(T_FINE_PROLONG_JACOBI, (3.0 * N + 0.125 * N) * S * nrhs + 2.0 * S * N, [&] {
            if (nrhs >= 2) tma3d_pro_launch<2>(op, x, b, Cc, xc, out, ld, nrhs);
            else tma3d_pro_launch<1>(op, x, b, Cc, xc, out, ld, nrhs);
        });
    }
    // the same with x = dinv .* b recomputed on the fly (cycle start with one pre-smoothing sweep); see k_fine3d_tma_prob
    template <int KB>
    void tma3d_prob_launch(const FineOp<T>& op, const C* b, const Level& Cc, const C* xc, C* out, int64_t ld, int nrhs) {
        typedef FineProBCfg<T, KB> Cfg;
        constexpr int NS = 4;
        constexpr size_t smem = (size_t)NS * Cfg::STAGE_BYTES + NS * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            HH_CUDA(cudaFuncSetAttribute(k_fine3d_tma_prob<T, KB, NS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            HH_CUDA(cudaFuncSetAttribute(k_fine3d_tma_prob<T, KB, NS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (pb.n[0] + Cfg::TX - 1) / Cfg::TX, ty = (pb.n[1] + Cfg::TY - 1) / Cfg::TY;
        int zchunk, nzc;
        zchunks(op.ze - op.zb, tx * ty, groups, 64, zchunk, nzc);
        dim3 g(tx * groups, ty, nzc);
        TmaDesc mb = make_tmap_n(b, op.sy, ld, Cfg::PX, Cfg::TY + 2, KB, nrhs);
        TmaDesc md = make_tmap_n(op.dinv, op.sy, ld, Cfg::PX, Cfg::TY + 2, 1, 0);
        TmaDesc mc = make_tmap_n(op.cdiag, op.sy, ld, Cfg::TX, Cfg::TY, 1, 0);
        TmaDesc mxc = make_tmap_g(xc, Cc.n, Cc.p0, Cc.N, Cfg::CTX, Cfg::CTY, KB, nrhs);
        if (pro_cache)
            k_fine3d_tma_prob<T, KB, NS, true><<<g, 256, smem, stream>>>(op, mb, md, mc, mxc, b, xc, out, ld, Cc.N, Cc.p0, Cc.n[1], nrhs, zchunk, groups);
        else
            k_fine3d_tma_prob<T, KB, NS, false><<<g, 256, smem, stream>>>(op, mb, md, mc, mxc, b, xc, out, ld, Cc.N, Cc.p0, Cc.n[1], nrhs, zchunk, groups);
    }
    bool can_fuse_prolong_b(const FineOp<T>& op, const C* b, const Level& Cc, const C* xc, int64_t ld) const {
        return fuse_recompute && can_fuse_first(op, b, ld) && tma_ok_level(Cc) && ((uintptr_t)xc % 16 == 0);
    }
    // b's halo planes were exchanged for the cycle start (fine_first) and b has not changed since: only xc travels
    void fine_prolong_jacobi_b(const FineOp<T>& op0, const C* b, const Level& Cc, const C* xc, C* out, int64_t ld, int nrhs) {
        with_halos({{1, xc, 0}}, nrhs, op0.zb, op0.ze, [&](int z0, int z1) {
            FineOp<T> o = op0;
            o.zb = z0;
            o.ze = z1;
            const double N = (double)pb.n[0] * pb.n[1] * (o.ze - o.zb);
            launch(T_FINE_PROLONG_JACOBI, (2.0 * N + 0.125 * N) * S * nrhs + 2.0 * S * N, [&] {
                if (nrhs >= 2) tma3d_prob_launch<2>(o, b, Cc, xc, out, ld, nrhs);
                else tma3d_prob_launch<1>(o, b, Cc, xc, out, ld, nrhs);
            });
        });
    }
    // fused coarse-grid correction + two post-smoothing sweeps; see k_fine3d_tma_pro2 (whole grids only: two-node halo)
    template <int KB>
    void tma3d_pro2_launch(const FineOp<T>& op, const C* x, const C* b, const Level& Cc, const C* xc, C* out, int64_t ld, int nrhs) {
        typedef FinePro2Cfg<T, KB> Cfg;
        constexpr size_t smem = (size_t)Cfg::NS * Cfg::STAGE_BYTES + 2 * Cfg::X1_SLOT + Cfg::NS * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            HH_CUDA(cudaFuncSetAttribute(k_fine3d_tma_pro2<T, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (pb.n[0] + Cfg::TX - 1) / Cfg::TX, ty = (pb.n[1] + Cfg::TY - 1) / Cfg::TY;
        // one CTA per SM, four extra planes per chunk: long, equal chunks in whole waves
        int zchunk, nzc;
        {
            const int nz = op.ze - op.zb;
            const int64_t per = (int64_t)tx * ty * groups;
            double best = 1e300;
            zchunk = nz;
            nzc = 1;
            for (int c = 1; c <= nz; ++c) {
                const int zc = (nz + c - 1) / c, cc = (nz + zc - 1) / zc;
                if (cc != c) continue;
                const double cost = std::ceil((double)per * cc / 148.0) * (zc + 4.0);
                if (cost < best) best = cost, zchunk = zc, nzc = cc;
                if (zc <= 8) break;
            }
        }
        dim3 g(tx * groups, ty, nzc);
        TmaDesc mx = make_tmap_n(x, op.sy, ld, Cfg::PX2, Cfg::PY2, KB, nrhs);
        TmaDesc mb = make_tmap_n(b, op.sy, ld, Cfg::PX1, Cfg::PY1, KB, nrhs);
        TmaDesc mc = make_tmap_n(op.cdiag, op.sy, ld, Cfg::PX1, Cfg::PY1, 1, 0);
        TmaDesc md = make_tmap_n(op.dinv, op.sy, ld, Cfg::PX1, Cfg::PY1, 1, 0);
        TmaDesc mxc = make_tmap_g(xc, Cc.n, Cc.p0, Cc.N, Cfg::CTX, Cfg::CTY, KB, nrhs);
        k_fine3d_tma_pro2<T, KB><<<g, 256, smem, stream>>>(op, mx, mb, mc, md, mxc, out, ld, nrhs, zchunk, groups);
    }
    bool can_fuse_post2(const FineOp<T>& op, const C* x, const C* b, const Level& Cc, const C* xc, int64_t ld) const {
        return fuse_post2 && !slab && can_fuse_prolong(op, x, b, Cc, xc, ld);
    }
    void fine_prolong_jacobi2(const FineOp<T>& op, const C* x, const C* b, const Level& Cc, const C* xc, C* out, int64_t ld, int nrhs) {
        const double N = (double)pb.N();
        launch(T_FINE_PROLONG_JACOBI2, (3.0 * N + 0.125 * N) * S * nrhs + 2.0 * S * N, [&] {
            if (nrhs >= 2) tma3d_pro2_launch<2>(op, x, b, Cc, xc, out, ld, nrhs);
            else tma3d_pro2_launch<1>(op, x, b, Cc, xc, out, ld, nrhs);
        });
    }
    template <int MODE>
    void tma3d_mode(const FineOp<T>& op, const C* x, const C* b, C* out, int64_t ld, int nrhs) {
        if (nrhs >= 2) tma3d_launch<MODE, 2>(op, x, b, out, ld, nrhs);
        else tma3d_launch<MODE, 1>(op, x, b, out, ld, nrhs);
    }
    void tma3d_dispatch(int mode, const FineOp<T>& op, const C* x, const C* b, C* out, int64_t ld, int nrhs) {
        if (mode == MODE_APPLY) tma3d_mode<MODE_APPLY>(op, x, b, out, ld, nrhs);
        else if (mode == MODE_RESID) tma3d_mode<MODE_RESID>(op, x, b, out, ld, nrhs);
        else tma3d_mode<MODE_JACOBI>(op, x, b, out, ld, nrhs);
    }
    bool tma_ok_level(const Level& L) const {
        if (pb.dim != 3 || fine_kernel != FK_TMA) return false;
        const int64_t es = 2 * sizeof(T);
        return (es * L.p0) % 16 == 0 && (es * L.p0 * L.n[1]) % 16 == 0 && (es * L.N) % 16 == 0;
    }
    // ---- 27-point TMA kernel: tile shape and z-chunking ------------------------------------------------------------
    // Tile shapes (TX x TY columns per CTA): 16 x 8, or the "alternative" 11 x 11 (ComplexF64) / 12 x 10 (ComplexF32,
    // whose TMA rows must be 16-byte multiples), which wastes far fewer lanes on the small 2^k+1 grids (65 = 4*16+1
    // fills 73 % of a 16 x 8 tiling, 92 % of an 11 x 11 one) at the price of two-way bank conflicts on rows that a
    // quarter-warp straddles.  HH_COARSE_TILE=16x8|alt forces one.
    static constexpr int ALT_TX = sizeof(T) == 8 ? 11 : 12, ALT_TY = sizeof(T) == 8 ? 11 : 10;
    int coarse_tile(const Level& L) const {
        if (force_tile >= 0) return force_tile;
        const double t16 = (double)((L.n[0] + 15) / 16) * ((L.n[1] + 7) / 8);
        const double talt = (double)((L.n[0] + ALT_TX - 1) / ALT_TX) * ((L.n[1] + ALT_TY - 1) / ALT_TY);
        return talt < 0.96 * t16 ? 1 : 0;
    }
    // z-chunks of equal length such that the waves of one-CTA-per-SM CTAs are full and the two extra input planes a
    // chunk stages (which cost a third of a plane each) stay a small share: minimise waves x (chunk + 2/3)
    static void balanced_chunks(int nz, int64_t ctas_per_chunk, int& zchunk, int& nzc) {
        double best = 1e300;
        zchunk = nz;
        nzc = 1;
        for (int c = 1; c <= nz; ++c) {
            const int zc = (nz + c - 1) / c, cc = (nz + zc - 1) / zc;
            if (cc != c) continue;  // same chunk length as a smaller count
            const double waves = std::ceil((double)ctas_per_chunk * cc / 148.0);
            const double cost = waves * (zc + 0.67) + 0.02 * cc;  // tie-break: fewer, longer chunks
            if (cost < best) best = cost, zchunk = zc, nzc = cc;
            if (zc <= 2) break;
        }
    }
    template <int MODE, int KB, int TX, int TY>
    void coarse_tma_launch(const Level& L, const C* coef, const C* x, const C* b, C* out, int nrhs) {
        typedef CoarseTmaCfg<T, MODE, KB, TX, TY> Cfg;
        constexpr size_t smem = (size_t)Cfg::NS * Cfg::STAGE_BYTES + Cfg::NS * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            HH_CUDA(cudaFuncSetAttribute(k_coarse3d_tma<T, MODE, KB, TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (L.n[0] + TX - 1) / TX, ty = (L.n[1] + TY - 1) / TY;
        int zchunk, nzc;
        balanced_chunks(zend(L) - zbeg(L), (int64_t)tx * ty * groups, zchunk, nzc);
        dim3 g(tx * groups, ty, nzc);
        TmaDesc mx = make_tmap_g(x, L.n, L.p0, L.N, Cfg::PX, TY + 2, KB, nrhs);
        TmaDesc mc = make_tmap_g(coef, L.n, L.p0, L.N, TX, TY, 9, 27);
        TmaDesc mb = (MODE != MODE_APPLY) ? make_tmap_g(b, L.n, L.p0, L.N, TX, TY, KB, nrhs) : mx;
        TmaDesc md = (MODE == MODE_JACOBI) ? make_tmap_g(L.dinv.p, L.n, L.p0, L.N, TX, TY, 1, 0) : mx;
        k_coarse3d_tma<T, MODE, KB, TX, TY><<<g, Cfg::THREADS, smem, stream>>>(mx, mc, mb, md, out, L.n[0], L.n[1], L.n[2], L.p0, L.N, nrhs,
                                                                                zchunk, groups, zbeg(L), zend(L));
    }
    template <int MODE, int KB>
    void coarse_tma_tile(const Level& L, const C* coef, const C* x, const C* b, C* out, int nrhs) {
        if (coarse_tile(L) == 1) coarse_tma_launch<MODE, KB, ALT_TX, ALT_TY>(L, coef, x, b, out, nrhs);
        else coarse_tma_launch<MODE, KB, 16, 8>(L, coef, x, b, out, nrhs);
    }
    template <int MODE>
    void coarse_tma_mode(const Level& L, const C* coef, const C* x, const C* b, C* out, int nrhs) {
        int kb = 1;
        while (kb * 2 <= nrhs && kb * 2 <= 8) kb *= 2;
        if (kb == 8) coarse_tma_tile<MODE, 8>(L, coef, x, b, out, nrhs);
        else if (kb == 4) coarse_tma_tile<MODE, 4>(L, coef, x, b, out, nrhs);
        else if (kb == 2) coarse_tma_tile<MODE, 2>(L, coef, x, b, out, nrhs);
        else coarse_tma_tile<MODE, 1>(L, coef, x, b, out, nrhs);
    }
    static int coarse_kb(int nrhs) {
        const int pref = sizeof(T) == 8 ? 4 : 8;
        int kb = 1;
        while (kb * 2 <= nrhs && kb * 2 <= pref) kb *= 2;
        return kb;
    }
    template <int MODE, int KB>
    void coarse3d_launch(const Level& L, const C* x, const C* b, C* out, int nrhs) {
        constexpr int TY = 8;
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (L.n[0] + 31) / 32, ty = (L.n[1] + TY - 1) / TY;
        int zchunk, nzc;
        zchunks(zend(L) - zbeg(L), tx * ty, groups, 16, zchunk, nzc);
        dim3 g(tx * groups, ty, nzc), blk(32, TY, 1);
        k_coarse3d_zmarch<T, MODE, KB, TY, COARSE_MINB><<<g, blk, 0, stream>>>(coarse_op(L), x, b, out, L.N, nrhs, zchunk, groups);
    }
    template <int MODE>
    void coarse3d_mode(const Level& L, const C* x, const C* b, C* out, int nrhs) {
        const int kb = coarse_kb(nrhs);
        if (kb == 8) coarse3d_launch<MODE, 8>(L, x, b, out, nrhs);
        else if (kb == 4) coarse3d_launch<MODE, 4>(L, x, b, out, nrhs);
        else if (kb == 2) coarse3d_launch<MODE, 2>(L, x, b, out, nrhs);
        else coarse3d_launch<MODE, 1>(L, x, b, out, nrhs);
    }
    // ---- 2-D grids: RHS-marching TMA kernel (k_stencil2d_tma) ----------------------------------------------------
    bool tma2d_ok(int sy, int n1, int64_t ld, const C* x, const C* b, int mode) const {
        if (pb.dim != 2 || fine_kernel != FK_TMA) return false;
        const int64_t es = 2 * sizeof(T);
        return (es * sy) % 16 == 0 && (es * ld) % 16 == 0 && ((uintptr_t)x % 16 == 0) && (mode == MODE_APPLY || (uintptr_t)b % 16 == 0) &&
               n1 >= 1;
    }
    template <int MODE, int KB, bool COARSE>
    void stencil2d_launch(const FineOp<T>& fop, const CoarseOp<T>& cop, const int* n, int sy, const C* x, const C* b, C* out,
                          int64_t ld, int nrhs, T damp) {
        typedef Rhs2dCfg<T, MODE, KB> Cfg;
        constexpr size_t smem = (size_t)Cfg::NS * Cfg::STAGE_BYTES + Cfg::NS * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            HH_CUDA(cudaFuncSetAttribute(k_stencil2d_tma<T, MODE, KB, COARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int tx = (n[0] + Cfg::TX - 1) / Cfg::TX, ty = (n[1] + Cfg::TY - 1) / Cfg::TY;
        // chunks of the RHS block (multiples of KB) so that small grids still fill the SMs
        int nch = std::max(1, std::min((nrhs + KB - 1) / KB, (2 * 148 + tx * ty - 1) / (tx * ty)));
        int rchunk = ((nrhs + nch - 1) / nch + KB - 1) / KB * KB;
        nch = (nrhs + rchunk - 1) / rchunk;
        const int nn[3] = {n[0], n[1], 1};
        TmaDesc mx = make_tmap_g(x, nn, sy, ld, Cfg::PX, Cfg::TY + 2, KB, nrhs);
        TmaDesc mb = (MODE != MODE_APPLY) ? make_tmap_g(b, nn, sy, ld, Cfg::TX, Cfg::TY, KB, nrhs) : mx;
        k_stencil2d_tma<T, MODE, KB, COARSE><<<dim3(tx, ty, nch), 256, smem, stream>>>(fop, cop, mx, mb, out, ld, nrhs, rchunk, damp);
    }
    template <bool COARSE>
    void stencil2d_dispatch(int mode, const FineOp<T>& fop, const CoarseOp<T>& cop, const int* n, int sy, const C* x, const C* b,
                            C* out, int64_t ld, int nrhs, T damp) {
        if (nrhs >= 2) {
            if (mode == MODE_APPLY) stencil2d_launch<MODE_APPLY, 2, COARSE>(fop, cop, n, sy, x, b, out, ld, nrhs, damp);
            else if (mode == MODE_RESID) stencil2d_launch<MODE_RESID, 2, COARSE>(fop, cop, n, sy, x, b, out, ld, nrhs, damp);
            else stencil2d_launch<MODE_JACOBI, 2, COARSE>(fop, cop, n, sy, x, b, out, ld, nrhs, damp);
        } else {
            if (mode == MODE_APPLY) stencil2d_launch<MODE_APPLY, 1, COARSE>(fop, cop, n, sy, x, b, out, ld, nrhs, damp);
            else if (mode == MODE_RESID) stencil2d_launch<MODE_RESID, 1, COARSE>(fop, cop, n, sy, x, b, out, ld, nrhs, damp);
            else stencil2d_launch<MODE_JACOBI, 1, COARSE>(fop, cop, n, sy, x, b, out, ld, nrhs, damp);
        }
    }
    void fine_jacobi0(const FineOp<T>& op, const C* b, C* out, int64_t ld, int nrhs, T damp) {
        dim3 g, blk;
        grid3(pb.n, pb.dim, g, blk);
        const double N = (double)pb.N();
        launch(T_FINE_JACOBI0, 2 * S * N * nrhs + 2.0 * CR * N, [&] {
            if (pb.dim == 3) k_fine_jacobi0<T, 3><<<g, blk, 0, stream>>>(op, b, out, ld, nrhs, damp);
            else k_fine_jacobi0<T, 2><<<dim3(g.x, g.y, rhs_slices(g, nrhs)), blk, 0, stream>>>(op, b, out, ld, nrhs, damp);
        });
    }
    CoarseOp<T> coarse_op(const Level& L) const {
        CoarseOp<T> op;
        op.coef = (use_scaled && L.scoef.p) ? L.scoef.p : L.coef.p;  // scaled: A diag(dinv), see k_scale_columns
        op.dinv = L.dinv.p;
        for (int d = 0; d < 3; ++d) op.n[d] = L.n[d];
        op.sy = L.p0;
        op.N = L.N;
        op.zb = zbeg(L);
        op.ze = zend(L);
        return op;
    }
    // scaled: apply the column-scaled copy A diag(dinv) of the level's operator (right-preconditioned Jacobi-GMRES)
    void coarse_stencil(int mode, const Level& L, const C* x, const C* b, C* out, int nrhs, bool scaled = false) {
        const int li = slab ? (int)(&L - levels.data()) : 0;
        HH_REQUIRE(!scaled || L.scoef.p != nullptr, HH_ERR_STATE, "no scaled operator on this level");
        use_scaled = scaled;
        with_halos({{li, x, 0}}, nrhs, L.zb, L.ze, [&](int z0, int z1) {
            rz_b = z0;  // the launchers below read the plane range through zbeg / zend
            rz_e = z1;
            coarse_stencil_range(mode, L, x, b, out, nrhs);
            rz_b = rz_e = -1;
        });
        use_scaled = false;
    }
    void coarse_stencil_range(int mode, const Level& L, const C* x, const C* b, C* out, int nrhs) {
        dim3 g, blk;
        grid3z(L.n, pb.dim, zbeg(L), zend(L), g, blk);
        const double N = pb.dim == 3 ? (double)L.n[0] * L.n[1] * (zend(L) - zbeg(L)) : (double)L.Nlog;
        const int NS = pb.dim == 3 ? 27 : 9;
        const bool use_tma = tma_ok_level(L) && ((uintptr_t)x % 16 == 0) && (mode == MODE_APPLY || (uintptr_t)b % 16 == 0);
        int KB = (pb.dim == 3 && fine_kernel != FK_SIMPLE) ? coarse_kb(nrhs) : 4;
        if (use_tma) {
            KB = 1;
            while (KB * 2 <= nrhs && KB * 2 <= 8) KB *= 2;
        }
        // algorithmic bytes: the 3^dim coefficients of a node once per launch (k = the whole batch shares
        // one read; re-reads by further RHS groups are L2 traffic the kernel has to earn, not algorithmic bytes)
        (void)KB;
        const double coefb = NS * S * N;
        double bytes;
        int tag;
        if (mode == MODE_APPLY) bytes = 2 * S * N * nrhs + coefb, tag = T_COARSE_APPLY;
        else if (mode == MODE_RESID) bytes = 3 * S * N * nrhs + coefb, tag = T_COARSE_RESID;
        else bytes = 3 * S * N * nrhs + coefb + S * N, tag = T_COARSE_JACOBI;
        CoarseOp<T> op = coarse_op(L);
        const int64_t ld = L.N;
        if (use_tma) {
            launch(tag, bytes, [&] {
                if (mode == MODE_APPLY) coarse_tma_mode<MODE_APPLY>(L, op.coef, x, b, out, nrhs);
                else if (mode == MODE_RESID) coarse_tma_mode<MODE_RESID>(L, op.coef, x, b, out, nrhs);
                else coarse_tma_mode<MODE_JACOBI>(L, op.coef, x, b, out, nrhs);
            });
            return;
        }
        if (tma2d_ok(L.p0, L.n[1], ld, x, b, mode)) {
            launch(tag, bytes, [&] { stencil2d_dispatch<true>(mode, FineOp<T>{}, op, L.n, L.p0, x, b, out, ld, nrhs, T(0)); });
            return;
        }
        if (pb.dim == 3 && fine_kernel != FK_SIMPLE) {
            launch(tag, bytes, [&] {
                if (mode == MODE_APPLY) coarse3d_mode<MODE_APPLY>(L, x, b, out, nrhs);
                else if (mode == MODE_RESID) coarse3d_mode<MODE_RESID>(L, x, b, out, nrhs);
                else coarse3d_mode<MODE_JACOBI>(L, x, b, out, nrhs);
            });
            return;
        }
        launch(tag, bytes, [&] {
            if (pb.dim == 3) {
                if (mode == MODE_APPLY) k_coarse_stencil<T, 3, MODE_APPLY, 4><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs);
                else if (mode == MODE_RESID) k_coarse_stencil<T, 3, MODE_RESID, 4><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs);
                else k_coarse_stencil<T, 3, MODE_JACOBI, 4><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs);
            } else {
                if (mode == MODE_APPLY) k_coarse_stencil<T, 2, MODE_APPLY, 4><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs);
                else if (mode == MODE_RESID) k_coarse_stencil<T, 2, MODE_RESID, 4><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs);
                else k_coarse_stencil<T, 2, MODE_JACOBI, 4><<<g, blk, 0, stream>>>(op, x, b, out, ld, nrhs);
            }
        });
    }
    void diag_scale(int tag, const C* dinv, const C* b, C* out, int64_t N, int nrhs) {
        launch(tag, 2 * S * (double)N * nrhs + S * (double)N, [&] {
            const unsigned nb = (unsigned)((N + 255) / 256);
            const unsigned ny = (unsigned)std::max<int64_t>(1, std::min<int64_t>(nrhs, (4 * 148 + nb - 1) / nb));
            k_diag_scale<T><<<dim3(nb, ny), 256, 0, stream>>>(dinv, b, out, N, N, nrhs);
        });
    }
    template <int KB>
    void restrict_tma_launch(const Level& F, const Level& Cc, const C* r, C* bc, int nrhs) {
        typedef RestrictCfg<T, KB> Cfg;
        constexpr size_t smem = (size_t)Cfg::NS * Cfg::STAGE_BYTES + Cfg::NS * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            HH_CUDA(cudaFuncSetAttribute(k_restrict3d_tma<T, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int groups = (nrhs + KB - 1) / KB;
        const int tx = (Cc.n[0] + Cfg::CTX - 1) / Cfg::CTX, ty = (Cc.n[1] + Cfg::CTY - 1) / Cfg::CTY;
        // two CTAs per SM: chunks of coarse planes such that the grid is a few waves of 296
        const int nk = Cc.ze - Cc.zb;
        int nzc = (int)std::max<int64_t>(1, std::min<int64_t>(nk, (4 * 296 + (int64_t)tx * ty * groups - 1) / ((int64_t)tx * ty * groups)));
        int kchunk = (nk + nzc - 1) / nzc;
        nzc = (nk + kchunk - 1) / kchunk;
        dim3 g(tx * groups, ty, nzc);
        TmaDesc mr = make_tmap_g(r, F.n, F.p0, F.N, Cfg::FX, Cfg::FY, KB, nrhs);
        k_restrict3d_tma<T, KB><<<g, 128, smem, stream>>>(mr, bc, Cc.n[0], Cc.n[1], Cc.p0, Cc.N, nrhs, kchunk, groups, Cc.zb, Cc.ze);
    }
    void restrict_to(const Level& F, const Level& Cc, const C* r, C* bc, int nrhs) {
        dim3 g, blk;
        grid3z(Cc.n, pb.dim, Cc.zb, Cc.ze, g, blk);
        halo_exchange((int)(&F - levels.data()), r, nrhs, 0, HALO_LOWER);
        if (pb.dim == 3 && tma_restrict && tma_ok_level(F) && ((uintptr_t)r % 16 == 0)) {
            launch(T_RESTRICT, S * ((double)F.Nlog + (double)Cc.Nlog) * nrhs, [&] {
                if (nrhs >= 2) restrict_tma_launch<2>(F, Cc, r, bc, nrhs);
                else restrict_tma_launch<1>(F, Cc, r, bc, nrhs);
            });
            return;
        }
        launch(T_RESTRICT, S * ((double)F.Nlog + (double)Cc.Nlog) * nrhs, [&] {
            if (pb.dim == 3)
                k_restrict<T, 3><<<g, blk, 0, stream>>>(r, bc, F.n[0], F.n[1], F.n[2], Cc.n[0], Cc.n[1], Cc.n[2], F.p0, Cc.p0, F.N, Cc.N, nrhs,
                                                        Cc.zb, Cc.ze, F.koff, F.n2g);
            else
                k_restrict<T, 2><<<dim3(g.x, g.y, rhs_slices(g, nrhs)), blk, 0, stream>>>(r, bc, F.n[0], F.n[1], 1, Cc.n[0], Cc.n[1], 1, F.p0, Cc.p0, F.N,
                                                                                        Cc.N, nrhs, 0, 1, 0, 1);
        });
    }
    void prolong_add(const Level& F, const Level& Cc, C* x, const C* xc, int nrhs) {
        dim3 g, blk;
        grid3z(F.n, pb.dim, F.zb, F.ze, g, blk);
        halo_exchange((int)(&Cc - levels.data()), xc, nrhs, 0, HALO_UPPER);
        launch(T_PROLONG, S * (2.0 * (double)F.Nlog + (double)Cc.Nlog) * nrhs, [&] {
            if (pb.dim == 3)
                k_prolong_add<T, 3><<<g, blk, 0, stream>>>(x, xc, F.n[0], F.n[1], F.n[2], F.p0, Cc.p0, Cc.n[1], F.N, Cc.N, nrhs, F.zb, F.ze);
            else
                k_prolong_add<T, 2><<<dim3(g.x, g.y, rhs_slices(g, nrhs)), blk, 0, stream>>>(x, xc, F.n[0], F.n[1], 1, F.p0, Cc.p0, Cc.n[1], F.N, Cc.N, nrhs,
                                                                                           0, 1);
        });
    }

    // ------------------------------------------------------------------ vector kernels
    void copy_vec(const C* in, C* out, int64_t N, int nrhs) {
        if (in == out) return;
        launch(T_COPY, 2 * S * (double)N * nrhs, [&] {
            HH_CUDA(cudaMemcpyAsync(out, in, (size_t)N * nrhs * sizeof(C), cudaMemcpyDeviceToDevice, stream));
        });
    }
    void zero_vec(C* v, int64_t N, int nrhs) {
        launch(T_COPY, S * (double)N * nrhs, [&] { HH_CUDA(cudaMemsetAsync(v, 0, (size_t)N * nrhs * sizeof(C), stream)); });
    }
    // partial sums of conj(V_i).w (i<nv) [+ |w|^2 if with_norm]; returns nblk
    int multidot(const C* const* V, int nv, const C* w, const Span& sp, int nrhs, bool with_norm, zc* partial) {
        const int nblk = multidot_launch(V, nv, w, sp, nrhs, with_norm, partial);
        return slab_reduce(partial, (nv + (with_norm ? 1 : 0)) * nrhs, nblk);
    }
    // the kernel only: the block partials stay in `partial` (layout of k_multidot), nblk is returned
    int multidot_launch(const C* const* V, int nv, const C* w, const Span& sp, int nrhs, bool with_norm, zc* partial) {
        HH_REQUIRE(nv >= 0 && nv <= HH_MAXV, HH_ERR_ARG, "multidot: too many vectors");
        const int64_t N = sp.len, ld = sp.ld;
        const int nblk = vec_blocks(N, nrhs);
        HH_REQUIRE((size_t)(nv + 1) * nrhs * nblk <= d_partial.n, HH_ERR_STATE, "multidot: partial buffer too small");
        VecList<T> L;
        for (int i = 0; i < HH_MAXV; ++i) L.v[i] = i < nv ? V[i] + sp.off : nullptr;
        w += sp.off;
        dim3 g(nblk, nrhs);
        launch(T_DOT, S * (double)N * nrhs * (nv + 1), [&] {
#define HH_MD(NV)                                                                              \
    case NV:                                                                                   \
        if (with_norm) k_multidot<T, NV, true><<<g, 256, 0, stream>>>(L, w, N, ld, partial);   \
        else k_multidot<T, NV, false><<<g, 256, 0, stream>>>(L, w, N, ld, partial);            \
        break;
            switch (nv) {
                HH_MD(0) HH_MD(1) HH_MD(2) HH_MD(3) HH_MD(4) HH_MD(5) HH_MD(6) HH_MD(7) HH_MD(8) HH_MD(9) HH_MD(10)
            }
#undef HH_MD
        });
        return nblk;
    }
    int multiaxpy(const C* const* V, int nv, C* w, const Span& sp, int nrhs, const zc* coef, int cstride, bool negate,
                  bool with_norm, zc* partial, const zc* post = nullptr) {
        const int nblk = multiaxpy_launch(V, nv, w, sp, nrhs, coef, cstride, negate, with_norm, partial, post);
        return with_norm ? slab_reduce(partial, nrhs, nblk) : nblk;
    }
    int multiaxpy_launch(const C* const* V, int nv, C* w, const Span& sp, int nrhs, const zc* coef, int cstride, bool negate,
                         bool with_norm, zc* partial, const zc* post = nullptr) {
        HH_REQUIRE(nv >= 1 && nv <= HH_MAXV, HH_ERR_ARG, "multiaxpy: bad vector count");
        const int64_t N = sp.len, ld = sp.ld;
        const int nblk = vec_blocks(N, nrhs);
        VecList<T> L;
        for (int i = 0; i < HH_MAXV; ++i) L.v[i] = i < nv ? V[i] + sp.off : nullptr;
        w += sp.off;
        dim3 g(nblk, nrhs);
        launch(T_AXPY, S * (double)N * nrhs * (nv + 2), [&] {
#define HH_MA(NV)                                                                                                 \
    case NV:                                                                                                      \
        if (with_norm) k_multiaxpy<T, NV, true><<<g, 256, 0, stream>>>(L, w, N, ld, coef, cstride, negate, post, partial); \
        else k_multiaxpy<T, NV, false><<<g, 256, 0, stream>>>(L, w, N, ld, coef, cstride, negate, post, partial);          \
        break;
            switch (nv) { HH_MA(1) HH_MA(2) HH_MA(3) HH_MA(4) HH_MA(5) HH_MA(6) HH_MA(7) HH_MA(8) HH_MA(9) HH_MA(10) }
#undef HH_MA
        });
        return nblk;
    }
    // x (+)= dinv .* sum_{i<nv} coef[r*cstride + i] V_i on the span (k_combine)
    void combine(const C* const* V, int nv, const C* dinv, C* x, const Span& sp, int nrhs, const zc* coef, int cstride,
                 bool accumulate, int tag) {
        HH_REQUIRE(nv >= 1 && nv <= HH_MAXV, HH_ERR_ARG, "combine: bad vector count");
        const int64_t N = sp.len, ld = sp.ld;
        const int nblk = vec_blocks(N, nrhs);
        VecList<T> L;
        for (int i = 0; i < HH_MAXV; ++i) L.v[i] = i < nv ? V[i] + sp.off : nullptr;
        dim3 g(nblk, nrhs);
        launch(tag, S * (double)N * (nrhs * (nv + (accumulate ? 2 : 1)) + 1), [&] {
#define HH_CB(NV)                                                                                                      \
    case NV:                                                                                                           \
        if (accumulate) k_combine<T, NV, true><<<g, 256, 0, stream>>>(L, dinv + sp.off, x + sp.off, N, ld, coef, cstride); \
        else k_combine<T, NV, false><<<g, 256, 0, stream>>>(L, dinv + sp.off, x + sp.off, N, ld, coef, cstride);           \
        break;
            switch (nv) { HH_CB(1) HH_CB(2) HH_CB(3) HH_CB(4) HH_CB(5) HH_CB(6) HH_CB(7) HH_CB(8) HH_CB(9) HH_CB(10) }
#undef HH_CB
        });
    }

    // ------------------------------------------------------------------ hierarchy (MGsetup)
    void clear() override {
        hoH.coef.release();
        levels.clear();
        have_hierarchy = false;
        inv_dense.release();
        kcap = 0;
        kry.release();
        kry_cap = 0;
    }
    bool hierarchy_exists() const override { return have_hierarchy; }
    void level_nodes(int level, int64_t* out) const override {
        HH_REQUIRE(have_hierarchy && level >= 0 && level < (int)levels.size(), HH_ERR_ARG, "bad level");
        for (int d = 0; d < pb.dim; ++d) out[d] = levels[level].n[d];
    }

    // HO mode.  hoH gets the un-shifted operator (always: it is the Krylov operator); levels[0] gets the shifted one and
    // damp/diag unless this solver only runs the Krylov method (mixed precision).  The stencils are evaluated on the
    // device from Float64 copies of m and gamma (k_ho_stencil: the closed form the host export hh_ho_stencil uses);
    // HH_HO_BUILD=host keeps the round-1 path (host loop + upload of 27 N coefficients) as an A/B.
    void build_ho_levels(const hh_mg_options& o) {
        const int NS = pb.dim == 3 ? 27 : 9, center = pb.dim == 3 ? 13 : 4;
        const int64_t Nd = pb.N();  // dense node count of the caller's model arrays
        HH_REQUIRE((int64_t)ho_m.size() == Nd && (int64_t)ho_g.size() == Nd, HH_ERR_STATE, "high-order operator: no model");
        int64_t nn[3] = {pb.n[0], pb.n[1], pb.n[2]};
        const Level& L0 = levels[0];
        for (int d = 0; d < 3; ++d) hoH.n[d] = L0.n[d];
        hoH.p0 = L0.p0;
        hoH.N = L0.N;
        hoH.Nlog = L0.Nlog;
        hoH.zb = L0.zb;
        hoH.ze = L0.ze;
        hoH.koff = L0.koff;
        hoH.n2g = L0.n2g;
        const double sw2 = o.shift[0] * pb.w_re * pb.w_re;
        const char* hb = getenv("HH_HO_BUILD");
        if (!(hb && !strcmp(hb, "host"))) {
            HH_REQUIRE(pb.w_re != 0.0, HH_ERR_ARG, "high-order operator: Re(omega) must be non-zero");
            const HoGeom g = ho_geom(pb.dim, nn, pb.h, pb.w_re, pb.w_im, pb.neumann_top, pb.sommerfeld, ho_beta);
            DevBuf<double> dm, dg;
            dm.alloc((size_t)Nd);
            dg.alloc((size_t)Nd);
            HH_CUDA(cudaMemcpyAsync(dm.p, ho_m.data(), (size_t)Nd * sizeof(double), cudaMemcpyHostToDevice, stream));
            HH_CUDA(cudaMemcpyAsync(dg.p, ho_g.data(), (size_t)Nd * sizeof(double), cudaMemcpyHostToDevice, stream));
            auto build = [&](DevBuf<C>& dst, double shift_w2) {
                dst.alloc((size_t)NS * L0.N);
                if (L0.N != Nd) HH_CUDA(cudaMemsetAsync(dst.p, 0, (size_t)NS * L0.N * sizeof(C), stream));  // ghost columns stay zero
                const unsigned nb = (unsigned)std::min<int64_t>((Nd + 255) / 256, 148 * 16);
                launch(T_SETUP, 0, [&] {
                    k_ho_stencil<T><<<nb, 256, 0, stream>>>(g, dm.p, dg.p, shift_w2, o.do_transpose ? 1 : 0, L0.p0, L0.N, dst.p);
                });
            };
            build(hoH.coef, 0.0);
            if (!krylov_only) {
                // shifted operator of the hierarchy: + i shift Re(w)^2 m on the diagonal (GetHelmholtzShiftOP, GetHelmholtz.jl:81-83)
                Level& Lm = levels[0];
                build(Lm.coef, sw2);
                Lm.dinv.alloc(Lm.N);
                launch(T_SETUP, 0, [&] {
                    k_coarse_dinv<T><<<(unsigned)((Lm.N + 255) / 256), 256, 0, stream>>>(Lm.coef.p + (int64_t)center * Lm.N, Lm.dinv.p, Lm.N, (T)o.relax_param);
                });
            }
            HH_CUDA(cudaStreamSynchronize(stream));  // dm / dg are released on return
            return;
        }
        std::vector<double> host((size_t)2 * NS * Nd);
        build_ho_stencil(pb.dim, nn, pb.h, ho_m.data(), ho_g.data(), pb.w_re, pb.w_im, pb.neumann_top, pb.sommerfeld, ho_beta,
                         host.data());
        std::vector<double> adj;  // transposed hierarchy (doTranspose = 1): the same kernels on the adjoint stencils
        auto upload = [&](DevBuf<C>& dst) {  // host Float64 dense -> device precision T in the level's row pitch
            const double* src = host.data();
            if (o.do_transpose) {
                adj.resize(host.size());
                adjoint_stencil(pb.dim, nn, host.data(), adj.data());
                src = adj.data();
            }
            std::vector<C> tmp((size_t)NS * Nd);
            for (size_t e = 0; e < tmp.size(); ++e) tmp[e] = mk<T>((T)src[2 * e], (T)src[2 * e + 1]);
            dst.alloc((size_t)NS * L0.N);
            if (L0.N == Nd) {
                HH_CUDA(cudaMemcpyAsync(dst.p, tmp.data(), tmp.size() * sizeof(C), cudaMemcpyHostToDevice, stream));
                HH_CUDA(cudaStreamSynchronize(stream));
                return;
            }
            DevBuf<C> dense;
            dense.alloc(tmp.size());
            HH_CUDA(cudaMemcpyAsync(dense.p, tmp.data(), tmp.size() * sizeof(C), cudaMemcpyHostToDevice, stream));
            HH_CUDA(cudaMemsetAsync(dst.p, 0, (size_t)NS * L0.N * sizeof(C), stream));  // ghost columns stay zero
            const int64_t rows = (int64_t)pb.n[1] * pb.n[2];
            launch(T_SETUP, 0, [&] {  // the NS coefficient arrays are repitched like NS right-hand sides
                k_repitch<C><<<dim3(148, NS), 256, 0, stream>>>(dense.p, dst.p, pb.n[0], rows, pb.n[0], L0.p0, Nd, L0.N);
            });
            HH_CUDA(cudaStreamSynchronize(stream));
        };
        upload(hoH.coef);
        if (krylov_only) return;
        // shifted operator of the hierarchy: + i shift Re(w)^2 m on the diagonal (GetHelmholtzShiftOP, GetHelmholtz.jl:81-83)
        for (int64_t p = 0; p < Nd; ++p) host[2 * ((int64_t)center * Nd + p) + 1] += sw2 * ho_m[p];
        Level& Lm = levels[0];
        upload(Lm.coef);
        Lm.dinv.alloc(Lm.N);
        launch(T_SETUP, 0, [&] {
            k_coarse_dinv<T><<<(unsigned)((Lm.N + 255) / 256), 256, 0, stream>>>(Lm.coef.p + (int64_t)center * Lm.N, Lm.dinv.p, Lm.N, (T)o.relax_param);
        });
    }

    void set_level_planes(int l) {
        Level& L = levels[l];
        if (slab) {
            L.zb = sgeo[l].zb;
            L.ze = sgeo[l].ze;
            L.koff = sgeo[l].koff;
            L.n2g = sgeo[l].n2g;
        } else {
            L.zb = 0;
            L.ze = L.n[2];
            L.koff = 0;
            L.n2g = L.n[2];
        }
    }

    void setup(const hh_mg_options& o) override {
        HH_CUDA(cudaSetDevice(device));
        auto t0 = std::chrono::steady_clock::now();
        HH_REQUIRE(o.levels >= 1 && o.levels <= HH_MAX_LEVELS, HH_ERR_ARG, "levels out of range");
        HH_REQUIRE(d_m.p != nullptr, HH_ERR_STATE, "no model set");
        HH_REQUIRE(o.relax_type == HH_RELAX_JAC || o.relax_type == HH_RELAX_JAC_GMRES, HH_ERR_ARG, "bad relax_type");
        HH_REQUIRE(o.cycle_type >= HH_CYCLE_V && o.cycle_type <= HH_CYCLE_K, HH_ERR_ARG, "bad cycle_type");
        HH_REQUIRE(o.coarse_type == HH_COARSE_LU || o.coarse_type == HH_COARSE_GMRES, HH_ERR_ARG, "bad coarse_type");
        if (slab) {
            HH_REQUIRE(pb.dim == 3 && fine_kernel != FK_SIMPLE, HH_ERR_UNSUPPORTED, "slab decomposition needs a 3-D grid");
            HH_REQUIRE((int)sgeo.size() == o.levels, HH_ERR_ARG,
                       "slab decomposition: hh_setup must use the number of levels given to hh_create_slab*");
            HH_REQUIRE(o.coarse_type == HH_COARSE_GMRES, HH_ERR_UNSUPPORTED,
                       "slab decomposition: the coarsest solve must be the inexact one (coarse_type GMRES)");
        }
        clear();
        opt = o;
        levels.resize(krylov_only ? 1 : o.levels);
        for (int d = 0; d < 3; ++d) levels[0].n[d] = pb.n[d];
        levels[0].p0 = fine_sy();
        levels[0].N = fineN();
        levels[0].Nlog = pb.N();
        set_level_planes(0);
        if (ho) {
            HH_REQUIRE(!slab, HH_ERR_UNSUPPORTED, "the high-order operator is not available on a slab handle");
            build_ho_levels(o);
        }
        if (krylov_only) {  // the cycle lives in another solver (prec_hook): only the fine-level geometry is needed
            have_hierarchy = true;
            setup_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            return;
        }
        for (int l = 1; l < o.levels; ++l) {
            for (int d = 0; d < 3; ++d) {
                const int nf = levels[l - 1].n[d];
                if (d < pb.dim && !(slab && d == 2)) {
                    HH_REQUIRE(nf >= 3 && (nf % 2) == 1, HH_ERR_ARG,
                               "cannot coarsen: every level needs an odd node count >= 3 in each dimension (cells "
                               "divisible by 2^(levels-1))");
                    levels[l].n[d] = (nf + 1) / 2;
                } else if (d >= pb.dim) {
                    levels[l].n[d] = 1;
                }
            }
            if (slab) levels[l].n[2] = sgeo[l].nloc;  // local planes of the slab (hh_slab.cuh)
            set_level_planes(l);
            levels[l].Nlog = (int64_t)levels[l].n[0] * levels[l].n[1] * levels[l].n[2];
            // the exact coarsest solve works on a dense unknown numbering: no padding on that level
            const bool dense_level = (l == o.levels - 1 && o.coarse_type == HH_COARSE_LU);
            levels[l].p0 = dense_level ? levels[l].n[0] : pitch_of(levels[l].n[0]);
            levels[l].N = (int64_t)levels[l].p0 * levels[l].n[1] * levels[l].n[2];
        }
        mg_fine = fine_op(o.shift[0], o.do_transpose, true);
        // (HO mode: no matrix-free diagonal arrays, so the fused matrix-free kernels stay off and level 0 takes the
        // generic stored-stencil path)
        if (o.levels > 1 && !ho) precompute_diag(mg_fine, mg_cdiag, &mg_dinv, (T)o.relax_param);
        const int NS = pb.dim == 3 ? 27 : 9;
        const int center = pb.dim == 3 ? 13 : 4;
        for (int l = 1; l < o.levels; ++l) {
            Level& Lc = levels[l];
            Level& Lf = levels[l - 1];
            Lc.coef.alloc((size_t)NS * Lc.N);
            Lc.dinv.alloc(Lc.N);
            HH_CUDA(cudaMemsetAsync(Lc.coef.p, 0, (size_t)NS * Lc.N * sizeof(C), stream));  // ghost columns stay zero
            const int64_t tot = (int64_t)Lc.n[0] * Lc.n[1] * (Lc.ze - Lc.zb) * NS;
            const unsigned nb = (unsigned)((tot + 127) / 128);
            if (slab && l >= 2) halo_exchange(l - 1, Lf.coef.p, NS, 0, HALO_LOWER);  // rows of the fine plane just below the slab
            launch(T_SETUP, 0, [&] {
                if (l == 1 && !ho) {
                    if (pb.dim == 3) {
                        FineCoef<T, 3> A{mg_fine};
                        k_galerkin<T, 3, FineCoef<T, 3>><<<nb, 128, 0, stream>>>(A, Lf.n[0], Lf.n[1], Lf.n[2], Lc.n[0], Lc.n[1], Lc.n[2], Lc.p0, Lc.N, Lc.coef.p,
                                                                                 Lc.zb, Lc.ze, Lf.koff, Lf.n2g, Lc.koff, Lc.n2g);
                    } else {
                        FineCoef<T, 2> A{mg_fine};
                        k_galerkin<T, 2, FineCoef<T, 2>><<<nb, 128, 0, stream>>>(A, Lf.n[0], Lf.n[1], 1, Lc.n[0], Lc.n[1], 1, Lc.p0, Lc.N, Lc.coef.p, 0, 1, 0, 1, 0, 1);
                    }
                } else {
                    if (pb.dim == 3) {
                        StoredCoef<T, 3> A{coarse_op(Lf)};
                        k_galerkin<T, 3, StoredCoef<T, 3>><<<nb, 128, 0, stream>>>(A, Lf.n[0], Lf.n[1], Lf.n[2], Lc.n[0], Lc.n[1], Lc.n[2], Lc.p0, Lc.N, Lc.coef.p,
                                                                                   Lc.zb, Lc.ze, Lf.koff, Lf.n2g, Lc.koff, Lc.n2g);
                    } else {
                        StoredCoef<T, 2> A{coarse_op(Lf)};
                        k_galerkin<T, 2, StoredCoef<T, 2>><<<nb, 128, 0, stream>>>(A, Lf.n[0], Lf.n[1], 1, Lc.n[0], Lc.n[1], 1, Lc.p0, Lc.N, Lc.coef.p, 0, 1, 0, 1, 0, 1);
                    }
                }
            });
            launch(T_SETUP, 0, [&] {
                k_coarse_dinv<T><<<(unsigned)((Lc.N + 255) / 256), 256, 0, stream>>>(Lc.coef.p + (int64_t)center * Lc.N, Lc.dinv.p, Lc.N, (T)o.relax_param);
            });
        }
        // levels whose Jacobi-preconditioned GMRES (smoother or inexact coarsest solve) runs on a stored stencil
        for (int l = (ho ? 0 : 1); l < o.levels && scaled_gmres; ++l) {
            const bool gm = (o.relax_type == HH_RELAX_JAC_GMRES && l < o.levels - 1) ||
                            (l == o.levels - 1 && o.coarse_type == HH_COARSE_GMRES);
            if (!gm) continue;
            Level& L = levels[l];
            if (slab) halo_exchange(l, L.dinv.p, 1, 0, HALO_BOTH);  // columns on the halo planes are scaled by the neighbour's dinv
            L.scoef.alloc((size_t)NS * L.N);
            HH_CUDA(cudaMemsetAsync(L.scoef.p, 0, (size_t)NS * L.N * sizeof(C), stream));
            const int64_t tot = (int64_t)L.n[0] * L.n[1] * L.n[2] * NS;
            CoarseOp<T> op = coarse_op(L);
            launch(T_SETUP, 0, [&] {
                if (pb.dim == 3) k_scale_columns<T, 3><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(op, L.scoef.p);
                else k_scale_columns<T, 2><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(op, L.scoef.p);
            });
        }
        if (!ho && (o.relax_type == HH_RELAX_JAC_GMRES || (o.levels == 1 && o.coarse_type == HH_COARSE_GMRES))) {
            Level& L0 = levels[0];
            L0.dinv.alloc(L0.N);
            HH_CUDA(cudaMemsetAsync(L0.dinv.p, 0, (size_t)L0.N * sizeof(C), stream));
            dim3 g, blk;
            grid3(pb.n, pb.dim, g, blk);
            launch(T_SETUP, 0, [&] {
                if (pb.dim == 3) k_fine_dinv<T, 3><<<g, blk, 0, stream>>>(mg_fine, L0.dinv.p, (T)o.relax_param);
                else k_fine_dinv<T, 2><<<g, blk, 0, stream>>>(mg_fine, L0.dinv.p, (T)o.relax_param);
            });
        }
        if (o.coarse_type == HH_COARSE_LU) build_dense_inverse();
        HH_CUDA(cudaStreamSynchronize(stream));
        have_hierarchy = true;
        setup_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    // exact coarsest solve: banded LU (device) -> explicit inverse (device), row-major in precision T
    void build_dense_inverse() {
        const int Lc = opt.levels - 1;
        HH_REQUIRE(Lc >= 1, HH_ERR_UNSUPPORTED,
                   "coarse_type LU needs levels >= 2 (the fine level is matrix-free; use GMRES for a 1-level solve)");
        Level& L = levels[Lc];
        const int64_t N = L.Nlog;  // == L.N: this level is never padded
        const int bw = pb.dim == 3 ? (1 + L.n[0] + L.n[0] * L.n[1]) : (1 + L.n[0]);
        const double need = (double)N * N * (16.0 + sizeof(C)) + (double)N * (2.0 * bw + 1) * 16.0;
        size_t fr = 0, tot = 0;
        HH_CUDA(cudaMemGetInfo(&fr, &tot));
        HH_REQUIRE(need < 0.5 * (double)fr && N <= 46000, HH_ERR_UNSUPPORTED,
                   "coarsest grid too large for the exact (LU) coarsest solve; use more levels or coarse_type GMRES");
        const int64_t W = 2 * (int64_t)bw + 1;
        DevBuf<zc> band, inv;
        band.alloc((size_t)N * W);
        inv.alloc((size_t)N * N);
        HH_CUDA(cudaMemsetAsync(band.p, 0, (size_t)N * W * sizeof(zc), stream));
        const int NS = pb.dim == 3 ? 27 : 9;
        const unsigned nb = (unsigned)((N * NS + 255) / 256);
        CoarseOp<T> op = coarse_op(L);
        launch(T_SETUP, 0, [&] {
            if (pb.dim == 3) k_band_fill<T, 3><<<nb, 256, 0, stream>>>(op, band.p, bw);
            else k_band_fill<T, 2><<<nb, 256, 0, stream>>>(op, band.p, bw);
        });
        DevBuf<int> lu_flag;
        lu_flag.alloc(1);
        HH_CUDA(cudaMemsetAsync(lu_flag.p, 0, sizeof(int), stream));
        DevBuf<double> diag0;
        diag0.alloc((size_t)N);
        launch(T_SETUP, 0, [&] { k_band_lu<<<1, 1024, 0, stream>>>(band.p, N, bw, lu_flag.p, diag0.p); });
        {
            int bad = 0;
            HH_CUDA(cudaMemcpyAsync(&bad, lu_flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
            HH_CUDA(cudaStreamSynchronize(stream));
            HH_REQUIRE(bad == 0, HH_ERR_UNSUPPORTED,
                       "exact coarsest solve: the LU factorisation (no pivoting) met a vanishing pivot at row " + std::to_string(bad - 1) +
                           "; the coarse operator is (nearly) singular or indefinite (shift = 0 without attenuation?) -- use a positive "
                           "shift or coarse_type GMRES");
        }
        for (int64_t c0 = 0; c0 < N; c0 += 32768) {
            const unsigned nc = (unsigned)std::min<int64_t>(32768, N - c0);
            launch(T_SETUP, 0, [&] { k_band_inverse<<<nc, 256, 0, stream>>>(band.p, N, bw, c0, inv.p); });
        }
        inv_dense.alloc((size_t)N * N);
        dim3 g((unsigned)((N + 31) / 32), (unsigned)((N + 31) / 32)), blk(32, 8);
        launch(T_SETUP, 0, [&] { k_inverse_pack<T><<<g, blk, 0, stream>>>(inv.p, inv_dense.p, N); });
        HH_CUDA(cudaStreamSynchronize(stream));
    }

    void get_level_stencil(int level, void* out) override {
        HH_REQUIRE(have_hierarchy && level >= (ho && !krylov_only ? 0 : 1) && level < (int)levels.size(), HH_ERR_ARG, "bad level");
        HH_CUDA(cudaSetDevice(device));
        HH_CUDA(cudaStreamSynchronize(stream));
        const Level& L = levels[level];
        if (L.N == L.Nlog) {
            HH_CUDA(cudaMemcpy(out, L.coef.p, L.coef.n * sizeof(C), cudaMemcpyDeviceToHost));
            return;
        }
        // padded level: strip the ghost column on the host
        std::vector<C> tmp(L.coef.n);
        HH_CUDA(cudaMemcpy(tmp.data(), L.coef.p, L.coef.n * sizeof(C), cudaMemcpyDeviceToHost));
        C* o = (C*)out;
        const int NS = pb.dim == 3 ? 27 : 9;
        const int64_t rows = (int64_t)L.n[1] * L.n[2];
        for (int s = 0; s < NS; ++s)
            for (int64_t r = 0; r < rows; ++r)
                std::memcpy(o + (int64_t)s * L.Nlog + r * L.n[0], tmp.data() + (int64_t)s * L.N + r * L.p0, L.n[0] * sizeof(C));
    }
    void get_diagonal(int shifted, double shift, double* out) override {
        HH_CUDA(cudaSetDevice(device));
        DevBuf<zc> d;
        d.alloc(pb.N());
        FineOp<T> op = fine_op(shifted ? shift : 0.0, 0);
        dim3 g, blk;
        grid3(pb.n, pb.dim, g, blk);
        launch(T_SETUP, 0, [&] {
            if (pb.dim == 3) k_fine_diag<T, 3><<<g, blk, 0, stream>>>(op, d.p);
            else k_fine_diag<T, 2><<<g, blk, 0, stream>>>(op, d.p);
        });
        HH_CUDA(cudaStreamSynchronize(stream));
        HH_CUDA(cudaMemcpy(out, d.p, pb.N() * sizeof(zc), cudaMemcpyDeviceToHost));
    }

    // ------------------------------------------------------------------ work memory (adjustMemoryForNumRHS)
    int gmres_small_steps(int l) const {
        int s = 0;
        const int L = opt.levels;
        if (opt.relax_type == HH_RELAX_JAC_GMRES && l < L - 1) s = std::max(s, std::max(opt.relax_pre[l], opt.relax_post[l]));
        if (l == L - 1 && opt.coarse_type == HH_COARSE_GMRES) s = std::max(s, opt.coarse_iters);
        return s;
    }
    bool kcycle_level(int l) const { return opt.cycle_type == HH_CYCLE_K && l >= 1 && l < opt.levels - 1; }
    double level_bytes_per_rhs() const {
        double b = 0;
        if (krylov_only) return 0.0;
        for (int l = 0; l < (int)levels.size(); ++l) {
            const double N = (double)levels[l].N;
            b += (l == 0 ? 1.0 : 3.0) * N * S;
            const int gs = gmres_small_steps(l);
            if (gs > 0) b += (gs + 1) * N * S;
            if (kcycle_level(l)) b += 5 * N * S;
        }
        return b;
    }
    void ensure_level_memory(int nrhs) {
        if (nrhs <= kcap) return;
        if (krylov_only) {
            kcap = nrhs;
            d_partial.alloc((size_t)(HH_MAXV + 1) * std::max(nrhs, 1) * (148 * 8 + 8));
            return;
        }
        for (int l = 0; l < (int)levels.size(); ++l) {
            Level& L = levels[l];
            alloc_zero(L.t, (size_t)L.N * nrhs);
            if (l >= 1) {
                alloc_zero(L.x, (size_t)L.N * nrhs);
                alloc_zero(L.b, (size_t)L.N * nrhs);
            }
            L.px = L.x.p;
            L.pb = L.b.p;
            L.pt = L.t.p;
            const int gs = gmres_small_steps(l);
            L.gs.steps = gs;
            if (gs > 0) {
                alloc_zero(L.gs.v, (size_t)(gs + 1) * L.N * nrhs);
                alloc_gmres_state(L.gs.g, gs, nrhs);
                L.gs.np.alloc((size_t)std::max(nrhs, 1) * (148 * 8 + 8));
            }
            L.ks.steps = kcycle_level(l) ? 2 : 0;
            if (L.ks.steps) {
                alloc_zero(L.ks.v, (size_t)3 * L.N * nrhs);
                alloc_zero(L.ks.z, (size_t)2 * L.N * nrhs);
                alloc_gmres_state(L.ks.g, 2, nrhs);
                L.ks.np.alloc((size_t)std::max(nrhs, 1) * (148 * 8 + 8));
            }
        }
        kcap = nrhs;
        const int nblk_max = 148 * 8 + 8;
        d_partial.alloc((size_t)(HH_MAXV + 1) * std::max(nrhs, 1) * nblk_max);
    }

    void alloc_zero(DevBuf<C>& b, size_t count) {
        b.alloc(count);
        HH_CUDA(cudaMemsetAsync(b.p, 0, count * sizeof(C), stream));
    }
    void alloc_gmres_state(GmresMem& g, int m, int nrhs) {
        g.H.alloc((size_t)nrhs * (m + 1) * m);
        g.cs.alloc((size_t)nrhs * m);
        g.sn.alloc((size_t)nrhs * m);
        g.s.alloc((size_t)nrhs * (m + 1));
        g.hcol.alloc((size_t)nrhs * (m + 1));
        g.y.alloc((size_t)nrhs * m);
        g.scale.alloc(nrhs);
        g.bnorm.alloc(nrhs);
        g.err.alloc(nrhs);
        g.done.alloc(nrhs);
        g.jdone.alloc(nrhs);
        g.nprec.alloc(nrhs);
        g.d.alloc((size_t)nrhs * (m + 1));
        g.acc.alloc((size_t)nrhs * 2);
        g.st = GmresState{g.H.p, g.cs.p, g.sn.p, g.s.p, g.hcol.p, g.y.p, g.bnorm.p, g.err.p, g.scale.p, g.done.p, g.jdone.p, g.nprec.p, m, g.d.p, g.acc.p};
    }

    // ------------------------------------------------------------------ generic operator access per level
    void level_apply(int l, int mode, const C* x, const C* b, C* out, int nrhs) {
        if (l == 0 && ho) coarse_stencil(mode, levels[0], x, b, out, nrhs);
        else if (l == 0) fine_stencil(mode, mg_fine, x, b, out, levels[0].N, nrhs, (T)opt.relax_param);
        else coarse_stencil(mode, levels[l], x, b, out, nrhs);
    }
    void level_jacobi0(int l, const C* b, C* out, int nrhs) {
        if (l == 0 && ho) diag_scale(T_COARSE_JACOBI0, levels[0].dinv.p, b, out, levels[0].N, nrhs);
        else if (l == 0) fine_jacobi0(mg_fine, b, out, levels[0].N, nrhs, (T)opt.relax_param);
        else diag_scale(T_COARSE_JACOBI0, levels[l].dinv.p, b, out, levels[l].N, nrhs);
    }

    // Orthogonalise w against V_0..V_j (classical Gram-Schmidt in one fused pass over the vectors),
    // then the Givens update.  Leaves 1/||w|| in st.scale.
    // V[0..j] are the stored (scaled) basis vectors, w = A z~_j on entry, the next basis vector on exit.
    // last_column: w is the last column of its cycle -- v_{j+1} would never be read, so the update pass over the vectors
    // is skipped and h_{j+1,j} comes from the dot pass alone (gmres_givens_dev, est).
    void gmres_orthogonalise(GmresMem& g, const C* const* V, int j, C* w, const Span& N, int nrhs, double tol,
                             bool last_column = false) {
        const C* vv[HH_MAXV];
        for (int i0 = 0; i0 <= j; i0 += HH_MAXV) {
            const int nv = std::min(HH_MAXV, j + 1 - i0);
            for (int i = 0; i < nv; ++i) vv[i] = V[i0 + i];
            const int nblk = multidot(vv, nv, w, N, nrhs, i0 == 0, d_partial.p);
            launch(T_SCALAR, 0, [&] {
                if (scalar_fast) k_gmres_hcol_mw<<<nrhs, 32 * (nv + (i0 == 0 ? 1 : 0)), 0, stream>>>(g.st, red_out, nblk, j, i0, nv);
                else k_gmres_hcol<<<nrhs, 32, 0, stream>>>(g.st, red_out, nblk, j, i0, nv);
            });
        }
        if (last_column && skip_last_update) {
            launch(T_SCALAR, 0, [&] {
                if (scalar_fast) k_gmres_givens_mw<<<nrhs, 32, 0, stream>>>(g.st, nullptr, 0, j, tol, 1);
                else k_gmres_givens<<<nrhs, 32, 0, stream>>>(g.st, nullptr, 0, j, tol, 1);
            });
            return;
        }
        int nblk = 0;
        for (int i0 = 0; i0 <= j; i0 += HH_MAXV) {
            const int nv = std::min(HH_MAXV, j + 1 - i0);
            for (int i = 0; i < nv; ++i) vv[i] = V[i0 + i];
            const bool last = (i0 + HH_MAXV > j);
            nblk = multiaxpy(vv, nv, w, N, nrhs, g.hcol.p + i0, g.st.m + 1, true, last, d_partial.p,
                             last ? g.scale.p : nullptr);
        }
        launch(T_SCALAR, 0, [&] {
            if (scalar_fast) k_gmres_givens_mw<<<nrhs, 32, 0, stream>>>(g.st, red_out, nblk, j, tol, 0);
            else k_gmres_givens<<<nrhs, 32, 0, stream>>>(g.st, red_out, nblk, j, tol, 0);
        });
    }
    // One step of a level's fixed-length GMRES with ONE scalar kernel and (slabs) ONE all-reduce: the Givens update of
    // column j-1 waits for the dot sums of column j (k_gmres_small_step); `pending` = nblk of the update pass of column
    // j-1 whose norm partials sit in ws.np (0: none).  The last column skips its update pass.
    void gmres_step_fused(SmallWs& ws, const C* const* V, int j, C* w, const Span& N, int nrhs, bool last_column, int& pending) {
        GmresMem& g = ws.g;
        const int nv = j + 1;
        const int nblk = multidot_launch(V, nv, w, N, nrhs, true, d_partial.p);
        const zc* dots = d_partial.p;
        const zc* norms = ws.np.p;
        int nb_d = nblk, nb_n = pending;
        if (slab && slab->nranks > 1) {
            const int nq = (nv + 1) * nrhs, nq2 = pending ? nrhs : 0;
            if (d_red.n < (size_t)(nq + nq2)) d_red.alloc((size_t)std::max(nq + nq2, 4096));
            launch(T_SCALAR, 0, [&] {
                k_sum_partials<<<(nq + 7) / 8, 256, 0, stream>>>(d_partial.p, nq, nblk, d_red.p);
                if (nq2) k_sum_partials<<<(nq2 + 7) / 8, 256, 0, stream>>>(ws.np.p, nq2, pending, d_red.p + nq);
            });
            launch(T_ALLREDUCE, 16.0 * (nq + nq2), [&] { slab->allreduce(stream, (double*)d_red.p, 2 * (nq + nq2), false); });
            dots = d_red.p;
            norms = d_red.p + nq;
            nb_d = nb_n = 1;
        }
        const int flags = (pending ? 1 : 0) | (last_column ? 2 : 0);
        launch(T_SCALAR, 0, [&] {
            if (scalar_fast) k_gmres_small_step_mw<<<nrhs, 32 * (j + 3), 0, stream>>>(g.st, dots, nb_d, norms, nb_n, j, flags);
            else k_gmres_small_step<<<nrhs, 32, 0, stream>>>(g.st, dots, nb_d, norms, nb_n, j, flags);
        });
        pending = 0;
        if (!last_column)
            pending = multiaxpy_launch(V, nv, w, N, nrhs, g.hcol.p, g.st.m + 1, true, true, ws.np.p, g.scale.p);
    }
    void gmres_begin(GmresMem& g, const C* r, const Span& N, int nrhs, bool first, double tol) {
        const int nblk = multidot(nullptr, 0, r, N, nrhs, true, d_partial.p);
        launch(T_SCALAR, 0, [&] { k_gmres_begin<<<nrhs, 32, 0, stream>>>(g.st, red_out, nblk, first ? 1 : 0, tol); });
    }
    void gmres_solve_y(GmresMem& g, int nrhs) {
        launch(T_SCALAR, 0, [&] { k_gmres_solve_y<<<(nrhs + 63) / 64, 64, 0, stream>>>(g.st, nrhs); });
    }

    // `nsteps` steps of GMRES on level l (one cycle, no convergence test), right-preconditioned by
    //   prec = 0: damped-Jacobi diagonal (Jac-GMRES smoother, inexact coarsest solve)
    //   prec = 1: one recursive multigrid cycle on level l (K-cycle)
    // x is updated in place (x_is_zero: x need not be initialised).
    void small_gmres(int l, int nsteps, int prec, const C* b, C* x, bool x_is_zero, int nrhs) {
        Level& L = levels[l];
        SmallWs& ws = prec == 0 ? L.gs : L.ks;
        GmresMem& g = ws.g;
        HH_REQUIRE(nsteps >= 1 && nsteps <= ws.steps, HH_ERR_STATE, "small_gmres workspace");
        const int64_t N = L.N;
        const Span sp = span(l);
        const int64_t vs = N * kcap;
        std::vector<C*> W(nsteps + 1);            // writable slots
        std::vector<const C*> V(nsteps + 1);      // the basis as read
        for (int i = 0; i <= nsteps; ++i) V[i] = W[i] = ws.v.p + (int64_t)i * vs;
        if (x_is_zero) V[0] = b;  // r0 = b: used in place, unscaled (d_0 = ||b||)
        else level_apply(l, MODE_RESID, x, b, W[0], nrhs);
        gmres_begin(g, V[0], sp, nrhs, true, 0.0);
        const bool fused = small_fused && nsteps <= HH_MAXV && ws.np.p != nullptr;
        int pending = 0;
        auto orthogonalise = [&](int j) {
            if (fused) gmres_step_fused(ws, V.data(), j, W[j + 1], sp, nrhs, j == nsteps - 1, pending);
            else gmres_orthogonalise(g, V.data(), j, W[j + 1], sp, nrhs, 0.0, j == nsteps - 1);
        };
        for (int j = 0; j < nsteps; ++j) {
            C* z;
            if (prec == 0 && L.scoef.p != nullptr) {
                // Jacobi, stored stencil: w = (A D^-1) v_j in one pass over the column-scaled coefficients
                coarse_stencil(MODE_APPLY, L, V[j], nullptr, W[j + 1], nrhs, true);
                orthogonalise(j);
                continue;
            }
            if (prec == 0) {
                z = L.pt;  // Jacobi: z = dinv .* v_j (not stored: x += dinv .* (V y) at the end)
                diag_scale(l == 0 ? T_FINE_JACOBI0 : T_COARSE_JACOBI0, L.dinv.p, V[j], z, N, nrhs);
            } else {
                z = ws.z.p + (int64_t)j * vs;
                cycle(l, V[j], z, true, nrhs);
            }
            level_apply(l, MODE_APPLY, z, nullptr, W[j + 1], nrhs);
            orthogonalise(j);
        }
        gmres_solve_y(g, nrhs);
        const C* vv[HH_MAXV];
        if (prec == 0 && fused) {
            // x (+)= dinv .* (V y) in one pass
            for (int i = 0; i < nsteps; ++i) vv[i] = V[i];
            combine(vv, nsteps, L.dinv.p, x, sp, nrhs, g.y.p, g.st.m, !x_is_zero, T_AXPY);
        } else if (prec == 0) {
            // t = sum_j y_j v_j  (accumulated into the last slot, free now), then x (+)= dinv .* t
            C* t = W[nsteps];
            zero_vec(t, N, nrhs);
            for (int i0 = 0; i0 < nsteps; i0 += HH_MAXV) {
                const int nv = std::min(HH_MAXV, nsteps - i0);
                for (int i = 0; i < nv; ++i) vv[i] = V[i0 + i];
                multiaxpy(vv, nv, t, sp, nrhs, g.y.p + i0, g.st.m, false, false, d_partial.p);
            }
            if (x_is_zero) {
                diag_scale(l == 0 ? T_FINE_JACOBI0 : T_COARSE_JACOBI0, L.dinv.p, t, x, N, nrhs);
            } else {
                diag_scale(l == 0 ? T_FINE_JACOBI0 : T_COARSE_JACOBI0, L.dinv.p, t, t, N, nrhs);
                const C* one[1] = {t};
                multiaxpy(one, 1, x, sp, nrhs, d_one.p, 0, false, false, d_partial.p);
            }
        } else {
            if (x_is_zero) zero_vec(x, N, nrhs);
            for (int i = 0; i < nsteps; ++i) vv[i] = ws.z.p + (int64_t)i * vs;
            multiaxpy(vv, nsteps, x, sp, nrhs, g.y.p, g.st.m, false, false, d_partial.p);
        }
    }

    // ------------------------------------------------------------------ multigrid cycle (recursiveCycle)
    // Jacobi sweeps on level l: result ends in `x`.  `t` is scratch of the same size.
    void smooth(int l, int nsweeps, const C* b, C*& x, C*& t, bool x_is_zero, bool can_swap, int nrhs) {
        if (nsweeps <= 0) {
            if (x_is_zero) zero_vec(x, levels[l].N, nrhs);
            return;
        }
        if (opt.relax_type == HH_RELAX_JAC_GMRES) {
            small_gmres(l, nsweeps, 0, b, x, x_is_zero, nrhs);
            return;
        }
        // ping-pong so that the final sweep writes x
        C* cur = x;  // holds the current iterate (if !x_is_zero)
        C* oth = t;
        for (int s = 0; s < nsweeps; ++s) {
            const int remaining = nsweeps - s;  // including this sweep
            if (s == 0 && x_is_zero) {
                // writes either buffer: choose so that parity ends in x
                C* dst = (remaining % 2 == 1) ? x : t;
                level_jacobi0(l, b, dst, nrhs);
                cur = dst;
                oth = (dst == x) ? t : x;
            } else {
                level_apply(l, MODE_JACOBI, cur, b, oth, nrhs);
                std::swap(cur, oth);
            }
        }
        if (cur != x) {
            if (can_swap) std::swap(x, t);
            else copy_vec(cur, x, levels[l].N, nrhs);
        }
    }

    void coarsest_solve(int l, const C* b, C* x, int nrhs) {
        Level& L = levels[l];
        if (opt.coarse_type == HH_COARSE_LU) {
            const int wpb = 8;
            launch(T_COARSEST_DENSE, S * ((double)L.N * L.N * ((nrhs + 3) / 4) + 2.0 * L.N * nrhs), [&] {
                k_dense_apply<T, 4><<<(unsigned)((L.N + wpb - 1) / wpb), wpb * 32, 0, stream>>>(inv_dense.p, b, x, L.N, L.N, nrhs);
            });
        } else {
            small_gmres(l, opt.coarse_iters, 0, b, x, true, nrhs);
        }
    }

    // one cycle on level l for A_l x = b; x is caller storage (level-owned for l >= 1).
    void cycle(int l, const C* b, C* x, bool x_is_zero, int nrhs) {
        const int Lmax = opt.levels - 1;
        if (l == Lmax) {
            coarsest_solve(l, b, x, nrhs);
            return;
        }
        Level& F = levels[l];
        Level& Cc = levels[l + 1];
        C* xx = x;
        C* tt = F.pt;
        const int npre = opt.relax_pre[l];
        const int npost = opt.relax_post[l];
        // Fused correction + two post-smoothing sweeps (k_fine3d_tma_pro2) reads the iterate from one buffer and writes the
        // result to another, and the result must land in the caller's x: the pre-smoothed iterate then lives in the
        // level's scratch vector and the residual (dead after the restriction) borrows x.
        bool post2 = false;
        // One pre-smoothing sweep from zero: x1 = dinv .* b is cheap to recompute, so the cycle start writes only the
        // residual and the correction + first post-sweep pass forms x1 from b again (k_fine3d_tma_prob).
        bool recompute = false;
        // Levels >= 1 own both their iterate and their scratch vector: after an odd number of ping-pong sweeps the two
        // swap roles (the callers read Level::px after the call) instead of a copy back.
        const bool swap_ok = level_swap && l >= 1 && x == F.px;
        if (l == 0 && x_is_zero && opt.relax_type == HH_RELAX_JAC && npre >= 1 && can_fuse_first(mg_fine, b, F.N)) {
            post2 = npost >= 2 && (npost % 2) == 0 && can_fuse_post2(mg_fine, tt, b, Cc, Cc.px, F.N) && ((uintptr_t)xx % 16 == 0);
            recompute = !post2 && npre == 1 && npost >= 1 && can_fuse_prolong_b(mg_fine, b, Cc, Cc.px, F.N);
            C* itb = post2 ? tt : xx;  // the iterate after pre-smoothing
            C* rsb = post2 ? xx : tt;  // the residual
            if (npre == 1) {
                fine_first(0, mg_fine, b, recompute ? nullptr : itb, rsb, F.N, nrhs);  // [x1 and] r = b - A x1 in one pass
            } else {
                // first two sweeps in one pass, written so that the remaining npre-2 ping-pong sweeps end in itb
                C* cur = ((npre - 2) % 2 == 0) ? itb : rsb;
                C* oth = (cur == itb) ? rsb : itb;
                fine_first(1, mg_fine, b, cur, nullptr, F.N, nrhs);
                for (int sw = 2; sw < npre; ++sw) {
                    level_apply(l, MODE_JACOBI, cur, b, oth, nrhs);
                    std::swap(cur, oth);
                }
                level_apply(l, MODE_RESID, itb, b, rsb, nrhs);
            }
            if (post2) std::swap(xx, tt);  // from here on: xx = iterate (the scratch vector), tt = residual (the caller's x)
        } else {
            smooth(l, npre, b, xx, tt, x_is_zero, swap_ok, nrhs);
            level_apply(l, MODE_RESID, xx, b, tt, nrhs);
        }
        restrict_to(F, Cc, tt, Cc.pb, nrhs);
        if (l + 1 == Lmax) {
            coarsest_solve(l + 1, Cc.pb, Cc.px, nrhs);
        } else if (opt.cycle_type == HH_CYCLE_V) {
            cycle(l + 1, Cc.pb, Cc.px, true, nrhs);
        } else if (opt.cycle_type == HH_CYCLE_W) {
            cycle(l + 1, Cc.pb, Cc.px, true, nrhs);
            cycle(l + 1, Cc.pb, Cc.px, false, nrhs);
        } else {
            small_gmres(l + 1, 2, 1, Cc.pb, Cc.px, true, nrhs);
        }
        if (post2) {
            // x2 = J(J(x + P xc)) in one pass, into the caller's x (tt here); further pairs of sweeps ping-pong back to it
            C* cur = tt;
            C* oth = xx;
            fine_prolong_jacobi2(mg_fine, xx, b, Cc, Cc.px, tt, F.N, nrhs);
            for (int sw = 2; sw < npost; ++sw) {
                level_apply(l, MODE_JACOBI, cur, b, oth, nrhs);
                std::swap(cur, oth);
            }
            return;  // npost is even: the last sweep wrote the caller's x
        }
        if (recompute) {
            // x' = dinv .* b + P xc and the first post-smoothing sweep in one pass; the residual buffer is free now, and
            // the first target is chosen so that the last sweep writes the caller's x
            C* cur = ((npost - 1) % 2 == 0) ? xx : tt;
            C* oth = (cur == xx) ? tt : xx;
            fine_prolong_jacobi_b(mg_fine, b, Cc, Cc.px, cur, F.N, nrhs);
            for (int sw = 1; sw < npost; ++sw) {
                level_apply(l, MODE_JACOBI, cur, b, oth, nrhs);
                std::swap(cur, oth);
            }
            return;
        }
        if (l == 0 && opt.relax_type == HH_RELAX_JAC && npost >= 1 && can_fuse_prolong(mg_fine, xx, b, Cc, Cc.px, F.N)) {
            // x' = x + P xc and the first post-smoothing sweep in one pass (x' never touches HBM)
            C* cur = tt;
            C* oth = xx;
            fine_prolong_jacobi(mg_fine, xx, b, Cc, Cc.px, tt, F.N, nrhs);
            for (int sw = 1; sw < npost; ++sw) {
                level_apply(l, MODE_JACOBI, cur, b, oth, nrhs);
                std::swap(cur, oth);
            }
            if (cur != xx) copy_vec(cur, xx, F.N, nrhs);
        } else {
            prolong_add(F, Cc, xx, Cc.px, nrhs);
            smooth(l, npost, b, xx, tt, false, swap_ok, nrhs);
        }
        if (swap_ok) {  // the iterate may have ended in the scratch vector: the two buffers swap roles
            F.px = xx;
            F.pt = tt;
        }
    }

    void precondition(const C* b, C* z, int nrhs) {
        if (prec_hook) prec_hook(b, z, nrhs);  // mixed precision: the cycle of the ComplexF32 companion solver
        else cycle(0, b, z, true, nrhs);
        ++n_prec;
    }
    void precondition_internal(const void* b, void* z, int nrhs) override {
        HH_REQUIRE(have_hierarchy && !krylov_only, HH_ERR_STATE, "no hierarchy");
        cycle(0, (const C*)b, (C*)z, true, nrhs);
    }
    void ensure_cycle_memory(int nrhs) override {
        ensure_level_memory(nrhs);
        ensure_const();
    }
    int64_t internal_ld() const override { return have_hierarchy ? levels[0].N : fineN(); }
    int internal_pitch() const override { return fine_sy(); }
    double per_rhs_bytes(const hh_solve_options& o) override {
        const int64_t Nf = have_hierarchy ? levels[0].N : pb.N();
        return level_bytes_per_rhs() + (double)(krylov_vectors(o) + (padded() ? 2 : 0)) * Nf * S;
    }
    double cycle_bytes_per_rhs() const override { return level_bytes_per_rhs(); }
    double held_bytes() const override { return (double)kcap * level_bytes_per_rhs() + (double)kry.n * sizeof(C); }

    void cycle_device(const void* dB, void* dZ, int64_t nrhs) override {
        HH_CUDA(cudaSetDevice(device));
        HH_REQUIRE(have_hierarchy, HH_ERR_STATE, "hh_setup has not been called");
        ensure_level_memory((int)nrhs);
        ensure_const();
        if (padded()) {
            DevBuf<C> bp, zp;
            alloc_zero(bp, (size_t)levels[0].N * nrhs);
            alloc_zero(zp, (size_t)levels[0].N * nrhs);
            repitch((const C*)dB, bp.p, (int)nrhs, true);
            precondition(bp.p, zp.p, (int)nrhs);
            repitch(zp.p, (C*)dZ, (int)nrhs, false);
            HH_CUDA(cudaStreamSynchronize(stream));
            return;
        }
        precondition((const C*)dB, (C*)dZ, (int)nrhs);
        HH_CUDA(cudaStreamSynchronize(stream));
    }
    // internal fine-level vectors are padded (ComplexF32 on an odd grid with the TMA kernels)?
    // (slab decomposition: caller blocks hold the owned planes only, internal ones the halo planes as well)
    bool padded() const { return have_hierarchy && (levels[0].N != pb.N() || slab); }
    // dense caller block (leading dimension caller_N()) <-> internal block (leading dimension levels[0].N, row pitch
    // p0, owned planes starting at plane zb)
    void repitch(const C* src, C* dst, int nrhs, bool to_padded) {
        const Level& L0 = levels[0];
        const int64_t rows = (int64_t)pb.n[1] * (L0.ze - L0.zb);
        const int n0 = pb.n[0], p0 = L0.p0;
        const int64_t ioff = (int64_t)L0.zb * p0 * pb.n[1];
        const int64_t cN = caller_N();
        dim3 g(std::max(1, 592 / std::max(nrhs, 1)), nrhs);
        launch(T_COPY, 2 * S * (double)cN * nrhs, [&] {
            if (to_padded) k_repitch<C><<<g, 256, 0, stream>>>(src, dst + ioff, n0, rows, n0, p0, cN, L0.N);
            else k_repitch<C><<<g, 256, 0, stream>>>(src + ioff, dst, n0, rows, p0, n0, L0.N, cN);
        });
    }

    void apply_device(const void* dX, void* dY, int64_t nrhs, int shifted, double shift, int transpose) override {
        HH_CUDA(cudaSetDevice(device));
        HH_REQUIRE(d_m.p != nullptr, HH_ERR_STATE, "no model set");
        if (ho) {  // stored stencils exist for the un-shifted operator and for the hierarchy's shift only
            HH_REQUIRE(have_hierarchy, HH_ERR_STATE, "high-order operator: hh_apply needs hh_setup first");
            HH_REQUIRE((transpose != 0) == (opt.do_transpose != 0), HH_ERR_UNSUPPORTED,
                       "high-order operator: hh_apply applies the operator the hierarchy was built for (do_transpose of hh_setup)");
            HH_REQUIRE(!shifted || shift == 0.0 || (!krylov_only && shift == opt.shift[0]), HH_ERR_UNSUPPORTED,
                       "high-order operator: hh_apply supports shift 0 and the shift given to hh_setup");
            const Level& Lop = (shifted && shift != 0.0) ? levels[0] : hoH;
            if (!padded()) {
                coarse_stencil(MODE_APPLY, Lop, (const C*)dX, nullptr, (C*)dY, (int)nrhs);
                HH_CUDA(cudaStreamSynchronize(stream));
                return;
            }
            DevBuf<C> xi, yi;
            alloc_zero(xi, (size_t)levels[0].N * nrhs);
            alloc_zero(yi, (size_t)levels[0].N * nrhs);
            repitch((const C*)dX, xi.p, (int)nrhs, true);
            coarse_stencil(MODE_APPLY, Lop, xi.p, nullptr, yi.p, (int)nrhs);
            repitch(yi.p, (C*)dY, (int)nrhs, false);
            HH_CUDA(cudaStreamSynchronize(stream));
            return;
        }
        if (slab) {  // caller blocks hold the owned planes: go through blocks with halo planes
            HH_REQUIRE(have_hierarchy, HH_ERR_STATE, "slab decomposition: hh_apply needs hh_setup first");
            FineOp<T> op = fine_op(shifted ? shift : 0.0, transpose, true);
            DevBuf<C> xi, yi;
            alloc_zero(xi, (size_t)levels[0].N * nrhs);
            alloc_zero(yi, (size_t)levels[0].N * nrhs);
            repitch((const C*)dX, xi.p, (int)nrhs, true);
            fine_stencil(MODE_APPLY, op, xi.p, nullptr, yi.p, levels[0].N, (int)nrhs, T(0));
            repitch(yi.p, (C*)dY, (int)nrhs, false);
            HH_CUDA(cudaStreamSynchronize(stream));
            return;
        }
        FineOp<T> op = fine_op(shifted ? shift : 0.0, transpose);
        fine_stencil(MODE_APPLY, op, (const C*)dX, nullptr, (C*)dY, pb.N(), (int)nrhs, T(0));
    }

    void scatter_point_sources(void* dB, const int64_t* idx0, const double* val, int64_t nrhs) override {
        HH_CUDA(cudaSetDevice(device));
        DevBuf<int64_t> di;
        DevBuf<zc> dv;
        di.alloc(nrhs);
        dv.alloc(nrhs);
        std::vector<int64_t> loc(idx0, idx0 + nrhs);
        if (slab) {  // global node index -> index inside the owned planes (or -1: the source lies in another slab)
            const int64_t plane = (int64_t)pb.n[0] * pb.n[1];
            for (auto& v : loc) v = (v >= sgeo[0].own0 * plane && v < sgeo[0].own1 * plane) ? v - sgeo[0].own0 * plane : -1;
        }
        const int64_t cN = caller_N();
        HH_CUDA(cudaMemcpyAsync(di.p, loc.data(), nrhs * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
        HH_CUDA(cudaMemcpyAsync(dv.p, val, nrhs * sizeof(zc), cudaMemcpyHostToDevice, stream));
        HH_CUDA(cudaMemsetAsync(dB, 0, (size_t)cN * nrhs * sizeof(C), stream));
        launch(T_COPY, 0, [&] { k_point_sources<T><<<(unsigned)((nrhs + 127) / 128), 128, 0, stream>>>((C*)dB, cN, di.p, dv.p, (int)nrhs); });
        HH_CUDA(cudaStreamSynchronize(stream));
    }

    void ensure_const() {
        if (d_one.p) return;
        d_one.alloc(1);
        zc one = mk<double>(1.0, 0.0);
        HH_CUDA(cudaMemcpy(d_one.p, &one, sizeof(zc), cudaMemcpyHostToDevice));
    }

    // the un-shifted operator of the outer Krylov method: matrix-free 5/7-point stencil, or the stored HO stencil
    void krylov_apply(int mode, const FineOp<T>& Hop, const C* x, const C* b, C* out, int64_t N, int nrhs) {
        if (ho) coarse_stencil(mode, hoH, x, b, out, nrhs);
        else fine_stencil(mode, Hop, x, b, out, N, nrhs, T(0));
    }

    // ------------------------------------------------------------------ outer Krylov
    int krylov_vectors(const hh_solve_options& o) const {
        return o.krylov == HH_KRYLOV_GMRES ? (2 * o.inner + 1) : 7;
    }
    // reusable_bytes: device memory the caller already holds and will reuse for this solve (host staging slots)
    int64_t max_rhs_per_batch(const hh_solve_options& o, double reusable_bytes = 0.0) override {
        HH_CUDA(cudaSetDevice(device));
        size_t fr = 0, tot = 0;
        HH_CUDA(cudaMemGetInfo(&fr, &tot));
        double avail = (double)fr + reusable_bytes;
        // memory we already hold for work vectors counts as available for re-use
        avail += (double)kcap * level_bytes_per_rhs() + (double)kry.n * sizeof(C);
        const int64_t Nf = have_hierarchy ? levels[0].N : pb.N();
        const double per = level_bytes_per_rhs() + (double)(krylov_vectors(o) + (padded() ? 2 : 0)) * Nf * S;
        int64_t k = (int64_t)std::floor(0.90 * avail / per);
        return std::max<int64_t>(k, 0);
    }
    // The arena is addressed with the vector stride N * kry_cap (fgmres, bicgstab, the padded B / X blocks), so what
    // must fit is vectors x N x kry_cap -- not x nrhs: Krylov method and restart length are per-call options, and a
    // later solve with more vectors but fewer right-hand sides would otherwise write past the end.
    void ensure_krylov_memory(const hh_solve_options& o, int nrhs) {
        const size_t nvec = (size_t)(krylov_vectors(o) + (padded() ? 2 : 0));
        if (nrhs <= kry_cap && nvec * levels[0].N * kry_cap <= kry.n) return;
        kry.release();
        alloc_zero(kry, nvec * levels[0].N * nrhs);
        kry_cap = nrhs;
    }

    int solve_device(const void* dB, void* dX, int64_t nrhs64, const hh_solve_options& o, int32_t* iters,
                     double* relres) override {
        HH_CUDA(cudaSetDevice(device));
        HH_REQUIRE(have_hierarchy, HH_ERR_STATE, "hh_setup has not been called");
        HH_REQUIRE(o.krylov == HH_KRYLOV_GMRES || o.krylov == HH_KRYLOV_BICGSTAB, HH_ERR_ARG, "bad krylov");
        HH_REQUIRE(o.krylov != HH_KRYLOV_GMRES || (o.inner >= 1 && o.inner <= 64), HH_ERR_ARG, "inner must be in 1..64");
        HH_REQUIRE(o.max_iter >= 1, HH_ERR_ARG, "max_iter must be >= 1");
        HH_REQUIRE((o.do_transpose != 0) == (opt.do_transpose != 0), HH_ERR_STATE,
                   "hierarchy was built for the other transpose state; call hh_setup with do_transpose");
        const int nrhs = (int)nrhs64;
        auto t0 = std::chrono::steady_clock::now();
        ensure_level_memory(nrhs);
        ensure_krylov_memory(o, nrhs);
        ensure_const();
        int rc;
        const C* Bs = (const C*)dB;
        C* Xs = (C*)dX;
        if (padded()) {  // work on padded copies (the last two blocks of the Krylov arena); ghost nodes stay zero
            const int64_t vs = levels[0].N * kry_cap;
            C* bp = kry.p + (int64_t)krylov_vectors(o) * vs;
            C* xp = bp + vs;
            repitch(Bs, bp, nrhs, true);
            Bs = bp;
            Xs = xp;
        }
        if (o.krylov == HH_KRYLOV_GMRES) rc = fgmres(Bs, Xs, nrhs, o, iters, relres);
        else rc = bicgstab(Bs, Xs, nrhs, o, iters, relres);
        if (padded()) {
            repitch(Xs, (C*)dX, nrhs, false);
            HH_CUDA(cudaStreamSynchronize(stream));
        }
        solve_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return rc;
    }

    // returns true when every RHS is done; throws on NaN
    // mapped pinned block [done | nprec | err] the scalar state is published into (k_publish_state)
    void ensure_state_block(int nrhs) {
        if (nrhs <= state_cap) return;
        if (h_state) HH_CUDA(cudaFreeHost(h_state));
        h_state = nullptr;
        state_cap = std::max(nrhs, 64);
        HH_CUDA(cudaHostAlloc((void**)&h_state, (size_t)state_cap * (2 * sizeof(int) + sizeof(double)), cudaHostAllocMapped));
        HH_CUDA(cudaHostGetDevicePointer((void**)&d_state, h_state, 0));
    }
    double* state_err(char* base) const { return reinterpret_cast<double*>(base); }                      // doubles first (alignment)
    int* state_done(char* base) const { return reinterpret_cast<int*>(base + (size_t)state_cap * sizeof(double)); }
    int* state_nprec(char* base) const { return state_done(base) + state_cap; }
    void publish_state(const int* d_done, const int* d_nprec, const double* d_err, int nrhs) {
        ensure_state_block(nrhs);
        launch(T_SCALAR, 0, [&] {
            k_publish_state<<<(nrhs + 127) / 128, 128, 0, stream>>>(d_done, d_nprec, d_err, state_done(d_state), state_nprec(d_state),
                                                                   state_err(d_state), nrhs);
        });
        HH_CUDA(cudaStreamSynchronize(stream));
    }
    bool fetch_done(const int* d_done, int nrhs) {
        if (!mapped_state) {  // HH_MAPPED_STATE=0: the copy-engine form (A/B)
            h_done.resize(nrhs);
            HH_CUDA(cudaMemcpyAsync(h_done.data(), d_done, nrhs * sizeof(int), cudaMemcpyDeviceToHost, stream));
            HH_CUDA(cudaStreamSynchronize(stream));
        } else {
            publish_state(d_done, nullptr, nullptr, nrhs);
            h_done.assign(state_done(h_state), state_done(h_state) + nrhs);
        }
        bool all = true;
        for (int r = 0; r < nrhs; ++r) {
            if (h_done[r] == 2) throw Error(HH_ERR_NAN, "NaN in the residual norm of right-hand side " + std::to_string(r));
            all = all && (h_done[r] != 0);
        }
        return all;
    }

    // Right-preconditioned restarted flexible GMRES, batched over the RHS block (each RHS keeps its own
    // Hessenberg / Givens scalars).  Counterpart of KrylovMethods.fgmres as called by solveGMRES_MG
    // (ShiftedLaplacianMultigridSolver.jl:89).
    int fgmres(const C* B, C* X, int nrhs, const hh_solve_options& o, int32_t* iters, double* relres) {
        const int m = o.inner;
        const int64_t N = levels[0].N;
        const Span sp = span(0);
        const int64_t vs = N * kry_cap;
        if (outer.st.m != m || outer_cap < nrhs) {
            alloc_gmres_state(outer, m, nrhs);
            outer_cap = nrhs;
        }
        std::vector<C*> W(m + 1), Z(m);
        std::vector<const C*> V(m + 1);
        for (int i = 0; i <= m; ++i) V[i] = W[i] = kry.p + (int64_t)i * vs;
        for (int i = 0; i < m; ++i) Z[i] = kry.p + (int64_t)(m + 1 + i) * vs;
        const FineOp<T> Hop = krylov_op(o.do_transpose);  // Afun: the un-shifted operator (GetHelmholtz.jl:85-95)
        zero_vec(X, N, nrhs);
        // r0 = b (x0 = 0): the first basis vector is B itself, used in place and unscaled (d_0 = ||b||)
        V[0] = B;
        gmres_begin(outer, B, sp, nrhs, true, o.rel_tol);
        bool all_done = fetch_done(outer.done.p, nrhs);
        for (int cyc = 0; cyc < o.max_iter && !all_done; ++cyc) {
            for (int j = 0; j < m; ++j) {
                precondition(V[j], Z[j], nrhs);
                krylov_apply(MODE_APPLY, Hop, Z[j], nullptr, W[j + 1], N, nrhs);
                gmres_orthogonalise(outer, V.data(), j, W[j + 1], sp, nrhs, o.rel_tol, j == m - 1);
                all_done = fetch_done(outer.done.p, nrhs);
                if (all_done) break;
            }
            gmres_solve_y(outer, nrhs);
            const C* zz[HH_MAXV];
            for (int i0 = 0; i0 < m; i0 += HH_MAXV) {
                const int nv = std::min(HH_MAXV, m - i0);
                for (int i = 0; i < nv; ++i) zz[i] = Z[i0 + i];
                multiaxpy(zz, nv, X, sp, nrhs, outer.y.p + i0, m, false, false, d_partial.p);
            }
            if (all_done || cyc + 1 == o.max_iter) break;
            // restart: r = b - H x  (again used unscaled as the first basis vector)
            V[0] = W[0];
            krylov_apply(MODE_RESID, Hop, X, B, W[0], N, nrhs);
            gmres_begin(outer, W[0], sp, nrhs, false, o.rel_tol);
            all_done = fetch_done(outer.done.p, nrhs);
        }
        return finish(outer.nprec.p, outer.err.p, nrhs, iters, relres, all_done);
    }

    int finish(const int* d_nprec, const double* d_err, int nrhs, int32_t* iters, double* relres, bool all_done) {
        std::vector<int> np(nrhs);
        std::vector<double> er(nrhs);
        if (!mapped_state) {
            HH_CUDA(cudaMemcpyAsync(np.data(), d_nprec, nrhs * sizeof(int), cudaMemcpyDeviceToHost, stream));
            HH_CUDA(cudaMemcpyAsync(er.data(), d_err, nrhs * sizeof(double), cudaMemcpyDeviceToHost, stream));
            HH_CUDA(cudaStreamSynchronize(stream));
        } else {
            publish_state(d_nprec /* any int array: the flags are not read here */, d_nprec, d_err, nrhs);
            np.assign(state_nprec(h_state), state_nprec(h_state) + nrhs);
            er.assign(state_err(h_state), state_err(h_state) + nrhs);
        }
        for (int r = 0; r < nrhs; ++r) {
            if (iters) iters[r] = np[r];
            if (relres) relres[r] = er[r];
        }
        if (prof.on) prof.flush(stream);
        return all_done ? HH_OK : HH_NOT_CONVERGED;
    }

    // Preconditioned BiCGSTAB, batched (KrylovMethods.bicgstb as called by solveBiCGSTAB_MG,
    // ShiftedLaplacianMultigridSolver.jl:92); two preconditioner applications per iteration.
    struct BicgMem {
        DevBuf<zc> rho, rho_old, alpha, omega, beta, neg_alpha, neg_omega, ao;
        DevBuf<double> bnorm, err;
        DevBuf<int> done, half, nprec, iters;
        BicgState st;
        int cap = 0;
    };
    int bicgstab(const C* B, C* X, int nrhs, const hh_solve_options& o, int32_t* iters, double* relres) {
        const int64_t N = levels[0].N;
        const Span sp = span(0);
        const int64_t vs = N * kry_cap;
        if (bicg.cap < nrhs) {
            BicgMem& g = bicg;
            g.rho.alloc(nrhs); g.rho_old.alloc(nrhs); g.alpha.alloc(nrhs); g.omega.alloc(nrhs); g.beta.alloc(nrhs);
            g.neg_alpha.alloc(nrhs); g.neg_omega.alloc(nrhs); g.ao.alloc(2 * (size_t)nrhs);
            g.bnorm.alloc(nrhs); g.err.alloc(nrhs); g.done.alloc(nrhs); g.half.alloc(nrhs); g.nprec.alloc(nrhs); g.iters.alloc(nrhs);
            g.st = BicgState{g.rho.p, g.rho_old.p, g.alpha.p, g.omega.p, g.beta.p, g.neg_alpha.p, g.neg_omega.p, g.ao.p,
                             g.bnorm.p, g.err.p, g.done.p, g.half.p, g.nprec.p, g.iters.p};
            g.cap = nrhs;
        }
        C* r = kry.p;
        C* rt = kry.p + vs;
        C* p = kry.p + 2 * vs;
        C* v = kry.p + 3 * vs;
        C* t = kry.p + 4 * vs;
        C* ph = kry.p + 5 * vs;
        C* sh = kry.p + 6 * vs;
        const FineOp<T> Hop = krylov_op(o.do_transpose);
        auto scalars = [&](int nblk, int stage) {
            launch(T_SCALAR, 0, [&] { k_bicg_scalars<<<nrhs, 32, 0, stream>>>(bicg.st, red_out, nblk, stage, o.rel_tol); });
        };
        zero_vec(X, N, nrhs);
        copy_vec(B, r, N, nrhs);
        copy_vec(B, rt, N, nrhs);
        zero_vec(p, N, nrhs);
        zero_vec(v, N, nrhs);
        scalars(multidot(nullptr, 0, B, sp, nrhs, true, d_partial.p), BICG_INIT);
        bool all_done = fetch_done(bicg.done.p, nrhs);
        const C* one[2];
        for (int it = 0; it < o.max_iter && !all_done; ++it) {
            one[0] = rt;
            scalars(multidot(one, 1, r, sp, nrhs, false, d_partial.p), BICG_RHO);
            {
                dim3 g(vec_blocks(sp.len, nrhs), nrhs);
                launch(T_AXPY, 4 * S * (double)sp.len * nrhs, [&] {
                    k_bicg_p<T><<<g, 256, 0, stream>>>(p + sp.off, r + sp.off, v + sp.off, sp.len, sp.ld, bicg.beta.p, bicg.omega.p);
                });
            }
            precondition(p, ph, nrhs);
            krylov_apply(MODE_APPLY, Hop, ph, nullptr, v, N, nrhs);
            one[0] = rt;
            scalars(multidot(one, 1, v, sp, nrhs, false, d_partial.p), BICG_ALPHA);
            one[0] = v;  // s = r - alpha v  (in place in r)
            scalars(multiaxpy(one, 1, r, sp, nrhs, bicg.neg_alpha.p, 1, false, true, d_partial.p), BICG_HALF);
            precondition(r, sh, nrhs);
            krylov_apply(MODE_APPLY, Hop, sh, nullptr, t, N, nrhs);
            one[0] = r;  // <s,t> and |t|^2
            scalars(multidot(one, 1, t, sp, nrhs, true, d_partial.p), BICG_OMEGA);
            one[0] = ph;
            one[1] = sh;
            multiaxpy(one, 2, X, sp, nrhs, bicg.ao.p, 2, false, false, d_partial.p);
            one[0] = t;  // r = s - omega t
            scalars(multiaxpy(one, 1, r, sp, nrhs, bicg.neg_omega.p, 1, false, true, d_partial.p), BICG_END);
            all_done = fetch_done(bicg.done.p, nrhs);
        }
        return finish(bicg.nprec.p, bicg.err.p, nrhs, iters, relres, all_done);
    }

   private:
    enum { FK_TMA = 0, FK_ZMARCH = 1, FK_SIMPLE = 2 };
    int fine_kernel = FK_TMA;
    bool fuse_first = true;
    bool use_pitch = false;
    DevBuf<T> d_mp, d_gp;
    DevBuf<C> mg_cdiag, mg_dinv, h_cdiag[2];
    FineOp<T> h_op[2];
    bool h_op_valid[2] = {false, false};
    DevBuf<T> d_m, d_g;
    std::vector<Level> levels;
    bool have_hierarchy = false;
    hh_mg_options opt{};
    FineOp<T> mg_fine{};
    DevBuf<C> inv_dense;
    int kcap = 0;
    DevBuf<C> kry;
    int kry_cap = 0;
    DevBuf<zc> d_partial, d_one, d_red;
    const zc* red_out = nullptr;         // where the last multidot / multiaxpy left its (all-reduced) sums: see slab_reduce
    Level hoH;  // HO mode: the un-shifted operator H as a stored stencil on the fine grid (the outer Krylov operator)
    cudaStream_t comm_stream = nullptr;  // halo exchanges that overlap the interior planes (with_halos)
    cudaEvent_t ev_ready = nullptr, ev_landed = nullptr;
    bool halo_overlap = false;           // HH_HALO_OVERLAP=1: exchange on a second stream while the interior planes run
    bool split_always = false;
    int rz_b = -1, rz_e = -1;            // plane range override of the coarse launchers (with_halos)
    bool use_scaled = false;             // coarse_op() hands out the column-scaled coefficients (coarse_stencil)
    int force_tile = -1;                 // HH_COARSE_TILE: 0 = 16x8, 1 = alternative tile (coarse_tile)
    bool scaled_gmres = true;
    bool tma_restrict = true;            // HH_TMA_RESTRICT=0: the one-thread-per-coarse-node restriction (A/B baseline)
    bool skip_last_update = true;        // HH_SKIP_LAST_UPDATE=0: orthogonalise the last column of a cycle like the others
    bool pro_cache = true;               // HH_PRO_CACHE=0: the interpolation is recomputed for every fine plane
    bool level_swap = true;              // HH_LEVEL_SWAP=0: copy back after an odd sweep count instead of swapping x / scratch
    bool fuse_recompute = true;          // HH_FUSE_RECOMPUTE=0: npre == 1 cycles recompute x1 = dinv .* b instead of storing it
    bool scalar_fast = true;             // HH_SCALAR_FAST=0: the one-warp, one-thread forms of the GMRES scalar kernels
    bool small_fused = true;             // HH_SMALL_FUSED=0: two scalar kernels / two reductions per step of a level's GMRES
    bool fuse_post2 = false;             // HH_FUSE_POST2=1: correction + BOTH post-sweeps in one pass (k_fine3d_tma_pro2; slower, see ctor)
    GmresMem outer;
    int outer_cap = 0;
    BicgMem bicg;
    std::vector<int> h_done;
    char* h_state = nullptr;             // mapped pinned host block the per-RHS state is published into
    char* d_state = nullptr;             // its device alias
    int state_cap = 0;
    bool mapped_state = true;            // HH_MAPPED_STATE=0: cudaMemcpyAsync of the flags (queues behind bulk D2H copies)
};

}  // namespace hh
