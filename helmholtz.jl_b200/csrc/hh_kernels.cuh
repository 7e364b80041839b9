// hh_kernels.cuh -- sm_100a kernels of the shifted-Laplacian multigrid Helmholtz solve.
//
// Operator (reference src/GetHelmholtz.jl:33-50,81-83,222-247 and
// src/PlainNodalLaplacian.jl:18-46):
//   (H u)_p = sum_d L_d(u)_p + c_p u_p,   L = -laplacian with ghost-eliminated Neumann rows,
//   c_p = -w^2 m_p (1 - i g_p / Re w) + i Re(w) sqrt(m_p) sum_faces 2/h_d  [+ i shift Re(w)^2 m_p]
// The fine level is matrix-free (m, gamma are the only arrays read besides the vectors); coarse
// levels hold the Galerkin 3^dim-point stencil as structure-of-arrays coef[s][node].
#pragma once
#include "hh_common.cuh"

namespace hh {

enum { MODE_APPLY = 0, MODE_RESID = 1, MODE_JACOBI = 2 };

template <typename T>
struct FineOp {
    const T* m;   // slowness squared, N reals
    const T* g;   // gamma (attenuation incl. absorbing layer), N reals
    T a, b;       // omega^2 = a + i b
    T inv_wr;     // 1 / Re(omega)
    T shift_w2;   // shift * Re(omega)^2   (0 for the un-shifted operator)
    T somm[3];    // Re(omega) * 2/h_d when Sommerfeld, else 0
    T ih2[3];     // 1/h_d^2
    T BC;         // 2 (second-order Neumann ghost) or 1
    int n[3];     // node counts (n[2] == 1 in 2-D)
    int sy;       // row pitch (elements) of every array this operator touches: n[0], or n[0]+1 when ComplexF32
                  // rows are padded to a 16-byte multiple for TMA; the ghost column is never read or written
    int neumann_top;
    int adj;      // 1: conjugate transpose
    // Slab decomposition along the last dimension (one slab per GPU): the arrays hold local planes 0..n[2]-1, local
    // plane z is global plane z + koff of n2g, and the kernels compute planes zb <= z < ze only (the planes outside
    // that range are halo / alignment planes filled by the halo exchange).  Whole grid: koff = 0, n2g = n[2], zb = 0,
    // ze = n[2].
    int koff, n2g, zb, ze;
    // optional precomputed diagonal arrays (k_fine_precompute): centre coefficient incl. the Laplacian
    // diagonal, and damp/centre.  Used by the TMA-staged production kernels.
    const cx<T>* cdiag;
    const cx<T>* dinv;
};

template <typename T>
struct CoarseOp {
    const cx<T>* coef;  // [3^dim][N] stencil coefficients
    const cx<T>* dinv;  // [N] damping / diagonal
    int n[3];
    int sy;             // row pitch (elements); N = padded node count sy*n[1]*n[2] = stride between coefficients
    int64_t N;
    int zb, ze;         // planes computed (slab decomposition: the owned planes; whole grid: 0, n[2]), see FineOp
};

// ---------------------------------------------------------------------------------------------
// fine-level coefficient evaluation
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM>
__device__ __forceinline__ cx<T> fine_center(const FineOp<T>& op, int64_t p, int i, int j, int kl) {
    const int k = kl + op.koff;  // global plane (boundary faces are global)
    const T mv = op.m[p];
    const T gv = op.g[p] * op.inv_wr;
    T re = -mv * (op.a + op.b * gv);
    T im = -mv * (op.b - op.a * gv) + op.shift_w2 * mv;
    T sf = T(0);
    const bool bi = (i == 0) | (i == op.n[0] - 1);
    const bool bj = (j == 0) | (j == op.n[1] - 1);
    if (DIM == 2) {
        if (bi) sf += op.somm[0];
        if ((j == 0 && !op.neumann_top) || j == op.n[1] - 1) sf += op.somm[1];
    } else {
        if (bi) sf += op.somm[0];
        if (bj) sf += op.somm[1];
        if ((k == 0 && !op.neumann_top) || k == op.n2g - 1) sf += op.somm[2];
    }
    if (sf != T(0)) im += sf * sqrt(mv);
    re += (bi ? op.BC : T(2)) * op.ih2[0];
    re += (bj ? op.BC : T(2)) * op.ih2[1];
    if (DIM == 3) re += ((k == 0 || k == op.n2g - 1) ? op.BC : T(2)) * op.ih2[2];
    if (op.adj) im = -im;
    return mk<T>(re, im);
}

// weight w such that the row of node `idx` (along dimension d, n nodes) holds -w on the neighbour at
// idx-1 (side 0) / idx+1 (side 1); 0 when that neighbour does not exist.
template <typename T>
__device__ __forceinline__ T fine_w(const FineOp<T>& op, int d, int side, int idx, int n) {
    if (side == 0) {
        if (idx == 0) return T(0);
        const bool bc = op.adj ? (idx - 1 == 0) : (idx == n - 1);
        return (bc ? op.BC : T(1)) * op.ih2[d];
    } else {
        if (idx == n - 1) return T(0);
        const bool bc = op.adj ? (idx + 1 == n - 1) : (idx == 0);
        return (bc ? op.BC : T(1)) * op.ih2[d];
    }
}

// same along the last dimension of a 3-D grid, for LOCAL plane z of a slab (weights follow the global plane)
template <typename T>
__device__ __forceinline__ T fine_wz(const FineOp<T>& op, int side, int z) {
    return fine_w(op, 2, side, z + op.koff, op.n2g);
}

// ---------------------------------------------------------------------------------------------
// K1/K2 (baseline form): fine-level stencil, one thread per node, KB right-hand sides per pass so
// that m/gamma and the index arithmetic are amortised.  MODE selects the fused epilogue:
//   APPLY : out = A x        RESID : out = b - A x        JACOBI : out = x + damp/diag (b - A x)
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM, int MODE, int KB>
__global__ void __launch_bounds__(256) k_fine_stencil(FineOp<T> op, const cx<T>* __restrict__ x,
                                                      const cx<T>* __restrict__ b, cx<T>* __restrict__ out,
                                                      int64_t ld, int nrhs, T damp) {
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = (DIM == 3) ? op.zb + blockIdx.z * blockDim.z + threadIdx.z : 0;
    if (i >= n0 || j >= n1 || (DIM == 3 && k >= op.ze)) return;
    (void)n2;
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const int64_t p = i + sy * j + sz * k;
    const cx<T> c = fine_center<T, DIM>(op, p, i, j, k);
    const T wxm = fine_w(op, 0, 0, i, n0), wxp = fine_w(op, 0, 1, i, n0);
    const T wym = fine_w(op, 1, 0, j, n1), wyp = fine_w(op, 1, 1, j, n1);
    const T wzm = (DIM == 3) ? fine_wz(op, 0, k) : T(0);
    const T wzp = (DIM == 3) ? fine_wz(op, 1, k) : T(0);
    // clamp neighbour offsets so that absent neighbours (weight 0) read the centre
    const int64_t oxm = wxm != T(0) ? -1 : 0, oxp = wxp != T(0) ? 1 : 0;
    const int64_t oym = wym != T(0) ? -sy : 0, oyp = wyp != T(0) ? sy : 0;
    const int64_t ozm = wzm != T(0) ? -sz : 0, ozp = wzp != T(0) ? sz : 0;
    cx<T> dinv = mk<T>(T(0), T(0));
    if (MODE == MODE_JACOBI) dinv = rdiv(damp, c);
    for (int r0 = 0; r0 < nrhs; r0 += KB) {
        cx<T> acc[KB], xc[KB];
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            if (r0 + q < nrhs) {
                const cx<T>* xr = x + (int64_t)(r0 + q) * ld + p;
                xc[q] = xr[0];
                cx<T> a = c * xc[q];
                rfma(a, -wxm, xr[oxm]);
                rfma(a, -wxp, xr[oxp]);
                rfma(a, -wym, xr[oym]);
                rfma(a, -wyp, xr[oyp]);
                if (DIM == 3) {
                    rfma(a, -wzm, xr[ozm]);
                    rfma(a, -wzp, xr[ozp]);
                }
                acc[q] = a;
            }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            if (r0 + q < nrhs) {
                const int64_t o = (int64_t)(r0 + q) * ld + p;
                if (MODE == MODE_APPLY) {
                    out[o] = acc[q];
                } else if (MODE == MODE_RESID) {
                    out[o] = b[o] - acc[q];
                } else {
                    out[o] = xc[q] + dinv * (b[o] - acc[q]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1/K2 (3-D production form): z-marching 7-point stencil.  A CTA owns a 32 x TY tile of (i,j)
// columns and walks a chunk of z planes; each thread keeps the z-1 / z / z+1 values of its column
// in registers, so every x value is fetched from L2/HBM once per sweep (the four in-plane
// neighbours were fetched as centres one step earlier by this warp or its neighbours and hit L1).
// KB right-hand sides share one evaluation of the coefficients (m, gamma, Sommerfeld, weights).
// blockIdx.x enumerates (RHS group, x tile) with the group fastest, so that the CTAs that share a tile's
// m/gamma are co-scheduled and those coefficients come from HBM once per sweep.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int KB, int TY, int MINB>
__global__ void __launch_bounds__(32 * TY, MINB) k_fine3d_zmarch(FineOp<T> op, const cx<T>* __restrict__ x,
                                                                 const cx<T>* __restrict__ b, cx<T>* __restrict__ out,
                                                                 int64_t ld, int nrhs, T damp, int zchunk, int groups) {
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int i = (blockIdx.x / groups) * blockDim.x + threadIdx.x;  // CTA shape (blockDim.x, 32*TY/blockDim.x)
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n0 || j >= n1) return;
    const int zc = blockIdx.z;
    const int r0 = (blockIdx.x % groups) * KB;
    const int z0 = op.zb + zc * zchunk;
    const int z1 = min(op.ze, z0 + zchunk);
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const int64_t pxy = i + sy * j;
    const T wxm = fine_w(op, 0, 0, i, n0), wxp = fine_w(op, 0, 1, i, n0);
    const T wym = fine_w(op, 1, 0, j, n1), wyp = fine_w(op, 1, 1, j, n1);
    const int64_t oxm = wxm != T(0) ? -1 : 0, oxp = wxp != T(0) ? 1 : 0;
    const int64_t oym = wym != T(0) ? -sy : 0, oyp = wyp != T(0) ? sy : 0;
    const bool bi = (i == 0) | (i == n0 - 1), bj = (j == 0) | (j == n1 - 1);
    const T lapxy = (bi ? op.BC : T(2)) * op.ih2[0] + (bj ? op.BC : T(2)) * op.ih2[1];
    const T sfxy = (bi ? op.somm[0] : T(0)) + (bj ? op.somm[1] : T(0));
    const cx<T>* xr[KB];
    cx<T> xm[KB], xc[KB], xp[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) {
        const int r = min(r0 + q, nrhs - 1);  // clamped: surplus slots recompute the last RHS, never stored
        xr[q] = x + (int64_t)r * ld + pxy;
        xc[q] = xr[q][(int64_t)z0 * sz];
        xm[q] = z0 > 0 ? xr[q][(int64_t)(z0 - 1) * sz] : mk<T>(T(0), T(0));
    }
#pragma unroll 1
    for (int z = z0; z < z1; ++z) {
        const int64_t zo = (int64_t)z * sz;
        const int64_t p = pxy + zo;
        const bool zlast = (z == n2 - 1);
#pragma unroll
        for (int q = 0; q < KB; ++q) xp[q] = zlast ? mk<T>(T(0), T(0)) : xr[q][zo + sz];
        // centre coefficient (fine_center, with the z-independent parts hoisted)
        const T mv = op.m[p];
        const T gv = op.g[p] * op.inv_wr;
        T re = -mv * (op.a + op.b * gv);
        T im = -mv * (op.b - op.a * gv) + op.shift_w2 * mv;
        T sf = sfxy;
        const int zg = z + op.koff;
        const bool gtop = (zg == 0), gbot = (zg == op.n2g - 1);
        if ((gtop && !op.neumann_top) || gbot) sf += op.somm[2];
        if (sf != T(0)) im += sf * sqrt(mv);
        re += lapxy + ((gtop || gbot) ? op.BC : T(2)) * op.ih2[2];
        if (op.adj) im = -im;
        const cx<T> c = mk<T>(re, im);
        const T wzm = fine_wz(op, 0, z), wzp = fine_wz(op, 1, z);
        cx<T> dinv = mk<T>(T(0), T(0));
        if (MODE == MODE_JACOBI) dinv = rdiv(damp, c);
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            const cx<T>* xq = xr[q] + zo;
            cx<T> a = c * xc[q];
            rfma(a, -wxm, xq[oxm]);
            rfma(a, -wxp, xq[oxp]);
            rfma(a, -wym, xq[oym]);
            rfma(a, -wyp, xq[oyp]);
            rfma(a, -wzm, xm[q]);
            rfma(a, -wzp, xp[q]);
            if (r0 + q < nrhs) {
                const int64_t o = (int64_t)(r0 + q) * ld + p;
                if (MODE == MODE_APPLY) {
                    out[o] = a;
                } else if (MODE == MODE_RESID) {
                    out[o] = b[o] - a;
                } else {
                    out[o] = xc[q] + dinv * (b[o] - a);
                }
            }
            xm[q] = xc[q];
            xc[q] = xp[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// coarse levels, 3-D production form: z-marching 27-point stencil with stored coefficients.  For
// every input plane a thread reads its 3 x 3 in-plane neighbourhood once and scatters it into three
// rolling accumulators (outputs z-1, z, z+1), so x is read 9 (not 27) times per node through L1 and
// once from L2/HBM; the 27 coefficients of a node are read once per KB right-hand sides.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int KB, int TY, int MINB>
__global__ void __launch_bounds__(32 * TY, MINB) k_coarse3d_zmarch(CoarseOp<T> op, const cx<T>* __restrict__ x,
                                                                   const cx<T>* __restrict__ b, cx<T>* __restrict__ out,
                                                                   int64_t ld, int nrhs, int zchunk, int groups) {
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int i = (blockIdx.x / groups) * blockDim.x + threadIdx.x;  // CTA shape (blockDim.x, 32*TY/blockDim.x)
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n0 || j >= n1) return;
    const int zc = blockIdx.z;
    const int r0 = (blockIdx.x % groups) * KB;
    const int z0 = op.zb + zc * zchunk;
    const int z1 = min(op.ze, z0 + zchunk);
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const int64_t N = op.N;
    const int64_t pxy = i + sy * j;
    const cx<T> zero = mk<T>(T(0), T(0));
    // acc[0]: output plane zi-1, acc[1]: zi, acc[2]: zi+1 while input plane zi is processed
    cx<T> acc[3][KB];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int q = 0; q < KB; ++q) acc[a][q] = zero;
    const cx<T>* xr[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) xr[q] = x + (int64_t)min(r0 + q, nrhs - 1) * ld + pxy;
    const int zi0 = max(z0 - 1, 0), zi1 = min(z1, n2 - 1);  // input planes zi0..zi1 inclusive
    for (int zi = zi0; zi <= zi1; ++zi) {
        const int64_t zo = (int64_t)zi * sz;
        // which outputs receive this plane
        const bool om = (zi - 1 >= z0) && (zi - 1 < z1);  // output zi-1 via dk = +1
        const bool oc = (zi >= z0) && (zi < z1);          // output zi   via dk =  0
        const bool op_ = (zi + 1 >= z0) && (zi + 1 < z1); // output zi+1 via dk = -1
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
            const bool okj = (unsigned)(j + dj) < (unsigned)n1;
#pragma unroll
            for (int di = -1; di <= 1; ++di) {
                const bool ok = okj && ((unsigned)(i + di) < (unsigned)n0);
                if (!ok) continue;
                const int sxy = (di + 1) + 3 * (dj + 1);
                const int64_t off = zo + di + sy * dj;
                cx<T> xv[KB];
#pragma unroll
                for (int q = 0; q < KB; ++q) xv[q] = xr[q][off];
                if (om) {
                    const cx<T> cf = op.coef[(int64_t)(sxy + 18) * N + pxy + zo - sz];
#pragma unroll
                    for (int q = 0; q < KB; ++q) cfma(acc[0][q], cf, xv[q]);
                }
                if (oc) {
                    const cx<T> cf = op.coef[(int64_t)(sxy + 9) * N + pxy + zo];
#pragma unroll
                    for (int q = 0; q < KB; ++q) cfma(acc[1][q], cf, xv[q]);
                }
                if (op_) {
                    const cx<T> cf = op.coef[(int64_t)sxy * N + pxy + zo + sz];
#pragma unroll
                    for (int q = 0; q < KB; ++q) cfma(acc[2][q], cf, xv[q]);
                }
            }
        }
        // output plane zi-1 is complete (or, at the top of the grid, output zi when zi is the last plane)
        if (om) {
            const int64_t p = pxy + zo - sz;
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                if (r0 + q < nrhs) {
                    const int64_t o = (int64_t)(r0 + q) * ld + p;
                    if (MODE == MODE_APPLY) out[o] = acc[0][q];
                    else if (MODE == MODE_RESID) out[o] = b[o] - acc[0][q];
                    else out[o] = x[o] + op.dinv[p] * (b[o] - acc[0][q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            acc[0][q] = acc[1][q];
            acc[1][q] = acc[2][q];
            acc[2][q] = zero;
        }
    }
    // the chunk's last output plane (z1-1) has no further input plane when z1 == n2
    if (z1 == n2) {
        const int64_t p = pxy + (int64_t)(n2 - 1) * sz;
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            if (r0 + q < nrhs) {
                const int64_t o = (int64_t)(r0 + q) * ld + p;
                if (MODE == MODE_APPLY) out[o] = acc[0][q];
                else if (MODE == MODE_RESID) out[o] = b[o] - acc[0][q];
                else out[o] = x[o] + op.dinv[p] * (b[o] - acc[0][q]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA plumbing of the TMA-staged production kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) accesses to the
// same bytes.  Needed wherever threads WRITE a staged tile that the TMA engine refills later (k_fine3d_tma_pro);
// executed by the writers, followed by the CTA barrier the issuing thread waits on.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 4-D / 3-D tiled TMA loads (cp.async.bulk.tensor, SASS UTMALDG): one instruction moves a whole box,
// out-of-range coordinates are zero-filled by the hardware (that is how tile halos at the domain
// boundary and surplus RHS slots are handled).
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// 128-byte tensor-map descriptor (CUtensorMap), passed by value as a __grid_constant__ parameter
struct alignas(64) TmaDesc {
    unsigned char bytes[128];
};

template <typename T, int MODE, int KB>
struct FineTmaCfg {
    static constexpr int TX = 32, TY = 8;
    // The TMA unit needs the innermost box coordinate to be a 16-byte multiple: a one-node x halo is fine for
    // ComplexF64 (16 B per node); ComplexF32 (8 B) starts the halo tile two nodes to the left instead.
    static constexpr int HX = sizeof(T) == 4 ? 2 : 1;
    static constexpr int PX = TX + 2 * HX;  // row pitch of the halo tile (elements)
    static constexpr int XT = (TY + 2) * PX;
    static constexpr int BT = TY * TX;
    static constexpr int ES = (int)sizeof(cx<T>);
    static constexpr int al(int b) { return (b + 127) / 128 * 128; }
    // byte offsets inside a stage (every TMA destination 128-byte aligned)
    static constexpr int OFF_X = 0;
    static constexpr int OFF_B = al(KB * XT * ES);
    static constexpr int OFF_C = OFF_B + (MODE != MODE_APPLY ? al(KB * BT * ES) : 0);
    static constexpr int OFF_D = OFF_C + al(BT * ES);
    static constexpr int STAGE_BYTES = OFF_D + (MODE == MODE_JACOBI ? al(BT * ES) : 0);
    static constexpr uint32_t TX_BYTES = KB * XT * ES + (MODE != MODE_APPLY ? KB * BT * ES : 0) + BT * ES +
                                         (MODE == MODE_JACOBI ? BT * ES : 0);
};

// K1/K2 (3-D, TMA production form).  Same z-marching decomposition as k_fine3d_zmarch, but every operand
// plane is staged into shared memory by the TMA engine with an NS-deep mbarrier ring: the x tile with a
// one-node halo (all KB right-hand sides in one box), the b tile and the precomputed diagonal tiles of
// planes z+1 .. z+NS-1 are in flight while plane z is computed, so memory-level parallelism no longer
// depends on occupancy or registers.  In-plane neighbours are read from the staged tile.
template <typename T, int MODE, int KB, int NS>
__global__ void __launch_bounds__(256) k_fine3d_tma(FineOp<T> op, const __grid_constant__ TmaDesc tm_x,
                                                    const __grid_constant__ TmaDesc tm_b,
                                                    const __grid_constant__ TmaDesc tm_c,
                                                    const __grid_constant__ TmaDesc tm_d, const cx<T>* __restrict__ x,
                                                    cx<T>* __restrict__ out, int64_t ld, int nrhs, int zchunk,
                                                    int groups) {
    typedef FineTmaCfg<T, MODE, KB> Cfg;
    constexpr int TX = Cfg::TX, TY = Cfg::TY, PX = Cfg::PX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * Cfg::STAGE_BYTES);
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = (blockIdx.x / groups) * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int r0 = (blockIdx.x % groups) * KB;
    const int z0 = op.zb + blockIdx.z * zchunk;
    const int z1 = min(op.ze, z0 + zchunk);
    const int zl = min(z1, n2 - 1);  // last plane that must be staged (z+1 halo of the chunk)
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const bool active = (i < n0) && (j < n1);
    auto issue = [&](int s, int z) {
        unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&bars[s], Cfg::TX_BYTES);
        tma_load_4d(st + Cfg::OFF_X, &tm_x, 2 * (i0 - Cfg::HX), j0 - 1, z, r0, &bars[s]);
        if (MODE != MODE_APPLY) tma_load_4d(st + Cfg::OFF_B, &tm_b, 2 * i0, j0, z, r0, &bars[s]);
        tma_load_3d(st + Cfg::OFF_C, &tm_c, 2 * i0, j0, z, &bars[s]);
        if (MODE == MODE_JACOBI) tma_load_3d(st + Cfg::OFF_D, &tm_d, 2 * i0, j0, z, &bars[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NS && z0 + s <= zl; ++s) issue(s, z0 + s);
    }
    __syncthreads();
    // z-invariant pieces
    const int ic = active ? i : 0, jc = active ? j : 0;
    const T wxm = fine_w(op, 0, 0, ic, n0), wxp = fine_w(op, 0, 1, ic, n0);
    const T wym = fine_w(op, 1, 0, jc, n1), wyp = fine_w(op, 1, 1, jc, n1);
    const int64_t pxy = ic + sy * jc;
    const int cidx = (ty + 1) * PX + (tx + Cfg::HX);  // centre of this thread inside an x tile
    const int bidx = ty * TX + tx;                    // inside a b / diagonal tile
    cx<T> xm[KB], xc[KB], xp[KB];
    mbar_wait(&bars[0], 0);
#pragma unroll
    for (int q = 0; q < KB; ++q) {
        xc[q] = reinterpret_cast<const cx<T>*>(smem_raw + Cfg::OFF_X)[q * Cfg::XT + cidx];
        const int r = min(r0 + q, nrhs - 1);
        xm[q] = (z0 > 0 && active) ? x[(int64_t)r * ld + pxy + (int64_t)(z0 - 1) * sz] : mk<T>(T(0), T(0));
    }
#pragma unroll 1
    for (int z = z0; z < z1; ++z) {
        const int s = (z - z0) % NS;
        const unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        const bool zlast = (z == n2 - 1);
        if (!zlast) {
            const int s1 = (z + 1 - z0) % NS;
            mbar_wait(&bars[s1], (uint32_t)(((z + 1 - z0) / NS) & 1));
            const cx<T>* x1 = reinterpret_cast<const cx<T>*>(smem_raw + (size_t)s1 * Cfg::STAGE_BYTES + Cfg::OFF_X);
#pragma unroll
            for (int q = 0; q < KB; ++q) xp[q] = x1[q * Cfg::XT + cidx];
        } else {
#pragma unroll
            for (int q = 0; q < KB; ++q) xp[q] = mk<T>(T(0), T(0));
        }
        if (active) {
            const cx<T>* sx = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_X);
            const cx<T>* sb = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_B);
            const cx<T> c = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_C)[bidx];
            cx<T> dinv = mk<T>(T(0), T(0));
            if (MODE == MODE_JACOBI) dinv = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_D)[bidx];
            const T wzm = fine_wz(op, 0, z), wzp = fine_wz(op, 1, z);
            const int64_t p = pxy + (int64_t)z * sz;
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                const cx<T>* xt = sx + q * Cfg::XT + cidx;
                cx<T> a = c * xc[q];
                rfma(a, -wxm, xt[-1]);  // halo cells outside the grid were zero-filled by the TMA unit
                rfma(a, -wxp, xt[1]);
                rfma(a, -wym, xt[-PX]);
                rfma(a, -wyp, xt[PX]);
                rfma(a, -wzm, xm[q]);
                rfma(a, -wzp, xp[q]);
                if (r0 + q < nrhs) {
                    const int64_t o = (int64_t)(r0 + q) * ld + p;
                    if (MODE == MODE_APPLY) {
                        out[o] = a;
                    } else {
                        const cx<T> bv = sb[q * Cfg::BT + bidx];
                        if (MODE == MODE_RESID) out[o] = bv - a;
                        else out[o] = xc[q] + dinv * (bv - a);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            xm[q] = xc[q];
            xc[q] = xp[q];
        }
        __syncthreads();  // every thread is done with stage s: refill it with plane z + NS
        if (threadIdx.x == 0 && z + NS <= zl) issue(s, z + NS);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused coarse-grid correction + first post-smoothing sweep on the fine level (3-D, TMA form):
//   x' = x + P xc ;  out = x' + dinv .* (b - A x')
// The tri-linear interpolation is applied to the staged x tile (with halo) in shared memory as each
// plane arrives -- the coarse tiles of the two coarse planes that plane touches ride in the same stage
// -- so x' is never written to and re-read from HBM.  Saves the 2S + S/8 pass of k_prolong_add.
// ---------------------------------------------------------------------------------------------
template <typename T, int KB>
struct FineProCfg {
    static constexpr int TX = 32, TY = 8;
    static constexpr int HX = sizeof(T) == 4 ? 2 : 1;  // see FineTmaCfg
    static constexpr int PX = TX + 2 * HX;
    static constexpr int XT = (TY + 2) * PX;
    static constexpr int BT = TY * TX;
    // coarse tile with halo: starts CH coarse nodes left of i0/2 (16-byte aligned start, even width for ComplexF32)
    static constexpr int CH = sizeof(T) == 4 ? 2 : 1;
    static constexpr int CTX = sizeof(T) == 4 ? TX / 2 + 4 : TX / 2 + 2, CTY = TY / 2 + 2, CT = CTX * CTY;
    static constexpr int ES = (int)sizeof(cx<T>);
    static constexpr int al(int b) { return (b + 127) / 128 * 128; }
    static constexpr int OFF_X = 0;
    static constexpr int OFF_B = al(KB * XT * ES);
    static constexpr int OFF_C = OFF_B + al(KB * BT * ES);
    static constexpr int OFF_D = OFF_C + al(BT * ES);
    static constexpr int OFF_XC = OFF_D + al(BT * ES);            // two coarse planes, KB right-hand sides each
    static constexpr int XC_PLANE = al(KB * CT * ES);
    static constexpr int STAGE_BYTES = OFF_XC + 2 * XC_PLANE;
    static constexpr uint32_t TX_BYTES = KB * XT * ES + KB * BT * ES + 2 * BT * ES + 2 * KB * CT * ES;
};

// (P xc)(i,j,k) from the global coarse array (used once per column for the plane below the chunk)
template <typename T>
__device__ __forceinline__ cx<T> prolong_point(const cx<T>* __restrict__ xc, int i, int j, int k, int csy, int nc1) {
    const int oi = i & 1, oj = j & 1, ok = k & 1;
    const cx<T>* c = xc + (i >> 1) + (int64_t)csy * ((j >> 1) + (int64_t)nc1 * (k >> 1));
    cx<T> acc = mk<T>(T(0), T(0));
    for (int a2 = 0; a2 <= ok; ++a2)
        for (int a1 = 0; a1 <= oj; ++a1)
            for (int a0 = 0; a0 <= oi; ++a0) acc = acc + c[a0 + (int64_t)csy * (a1 + (int64_t)nc1 * a2)];
    return (T(1) / T(1 << (oi + oj + ok))) * acc;
}

template <typename T, int KB, int NS>
__global__ void __launch_bounds__(256) k_fine3d_tma_pro(FineOp<T> op, const __grid_constant__ TmaDesc tm_x,
                                                        const __grid_constant__ TmaDesc tm_b,
                                                        const __grid_constant__ TmaDesc tm_c,
                                                        const __grid_constant__ TmaDesc tm_d,
                                                        const __grid_constant__ TmaDesc tm_xc,
                                                        const cx<T>* __restrict__ x, const cx<T>* __restrict__ xcg,
                                                        cx<T>* __restrict__ out, int64_t ld, int64_t ldc, int csy,
                                                        int nc1, int nrhs, int zchunk, int groups) {
    typedef FineProCfg<T, KB> Cfg;
    constexpr int TX = Cfg::TX, TY = Cfg::TY, PX = Cfg::PX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * Cfg::STAGE_BYTES);
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = (blockIdx.x / groups) * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int r0 = (blockIdx.x % groups) * KB;
    const int z0 = op.zb + blockIdx.z * zchunk;
    const int z1 = min(op.ze, z0 + zchunk);
    const int zl = min(z1, n2 - 1);
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const bool active = (i < n0) && (j < n1);
    const int Is = (i0 >> 1) - Cfg::CH, Js = (j0 >> 1) - 1;  // origin of the coarse tile
    auto issue = [&](int s, int z) {
        unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&bars[s], Cfg::TX_BYTES);
        tma_load_4d(st + Cfg::OFF_X, &tm_x, 2 * (i0 - Cfg::HX), j0 - 1, z, r0, &bars[s]);
        tma_load_4d(st + Cfg::OFF_B, &tm_b, 2 * i0, j0, z, r0, &bars[s]);
        tma_load_3d(st + Cfg::OFF_C, &tm_c, 2 * i0, j0, z, &bars[s]);
        tma_load_3d(st + Cfg::OFF_D, &tm_d, 2 * i0, j0, z, &bars[s]);
        tma_load_4d(st + Cfg::OFF_XC, &tm_xc, 2 * Is, Js, z >> 1, r0, &bars[s]);
        tma_load_4d(st + Cfg::OFF_XC + Cfg::XC_PLANE, &tm_xc, 2 * Is, Js, (z >> 1) + 1, r0, &bars[s]);
    };
    // Interpolation geometry does not depend on z: computed once per thread.
    //  * its own column (centre cell of the tile): corrected when the plane's centre value enters the register
    //    pipeline, and written back so that the neighbours see x' one iteration later;
    //  * one cell of the halo ring (84 cells per RHS) for the first 84*KB threads.
    // Both corrections of plane z+1 happen during iteration z and are ordered before their first use
    // (iteration z+1) by the __syncthreads that ends every iteration: no extra barrier.
    auto geom = [&](int fi, int fj, int q, int& coff, int& par) {
        coff = 0;
        par = 0;
        if ((unsigned)fi < (unsigned)n0 && (unsigned)fj < (unsigned)n1) {
            coff = q * Cfg::CT + ((fj >> 1) - Js) * Cfg::CTX + ((fi >> 1) - Is);
            par = 4 | (fi & 1) | ((fj & 1) << 1);
        }
    };
    auto interp = [&](const cx<T>* c0, int coff, int par, int ok) -> cx<T> {
        const int oi = par & 1, oj = (par >> 1) & 1;
        const cx<T>* p0 = c0 + coff;
        cx<T> acc = p0[0];
        if (oi) acc = acc + p0[1];
        if (oj) {
            acc = acc + p0[Cfg::CTX];
            if (oi) acc = acc + p0[Cfg::CTX + 1];
        }
        if (ok) {
            const cx<T>* p1 = p0 + Cfg::XC_PLANE / Cfg::ES;
            acc = acc + p1[0];
            if (oi) acc = acc + p1[1];
            if (oj) {
                acc = acc + p1[Cfg::CTX];
                if (oi) acc = acc + p1[Cfg::CTX + 1];
            }
        }
        return (T(1) / T(1 << (oi + oj + ok))) * acc;
    };
    const int cidx = (ty + 1) * PX + (tx + Cfg::HX);
    const int bidx = ty * TX + tx;
    int ccoff[KB], cpar;  // centre
    {
        int par0 = 0;
#pragma unroll
        for (int q = 0; q < KB; ++q) geom(i, j, q, ccoff[q], par0);
        cpar = par0;
    }
    int hoff = -1, hcoff = 0, hpar = 0;  // halo-ring cell of this thread (if any)
    if (threadIdx.x < 84 * KB) {
        const int q = threadIdx.x / 84, t = threadIdx.x - q * 84;
        constexpr int RW = TX + 2;  // width of the one-node ring rows
        int row, col;              // col counted from the node i0-1
        if (t < RW) {
            row = 0;
            col = t;
        } else if (t < 2 * RW) {
            row = TY + 1;
            col = t - RW;
        } else if (t < 2 * RW + TY) {
            row = 1 + (t - 2 * RW);
            col = 0;
        } else {
            row = 1 + (t - 2 * RW - TY);
            col = TX + 1;
        }
        geom(i0 - 1 + col, j0 - 1 + row, q, hcoff, hpar);
        if (hpar & 4) hoff = q * Cfg::XT + row * PX + (col + Cfg::HX - 1);
    }
    auto correct_plane = [&](unsigned char* st, int z, cx<T>* xv) {
        cx<T>* xs = reinterpret_cast<cx<T>*>(st + Cfg::OFF_X);
        const cx<T>* c0 = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_XC);
        const int ok = z & 1;
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            cx<T> v = xs[q * Cfg::XT + cidx];
            if (cpar & 4) v = v + interp(c0, ccoff[q], cpar, ok);
            xs[q * Cfg::XT + cidx] = v;
            xv[q] = v;
        }
        if (hoff >= 0) xs[hoff] = xs[hoff] + interp(c0, hcoff, hpar, ok);
        fence_proxy_async();  // these generic stores precede the TMA refill of this stage (after a CTA barrier)
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NS && z0 + s <= zl; ++s) issue(s, z0 + s);
    }
    __syncthreads();
    const int ic = active ? i : 0, jc = active ? j : 0;
    const T wxm = fine_w(op, 0, 0, ic, n0), wxp = fine_w(op, 0, 1, ic, n0);
    const T wym = fine_w(op, 1, 0, jc, n1), wyp = fine_w(op, 1, 1, jc, n1);
    const int64_t pxy = ic + sy * jc;
    cx<T> xm[KB], xc[KB], xp[KB];
    mbar_wait(&bars[0], 0);
    correct_plane(smem_raw, z0, xc);
    __syncthreads();  // once per chunk: plane z0 is used in the first iteration already
#pragma unroll
    for (int q = 0; q < KB; ++q) {
        xm[q] = mk<T>(T(0), T(0));
        if (z0 > 0 && active) {
            const int r = min(r0 + q, nrhs - 1);
            xm[q] = x[(int64_t)r * ld + pxy + (int64_t)(z0 - 1) * sz] +
                    prolong_point<T>(xcg + (int64_t)r * ldc, ic, jc, z0 - 1, csy, nc1);
        }
    }
#pragma unroll 1
    for (int z = z0; z < z1; ++z) {
        const int s = (z - z0) % NS;
        const unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        const bool zlast = (z == n2 - 1);
        if (!zlast) {
            const int s1 = (z + 1 - z0) % NS;
            mbar_wait(&bars[s1], (uint32_t)(((z + 1 - z0) / NS) & 1));
            correct_plane(smem_raw + (size_t)s1 * Cfg::STAGE_BYTES, z + 1, xp);
        } else {
#pragma unroll
            for (int q = 0; q < KB; ++q) xp[q] = mk<T>(T(0), T(0));
        }
        if (active) {
            const cx<T>* sx = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_X);
            const cx<T>* sb = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_B);
            const cx<T> c = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_C)[bidx];
            const cx<T> dinv = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_D)[bidx];
            const T wzm = fine_wz(op, 0, z), wzp = fine_wz(op, 1, z);
            const int64_t p = pxy + (int64_t)z * sz;
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                const cx<T>* xt = sx + q * Cfg::XT + cidx;
                cx<T> a = c * xc[q];
                rfma(a, -wxm, xt[-1]);
                rfma(a, -wxp, xt[1]);
                rfma(a, -wym, xt[-PX]);
                rfma(a, -wyp, xt[PX]);
                rfma(a, -wzm, xm[q]);
                rfma(a, -wzp, xp[q]);
                if (r0 + q < nrhs) {
                    const int64_t o = (int64_t)(r0 + q) * ld + p;
                    out[o] = xc[q] + dinv * (sb[q * Cfg::BT + bidx] - a);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            xm[q] = xc[q];
            xc[q] = xp[q];
        }
        __syncthreads();
        if (threadIdx.x == 0 && z + NS <= zl) issue(s, z + NS);
    }
}

// ---------------------------------------------------------------------------------------------
// The same fusion when the iterate that is corrected is the FIRST sweep from zero, x1 = dinv .* b (pre-smoothing count 1,
// the bench's W(1,2) / V(1,*) cycles):   x' = dinv .* b + P xc ;  out = x' + dinv .* (b - A x')
// x1 is recomputed from b and dinv (both staged WITH halo, as k_fine3d_tma_first stages them) instead of being written
// by the first kernel and read back here: the cycle start writes only the residual (2S instead of 3S per node and
// right-hand side) and this pass reads b once instead of x and b (2S + S/8 instead of 3S + S/8).  The staged b tile
// becomes the x' tile in place (the centre thread keeps b and dinv of its column in registers for the Jacobi update),
// the stage is 27 KB instead of 34 KB, so the ring is 4 deep at two CTAs per SM.  Values are those of the two-kernel
// path: x1 is formed by the same multiplication.
// ---------------------------------------------------------------------------------------------
template <typename T, int KB>
struct FineProBCfg {
    static constexpr int TX = 32, TY = 8;
    static constexpr int HX = sizeof(T) == 4 ? 2 : 1;  // see FineTmaCfg
    static constexpr int PX = TX + 2 * HX;
    static constexpr int XT = (TY + 2) * PX;
    static constexpr int BT = TY * TX;
    static constexpr int CH = sizeof(T) == 4 ? 2 : 1;
    static constexpr int CTX = sizeof(T) == 4 ? TX / 2 + 4 : TX / 2 + 2, CTY = TY / 2 + 2, CT = CTX * CTY;
    static constexpr int ES = (int)sizeof(cx<T>);
    static constexpr int al(int b) { return (b + 127) / 128 * 128; }
    static constexpr int OFF_B = 0;                            // b tiles with halo (become the x' tiles), KB right-hand sides
    static constexpr int OFF_D = al(KB * XT * ES);             // dinv tile with halo
    static constexpr int OFF_C = OFF_D + al(XT * ES);          // centre coefficient tile
    static constexpr int OFF_XC = OFF_C + al(BT * ES);         // two coarse planes, KB right-hand sides each
    static constexpr int XC_PLANE = al(KB * CT * ES);
    static constexpr int STAGE_BYTES = OFF_XC + 2 * XC_PLANE;
    static constexpr uint32_t TX_BYTES = KB * XT * ES + XT * ES + BT * ES + 2 * KB * CT * ES;
};

template <typename T, int KB, int NS, bool CACHE>
__global__ void __launch_bounds__(256) k_fine3d_tma_prob(FineOp<T> op, const __grid_constant__ TmaDesc tm_b,
                                                         const __grid_constant__ TmaDesc tm_d,
                                                         const __grid_constant__ TmaDesc tm_c,
                                                         const __grid_constant__ TmaDesc tm_xc,
                                                         const cx<T>* __restrict__ b, const cx<T>* __restrict__ xcg,
                                                         cx<T>* __restrict__ out, int64_t ld, int64_t ldc, int csy,
                                                         int nc1, int nrhs, int zchunk, int groups) {
    typedef FineProBCfg<T, KB> Cfg;
    constexpr int TX = Cfg::TX, TY = Cfg::TY, PX = Cfg::PX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * Cfg::STAGE_BYTES);
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = (blockIdx.x / groups) * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int r0 = (blockIdx.x % groups) * KB;
    const int z0 = op.zb + blockIdx.z * zchunk;
    const int z1 = min(op.ze, z0 + zchunk);
    const int zl = min(z1, n2 - 1);
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const bool active = (i < n0) && (j < n1);
    const int Is = (i0 >> 1) - Cfg::CH, Js = (j0 >> 1) - 1;  // origin of the coarse tile
    auto issue = [&](int s, int z) {
        unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&bars[s], Cfg::TX_BYTES);
        tma_load_4d(st + Cfg::OFF_B, &tm_b, 2 * (i0 - Cfg::HX), j0 - 1, z, r0, &bars[s]);
        tma_load_3d(st + Cfg::OFF_D, &tm_d, 2 * (i0 - Cfg::HX), j0 - 1, z, &bars[s]);
        tma_load_3d(st + Cfg::OFF_C, &tm_c, 2 * i0, j0, z, &bars[s]);
        tma_load_4d(st + Cfg::OFF_XC, &tm_xc, 2 * Is, Js, z >> 1, r0, &bars[s]);
        tma_load_4d(st + Cfg::OFF_XC + Cfg::XC_PLANE, &tm_xc, 2 * Is, Js, (z >> 1) + 1, r0, &bars[s]);
    };
    // interpolation geometry: as in k_fine3d_tma_pro (the centre cell of this thread, and one cell of the halo ring for the
    // first 84*KB threads)
    auto geom = [&](int fi, int fj, int q, int& coff, int& par) {
        coff = 0;
        par = 0;
        if ((unsigned)fi < (unsigned)n0 && (unsigned)fj < (unsigned)n1) {
            coff = q * Cfg::CT + ((fj >> 1) - Js) * Cfg::CTX + ((fi >> 1) - Is);
            par = 4 | (fi & 1) | ((fj & 1) << 1);
        }
    };
    auto interp = [&](const cx<T>* c0, int coff, int par, int ok) -> cx<T> {
        const int oi = par & 1, oj = (par >> 1) & 1;
        const cx<T>* p0 = c0 + coff;
        cx<T> acc = p0[0];
        if (oi) acc = acc + p0[1];
        if (oj) {
            acc = acc + p0[Cfg::CTX];
            if (oi) acc = acc + p0[Cfg::CTX + 1];
        }
        if (ok) {
            const cx<T>* p1 = p0 + Cfg::XC_PLANE / Cfg::ES;
            acc = acc + p1[0];
            if (oi) acc = acc + p1[1];
            if (oj) {
                acc = acc + p1[Cfg::CTX];
                if (oi) acc = acc + p1[Cfg::CTX + 1];
            }
        }
        return (T(1) / T(1 << (oi + oj + ok))) * acc;
    };
    const int cidx = (ty + 1) * PX + (tx + Cfg::HX);
    const int bidx = ty * TX + tx;
    int ccoff[KB], cpar;  // centre
    {
        int par0 = 0;
#pragma unroll
        for (int q = 0; q < KB; ++q) geom(i, j, q, ccoff[q], par0);
        cpar = par0;
    }
    int hoff = -1, hdoff = 0, hcoff = 0, hpar = 0;  // halo-ring cell of this thread (if any)
    if (threadIdx.x < 84 * KB) {
        const int q = threadIdx.x / 84, t = threadIdx.x - q * 84;
        constexpr int RW = TX + 2;  // width of the one-node ring rows
        int row, col;              // col counted from the node i0-1
        if (t < RW) {
            row = 0;
            col = t;
        } else if (t < 2 * RW) {
            row = TY + 1;
            col = t - RW;
        } else if (t < 2 * RW + TY) {
            row = 1 + (t - 2 * RW);
            col = 0;
        } else {
            row = 1 + (t - 2 * RW - TY);
            col = TX + 1;
        }
        geom(i0 - 1 + col, j0 - 1 + row, q, hcoff, hpar);
        if (hpar & 4) {
            hdoff = row * PX + (col + Cfg::HX - 1);
            hoff = q * Cfg::XT + hdoff;
        }
    }
    // CACHE: the in-plane part of the interpolation of a coarse plane K serves the fine planes 2K-1, 2K, 2K+1, so every
    // thread keeps it in registers for its cells -- I_K while the march is at coarse plane K = z >> 1, and on odd planes
    // I_{K+1}, which becomes I_K two planes later: 1.1 instead of 3.4 shared-memory loads per cell and plane.
    auto inplane = [&](const cx<T>* c0, int coff, int par, int pl) -> cx<T> {
        const int oi = par & 1, oj = (par >> 1) & 1;
        const cx<T>* p0 = c0 + coff + pl * (Cfg::XC_PLANE / Cfg::ES);
        cx<T> acc = p0[0];
        if (oi) acc = acc + p0[1];
        if (oj) {
            acc = acc + p0[Cfg::CTX];
            if (oi) acc = acc + p0[Cfg::CTX + 1];
        }
        return (T(1) / T(1 << (oi + oj))) * acc;
    };
    cx<T> ik[KB], inx[KB], hik = mk<T>(T(0), T(0)), hinx = mk<T>(T(0), T(0));
#pragma unroll
    for (int q = 0; q < KB; ++q) ik[q] = inx[q] = mk<T>(T(0), T(0));
    auto cached = [&](const cx<T>* c0, int coff, int par, int ok, bool first, cx<T>& cur, cx<T>& nxt) -> cx<T> {
        if (first) cur = inplane(c0, coff, par, 0);
        else if (!ok) cur = nxt;  // an even plane after an odd one: the coarse plane advanced
        if (!ok) return cur;
        nxt = inplane(c0, coff, par, 1);
        return T(0.5) * (cur + nxt);
    };
    // plane z of stage st: the b tile becomes x' = dinv .* b + P xc; xv / bv / dv = x', b, dinv of this thread's column
    auto correct_plane = [&](unsigned char* st, int z, cx<T>* xv, cx<T>* bv, cx<T>& dv, bool first) {
        cx<T>* xs = reinterpret_cast<cx<T>*>(st + Cfg::OFF_B);
        const cx<T>* sd = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_D);
        const cx<T>* c0 = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_XC);
        const int ok = z & 1;
        dv = sd[cidx];
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            const cx<T> bq = xs[q * Cfg::XT + cidx];
            cx<T> v = dv * bq;
            if (cpar & 4) v = v + (CACHE ? cached(c0, ccoff[q], cpar, ok, first, ik[q], inx[q]) : interp(c0, ccoff[q], cpar, ok));
            xs[q * Cfg::XT + cidx] = v;
            xv[q] = v;
            bv[q] = bq;
        }
        if (hoff >= 0)
            xs[hoff] = sd[hdoff] * xs[hoff] + (CACHE ? cached(c0, hcoff, hpar, ok, first, hik, hinx) : interp(c0, hcoff, hpar, ok));
        fence_proxy_async();  // these generic stores precede the TMA refill of this stage (after a CTA barrier)
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NS && z0 + s <= zl; ++s) issue(s, z0 + s);
    }
    __syncthreads();
    const int ic = active ? i : 0, jc = active ? j : 0;
    const T wxm = fine_w(op, 0, 0, ic, n0), wxp = fine_w(op, 0, 1, ic, n0);
    const T wym = fine_w(op, 1, 0, jc, n1), wyp = fine_w(op, 1, 1, jc, n1);
    const int64_t pxy = ic + sy * jc;
    cx<T> xm[KB], xc[KB], xp[KB], bc[KB], bp[KB];
    cx<T> dc, dp = mk<T>(T(0), T(0));
    mbar_wait(&bars[0], 0);
    correct_plane(smem_raw, z0, xc, bc, dc, true);
    __syncthreads();  // once per chunk: plane z0 is used in the first iteration already
#pragma unroll
    for (int q = 0; q < KB; ++q) {
        xm[q] = mk<T>(T(0), T(0));
        bp[q] = mk<T>(T(0), T(0));
        if (z0 > 0 && active) {
            const int r = min(r0 + q, nrhs - 1);
            const int64_t pm = pxy + (int64_t)(z0 - 1) * sz;
            xm[q] = op.dinv[pm] * b[(int64_t)r * ld + pm] + prolong_point<T>(xcg + (int64_t)r * ldc, ic, jc, z0 - 1, csy, nc1);
        }
    }
#pragma unroll 1
    for (int z = z0; z < z1; ++z) {
        const int s = (z - z0) % NS;
        const unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        const bool zlast = (z == n2 - 1);
        if (!zlast) {
            const int s1 = (z + 1 - z0) % NS;
            mbar_wait(&bars[s1], (uint32_t)(((z + 1 - z0) / NS) & 1));
            correct_plane(smem_raw + (size_t)s1 * Cfg::STAGE_BYTES, z + 1, xp, bp, dp, false);
        } else {
#pragma unroll
            for (int q = 0; q < KB; ++q) xp[q] = mk<T>(T(0), T(0));
        }
        if (active) {
            const cx<T>* sx = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_B);
            const cx<T> c = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_C)[bidx];
            const T wzm = fine_wz(op, 0, z), wzp = fine_wz(op, 1, z);
            const int64_t p = pxy + (int64_t)z * sz;
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                const cx<T>* xt = sx + q * Cfg::XT + cidx;
                cx<T> a = c * xc[q];
                rfma(a, -wxm, xt[-1]);
                rfma(a, -wxp, xt[1]);
                rfma(a, -wym, xt[-PX]);
                rfma(a, -wyp, xt[PX]);
                rfma(a, -wzm, xm[q]);
                rfma(a, -wzp, xp[q]);
                if (r0 + q < nrhs) {
                    const int64_t o = (int64_t)(r0 + q) * ld + p;
                    out[o] = xc[q] + dc * (bc[q] - a);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            xm[q] = xc[q];
            xc[q] = xp[q];
            bc[q] = bp[q];
        }
        dc = dp;
        __syncthreads();
        if (threadIdx.x == 0 && z + NS <= zl) issue(s, z + NS);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused coarse-grid correction + TWO post-smoothing sweeps on the fine level (3-D, TMA form):
//   x' = x + P xc ;  x1 = x' + dinv .* (b - A x') ;  out = x1 + dinv .* (b - A x1)
// Neither x' nor x1 touches HBM: per node and right-hand side the pass reads x, b (and xc/8) and writes x2 --
// 3.125 S instead of the 6.125 S of k_fine3d_tma_pro followed by k_fine3d_tma<JACOBI>.  The price is a two-node
// halo: the x tile is staged with two halo nodes per side, corrected in shared memory (all 36 x 12 entries), x1 is
// formed on the tile plus a one-node ring (the ring nodes by the first 84*KB threads) into a two-slot plane buffer in
// shared memory, and x2 on the tile proper.  In z the kernel runs three planes deep: when plane p arrives it
// completes x'(p), x1(p-1) and x2(p-2); a chunk of output planes [z0,z1) therefore stages planes z0-2 .. z1+1.
// The z-neighbours of a column (x' and x1 of the planes before) ride in registers, b / c / dinv of the plane a
// thread has just used for x1 are carried in registers for its x2 one iteration later, so only two stages are
// live at a time.  Not used under slab decomposition (the halo planes are one deep there).
// ---------------------------------------------------------------------------------------------
template <typename T, int KB>
struct FinePro2Cfg {
    static constexpr int TX = 32, TY = 8;
    static constexpr int H2 = 2;                        // halo of the x tile (even: 16-byte aligned box start in both precisions)
    static constexpr int PX2 = TX + 2 * H2, PY2 = TY + 2 * H2, XT2 = PX2 * PY2;
    static constexpr int H1 = sizeof(T) == 4 ? 2 : 1;   // x halo of the b / c / dinv tiles and of the x1 plane buffer
    static constexpr int PX1 = TX + 2 * H1, PY1 = TY + 2, ET1 = PX1 * PY1;
    static constexpr int CH = sizeof(T) == 4 ? 2 : 1;   // the coarse tile starts CH coarse nodes left of i0/2
    static constexpr int CTX = TX / 2 + 2 + CH, CTY = TY / 2 + 3, CT = CTX * CTY;
    static constexpr int ES = (int)sizeof(cx<T>);
    static constexpr int al(int b) { return (b + 127) / 128 * 128; }
    static constexpr int OFF_X = 0;
    static constexpr int OFF_B = al(KB * XT2 * ES);
    static constexpr int OFF_C = OFF_B + al(KB * ET1 * ES);
    static constexpr int OFF_D = OFF_C + al(ET1 * ES);
    static constexpr int OFF_XC = OFF_D + al(ET1 * ES);
    static constexpr int XC_PLANE = al(KB * CT * ES);
    static constexpr int STAGE_BYTES = OFF_XC + 2 * XC_PLANE;
    static constexpr uint32_t TX_BYTES = KB * XT2 * ES + KB * ET1 * ES + 2 * ET1 * ES + 2 * KB * CT * ES;
    static constexpr int X1_SLOT = al(KB * ET1 * ES);   // x1 of one plane on the tile + ring
    static constexpr int NS = (4 * STAGE_BYTES + 2 * X1_SLOT + 64 <= 227 * 1024) ? 4 : 3;
    static constexpr int NCORR = (KB * XT2 + 255) / 256;  // entries of the x tile each thread corrects
    static constexpr int RING = 2 * (TX + 2) + 2 * TY;    // nodes of the one-node ring (84)
    static_assert((PX2 * ES) % 16 == 0 && (PX1 * ES) % 16 == 0 && (CTX * ES) % 16 == 0, "TMA box rows are 16-byte multiples");
    static_assert(RING * KB <= 256, "one thread per ring node and right-hand side");
};

template <typename T, int KB>
__global__ void __launch_bounds__(256, 1) k_fine3d_tma_pro2(FineOp<T> op, const __grid_constant__ TmaDesc tm_x,
                                                            const __grid_constant__ TmaDesc tm_b,
                                                            const __grid_constant__ TmaDesc tm_c,
                                                            const __grid_constant__ TmaDesc tm_d,
                                                            const __grid_constant__ TmaDesc tm_xc, cx<T>* __restrict__ out,
                                                            int64_t ld, int nrhs, int zchunk, int groups) {
    typedef FinePro2Cfg<T, KB> Cfg;
    constexpr int TX = Cfg::TX, TY = Cfg::TY, PX2 = Cfg::PX2, PX1 = Cfg::PX1, NS = Cfg::NS, H1 = Cfg::H1, H2 = Cfg::H2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* x1buf = smem_raw + (size_t)NS * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(x1buf + 2 * Cfg::X1_SLOT);
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = (blockIdx.x / groups) * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int r0 = (blockIdx.x % groups) * KB;
    const int z0 = op.zb + blockIdx.z * zchunk;
    const int z1 = min(op.ze, z0 + zchunk);
    const int pf = z0 - 2, pl = z1 + 1;  // first / last staged plane
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const bool active = (i < n0) && (j < n1);
    const int Is = (i0 >> 1) - Cfg::CH, Js = (j0 >> 1) - 1;  // origin of the coarse tile
    auto issue = [&](int s, int p) {
        unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&bars[s], Cfg::TX_BYTES);
        tma_load_4d(st + Cfg::OFF_X, &tm_x, 2 * (i0 - H2), j0 - H2, p, r0, &bars[s]);
        tma_load_4d(st + Cfg::OFF_B, &tm_b, 2 * (i0 - H1), j0 - 1, p, r0, &bars[s]);
        tma_load_3d(st + Cfg::OFF_C, &tm_c, 2 * (i0 - H1), j0 - 1, p, &bars[s]);
        tma_load_3d(st + Cfg::OFF_D, &tm_d, 2 * (i0 - H1), j0 - 1, p, &bars[s]);
        tma_load_4d(st + Cfg::OFF_XC, &tm_xc, 2 * Is, Js, p >> 1, r0, &bars[s]);
        tma_load_4d(st + Cfg::OFF_XC + Cfg::XC_PLANE, &tm_xc, 2 * Is, Js, (p >> 1) + 1, r0, &bars[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NS && pf + s <= pl; ++s) issue(s, pf + s);
    }
    // ---- correction geometry of the x-tile entries this thread owns (z-invariant) ----
    int coff[Cfg::NCORR], cinfo[Cfg::NCORR];  // x-tile offset (or -1), coarse-tile offset << 2 | parity bits
#pragma unroll
    for (int k = 0; k < Cfg::NCORR; ++k) {
        const int e = threadIdx.x + 256 * k;
        coff[k] = -1;
        cinfo[k] = 0;
        if (e < KB * Cfg::XT2) {
            const int q = e / Cfg::XT2, r = e - q * Cfg::XT2;
            const int row = r / PX2, col = r - row * PX2;
            const int fi = i0 - H2 + col, fj = j0 - H2 + row;
            if ((unsigned)fi < (unsigned)n0 && (unsigned)fj < (unsigned)n1) {
                coff[k] = e;
                cinfo[k] = ((q * Cfg::CT + ((fj >> 1) - Js) * Cfg::CTX + ((fi >> 1) - Is)) << 2) | (fi & 1) | ((fj & 1) << 1);
            }
        }
    }
    auto interp = [&](const cx<T>* c0, int info, int ok) -> cx<T> {
        const int oi = info & 1, oj = (info >> 1) & 1;
        const cx<T>* p0 = c0 + (info >> 2);
        cx<T> acc = p0[0];
        if (oi) acc = acc + p0[1];
        if (oj) {
            acc = acc + p0[Cfg::CTX];
            if (oi) acc = acc + p0[Cfg::CTX + 1];
        }
        if (ok) {
            const cx<T>* p1 = p0 + Cfg::XC_PLANE / Cfg::ES;
            acc = acc + p1[0];
            if (oi) acc = acc + p1[1];
            if (oj) {
                acc = acc + p1[Cfg::CTX];
                if (oi) acc = acc + p1[Cfg::CTX + 1];
            }
        }
        return (T(1) / T(1 << (oi + oj + ok))) * acc;
    };
    // ---- the ring node of this thread (threads < 84*KB): one node of the one-node ring, one right-hand side ----
    const bool ring_thread = threadIdx.x < Cfg::RING * KB;
    int rq = 0, ri = 0, rj = 0, rrow = 0, rcol = 0;  // RHS slot, fine coordinates, position relative to (i0-1, j0-1)
    if (ring_thread) {
        rq = threadIdx.x / Cfg::RING;
        const int u = threadIdx.x - rq * Cfg::RING;
        constexpr int RW = TX + 2;
        if (u < RW) {
            rrow = 0;
            rcol = u;
        } else if (u < 2 * RW) {
            rrow = TY + 1;
            rcol = u - RW;
        } else if (u < 2 * RW + TY) {
            rrow = 1 + (u - 2 * RW);
            rcol = 0;
        } else {
            rrow = 1 + (u - 2 * RW - TY);
            rcol = TX + 1;
        }
        ri = i0 - 1 + rcol;
        rj = j0 - 1 + rrow;
    }
    const bool ring_ok = ring_thread && (unsigned)ri < (unsigned)n0 && (unsigned)rj < (unsigned)n1;
    // offsets of the two roles inside an x tile (two-node halo) and inside an E1 tile (b, c, dinv, x1 buffer)
    const int xi_c = (ty + H2) * PX2 + (tx + H2), e1_c = (ty + 1) * PX1 + (tx + H1);
    const int xi_r = (rrow + H2 - 1) * PX2 + (rcol + H2 - 1) + rq * Cfg::XT2, e1_r = rrow * PX1 + (rcol + H1 - 1);
    const int ic = active ? i : 0, jc = active ? j : 0;
    const T wxm = fine_w(op, 0, 0, ic, n0), wxp = fine_w(op, 0, 1, ic, n0);
    const T wym = fine_w(op, 1, 0, jc, n1), wyp = fine_w(op, 1, 1, jc, n1);
    const int ir = ring_ok ? ri : 0, jr = ring_ok ? rj : 0;
    const T rwxm = fine_w(op, 0, 0, ir, n0), rwxp = fine_w(op, 0, 1, ir, n0);
    const T rwym = fine_w(op, 1, 0, jr, n1), rwyp = fine_w(op, 1, 1, jr, n1);
    const int64_t pxy = ic + sy * jc;
    const cx<T> zero = mk<T>(T(0), T(0));
    cx<T> xa[KB], xb[KB], y3[KB], y2[KB], b2[KB];  // x'(p-2), x'(p-1); x1(p-3), x1(p-2); b(p-2) of the own column
    cx<T> c2 = zero, d2 = zero;                    // c(p-2), dinv(p-2)
    cx<T> ra = zero, rb = zero;                    // x'(p-2), x'(p-1) of the ring node
#pragma unroll
    for (int q = 0; q < KB; ++q) xa[q] = xb[q] = y3[q] = y2[q] = b2[q] = zero;
    __syncthreads();
#pragma unroll 1
    for (int p = pf; p <= pl; ++p) {
        const int it = p - pf;
        const int s = it % NS;
        unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        const unsigned char* stm = smem_raw + (size_t)((it + NS - 1) % NS) * Cfg::STAGE_BYTES;  // stage of plane p-1
        mbar_wait(&bars[s], (uint32_t)((it / NS) & 1));
        // (1) x'(p) = x(p) + (P xc)(p) on the whole staged tile (planes outside the grid stay zero)
        cx<T>* xs = reinterpret_cast<cx<T>*>(st + Cfg::OFF_X);
        if ((unsigned)p < (unsigned)n2) {
            const cx<T>* cc = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_XC);
            const int ok = p & 1;
#pragma unroll
            for (int k = 0; k < Cfg::NCORR; ++k)
                if (coff[k] >= 0) xs[coff[k]] = xs[coff[k]] + interp(cc, cinfo[k], ok);
        }
        fence_proxy_async();  // generic stores into a stage the TMA engine refills later
        __syncthreads();
        // (2) x1(p-1) on the tile and its ring, from x'(p-2) [registers], x'(p-1) [stage p-1], x'(p) [stage p]
        cx<T> y1[KB];
#pragma unroll
        for (int q = 0; q < KB; ++q) y1[q] = zero;
        cx<T> bn[KB], cn = zero, dn = zero;  // b, c, dinv of plane p-1 at the own node
#pragma unroll
        for (int q = 0; q < KB; ++q) bn[q] = zero;
        cx<T>* x1w = reinterpret_cast<cx<T>*>(x1buf + (size_t)((p - 1) & 1) * Cfg::X1_SLOT);
        if (it >= 2) {  // planes p-2, p-1, p are all staged
            const cx<T>* xm1 = reinterpret_cast<const cx<T>*>(stm + Cfg::OFF_X);
            const cx<T>* sb = reinterpret_cast<const cx<T>*>(stm + Cfg::OFF_B);
            const cx<T>* sc = reinterpret_cast<const cx<T>*>(stm + Cfg::OFF_C);
            const cx<T>* sd = reinterpret_cast<const cx<T>*>(stm + Cfg::OFF_D);
            const T wzm = fine_wz(op, 0, p - 1), wzp = fine_wz(op, 1, p - 1);
            cn = sc[e1_c];
            dn = sd[e1_c];
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                const cx<T>* xt = xm1 + q * Cfg::XT2 + xi_c;
                bn[q] = sb[q * Cfg::ET1 + e1_c];
                cx<T> a = cn * xb[q];
                rfma(a, -wxm, xt[-1]);
                rfma(a, -wxp, xt[1]);
                rfma(a, -wym, xt[-PX2]);
                rfma(a, -wyp, xt[PX2]);
                rfma(a, -wzm, xa[q]);
                rfma(a, -wzp, xs[q * Cfg::XT2 + xi_c]);
                y1[q] = xb[q] + dn * (bn[q] - a);
                x1w[q * Cfg::ET1 + e1_c] = y1[q];
            }
            if (ring_thread) {
                cx<T> v = zero;
                if (ring_ok) {
                    const cx<T>* xt = xm1 + xi_r;
                    cx<T> a = sc[e1_r] * rb;
                    rfma(a, -rwxm, xt[-1]);
                    rfma(a, -rwxp, xt[1]);
                    rfma(a, -rwym, xt[-PX2]);
                    rfma(a, -rwyp, xt[PX2]);
                    rfma(a, -wzm, ra);
                    rfma(a, -wzp, xs[xi_r]);
                    v = rb + sd[e1_r] * (sb[rq * Cfg::ET1 + e1_r] - a);
                }
                x1w[rq * Cfg::ET1 + e1_r] = v;
            }
        }
        // (3) x2(p-2) on the tile, from x1(p-3), x1(p-2) [registers; neighbours in the other x1 slot], x1(p-1)
        const int zo = p - 2;
        if (active && zo >= z0 && zo < z1) {
            const cx<T>* x1r = reinterpret_cast<const cx<T>*>(x1buf + (size_t)(zo & 1) * Cfg::X1_SLOT) + e1_c;
            const T wzm = fine_wz(op, 0, zo), wzp = fine_wz(op, 1, zo);
            const int64_t pn = pxy + (int64_t)zo * sz;
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                const cx<T>* yt = x1r + q * Cfg::ET1;
                cx<T> a = c2 * y2[q];
                rfma(a, -wxm, yt[-1]);
                rfma(a, -wxp, yt[1]);
                rfma(a, -wym, yt[-PX1]);
                rfma(a, -wyp, yt[PX1]);
                rfma(a, -wzm, y3[q]);
                rfma(a, -wzp, y1[q]);
                if (r0 + q < nrhs) out[(int64_t)(r0 + q) * ld + pn] = y2[q] + d2 * (b2[q] - a);
            }
        }
        // roll the registers
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            xa[q] = xb[q];
            xb[q] = xs[q * Cfg::XT2 + xi_c];
            y3[q] = y2[q];
            y2[q] = y1[q];
            b2[q] = bn[q];
        }
        c2 = cn;
        d2 = dn;
        if (ring_thread) {
            ra = rb;
            rb = xs[xi_r];
        }
        __syncthreads();  // stage of plane p-1 and the x1 slot just read are free
        if (threadIdx.x == 0 && it >= 1 && p - 1 + NS <= pl) issue((it + NS - 1) % NS, p - 1 + NS);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused start of a cycle on the fine level (3-D, TMA form): the first Jacobi sweep from a zero guess,
// x1 = dinv .* b, is never written and re-read -- the kernel stages b and dinv WITH halo, forms x1 on
// the fly at the centre and the six neighbours, and directly produces
//   SECOND = 0 :  x1 and the residual r = b - A x1          (pre-smoothing count 1: x1, r in one pass)
//   SECOND = 1 :  x2 = x1 + dinv .* (b - A x1)               (the first two sweeps in one pass)
// Algorithmic bytes per node per RHS: S read + 2S / S written, instead of 2S + 3S for two kernels.
// ---------------------------------------------------------------------------------------------
template <typename T, int KB>
struct FineFirstCfg {
    static constexpr int TX = 32, TY = 8;
    static constexpr int HX = sizeof(T) == 4 ? 2 : 1;  // see FineTmaCfg
    static constexpr int PX = TX + 2 * HX;
    static constexpr int XT = (TY + 2) * PX;
    static constexpr int BT = TY * TX;
    static constexpr int ES = (int)sizeof(cx<T>);
    static constexpr int al(int b) { return (b + 127) / 128 * 128; }
    static constexpr int OFF_B = 0;                       // b tiles with halo, KB right-hand sides
    static constexpr int OFF_D = al(KB * XT * ES);        // dinv tile with halo
    static constexpr int OFF_C = OFF_D + al(XT * ES);     // centre coefficient tile
    static constexpr int STAGE_BYTES = OFF_C + al(BT * ES);
    static constexpr uint32_t TX_BYTES = KB * XT * ES + XT * ES + BT * ES;
};

template <typename T, int SECOND, int KB, int NS>
__global__ void __launch_bounds__(256) k_fine3d_tma_first(FineOp<T> op, const __grid_constant__ TmaDesc tm_b,
                                                          const __grid_constant__ TmaDesc tm_d,
                                                          const __grid_constant__ TmaDesc tm_c,
                                                          const cx<T>* __restrict__ b, cx<T>* __restrict__ out,
                                                          cx<T>* __restrict__ out2, int64_t ld, int nrhs, int zchunk,
                                                          int groups) {
    typedef FineFirstCfg<T, KB> Cfg;
    constexpr int TX = Cfg::TX, TY = Cfg::TY, PX = Cfg::PX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * Cfg::STAGE_BYTES);
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = (blockIdx.x / groups) * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int r0 = (blockIdx.x % groups) * KB;
    const int z0 = op.zb + blockIdx.z * zchunk;
    const int z1 = min(op.ze, z0 + zchunk);
    const int zl = min(z1, n2 - 1);
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const bool active = (i < n0) && (j < n1);
    auto issue = [&](int s, int z) {
        unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&bars[s], Cfg::TX_BYTES);
        tma_load_4d(st + Cfg::OFF_B, &tm_b, 2 * (i0 - Cfg::HX), j0 - 1, z, r0, &bars[s]);
        tma_load_3d(st + Cfg::OFF_D, &tm_d, 2 * (i0 - Cfg::HX), j0 - 1, z, &bars[s]);
        tma_load_3d(st + Cfg::OFF_C, &tm_c, 2 * i0, j0, z, &bars[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NS && z0 + s <= zl; ++s) issue(s, z0 + s);
    }
    __syncthreads();
    const int ic = active ? i : 0, jc = active ? j : 0;
    const T wxm = fine_w(op, 0, 0, ic, n0), wxp = fine_w(op, 0, 1, ic, n0);
    const T wym = fine_w(op, 1, 0, jc, n1), wyp = fine_w(op, 1, 1, jc, n1);
    const int64_t pxy = ic + sy * jc;
    const int cidx = (ty + 1) * PX + (tx + Cfg::HX);
    const int bidx = ty * TX + tx;
    cx<T> tm_[KB], tc[KB], tp[KB];  // x1 = dinv .* b at planes z-1, z, z+1 of this column
    mbar_wait(&bars[0], 0);
    {
        const cx<T> d0 = reinterpret_cast<const cx<T>*>(smem_raw + Cfg::OFF_D)[cidx];
        const cx<T> dm = (z0 > 0 && active) ? op.dinv[pxy + (int64_t)(z0 - 1) * sz] : mk<T>(T(0), T(0));
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            tc[q] = d0 * reinterpret_cast<const cx<T>*>(smem_raw + Cfg::OFF_B)[q * Cfg::XT + cidx];
            const int r = min(r0 + q, nrhs - 1);
            tm_[q] = (z0 > 0 && active) ? dm * b[(int64_t)r * ld + pxy + (int64_t)(z0 - 1) * sz] : mk<T>(T(0), T(0));
        }
    }
#pragma unroll 1
    for (int z = z0; z < z1; ++z) {
        const int s = (z - z0) % NS;
        const unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        const bool zlast = (z == n2 - 1);
        if (!zlast) {
            const int s1 = (z + 1 - z0) % NS;
            mbar_wait(&bars[s1], (uint32_t)(((z + 1 - z0) / NS) & 1));
            const unsigned char* st1 = smem_raw + (size_t)s1 * Cfg::STAGE_BYTES;
            const cx<T> d1 = reinterpret_cast<const cx<T>*>(st1 + Cfg::OFF_D)[cidx];
#pragma unroll
            for (int q = 0; q < KB; ++q) tp[q] = d1 * reinterpret_cast<const cx<T>*>(st1 + Cfg::OFF_B)[q * Cfg::XT + cidx];
        } else {
#pragma unroll
            for (int q = 0; q < KB; ++q) tp[q] = mk<T>(T(0), T(0));
        }
        if (active) {
            const cx<T>* sb = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_B);
            const cx<T>* sd = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_D) + cidx;
            const cx<T> c = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_C)[bidx];
            const cx<T> dW = sd[-1], dE = sd[1], dS = sd[-PX], dN = sd[PX], dC = sd[0];
            const T wzm = fine_wz(op, 0, z), wzp = fine_wz(op, 1, z);
            const int64_t p = pxy + (int64_t)z * sz;
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                const cx<T>* bt = sb + q * Cfg::XT + cidx;
                cx<T> a = c * tc[q];
                rfma(a, -wxm, dW * bt[-1]);  // halo cells outside the grid are zero-filled (b and dinv)
                rfma(a, -wxp, dE * bt[1]);
                rfma(a, -wym, dS * bt[-PX]);
                rfma(a, -wyp, dN * bt[PX]);
                rfma(a, -wzm, tm_[q]);
                rfma(a, -wzp, tp[q]);
                if (r0 + q < nrhs) {
                    const int64_t o = (int64_t)(r0 + q) * ld + p;
                    const cx<T> res = bt[0] - a;
                    if (SECOND == 0) {
                        if (out != nullptr) out[o] = tc[q];  // nullptr: x1 is recomputed by k_fine3d_tma_prob
                        out2[o] = res;
                    } else {
                        out[o] = tc[q] + dC * res;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            tm_[q] = tc[q];
            tc[q] = tp[q];
        }
        __syncthreads();
        if (threadIdx.x == 0 && z + NS <= zl) issue(s, z + NS);
    }
}

// ---------------------------------------------------------------------------------------------
// Coarse levels, 3-D, TMA production form: 27-point stencil with stored Galerkin coefficients.
// A CTA owns a TX x TY tile of (i,j) columns (TX*TY <= 128) and walks a chunk of z planes in the
// scatter form of k_coarse3d_zmarch (three rolling accumulators per right-hand side).  Each
// iteration consumes one stage staged by the TMA engine: the x tile of input plane zi with a
// one-node halo for all KB right-hand sides, the three 9-coefficient sets that plane feeds (dk = -1
// of plane zi+1, dk = 0 of plane zi, dk = +1 of plane zi-1), and b / dinv of the output plane zi-1
// that completes in this iteration.  Out-of-range planes, halos and surplus RHS slots are zero-filled
// by the hardware, so the loop has no boundary cases.  The 27 coefficients of a node are read once
// per KB (up to 8) right-hand sides.
//
// The kernel is bound by the FP64 pipe and the shared-memory pipe together (108 DFMA and ~12 LDS.128
// per node and right-hand side), not by HBM, unless both pipes stay busy: with KB >= 2 the CTA runs
// TWO warpgroups of 128 threads on the same staged tile, each taking KB/2 right-hand sides -- 8 warps
// per SM (2 per scheduler) instead of 4, half the accumulator registers per thread -- and the tile
// shape is a launch-time choice (hh_solver.cuh: coarse_tile) so that 2^k+1 grids do not leave most of
// the last tile column idle (129 = 4*32+1 wastes 19 % of a 32-wide tiling, 9 % of an 11-wide one).
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int KB, int TX_, int TY_>
struct CoarseTmaCfg {
    static constexpr int TX = TX_, TY = TY_;
    static constexpr int NWG = KB >= 2 ? 2 : 1;   // warpgroups (128 threads each) working on one staged tile
    static constexpr int KH = KB / NWG;           // right-hand sides per warpgroup
    static constexpr int THREADS = 128 * NWG;
    static constexpr int HX = sizeof(T) == 4 ? 2 : 1;  // see FineTmaCfg
    // Row pitch of the staged x tile.  Threads are numbered row-major over the TX-wide tile, so a quarter-warp (the unit
    // of a 16-byte shared load) that straddles two tile rows is conflict-free only if the pitch exceeds TX by a multiple
    // of 8 entries (128 bytes): the 11-wide ComplexF64 tile pads its rows from 13 to 19 entries (the box simply fetches
    // six more columns; ncu showed 26 % of the shared wavefronts of the unpadded tile were bank conflicts and the shared
    // pipe at 77 % -- the binding pipe of this kernel).
    static constexpr int PADX = (sizeof(T) == 8 && (TX % 8) != 0) ? (8 - (2 * HX) % 8) % 8 : 0;
    static constexpr int PX = TX + 2 * HX + PADX;
    static constexpr int XT = (TY + 2) * PX;
    static constexpr int BT = TY * TX;
    static constexpr int ES = (int)sizeof(cx<T>);
    static constexpr int al(int b) { return (b + 127) / 128 * 128; }
    static constexpr int OFF_X = 0;
    static constexpr int OFF_C = al(KB * XT * ES);             // 3 sets of 9 coefficient tiles
    static constexpr int CSET = al(9 * BT * ES);
    static constexpr int OFF_B = OFF_C + 3 * CSET;
    static constexpr int OFF_D = OFF_B + (MODE != MODE_APPLY ? al(KB * BT * ES) : 0);
    static constexpr int STAGE_BYTES = OFF_D + (MODE == MODE_JACOBI ? al(BT * ES) : 0);
    static constexpr uint32_t TX_BYTES = KB * XT * ES + 27 * BT * ES + (MODE != MODE_APPLY ? KB * BT * ES : 0) +
                                         (MODE == MODE_JACOBI ? BT * ES : 0);
    static constexpr int NS = (3 * STAGE_BYTES + 64 <= 227 * 1024) ? 3 : 2;
    static_assert(TX * TY <= 128, "a warpgroup covers the tile with one thread per column");
    static_assert((TX * ES) % 16 == 0 && (PX * ES) % 16 == 0, "TMA box rows are 16-byte multiples");
    static_assert(2 * STAGE_BYTES + 64 <= 227 * 1024, "two stages must fit in shared memory");
};

// one input plane scattered into the three rolling accumulators: acc[2] += c(dk=-1) x, acc[1] += c(dk=0) x,
// acc[0] += c(dk=+1) x over the 3 x 3 in-plane neighbourhood (sx, cm, c0, cp already point at this thread's entries)
template <typename T, int KH, int XT, int BT, int PX, bool UM, bool U0, bool UP>
__device__ __forceinline__ void coarse_scatter(cx<T> (&acc)[3][KH], const cx<T>* __restrict__ sx, const cx<T>* __restrict__ cm,
                                               const cx<T>* __restrict__ c0, const cx<T>* __restrict__ cp) {
#pragma unroll
    for (int dj = -1; dj <= 1; ++dj) {
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
            const int sxy = (di + 1) + 3 * (dj + 1);
            cx<T> fm, f0, fp;
            if (UM) fm = cm[sxy * BT];
            if (U0) f0 = c0[sxy * BT];
            if (UP) fp = cp[sxy * BT];
#pragma unroll
            for (int q = 0; q < KH; ++q) {
                const cx<T> xv = sx[q * XT + di + dj * PX];
                if (UM) cfma(acc[2][q], fm, xv);
                if (U0) cfma(acc[1][q], f0, xv);
                if (UP) cfma(acc[0][q], fp, xv);
            }
        }
    }
}

template <typename T, int MODE, int KB, int TX_, int TY_>
__global__ void __launch_bounds__(CoarseTmaCfg<T, MODE, KB, TX_, TY_>::THREADS, 1)
    k_coarse3d_tma(const __grid_constant__ TmaDesc tm_x, const __grid_constant__ TmaDesc tm_c,
                   const __grid_constant__ TmaDesc tm_b, const __grid_constant__ TmaDesc tm_d, cx<T>* __restrict__ out,
                   int n0, int n1, int n2, int sy, int64_t ld, int nrhs, int zchunk, int groups, int zb, int ze) {
    typedef CoarseTmaCfg<T, MODE, KB, TX_, TY_> Cfg;
    constexpr int TX = Cfg::TX, TY = Cfg::TY, PX = Cfg::PX, NS = Cfg::NS, KH = Cfg::KH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * Cfg::STAGE_BYTES);
    const int wg = threadIdx.x >> 7;                     // warpgroup: right-hand sides q0 .. q0+KH-1 of the group
    const int lt = threadIdx.x & 127;
    const bool lane_ok = lt < TX * TY;                    // tiles smaller than 128 columns leave the last lanes idle
    const int lc = lane_ok ? lt : 0;
    const int tx = lc % TX, ty = lc / TX;
    const int i0 = (blockIdx.x / groups) * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int r0 = (blockIdx.x % groups) * KB;
    const int q0 = wg * KH;
    const int z0 = zb + blockIdx.z * zchunk;
    const int z1 = min(ze, z0 + zchunk);
    (void)n2;
    const int niter = z1 - z0 + 2;  // input planes z0-1 .. z1
    const bool active = lane_ok && (i < n0) && (j < n1);
    auto issue = [&](int s, int zi) {
        unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&bars[s], Cfg::TX_BYTES);
        tma_load_4d(st + Cfg::OFF_X, &tm_x, 2 * (i0 - Cfg::HX), j0 - 1, zi, r0, &bars[s]);
        tma_load_4d(st + Cfg::OFF_C, &tm_c, 2 * i0, j0, zi + 1, 0, &bars[s]);                  // dk = -1 of plane zi+1
        tma_load_4d(st + Cfg::OFF_C + Cfg::CSET, &tm_c, 2 * i0, j0, zi, 9, &bars[s]);          // dk =  0 of plane zi
        tma_load_4d(st + Cfg::OFF_C + 2 * Cfg::CSET, &tm_c, 2 * i0, j0, zi - 1, 18, &bars[s]); // dk = +1 of plane zi-1
        if (MODE != MODE_APPLY) tma_load_4d(st + Cfg::OFF_B, &tm_b, 2 * i0, j0, zi - 1, r0, &bars[s]);
        if (MODE == MODE_JACOBI) tma_load_3d(st + Cfg::OFF_D, &tm_d, 2 * i0, j0, zi - 1, &bars[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NS && s < niter; ++s) issue(s, z0 - 1 + s);
    }
    __syncthreads();
    const cx<T> zero = mk<T>(T(0), T(0));
    cx<T> acc[3][KH], xprev[KH];
#pragma unroll
    for (int q = 0; q < KH; ++q) {
        acc[0][q] = acc[1][q] = acc[2][q] = zero;
        xprev[q] = zero;
    }
    const int cidx = (ty + 1) * PX + (tx + Cfg::HX);
    const int bidx = ty * TX + tx;
    const int64_t pxy = (active ? i : 0) + (int64_t)sy * (active ? j : 0);
    const int64_t sz = (int64_t)sy * n1;
#pragma unroll 1
    for (int it = 0; it < niter; ++it) {
        const int zi = z0 - 1 + it;
        const int s = it % NS;
        const unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_wait(&bars[s], (uint32_t)((it / NS) & 1));
        const cx<T>* sx = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_X) + q0 * Cfg::XT;
        const cx<T>* cm = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_C);
        const cx<T>* c0 = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_C + Cfg::CSET);
        const cx<T>* cp = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_C + 2 * Cfg::CSET);
        // the first / last input plane of a chunk feeds one output plane only: skip the other two coefficient sets
        if (it == 0) coarse_scatter<T, KH, Cfg::XT, Cfg::BT, PX, true, false, false>(acc, sx + cidx, cm + bidx, c0 + bidx, cp + bidx);
        else if (it == niter - 1) coarse_scatter<T, KH, Cfg::XT, Cfg::BT, PX, false, false, true>(acc, sx + cidx, cm + bidx, c0 + bidx, cp + bidx);
        else coarse_scatter<T, KH, Cfg::XT, Cfg::BT, PX, true, true, true>(acc, sx + cidx, cm + bidx, c0 + bidx, cp + bidx);
        // output plane zi-1 is complete
        const int zo = zi - 1;
        if (active && zo >= z0 && zo < z1) {
            const int64_t p = pxy + (int64_t)zo * sz;
            const cx<T>* sb = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_B) + q0 * Cfg::BT;
            cx<T> dv = zero;
            if (MODE == MODE_JACOBI) dv = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_D)[bidx];
#pragma unroll
            for (int q = 0; q < KH; ++q) {
                if (r0 + q0 + q < nrhs) {
                    const int64_t o = (int64_t)(r0 + q0 + q) * ld + p;
                    if (MODE == MODE_APPLY) {
                        out[o] = acc[0][q];
                    } else {
                        const cx<T> bv = sb[q * Cfg::BT + bidx];
                        if (MODE == MODE_RESID) out[o] = bv - acc[0][q];
                        else out[o] = xprev[q] + dv * (bv - acc[0][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KH; ++q) {
            acc[0][q] = acc[1][q];
            acc[1][q] = acc[2][q];
            acc[2][q] = zero;
            xprev[q] = sx[q * Cfg::XT + cidx];
        }
        __syncthreads();
        if (threadIdx.x == 0 && it + NS < niter) issue(s, zi + NS);
    }
}

// Column-scaled copy of a stored stencil: As[s][p] = A[s][p] * dinv[p + off(s)], i.e. As = A * diag(dinv).  With it the
// right-preconditioned Jacobi-GMRES of a level (inexact coarsest solve, Jac-GMRES smoother) applies w = (A D^-1) v in
// ONE stencil pass instead of z = D^-1 v followed by w = A z (saves a 2S pass and a launch per GMRES step).
template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_scale_columns(CoarseOp<T> op, cx<T>* __restrict__ scaled) {
    const int NS = (DIM == 3) ? 27 : 9;
    const int64_t nodes = (int64_t)op.n[0] * op.n[1] * op.n[2];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nodes * NS) return;
    const int s = (int)(t / nodes);
    const int64_t q = t - (int64_t)s * nodes;
    const int i = (int)(q % op.n[0]), j = (int)((q / op.n[0]) % op.n[1]), k = (int)(q / ((int64_t)op.n[0] * op.n[1]));
    const int di = s % 3 - 1, dj = (s / 3) % 3 - 1, dk = (DIM == 3) ? (s / 9 - 1) : 0;
    const int64_t p = i + (int64_t)op.sy * j + (int64_t)op.sy * op.n[1] * k;
    cx<T> v = mk<T>(T(0), T(0));
    if ((unsigned)(i + di) < (unsigned)op.n[0] && (unsigned)(j + dj) < (unsigned)op.n[1] && (unsigned)(k + dk) < (unsigned)op.n[2])
        v = op.coef[(int64_t)s * op.N + p] * op.dinv[p + di + (int64_t)op.sy * dj + (int64_t)op.sy * op.n[1] * dk];
    scaled[(int64_t)s * op.N + p] = v;
}

// ---------------------------------------------------------------------------------------------
// 2-D grids (BASELINE configs 1-2): the marching dimension of the TMA ring is the RIGHT-HAND SIDE.  A CTA owns a
// 32 x 8 tile of nodes and a chunk of the RHS block; every thread evaluates the coefficients of its node once --
// the matrix-free 5-point row (centre with mass / absorbing layer / Sommerfeld / shift, four weights) on the fine
// level, the 9 stored Galerkin coefficients on a coarse level -- and keeps them in registers while the x tiles
// (one-node halo) and b tiles of KB right-hand sides at a time stream through an NS-deep mbarrier ring.  Halos at
// the domain boundary and surplus RHS slots are zero-filled by the TMA unit.  With 64 sources per batch (config 2)
// the coefficient work is amortised 64 times and a sweep moves exactly 3S per node and right-hand side.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int KB>
struct Rhs2dCfg {
    static constexpr int TX = 32, TY = 8;
    static constexpr int HX = sizeof(T) == 4 ? 2 : 1;  // see FineTmaCfg
    static constexpr int PX = TX + 2 * HX;
    static constexpr int XT = (TY + 2) * PX;
    static constexpr int BT = TY * TX;
    static constexpr int ES = (int)sizeof(cx<T>);
    static constexpr int al(int b) { return (b + 127) / 128 * 128; }
    static constexpr int OFF_X = 0;
    static constexpr int OFF_B = al(KB * XT * ES);
    static constexpr int STAGE_BYTES = OFF_B + (MODE != MODE_APPLY ? al(KB * BT * ES) : 0);
    static constexpr uint32_t TX_BYTES = KB * XT * ES + (MODE != MODE_APPLY ? KB * BT * ES : 0);
    static constexpr int NS = 4;
};

// COARSE = false: matrix-free fine operator `fop`; COARSE = true: stored 9-point stencil `cop` (+ cop.dinv)
template <typename T, int MODE, int KB, bool COARSE>
__global__ void __launch_bounds__(256) k_stencil2d_tma(FineOp<T> fop, CoarseOp<T> cop, const __grid_constant__ TmaDesc tm_x,
                                                       const __grid_constant__ TmaDesc tm_b, cx<T>* __restrict__ out,
                                                       int64_t ld, int nrhs, int rchunk, T damp) {
    typedef Rhs2dCfg<T, MODE, KB> Cfg;
    constexpr int TX = Cfg::TX, TY = Cfg::TY, PX = Cfg::PX, NS = Cfg::NS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * Cfg::STAGE_BYTES);
    const int n0 = COARSE ? cop.n[0] : fop.n[0], n1 = COARSE ? cop.n[1] : fop.n[1];
    const int sy = COARSE ? cop.sy : fop.sy;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int rb = blockIdx.z * rchunk, re = min(nrhs, rb + rchunk);
    const int niter = (re - rb + KB - 1) / KB;
    const bool active = (i < n0) && (j < n1);
    auto issue = [&](int s, int it) {
        unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&bars[s], Cfg::TX_BYTES);
        tma_load_4d(st + Cfg::OFF_X, &tm_x, 2 * (i0 - Cfg::HX), j0 - 1, 0, rb + it * KB, &bars[s]);
        if (MODE != MODE_APPLY) tma_load_4d(st + Cfg::OFF_B, &tm_b, 2 * i0, j0, 0, rb + it * KB, &bars[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NS && s < niter; ++s) issue(s, s);
    }
    __syncthreads();
    const int ic = active ? i : 0, jc = active ? j : 0;
    const int64_t p = ic + (int64_t)sy * jc;
    // the row of this node, in registers for every right-hand side of the chunk
    cx<T> cf[9];
    cx<T> dinv = mk<T>(T(0), T(0));
    if (COARSE) {
#pragma unroll
        for (int s9 = 0; s9 < 9; ++s9) cf[s9] = cop.coef[(int64_t)s9 * cop.N + p];  // absent neighbours hold 0
        if (MODE == MODE_JACOBI) dinv = cop.dinv[p];
    } else {
        const cx<T> c = fine_center<T, 2>(fop, p, ic, jc, 0);
#pragma unroll
        for (int s9 = 0; s9 < 9; ++s9) cf[s9] = mk<T>(T(0), T(0));
        cf[4] = c;
        cf[3] = mk<T>(-fine_w(fop, 0, 0, ic, n0), T(0));
        cf[5] = mk<T>(-fine_w(fop, 0, 1, ic, n0), T(0));
        cf[1] = mk<T>(-fine_w(fop, 1, 0, jc, n1), T(0));
        cf[7] = mk<T>(-fine_w(fop, 1, 1, jc, n1), T(0));
        if (MODE == MODE_JACOBI) dinv = rdiv(damp, c);
    }
    const int cidx = (ty + 1) * PX + (tx + Cfg::HX);
    const int bidx = ty * TX + tx;
#pragma unroll 1
    for (int it = 0; it < niter; ++it) {
        const int s = it % NS;
        const unsigned char* st = smem_raw + (size_t)s * Cfg::STAGE_BYTES;
        mbar_wait(&bars[s], (uint32_t)((it / NS) & 1));
        if (active) {
            const cx<T>* sx = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_X) + cidx;
            const cx<T>* sb = reinterpret_cast<const cx<T>*>(st + Cfg::OFF_B) + bidx;
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                const int r = rb + it * KB + q;
                const cx<T>* xt = sx + q * Cfg::XT;
                const cx<T> xc = xt[0];
                cx<T> a = cf[4] * xc;
                if (COARSE) {
                    cfma(a, cf[0], xt[-PX - 1]);
                    cfma(a, cf[1], xt[-PX]);
                    cfma(a, cf[2], xt[-PX + 1]);
                    cfma(a, cf[3], xt[-1]);
                    cfma(a, cf[5], xt[1]);
                    cfma(a, cf[6], xt[PX - 1]);
                    cfma(a, cf[7], xt[PX]);
                    cfma(a, cf[8], xt[PX + 1]);
                } else {
                    rfma(a, cf[3].x, xt[-1]);
                    rfma(a, cf[5].x, xt[1]);
                    rfma(a, cf[1].x, xt[-PX]);
                    rfma(a, cf[7].x, xt[PX]);
                }
                if (r < re) {
                    const int64_t o = (int64_t)r * ld + p;
                    if (MODE == MODE_APPLY) {
                        out[o] = a;
                    } else {
                        const cx<T> bv = sb[q * Cfg::BT];
                        if (MODE == MODE_RESID) out[o] = bv - a;
                        else out[o] = xc + dinv * (bv - a);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && it + NS < niter) issue(s, it + NS);
    }
}

// ---------------------------------------------------------------------------------------------
// Device-side set-up: gamma <- gamma0 + getABL(n, NeumannAtFirstDim, ABLpad, ABLamp) evaluated from
// the separable 1-D profiles of src/GetHelmholtz.jl:141-218 (tables built on the host in Float64: they are O(n)).
//   3-D (:164-218): gamma = min(amp (g0[i] + g1[j] + g2[k]), amp), g_d already normalised by (max + 1e-5)
//   2-D (:141-163): gamma = amp ( [top[j] (1 - l[i] - r[i])] + bot[j] + l[i] + r[i] - l[i] bot[j] - r[i] bot[j] )
//                   with tab0 = l, tab1 = r (dim 1 ramps), tab2 = top (zero when Neumann), tab3 = bot
// koff: global index of local plane 0 (slab decomposition).
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_gamma_abl(T* __restrict__ g, int n0, int n1, int n2, int koff, double gamma0, double amp,
                                                   const double* __restrict__ tab0, const double* __restrict__ tab1,
                                                   const double* __restrict__ tab2, const double* __restrict__ tab3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = (DIM == 3) ? blockIdx.z * blockDim.z + threadIdx.z : 0;
    if (i >= n0 || j >= n1 || k >= n2) return;
    double v;
    if (DIM == 3) {
        v = (tab0[i] + tab1[j] + tab2[k + koff]) * amp;
        if (v >= amp) v = amp;
    } else {
        const double l = tab0[i], r = tab1[i], top = tab2[j], bot = tab3[j];
        v = (top - l * top - r * top + bot + l + r - l * bot - r * bot) * amp;
    }
    g[i + (int64_t)n0 * (j + (int64_t)n1 * k)] = (T)(gamma0 + v);
}

// max of a real array (getMaximalFrequency, src/GetHelmholtz.jl:75-79): per-block maxima, finished on the host
template <typename T>
__global__ void __launch_bounds__(256) k_max_partial(const T* __restrict__ a, int64_t n, double* __restrict__ out) {
    __shared__ double sm[32];
    double v = -1e300;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) v = fmax(v, (double)a[p]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = fmax(v, sm[w]);
        out[blockIdx.x] = v;
    }
}

// centre coefficient (incl. Laplacian diagonal) and damp/centre as arrays, for the TMA-staged kernels
template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_fine_precompute(FineOp<T> op, cx<T>* __restrict__ cdiag, cx<T>* __restrict__ dinv,
                                                         T damp) {
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = (DIM == 3) ? blockIdx.z * blockDim.z + threadIdx.z : 0;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const int64_t p = i + (int64_t)op.sy * j + (int64_t)op.sy * n1 * k;
    const cx<T> c = fine_center<T, DIM>(op, p, i, j, k);
    cdiag[p] = c;
    if (dinv) dinv[p] = rdiv(damp, c);
}

// first Jacobi sweep from a zero guess on the fine level: out = damp/diag * b
template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_fine_jacobi0(FineOp<T> op, const cx<T>* __restrict__ b,
                                                      cx<T>* __restrict__ out, int64_t ld, int nrhs, T damp) {
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = (DIM == 3) ? blockIdx.z * blockDim.z + threadIdx.z : 0;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const int64_t p = i + (int64_t)op.sy * j + (int64_t)op.sy * n1 * k;
    const cx<T> dinv = rdiv(damp, fine_center<T, DIM>(op, p, i, j, k));
    for (int r = (DIM == 2 ? blockIdx.z : 0); r < nrhs; r += (DIM == 2 ? gridDim.z : 1)) out[(int64_t)r * ld + p] = dinv * b[(int64_t)r * ld + p];
}

// complex diagonal c_p (without the Laplacian part) in double, for hh_get_diagonal
template <typename T, int DIM>
__global__ void k_fine_diag(FineOp<T> op, zc* __restrict__ out) {
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = (DIM == 3) ? blockIdx.z * blockDim.z + threadIdx.z : 0;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const int64_t p = i + (int64_t)op.sy * j + (int64_t)op.sy * n1 * k;
    cx<T> c = fine_center<T, DIM>(op, p, i, j, k);
    const bool bi = (i == 0) | (i == n0 - 1), bj = (j == 0) | (j == n1 - 1);
    T lap = (bi ? op.BC : T(2)) * op.ih2[0] + (bj ? op.BC : T(2)) * op.ih2[1];
    if (DIM == 3) lap += ((k == 0 || k == n2 - 1) ? op.BC : T(2)) * op.ih2[2];
    out[p] = mk<double>((double)(c.x - lap), (double)c.y);
}

// ---------------------------------------------------------------------------------------------
// coarse levels (baseline form): stored 3^DIM-point stencil, one thread per node, KB RHS per pass
// so that every coefficient is loaded once per KB right-hand sides.
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM, int MODE, int KB>
__global__ void __launch_bounds__(256) k_coarse_stencil(CoarseOp<T> op, const cx<T>* __restrict__ x,
                                                        const cx<T>* __restrict__ b, cx<T>* __restrict__ out,
                                                        int64_t ld, int nrhs) {
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = (DIM == 3) ? op.zb + blockIdx.z * blockDim.z + threadIdx.z : 0;
    if (i >= n0 || j >= n1 || (DIM == 3 && k >= op.ze)) return;
    const int64_t sy = op.sy, sz = (int64_t)op.sy * n1;
    const int64_t N = op.N;
    const int64_t p = i + sy * j + sz * k;
    for (int r0 = 0; r0 < nrhs; r0 += KB) {
        cx<T> acc[KB];
#pragma unroll
        for (int q = 0; q < KB; ++q) acc[q] = mk<T>(T(0), T(0));
#pragma unroll
        for (int dk = (DIM == 3 ? -1 : 0); dk <= (DIM == 3 ? 1 : 0); ++dk) {
            const bool okk = (unsigned)(k + dk) < (unsigned)n2;
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                const bool okj = (unsigned)(j + dj) < (unsigned)n1;
#pragma unroll
                for (int di = -1; di <= 1; ++di) {
                    const bool ok = okk && okj && ((unsigned)(i + di) < (unsigned)n0);
                    if (ok) {
                        const int s = (di + 1) + 3 * (dj + 1) + (DIM == 3 ? 9 * (dk + 1) : 0);
                        const cx<T> cf = op.coef[(int64_t)s * N + p];
                        const int64_t off = p + di + sy * dj + sz * dk;
#pragma unroll
                        for (int q = 0; q < KB; ++q)
                            if (r0 + q < nrhs) cfma(acc[q], cf, x[(int64_t)(r0 + q) * ld + off]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            if (r0 + q < nrhs) {
                const int64_t o = (int64_t)(r0 + q) * ld + p;
                if (MODE == MODE_APPLY) {
                    out[o] = acc[q];
                } else if (MODE == MODE_RESID) {
                    out[o] = b[o] - acc[q];
                } else {
                    out[o] = x[o] + op.dinv[p] * (b[o] - acc[q]);
                }
            }
        }
    }
}

// out = dinv .* b   (first Jacobi sweep from zero on a coarse level; also the Jacobi right
// preconditioner of the inexact coarsest GMRES and the Jac-GMRES smoother)
template <typename T>
__global__ void __launch_bounds__(256) k_diag_scale(const cx<T>* __restrict__ dinv, const cx<T>* __restrict__ b,
                                                    cx<T>* __restrict__ out, int64_t N, int64_t ld, int nrhs) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const cx<T> d = dinv[p];
    for (int r = blockIdx.y; r < nrhs; r += gridDim.y) out[(int64_t)r * ld + p] = d * b[(int64_t)r * ld + p];
}

// fine-level damp/diag as an explicit array (used by Jac-GMRES on the fine level)
template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_fine_dinv(FineOp<T> op, cx<T>* __restrict__ dinv, T damp) {
    const int n0 = op.n[0], n1 = op.n[1], n2 = op.n[2];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = (DIM == 3) ? blockIdx.z * blockDim.z + threadIdx.z : 0;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const int64_t p = i + (int64_t)op.sy * j + (int64_t)op.sy * n1 * k;
    dinv[p] = rdiv(damp, fine_center<T, DIM>(op, p, i, j, k));
}

// ---------------------------------------------------------------------------------------------
// K3 (baseline form): full-weighting restriction  bc = R r,  R = 2^-DIM P^T  (1-D weights 1/4 1/2 1/4,
// truncated at the boundary).  One thread per coarse node.
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_restrict(const cx<T>* __restrict__ r, cx<T>* __restrict__ bc, int nf0, int nf1,
                                                  int nf2, int nc0, int nc1, int nc2, int fsy, int csy, int64_t ldf,
                                                  int64_t ldc, int nrhs, int Kb, int Ke, int fkoff, int fn2g) {
    // slab decomposition: coarse planes Kb <= K < Ke are computed; fine plane fk (local) exists iff its global index
    // fk + fkoff lies in [0, fn2g).  Whole grid: Kb = 0, Ke = nc2, fkoff = 0, fn2g = nf2.
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    const int J = blockIdx.y * blockDim.y + threadIdx.y;
    const int K = (DIM == 3) ? Kb + blockIdx.z * blockDim.z + threadIdx.z : 0;
    if (I >= nc0 || J >= nc1 || (DIM == 3 && K >= Ke)) return;
    (void)nc2;
    (void)nf2;
    const int64_t sy = fsy, sz = (int64_t)fsy * nf1;
    const int64_t pc = I + (int64_t)csy * J + (int64_t)csy * nc1 * K;
    const int fi = 2 * I, fj = 2 * J, fk = (DIM == 3) ? 2 * K : 0;
    // 2-D: the grid's z dimension strides over the right-hand sides (a 2-D grid alone cannot fill the SMs)
    for (int q = (DIM == 2 ? blockIdx.z : 0); q < nrhs; q += (DIM == 2 ? gridDim.z : 1)) {
        const cx<T>* rr = r + (int64_t)q * ldf;
        cx<T> acc = mk<T>(T(0), T(0));
#pragma unroll
        for (int dk = (DIM == 3 ? -1 : 0); dk <= (DIM == 3 ? 1 : 0); ++dk) {
            if (DIM == 3 && (unsigned)(fk + dk + fkoff) >= (unsigned)fn2g) continue;
            const T wk = (DIM == 3) ? (dk == 0 ? T(0.5) : T(0.25)) : T(1);
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                if ((unsigned)(fj + dj) >= (unsigned)nf1) continue;
                const T wj = wk * (dj == 0 ? T(0.5) : T(0.25));
#pragma unroll
                for (int di = -1; di <= 1; ++di) {
                    if ((unsigned)(fi + di) >= (unsigned)nf0) continue;
                    const T w = wj * (di == 0 ? T(0.5) : T(0.25));
                    rfma(acc, w, rr[(fi + di) + sy * (fj + dj) + sz * (fk + dk)]);
                }
            }
        }
        bc[(int64_t)q * ldc + pc] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// K3 (3-D production form): full-weighting restriction with the fine planes staged by TMA.
// A CTA owns a CTX x CTY tile of coarse columns and marches over a chunk of coarse planes; every fine plane it
// touches (2K-1, 2K, 2K+1) is loaded once, as one box of (2 CTX + 1) x (2 CTY + 1) nodes per right-hand side, into an
// NS-deep mbarrier ring; a thread forms the in-plane restriction s(z) of its column from the staged tile and rolls
// bc(K) = 1/4 s(2K-1) + 1/2 s(2K) + 1/4 s(2K+1) in registers.  Rows, columns and planes outside the grid are
// zero-filled by the hardware, which IS the truncated boundary stencil of R = 2^-3 P^T.  Replaces one thread per
// coarse node doing 27 stride-2 global loads (3.4 TB/s, latency-bound) by coalesced bulk loads.
// ---------------------------------------------------------------------------------------------
template <typename T, int KB>
struct RestrictCfg {
    static constexpr int CTX = 16, CTY = 8;
    static constexpr int HX = sizeof(T) == 4 ? 2 : 1;          // the box starts HX fine nodes left of 2*I0 (16-byte aligned start)
    static constexpr int FX = 2 * CTX + 1 + (HX - 1);          // 33 (ComplexF64) / 34 (ComplexF32: even width)
    static constexpr int FY = 2 * CTY + 1;
    static constexpr int FT = FX * FY;
    static constexpr int ES = (int)sizeof(cx<T>);
    static constexpr int al(int b) { return (b + 127) / 128 * 128; }
    static constexpr int STAGE_BYTES = al(KB * FT * ES);
    static constexpr uint32_t TX_BYTES = KB * FT * ES;
    static constexpr int NS = 6;
    static_assert((FX * ES) % 16 == 0, "TMA box rows are 16-byte multiples");
};

template <typename T, int KB>
__global__ void __launch_bounds__(128) k_restrict3d_tma(const __grid_constant__ TmaDesc tm_r, cx<T>* __restrict__ bc,
                                                        int nc0, int nc1, int csy, int64_t ldc, int nrhs, int kchunk,
                                                        int groups, int Kb, int Ke) {
    typedef RestrictCfg<T, KB> Cfg;
    constexpr int NS = Cfg::NS, FX = Cfg::FX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * Cfg::STAGE_BYTES);
    const int tx = threadIdx.x % Cfg::CTX, ty = threadIdx.x / Cfg::CTX;
    const int I0 = (blockIdx.x / groups) * Cfg::CTX, J0 = blockIdx.y * Cfg::CTY;
    const int r0 = (blockIdx.x % groups) * KB;
    const int K0 = Kb + blockIdx.z * kchunk, K1 = min(Ke, K0 + kchunk);
    const int zf0 = 2 * K0 - 1;                 // first fine plane of the chunk (may be -1: zero-filled)
    const int niter = 2 * (K1 - K0) + 1;        // fine planes 2K0-1 .. 2K1-1
    const int I = I0 + tx, J = J0 + ty;
    const bool active = I < nc0 && J < nc1;
    auto issue = [&](int s, int z) {
        mbar_expect_tx(&bars[s], Cfg::TX_BYTES);
        tma_load_4d(smem_raw + (size_t)s * Cfg::STAGE_BYTES, &tm_r, 2 * (2 * I0 - Cfg::HX), 2 * J0 - 1, z, r0, &bars[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NS && s < niter; ++s) issue(s, zf0 + s);
    }
    __syncthreads();
    const cx<T> zero = mk<T>(T(0), T(0));
    cx<T> cur[KB], nxt[KB];  // bc(K) being accumulated, and the quarter of the odd plane that belongs to bc(K+1)
#pragma unroll
    for (int q = 0; q < KB; ++q) cur[q] = nxt[q] = zero;
    const int cidx = (2 * ty + 1) * FX + (2 * tx + Cfg::HX);  // the node (2I, 2J) inside the staged tile
    const int64_t pc = (active ? I : 0) + (int64_t)csy * (active ? J : 0);
    const int64_t csz = (int64_t)csy * nc1;
#pragma unroll 1
    for (int it = 0; it < niter; ++it) {
        const int s = it % NS;
        mbar_wait(&bars[s], (uint32_t)((it / NS) & 1));
        const cx<T>* t = reinterpret_cast<const cx<T>*>(smem_raw + (size_t)s * Cfg::STAGE_BYTES) + cidx;
        const bool odd = (it & 1) == 0;  // fine plane zf0 + it is odd when it is even
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            const cx<T>* tq = t + q * Cfg::FT;
            cx<T> a = tq[0];                                                     // 1/2 * 1/2
            cx<T> e = (tq[-1] + tq[1]) + (tq[-FX] + tq[FX]);                      // 1/2 * 1/4
            cx<T> c = (tq[-FX - 1] + tq[-FX + 1]) + (tq[FX - 1] + tq[FX + 1]);    // 1/4 * 1/4
            cx<T> sxy = T(0.25) * a + T(0.125) * e + T(0.0625) * c;
            if (odd) {
                cur[q] = cur[q] + T(0.25) * sxy;   // closes bc(K) of the plane below ... (K = (z-1)/2)
                nxt[q] = T(0.25) * sxy;            // ... and opens bc(K+1)
            } else {
                cur[q] = cur[q] + T(0.5) * sxy;
            }
        }
        if (odd && it > 0) {  // plane 2K+1 done: bc(K) complete, K = K0 + it/2 - 1
            const int K = K0 + (it >> 1) - 1;
            if (active) {
#pragma unroll
                for (int q = 0; q < KB; ++q)
                    if (r0 + q < nrhs) bc[(int64_t)(r0 + q) * ldc + pc + (int64_t)K * csz] = cur[q];
            }
        }
        if (odd) {
#pragma unroll
            for (int q = 0; q < KB; ++q) cur[q] = nxt[q];
        }
        __syncthreads();
        if (threadIdx.x == 0 && it + NS < niter) issue(s, zf0 + it + NS);
    }
}

// ---------------------------------------------------------------------------------------------
// K4: prolongation + correction  x += P xc  (bi/tri-linear).  One thread per fine node.
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM>
__global__ void __launch_bounds__(256) k_prolong_add(cx<T>* __restrict__ x, const cx<T>* __restrict__ xc, int nf0,
                                                     int nf1, int nf2, int fsy, int csy, int nc1, int64_t ldf,
                                                     int64_t ldc, int nrhs, int zb, int ze) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = (DIM == 3) ? zb + blockIdx.z * blockDim.z + threadIdx.z : 0;  // fine planes zb <= k < ze
    if (i >= nf0 || j >= nf1 || (DIM == 3 && k >= ze)) return;
    (void)nf2;
    const int64_t p = i + (int64_t)fsy * j + (int64_t)fsy * nf1 * k;
    const int64_t cy = csy, cz = (int64_t)csy * nc1;
    const int I0 = i >> 1, J0 = j >> 1, K0 = k >> 1;
    const int oi = i & 1, oj = j & 1, ok = (DIM == 3) ? (k & 1) : 0;
    const T w = T(1) / T((1 << oi) * (1 << oj) * (1 << ok));
    const int64_t base = I0 + cy * J0 + cz * K0;
    for (int q = (DIM == 2 ? blockIdx.z : 0); q < nrhs; q += (DIM == 2 ? gridDim.z : 1)) {
        const cx<T>* c = xc + (int64_t)q * ldc + base;
        cx<T> acc = mk<T>(T(0), T(0));
#pragma unroll
        for (int dk = 0; dk <= 1; ++dk) {
            if (dk > ok) continue;
#pragma unroll
            for (int dj = 0; dj <= 1; ++dj) {
                if (dj > oj) continue;
#pragma unroll
                for (int di = 0; di <= 1; ++di) {
                    if (di > oi) continue;
                    acc = acc + c[di + cy * dj + cz * dk];
                }
            }
        }
        const int64_t o = (int64_t)q * ldf + p;
        cx<T> v = x[o];
        rfma(v, w, acc);
        x[o] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// K5: Galerkin coarse operator  A_c = R A P  as a 3^DIM-point stencil (set-up, once per model/omega).
// One thread per (coarse node, coarse offset).  A(i, t) is supplied by the functor `Coef`.
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM>
struct FineCoef {
    FineOp<T> op;
    __device__ __forceinline__ bool sparse7() const { return true; }
    __device__ __forceinline__ cx<double> get(int i, int j, int k, int di, int dj, int dk) const {
        const int nz = (di != 0) + (dj != 0) + (dk != 0);
        if (nz == 0) {
            const int64_t p = i + (int64_t)op.sy * j + (int64_t)op.sy * op.n[1] * k;
            cx<T> c = fine_center<T, DIM>(op, p, i, j, k);
            return mk<double>((double)c.x, (double)c.y);
        }
        if (nz > 1) return mk<double>(0.0, 0.0);
        T w;
        if (di != 0) w = fine_w(op, 0, di > 0, i, op.n[0]);
        else if (dj != 0) w = fine_w(op, 1, dj > 0, j, op.n[1]);
        else w = fine_wz(op, dk > 0, k);
        return mk<double>(-(double)w, 0.0);
    }
};

template <typename T, int DIM>
struct StoredCoef {
    CoarseOp<T> op;
    __device__ __forceinline__ bool sparse7() const { return false; }
    __device__ __forceinline__ cx<double> get(int i, int j, int k, int di, int dj, int dk) const {
        const int64_t p = i + (int64_t)op.sy * j + (int64_t)op.sy * op.n[1] * k;
        const int s = (di + 1) + 3 * (dj + 1) + (DIM == 3 ? 9 * (dk + 1) : 0);
        cx<T> c = op.coef[(int64_t)s * op.N + p];
        return mk<double>((double)c.x, (double)c.y);
    }
};

template <typename T, int DIM, typename Coef>
__global__ void __launch_bounds__(128) k_galerkin(Coef A, int nf0, int nf1, int nf2, int nc0, int nc1, int nc2,
                                                  int csy, int64_t cN, cx<T>* __restrict__ coefc, int Kb, int Ke,
                                                  int fkoff, int fn2g, int ckoff, int cn2g) {
    // slab decomposition (3-D): rows of the coarse planes Kb <= K < Ke are computed; a fine / coarse plane exists iff
    // its global index (local + fkoff / ckoff) lies inside the global grid (fn2g / cn2g planes); the local arrays
    // hold the halo planes those rows touch.  Whole grid: Kb = 0, Ke = nc2, offsets 0, fn2g = nf2, cn2g = nc2.
    const int64_t Nc = (int64_t)nc0 * nc1 * (Ke - Kb);
    const int NS = (DIM == 3) ? 27 : 9;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Nc * NS) return;
    (void)nf2;
    (void)nc2;
    const int s = (int)(t / Nc);
    int64_t pc = t - (int64_t)s * Nc;
    const int I = (int)(pc % nc0);
    const int J = (int)((pc / nc0) % nc1);
    const int K = Kb + (int)(pc / ((int64_t)nc0 * nc1));
    const int dI = s % 3 - 1, dJ = (s / 3) % 3 - 1, dK = (DIM == 3) ? (s / 9 - 1) : 0;
    const int JI = I + dI, JJ = J + dJ, JK = K + dK;  // coarse column node
    cx<double> acc = mk<double>(0.0, 0.0);
    if ((unsigned)JI < (unsigned)nc0 && (unsigned)JJ < (unsigned)nc1 && (unsigned)(JK + ckoff) < (unsigned)cn2g) {
        const double rscale = (DIM == 3) ? 0.125 : 0.25;
        for (int ek = (DIM == 3 ? -1 : 0); ek <= (DIM == 3 ? 1 : 0); ++ek) {
            const int fk = 2 * K + ek;
            if ((unsigned)(fk + fkoff) >= (unsigned)fn2g) continue;
            for (int ej = -1; ej <= 1; ++ej) {
                const int fj = 2 * J + ej;
                if ((unsigned)fj >= (unsigned)nf1) continue;
                for (int ei = -1; ei <= 1; ++ei) {
                    const int fi = 2 * I + ei;
                    if ((unsigned)fi >= (unsigned)nf0) continue;
                    const double wR = rscale * (ei ? 0.5 : 1.0) * (ej ? 0.5 : 1.0) * (ek ? 0.5 : 1.0);
                    // columns j = i + t of A that interpolate from coarse node (JI,JJ,JK)
                    for (int tk = (DIM == 3 ? -1 : 0); tk <= (DIM == 3 ? 1 : 0); ++tk) {
                        const int gk = fk + tk;
                        if ((unsigned)(gk + fkoff) >= (unsigned)fn2g) continue;
                        const int qk = (DIM == 3) ? gk - 2 * JK : 0;
                        if (qk < -1 || qk > 1) continue;
                        for (int tj = -1; tj <= 1; ++tj) {
                            const int gj = fj + tj;
                            if ((unsigned)gj >= (unsigned)nf1) continue;
                            const int qj = gj - 2 * JJ;
                            if (qj < -1 || qj > 1) continue;
                            for (int ti = -1; ti <= 1; ++ti) {
                                const int gi = fi + ti;
                                if ((unsigned)gi >= (unsigned)nf0) continue;
                                const int qi = gi - 2 * JI;
                                if (qi < -1 || qi > 1) continue;
                                if (A.sparse7() && ((ti != 0) + (tj != 0) + (tk != 0)) > 1) continue;
                                const double wP = (qi ? 0.5 : 1.0) * (qj ? 0.5 : 1.0) * (qk ? 0.5 : 1.0);
                                const cx<double> a = A.get(fi, fj, fk, ti, tj, tk);
                                const double w = wR * wP;
                                acc.x += w * a.x;
                                acc.y += w * a.y;
                            }
                        }
                    }
                }
            }
        }
    }
    coefc[(int64_t)s * cN + I + (int64_t)csy * J + (int64_t)csy * nc1 * K] = mk<T>((T)acc.x, (T)acc.y);
}

template <typename T>
__global__ void k_coarse_dinv(const cx<T>* __restrict__ center, cx<T>* __restrict__ dinv, int64_t N, T damp) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) {
        const cx<T> c = center[p];
        dinv[p] = (c.x == T(0) && c.y == T(0)) ? mk<T>(T(0), T(0)) : rdiv(damp, c);  // ghost nodes of a padded level stay 0
    }
}

// ---------------------------------------------------------------------------------------------
// coarsest level, exact: banded LU without pivoting (the shifted operator has its numerical range in
// a half plane, so every leading minor is non-singular), then an explicit inverse so that the
// per-cycle coarsest solve is one dense, HBM-streaming matrix product instead of a serial
// triangular solve.
// band[row*W + (col-row+bw)], W = 2 bw + 1, in double.
// ---------------------------------------------------------------------------------------------
template <typename T, int DIM>
__global__ void k_band_fill(CoarseOp<T> op, zc* __restrict__ band, int bw) {
    const int64_t N = (int64_t)op.n[0] * op.n[1] * op.n[2];  // logical unknowns of the dense band matrix
    const int NS = (DIM == 3) ? 27 : 9;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * NS) return;
    const int s = (int)(t / N);
    const int64_t p = t - (int64_t)s * N;
    const int i = (int)(p % op.n[0]), j = (int)((p / op.n[0]) % op.n[1]), k = (int)(p / ((int64_t)op.n[0] * op.n[1]));
    const int di = s % 3 - 1, dj = (s / 3) % 3 - 1, dk = (DIM == 3) ? (s / 9 - 1) : 0;
    if ((unsigned)(i + di) >= (unsigned)op.n[0] || (unsigned)(j + dj) >= (unsigned)op.n[1] ||
        (unsigned)(k + dk) >= (unsigned)op.n[2])
        return;
    const int64_t off = di + (int64_t)op.n[0] * dj + (int64_t)op.n[0] * op.n[1] * dk;
    const cx<T> c = op.coef[(int64_t)s * op.N + i + (int64_t)op.sy * j + (int64_t)op.sy * op.n[1] * k];
    band[p * (2 * (int64_t)bw + 1) + (off + bw)] = mk<double>((double)c.x, (double)c.y);
}

// right-looking banded LU, one CTA (set-up only).  L (unit lower) and U overwrite the band.
// The factorisation does not pivot (see above); with shift == 0 and no attenuation the coarse operator is indefinite
// and a pivot can collapse.  flag[0] is set to 1 + the row of the first pivot that is not finite or has lost ten digits
// against the diagonal entry it started from (diag0, scratch of N doubles); the host turns that into
// HH_ERR_UNSUPPORTED instead of returning Inf/NaN.
__global__ void __launch_bounds__(1024) k_band_lu(zc* __restrict__ band, int64_t N, int bw, int* __restrict__ flag,
                                                  double* __restrict__ diag0) {
    const int64_t W = 2 * (int64_t)bw + 1;
    for (int64_t k = threadIdx.x; k < N; k += blockDim.x) diag0[k] = fabs(band[k * W + bw].x) + fabs(band[k * W + bw].y);
    __syncthreads();
    for (int64_t k = 0; k < N; ++k) {
        const int nrow = (int)min((int64_t)bw, N - 1 - k);
        const zc piv = band[k * W + bw];
        {
            const double pa = fabs(piv.x) + fabs(piv.y);
            if (threadIdx.x == 0 && flag[0] == 0 && !(pa > 1e-10 * diag0[k] && pa < 1e300 && pa > 0.0)) flag[0] = (int)min(k + 1, (int64_t)INT_MAX);
        }
        for (int t = threadIdx.x; t < nrow; t += blockDim.x) {
            const int64_t i = k + 1 + t;
            zc* e = &band[i * W + (k - i + bw)];
            *e = cdiv(*e, piv);
        }
        __syncthreads();
        const int64_t tot = (int64_t)nrow * nrow;
        for (int64_t t = threadIdx.x; t < tot; t += blockDim.x) {
            const int ri = (int)(t / nrow), cj = (int)(t % nrow);
            const int64_t i = k + 1 + ri, j = k + 1 + cj;
            const zc l = band[i * W + (k - i + bw)];
            const zc u = band[k * W + (j - k + bw)];
            zc* e = &band[i * W + (j - i + bw)];
            zc v = *e;
            v.x -= l.x * u.x - l.y * u.y;
            v.y -= l.x * u.y + l.y * u.x;
            *e = v;
        }
        __syncthreads();
    }
}

// Columns [c0, c0+gridDim.x) of the inverse: solve L U x = e_c with the band factors.  One CTA per
// column; x lives in global memory (column c of `inv`, column-major N x N, double).
__global__ void __launch_bounds__(256) k_band_inverse(const zc* __restrict__ band, int64_t N, int bw, int64_t c0,
                                                      zc* __restrict__ inv) {
    const int64_t W = 2 * (int64_t)bw + 1;
    const int64_t c = c0 + blockIdx.x;
    if (c >= N) return;
    zc* x = inv + c * N;
    for (int64_t p = threadIdx.x; p < N; p += blockDim.x) x[p] = mk<double>(p == c ? 1.0 : 0.0, 0.0);
    __syncthreads();
    // forward substitution (unit lower); entries before row c stay zero
    for (int64_t k = c; k < N; ++k) {
        const zc yk = x[k];
        const int nrow = (int)min((int64_t)bw, N - 1 - k);
        if (yk.x != 0.0 || yk.y != 0.0) {
            for (int t = threadIdx.x; t < nrow; t += blockDim.x) {
                const int64_t i = k + 1 + t;
                const zc l = band[i * W + (k - i + bw)];
                zc v = x[i];
                v.x -= l.x * yk.x - l.y * yk.y;
                v.y -= l.x * yk.y + l.y * yk.x;
                x[i] = v;
            }
        }
        __syncthreads();
    }
    // backward substitution
    for (int64_t k = N - 1; k >= 0; --k) {
        if (threadIdx.x == 0) x[k] = cdiv(x[k], band[k * W + bw]);
        __syncthreads();
        const zc xk = x[k];
        const int nrow = (int)min((int64_t)bw, k);
        for (int t = threadIdx.x; t < nrow; t += blockDim.x) {
            const int64_t i = k - 1 - t;
            const zc u = band[i * W + (k - i + bw)];
            zc v = x[i];
            v.x -= u.x * xk.x - u.y * xk.y;
            v.y -= u.x * xk.y + u.y * xk.x;
            x[i] = v;
        }
        __syncthreads();
    }
}

// inverse (column-major, double) -> row-major copy in the solve precision
template <typename T>
__global__ void k_inverse_pack(const zc* __restrict__ inv, cx<T>* __restrict__ invT, int64_t N) {
    __shared__ zc tile[32][33];
    const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t row = bx + threadIdx.x, col = by + r;  // inv[row + col*N]
        if (row < N && col < N) tile[r][threadIdx.x] = inv[row + col * N];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t row = bx + r, col = by + threadIdx.x;  // invT[row*N + col]
        if (row < N && col < N) {
            const zc v = tile[threadIdx.x][r];
            invT[row * N + col] = mk<T>((T)v.x, (T)v.y);
        }
    }
}

// K6 (exact): xc = Ainv * bc.  One warp per output row, KB right-hand sides per pass; the row of
// the inverse is streamed once from HBM and reused for KB right-hand sides.
template <typename T, int KB>
__global__ void __launch_bounds__(256) k_dense_apply(const cx<T>* __restrict__ invT, const cx<T>* __restrict__ b,
                                                     cx<T>* __restrict__ x, int64_t N, int64_t ld, int nrhs) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const cx<T>* a = invT + row * N;
    for (int r0 = 0; r0 < nrhs; r0 += KB) {
        double ax[KB], ay[KB];
#pragma unroll
        for (int q = 0; q < KB; ++q) ax[q] = ay[q] = 0.0;
        for (int64_t c = lane; c < N; c += 32) {
            const cx<T> av = a[c];
#pragma unroll
            for (int q = 0; q < KB; ++q) {
                if (r0 + q < nrhs) {
                    const cx<T> bv = b[(int64_t)(r0 + q) * ld + c];
                    ax[q] += (double)av.x * bv.x - (double)av.y * bv.y;
                    ay[q] += (double)av.x * bv.y + (double)av.y * bv.x;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            const double sx = warp_sum(ax[q]), sy = warp_sum(ay[q]);
            if (lane == 0 && r0 + q < nrhs) x[(int64_t)(r0 + q) * ld + row] = mk<T>((T)sx, (T)sy);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K7: batched Krylov vector kernels.  Vectors are N x nrhs (leading dimension ld); every
// right-hand side has its own scalars (batched, not block, Krylov).
// ---------------------------------------------------------------------------------------------
#define HH_MAXV 10
template <typename T>
struct VecList {
    const cx<T>* v[HH_MAXV];
};

// partial[(i*nrhs + r)*nblk + blk] = sum over the block's slice of conj(V_i) .* w  for i < NV, and, if
// WITH_NORM, entry i = NV holds sum |w|^2.  Grid (nblk, nrhs).  Accumulation in double.
template <typename T, int NV, bool WITH_NORM>
__global__ void __launch_bounds__(256) k_multidot(VecList<T> V, const cx<T>* __restrict__ w, int64_t N, int64_t ld,
                                                  zc* __restrict__ partial) {
    __shared__ double sm[32];
    const int r = blockIdx.y, nrhs = gridDim.y, nblk = gridDim.x;
    const int64_t base = (int64_t)r * ld;
    double ax[NV + 1], ay[NV + 1];
#pragma unroll
    for (int i = 0; i <= NV; ++i) ax[i] = ay[i] = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)nblk * blockDim.x) {
        const cx<T> wv = w[base + p];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const cx<T> vv = V.v[i][base + p];
            ax[i] += (double)vv.x * wv.x + (double)vv.y * wv.y;
            ay[i] += (double)vv.x * wv.y - (double)vv.y * wv.x;
        }
        if (WITH_NORM) ax[NV] += (double)wv.x * wv.x + (double)wv.y * wv.y;
    }
    const int nout = WITH_NORM ? NV + 1 : NV;
#pragma unroll
    for (int i = 0; i < nout; ++i) {
        const double sx = block_sum(ax[i], sm);
        const double sy = (i < NV) ? block_sum(ay[i], sm) : 0.0;
        if (threadIdx.x == 0) partial[((int64_t)i * nrhs + r) * nblk + blockIdx.x] = mk<double>(sx, sy);
    }
}

// w <- (w + sum_i coef[r*cstride + i] * V_i) * post[r]   (i < NV; post == nullptr: no scaling), optionally
// accumulating |w_new|^2 partials into partial[r*nblk + blk].  `negate` subtracts (Gram-Schmidt).  The
// post-scale writes the next Arnoldi vector already (approximately) normalised, so no separate scaling
// pass over the vector is needed (scaled-basis GMRES, see k_gmres_hcol).
template <typename T, int NV, bool WITH_NORM>
__global__ void __launch_bounds__(256) k_multiaxpy(VecList<T> V, cx<T>* __restrict__ w, int64_t N, int64_t ld,
                                                   const zc* __restrict__ coef, int cstride, int negate,
                                                   const zc* __restrict__ post, zc* __restrict__ partial) {
    __shared__ double sm[32];
    const int r = blockIdx.y, nblk = gridDim.x;
    const int64_t base = (int64_t)r * ld;
    cx<T> c[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const zc cc = coef[(int64_t)r * cstride + i];
        c[i] = negate ? mk<T>((T)-cc.x, (T)-cc.y) : mk<T>((T)cc.x, (T)cc.y);
    }
    const T ps = post ? (T)post[r].x : T(1);
    double nrm = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)nblk * blockDim.x) {
        cx<T> wv = w[base + p];
#pragma unroll
        for (int i = 0; i < NV; ++i) cfma(wv, c[i], V.v[i][base + p]);
        wv = ps * wv;
        w[base + p] = wv;
        if (WITH_NORM) nrm += (double)wv.x * wv.x + (double)wv.y * wv.y;
    }
    if (WITH_NORM) {
        const double s = block_sum(nrm, sm);
        if (threadIdx.x == 0) partial[(int64_t)r * nblk + blockIdx.x] = mk<double>(s, 0.0);
    }
}

// x (+)= dinv .* sum_i coef[r*cstride + i] * V_i   (i < NV): the solution update of a Jacobi-preconditioned GMRES of a
// level, x = x0 + D^-1 (V y), in one pass (instead of zero + multiaxpy + diagonal scaling [+ axpy]).  dinv has one
// entry per node (shared by the right-hand sides).  ACCUM: add to x, else overwrite.
template <typename T, int NV, bool ACCUM>
__global__ void __launch_bounds__(256) k_combine(VecList<T> V, const cx<T>* __restrict__ dinv, cx<T>* __restrict__ x,
                                                 int64_t N, int64_t ld, const zc* __restrict__ coef, int cstride) {
    const int r = blockIdx.y, nblk = gridDim.x;
    const int64_t base = (int64_t)r * ld;
    cx<T> c[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const zc cc = coef[(int64_t)r * cstride + i];
        c[i] = mk<T>((T)cc.x, (T)cc.y);
    }
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)nblk * blockDim.x) {
        cx<T> acc = mk<T>(T(0), T(0));
#pragma unroll
        for (int i = 0; i < NV; ++i) cfma(acc, c[i], V.v[i][base + p]);
        acc = dinv[p] * acc;
        if (ACCUM) acc = x[base + p] + acc;
        x[base + p] = acc;
    }
}

// dst (row pitch dsy, leading dimension dld) <- src (row pitch ssy, leading dimension sld); rows of n0 nodes
template <typename U>
__global__ void __launch_bounds__(256) k_repitch(const U* __restrict__ src, U* __restrict__ dst, int n0, int64_t rows,
                                                 int ssy, int dsy, int64_t sld, int64_t dld) {
    const int r = blockIdx.y;
    const int64_t tot = rows * n0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / n0;
        const int i = (int)(t - row * n0);
        dst[(int64_t)r * dld + row * dsy + i] = src[(int64_t)r * sld + row * ssy + i];
    }
}

// precision conversion between blocks with different row pitches (mixed precision: ComplexF64 Krylov vectors <->
// ComplexF32 multigrid vectors); rows of n0 nodes, ghost nodes of a padded destination are left untouched (zero)
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) k_convert(const cx<TS>* __restrict__ src, cx<TD>* __restrict__ dst, int n0,
                                                 int64_t rows, int ssy, int dsy, int64_t sld, int64_t dld) {
    const int r = blockIdx.y;
    const int64_t tot = rows * n0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / n0;
        const int i = (int)(t - row * n0);
        const cx<TS> v = src[(int64_t)r * sld + row * ssy + i];
        dst[(int64_t)r * dld + row * dsy + i] = mk<TD>((TD)v.x, (TD)v.y);
    }
}

// scatter point sources: B[idx[r] + r*ld] = val[r]  (B zeroed beforehand)
template <typename T>
__global__ void k_point_sources(cx<T>* __restrict__ B, int64_t ld, const int64_t* __restrict__ idx,
                                const zc* __restrict__ val, int nrhs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrhs && idx[r] >= 0) B[(int64_t)r * ld + idx[r]] = mk<T>((T)val[r].x, (T)val[r].y);  // < 0: another slab's node
}

// out[q] = sum over the nblk partials of quantity q (slab decomposition: the local sums that are all-reduced over the
// slabs; the scalar kernels then run with nblk = 1).  One warp per quantity.
__global__ void k_sum_partials(const zc* __restrict__ partial, int nq, int nblk, zc* __restrict__ out) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int lane = threadIdx.x & 31;
    double sx = 0.0, sy = 0.0;
    for (int t = lane; t < nblk; t += 32) {
        const zc v = partial[(int64_t)q * nblk + t];
        sx += v.x;
        sy += v.y;
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    if (lane == 0) out[q] = mk<double>(sx, sy);
}

// ---- small per-RHS scalar kernels (one warp per right-hand side; all state in double) ----------

// sum `nblk` partials for quantity (i, r):  partial[(i*nrhs + r)*nblk + .]
__device__ __forceinline__ zc reduce_partials(const zc* partial, int64_t idx, int nblk) {
    const int lane = threadIdx.x & 31;
    double sx = 0.0, sy = 0.0;
    for (int t = lane; t < nblk; t += 32) {
        const zc v = partial[idx * nblk + t];
        sx += v.x;
        sy += v.y;
    }
    return mk<double>(warp_sum(sx), warp_sum(sy));
}

// Scaled-basis GMRES.  The stored Arnoldi vectors v~_i = d_i v_i are only approximately normalised
// (d_0 = ||r||, d_{i+1} ~ 1): the next vector is written as (w - sum c_i v~_i) / beta_est in the same pass
// that orthogonalises it, with beta_est^2 = ||w||^2 - sum |<v~_i,w>|^2/d_i^2 known from the dot pass, and its
// true norm d_{i+1} is measured in that pass.  All relations are exact for any positive beta_est (a poor
// estimate only makes d_{i+1} differ from 1), the Hessenberg matrix is that of the unit basis
// (h_ij = g_i/(d_i d_j), h_{j+1,j} = beta_est d_{j+1}/d_j), and the update uses y_j/d_j on the stored
// preconditioned vectors z~_j = M v~_j.  This removes the separate "v = w/||w||" pass over the vectors.
struct GmresState {
    // per RHS r:  H[r][(m+1)*m] column-major (ldh = m+1), cs[r][m] (real in .x), sn[r][m], s[r][m+1],
    // hcol[r][m+1]: Gram-Schmidt update coefficients g_i/d_i^2 of the current column, y[r][m]: y_j/d_j
    zc *H, *cs, *sn, *s, *hcol, *y;
    double *bnorm, *err;  // per RHS
    zc* scale;            // per RHS: 1/beta_est of the current column (0 when frozen)
    int *done, *jdone, *nprec;
    int m;       // restart length
    double* d;   // [r][m+1] norms of the stored basis vectors
    double* acc; // [r][2]: ||w||^2 and sum |g_i|^2/d_i^2 of the current column
};

// after a multidot group: g_i = <v~_i, w>, i0 <= i < i0+nv (<= j); the first group also carries ||w||^2.
// Called by every lane of the warp that owns right-hand side r (lane 0 writes).
__device__ __forceinline__ void gmres_hcol_dev(const GmresState& st, const zc* __restrict__ partial, int nblk, int r, int nrhs,
                                               int j, int i0, int nv) {
    const int ldh = st.m + 1;
    const bool lead = (threadIdx.x & 31) == 0;
    double sumsq = 0.0;
    for (int i = 0; i < nv; ++i) {
        const zc g = reduce_partials(partial, (int64_t)i * nrhs + r, nblk);
        if (lead) {
            const double di = st.d[(int64_t)r * ldh + i0 + i], dj = st.d[(int64_t)r * ldh + j];
            const bool ok = !st.done[r] && di > 0.0 && dj > 0.0;
            st.hcol[(int64_t)r * ldh + i0 + i] = ok ? (1.0 / (di * di)) * g : mk<double>(0.0, 0.0);
            st.H[(int64_t)r * ldh * st.m + (int64_t)j * ldh + i0 + i] = ok ? (1.0 / (di * dj)) * g : mk<double>(0.0, 0.0);
            if (ok) sumsq += (g.x * g.x + g.y * g.y) / (di * di);
        }
    }
    zc nw = mk<double>(0.0, 0.0);
    if (i0 == 0) nw = reduce_partials(partial, (int64_t)nv * nrhs + r, nblk);
    if (!lead) return;
    double* acc = st.acc + 2 * (int64_t)r;
    if (i0 == 0) {
        acc[0] = nw.x;
        acc[1] = 0.0;
    }
    acc[1] += sumsq;
    if (i0 + nv == j + 1) {  // last group: the estimate of ||w - sum ...||
        double be2 = acc[0] - acc[1];
        const double floor2 = 1e-6 * acc[0];  // any positive value is valid; keeps d_{j+1} within [~1e-3, 1]
        if (!(be2 > floor2)) be2 = floor2;
        st.scale[r] = mk<double>((st.done[r] || !(be2 > 0.0)) ? 0.0 : 1.0 / sqrt(be2), 0.0);
    }
}
__global__ void k_gmres_hcol(GmresState st, const zc* __restrict__ partial, int nblk, int j, int i0, int nv) {
    gmres_hcol_dev(st, partial, nblk, blockIdx.x, gridDim.x, j, i0, nv);
}

// after the orthogonalisation pass: d_{j+1} = ||v~_{j+1}||, h_{j+1,j}, Givens rotations, residual estimate.
// One thread per right-hand side.  nn2 = ||v~_{j+1}||^2 as measured by the update pass.
// est != 0: the update pass was NOT run (the last column of a cycle: v_{j+1} is never used, only h_{j+1,j} is) and
// h_{j+1,j} = ||w - sum <v_i,w> v_i|| / d_j comes from the dot pass alone: ||w||^2 - sum |g_i|^2/d_i^2 (orthonormal v_i).
// The cancellation error of that difference is ~1e-15 ||w||^2, i.e. relative 1e-15 (||w||/h)^2 in h^2; the floor keeps a
// near-breakdown column (h < 1e-6 ||w||, the Krylov space already holds the solution) from reading as h = 0, which would
// report a zero residual: the estimate errs towards "not converged" and the restart's true residual decides.
__device__ __forceinline__ void gmres_givens_dev(const GmresState& st, int r, int j, double tol, double nn2, int est) {
    const int m = st.m, ldh = m + 1;
    if (st.done[r]) return;
    zc* Hc = st.H + (int64_t)r * ldh * m + (int64_t)j * ldh;
    zc* cs = st.cs + (int64_t)r * m;
    zc* sn = st.sn + (int64_t)r * m;
    zc* s = st.s + (int64_t)r * ldh;
    double* d = st.d + (int64_t)r * ldh;
    double hn;
    if (est) {
        const double* acc = st.acc + 2 * (int64_t)r;
        double be2 = acc[0] - acc[1];
        const double floor2 = 1e-12 * acc[0];
        if (!(be2 > floor2)) be2 = floor2;
        d[j + 1] = 0.0;  // no stored vector
        hn = (d[j] > 0.0) ? sqrt(be2) / d[j] : 0.0;
    } else {
        const double dn = sqrt(nn2);
        d[j + 1] = dn;
        const double sc = st.scale[r].x;
        hn = (sc > 0.0 && d[j] > 0.0) ? dn / (sc * d[j]) : 0.0;
    }
    Hc[j + 1] = mk<double>(hn, 0.0);
    for (int k = 0; k < j; ++k) {
        const zc t = cs[k].x * Hc[k] + sn[k] * Hc[k + 1];
        Hc[k + 1] = cs[k].x * Hc[k + 1] - conj(sn[k]) * Hc[k];
        Hc[k] = t;
    }
    const zc a = Hc[j];
    const double aa = sqrt(a.x * a.x + a.y * a.y);
    const double den = sqrt(aa * aa + hn * hn);
    double c;
    zc sgn;
    if (den == 0.0) {
        c = 1.0;
        sgn = mk<double>(0.0, 0.0);
    } else if (aa == 0.0) {
        c = 0.0;
        sgn = mk<double>(1.0, 0.0);
    } else {
        c = aa / den;
        sgn = (hn / (den * aa)) * a;  // (a/|a|) * conj(b)/den with b = hn real
    }
    cs[j] = mk<double>(c, 0.0);
    sn[j] = sgn;
    Hc[j] = c * a + hn * sgn;
    Hc[j + 1] = mk<double>(0.0, 0.0);
    const zc sj = s[j];
    s[j + 1] = mk<double>(0.0, 0.0) - conj(sgn) * sj;
    s[j] = c * sj;
    const zc sn1 = s[j + 1];
    const double err = sqrt(sn1.x * sn1.x + sn1.y * sn1.y) / st.bnorm[r];
    st.err[r] = err;
    st.jdone[r] = j + 1;
    st.nprec[r] += 1;
    if (!(err > tol)) st.done[r] = 1;  // also catches NaN -> stops; host checks for NaN separately
    if (err != err) st.done[r] = 2;
}
__global__ void k_gmres_givens(GmresState st, const zc* __restrict__ partial, int nblk, int j, double tol, int est) {
    const int r = blockIdx.x;
    zc nn = mk<double>(0.0, 0.0);
    if (!est) nn = reduce_partials(partial, r, nblk);
    if ((threadIdx.x & 31) != 0) return;
    gmres_givens_dev(st, r, j, tol, nn.x, est);
}

// One scalar kernel per step of the fixed-length GMRES of a level (smoother / coarsest solve / K-cycle: no
// convergence test, so the Givens update of a column can wait for the next column's dot sums):
//   flags & 1: column j-1's update pass has run; its ||v~_j||^2 partials are in `norm` -> Givens of column j-1
//   then the Gram-Schmidt coefficients of column j from `dots` (one multidot group with ||w||^2, nv = j+1)
//   flags & 2: column j is the last one -> its Givens update from the estimate (no update pass follows)
__global__ void k_gmres_small_step(GmresState st, const zc* __restrict__ dots, int nblk_d, const zc* __restrict__ norm,
                                   int nblk_n, int j, int flags) {
    const int r = blockIdx.x, nrhs = gridDim.x;
    const bool lead = (threadIdx.x & 31) == 0;
    if (flags & 1) {
        const zc nn = reduce_partials(norm, r, nblk_n);
        if (lead) gmres_givens_dev(st, r, j - 1, 0.0, nn.x, 0);
    }
    __syncwarp();
    gmres_hcol_dev(st, dots, nblk_d, r, nrhs, j, 0, j + 1);
    if ((flags & 2) && lead) gmres_givens_dev(st, r, j, 0.0, 0.0, 1);
}

// ---- latency-lean forms of the scalar kernels --------------------------------------------------------------------
// The one-warp kernels above spend their ~15-20 us in chains of dependent global loads (one quantity after the other,
// then one Givens rotation after the other).  Here one warp per QUANTITY sums its block partials (all loads in flight at
// once), and warp 0 then does the per-RHS algebra with lane-parallel loads of the Hessenberg column and the stored
// rotations (lane k holds entry k; the serial recurrences read them by shuffle).  Same arithmetic, same order.
__device__ __forceinline__ zc shfl_zc(zc v, int src) {
    return mk<double>(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
// warp 0, all lanes: Gram-Schmidt coefficients of column j from the sums g[0..nv) (g[nv] = ||w||^2 when i0 == 0)
__device__ __forceinline__ void gmres_hcol_w0(const GmresState& st, const zc* g, int r, int j, int i0, int nv) {
    const int lane = threadIdx.x & 31;
    const int ldh = st.m + 1;
    const int done = st.done[r];
    const double dj = st.d[(int64_t)r * ldh + j];
    double term = 0.0;
    if (lane < nv) {
        const double di = st.d[(int64_t)r * ldh + i0 + lane];
        const zc gi = g[lane];
        const bool ok = !done && di > 0.0 && dj > 0.0;
        st.hcol[(int64_t)r * ldh + i0 + lane] = ok ? (1.0 / (di * di)) * gi : mk<double>(0.0, 0.0);
        st.H[(int64_t)r * ldh * st.m + (int64_t)j * ldh + i0 + lane] = ok ? (1.0 / (di * dj)) * gi : mk<double>(0.0, 0.0);
        if (ok) term = (gi.x * gi.x + gi.y * gi.y) / (di * di);
    }
    double sumsq = 0.0;
    for (int i = 0; i < nv; ++i) sumsq += __shfl_sync(0xffffffffu, term, i);  // the order of the one-thread version
    if (lane != 0) return;
    double* acc = st.acc + 2 * (int64_t)r;
    double a0 = acc[0], a1 = acc[1];
    if (i0 == 0) {
        a0 = g[nv].x;
        a1 = 0.0;
    }
    a1 += sumsq;
    acc[0] = a0;
    acc[1] = a1;
    if (i0 + nv == j + 1) {
        double be2 = a0 - a1;
        const double floor2 = 1e-6 * a0;
        if (!(be2 > floor2)) be2 = floor2;
        st.scale[r] = mk<double>((done || !(be2 > 0.0)) ? 0.0 : 1.0 / sqrt(be2), 0.0);
    }
}
// warp 0, all lanes (uniform control flow): the Givens update of column j (see gmres_givens_dev)
__device__ __forceinline__ void gmres_givens_w0(const GmresState& st, int r, int j, double tol, double nn2, int est) {
    const int lane = threadIdx.x & 31;
    const int m = st.m, ldh = m + 1;
    if (j + 1 > 32) {  // restart lengths beyond a warp: the one-thread form
        if (lane == 0) gmres_givens_dev(st, r, j, tol, nn2, est);
        return;
    }
    if (st.done[r]) return;
    zc* Hc = st.H + (int64_t)r * ldh * m + (int64_t)j * ldh;
    zc* cs = st.cs + (int64_t)r * m;
    zc* sn = st.sn + (int64_t)r * m;
    zc* s = st.s + (int64_t)r * ldh;
    double* d = st.d + (int64_t)r * ldh;
    // every load of the update, issued together
    const zc hk = lane <= j ? Hc[lane] : mk<double>(0.0, 0.0);
    const double ck = lane < j ? cs[lane].x : 0.0;
    const zc sk = lane < j ? sn[lane] : mk<double>(0.0, 0.0);
    const double dj = d[j];
    const zc sj = s[j];
    const double bn = st.bnorm[r];
    const double sc = st.scale[r].x;
    const double a0 = st.acc[2 * (int64_t)r], a1 = st.acc[2 * (int64_t)r + 1];
    const int np0 = st.nprec[r];
    double hn, dn1;
    if (est) {
        double be2 = a0 - a1;
        const double floor2 = 1e-12 * a0;
        if (!(be2 > floor2)) be2 = floor2;
        dn1 = 0.0;  // no stored vector
        hn = (dj > 0.0) ? sqrt(be2) / dj : 0.0;
    } else {
        dn1 = sqrt(nn2);
        hn = (sc > 0.0 && dj > 0.0) ? dn1 / (sc * dj) : 0.0;
    }
    zc cur = shfl_zc(hk, 0);
    for (int k = 0; k < j; ++k) {
        const zc nxt = shfl_zc(hk, k + 1);
        const double c = __shfl_sync(0xffffffffu, ck, k);
        const zc sg = shfl_zc(sk, k);
        const zc t = c * cur + sg * nxt;
        cur = c * nxt - conj(sg) * cur;
        if (lane == 0) Hc[k] = t;
    }
    if (lane != 0) return;
    d[j + 1] = dn1;
    const zc a = cur;
    const double aa = sqrt(a.x * a.x + a.y * a.y);
    const double den = sqrt(aa * aa + hn * hn);
    double c;
    zc sgn;
    if (den == 0.0) {
        c = 1.0;
        sgn = mk<double>(0.0, 0.0);
    } else if (aa == 0.0) {
        c = 0.0;
        sgn = mk<double>(1.0, 0.0);
    } else {
        c = aa / den;
        sgn = (hn / (den * aa)) * a;
    }
    cs[j] = mk<double>(c, 0.0);
    sn[j] = sgn;
    Hc[j] = c * a + hn * sgn;
    Hc[j + 1] = mk<double>(0.0, 0.0);
    const zc s1 = mk<double>(0.0, 0.0) - conj(sgn) * sj;
    s[j + 1] = s1;
    s[j] = c * sj;
    const double err = sqrt(s1.x * s1.x + s1.y * s1.y) / bn;
    st.err[r] = err;
    st.jdone[r] = j + 1;
    st.nprec[r] = np0 + 1;
    if (!(err > tol)) st.done[r] = 1;
    if (err != err) st.done[r] = 2;
}
// grid nrhs, block 32*(nv+1) [i0 == 0] or 32*nv threads: warp w sums quantity w of the multidot group
__global__ void __launch_bounds__(32 * (HH_MAXV + 2)) k_gmres_hcol_mw(GmresState st, const zc* __restrict__ partial, int nblk, int j, int i0,
                                                                     int nv) {
    __shared__ zc g[HH_MAXV + 2];
    const int r = blockIdx.x, nrhs = gridDim.x, w = threadIdx.x >> 5;
    const zc v = reduce_partials(partial, (int64_t)w * nrhs + r, nblk);
    if ((threadIdx.x & 31) == 0) g[w] = v;
    __syncthreads();
    if (w == 0) gmres_hcol_w0(st, g, r, j, i0, nv);
}
// grid nrhs, block 32 threads
__global__ void k_gmres_givens_mw(GmresState st, const zc* __restrict__ partial, int nblk, int j, double tol, int est) {
    const int r = blockIdx.x;
    zc nn = mk<double>(0.0, 0.0);
    if (!est) nn = reduce_partials(partial, r, nblk);
    gmres_givens_w0(st, r, j, tol, nn.x, est);
}
// k_gmres_small_step with one warp per quantity: grid nrhs, block 32*(j+3) threads -- warps 0..j the dots of column j,
// warp j+1 its ||w||^2, warp j+2 the pending ||v~_j||^2 of column j-1 (idle when flags & 1 == 0)
__global__ void __launch_bounds__(32 * (HH_MAXV + 2)) k_gmres_small_step_mw(GmresState st, const zc* __restrict__ dots, int nblk_d,
                                                                           const zc* __restrict__ norm, int nblk_n, int j, int flags) {
    __shared__ zc g[HH_MAXV + 2];
    const int r = blockIdx.x, nrhs = gridDim.x, w = threadIdx.x >> 5;
    const int nv = j + 1;
    zc v = mk<double>(0.0, 0.0);
    if (w <= nv) v = reduce_partials(dots, (int64_t)w * nrhs + r, nblk_d);
    else if (flags & 1) v = reduce_partials(norm, r, nblk_n);
    if ((threadIdx.x & 31) == 0) g[w] = v;
    __syncthreads();
    if (w != 0) return;
    if (flags & 1) gmres_givens_w0(st, r, j - 1, 0.0, g[nv + 1].x, 0);
    __syncwarp();
    gmres_hcol_w0(st, g, r, j, 0, nv);
    __syncwarp();
    if (flags & 2) gmres_givens_w0(st, r, j, 0.0, 0.0, 1);
}

// y = H(1:jd,1:jd) \ s(1:jd); stored as y_i/d_i (coefficients of the stored z~_i), zero beyond jd
__global__ void k_gmres_solve_y(GmresState st, int nrhs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrhs) return;
    const int m = st.m, ldh = m + 1;
    const int jd = st.jdone[r];
    const zc* H = st.H + (int64_t)r * ldh * m;
    const zc* s = st.s + (int64_t)r * ldh;
    const double* d = st.d + (int64_t)r * ldh;
    zc* y = st.y + (int64_t)r * m;
    for (int i = 0; i < m; ++i) y[i] = mk<double>(0.0, 0.0);
    for (int i = jd - 1; i >= 0; --i) {
        zc acc = s[i];
        for (int k = i + 1; k < jd; ++k) acc = acc - H[(int64_t)k * ldh + i] * y[k];
        y[i] = cdiv(acc, H[(int64_t)i * ldh + i]);
    }
    for (int i = 0; i < jd; ++i) y[i] = (d[i] > 0.0 ? 1.0 / d[i] : 0.0) * y[i];
    st.jdone[r] = 0;
}

// start of a cycle: beta = ||r|| from norm partials; v~_0 = r is used unscaled (d_0 = beta), s = beta e1,
// err = beta/bnorm.  first != 0: this is ||b||: record bnorm, mark zero right-hand sides done.
__global__ void k_gmres_begin(GmresState st, const zc* __restrict__ partial, int nblk, int first, double tol) {
    const int r = blockIdx.x;
    const zc nn = reduce_partials(partial, r, nblk);
    if ((threadIdx.x & 31) != 0) return;
    const int ldh = st.m + 1;
    const double beta = sqrt(nn.x);
    if (first) {
        st.bnorm[r] = beta;
        st.nprec[r] = 0;
        st.jdone[r] = 0;
        st.done[r] = (beta == 0.0) ? 1 : 0;
        st.err[r] = (beta == 0.0) ? 0.0 : 1.0;
        if (beta != beta) st.done[r] = 2;
    } else if (!st.done[r]) {
        const double err = beta / st.bnorm[r];
        st.err[r] = err;
        if (!(err > tol)) st.done[r] = 1;
        if (err != err) st.done[r] = 2;
    }
    zc* s = st.s + (int64_t)r * ldh;
    for (int i = 0; i < ldh; ++i) s[i] = mk<double>(0.0, 0.0);
    s[0] = mk<double>(beta, 0.0);
    st.d[(int64_t)r * ldh] = beta;
    st.scale[r] = mk<double>(0.0, 0.0);
}

// Per-RHS solver state for the host without the copy engine: the values are written into mapped pinned host memory by
// this kernel (then a stream synchronise).  A cudaMemcpyAsync of a few bytes queues behind any bulk device-to-host
// transfer that is in flight on another stream (the previous sub-batch's solutions on their way to the caller): the
// Krylov loop's convergence check then waits for gigabytes to drain and the solve no longer overlaps the copy.
__global__ void k_publish_state(const int* __restrict__ done, const int* __restrict__ nprec, const double* __restrict__ err,
                                int* __restrict__ h_done, int* __restrict__ h_nprec, double* __restrict__ h_err, int n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        h_done[r] = done[r];
        if (nprec) h_nprec[r] = nprec[r];
        if (err) h_err[r] = err[r];
    }
    __threadfence_system();
}

// ---- BiCGSTAB: per-RHS scalars and the p-update -------------------------------------------------
struct BicgState {
    zc *rho, *rho_old, *alpha, *omega, *beta;  // per RHS
    zc *neg_alpha, *neg_omega, *ao;            // helper coefficient slots; ao[r][2] = (alpha, omega)
    double *bnorm, *err;
    int *done, *half, *nprec, *iters;
};

enum { BICG_INIT = 0, BICG_RHO = 1, BICG_ALPHA = 2, BICG_HALF = 3, BICG_OMEGA = 4, BICG_END = 5 };

// stage machine of preconditioned BiCGSTAB; `partial` holds the reductions the stage needs
// (layout of k_multidot: quantity i, RHS r -> partial[(i*nrhs + r)*nblk + .]).
__global__ void k_bicg_scalars(BicgState st, const zc* __restrict__ partial, int nblk, int stage, double tol) {
    const int r = blockIdx.x, nrhs = gridDim.x;
    const zc q0 = reduce_partials(partial, r, nblk);
    zc q1 = mk<double>(0.0, 0.0);
    if (stage == BICG_OMEGA) q1 = reduce_partials(partial, (int64_t)nrhs + r, nblk);
    if ((threadIdx.x & 31) != 0) return;
    const zc zero = mk<double>(0.0, 0.0);
    if (stage == BICG_INIT) {  // q0 = |b|^2
        const double bn = sqrt(q0.x);
        st.bnorm[r] = bn;
        st.err[r] = bn == 0.0 ? 0.0 : 1.0;
        st.done[r] = bn == 0.0 ? 1 : (bn != bn ? 2 : 0);
        st.half[r] = 0;
        st.nprec[r] = 0;
        st.iters[r] = 0;
        st.rho_old[r] = st.alpha[r] = st.omega[r] = mk<double>(1.0, 0.0);
        st.beta[r] = zero;
        return;
    }
    if (st.done[r]) {
        st.beta[r] = zero;
        st.neg_alpha[r] = zero;
        st.neg_omega[r] = zero;
        st.ao[2 * r] = zero;
        st.ao[2 * r + 1] = zero;
        return;
    }
    if (stage == BICG_RHO) {  // q0 = <rt, r>
        st.rho[r] = q0;
        if (st.iters[r] == 0) {
            st.beta[r] = zero;
        } else {
            st.beta[r] = cdiv(q0, st.rho_old[r]) * cdiv(st.alpha[r], st.omega[r]);
        }
        st.nprec[r] += 1;  // the p-hat preconditioner application that follows
    } else if (stage == BICG_ALPHA) {  // q0 = <rt, v>
        const zc a = cdiv(st.rho[r], q0);
        st.alpha[r] = a;
        st.neg_alpha[r] = zero - a;
    } else if (stage == BICG_HALF) {  // q0 = |s|^2
        const double err = sqrt(q0.x) / st.bnorm[r];
        st.err[r] = err;
        if (!(err > tol)) st.half[r] = 1;
        if (err != err) st.done[r] = 2;
        if (!st.half[r]) st.nprec[r] += 1;  // the s-hat application is only counted when needed
    } else if (stage == BICG_OMEGA) {  // q0 = <s, t>, q1 = |t|^2
        zc w = zero;
        if (!st.half[r] && q1.x > 0.0) w = (1.0 / q1.x) * conj(q0);
        st.omega[r] = w;
        st.neg_omega[r] = zero - w;
        st.ao[2 * r] = st.alpha[r];
        st.ao[2 * r + 1] = w;
    } else if (stage == BICG_END) {  // q0 = |r|^2
        st.iters[r] += 1;
        if (st.half[r]) {
            st.done[r] = 1;
        } else {
            const double err = sqrt(q0.x) / st.bnorm[r];
            st.err[r] = err;
            if (!(err > tol)) st.done[r] = 1;
            if (err != err) st.done[r] = 2;
        }
        st.rho_old[r] = st.rho[r];
    }
}

// p = r + beta (p - omega v)
template <typename T>
__global__ void __launch_bounds__(256) k_bicg_p(cx<T>* __restrict__ p, const cx<T>* __restrict__ r,
                                                const cx<T>* __restrict__ v, int64_t N, int64_t ld,
                                                const zc* __restrict__ beta, const zc* __restrict__ omega) {
    const int c = blockIdx.y;
    const cx<T> be = mk<T>((T)beta[c].x, (T)beta[c].y);
    const cx<T> om = mk<T>((T)omega[c].x, (T)omega[c].y);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < N; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = (int64_t)c * ld + q;
        const cx<T> t = p[o] - om * v[o];
        p[o] = r[o] + be * t;
    }
}

}  // namespace hh
