// ubench2.cu -- isolates which access of the z-marching stencil limits bandwidth (tuning only).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "../helmholtz.jl_b200/csrc/hh_common.cuh"
using namespace hh;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
enum { F_B = 1, F_MG = 2, F_Y = 4, F_X = 8, F_MATH = 16, F_DIV = 32, F_CARR = 64 };
template <typename T, int KB, int FLAGS, int MINB>
__global__ void __launch_bounds__(256, MINB) k_test(const cx<T>* __restrict__ x, const cx<T>* __restrict__ b, const T* __restrict__ m,
        const T* __restrict__ g, cx<T>* __restrict__ out, int n0, int n1, int n2, int64_t ld, int nrhs, int zchunk, int groups) {
    const int i = (blockIdx.x / groups) * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= n0 || j >= n1) return;
    const int r0 = (blockIdx.x % groups) * KB;
    const int z0 = blockIdx.z * zchunk, z1 = min(n2, z0 + zchunk);
    const int64_t sy = n0, sz = (int64_t)n0 * n1, pxy = i + sy * j;
    const int64_t oxm = i > 0 ? -1 : 0, oxp = i < n0 - 1 ? 1 : 0, oym = j > 0 ? -sy : 0, oyp = j < n1 - 1 ? sy : 0;
    const cx<T>* xr[KB]; cx<T> xm[KB], xc[KB], xp[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) { xr[q] = x + (int64_t)min(r0 + q, nrhs - 1) * ld + pxy; xc[q] = xr[q][(int64_t)z0 * sz]; xm[q] = xc[q]; }
#pragma unroll 1
    for (int z = z0; z < z1; ++z) {
        const int64_t zo = (int64_t)z * sz, p = pxy + zo;
#pragma unroll
        for (int q = 0; q < KB; ++q) xp[q] = z == n2 - 1 ? xc[q] : xr[q][zo + sz];
        cx<T> c = mk<T>(T(1), T(0));
        if (FLAGS & F_MG) c.x = m[p] * T(0.5) + g[p];
        if (FLAGS & F_CARR) c = reinterpret_cast<const cx<T>*>(m)[p];
        if (FLAGS & F_MATH) {
            const T mv = m[p], gv = g[p] * T(0.1);
            T re = -mv * (T(88.8) + T(0.01) * gv), im = -mv * (T(0.01) - T(88.8) * gv) + T(17.7) * mv;
            T sf = (i == 0 || i == n0 - 1 ? T(188) : T(0)) + (j == 0 || j == n1 - 1 ? T(188) : T(0));
            if (z == n2 - 1) sf += T(188);
            if (sf != T(0)) im += sf * sqrt(mv);
            re += T(400) + ((z == 0 || z == n2 - 1) ? T(2) : T(2)) * T(100);
            c = mk<T>(re, im);
        }
        cx<T> dinv = mk<T>(T(1), T(0));
        if (FLAGS & F_DIV) dinv = rdiv(T(0.8), c);
#pragma unroll
        for (int q = 0; q < KB; ++q) {
            const cx<T>* xq = xr[q] + zo;
            cx<T> a = c * xc[q];
            if (FLAGS & F_DIV) a = dinv * a;
            if (FLAGS & F_X) { rfma(a, T(-1), xq[oxm]); rfma(a, T(-1), xq[oxp]); }
            if (FLAGS & F_Y) { rfma(a, T(-1), xq[oym]); rfma(a, T(-1), xq[oyp]); }
            rfma(a, T(-1), xm[q]); rfma(a, T(-1), xp[q]);
            if (r0 + q < nrhs) {
                const int64_t o = (int64_t)(r0 + q) * ld + p;
                if (FLAGS & F_B) out[o] = b[o] - a; else out[o] = a;
            }
            xm[q] = xc[q]; xc[q] = xp[q];
        }
    }
}
// flat streaming reference: out = b - c*x elementwise over everything
template <typename T>
__global__ void __launch_bounds__(256) k_stream(const cx<T>* __restrict__ x, const cx<T>* __restrict__ b, cx<T>* __restrict__ out, int64_t n) {
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256) out[p] = b[p] - x[p];
}
template <typename T, int KB, int FLAGS, int MINB>
void run(const char* name, const cx<T>* x, const cx<T>* b, const T* m, const T* g, cx<T>* o, int n, int nrhs, int pref, int bx) {
    const int by = 256 / bx; const int groups = (nrhs + KB - 1) / KB; const int tx = (n + bx - 1) / bx, ty = (n + by - 1) / by;
    int nzc = std::max(1, (n + pref - 1) / pref), zchunk = (n + nzc - 1) / nzc; nzc = (n + zchunk - 1) / zchunk;
    dim3 gr(tx * groups, ty, nzc), blk(bx, by, 1);
    int64_t N = (int64_t)n * n * n;
    cudaEvent_t a, c; cudaEventCreate(&a); cudaEventCreate(&c);
    k_test<T, KB, FLAGS, MINB><<<gr, blk>>>(x, b, m, g, o, n, n, n, N, nrhs, zchunk, groups); CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) k_test<T, KB, FLAGS, MINB><<<gr, blk>>>(x, b, m, g, o, n, n, n, N, nrhs, zchunk, groups);
    cudaEventRecord(c); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, c); ms /= 5;
    const double S = sizeof(cx<T>);
    double bytes = (2 + ((FLAGS & F_B) ? 1 : 0)) * S * N * nrhs + ((FLAGS & F_MG) ? 2.0 * sizeof(T) * N : 0);
    printf("%-40s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6); fflush(stdout);
}
template <typename T> void bench(int n, int nrhs) {
    const int64_t N = (int64_t)n * n * n; const double S = sizeof(cx<T>);
    cx<T>*x, *b, *o; T *m, *g;
    CK(cudaMalloc(&x, N * nrhs * S)); CK(cudaMalloc(&b, N * nrhs * S)); CK(cudaMalloc(&o, N * nrhs * S)); CK(cudaMalloc(&m, 2 * N * sizeof(T))); CK(cudaMalloc(&g, N * sizeof(T)));
    cudaMemset(x, 0, N * nrhs * S); cudaMemset(b, 0, N * nrhs * S); cudaMemset(m, 0, 2 * N * sizeof(T)); cudaMemset(g, 0, N * sizeof(T));
    printf("== %s n=%d nrhs=%d\n", sizeof(T) == 8 ? "c128" : "c64", n, nrhs);
    {
        cudaEvent_t a, c; cudaEventCreate(&a); cudaEventCreate(&c);
        for (int grid : {148 * 8, 148 * 32, 148 * 128}) {
            k_stream<T><<<grid, 256>>>(x, b, o, N * nrhs); CK(cudaDeviceSynchronize());
            cudaEventRecord(a); for (int r = 0; r < 5; ++r) k_stream<T><<<grid, 256>>>(x, b, o, N * nrhs); cudaEventRecord(c); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, a, c); ms /= 5; printf("stream out=b-x grid %6d                 %8.3f ms  %7.1f GB/s\n", grid, ms, 3 * S * N * nrhs / ms / 1e6);
        }
    }
#define R(KB, FL, MINB, PREF, BX) run<T, KB, FL, MINB>("KB" #KB " flags " #FL " MINB" #MINB " P" #PREF " bx" #BX, x, b, m, g, o, n, nrhs, PREF, BX);
    R(2, 0, 4, 32, 32) R(2, F_B, 4, 32, 32) R(2, F_B | F_MG, 4, 32, 32) R(2, F_B | F_MG | F_Y, 4, 32, 32) R(2, F_B | F_MG | F_X, 4, 32, 32) R(2, F_B | F_MG | F_X | F_Y, 4, 32, 32)
    R(2, F_B | F_MG | F_X | F_Y | F_MATH, 4, 32, 32) R(2, F_B | F_MG | F_X | F_Y | F_MATH | F_DIV, 4, 32, 32) R(2, F_B | F_CARR | F_X | F_Y, 4, 32, 32) R(2, F_B | F_CARR | F_X | F_Y | F_DIV, 4, 32, 32)
    R(1, F_B | F_MG | F_X | F_Y | F_MATH | F_DIV, 6, 32, 32) R(1, F_B | F_CARR | F_X | F_Y | F_DIV, 6, 32, 32) R(1, F_B | F_CARR | F_X | F_Y, 6, 32, 32)
    cudaFree(x); cudaFree(b); cudaFree(o); cudaFree(m); cudaFree(g);
}
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 257, nrhs = argc > 2 ? atoi(argv[2]) : 8;
    bench<double>(n, nrhs); bench<float>(n, nrhs);
    return 0;
}
