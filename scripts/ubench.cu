// ubench.cu -- stand-alone tuning harness for the z-marching stencil kernels (not part of the product).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/ubench scripts/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../helmholtz.jl_b200/csrc/hh_kernels.cuh"
using namespace hh;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <typename T> FineOp<T> make_op(int n, const T* m, const T* g) {
    FineOp<T> op; op.m = m; op.g = g; op.a = 88.8; op.b = 0; op.inv_wr = 1.0 / 9.42; op.shift_w2 = 0.2 * 88.8;
    for (int d = 0; d < 3; ++d) { op.somm[d] = 9.42 * 2 / 0.1; op.ih2[d] = 100.0; op.n[d] = n; }
    op.BC = 2; op.neumann_top = 1; op.adj = 0; return op;
}
static void zchunks(int n2, int tiles, int groups, int pref, int& zchunk, int& nzc) {
    int want = (592 + tiles * groups - 1) / (tiles * groups);
    nzc = std::max((n2 + pref - 1) / pref, want); nzc = std::max(1, std::min(nzc, n2));
    zchunk = (n2 + nzc - 1) / nzc; nzc = (n2 + zchunk - 1) / zchunk;
}
static int g_txw = 32;
template <typename T, int MODE, int KB, int TY, int MINB>
float run_fine(const FineOp<T>& op, const cx<T>* x, const cx<T>* b, cx<T>* out, int n, int nrhs, int pref, int reps) {
    const int bx = g_txw, by = 32 * TY / g_txw;
    const int groups = (nrhs + KB - 1) / KB; const int tx = (n + bx - 1) / bx, ty = (n + by - 1) / by;
    int zchunk, nzc; zchunks(n, tx * ty, groups, pref, zchunk, nzc);
    dim3 g(tx * groups, ty, nzc), blk(bx, by, 1);
    cudaEvent_t a, c; cudaEventCreate(&a); cudaEventCreate(&c);
    int64_t N = (int64_t)n * n * n;
    k_fine3d_zmarch<T, MODE, KB, TY, MINB><<<g, blk>>>(op, x, b, out, N, nrhs, (T)0.8, zchunk, groups);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) k_fine3d_zmarch<T, MODE, KB, TY, MINB><<<g, blk>>>(op, x, b, out, N, nrhs, (T)0.8, zchunk, groups);
    cudaEventRecord(c); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, c); return ms / reps;
}
template <typename T, int MODE, int KB, int TY, int MINB>
float run_coarse(const CoarseOp<T>& op, const cx<T>* x, const cx<T>* b, cx<T>* out, int n, int nrhs, int pref, int reps) {
    const int groups = (nrhs + KB - 1) / KB; const int tx = (n + 31) / 32, ty = (n + TY - 1) / TY;
    int zchunk, nzc; zchunks(n, tx * ty, groups, pref, zchunk, nzc);
    dim3 g(tx * groups, ty, nzc), blk(32, TY, 1);
    cudaEvent_t a, c; cudaEventCreate(&a); cudaEventCreate(&c);
    int64_t N = (int64_t)n * n * n;
    k_coarse3d_zmarch<T, MODE, KB, TY, MINB><<<g, blk>>>(op, x, b, out, N, nrhs, zchunk, groups);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) k_coarse3d_zmarch<T, MODE, KB, TY, MINB><<<g, blk>>>(op, x, b, out, N, nrhs, zchunk, groups);
    cudaEventRecord(c); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, c); return ms / reps;
}
template <typename T> void bench(int n, int nc, int nrhs) {
    const int64_t N = (int64_t)n * n * n, Nc = (int64_t)nc * nc * nc;
    const double S = sizeof(cx<T>), CR = sizeof(T);
    cx<T>*x, *b, *o; T *m, *g; cx<T>* coef; cx<T>* dinv;
    CK(cudaMalloc(&x, N * nrhs * S)); CK(cudaMalloc(&b, N * nrhs * S)); CK(cudaMalloc(&o, N * nrhs * S));
    CK(cudaMalloc(&m, N * CR)); CK(cudaMalloc(&g, N * CR)); CK(cudaMalloc(&coef, Nc * 27 * S)); CK(cudaMalloc(&dinv, Nc * S));
    cudaMemset(x, 0, N * nrhs * S); cudaMemset(b, 0, N * nrhs * S); cudaMemset(m, 0, N * CR); cudaMemset(g, 0, N * CR);
    cudaMemset(coef, 0, Nc * 27 * S); cudaMemset(dinv, 0, Nc * S);
    FineOp<T> op = make_op<T>(n, m, g);
    CoarseOp<T> cop; cop.coef = coef; cop.dinv = dinv; cop.n[0] = cop.n[1] = cop.n[2] = nc;
    const char* tn = sizeof(T) == 8 ? "c128" : "c64";
    auto rf = [&](const char* name, float ms, double bytes) { printf("%s %-34s %8.3f ms  %7.1f GB/s\n", tn, name, ms, bytes / ms / 1e6); fflush(stdout); };
    const double ba = (2 * S * nrhs + 2 * CR) * N, bj = (3 * S * nrhs + 2 * CR) * N;
#define FA(KB, TY, MINB, PREF) rf("fine apply  KB" #KB " TY" #TY " MINB" #MINB " P" #PREF, run_fine<T, MODE_APPLY, KB, TY, MINB>(op, x, b, o, n, nrhs, PREF, 5), ba);
#define FJ(KB, TY, MINB, PREF) rf("fine jacobi KB" #KB " TY" #TY " MINB" #MINB " P" #PREF, run_fine<T, MODE_JACOBI, KB, TY, MINB>(op, x, b, o, n, nrhs, PREF, 5), bj);
    for (int txw : {32, 64, 128, 256}) {
        g_txw = txw;
        printf("tile %d x %d\n", txw, 256 / txw);
        FA(2, 8, 4, 32) FJ(2, 8, 4, 32) FJ(1, 8, 6, 32)
    }
    cudaFree(x); cudaFree(b); cudaFree(o); cudaFree(m); cudaFree(g); cudaFree(coef); cudaFree(dinv);
}
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 257, nrhs = argc > 2 ? atoi(argv[2]) : 8;
    bench<double>(n, (n + 1) / 2, nrhs);
    bench<float>(n, (n + 1) / 2, nrhs);
    return 0;
}
