// hh_common.cuh -- shared types, complex arithmetic and reduction helpers.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>

namespace hh {

// Interleaved complex number, same memory layout as Julia's Complex{T} (re,im).
// Aligned to its size so that one element is one 8/16-byte vector load.
template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T x, y;
};

template <typename T>
__host__ __device__ __forceinline__ cx<T> mk(T a, T b) {
    cx<T> r;
    r.x = a;
    r.y = b;
    return r;
}
template <typename T>
__host__ __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T>
__host__ __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T>
__host__ __device__ __forceinline__ cx<T> operator*(cx<T> a, cx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T>
__host__ __device__ __forceinline__ cx<T> operator*(T s, cx<T> a) { return mk<T>(s * a.x, s * a.y); }
template <typename T>
__host__ __device__ __forceinline__ cx<T> conj(cx<T> a) { return mk<T>(a.x, -a.y); }
__host__ __device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__host__ __device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
// acc += a*b as FOUR chained fused multiply-adds.  Written out because `acc.x += a.x*b.x - a.y*b.y` compiles to
// DMUL + DFMA + DADD per component (the compiler may not re-associate), i.e. 6 instead of 4 FP64 instructions per
// complex multiply-add: the 27-point kernels are bound by exactly that pipe.
template <typename T>
__host__ __device__ __forceinline__ void cfma(cx<T>& acc, cx<T> a, cx<T> b) {
    acc.x = fma_t(a.x, b.x, acc.x);
    acc.x = fma_t(-a.y, b.y, acc.x);
    acc.y = fma_t(a.x, b.y, acc.y);
    acc.y = fma_t(a.y, b.x, acc.y);
}
// acc += s*b (s real)
template <typename T>
__host__ __device__ __forceinline__ void rfma(cx<T>& acc, T s, cx<T> b) {
    acc.x += s * b.x;
    acc.y += s * b.y;
}
// s / c  (s real)
template <typename T>
__host__ __device__ __forceinline__ cx<T> rdiv(T s, cx<T> c) {
    T d = s / (c.x * c.x + c.y * c.y);
    return mk<T>(c.x * d, -c.y * d);
}
template <typename T>
__host__ __device__ __forceinline__ cx<T> cdiv(cx<T> a, cx<T> b) {
    T d = T(1) / (b.x * b.x + b.y * b.y);
    return mk<T>((a.x * b.x + a.y * b.y) * d, (a.y * b.x - a.x * b.y) * d);
}

typedef cx<double> zc;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum `v` over the thread block; result valid in thread 0.  `sm` needs >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sm) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect sm reuse between successive calls
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? sm[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

}  // namespace hh
