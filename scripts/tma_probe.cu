// tma_probe.cu -- which FLOAT32 tensor-map boxes does the TMA unit accept? (debugging aid, not product)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "../helmholtz.jl_b200/csrc/hh_kernels.cuh"
using namespace hh;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ TmaDesc tm, int c0, int c1, int c2, int c3, uint32_t bytes, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 32768);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, bytes);
        tma_load_4d(sm, &tm, c0, c1, c2, c3, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    if (threadIdx.x < 8) out[threadIdx.x] = reinterpret_cast<float*>(sm)[threadIdx.x];
}
int main(int argc, char** argv) {
    int only = argc > 1 ? atoi(argv[1]) : -1; int ci = -1;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    float* d; cudaMalloc(&d, 1 << 24); float* o; cudaMalloc(&o, 64);
    float h[64]; for (int i = 0; i < 64; ++i) h[i] = i + 1; cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    struct Case { int sy2, n0x2, box0, c0; CUtensorMapDataType dt; int es; const char* name; } cases[] = {
        {36, 34, 68, -2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "f32 dims34 box68 c0=-2"},
        {36, 34, 68, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "f32 dims34 box68 c0=0"},
        {36, 34, 64, -2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "f32 dims34 box64 c0=-2"},
        {36, 34, 64, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "f32 dims34 box64 c0=0"},
        {36, 36, 68, -2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "f32 dims36 box68 c0=-2"},
        {36, 36, 68, -4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "f32 dims36 box68 c0=-4"},
        {36, 34, 68, -4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "f32 dims34 box68 c0=-4"},
        {34, 34, 68, -2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64 dims34 box68 c0=-2"},
        {18, 17, 34, -1, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, "f64-as-c64 dims17 box34 c0=-1"},
    };
    for (auto& c : cases) {
        ++ci; if (only >= 0 && ci != only) continue;
        TmaDesc tm;
        cuuint64_t dims[4] = {(cuuint64_t)c.n0x2, 17, 17, 2};
        cuuint64_t str[3] = {(cuuint64_t)c.es * c.sy2, (cuuint64_t)c.es * c.sy2 * 17, (cuuint64_t)c.es * c.sy2 * 17 * 17};
        cuuint32_t box[4] = {(cuuint32_t)c.box0, 10, 1, 2}; cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc((CUtensorMap*)&tm, c.dt, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-34s encode failed %d\n", c.name, (int)r); continue; }
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
        k<<<1, 32, 32768 + 64>>>(tm, c.c0, -1, 0, 0, (uint32_t)(c.box0 * 10 * 2 * c.es), o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%-34s %s\n", c.name, cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("(context lost; remaining cases skipped)\n"); break; }
    }
    return 0;
}
