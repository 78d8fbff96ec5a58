// Probe matrix: which TMA / bulk-copy forms run on this box?  One variant per process (a fault kills the context).
//   0  cp.async.bulk 1-D (no tensor map)                 1  libcu++ 2-D tensor copy, map = __grid_constant__
//   2  raw PTX 2-D                                       3  raw PTX 4-D, map = __grid_constant__
//   4  raw PTX 4-D, map in global memory                 5  libcu++ 4-D
//   6  raw PTX 4-D, box depth 1 (c box = 1)              7  raw PTX 3-D
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cde = cuda::device::experimental;
using barrier_t = cuda::barrier<cuda::thread_scope_block>;

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(unsigned long long *bar, unsigned ph) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(s32(bar)), "r"(ph) : "memory");
}
__device__ int g_fence = 1;
__device__ __forceinline__ void bar_setup(unsigned long long *bar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (g_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
}

__global__ void k_bulk1d(const float *src, float *out) {
    __shared__ __align__(128) float tile[256];
    __shared__ __align__(8) unsigned long long bar;
    bar_setup(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(tile)), "l"(src), "r"(1024), "r"(s32(&bar)) : "memory");
    }
    wait(&bar, 0);
    if (threadIdx.x < 4) out[threadIdx.x] = tile[threadIdx.x];
}

template <int ND, bool GLOBAL_MAP>
__global__ void k_raw(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, int bytes, float *out, int x) {
    __shared__ __align__(128) float tile[2048];
    __shared__ __align__(8) unsigned long long bar;
    bar_setup(&bar);
    if (threadIdx.x == 0) {
        const CUtensorMap *mp = GLOBAL_MAP ? gmap : &pmap;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        if (ND == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(tile)), "l"((unsigned long long)mp), "r"(s32(&bar)), "r"(x), "r"(2) : "memory");
        else if (ND == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(tile)), "l"((unsigned long long)mp), "r"(s32(&bar)), "r"(x), "r"(2), "r"(1) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(tile)), "l"((unsigned long long)mp), "r"(s32(&bar)), "r"(x), "r"(2), "r"(1), "r"(0) : "memory");
    }
    wait(&bar, 0);
    if (threadIdx.x < 4) out[threadIdx.x] = tile[threadIdx.x];
}

template <int ND>
__global__ void k_cxx(const __grid_constant__ CUtensorMap pmap, int bytes, float *out) {
    __shared__ __align__(128) float tile[2048];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier_t bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier_t::arrival_token tok;
    if (threadIdx.x == 0) {
        if (ND == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(tile, &pmap, 4, 2, bar);
        else cde::cp_async_bulk_tensor_4d_global_to_shared(tile, &pmap, 4, 2, 1, 0, bar);
        tok = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else {
        tok = bar.arrive();
    }
    bar.wait(std::move(tok));
    if (threadIdx.x < 4) out[threadIdx.x] = tile[threadIdx.x];
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int x = argc > 2 ? atoi(argv[2]) : 4;
    const int promo = argc > 3 ? atoi(argv[3]) : 0;
    const int fence = argc > 4 ? atoi(argv[4]) : 1;
    cudaMemcpyToSymbol(g_fence, &fence, sizeof(int));
    printf("x=%d promo=%d fence=%d\n", x, promo, fence);
    const int W = 72, H = 60, C = 128, B = 2;
    const long n = (long)W * H * C * B;
    float *img, *out;
    cudaMalloc(&img, sizeof(float) * n);
    cudaMalloc(&out, 64);
    cudaMemset(out, 0, 64);
    float *h = (float *)malloc(sizeof(float) * n);
    for (long i = 0; i < n; ++i) h[i] = (float)(i % 100003);
    cudaMemcpy(img, h, sizeof(float) * n, cudaMemcpyHostToDevice);
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("variant %d: entry point %s q=%d p=%p\n", variant, cudaGetErrorString(ge), (int)q, p);
    Enc enc = (Enc)p;
    int drv = 0, rt = 0;
    cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("driver %d runtime %d device %s cc %d.%d\n", drv, rt, prop.name, prop.major, prop.minor);

    CUtensorMap map;
    int nd = 4;
    if (variant == 1 || variant == 2) nd = 2;
    if (variant == 7) nd = 3;
    const int cdepth = (variant == 6) ? 1 : 4;
    cuuint64_t gd[4] = {W, H, C, B};
    cuuint64_t gs[3] = {W * 4ull, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t box[4] = {8, 8, (cuuint32_t)cdepth, 1}, es[4] = {1, 1, 1, 1};
    if (nd == 2) gd[1] = (cuuint64_t)H * C * B;            // rows of all planes stacked
    if (nd == 3) gd[2] = (cuuint64_t)C * B;
    int bytes = 8 * 8 * 4;
    if (nd >= 3) bytes *= cdepth;
    if (variant != 0) {
        CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, nd, img, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode nd=%d rc=%d bytes=%d\n", nd, (int)rc, bytes);
        if (rc != CUDA_SUCCESS) return 2;
    }
    CUtensorMap *gmap;
    cudaMalloc(&gmap, sizeof(CUtensorMap));
    cudaMemcpy(gmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    long base = 0;
    switch (variant) {
        case 0: k_bulk1d<<<1, 32>>>(img + 256, out); base = 256; break;
        case 1: k_cxx<2><<<1, 32>>>(map, bytes, out); base = 2L * W + 4; break;
        case 2: k_raw<2, false><<<1, 32>>>(map, gmap, bytes, out, x); base = 2L * W + x; break;
        case 3: case 6: k_raw<4, false><<<1, 32>>>(map, gmap, bytes, out, x); base = ((0L * C + 1) * H + 2) * W + x; break;
        case 4: k_raw<4, true><<<1, 32>>>(map, gmap, bytes, out, x); base = ((0L * C + 1) * H + 2) * W + x; break;
        case 5: k_cxx<4><<<1, 32>>>(map, bytes, out); base = ((0L * C + 1) * H + 2) * W + 4; break;
        case 7: k_raw<3, false><<<1, 32>>>(map, gmap, bytes, out, x); base = ((0L * C + 1) * H + 2) * W + x; break;
    }
    cudaError_t le = cudaGetLastError();
    cudaError_t e = cudaDeviceSynchronize();
    float r[4] = {-1, -1, -1, -1};
    cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost);
    printf("variant %d: launch=%s sync=%s  got %.0f %.0f  want %.0f %.0f  => %s\n", variant, cudaGetErrorString(le), cudaGetErrorString(e), r[0], r[1],
           h[base], h[base + 1], (e == cudaSuccess && r[0] == h[base] && r[1] == h[base + 1]) ? "PASS" : "FAIL");
    return e == cudaSuccess ? 0 : 1;
}
