// RoIAlign forward for contiguous NCHW maps (the reference's own layout), TMA-staged.
//
// In NCHW the four taps of a sample are two 8-byte pairs in two rows of ONE channel plane, and the 256 channels of a
// pixel are H*W*4 bytes apart: a thread-per-output gather (the reference kernel, and the plain NCHW kernel in
// roi_align.cu) issues 4 scattered scalar loads per output.  Here one CTA owns one box: the rectangle of the feature
// map that the box's taps touch, [ymin..ymax] x [xmin..xmax] x CC channels, is fetched by ONE `cp.async.bulk.tensor.4d`
// (TMA) per channel chunk straight into shared memory -- any alignment of the corner, rows coalesced by the copy engine,
// out-of-image elements zero-filled, no registers and no LSU instructions spent on the load -- double-buffered behind an
// mbarrier so chunk k+1 streams in while chunk k is interpolated.  The interpolation reads shared memory (4 LDS per
// output) and the crop is written fully coalesced ([C,P,P] of a box is contiguous).
//
// Tile shapes are baked into the tensor maps, so nine maps (8/16/32 x 8/16/32 pixels, channel depth chosen for ~32 KB
// per stage) are encoded per call and the CTA picks the smallest one that covers its rectangle.  Boxes whose footprint
// exceeds 32x32 pixels (RoIs pooled on much finer maps) take the direct-load path of roi_align.cu.
// Needs W % 4 == 0 (TMA global strides are multiples of 16 B), C % 64 == 0 and a 16-byte aligned map; the tile's x origin is
// rounded down to a multiple of 4 pixels (16 B), which the copy engine requires of the innermost coordinate.
#include <cuda.h>

#include "fi_common.cuh"

namespace fi {

constexpr int kTmaThreads = 256;
constexpr int kTmaMaxP = 32;                 // taps per axis kept in shared memory
constexpr int kStageFloats = 8192;           // 32 KB per stage
constexpr int kNumTileCfg = 9;

struct TileCfg { int bw, bh, cc; };
__host__ __device__ inline TileCfg tile_cfg(int k) {
    TileCfg c;
    c.bw = 8 << (k % 3);
    c.bh = 8 << (k / 3);
    int cc = kStageFloats / (c.bw * c.bh);
    c.cc = cc > 64 ? 64 : cc;
    return c;
}

struct TmaMaps {
    CUtensorMap m[kNumTileCfg];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int c, int b) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(b)
                 : "memory");
}

struct TmaTaps {
    int ylo[kTmaMaxP], yhi[kTmaMaxP], xlo[kTmaMaxP], xhi[kTmaMaxP];
    float yf[kTmaMaxP], xf[kTmaMaxP];
    unsigned char yin[kTmaMaxP], xin[kTmaMaxP];
};

__global__ void __launch_bounds__(kTmaThreads) crop_fwd_nchw_tma_kernel(const __grid_constant__ TmaMaps maps, const float *__restrict__ image,
                                                                       const float *__restrict__ boxes, const int *__restrict__ box_ind,
                                                                       const int *__restrict__ dst_row, int B, int H, int W, int ph, int pw, int C,
                                                                       int c_per_block, float extrap, float *__restrict__ crops) {
    extern __shared__ __align__(128) float stage_mem[];          // 2 stages x 32 KB (dynamic: above the 48 KB static limit)
    __shared__ TmaTaps t;
    __shared__ __align__(8) unsigned long long bar[2];
    __shared__ int rect[5];                      // xmin, ymin, tile cfg (-1: direct path, -2: nothing inside)
    const int r = blockIdx.x;
    const int c_begin = blockIdx.y * c_per_block;
    const int pp = ph * pw;
    const int b = box_ind[r];
    const long orow = dst_row ? (long)dst_row[r] : (long)r;
    float *out = crops + (orow * C + c_begin) * (long)pp;
    if (b < 0 || b >= B) {                       // reference leaves such rows at their zero fill (crop_and_resize_kernel.cu:34-38)
        for (int e = threadIdx.x; e < c_per_block * pp; e += kTmaThreads) out[e] = 0.f;
        return;
    }
    {
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const float sy = axis_step(y1, y2, H, ph), sx = axis_step(x1, x2, W, pw);
        for (int k = threadIdx.x; k < ph; k += kTmaThreads) {
            const AxisTap a = axis_sample(y1, y2, sy, k, H, ph);
            t.ylo[k] = a.lo; t.yhi[k] = a.hi; t.yf[k] = a.frac; t.yin[k] = a.inside;
        }
        for (int k = threadIdx.x; k < pw; k += kTmaThreads) {
            const AxisTap a = axis_sample(x1, x2, sx, k, W, pw);
            t.xlo[k] = a.lo; t.xhi[k] = a.hi; t.xf[k] = a.frac; t.xin[k] = a.inside;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1;
        for (int k = 0; k < ph; ++k) if (t.yin[k]) { ymin = min(ymin, t.ylo[k]); ymax = max(ymax, t.yhi[k]); }
        for (int k = 0; k < pw; ++k) if (t.xin[k]) { xmin = min(xmin, t.xlo[k]); xmax = max(xmax, t.xhi[k]); }
        int cfg = -2;
        if (xmax >= 0 && ymax >= 0) {
            // The innermost start coordinate of a tiled copy must be 16-byte aligned (x % 4 == 0 for fp32): an unaligned x
            // faults with "illegal instruction" on B200 (tools/probe/tma_probe.cu).  Rows and planes may start anywhere.
            xmin &= ~3;
            const int w = xmax - xmin + 1, h = ymax - ymin + 1;
            cfg = -1;
            if (w <= 32 && h <= 32) cfg = (h <= 8 ? 0 : (h <= 16 ? 1 : 2)) * 3 + (w <= 8 ? 0 : (w <= 16 ? 1 : 2));
        }
        rect[0] = xmin; rect[1] = ymin; rect[2] = cfg;
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int cfg = rect[2];
    const float *img = image + ((long)b * C + c_begin) * H * W;
    if (cfg < 0) {
        // nothing inside (all extrapolated) or a footprint larger than the largest tile: direct loads
        for (int e = threadIdx.x; e < c_per_block * pp; e += kTmaThreads) {
            const int c = e / pp, s = e - c * pp;
            const int i = s / pw, j = s - i * pw;
            float v = extrap;
            if (cfg == -1 && t.yin[i] && t.xin[j]) {
                const float *p = img + (long)c * H * W;
                const float *rt = p + (long)t.ylo[i] * W, *rb = p + (long)t.yhi[i] * W;
                const float top = lerp_rn(__ldg(rt + t.xlo[j]), __ldg(rt + t.xhi[j]), t.xf[j]);
                const float bot = lerp_rn(__ldg(rb + t.xlo[j]), __ldg(rb + t.xhi[j]), t.xf[j]);
                v = lerp_rn(top, bot, t.yf[i]);
            }
            __stcs(out + e, v);
        }
        return;
    }
    const TileCfg tc = tile_cfg(cfg);
    const int xmin = rect[0], ymin = rect[1];
    const int nchunks = c_per_block / tc.cc;
    const unsigned stage_bytes = (unsigned)(tc.bw * tc.bh * tc.cc) * 4u;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], stage_bytes);
        tma_load_4d(stage_mem, &maps.m[cfg], &bar[0], xmin, ymin, c_begin, b);
    }
    for (int k = 0; k < nchunks; ++k) {
        if (threadIdx.x == 0 && k + 1 < nchunks) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of that stage (iteration k-1) are done
            mbar_expect_tx(&bar[(k + 1) & 1], stage_bytes);
            tma_load_4d(stage_mem + ((k + 1) & 1) * kStageFloats, &maps.m[cfg], &bar[(k + 1) & 1], xmin, ymin, c_begin + (k + 1) * tc.cc, b);
        }
        mbar_wait(&bar[k & 1], (unsigned)((k >> 1) & 1));
        const float *tile = stage_mem + (k & 1) * kStageFloats;
        float *o = out + (long)k * tc.cc * pp;
        for (int e = threadIdx.x; e < tc.cc * pp; e += kTmaThreads) {
            const int c = e / pp, s = e - c * pp;
            const int i = s / pw, j = s - i * pw;
            float v = extrap;
            if (t.yin[i] && t.xin[j]) {
                const float *pl = tile + (c * tc.bh) * tc.bw;
                const float *rt = pl + (t.ylo[i] - ymin) * tc.bw - xmin, *rb = pl + (t.yhi[i] - ymin) * tc.bw - xmin;
                const float top = lerp_rn(rt[t.xlo[j]], rt[t.xhi[j]], t.xf[j]);          // crop_and_resize.c:102
                const float bot = lerp_rn(rb[t.xlo[j]], rb[t.xhi[j]], t.xf[j]);          // :103-104
                v = lerp_rn(top, bot, t.yf[i]);                                          // :106
            }
            __stcs(o + e, v);
        }
        __syncthreads();                          // the stage may be overwritten by the copy issued in the next iteration
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace fi

using namespace fi;

// Returns FI_ERR_UNSUPPORTED when the shape does not qualify (caller falls back to the plain NCHW kernel).
int fi_crop_forward_nchw_tma(const float *image, const float *boxes, const int *box_ind, const int *dst_row, int R, int B, int H, int W, int ph,
                             int pw, int C, float extrap, float *crops, cudaStream_t stream) {
    if (W % 4 != 0 || C % 64 != 0 || ph > kTmaMaxP || pw > kTmaMaxP || ((uintptr_t)image % 16) != 0 || R == 0) return FI_ERR_UNSUPPORTED;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return FI_ERR_UNSUPPORTED;
    TmaMaps maps;
    const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t gstride[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    const cuuint32_t estride[4] = {1, 1, 1, 1};
    for (int k = 0; k < kNumTileCfg; ++k) {
        const TileCfg tc = tile_cfg(k);
        const cuuint32_t box[4] = {(cuuint32_t)tc.bw, (cuuint32_t)tc.bh, (cuuint32_t)tc.cc, 1};
        const CUresult rc = enc(&maps.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(image), gdim, gstride, box, estride,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return FI_ERR_UNSUPPORTED;
    }
    // channels per block: a multiple of 64 (every tile depth divides 64); two blocks per box when C allows, for balance
    const int c_per_block = (C % 128 == 0) ? C / 2 : C;
    dim3 grid(R, C / c_per_block);
    const int smem = 2 * kStageFloats * (int)sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(crop_fwd_nchw_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); return FI_ERR_UNSUPPORTED; }
        attr_set = true;
    }
    crop_fwd_nchw_tma_kernel<<<grid, kTmaThreads, smem, stream>>>(maps, image, boxes, box_ind, dst_row, B, H, W, ph, pw, C, c_per_block, extrap, crops);
    return check_launch("fi_crop_and_resize_forward[nchw tma]");
}
