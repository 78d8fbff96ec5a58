// Caffe-style max RoIPool (the ROIS.METHOD='roi_pool' alternative, lib/roi_pooling/src/roi_pooling_kernel.cu).
// Forward keeps the reference's integer bin arithmetic exactly (roi_pooling_kernel.cu:45-91).  Backward is an
// argmax scatter -- O(outputs) -- instead of the reference's per-input-pixel loop over all RoIs
// (O(B*C*H*W*R), roi_pooling_kernel.cu:147-200), with the reference's own feasibility tests applied to the
// argmax element so the result is the same (see oracle/fi_oracle.c::fi_oracle_roi_pool_bwd).
#include <float.h>

#include "fi_common.cuh"

namespace fi {

struct RoiBins {
    int b, sw, sh, ew, eh;
    float bin_h, bin_w;
};

__device__ __forceinline__ RoiBins roi_bins(const float *roi, float scale, int ph, int pw) {
    RoiBins r;
    r.b = (int)roi[0];
    r.sw = (int)roundf(__fmul_rn(roi[1], scale));
    r.sh = (int)roundf(__fmul_rn(roi[2], scale));
    r.ew = (int)roundf(__fmul_rn(roi[3], scale));
    r.eh = (int)roundf(__fmul_rn(roi[4], scale));
    const int rw = (int)fmaxf((float)(r.ew - r.sw + 1), 1.f), rh = (int)fmaxf((float)(r.eh - r.sh + 1), 1.f);   // malformed -> 1x1
    r.bin_h = __fdiv_rn((float)rh, (float)ph);
    r.bin_w = __fdiv_rn((float)rw, (float)pw);
    return r;
}

__global__ void roi_pool_fwd_kernel(const float *__restrict__ bottom, float scale, long total, int H, int W, int C, int ph, int pw,
                                    const float *__restrict__ rois, float *__restrict__ top, int *__restrict__ argmax) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % pw);
        const int i = (int)((idx / pw) % ph);
        const int c = (int)((idx / pw / ph) % C);
        const int n = (int)(idx / pw / ph / C);
        const RoiBins r = roi_bins(rois + 5 * n, scale, ph, pw);
        int h0 = (int)floorf(__fmul_rn((float)i, r.bin_h)), w0 = (int)floorf(__fmul_rn((float)j, r.bin_w));
        int h1 = (int)ceilf(__fmul_rn((float)(i + 1), r.bin_h)), w1 = (int)ceilf(__fmul_rn((float)(j + 1), r.bin_w));
        h0 = min(max(h0 + r.sh, 0), H); h1 = min(max(h1 + r.sh, 0), H);
        w0 = min(max(w0 + r.sw, 0), W); w1 = min(max(w1 + r.sw, 0), W);
        const bool empty = (h1 <= h0) || (w1 <= w0);
        float best = empty ? 0.f : -FLT_MAX;
        int where = -1;
        const int base = (r.b * C + c) * H * W;
        for (int h = h0; h < h1; ++h)
            for (int w = w0; w < w1; ++w) {
                const float v = __ldg(bottom + base + h * W + w);
                if (v > best) { best = v; where = base + h * W + w; }
            }
        top[idx] = best;
        if (argmax) argmax[idx] = where;
    }
}

__global__ void roi_pool_bwd_kernel(const float *__restrict__ top_diff, const int *__restrict__ argmax, float scale, long total, int H, int W,
                                    int C, int ph, int pw, const float *__restrict__ rois, float *__restrict__ bottom_diff) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int where = argmax[idx];
        if (where < 0) continue;
        const int j = (int)(idx % pw);
        const int i = (int)((idx / pw) % ph);
        const int c = (int)((idx / pw / ph) % C);
        const int n = (int)(idx / pw / ph / C);
        const RoiBins r = roi_bins(rois + 5 * n, scale, ph, pw);
        const int w = where % W, h = (where / W) % H, cc = (where / (W * H)) % C, nn = where / (W * H * C);
        if (nn != r.b || cc != c) continue;                                            // roi_pooling_kernel.cu:153-156
        if (!(w >= r.sw && w <= r.ew && h >= r.sh && h <= r.eh)) continue;             // :164-168
        int p0 = (int)floorf(__fdiv_rn((float)(h - r.sh), r.bin_h)), p1 = (int)ceilf(__fdiv_rn((float)(h - r.sh + 1), r.bin_h));
        int q0 = (int)floorf(__fdiv_rn((float)(w - r.sw), r.bin_w)), q1 = (int)ceilf(__fdiv_rn((float)(w - r.sw + 1), r.bin_w));
        p0 = min(max(p0, 0), ph); p1 = min(max(p1, 0), ph);
        q0 = min(max(q0, 0), pw); q1 = min(max(q1, 0), pw);
        if (i < p0 || i >= p1 || j < q0 || j >= q1) continue;                          // :185-193
        atomicAdd(bottom_diff + where, top_diff[idx]);
    }
}

// ------------------------------------------------------------------------------------------------
// NHWC (torch.channels_last) forward: warp = (RoI, bin, 128-channel slab), lane = 4 channels.  The reference's kernel (and the NCHW
// one above) walks every bin once per CHANNEL with scalar loads H*W*4 bytes apart; here a bin's pixels are read once per slab as
// fully coalesced 512-byte rows, two pixels in flight, and every lane keeps the running max / arg-max of its 4 channels.  Same
// bin arithmetic, same scan order (rows, then columns), same strict `>`: values and arg-max indices (the reference's flat NCHW
// offsets, layout-independent) are identical.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) roi_pool_fwd_nhwc_kernel(const float *__restrict__ bottom, float scale, long units, int H, int W, int C,
                                                               int ph, int pw, const float *__restrict__ rois, float *__restrict__ top,
                                                               int *__restrict__ argmax) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int slabs = C / 128;
    for (long u = warp; u < units; u += nwarps) {
        const int slab = (int)(u % slabs);
        const long q = u / slabs;
        const int j = (int)(q % pw), i = (int)((q / pw) % ph), n = (int)(q / ((long)pw * ph));
        const RoiBins r = roi_bins(rois + 5 * n, scale, ph, pw);
        int h0 = (int)floorf(__fmul_rn((float)i, r.bin_h)), w0 = (int)floorf(__fmul_rn((float)j, r.bin_w));
        int h1 = (int)ceilf(__fmul_rn((float)(i + 1), r.bin_h)), w1 = (int)ceilf(__fmul_rn((float)(j + 1), r.bin_w));
        h0 = min(max(h0 + r.sh, 0), H); h1 = min(max(h1 + r.sh, 0), H);
        w0 = min(max(w0 + r.sw, 0), W); w1 = min(max(w1 + r.sw, 0), W);
        const bool empty = (h1 <= h0) || (w1 <= w0);
        const float init = empty ? 0.f : -FLT_MAX;
        float4 best = make_float4(init, init, init, init);
        int4 where = make_int4(-1, -1, -1, -1);
        const int coff = slab * 128 + lane * 4;
        const float *img = bottom + (long)r.b * H * W * C + coff;
        const int bw = w1 - w0, npx = empty ? 0 : (h1 - h0) * bw;
        auto upd = [&](const float4 v, int pos) {
            if (v.x > best.x) { best.x = v.x; where.x = pos; }
            if (v.y > best.y) { best.y = v.y; where.y = pos; }
            if (v.z > best.z) { best.z = v.z; where.z = pos; }
            if (v.w > best.w) { best.w = v.w; where.w = pos; }
        };
        int p = 0;
        for (; p + 2 <= npx; p += 2) {
            const int ha = h0 + p / bw, wa = w0 + p % bw, hb = h0 + (p + 1) / bw, wb = w0 + (p + 1) % bw;
            const float4 va = __ldg(reinterpret_cast<const float4 *>(img + ((long)ha * W + wa) * C));
            const float4 vb = __ldg(reinterpret_cast<const float4 *>(img + ((long)hb * W + wb) * C));
            upd(va, ha * W + wa);
            upd(vb, hb * W + wb);
        }
        if (p < npx) {
            const int ha = h0 + p / bw, wa = w0 + p % bw;
            upd(__ldg(reinterpret_cast<const float4 *>(img + ((long)ha * W + wa) * C)), ha * W + wa);
        }
        const long o = (((long)n * ph + i) * pw + j) * C + coff;
        *reinterpret_cast<float4 *>(top + o) = best;
        if (argmax) {
            const int plane = H * W, base = (r.b * C + coff) * plane;          // flat NCHW offset of (b, c, h, w), like the reference
            *reinterpret_cast<int4 *>(argmax + o) = make_int4(where.x < 0 ? -1 : base + where.x, where.y < 0 ? -1 : base + plane + where.y,
                                                              where.z < 0 ? -1 : base + 2 * plane + where.z, where.w < 0 ? -1 : base + 3 * plane + where.w);
        }
    }
}

// NHWC backward: thread per output element (channel fastest: coalesced reads of the gradient), arg-max scatter with the reference's
// feasibility tests; the flat NCHW arg-max is re-addressed into the NHWC gradient map.
__global__ void roi_pool_bwd_nhwc_kernel(const float *__restrict__ top_diff, const int *__restrict__ argmax, float scale, long total, int H, int W,
                                         int C, int ph, int pw, const float *__restrict__ rois, float *__restrict__ bottom_diff) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int where = argmax[idx];
        if (where < 0) continue;
        const int c = (int)(idx % C);
        const int j = (int)((idx / C) % pw);
        const int i = (int)((idx / C / pw) % ph);
        const int n = (int)(idx / C / pw / ph);
        const RoiBins r = roi_bins(rois + 5 * n, scale, ph, pw);
        const int w = where % W, h = (where / W) % H, cc = (where / (W * H)) % C, nn = where / (W * H * C);
        if (nn != r.b || cc != c) continue;                                            // roi_pooling_kernel.cu:153-156
        if (!(w >= r.sw && w <= r.ew && h >= r.sh && h <= r.eh)) continue;             // :164-168
        int p0 = (int)floorf(__fdiv_rn((float)(h - r.sh), r.bin_h)), p1 = (int)ceilf(__fdiv_rn((float)(h - r.sh + 1), r.bin_h));
        int q0 = (int)floorf(__fdiv_rn((float)(w - r.sw), r.bin_w)), q1 = (int)ceilf(__fdiv_rn((float)(w - r.sw + 1), r.bin_w));
        p0 = min(max(p0, 0), ph); p1 = min(max(p1, 0), ph);
        q0 = min(max(q0, 0), pw); q1 = min(max(q1, 0), pw);
        if (i < p0 || i >= p1 || j < q0 || j >= q1) continue;                          // :185-193
        atomicAdd(bottom_diff + (((long)nn * H + h) * W + w) * C + cc, top_diff[idx]);
    }
}

static int launch_grid(long total) {
    long g = (total + 255) / 256;
    if (g > kNumSMs * 8) g = kNumSMs * 8;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace fi

using namespace fi;

FI_API int fi_roi_pool_forward(const float *bottom, float spatial_scale, int batch, int num_rois, int height, int width, int channels,
                               int pooled_h, int pooled_w, const float *rois, float *top, int *argmax, cudaStream_t stream) {
    FI_REQUIRE(batch > 0 && num_rois >= 0 && height > 0 && width > 0 && channels > 0 && pooled_h > 0 && pooled_w > 0, "fi_roi_pool_forward: bad sizes");
    FI_REQUIRE((long)batch * channels * height * width < 2147483647L, "fi_roi_pool_forward: argmax is int32 (like the reference); feature map too large");
    if (num_rois == 0) return ok();
    FI_REQUIRE(bottom && rois && top, "fi_roi_pool_forward: null pointer");
    const long total = (long)num_rois * channels * pooled_h * pooled_w;
    roi_pool_fwd_kernel<<<launch_grid(total), 256, 0, stream>>>(bottom, spatial_scale, total, height, width, channels, pooled_h, pooled_w, rois, top, argmax);
    return check_launch("fi_roi_pool_forward");
}

FI_API int fi_roi_pool_backward(const float *top_diff, float spatial_scale, int batch, int num_rois, int height, int width, int channels,
                                int pooled_h, int pooled_w, const float *rois, float *bottom_diff, const int *argmax, cudaStream_t stream) {
    FI_REQUIRE(batch > 0 && num_rois >= 0 && height > 0 && width > 0 && channels > 0 && pooled_h > 0 && pooled_w > 0 && bottom_diff, "fi_roi_pool_backward: bad arguments");
    cudaError_t e = cudaMemsetAsync(bottom_diff, 0, sizeof(float) * (size_t)batch * channels * height * width, stream);
    if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_roi_pool_backward: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    if (num_rois == 0) return ok();
    FI_REQUIRE(top_diff && rois && argmax, "fi_roi_pool_backward: null pointer");
    const long total = (long)num_rois * channels * pooled_h * pooled_w;
    roi_pool_bwd_kernel<<<launch_grid(total), 256, 0, stream>>>(top_diff, argmax, spatial_scale, total, height, width, channels, pooled_h, pooled_w, rois, bottom_diff);
    return check_launch("fi_roi_pool_backward");
}

// ---- reference-named launchers (lib/roi_pooling/src/roi_pooling_kernel.h:8-18): 1 = success ----------
FI_API int ROIPoolForwardLaucher(const float *bottom_data, const float spatial_scale, const int num_rois, const int height, const int width,
                                 const int channels, const int pooled_height, const int pooled_width, const float *bottom_rois,
                                 float *top_data, int *argmax_data, cudaStream_t stream) {
    // the reference launcher is not told the batch size; it is only used for the int32 range check
    return fi_roi_pool_forward(bottom_data, spatial_scale, 1, num_rois, height, width, channels, pooled_height, pooled_width, bottom_rois,
                               top_data, argmax_data, stream) == FI_OK ? 1 : 0;
}

FI_API int ROIPoolBackwardLaucher(const float *top_diff, const float spatial_scale, const int batch_size, const int num_rois, const int height,
                                  const int width, const int channels, const int pooled_height, const int pooled_width,
                                  const float *bottom_rois, float *bottom_diff, const int *argmax_data, cudaStream_t stream) {
    return fi_roi_pool_backward(top_diff, spatial_scale, batch_size, num_rois, height, width, channels, pooled_height, pooled_width,
                                bottom_rois, bottom_diff, argmax_data, stream) == FI_OK ? 1 : 0;
}

FI_API int fi_roi_pool_forward_nhwc(const float *bottom, float spatial_scale, int batch, int num_rois, int height, int width, int channels,
                                    int pooled_h, int pooled_w, const float *rois, float *top, int *argmax, cudaStream_t stream) {
    FI_REQUIRE(batch > 0 && num_rois >= 0 && height > 0 && width > 0 && channels > 0 && pooled_h > 0 && pooled_w > 0, "fi_roi_pool_forward_nhwc: bad sizes");
    FI_REQUIRE((long)batch * channels * height * width < 2147483647L, "fi_roi_pool_forward_nhwc: argmax is int32 (like the reference); feature map too large");
    if (channels % 128 != 0 || ((uintptr_t)bottom % 16) || ((uintptr_t)top % 16) || ((uintptr_t)argmax % 16)) {
        set_error(FI_ERR_UNSUPPORTED, "fi_roi_pool_forward_nhwc: needs channels %% 128 == 0 and 16-byte aligned tensors");
        return FI_ERR_UNSUPPORTED;
    }
    if (num_rois == 0) return ok();
    FI_REQUIRE(bottom && rois && top, "fi_roi_pool_forward_nhwc: null pointer");
    const long units = (long)num_rois * pooled_h * pooled_w * (channels / 128);
    long grid = (units + 7) / 8;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    roi_pool_fwd_nhwc_kernel<<<(int)grid, 256, 0, stream>>>(bottom, spatial_scale, units, height, width, channels, pooled_h, pooled_w, rois, top, argmax);
    return check_launch("fi_roi_pool_forward_nhwc");
}

FI_API int fi_roi_pool_backward_nhwc(const float *top_diff, float spatial_scale, int batch, int num_rois, int height, int width, int channels,
                                     int pooled_h, int pooled_w, const float *rois, float *bottom_diff, const int *argmax, cudaStream_t stream) {
    FI_REQUIRE(batch > 0 && num_rois >= 0 && height > 0 && width > 0 && channels > 0 && pooled_h > 0 && pooled_w > 0 && bottom_diff, "fi_roi_pool_backward_nhwc: bad arguments");
    cudaError_t e = cudaMemsetAsync(bottom_diff, 0, sizeof(float) * (size_t)batch * channels * height * width, stream);
    if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_roi_pool_backward_nhwc: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    if (num_rois == 0) return ok();
    FI_REQUIRE(top_diff && rois && argmax, "fi_roi_pool_backward_nhwc: null pointer");
    const long total = (long)num_rois * channels * pooled_h * pooled_w;
    roi_pool_bwd_nhwc_kernel<<<launch_grid(total), 256, 0, stream>>>(top_diff, argmax, spatial_scale, total, height, width, channels, pooled_h, pooled_w, rois, bottom_diff);
    return check_launch("fi_roi_pool_backward_nhwc");
}
