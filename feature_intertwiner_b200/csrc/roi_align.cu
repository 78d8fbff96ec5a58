// RoIAlign (TF crop_and_resize, one bilinear sample per bin) for B200 / sm_100a.
//
// Replaces lib/roi_align/src/cuda/crop_and_resize_kernel.cu (thread-per-output NCHW gather, 4 scalar
// uncoalesced loads per element, 4 scalar atomics per element in backward, full memsets of both outputs).
//
// Layouts.  The op is a gather of C-vectors: each sample reads 4 feature pixels and every channel of a
// pixel shares the same coordinates.  In NCHW those C values are H*W*4 bytes apart; in NHWC
// (torch.channels_last -- what cuDNN produces natively on this part) they are ONE contiguous C*4-byte run
// (1 KB for FPN's C=256), so a tap is a fully coalesced 128-bit-per-lane read and a backward tap is a
// coalesced vector reduction.  NHWC is therefore the native layout here; NCHW is kept for drop-in parity
// with callers that hand over contiguous NCHW tensors.
//
// Coordinate math is shared with fi_common.cuh::axis_sample (bit-exact tap indices vs the reference).
#include "fi_common.cuh"

namespace fi {

// ------------------------------------------------------------------------------------------------
// Tap table (test / verification entry point).
// ------------------------------------------------------------------------------------------------
__global__ void crop_taps_kernel(const float *__restrict__ boxes, int R, int H, int W, int ph, int pw, int *__restrict__ taps) {
    const long total = (long)R * ph * pw;
    for (long s = blockIdx.x * (long)blockDim.x + threadIdx.x; s < total; s += (long)gridDim.x * blockDim.x) {
        const int r = (int)(s / (ph * pw));
        const int rem = (int)(s - (long)r * ph * pw);
        const int i = rem / pw, j = rem - i * pw;
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        const AxisTap tx = axis_sample(x1, x2, axis_step(x1, x2, W, pw), j, W, pw);
        int *t = taps + s * 5;
        t[0] = ty.lo; t[1] = ty.hi; t[2] = tx.lo; t[3] = tx.hi; t[4] = (ty.inside && tx.inside) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// NHWC forward.  `lanes` threads cooperate on one sample, each moving VEC floats per channel step;
// a block works on blockDim.x/lanes samples at a time.  Per float4 of output: 4 coalesced 128-bit
// read-only loads (L1/L2 absorb the tap overlap between neighbouring samples and boxes) and one
// streaming 128-bit store.  No shared memory: there is nothing to transpose in this layout.
// ------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) crop_fwd_nhwc_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                                                           const int *__restrict__ box_ind, const int *__restrict__ dst_row,
                                                           long nsamples, int B, int H, int W, int ph, int pw, int C, int lanes,
                                                           float extrap, float *__restrict__ crops) {
    const int spb = blockDim.x / lanes;           // samples in flight per block
    const int ls = threadIdx.x / lanes;
    const int lc = threadIdx.x - ls * lanes;
    if (ls >= spb) return;
    const int pp = ph * pw;
    const int CV = C / VEC;
    for (long s = (long)blockIdx.x * spb + ls; s < nsamples; s += (long)gridDim.x * spb) {
        const int r = (int)(s / pp);
        const int rem = (int)(s - (long)r * pp);
        const int i = rem / pw, j = rem - i * pw;
        const int b = box_ind[r];
        const long orow = dst_row ? (long)dst_row[r] : (long)r;
        float *out = crops + (orow * pp + rem) * (long)C;
        if (b < 0 || b >= B) {   // reference leaves such rows at their zero fill (crop_and_resize_kernel.cu:34-38)
            for (int cv = lc; cv < CV; cv += lanes) {
                if (VEC == 4) st_stream4(out + cv * 4, make_float4(0.f, 0.f, 0.f, 0.f));
                else out[cv] = 0.f;
            }
            continue;
        }
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        const AxisTap tx = axis_sample(x1, x2, axis_step(x1, x2, W, pw), j, W, pw);
        if (!(ty.inside && tx.inside)) {
            for (int cv = lc; cv < CV; cv += lanes) {
                if (VEC == 4) st_stream4(out + cv * 4, make_float4(extrap, extrap, extrap, extrap));
                else out[cv] = extrap;
            }
            continue;
        }
        const float *img = image + (long)b * H * W * C;
        const float *ptl = img + ((long)ty.lo * W + tx.lo) * C;
        const float *ptr = img + ((long)ty.lo * W + tx.hi) * C;
        const float *pbl = img + ((long)ty.hi * W + tx.lo) * C;
        const float *pbr = img + ((long)ty.hi * W + tx.hi) * C;
        for (int cv = lc; cv < CV; cv += lanes) {
            if (VEC == 4) {
                const float4 tl = ldg4(ptl + cv * 4), tr = ldg4(ptr + cv * 4);
                const float4 bl = ldg4(pbl + cv * 4), br = ldg4(pbr + cv * 4);
                const float4 top = lerp_rn(tl, tr, tx.frac);
                const float4 bot = lerp_rn(bl, br, tx.frac);
                st_stream4(out + cv * 4, lerp_rn(top, bot, ty.frac));
            } else {
                const float top = lerp_rn(__ldg(ptl + cv), __ldg(ptr + cv), tx.frac);
                const float bot = lerp_rn(__ldg(pbl + cv), __ldg(pbr + cv), tx.frac);
                out[cv] = lerp_rn(top, bot, ty.frac);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NHWC backward (scatter form).  One thread owns VEC channels of one crop ROW (i fixed) and walks the
// row's samples j = 0..pw-1 in order, keeping the contributions to the current (x_lo, x_hi) pixel pair
// of the top and of the bottom image row in registers.  They are flushed with vector reductions
// (red.global.add.v4.f32, coalesced across the C-lanes of the pixel) only when the pixel pair changes.
// Upsampling crops (step < 1 px: every small box pooled at 14x14) and degenerate / zero-padded RoIs (step 0:
// all samples on one pixel, which in the reference serialise hundreds of atomics on one address) thus
// issue one reduction per DISTINCT pixel instead of one per tap.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(float *p, float4 v) {
    // sm_90+ 128-bit vector reduction; result unused -> RED, not ATOM
    atomicAdd(reinterpret_cast<float4 *>(p), v);
}
__device__ __forceinline__ void red_add(float *p, float v) { atomicAdd(p, v); }

__device__ __forceinline__ float4 f4_scale(float4 a, float w) {
    return make_float4(__fmul_rn(a.x, w), __fmul_rn(a.y, w), __fmul_rn(a.z, w), __fmul_rn(a.w, w));
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float f4_scale(float a, float w) { return __fmul_rn(a, w); }
__device__ __forceinline__ float f4_add(float a, float b) { return a + b; }
template <typename T> __device__ __forceinline__ T vzero();
template <> __device__ __forceinline__ float4 vzero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float vzero<float>() { return 0.f; }
__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }

template <typename VT>
__global__ void __launch_bounds__(256) crop_bwd_nhwc_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                                                           const int *__restrict__ box_ind, const int *__restrict__ src_row,
                                                           long nrows, int B, int H, int W, int ph, int pw, int C, int lanes,
                                                           float *__restrict__ gimg) {
    constexpr int VEC = sizeof(VT) / sizeof(float);
    const int rpb = blockDim.x / lanes;           // crop rows in flight per block
    const int lr = threadIdx.x / lanes;
    const int lc = threadIdx.x - lr * lanes;
    if (lr >= rpb) return;
    const int CV = C / VEC;
    for (long q = (long)blockIdx.x * rpb + lr; q < nrows; q += (long)gridDim.x * rpb) {
        const int r = (int)(q / ph);
        const int i = (int)(q - (long)r * ph);
        const int b = box_ind[r];
        if (b < 0 || b >= B) continue;
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        if (!ty.inside) continue;
        const float sx = axis_step(x1, x2, W, pw);
        const float wy_hi = ty.frac, wy_lo = __fsub_rn(1.f, ty.frac);     // crop_and_resize.c:241,245
        const long grow = src_row ? (long)src_row[r] : (long)r;
        const float *g = grads + ((grow * ph + i) * (long)pw) * C;
        float *img_top = gimg + ((long)b * H + ty.lo) * (long)W * C;
        float *img_bot = gimg + ((long)b * H + ty.hi) * (long)W * C;
        for (int cv = lc; cv < CV; cv += lanes) {
            VT t_lo = vzero<VT>(), t_hi = vzero<VT>(), b_lo = vzero<VT>(), b_hi = vzero<VT>();
            int cur_lo = -1, cur_hi = -1;
            for (int j = 0; j < pw; ++j) {
                const AxisTap tx = axis_sample(x1, x2, sx, j, W, pw);
                if (!tx.inside) continue;
                if (tx.lo != cur_lo || tx.hi != cur_hi) {
                    if (cur_lo >= 0) {
                        red_add(img_top + (long)cur_lo * C + cv * VEC, t_lo);
                        red_add(img_bot + (long)cur_lo * C + cv * VEC, b_lo);
                        if (tx.lo == cur_hi && tx.hi != cur_hi) {
                            // window slides by one pixel: the old `hi` column becomes the new `lo`
                            t_lo = t_hi; b_lo = b_hi;
                        } else {
                            red_add(img_top + (long)cur_hi * C + cv * VEC, t_hi);
                            red_add(img_bot + (long)cur_hi * C + cv * VEC, b_hi);
                            t_lo = vzero<VT>(); b_lo = vzero<VT>();
                        }
                        t_hi = vzero<VT>(); b_hi = vzero<VT>();
                    }
                    cur_lo = tx.lo; cur_hi = tx.hi;
                }
                const VT gv = ld_stream(reinterpret_cast<const VT *>(g + (long)j * C) + cv);
                const VT dtop = f4_scale(gv, wy_lo), dbot = f4_scale(gv, wy_hi);
                const float wx_hi = tx.frac, wx_lo = __fsub_rn(1.f, tx.frac);
                t_lo = f4_add(t_lo, f4_scale(dtop, wx_lo)); t_hi = f4_add(t_hi, f4_scale(dtop, wx_hi));
                b_lo = f4_add(b_lo, f4_scale(dbot, wx_lo)); b_hi = f4_add(b_hi, f4_scale(dbot, wx_hi));
            }
            if (cur_lo >= 0) {
                red_add(img_top + (long)cur_lo * C + cv * VEC, t_lo);
                red_add(img_bot + (long)cur_lo * C + cv * VEC, b_lo);
                red_add(img_top + (long)cur_hi * C + cv * VEC, t_hi);
                red_add(img_bot + (long)cur_hi * C + cv * VEC, b_hi);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NCHW forward / backward (drop-in layout).  One block per (box, channel chunk): the P*P tap table of
// the box is computed once into shared memory, then threads sweep (channel, i, j) with j fastest so the
// crop is written fully coalesced.  Reads are 4 L1-cached scalar loads per output.
// ------------------------------------------------------------------------------------------------
constexpr int kNchwMaxTaps = 64;       // per axis, held in shared memory
constexpr int kNchwChunk = 32;         // channels per block

struct SmemTaps {
    int ylo[kNchwMaxTaps], yhi[kNchwMaxTaps], xlo[kNchwMaxTaps], xhi[kNchwMaxTaps];
    float yf[kNchwMaxTaps], xf[kNchwMaxTaps];
    unsigned char yin[kNchwMaxTaps], xin[kNchwMaxTaps];
};

__device__ __forceinline__ void fill_taps(SmemTaps &t, const float *boxes, int r, int H, int W, int ph, int pw) {
    const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
    const float sy = axis_step(y1, y2, H, ph), sx = axis_step(x1, x2, W, pw);
    for (int k = threadIdx.x; k < ph; k += blockDim.x) {
        const AxisTap a = axis_sample(y1, y2, sy, k, H, ph);
        t.ylo[k] = a.lo; t.yhi[k] = a.hi; t.yf[k] = a.frac; t.yin[k] = a.inside;
    }
    for (int k = threadIdx.x; k < pw; k += blockDim.x) {
        const AxisTap a = axis_sample(x1, x2, sx, k, W, pw);
        t.xlo[k] = a.lo; t.xhi[k] = a.hi; t.xf[k] = a.frac; t.xin[k] = a.inside;
    }
}

__global__ void __launch_bounds__(256) crop_fwd_nchw_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                                                           const int *__restrict__ box_ind, const int *__restrict__ dst_row,
                                                           int B, int H, int W, int ph, int pw, int C, float extrap,
                                                           float *__restrict__ crops) {
    __shared__ SmemTaps t;
    const int r = blockIdx.x;
    const int c0 = blockIdx.y * kNchwChunk;
    const int nc = min(kNchwChunk, C - c0);
    const int pp = ph * pw;
    const int b = box_ind[r];
    const long orow = dst_row ? (long)dst_row[r] : (long)r;
    float *out = crops + (orow * C + c0) * (long)pp;
    if (b < 0 || b >= B) {
        for (int e = threadIdx.x; e < nc * pp; e += blockDim.x) out[e] = 0.f;
        return;
    }
    fill_taps(t, boxes, r, H, W, ph, pw);
    __syncthreads();
    const float *img = image + ((long)b * C + c0) * H * W;
    for (int e = threadIdx.x; e < nc * pp; e += blockDim.x) {
        const int c = e / pp;
        const int s = e - c * pp;
        const int i = s / pw, j = s - i * pw;
        float v = extrap;
        if (t.yin[i] && t.xin[j]) {
            const float *p = img + (long)c * H * W;
            const float *rt = p + (long)t.ylo[i] * W, *rb = p + (long)t.yhi[i] * W;
            const float top = lerp_rn(__ldg(rt + t.xlo[j]), __ldg(rt + t.xhi[j]), t.xf[j]);
            const float bot = lerp_rn(__ldg(rb + t.xlo[j]), __ldg(rb + t.xhi[j]), t.xf[j]);
            v = lerp_rn(top, bot, t.yf[i]);
        }
        __stcs(out + e, v);
    }
}

__global__ void __launch_bounds__(256) crop_bwd_nchw_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                                                           const int *__restrict__ box_ind, const int *__restrict__ src_row,
                                                           int B, int H, int W, int ph, int pw, int C, float *__restrict__ gimg) {
    __shared__ SmemTaps t;
    const int r = blockIdx.x;
    const int c0 = blockIdx.y * kNchwChunk;
    const int nc = min(kNchwChunk, C - c0);
    const int pp = ph * pw;
    const int b = box_ind[r];
    if (b < 0 || b >= B) return;
    fill_taps(t, boxes, r, H, W, ph, pw);
    __syncthreads();
    const long grow = src_row ? (long)src_row[r] : (long)r;
    const float *g = grads + (grow * C + c0) * (long)pp;
    float *img = gimg + ((long)b * C + c0) * H * W;
    for (int e = threadIdx.x; e < nc * pp; e += blockDim.x) {
        const int c = e / pp;
        const int s = e - c * pp;
        const int i = s / pw, j = s - i * pw;
        if (!(t.yin[i] && t.xin[j])) continue;
        const float gv = __ldcs(g + e);
        float *p = img + (long)c * H * W;
        float *rt = p + (long)t.ylo[i] * W, *rb = p + (long)t.yhi[i] * W;
        const float dtop = __fmul_rn(__fsub_rn(1.f, t.yf[i]), gv), dbot = __fmul_rn(t.yf[i], gv);
        const float wl = __fsub_rn(1.f, t.xf[j]), wh = t.xf[j];
        atomicAdd(rt + t.xlo[j], __fmul_rn(wl, dtop));
        atomicAdd(rt + t.xhi[j], __fmul_rn(wh, dtop));
        atomicAdd(rb + t.xlo[j], __fmul_rn(wl, dbot));
        atomicAdd(rb + t.xhi[j], __fmul_rn(wh, dbot));
    }
}

// Generic fallback for crops wider than the shared tap table: thread per output element.
__global__ void crop_fwd_nchw_generic_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                                             const int *__restrict__ box_ind, const int *__restrict__ dst_row, long total,
                                             int B, int H, int W, int ph, int pw, int C, float extrap, float *__restrict__ crops) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long idx = e;
        const int j = (int)(idx % pw); idx /= pw;
        const int i = (int)(idx % ph); idx /= ph;
        const int c = (int)(idx % C);
        const int r = (int)(idx / C);
        const long orow = dst_row ? (long)dst_row[r] : (long)r;
        float *o = crops + ((orow * C + c) * ph + i) * (long)pw + j;
        const int b = box_ind[r];
        if (b < 0 || b >= B) { *o = 0.f; continue; }
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        const AxisTap tx = axis_sample(x1, x2, axis_step(x1, x2, W, pw), j, W, pw);
        float v = extrap;
        if (ty.inside && tx.inside) {
            const float *p = image + ((long)b * C + c) * H * W;
            const float top = lerp_rn(__ldg(p + (long)ty.lo * W + tx.lo), __ldg(p + (long)ty.lo * W + tx.hi), tx.frac);
            const float bot = lerp_rn(__ldg(p + (long)ty.hi * W + tx.lo), __ldg(p + (long)ty.hi * W + tx.hi), tx.frac);
            v = lerp_rn(top, bot, ty.frac);
        }
        *o = v;
    }
}

__global__ void crop_bwd_nchw_generic_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                                             const int *__restrict__ box_ind, const int *__restrict__ src_row, long total,
                                             int B, int H, int W, int ph, int pw, int C, float *__restrict__ gimg) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long idx = e;
        const int j = (int)(idx % pw); idx /= pw;
        const int i = (int)(idx % ph); idx /= ph;
        const int c = (int)(idx % C);
        const int r = (int)(idx / C);
        const int b = box_ind[r];
        if (b < 0 || b >= B) continue;
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        const AxisTap tx = axis_sample(x1, x2, axis_step(x1, x2, W, pw), j, W, pw);
        if (!(ty.inside && tx.inside)) continue;
        const long grow = src_row ? (long)src_row[r] : (long)r;
        const float gv = grads[((grow * C + c) * ph + i) * (long)pw + j];
        float *p = gimg + ((long)b * C + c) * H * W;
        const float dtop = __fmul_rn(__fsub_rn(1.f, ty.frac), gv), dbot = __fmul_rn(ty.frac, gv);
        const float wl = __fsub_rn(1.f, tx.frac), wh = tx.frac;
        atomicAdd(p + (long)ty.lo * W + tx.lo, __fmul_rn(wl, dtop));
        atomicAdd(p + (long)ty.lo * W + tx.hi, __fmul_rn(wh, dtop));
        atomicAdd(p + (long)ty.hi * W + tx.lo, __fmul_rn(wl, dbot));
        atomicAdd(p + (long)ty.hi * W + tx.hi, __fmul_rn(wh, dbot));
    }
}

// lanes cooperating on one sample/row: the largest power of two <= min(CV, 256) that divides 256
static int pick_lanes(int CV) {
    int l = 1;
    while (l * 2 <= CV && l * 2 <= 256) l *= 2;
    return l;
}

static int grid_for(long work_items, int per_block, int blocks_per_sm) {
    long need = (work_items + per_block - 1) / per_block;
    long cap = (long)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static int check_common(const void *a, const void *boxes, const void *box_ind, const void *out, int R, int B, int H, int W,
                        int ph, int pw, int C) {
    FI_REQUIRE(R >= 0 && B > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && C > 0, "crop_and_resize: bad sizes R=%d B=%d H=%d W=%d ph=%d pw=%d C=%d", R, B, H, W, ph, pw, C);
    FI_REQUIRE(R == 0 || (a && boxes && box_ind && out), "crop_and_resize: null pointer");
    return FI_OK;
}

}  // namespace fi

using namespace fi;

FI_API int fi_crop_taps(const float *boxes, int num_boxes, int H, int W, int ph, int pw, int *taps, cudaStream_t stream) {
    FI_REQUIRE(num_boxes >= 0 && H > 0 && W > 0 && ph > 0 && pw > 0, "fi_crop_taps: bad sizes");
    if (num_boxes == 0) return ok();
    FI_REQUIRE(boxes && taps, "fi_crop_taps: null pointer");
    const long total = (long)num_boxes * ph * pw;
    crop_taps_kernel<<<grid_for(total, 256, 8), 256, 0, stream>>>(boxes, num_boxes, H, W, ph, pw, taps);
    return check_launch("fi_crop_taps");
}

FI_API int fi_crop_and_resize_forward(const float *image, int image_layout, const float *boxes, const int *box_ind,
                                      const int *dst_row, int R, int B, int H, int W, int ph, int pw, int C, float extrap,
                                      float *crops, int crops_layout, cudaStream_t stream) {
    if (int e = check_common(image, boxes, box_ind, crops, R, B, H, W, ph, pw, C)) return e;
    if (R == 0) return ok();
    if (image_layout != crops_layout) {
        set_error(FI_ERR_UNSUPPORTED, "fi_crop_and_resize_forward: mixed layouts (image %d, crops %d)", image_layout, crops_layout);
        return FI_ERR_UNSUPPORTED;
    }
    if (image_layout == FI_LAYOUT_NHWC) {
        const long nsamples = (long)R * ph * pw;
        const bool vec = (C % 4 == 0) && ((uintptr_t)image % 16 == 0) && ((uintptr_t)crops % 16 == 0);
        if (vec) {
            const int lanes = pick_lanes(C / 4);
            crop_fwd_nhwc_kernel<4><<<grid_for(nsamples, 256 / lanes, 8), 256, 0, stream>>>(
                image, boxes, box_ind, dst_row, nsamples, B, H, W, ph, pw, C, lanes, extrap, crops);
        } else {
            const int lanes = pick_lanes(C);
            crop_fwd_nhwc_kernel<1><<<grid_for(nsamples, 256 / lanes, 8), 256, 0, stream>>>(
                image, boxes, box_ind, dst_row, nsamples, B, H, W, ph, pw, C, lanes, extrap, crops);
        }
        return check_launch("fi_crop_and_resize_forward[nhwc]");
    }
    if (image_layout == FI_LAYOUT_NCHW) {
        if (ph <= kNchwMaxTaps && pw <= kNchwMaxTaps) {
            dim3 grid(R, ceil_div(C, kNchwChunk));
            crop_fwd_nchw_kernel<<<grid, 256, 0, stream>>>(image, boxes, box_ind, dst_row, B, H, W, ph, pw, C, extrap, crops);
        } else {
            const long total = (long)R * C * ph * pw;
            crop_fwd_nchw_generic_kernel<<<grid_for(total, 256, 8), 256, 0, stream>>>(image, boxes, box_ind, dst_row, total, B, H, W,
                                                                                     ph, pw, C, extrap, crops);
        }
        return check_launch("fi_crop_and_resize_forward[nchw]");
    }
    set_error(FI_ERR_INVALID, "fi_crop_and_resize_forward: unknown layout %d", image_layout);
    return FI_ERR_INVALID;
}

FI_API int fi_crop_and_resize_backward(const float *grads, int grads_layout, const float *boxes, const int *box_ind,
                                       const int *src_row, int R, int B, int H, int W, int ph, int pw, int C,
                                       float *gimg, int image_layout, int accumulate, cudaStream_t stream) {
    FI_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && gimg, "fi_crop_and_resize_backward: bad image");
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(gimg, 0, sizeof(float) * (size_t)B * C * H * W, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_crop_and_resize_backward: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    }
    if (int e = check_common(grads, boxes, box_ind, gimg, R, B, H, W, ph, pw, C)) return e;
    if (R == 0) return ok();
    if (image_layout != grads_layout) {
        set_error(FI_ERR_UNSUPPORTED, "fi_crop_and_resize_backward: mixed layouts (grads %d, image %d)", grads_layout, image_layout);
        return FI_ERR_UNSUPPORTED;
    }
    if (image_layout == FI_LAYOUT_NHWC) {
        const long nrows = (long)R * ph;
        const bool vec = (C % 4 == 0) && ((uintptr_t)gimg % 16 == 0) && ((uintptr_t)grads % 16 == 0);
        if (vec) {
            const int lanes = pick_lanes(C / 4);
            crop_bwd_nhwc_kernel<float4><<<grid_for(nrows, 256 / lanes, 8), 256, 0, stream>>>(
                grads, boxes, box_ind, src_row, nrows, B, H, W, ph, pw, C, lanes, gimg);
        } else {
            const int lanes = pick_lanes(C);
            crop_bwd_nhwc_kernel<float><<<grid_for(nrows, 256 / lanes, 8), 256, 0, stream>>>(
                grads, boxes, box_ind, src_row, nrows, B, H, W, ph, pw, C, lanes, gimg);
        }
        return check_launch("fi_crop_and_resize_backward[nhwc]");
    }
    if (image_layout == FI_LAYOUT_NCHW) {
        if (ph <= kNchwMaxTaps && pw <= kNchwMaxTaps) {
            dim3 grid(R, ceil_div(C, kNchwChunk));
            crop_bwd_nchw_kernel<<<grid, 256, 0, stream>>>(grads, boxes, box_ind, src_row, B, H, W, ph, pw, C, gimg);
        } else {
            const long total = (long)R * C * ph * pw;
            crop_bwd_nchw_generic_kernel<<<grid_for(total, 256, 8), 256, 0, stream>>>(grads, boxes, box_ind, src_row, total, B, H, W,
                                                                                     ph, pw, C, gimg);
        }
        return check_launch("fi_crop_and_resize_backward[nchw]");
    }
    set_error(FI_ERR_INVALID, "fi_crop_and_resize_backward: unknown layout %d", image_layout);
    return FI_ERR_INVALID;
}

// ---- reference-named launchers (lib/roi_align/src/cuda/crop_and_resize_kernel.h:8-18) ------------
FI_API void CropAndResizeLaucher(const float *image_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes, int batch,
                                 int image_height, int image_width, int crop_height, int crop_width, int depth,
                                 float extrapolation_value, float *crops_ptr, cudaStream_t stream) {
    fi_crop_and_resize_forward(image_ptr, FI_LAYOUT_NCHW, boxes_ptr, box_ind_ptr, nullptr, num_boxes, batch, image_height,
                               image_width, crop_height, crop_width, depth, extrapolation_value, crops_ptr, FI_LAYOUT_NCHW, stream);
}

FI_API void CropAndResizeBackpropImageLaucher(const float *grads_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes,
                                              int batch, int image_height, int image_width, int crop_height, int crop_width,
                                              int depth, float *grads_image_ptr, cudaStream_t stream) {
    fi_crop_and_resize_backward(grads_ptr, FI_LAYOUT_NCHW, boxes_ptr, box_ind_ptr, nullptr, num_boxes, batch, image_height,
                                image_width, crop_height, crop_width, depth, grads_image_ptr, FI_LAYOUT_NCHW, /*accumulate=*/1, stream);
}
