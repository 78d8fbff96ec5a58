// Per-unit device code of the NHWC RoIAlign kernels of roi_align.cu (plain and level-batched launches).  One unit = (box r, crop row i, 128-channel slab) = one warp.
#pragma once
#include "fi_common.cuh"

namespace fi {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ AxisTap shfl_tap(const AxisTap &t, int src) {
    AxisTap o;
    o.lo = __shfl_sync(0xffffffffu, t.lo, src);
    o.hi = __shfl_sync(0xffffffffu, t.hi, src);
    o.frac = __shfl_sync(0xffffffffu, t.frac, src);
    o.inside = __shfl_sync(0xffffffffu, (int)t.inside, src) != 0;
    return o;
}

// Forward.  Unit of work = (box r, crop row i, 128-channel slab): one warp, one float4 per lane per tap.
// The row is walked in batches of U samples: all 4*U tap loads of a batch are issued back to back
// (branch-free: outside samples read a clamped pixel and are overwritten with the extrapolation value
// afterwards), then the batch is interpolated and stored.  ncu on the first, branchy version of this kernel
// showed 16 % issue-active with every warp parked on long_scoreboard: the bound was memory-level
// parallelism, not bytes, so the loop is written to keep 4*U*512 B in flight per warp.
// Overlap between neighbouring samples / rows / boxes is left to L1 and L2.
struct FwdSet {                  // one forward crop set (device view)
    const float *image, *boxes;
    const int *box_ind, *dst_row, *R_dev;     // R_dev: NULL, or the actual number of boxes on the device (units past it exit)
    float *crops, *crops2;
    int B, H, W, C, ph, pw, slabs, R;         // R: number of boxes (capacity of the lists when R_dev is given)
    float extrap;
};

template <int U>
__device__ __forceinline__ void fwd_unit(const FwdSet &S, long u, int lane) {
    const float *__restrict__ image = S.image;
    const float *__restrict__ boxes = S.boxes;
    const int *__restrict__ box_ind = S.box_ind;
    const int *__restrict__ dst_row = S.dst_row;
    float *__restrict__ crops = S.crops;
    float *__restrict__ crops2 = S.crops2;
    const int slabs = S.slabs, B = S.B, H = S.H, W = S.W, ph = S.ph, pw = S.pw, C = S.C;
    const float extrap = S.extrap;
    const int slab = (int)(u % slabs);
    const long q = u / slabs;
    const int r = (int)(q / ph), i = (int)(q - (long)r * ph);
    if (S.R_dev && r >= *S.R_dev) return;
    const int b = box_ind[r];
    const long orow = dst_row ? (long)dst_row[r] : (long)r;
    const int coff = slab * 128 + lane * 4;
    float *out = crops + ((orow * ph + i) * (long)pw) * C + coff;
    float *out2 = crops2 ? crops2 + (((long)r * ph + i) * (long)pw) * C + coff : nullptr;   // optional compact copy (row r)
    const bool bad = (b < 0 || b >= B);     // reference leaves such rows at their zero fill (crop_and_resize_kernel.cu:34-38)
    const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
    const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
    if (bad || !ty.inside) {
        const float v = bad ? 0.f : extrap;
        const float4 v4 = make_float4(v, v, v, v);
        for (int j = 0; j < pw; ++j) {
            st_stream4(out + (long)j * C, v4);
            if (out2) st_stream4(out2 + (long)j * C, v4);
        }
        return;
    }
    const float sx = axis_step(x1, x2, W, pw);
    const float *rowT = image + ((long)b * H + ty.lo) * (long)W * C + coff;
    const float *rowB = image + ((long)b * H + ty.hi) * (long)W * C + coff;
    for (int jb = 0; jb < pw; jb += 32) {                        // lane l computes the x tap of sample jb + l
        const AxisTap mine = axis_sample(x1, x2, sx, jb + lane, W, pw);
        const int lo_c = min(max(mine.lo, 0), W - 1), hi_c = min(max(mine.hi, 0), W - 1);
        const int packed = lo_c | (hi_c << 15) | (mine.inside ? (1 << 30) : 0);      // W <= 32768
        const int jn = min(32, pw - jb);
        for (int j0 = 0; j0 < jn; j0 += U) {
            float4 tl[U], tr[U], bl[U], br[U];
            int pk[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int j = min(j0 + k, jn - 1);               // tail lanes re-read the last sample (discarded)
                pk[k] = __shfl_sync(0xffffffffu, packed, j);
                const long xl = (long)(pk[k] & 0x7fff) * C, xh = (long)((pk[k] >> 15) & 0x7fff) * C;
                tl[k] = ldg4(rowT + xl); tr[k] = ldg4(rowT + xh);
                bl[k] = ldg4(rowB + xl); br[k] = ldg4(rowB + xh);
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int j = j0 + k;
                const float fx = __shfl_sync(0xffffffffu, mine.frac, min(j, jn - 1));
                if (j < jn) {
                    const float4 top = lerp_rn(tl[k], tr[k], fx);            // crop_and_resize.c:102
                    const float4 bot = lerp_rn(bl[k], br[k], fx);            // :103-104
                    float4 v = lerp_rn(top, bot, ty.frac);                   // :106
                    if (!(pk[k] & (1 << 30))) v = make_float4(extrap, extrap, extrap, extrap);
                    st_stream4(out + (long)(jb + j) * C, v);
                    if (out2) st_stream4(out2 + (long)(jb + j) * C, v);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NHWC backward, scatter form.  Same warp-per-crop-row walk.  The contributions to the current
// (x_lo, x_hi) pixel pair of the top and of the bottom image row are kept in registers and flushed with
// 128-bit vector reductions (red.global.add.v4.f32, coalesced across the channel lanes) only when the pair
// changes: up-sampling crops and degenerate / zero-padded RoIs (all samples on one pixel -- in the
// reference hundreds of serialised atomics on one address) issue one reduction per DISTINCT pixel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(float *p, float4 v) {
    atomicAdd(reinterpret_cast<float4 *>(p), v);     // sm_90+: REDG.E.ADD.F32x4 (result unused)
}
__device__ __forceinline__ float4 f4_scale(float4 a, float w) {
    return make_float4(__fmul_rn(a.x, w), __fmul_rn(a.y, w), __fmul_rn(a.z, w), __fmul_rn(a.w, w));
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

struct BwdSet {                  // one backward crop set (device view)
    const float *grads, *grads2, *boxes;
    const int *box_ind, *src_row;
    float *gimg;
    int B, H, W, C, ph, pw, slabs;
};

// unit = (box r, crop row i, 128-channel slab); gradient loads of a batch of U samples are issued together
template <int U>
__device__ __forceinline__ void bwd_unit(const BwdSet &S, long u, int lane) {
    const float *__restrict__ grads = S.grads;
    const float *__restrict__ grads2 = S.grads2;
    const float *__restrict__ boxes = S.boxes;
    const int *__restrict__ box_ind = S.box_ind;
    const int *__restrict__ src_row = S.src_row;
    float *__restrict__ gimg = S.gimg;
    const int slabs = S.slabs, B = S.B, H = S.H, W = S.W, ph = S.ph, pw = S.pw, C = S.C;
    const int slab = (int)(u % slabs);
    const long q = u / slabs;
    const int r = (int)(q / ph), i = (int)(q - (long)r * ph);
    const int b = box_ind[r];
    if (b < 0 || b >= B) return;
    const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
    const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
    if (!ty.inside) return;
    const float sx = axis_step(x1, x2, W, pw);
    const float wy_hi = ty.frac, wy_lo = __fsub_rn(1.f, ty.frac);            // crop_and_resize.c:241,245
    const long grow = src_row ? (long)src_row[r] : (long)r;
    const int coff = slab * 128 + lane * 4;
    const float *g = grads + ((grow * ph + i) * (long)pw) * C + coff;
    const float *g2 = grads2 ? grads2 + (((long)r * ph + i) * (long)pw) * C + coff : nullptr;
    float *rowT = gimg + ((long)b * H + ty.lo) * (long)W * C + coff;
    float *rowB = gimg + ((long)b * H + ty.hi) * (long)W * C + coff;
    float4 Lt = f4_zero(), Lb = f4_zero(), Rt = f4_zero(), Rb = f4_zero();
    int cur_lo = -1, cur_hi = -1;
    for (int jb = 0; jb < pw; jb += 32) {
        const AxisTap mine = axis_sample(x1, x2, sx, jb + lane, W, pw);
        const int jn = min(32, pw - jb);
        for (int j0 = 0; j0 < jn; j0 += U) {
            float4 gv[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int j = jb + min(j0 + k, jn - 1);
                gv[k] = __ldcs(reinterpret_cast<const float4 *>(g + (long)j * C));
                if (g2) gv[k] = f4_add(gv[k], __ldcs(reinterpret_cast<const float4 *>(g2 + (long)j * C)));
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if (j0 + k >= jn) break;
                const AxisTap tx = shfl_tap(mine, j0 + k);
                if (!tx.inside) continue;
                if (!(tx.lo == cur_lo && tx.hi == cur_hi)) {
                    if (cur_lo >= 0) {
                        red_add(rowT + (long)cur_lo * C, Lt);
                        red_add(rowB + (long)cur_lo * C, Lb);
                        if ((tx.lo == cur_hi) && (tx.hi != cur_hi)) { Lt = Rt; Lb = Rb; }     // old right column becomes the new left one
                        else {
                            red_add(rowT + (long)cur_hi * C, Rt);
                            red_add(rowB + (long)cur_hi * C, Rb);
                            Lt = f4_zero(); Lb = f4_zero();
                        }
                        Rt = f4_zero(); Rb = f4_zero();
                    }
                    cur_lo = tx.lo; cur_hi = tx.hi;
                }
                const float wx_hi = tx.frac, wx_lo = __fsub_rn(1.f, tx.frac);
                const float4 dtop = f4_scale(gv[k], wy_lo), dbot = f4_scale(gv[k], wy_hi);
                Lt = f4_add(Lt, f4_scale(dtop, wx_lo)); Rt = f4_add(Rt, f4_scale(dtop, wx_hi));
                Lb = f4_add(Lb, f4_scale(dbot, wx_lo)); Rb = f4_add(Rb, f4_scale(dbot, wx_hi));
            }
        }
    }
    if (cur_lo >= 0) {
        red_add(rowT + (long)cur_lo * C, Lt);
        red_add(rowB + (long)cur_lo * C, Lb);
        red_add(rowT + (long)cur_hi * C, Rt);
        red_add(rowB + (long)cur_hi * C, Rb);
    }
}

}  // namespace fi
