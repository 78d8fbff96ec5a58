// RoIAlign backward as a GATHER: every pixel of the dense gradient map is computed by exactly one warp and
// written exactly once.
//
// The reference (crop_and_resize_kernel.cu:84-165) zero-fills the dense map and then issues 4 scalar
// atomicAdds per crop element: memset + read-modify-write traffic on the map, contention where RoIs
// overlap, and a summation order that changes from run to run.  The scatter kernels in roi_align.cu keep
// that formulation (with vector reductions); measured on B200 they stop at ~3.4 TB/s of reduction payload,
// i.e. well under the 70 % HBM bar whenever taps do not coalesce (the "big" boxes pooled on finer maps).
//
// Here the map is cut into 8x8-pixel tiles x 128-channel slabs.  One CTA owns a tile, one WARP owns one
// pixel row of it (8 pixels x 128 channels = 8 float4 accumulators per lane, in registers), lane = channel
// quad.  For every RoI whose tap footprint overlaps the tile -- found by a per-tile scan of precomputed
// per-box pixel bounds, kept in box order -- the warp looks up the box's tap table (one 8-byte entry per
// lane: lanes 0..15 hold the y taps, lanes 16..31 the x taps), derives with two ballots which crop rows /
// columns touch its pixels, loads exactly those crop gradients (coalesced 512 B per sample) and adds
// them, in (box, i, j, TL->TR->BL->BR) order, with un-fused fp32 mul/add.  That is the order and the
// arithmetic of the reference's serial CPU loop (crop_and_resize.c:190-250), so the result is
// DETERMINISTIC and BIT-IDENTICAL to the CPU reference; no memset, no atomics, the map is written once.
//
// Several crop sets that read the same feature map (the 7x7 and the 14x14 crops of a level's made-up
// map) can be folded into one pass: fi_crop_and_resize_backward_multi.
#include <stdlib.h>

#include "fi_common.cuh"

namespace fi {

constexpr int kTile = 8;               // pixels per tile edge; also warps per CTA
constexpr int kMaxCrop = 16;           // crop_h, crop_w <= 16 on this path (7 and 14 in the model)
constexpr short kNoTap = -32768;

struct TapEntry {                      // 8 bytes
    short lo, hi;
    float frac;
};

struct SetDev {                        // one crop set as the tile kernel sees it
    const float *grads;                // [rows, ph, pw, C] NHWC; row = src_row[r] (or r)
    const float *grads2;               // optional second gradient, compact rows (row = r), added to `grads`
    const int *box_ind;
    const int *src_row;
    const TapEntry *taps;              // [R, 32]
    const int4 *bounds;                // [R] (ymin, ymax, xmin, xmax) of the tap footprint; ymin > ymax = empty
    int R, ph, pw;
};

constexpr int kMaxSets = 4;
struct SetsDev {
    SetDev s[kMaxSets];
    int n;
};

// ---- prep: one warp per box: tap table + footprint bounds --------------------------------------------
__global__ void __launch_bounds__(256) bwd_prep_kernel(const float *__restrict__ boxes, const int *__restrict__ box_ind, int R, int B, int H,
                                                      int W, int ph, int pw, TapEntry *__restrict__ taps, int4 *__restrict__ bounds) {
    const int lane = threadIdx.x & 31;
    const int r = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= R) return;
    const int b = box_ind[r];
    const bool bad = (b < 0 || b >= B);
    const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
    const bool is_y = lane < 16;
    const int k = is_y ? lane : lane - 16;
    const int crop = is_y ? ph : pw, extent = is_y ? H : W;
    const float c1 = is_y ? y1 : x1, c2 = is_y ? y2 : x2;
    TapEntry e;
    e.lo = kNoTap; e.hi = kNoTap; e.frac = 0.f;
    int lo = 1 << 30, hi = -(1 << 30);
    if (!bad && k < crop) {
        const AxisTap t = axis_sample(c1, c2, axis_step(c1, c2, extent, crop), k, extent, crop);
        if (t.inside) { e.lo = (short)t.lo; e.hi = (short)t.hi; e.frac = t.frac; lo = t.lo; hi = t.hi; }
    }
    taps[(long)r * 32 + lane] = e;
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) {             // min/max inside each half-warp
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    const int xlo = __shfl_sync(0xffffffffu, lo, 16), xhi = __shfl_sync(0xffffffffu, hi, 16);
    if (lane == 0) {
        const bool empty = (lo > hi) || (xlo > xhi);
        bounds[r] = empty ? make_int4(1, 0, 1, 0) : make_int4(lo, hi, xlo, xhi);
    }
}

__device__ __forceinline__ float4 add_rn4(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 mul_rn4(float w, float4 a) {
    return make_float4(__fmul_rn(w, a.x), __fmul_rn(w, a.y), __fmul_rn(w, a.z), __fmul_rn(w, a.w));
}

// acc[x] += v with x warp-uniform: a uniform branch tree instead of eight predicated adds
__device__ __forceinline__ void acc_add(float4 (&acc)[kTile], int x, float4 v) {
    switch (x) {
        case 0: acc[0] = add_rn4(acc[0], v); break;
        case 1: acc[1] = add_rn4(acc[1], v); break;
        case 2: acc[2] = add_rn4(acc[2], v); break;
        case 3: acc[3] = add_rn4(acc[3], v); break;
        case 4: acc[4] = add_rn4(acc[4], v); break;
        case 5: acc[5] = add_rn4(acc[5], v); break;
        case 6: acc[6] = add_rn4(acc[6], v); break;
        case 7: acc[7] = add_rn4(acc[7], v); break;
        default: break;
    }
}

// ---- tile kernel ---------------------------------------------------------------------------------------
// grid (tiles_x, tiles_y, B * slabs), 256 threads: warp w = pixel row Y0 + w, lane = channel quad of the slab.
__global__ void __launch_bounds__(kTile * 32) bwd_tile_kernel(const SetsDev sets, int B, int H, int W, int C, int slabs, int accumulate,
                                                             float *__restrict__ gimg) {
    __shared__ int list[256];
    __shared__ int warp_hits[kTile];
    __shared__ int nlist;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.z / slabs, slab = blockIdx.z - b * slabs;
    const int X0 = blockIdx.x * kTile, Y0 = blockIdx.y * kTile;
    const int y = Y0 + w;
    const int coff = slab * 128 + lane * 4;
    float4 acc[kTile];
#pragma unroll
    for (int x = 0; x < kTile; ++x) acc[x] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int si = 0; si < sets.n; ++si) {
        const SetDev &S = sets.s[si];
        for (int base = 0; base < S.R; base += 256) {
            // ---- ordered compaction of the boxes of this chunk whose footprint overlaps the tile
            const int r = base + threadIdx.x;
            bool hit = false;
            if (r < S.R && S.box_ind[r] == b) {
                const int4 bd = S.bounds[r];
                hit = bd.x <= Y0 + kTile - 1 && bd.y >= Y0 && bd.z <= X0 + kTile - 1 && bd.w >= X0;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) warp_hits[w] = __popc(m);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int q = 0; q < kTile; ++q) { const int h = warp_hits[q]; if (q < w) before += h; total += h; }
            if (hit) list[before + __popc(m & ((1u << lane) - 1u))] = r;
            if (threadIdx.x == 0) nlist = total;
            __syncthreads();
            const int nl = nlist;
            // ---- every warp walks the hit list for its own pixel row
            if (y < H) {
                for (int k = 0; k < nl; ++k) {
                    const int rr = list[k];
                    const int4 bd = S.bounds[rr];
                    if (y < bd.x || y > bd.y) continue;                                   // warp-uniform
                    const TapEntry e = S.taps[(long)rr * 32 + lane];
                    const unsigned ymask = __ballot_sync(0xffffffffu, lane < 16 && (e.lo == y || e.hi == y));
                    const unsigned xmask = __ballot_sync(0xffffffffu, lane >= 16 && ((e.lo >= X0 && e.lo < X0 + kTile) || (e.hi >= X0 && e.hi < X0 + kTile))) >> 16;
                    if (ymask == 0 || xmask == 0) continue;
                    const long grow = S.src_row ? (long)S.src_row[rr] : (long)rr;
                    unsigned ym = ymask;
                    while (ym) {
                        const int i = __ffs(ym) - 1;
                        ym &= ym - 1;
                        const int ylo = __shfl_sync(0xffffffffu, (int)e.lo, i), yhi = __shfl_sync(0xffffffffu, (int)e.hi, i);
                        const float fy = __shfl_sync(0xffffffffu, e.frac, i);
                        const float wy_top = __fsub_rn(1.f, fy);                              // crop_and_resize.c:241
                        const float *g1 = S.grads + ((grow * S.ph + i) * (long)S.pw) * C + coff;
                        const float *g2 = S.grads2 ? S.grads2 + (((long)rr * S.ph + i) * (long)S.pw) * C + coff : nullptr;
                        unsigned xm = xmask;
                        while (xm) {
                            // up to 4 samples of this crop row per batch: loads first, then the ordered adds
                            int js[4];
                            float4 gv[4];
                            int nb = 0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (xm) { js[q] = __ffs(xm) - 1; xm &= xm - 1; nb = q + 1; } else js[q] = js[q > 0 ? q - 1 : 0];
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (q < nb) {
                                    gv[q] = __ldg(reinterpret_cast<const float4 *>(g1 + (long)js[q] * C));
                                    if (g2) gv[q] = add_rn4(gv[q], __ldg(reinterpret_cast<const float4 *>(g2 + (long)js[q] * C)));
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (q < nb) {
                                    const int j = js[q];
                                    const int xlo = __shfl_sync(0xffffffffu, (int)e.lo, 16 + j) - X0, xhi = __shfl_sync(0xffffffffu, (int)e.hi, 16 + j) - X0;
                                    const float fx = __shfl_sync(0xffffffffu, e.frac, 16 + j);
                                    const float wx_lo = __fsub_rn(1.f, fx);
                                    if (ylo == y) {                                           // TL, TR   (:242-243)
                                        const float4 dtop = mul_rn4(wy_top, gv[q]);
                                        acc_add(acc, xlo, mul_rn4(wx_lo, dtop));
                                        acc_add(acc, xhi, mul_rn4(fx, dtop));
                                    }
                                    if (yhi == y) {                                           // BL, BR   (:245-247)
                                        const float4 dbot = mul_rn4(fy, gv[q]);
                                        acc_add(acc, xlo, mul_rn4(wx_lo, dbot));
                                        acc_add(acc, xhi, mul_rn4(fx, dbot));
                                    }
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();           // the list is rebuilt by the next chunk
        }
    }
    if (y < H) {
        float *dst = gimg + (((long)b * H + y) * (long)W + X0) * C + coff;
#pragma unroll
        for (int x = 0; x < kTile; ++x) {
            if (X0 + x < W) {
                float4 v = acc[x];
                if (accumulate) v = add_rn4(*reinterpret_cast<const float4 *>(dst + (long)x * C), v);
                *reinterpret_cast<float4 *>(dst + (long)x * C) = v;
            }
        }
    }
}

}  // namespace fi

using namespace fi;

// Host side of the gather path.  Returns FI_ERR_UNSUPPORTED (without touching gimg) when a set does not
// qualify, so the caller can fall back to the scatter kernels.
static int gather_backward(const fi_crop_set *sets, int nsets, int B, int H, int W, int C, float *gimg, int accumulate, cudaStream_t stream) {
    if (nsets < 1 || nsets > kMaxSets || C % 128 != 0 || H > 32767 || W > 32767 || ((uintptr_t)gimg % 16) != 0) return FI_ERR_UNSUPPORTED;
    size_t bytes = 0;
    for (int i = 0; i < nsets; ++i) {
        const fi_crop_set &s = sets[i];
        if (s.crop_height > kMaxCrop || s.crop_width > kMaxCrop || s.crop_height < 1 || s.crop_width < 1) return FI_ERR_UNSUPPORTED;
        if (((uintptr_t)s.grads % 16) != 0 || ((uintptr_t)s.grads2 % 16) != 0) return FI_ERR_UNSUPPORTED;
        bytes += (size_t)s.num_boxes * (32 * sizeof(TapEntry) + sizeof(int4));
    }
    char *ws = nullptr;
    if (bytes) {
        cudaError_t e = cudaMallocAsync((void **)&ws, bytes, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "crop_and_resize backward: workspace (%zu B): %s", bytes, cudaGetErrorString(e)); return FI_ERR_CUDA; }
    }
    SetsDev dev;
    dev.n = 0;
    char *p = ws;
    for (int i = 0; i < nsets; ++i) {
        const fi_crop_set &s = sets[i];
        if (s.num_boxes == 0) continue;
        SetDev &d = dev.s[dev.n++];
        d.grads = s.grads; d.grads2 = s.grads2; d.box_ind = s.box_ind; d.src_row = s.src_row;
        d.R = s.num_boxes; d.ph = s.crop_height; d.pw = s.crop_width;
        d.bounds = reinterpret_cast<const int4 *>(p); p += (size_t)s.num_boxes * sizeof(int4);
        d.taps = reinterpret_cast<const TapEntry *>(p); p += (size_t)s.num_boxes * 32 * sizeof(TapEntry);
        bwd_prep_kernel<<<ceil_div(s.num_boxes, 8), 256, 0, stream>>>(s.boxes, s.box_ind, s.num_boxes, B, H, W, s.crop_height, s.crop_width,
                                                                     const_cast<TapEntry *>(d.taps), const_cast<int4 *>(d.bounds));
        if (int e = check_launch("crop_and_resize backward[prep]")) { if (ws) cudaFreeAsync(ws, stream); return e; }
    }
    const int slabs = C / 128;
    dim3 grid(ceil_div(W, kTile), ceil_div(H, kTile), B * slabs);
    if (grid.z > 65535 || grid.y > 65535) { if (ws) cudaFreeAsync(ws, stream); return FI_ERR_UNSUPPORTED; }
    bwd_tile_kernel<<<grid, kTile * 32, 0, stream>>>(dev, B, H, W, C, slabs, accumulate, gimg);
    const int rc = check_launch("crop_and_resize backward[tile]");
    if (ws) cudaFreeAsync(ws, stream);
    return rc;
}

int fi_scatter_backward_nhwc(const float *grads, const float *boxes, const int *box_ind, const int *src_row, int R, int B, int H, int W, int ph,
                             int pw, int C, float *gimg, cudaStream_t stream);   // roi_align.cu

FI_API int fi_crop_and_resize_backward_multi(const fi_crop_set *sets, int num_sets, int batch, int image_height, int image_width, int depth,
                                             float *grads_image, int accumulate, cudaStream_t stream) {
    FI_REQUIRE(sets && num_sets >= 1 && batch > 0 && image_height > 0 && image_width > 0 && depth > 0 && grads_image, "fi_crop_and_resize_backward_multi: bad arguments");
    for (int i = 0; i < num_sets; ++i) {
        const fi_crop_set &s = sets[i];
        FI_REQUIRE(s.num_boxes >= 0 && s.crop_height > 0 && s.crop_width > 0, "fi_crop_and_resize_backward_multi: bad set %d", i);
        FI_REQUIRE(s.num_boxes == 0 || (s.grads && s.boxes && s.box_ind), "fi_crop_and_resize_backward_multi: null pointer in set %d", i);
    }
    const char *force = getenv("FI_BWD");      // "scatter" forces the reduction kernels (A/B measurements)
    int rc = (force && force[0] == 's') ? FI_ERR_UNSUPPORTED : gather_backward(sets, num_sets, batch, image_height, image_width, depth, grads_image, accumulate, stream);
    if (rc != FI_ERR_UNSUPPORTED) return rc;
    // fallback: zero-fill + scatter, set by set
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(grads_image, 0, sizeof(float) * (size_t)batch * depth * image_height * image_width, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_crop_and_resize_backward_multi: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    }
    for (int i = 0; i < num_sets; ++i) {
        const fi_crop_set &s = sets[i];
        if (s.num_boxes == 0) continue;
        rc = fi_scatter_backward_nhwc(s.grads, s.boxes, s.box_ind, s.src_row, s.num_boxes, batch, image_height, image_width, s.crop_height,
                                      s.crop_width, depth, grads_image, stream);
        if (rc) return rc;
        if (s.grads2) {
            rc = fi_scatter_backward_nhwc(s.grads2, s.boxes, s.box_ind, nullptr, s.num_boxes, batch, image_height, image_width, s.crop_height,
                                          s.crop_width, depth, grads_image, stream);
            if (rc) return rc;
        }
    }
    return ok();
}
