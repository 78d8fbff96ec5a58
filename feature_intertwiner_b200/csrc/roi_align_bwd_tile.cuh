// Shared declarations of the tile-owner RoIAlign backward (roi_align_bwd_tile.cu: prep / collapse / enumerate and the
// shared-memory accumulate kernels; roi_align_bwd_pix.cu: the bulk-copy staged, register-accumulating accumulate kernel).
#pragma once
#include "fi_common.cuh"

namespace fi {
namespace tile {

constexpr int kQ = 64;                             // per-warp sample queue
constexpr int kMaxCrop = 16;                       // crop_h, crop_w <= 16 (the model uses 7 and 14)
constexpr int kNoTap = -32768;
constexpr int kMaxSets = 12, kMaxMaps = 8;
// per-tile sample lists (enumerate -> accumulate): 64-entry chunks; entry = (gradient row, pk2) + 4 tap weights.
// pk2: tap flags TL|TR|BL|BR (bits 0-3), kDefer2 (bit 4), source index 3 * set + {grads, grads2, collapse rows} (bits 5-10),
//      TL pixel index inside the tile, signed (bits 11..).
constexpr int kDefer2 = 1 << 4;
constexpr int kChunk = 64;

struct TSet {
    const float *grads, *grads2;
    const float4 *boxes;
    const int *box_ind, *src_row;
    int4 *rec;                         // [R] (ymin | ymax << 16, xmin | xmax << 16, image or -1, degenerate)
    float4 *geom;                      // [R] (y of sample row 0, y step, x of sample column 0, x step) in pixels
    unsigned *range;                   // [B,2]: min box index of image b, ~(max box index); memset 0xFF = "none"
    float *coll;                       // [R,4,C] corner sums of degenerate boxes (only those rows are written)
    const int *R_dev;                  // NULL, or the actual number of boxes (<= R) on the device
    int R, ph, pw, map;                // R: capacity of the lists above
};
struct TMap {
    float *gimg;
    int B, H, W, C, tiles_x, tiles_y, first_tile, set_begin, set_end;
};
struct BinWs {                         // per-tile sample lists of the two-kernel form (enumerate -> accumulate)
    int *tile_head;                    // [tiles] first chunk of the tile's list
    int *tile_n;                       // [tiles] entries in the list (a multiple of 8)
    int *chunk_next;                   // [pool]  chunk -> next chunk of the same tile
    uint2 *qa;                         // [pool * 64] (gradient row, pk2)
    float4 *qw;                        // [pool * 64] tap weights
    int *cursor;                       // [0] chunks handed out, [1] overflow flag (a list chunk did not fit the pool)
    int *work;                         // [32] tile counters of the persistent accumulate kernel (one per channel block)
    int pool, total_tiles, pix_group;
};
struct TParams {
    TSet s[kMaxSets];
    TMap m[kMaxMaps];
    int *deg_list;                     // [0] = count, then (set << 24 | box) entries
    BinWs bin;
    size_t range_bytes;                // the sets' `range` arrays are one block of this size, starting at s[0].range
    int nsets, nmaps, accumulate, collapse;
};

struct Tap {                           // one axis tap as the tile kernel uses it
    int lo, hi;                        // kNoTap when the sample is outside the image
    float frac;
};
// One axis tap from the per-box geometry record (base, step): the same fp32 operations as fi_common.cuh::axis_sample.
__device__ __forceinline__ Tap geom_tap(float base, float step, int k, int extent) {
    const float pos = __fadd_rn(base, __fmul_rn((float)k, step));
    Tap t;
    t.lo = kNoTap; t.hi = kNoTap; t.frac = 0.f;
    if (!(pos < 0.f || pos > (float)(extent - 1))) {
        t.lo = (int)floorf(pos);
        t.hi = (int)ceilf(pos);
        t.frac = __fsub_rn(pos, (float)t.lo);
    }
    return t;
}
__device__ __forceinline__ float geom_base(float c1, float c2, int extent, int crop) {      // crop_and_resize.c:52-56
    if (crop > 1) return __fmul_rn(c1, (float)(extent - 1));
    return (float)(0.5 * (double)__fadd_rn(c1, c2) * (double)(extent - 1));
}

// ---- arithmetic -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 add_rn4(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 mul_rn4(float w, float4 a) {
    return make_float4(__fmul_rn(w, a.x), __fmul_rn(w, a.y), __fmul_rn(w, a.z), __fmul_rn(w, a.w));
}
// a + g * w on two packed pairs (FFMA2)
__device__ __forceinline__ float4 fma4(float4 g, float w, float4 a) {
    unsigned long long g0, g1, a0, a1, ww, r0, r1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(g0) : "f"(g.x), "f"(g.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(g1) : "f"(g.z), "f"(g.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a0) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a1) : "f"(a.z), "f"(a.w));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r0) : "l"(g0), "l"(ww), "l"(a0));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r1) : "l"(g1), "l"(ww), "l"(a1));
    float4 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(r0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.z), "=f"(r.w) : "l"(r1));
    return r;
}

}  // namespace tile
}  // namespace fi
