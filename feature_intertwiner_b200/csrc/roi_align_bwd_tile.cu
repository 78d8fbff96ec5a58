// RoIAlign backward, tile-owner form: the dense gradient maps are cut into 4x8-pixel tiles, every tile is
// accumulated in SHARED MEMORY by the one warp that owns it (per 128-channel slab) and written to HBM exactly once.
//
// Why: the reference (crop_and_resize_kernel.cu:84-165) zero-fills the map and issues 4 atomics per crop element.
// The vector-reduction rewrite of that formulation (roi_align.cu) is bound by L2 reduction throughput (~3.5 TB/s of
// RED payload, ~5 GB per C2 step) plus a 1.5 GB zero fill -- 37 % of the HBM roofline.  Here the only HBM traffic is
// the algorithmic one: every crop gradient is read (once, plus re-reads by neighbouring tiles that L2 absorbs) and
// every map pixel is written once; no memset, no atomics, no read-modify-write in L2/HBM.
//
// Work decomposition.  CTA = 2 warps = one tile x two 128-channel slabs; the warps never synchronise with each other
// (each owns its slab of the tile in shared memory).  Per warp:
//   scan     the boxes of the tile's image (index range from the prep kernel), 32 at a time, footprint-bounds test,
//            ballot-compacted IN BOX ORDER into a hit list;
//   expand   per hit box the warp loads the box's tap table (lanes 0-15 = y taps, 16-31 = x taps, 8 B per lane),
//            two ballots give the crop rows / columns that touch the tile; lanes then describe the samples of that
//            rectangle in parallel (tap offsets, in-tile flags, lerp weights, gradient row) and append the ones with at
//            least one tap inside the tile to a per-warp queue -- in (box, crop row, crop column) order;
//   drain    the queue is consumed in order, 8 (4 for two-source sets) 512-byte gradient loads in flight ahead of the
//            adds (double-buffered in registers); each tap is one conflict-free LDS.128 / add / STS.128;
//   store    the tile is streamed out, 512 B per warp instruction.
// Because every pixel is summed by ONE warp in the order (box, crop row, crop column, TL->TR->BL->BR) the result is
// run-to-run deterministic; with EXACT arithmetic (un-fused fp32 mul then add, crop_and_resize.c:241-247) it is
// bit-identical to the reference's serial CPU loop (crop_and_resize.c:190-250).  The default mode uses packed FMAs
// (fma.rn.f32x2 -> FFMA2) with pre-multiplied weights: same order, one rounding fewer per contribution.
#include <stdlib.h>

#include "fi_common.cuh"

namespace fi {
namespace tile {

constexpr int kTX = 8, kTY = 4, kTP = kTX * kTY;   // tile: 4 rows x 8 pixels; a tile row is 8 KB contiguous in NHWC (C=256)
constexpr int kQ = 64;                             // per-warp sample queue
constexpr int kList = 64;                          // per-warp hit list
constexpr int kMaxCrop = 16;                       // crop_h, crop_w <= 16 (the model uses 7 and 14)
constexpr short kNoTap = -32768;
constexpr int kMaxSets = 12, kMaxMaps = 8;

struct TapEntry {                      // 8 bytes
    short lo, hi;
    float frac;
};

struct TSet {
    const float *grads, *grads2, *boxes;
    const int *box_ind, *src_row;
    TapEntry *taps;                    // [R,32]
    int4 *bounds;                      // [R] (ymin, ymax, xmin, xmax) of the tap footprint; ymin > ymax: empty
    unsigned *range;                   // [B,2]: min box index of image b, ~(max box index); memset 0xFF = "none"
    int R, ph, pw, map;
};
struct TMap {
    float *gimg;
    int B, H, W, C, tiles_x, tiles_y, first_tile, set_begin, set_end;
};
struct TParams {
    TSet s[kMaxSets];
    TMap m[kMaxMaps];
    int nsets, nmaps, accumulate;
};

// ---- prep: one warp per box (all sets in one launch): tap table, footprint bounds, per-image index range ----------
__global__ void __launch_bounds__(256) tile_prep_kernel(const TParams P) {
    const TSet &S = P.s[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= S.R) return;
    const int B = P.m[S.map].B, H = P.m[S.map].H, W = P.m[S.map].W;
    const int b = S.box_ind[r];
    const bool bad = (b < 0 || b >= B);            // skipped like crop_and_resize_kernel.cu:34-38
    if (!bad && lane == 0) { atomicMin(S.range + 2 * b, (unsigned)r); atomicMin(S.range + 2 * b + 1, ~(unsigned)r); }
    const float y1 = S.boxes[4 * r + 0], x1 = S.boxes[4 * r + 1], y2 = S.boxes[4 * r + 2], x2 = S.boxes[4 * r + 3];
    const bool is_y = lane < 16;
    const int k = is_y ? lane : lane - 16;
    const int crop = is_y ? S.ph : S.pw, extent = is_y ? H : W;
    const float c1 = is_y ? y1 : x1, c2 = is_y ? y2 : x2;
    TapEntry e;
    e.lo = kNoTap; e.hi = kNoTap; e.frac = 0.f;
    int lo = 1 << 30, hi = -(1 << 30);
    if (!bad && k < crop) {
        const AxisTap t = axis_sample(c1, c2, axis_step(c1, c2, extent, crop), k, extent, crop);
        if (t.inside) { e.lo = (short)t.lo; e.hi = (short)t.hi; e.frac = t.frac; lo = t.lo; hi = t.hi; }
    }
    S.taps[(long)r * 32 + lane] = e;
#pragma unroll
    for (int d = 8; d > 0; d >>= 1) {              // min / max inside each half-warp
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    const int xlo = __shfl_sync(0xffffffffu, lo, 16), xhi = __shfl_sync(0xffffffffu, hi, 16);
    if (lane == 0) {
        const bool empty = (lo > hi) || (xlo > xhi);
        S.bounds[r] = empty ? make_int4(1, 0, 1, 0) : make_int4(lo, hi, xlo, xhi);
    }
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

struct __align__(16) WarpSmem {
    float tile[kTP * 128];             // this warp's slab of the tile: [pixel][128 channels]
    uint4 q4[kQ];                      // sample queue: (row of grads, row of grads2, fy bits, fx bits)
    int qp[kQ];                        //               in-tile flags TL|TR|BL|BR (bits 0-3), TL pixel offset + 16 (bits 4..)
    int list[kList];                   // hit boxes, in index order
};

// One queued sample: up to 4 read-modify-writes of this lane's float4 in the tile.  Flags / offsets are warp-uniform.
template <bool EXACT>
__device__ __forceinline__ void apply_sample(float *tile_lane, int pk, float fy, float fx, float4 g) {
    float4 *tp = reinterpret_cast<float4 *>(tile_lane + ((pk >> 4) - 16) * 128);
    const float wy0 = __fsub_rn(1.f, fy), wx0 = __fsub_rn(1.f, fx);                 // crop_and_resize.c:241-247
    if (EXACT) {
        const float4 dtop = mul_rn4(wy0, g), dbot = mul_rn4(fy, g);
        if (pk & 1) tp[0] = add_rn4(tp[0], mul_rn4(wx0, dtop));
        if (pk & 2) tp[32] = add_rn4(tp[32], mul_rn4(fx, dtop));
        if (pk & 4) tp[kTX * 32] = add_rn4(tp[kTX * 32], mul_rn4(wx0, dbot));
        if (pk & 8) tp[kTX * 32 + 32] = add_rn4(tp[kTX * 32 + 32], mul_rn4(fx, dbot));
    } else {
        if (pk & 1) tp[0] = fma4(g, wy0 * wx0, tp[0]);
        if (pk & 2) tp[32] = fma4(g, wy0 * fx, tp[32]);
        if (pk & 4) tp[kTX * 32] = fma4(g, fy * wx0, tp[kTX * 32]);
        if (pk & 8) tp[kTX * 32 + 32] = fma4(g, fy * fx, tp[kTX * 32 + 32]);
    }
}

// Consume the queue in order.  Groups of U samples; the loads of group i+1 are in flight while group i is added.
template <bool EXACT, bool DUAL>
__device__ __forceinline__ void drain(WarpSmem &ws, int qn, const float *__restrict__ G1, const float *__restrict__ G2, int C,
                                      float *tile_lane) {
    constexpr int U = DUAL ? 4 : 8;
    constexpr int UB = DUAL ? U : 1;
    if (qn <= 0) return;
    float4 a0[U], a1[U], b0[UB], b1[UB];
    auto load = [&](int g0, float4(&a)[U], float4(&b)[UB]) {
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int idx = min(g0 + k, qn - 1);           // the tail re-reads the last sample (discarded): branch-free
            const uint2 rows = *reinterpret_cast<const uint2 *>(&ws.q4[idx]);
            a[k] = __ldcg(reinterpret_cast<const float4 *>(G1 + (size_t)rows.x * C));
            if (DUAL) b[k] = __ldcg(reinterpret_cast<const float4 *>(G2 + (size_t)rows.y * C));
        }
    };
    auto process = [&](int g0, float4(&a)[U], float4(&b)[UB]) {
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int idx = g0 + k;
            if (idx < qn) {
                const uint4 d = ws.q4[idx];
                const int pk = ws.qp[idx];
                const float4 g = DUAL ? add_rn4(a[k], b[k]) : a[k];
                apply_sample<EXACT>(tile_lane, pk, __uint_as_float(d.z), __uint_as_float(d.w), g);
            }
        }
    };
    load(0, a0, b0);
    for (int g0 = 0; g0 < qn; g0 += 2 * U) {
        const bool more = g0 + U < qn;
        if (more) load(g0 + U, a1, b1);
        process(g0, a0, b0);
        if (more) {
            if (g0 + 2 * U < qn) load(g0 + 2 * U, a0, b0);
            process(g0 + U, a1, b1);
        }
    }
}

template <bool EXACT>
__device__ __noinline__ void drain_any(WarpSmem &ws, int qn, const float *G1, const float *G2, int C, float *tile_lane, bool zero) {
    if (zero) {                                        // first use of the tile
#pragma unroll
        for (int p = 0; p < kTP; ++p) *reinterpret_cast<float4 *>(tile_lane + p * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    if (G2) drain<EXACT, true>(ws, qn, G1, G2, C, tile_lane);
    else drain<EXACT, false>(ws, qn, G1, nullptr, C, tile_lane);
    __syncwarp();
}

// grid (total tiles of all maps, ceil(slabs / 2)), 64 threads
template <bool EXACT>
__global__ void __launch_bounds__(64, 6) bwd_smem_tile_kernel(const TParams P) {
    __shared__ WarpSmem smem[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int t = blockIdx.x, mi = 0;
    while (mi + 1 < P.nmaps && t >= P.m[mi + 1].first_tile) ++mi;
    const TMap &M = P.m[mi];
    const int H = M.H, W = M.W, C = M.C;
    const int slab = blockIdx.y * 2 + w;
    if (slab * 128 >= C) return;                       // the warps of a CTA never synchronise with each other
    t -= M.first_tile;
    const int per_img = M.tiles_x * M.tiles_y;
    const int b = t / per_img;
    t -= b * per_img;
    const int tyi = t / M.tiles_x, txi = t - tyi * M.tiles_x;
    const int X0 = txi * kTX, Y0 = tyi * kTY;
    const int coff = slab * 128 + lane * 4;
    WarpSmem &ws = smem[w];
    float *tile_lane = ws.tile + lane * 4;
    const unsigned lt = (1u << lane) - 1u;
    bool dirty = false;                                // tile zeroed lazily, on the first queued sample

    for (int si = M.set_begin; si < M.set_end; ++si) {
        const TSet &S = P.s[si];
        const unsigned first = S.range[2 * b], last = ~S.range[2 * b + 1];
        if (first > last) continue;                    // no box of this set lives in image b
        const int ph = S.ph, pw = S.pw;
        int nl = 0, qn = 0;
        for (unsigned base = first & ~31u; base <= last; base += 32) {
            // ---- scan: ordered compaction of the boxes whose footprint overlaps the tile
            const unsigned r = base + lane;
            bool hit = false;
            if (r >= first && r <= last && S.box_ind[r] == b) {
                const int4 bd = S.bounds[r];
                hit = bd.x <= Y0 + kTY - 1 && bd.y >= Y0 && bd.z <= X0 + kTX - 1 && bd.w >= X0;
            }
            const unsigned hm = __ballot_sync(0xffffffffu, hit);
            if (hit) ws.list[nl + __popc(hm & lt)] = (int)r;
            nl += __popc(hm);
            const bool last_chunk = base + 32 > last || base + 32 < base;
            if (!(nl > kList - 32 || last_chunk)) continue;
            __syncwarp();
            // ---- expand the hit boxes into queued samples
            long long e_next = 0;
            int grow_next = 0;
            if (nl > 0) {
                const int r0 = ws.list[0];
                e_next = *reinterpret_cast<const long long *>(S.taps + (long)r0 * 32 + lane);
                grow_next = S.src_row ? S.src_row[r0] : r0;
            }
            for (int k = 0; k < nl; ++k) {
                const int rr = ws.list[k];
                const long long e_cur = e_next;
                const int grow = grow_next;
                if (k + 1 < nl) {                                                  // prefetch the next box's taps
                    const int rn = ws.list[k + 1];
                    e_next = *reinterpret_cast<const long long *>(S.taps + (long)rn * 32 + lane);
                    grow_next = S.src_row ? S.src_row[rn] : rn;
                }
                const int lohi = (int)(e_cur & 0xffffffffll);
                const float frac = __int_as_float((int)(e_cur >> 32));
                const int lo = (short)(lohi & 0xffff), hi = (short)(lohi >> 16);
                const bool touch = lane < 16 ? ((lo >= Y0 && lo < Y0 + kTY) || (hi >= Y0 && hi < Y0 + kTY))
                                             : ((lo >= X0 && lo < X0 + kTX) || (hi >= X0 && hi < X0 + kTX));
                const unsigned bal = __ballot_sync(0xffffffffu, touch);
                const unsigned ymask = bal & 0xffffu, xmask = bal >> 16;
                if (ymask == 0 || xmask == 0) continue;
                const int iy0 = __ffs(ymask) - 1, iy1 = 31 - __clz(ymask);
                const int ix0 = __ffs(xmask) - 1, ix1 = 31 - __clz(xmask);
                const int nx = ix1 - ix0 + 1, n = (iy1 - iy0 + 1) * nx;          // <= 256 samples of this box may touch the tile
                const unsigned inv = (65536u + nx - 1) / nx;                       // s / nx == (s * inv) >> 16 for s < 256, nx <= 16
                for (int c0 = 0; c0 < n; c0 += 32) {
                    const int s = c0 + lane;
                    const bool valid = s < n;
                    const int ii = valid ? (int)((s * inv) >> 16) : 0;
                    const int jj = valid ? s - ii * nx : 0;
                    const int iy = iy0 + ii, ix = ix0 + jj;
                    const int py = __shfl_sync(0xffffffffu, lohi, iy), px = __shfl_sync(0xffffffffu, lohi, 16 + ix);
                    const float fy = __shfl_sync(0xffffffffu, frac, iy), fx = __shfl_sync(0xffffffffu, frac, 16 + ix);
                    const int ylo = (short)(py & 0xffff), yhi = (short)(py >> 16);
                    const int xlo = (short)(px & 0xffff), xhi = (short)(px >> 16);
                    // a tap that coincides with its partner (integer sample position) carries weight 0: dropped
                    const bool top = ylo >= Y0 && ylo < Y0 + kTY, bot = yhi >= Y0 && yhi < Y0 + kTY && yhi != ylo;
                    const bool lef = xlo >= X0 && xlo < X0 + kTX, rig = xhi >= X0 && xhi < X0 + kTX && xhi != xlo;
                    int f = (top && lef ? 1 : 0) | (top && rig ? 2 : 0) | (bot && lef ? 4 : 0) | (bot && rig ? 8 : 0);
                    if (!valid) f = 0;
                    const unsigned am = __ballot_sync(0xffffffffu, f != 0);
                    if (f) {
                        const int pos = qn + __popc(am & lt);
                        const int tl = (ylo - Y0) * kTX + (xlo - X0) + 16;         // >= 16 - 9
                        const unsigned row1 = (unsigned)((grow * ph + iy) * pw + ix);
                        const unsigned row2 = (unsigned)((rr * ph + iy) * pw + ix);
                        ws.q4[pos] = make_uint4(row1, row2, __float_as_uint(fy), __float_as_uint(fx));
                        ws.qp[pos] = f | (tl << 4);
                    }
                    qn += __popc(am);
                    if (qn > kQ - 32) {
                        drain_any<EXACT>(ws, qn, S.grads + coff, S.grads2 ? S.grads2 + coff : nullptr, C, tile_lane, !dirty);
                        dirty = true;
                        qn = 0;
                    }
                }
            }
            nl = 0;
            __syncwarp();
        }
        if (qn > 0) {                                   // set boundary: the queue holds rows of one set only
            drain_any<EXACT>(ws, qn, S.grads + coff, S.grads2 ? S.grads2 + coff : nullptr, C, tile_lane, !dirty);
            dirty = true;
        }
    }
    // ---- store the tile (or zeros), once
    __syncwarp();
    float *dst0 = M.gimg + (((long)b * H + Y0) * (long)W + X0) * C + coff;
#pragma unroll 8
    for (int p = 0; p < kTP; ++p) {
        const int yy = p / kTX, xx = p % kTX;
        if (Y0 + yy < H && X0 + xx < W) {
            float *dst = dst0 + ((long)yy * W + xx) * C;
            float4 v = dirty ? *reinterpret_cast<const float4 *>(tile_lane + p * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (P.accumulate) v = add_rn4(*reinterpret_cast<const float4 *>(dst), v);
            __stcs(reinterpret_cast<float4 *>(dst), v);
        }
    }
}

}  // namespace tile
}  // namespace fi

using namespace fi;
using namespace fi::tile;

// Host side.  Returns FI_ERR_UNSUPPORTED (nothing touched) when the sets do not qualify so that the caller can use the
// reduction kernels.  exact != 0: arithmetic and order of crop_and_resize.c:190-250 (bit-identical for one set per map).
int fi_tile_backward(const fi_bwd_set *sets, int num_sets, int accumulate, int exact, cudaStream_t stream) {
    if (num_sets < 1 || num_sets > kMaxSets) return FI_ERR_UNSUPPORTED;
    TParams P;
    P.nsets = 0; P.nmaps = 0; P.accumulate = accumulate ? 1 : 0;
    // group the sets by map, maps in order of first appearance
    for (int i = 0; i < num_sets; ++i) {
        const fi_bwd_set &h = sets[i];
        if (h.depth % 128 != 0 || h.image_height > 32767 || h.image_width > 32767 || h.crop_height > kMaxCrop || h.crop_width > kMaxCrop ||
            h.crop_height < 1 || h.crop_width < 1) return FI_ERR_UNSUPPORTED;
        if (((uintptr_t)h.grads_image % 16) || ((uintptr_t)h.grads % 16) || ((uintptr_t)h.grads2 % 16)) return FI_ERR_UNSUPPORTED;
        if ((long)h.num_boxes * h.crop_height * h.crop_width >= (1L << 31)) return FI_ERR_UNSUPPORTED;
        bool seen = false;
        for (int q = 0; q < i; ++q) seen = seen || (sets[q].grads_image == h.grads_image);
        if (seen) continue;
        if (P.nmaps == kMaxMaps) return FI_ERR_UNSUPPORTED;
        TMap &M = P.m[P.nmaps];
        M.gimg = h.grads_image; M.B = h.batch; M.H = h.image_height; M.W = h.image_width; M.C = h.depth;
        M.tiles_x = ceil_div(M.W, kTX); M.tiles_y = ceil_div(M.H, kTY);
        M.set_begin = P.nsets;
        for (int q = i; q < num_sets; ++q) {
            const fi_bwd_set &g = sets[q];
            if (g.grads_image != h.grads_image) continue;
            if (g.batch != h.batch || g.image_height != h.image_height || g.image_width != h.image_width || g.depth != h.depth) {
                set_error(FI_ERR_INVALID, "crop backward: sets %d and %d name the same map with different shapes", i, q);
                return FI_ERR_INVALID;
            }
            TSet &S = P.s[P.nsets++];
            S.grads = g.grads; S.grads2 = g.grads2; S.boxes = g.boxes; S.box_ind = g.box_ind; S.src_row = g.src_row;
            S.R = g.num_boxes; S.ph = g.crop_height; S.pw = g.crop_width; S.map = P.nmaps;
        }
        M.set_end = P.nsets;
        ++P.nmaps;
    }
    long tiles = 0;
    int max_slabs = 1, max_R = 0;
    for (int m = 0; m < P.nmaps; ++m) {
        TMap &M = P.m[m];
        if (tiles + (long)M.B * M.tiles_x * M.tiles_y >= (1L << 31)) return FI_ERR_UNSUPPORTED;
        M.first_tile = (int)tiles;
        tiles += (long)M.B * M.tiles_x * M.tiles_y;
        max_slabs = max_slabs > M.C / 128 ? max_slabs : M.C / 128;
    }
    // workspace: [ranges of all sets][bounds + taps per set]
    size_t range_bytes = 0, bytes = 0;
    for (int i = 0; i < P.nsets; ++i) range_bytes += ((size_t)P.m[P.s[i].map].B * 2 * sizeof(unsigned) + 15) / 16 * 16;
    bytes = range_bytes;
    for (int i = 0; i < P.nsets; ++i) {
        bytes += (size_t)P.s[i].R * (sizeof(int4) + 32 * sizeof(TapEntry));
        max_R = max_R > P.s[i].R ? max_R : P.s[i].R;
    }
    static bool pool_ready = false;      // keep freed workspace cached in the stream-ordered pool across synchronisations
    if (!pool_ready) {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_ready = true;
    }
    char *ws = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&ws, bytes, stream);
    if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "crop backward: workspace (%zu B): %s", bytes, cudaGetErrorString(e)); return FI_ERR_CUDA; }
    e = cudaMemsetAsync(ws, 0xFF, range_bytes, stream);
    if (e != cudaSuccess) { cudaFreeAsync(ws, stream); set_error(FI_ERR_CUDA, "crop backward: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    char *p = ws;
    for (int i = 0; i < P.nsets; ++i) {
        P.s[i].range = reinterpret_cast<unsigned *>(p);
        p += ((size_t)P.m[P.s[i].map].B * 2 * sizeof(unsigned) + 15) / 16 * 16;
    }
    for (int i = 0; i < P.nsets; ++i) {
        P.s[i].bounds = reinterpret_cast<int4 *>(p); p += (size_t)P.s[i].R * sizeof(int4);
        P.s[i].taps = reinterpret_cast<TapEntry *>(p); p += (size_t)P.s[i].R * 32 * sizeof(TapEntry);
    }
    int rc = ok();
    if (max_R > 0) {
        tile_prep_kernel<<<dim3(ceil_div(max_R, 8), P.nsets), 256, 0, stream>>>(P);
        rc = check_launch("crop backward[prep]");
    }
    if (rc == FI_OK && tiles > 0) {
        const dim3 grid((unsigned)tiles, ceil_div(max_slabs, 2));
        if (exact) bwd_smem_tile_kernel<true><<<grid, 64, 0, stream>>>(P);
        else bwd_smem_tile_kernel<false><<<grid, 64, 0, stream>>>(P);
        rc = check_launch("crop backward[tile]");
    }
    cudaFreeAsync(ws, stream);
    return rc;
}
