// RoIAlign backward, tile-owner form: the dense gradient maps are cut into 4x8-pixel tiles, every tile is
// accumulated in SHARED MEMORY by the one warp that owns it (per 128-channel slab) and written to HBM exactly once.
//
// Why: the reference (crop_and_resize_kernel.cu:84-165) zero-fills the map and issues 4 atomics per crop element.
// The vector-reduction rewrite of that formulation (roi_align.cu) is bound by L2 reduction throughput (~3.5 TB/s of
// RED payload, ~5 GB per C2 step) plus a 1.5 GB zero fill -- 37 % of the HBM roofline.  Here the only HBM traffic is
// the algorithmic one: every crop gradient is read (once, plus re-reads by neighbouring tiles that L2 absorbs) and
// every map pixel is written once; no memset, no atomics, no read-modify-write in L2/HBM (ncu: 4.40 GB for 4.06 GB).
//
// Launches of one backward (all crop sets / maps of a Dev.forward pass): tile_prep (thread per box: sample geometry,
// footprint bounds, per-image index range, list of degenerate boxes) -> tile_collapse (degenerate boxes only, see below)
// -> bin_enumerate + bin_accumulate (default), or the fused bwd_smem_tile_kernel (FI_BWD_TILE=fused).
//
// What a tile's work consists of (both forms; the two-kernel form splits it after "expand", see further down):
//   scan     the boxes of the tile's image (index range from prep), 32 at a time: one 16 B record (bounds, image) per lane,
//            footprint-bounds test, ballot-compacted IN BOX ORDER into a hit list; next chunk's records in flight;
//   expand   a batch of up to 32 hit boxes, lane i <-> box i: each lane loads its box's geometry record and finds the crop
//            rows / columns whose taps touch the tile (positions are monotone: contiguous ranges); a warp scan of the
//            rectangle sizes orders the batch's samples as one (box, crop row, crop column) stream; lane l then describes
//            sample g0 + l of that stream (binary search of its box by shuffle, taps from the geometry with the forward's
//            fp32 operations, in-tile flags, pre-multiplied tap weights, gradient row) and appends it to the queue;
//   drain    the queue is consumed in order, gradient rows requested into L2 a window ahead, 8 512-byte loads in flight
//            ahead of the adds (double-buffered in registers); each tap is one conflict-free LDS.128 / add / STS.128 on
//            the warp's 16 KB accumulator (tile x 128-channel slab);
//   store    the tile is streamed out, 512 B per warp instruction.
// Because every pixel is summed by ONE warp in the order (box, crop row, crop column, TL->TR->BL->BR) the result is
// run-to-run deterministic; with EXACT arithmetic (un-fused fp32 mul then add, crop_and_resize.c:241-247) it is
// bit-identical to the reference's serial CPU loop (crop_and_resize.c:190-250).  The default mode uses packed FMAs
// (fma.rn.f32x2 -> FFMA2) with pre-multiplied weights: same order, one rounding fewer per contribution.
//
// Degenerate boxes.  The reference zero-pads its RoI lists (lib/layers.py:413,427): every all-zero box puts ALL its
// P*P samples on pixel (0,0) of its image, so one tile per image would have to stream hundreds of crops through a
// single warp (measured: a 1.6 ms serial tail).  In the default mode boxes whose footprint is at most 2x2 pixels are
// therefore pre-reduced by the collapse kernel (one CTA per box, all warps loading in parallel) into 4 corner sums
// that the tile kernel consumes as 4 unit-weight samples.  EXACT mode keeps the reference's sample-by-sample order.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "roi_align_bwd_tile.cuh"

#ifndef FI_SCAN_DEPTH
#define FI_SCAN_DEPTH 1     // chunks of box records in flight ahead of the scan (3 measured 1.58 ms vs 1.52 ms: register pressure)
#endif
#ifndef FI_TILE_MINB
#define FI_TILE_MINB 6      // resident CTAs per SM the 4x8 kernel is compiled for (register cap)
#endif

namespace fi {
namespace tile {

// ---- prep: one thread per box (all sets in one launch): footprint bounds, per-image index range, degenerate list ----
__global__ void __launch_bounds__(128) tile_prep_kernel(const TParams P) {
    const TSet &S = P.s[blockIdx.y];
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (S.R_dev ? min(*S.R_dev, S.R) : S.R)) return;
    const int B = P.m[S.map].B, H = P.m[S.map].H, W = P.m[S.map].W;
    const int b = S.box_ind[r];
    const float4 bx = S.boxes[r];                  // (y1, x1, y2, x2)
    const float4 geom = make_float4(geom_base(bx.x, bx.z, H, S.ph), axis_step(bx.x, bx.z, H, S.ph), geom_base(bx.y, bx.w, W, S.pw),
                                    axis_step(bx.y, bx.w, W, S.pw));
    S.geom[r] = geom;
    int ymin = 1 << 30, ymax = -(1 << 30), xmin = 1 << 30, xmax = -(1 << 30);
    if (b >= 0 && b < B) {                         // others are skipped like crop_and_resize_kernel.cu:34-38
        for (int k = 0; k < S.ph; ++k) {
            const Tap t = geom_tap(geom.x, geom.y, k, H);
            if (t.lo != kNoTap) { ymin = min(ymin, t.lo); ymax = max(ymax, t.hi); }
        }
        for (int k = 0; k < S.pw; ++k) {
            const Tap t = geom_tap(geom.z, geom.w, k, W);
            if (t.lo != kNoTap) { xmin = min(xmin, t.lo); xmax = max(xmax, t.hi); }
        }
    }
    const bool live = ymin <= ymax && xmin <= xmax;
    int4 rec = make_int4(0, 0, -1, 0);
    if (live) {
        const bool degenerate = (ymax - ymin <= 1) && (xmax - xmin <= 1) && S.ph * S.pw > 4;
        rec = make_int4(ymin | (ymax << 16), xmin | (xmax << 16), b, degenerate ? 1 : 0);
        atomicMin(S.range + 2 * b, (unsigned)r);
        atomicMin(S.range + 2 * b + 1, ~(unsigned)r);
        if (degenerate && P.collapse) P.deg_list[1 + atomicAdd(P.deg_list, 1)] = ((int)blockIdx.y << 24) | r;
    }
    S.rec[r] = rec;
}

// ---- collapse: corner sums of degenerate boxes (footprint <= 2x2 pixels), one CTA per box -------------------------
// coll[r][corner][c] = sum over the box's samples of (tap weight on that corner) * (grads + grads2); corner =
// 2 * (y - ymin) + (x - xmin).  8 warps split the samples, lanes hold float4 of a 128-channel slab.
__global__ void __launch_bounds__(256, 2) tile_collapse_kernel(const TParams P) {
    __shared__ float4 part[8][4][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int count = P.deg_list[0];
    for (int it = blockIdx.x; it < count; it += gridDim.x) {
        const int code = P.deg_list[1 + it];
        const TSet &S = P.s[code >> 24];
        const int r = code & 0xffffff;
        const int H = P.m[S.map].H, W = P.m[S.map].W, C = P.m[S.map].C;
        const int4 rec = S.rec[r];
        const int ymin = (short)(rec.x & 0xffff), xmin = (short)(rec.y & 0xffff);
        const float4 geom = S.geom[r];
        const long grow = S.src_row ? (long)S.src_row[r] : (long)r;
        const int pp = S.ph * S.pw;
        // Two 128-channel slabs and 8 samples per warp per round: the kernel is one CTA per box and as long as its chain of
        // dependent DRAM round trips (0.048 ms with one slab and 4 samples per round).  Every warp still sums its samples
        // w, w + 8, w + 16, ... in that order and the 8 partial sums meet in the same order: the bits do not change.
        for (int slab = 0; slab * 128 < C; slab += 2) {
            const int coff = slab * 128 + lane * 4;
            const bool two = (slab + 1) * 128 < C;
            float4 acc[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[h][c] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s0 = w; s0 < pp; s0 += 64) {                      // 8 samples per warp per round, loads first
                float4 g[8][2];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int sc = min(s0 + 8 * q, pp - 1);
                    const int i = sc / S.pw, j = sc - i * S.pw;
                    const float *p1 = S.grads + ((grow * S.ph + i) * S.pw + j) * (long)C + coff;
                    g[q][0] = __ldcs(reinterpret_cast<const float4 *>(p1));
                    g[q][1] = two ? __ldcs(reinterpret_cast<const float4 *>(p1 + 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (S.grads2) {
                        const float *p2 = S.grads2 + (((long)r * S.ph + i) * S.pw + j) * (long)C + coff;
                        g[q][0] = add_rn4(g[q][0], __ldcs(reinterpret_cast<const float4 *>(p2)));
                        if (two) g[q][1] = add_rn4(g[q][1], __ldcs(reinterpret_cast<const float4 *>(p2 + 128)));
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {                          // the taps are formed here, after the loads: fewer live registers
                    const int s = s0 + 8 * q;
                    if (s >= pp) break;
                    const int i = s / S.pw, j = s - i * S.pw;
                    const Tap tyq = geom_tap(geom.x, geom.y, i, H), txq = geom_tap(geom.z, geom.w, j, W);
                    if (tyq.lo == kNoTap || txq.lo == kNoTap) continue;
                    const float wy1 = tyq.frac, wy0 = 1.f - wy1, wx1 = txq.frac, wx0 = 1.f - wx1;
                    const int cy0 = tyq.lo - ymin, cy1 = tyq.hi - ymin, cx0 = txq.lo - xmin, cx1 = txq.hi - xmin;   // each 0 or 1
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int cy = c >> 1, cx = c & 1;
                        float wgt = 0.f;                                   // coinciding taps: the `hi` one carries weight 0
                        if (cy == cy0 && cx == cx0) wgt += wy0 * wx0;
                        if (cy == cy0 && cx == cx1 && cx1 != cx0) wgt += wy0 * wx1;
                        if (cy == cy1 && cy1 != cy0 && cx == cx0) wgt += wy1 * wx0;
                        if (cy == cy1 && cy1 != cy0 && cx == cx1 && cx1 != cx0) wgt += wy1 * wx1;
                        acc[0][c] = fma4(g[q][0], wgt, acc[0][c]);
                        acc[1][c] = fma4(g[q][1], wgt, acc[1][c]);
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1 && !two) break;
#pragma unroll
                for (int c = 0; c < 4; ++c) part[w][c][lane] = acc[h][c];
                __syncthreads();
                if (w < 4) {
                    float4 v = part[0][w][lane];
#pragma unroll
                    for (int q = 1; q < 8; ++q) v = add_rn4(v, part[q][w][lane]);
                    *reinterpret_cast<float4 *>(S.coll + ((long)r * 4 + w) * C + coff + h * 128) = v;
                }
                __syncthreads();
            }
        }
    }
}

template <int TY, int TX>
struct __align__(16) WarpSmemT {
    float tile[TY * TX * 128];         // this warp's slab of the tile: [pixel][128 channels]
    uint4 qa[kQ];                      // sample queue: (row of grads, row of grads2, pk, -)
    float4 qw[kQ];                     //               tap weights (TL, TR, BL, BR); EXACT: (1 - fy, fy, 1 - fx, fx)
    int list[64];                      // hit boxes, in index order
};
// pk: in-tile flags TL|TR|BL|BR (bits 0-3), kCollapsed (bit 4), byte offset of the TL pixel inside the warp's tile (bits 9.., signed)
constexpr int kCollapsed = 1 << 4;

// One queued sample: up to 4 read-modify-writes of this lane's float4 in the tile.  Flags / offsets are warp-uniform.
struct TapVals { float4 v0, v1, v2, v3; };

template <int TX>
__device__ __forceinline__ TapVals taps_load(const float *tile_lane, int pk) {
    const float4 *tp = reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(tile_lane) + (pk & ~0x1ff));
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    TapVals t;
    t.v0 = z; t.v1 = z; t.v2 = z; t.v3 = z;
    if (pk & 1) t.v0 = tp[0];
    if (pk & 2) t.v1 = tp[32];
    if (pk & 4) t.v2 = tp[TX * 32];
    if (pk & 8) t.v3 = tp[TX * 32 + 32];
    return t;
}
template <bool EXACT>
__device__ __forceinline__ void taps_add(TapVals &t, float4 w, float4 g) {
    if (EXACT) {                                                                     // crop_and_resize.c:241-247
        const float4 dtop = mul_rn4(w.x, g), dbot = mul_rn4(w.y, g);
        t.v0 = add_rn4(t.v0, mul_rn4(w.z, dtop));
        t.v1 = add_rn4(t.v1, mul_rn4(w.w, dtop));
        t.v2 = add_rn4(t.v2, mul_rn4(w.z, dbot));
        t.v3 = add_rn4(t.v3, mul_rn4(w.w, dbot));
    } else {
        t.v0 = fma4(g, w.x, t.v0);
        t.v1 = fma4(g, w.y, t.v1);
        t.v2 = fma4(g, w.z, t.v2);
        t.v3 = fma4(g, w.w, t.v3);
    }
}
template <int TX>
__device__ __forceinline__ void taps_store(float *tile_lane, int pk, const TapVals &t) {
    float4 *tp = reinterpret_cast<float4 *>(reinterpret_cast<char *>(tile_lane) + (pk & ~0x1ff));
    if (pk & 1) tp[0] = t.v0;
    if (pk & 2) tp[32] = t.v1;
    if (pk & 4) tp[TX * 32] = t.v2;
    if (pk & 8) tp[TX * 32 + 32] = t.v3;
}

// One queued sample: up to 4 read-modify-writes of this lane's float4 in the tile.  Flags / offsets are warp-uniform.  The
// four taps are distinct pixels (coinciding ones were dropped by the emitter): read all, then add, then write.
template <bool EXACT, int TX>
__device__ __forceinline__ void apply_sample(float *tile_lane, int pk, float4 w, float4 g) {
    TapVals t = taps_load<TX>(tile_lane, pk);
    taps_add<EXACT>(t, w, g);
    taps_store<TX>(tile_lane, pk, t);
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Consume the queue in order.  Groups of U samples; the loads of group i+1 are in flight while group i is added.
// qn is a multiple of U: the emitter pads the tail with flag-less copies of the last sample.
template <bool EXACT, bool DUAL, int TX>
__device__ __forceinline__ void drain(const uint4 *qa, const float4 *qw, int qn, const float *__restrict__ G1, const float *__restrict__ G2,
                                      const float *__restrict__ COLL, int C, float *tile_lane) {
    constexpr int U = DUAL ? 4 : 8;
    constexpr int UB = DUAL ? U : 1;
    constexpr int kAhead = 24;                         // samples kept on their way into L2 ahead of the register loads
    const int lane = threadIdx.x & 31;
    float4 a0[U], a1[U], b0[UB], b1[UB];
    // 8 samples per call: lane -> (sample e0 + lane / 4, 128-byte line lane % 4 of the warp's 512-byte slab run).  Register
    // double-buffering alone keeps ~4 KB per warp in flight -- at 12 warps / SM that is latency-bound against HBM -- so the
    // lines are requested into L2 a window ahead and the register loads then see L2 latency.
    auto prefetch8 = [&](int e0) {
        const int idx = e0 + (lane >> 2);
        if (idx < qn) {
            const uint4 e = qa[idx];
            const bool coll = (!EXACT) && (e.z & kCollapsed);
            const float *src = (coll ? COLL : G1) - lane * 4 + (lane & 3) * 32;
            prefetch_l2(src + (size_t)e.x * C);
            if (DUAL && !coll) prefetch_l2(G2 - lane * 4 + (lane & 3) * 32 + (size_t)e.y * C);
        }
    };
    auto load = [&](int g0, float4(&a)[U], float4(&b)[UB]) {
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const uint4 e = qa[g0 + k];
            const bool coll = (!EXACT) && (e.z & kCollapsed);
            const float *src = coll ? COLL : G1;
            a[k] = __ldcg(reinterpret_cast<const float4 *>(src + (size_t)e.x * C));
            if (DUAL) b[k] = coll ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcg(reinterpret_cast<const float4 *>(G2 + (size_t)e.y * C));
        }
    };
    auto process = [&](int g0, float4(&a)[U], float4(&b)[UB]) {
#pragma unroll
        for (int k = 0; k < U; k += 2) {
            const int pkA = (int)qa[g0 + k].z, pkB = (int)qa[g0 + k + 1].z;
            const float4 wA = qw[g0 + k], wB = qw[g0 + k + 1];
            const float4 gA = DUAL ? add_rn4(a[k], b[k]) : a[k];
            const float4 gB = DUAL ? add_rn4(a[k + 1], b[k + 1]) : a[k + 1];
            apply_sample<EXACT, TX>(tile_lane, pkA, wA, gA);
            apply_sample<EXACT, TX>(tile_lane, pkB, wB, gB);
        }
    };
#pragma unroll
    for (int e0 = U; e0 < U + kAhead + (DUAL ? 8 : 0); e0 += 8) prefetch8(e0);
    load(0, a0, b0);
    for (int g0 = 0; g0 < qn; g0 += 2 * U) {
        const bool more = g0 + U < qn;
        prefetch8(g0 + U + kAhead + (DUAL ? 8 : 0));
        if (more) load(g0 + U, a1, b1);
        process(g0, a0, b0);
        if (more) {
            if (!DUAL) prefetch8(g0 + U + kAhead + 8);
            if (g0 + 2 * U < qn) load(g0 + 2 * U, a0, b0);
            process(g0 + U, a1, b1);
        }
    }
}

template <bool EXACT, int TY, int TX>
__device__ __noinline__ void drain_any(uint4 *qa, float4 *qw, int qn, const float *G1, const float *G2, const float *COLL, int C,
                                       float *tile_lane, bool zero) {
    const int lane = threadIdx.x & 31;
    if (zero) {                                        // first use of the tile
#pragma unroll
        for (int p = 0; p < TY * TX; ++p) *reinterpret_cast<float4 *>(tile_lane + p * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    const int qnp = (qn + 7) & ~7;                     // pad to whole groups: flag-less copies of the last sample (valid address)
    if (qn + lane < qnp) {
        uint4 e = qa[qn - 1];
        e.z &= kCollapsed;
        qa[qn + lane] = e;
        qw[qn + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    if (G2) drain<EXACT, true, TX>(qa, qw, qnp, G1, G2, COLL, C, tile_lane);
    else drain<EXACT, false, TX>(qa, qw, qnp, G1, nullptr, COLL, C, tile_lane);
    __syncwarp();
}

// grid (total tiles of all maps, ceil(slabs / 2)), 64 threads
template <bool EXACT, int TY, int TX, int MINB>
__global__ void __launch_bounds__(64, MINB) bwd_smem_tile_kernel(const TParams P) {
    __shared__ WarpSmemT<TY, TX> smem[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int t = (int)gridDim.x - 1 - (int)blockIdx.x, mi = 0;          // heaviest (coarsest-level) tiles first, see bin_accumulate_kernel
    while (mi + 1 < P.nmaps && t >= P.m[mi + 1].first_tile) ++mi;
    const TMap &M = P.m[mi];
    const int H = M.H, W = M.W, C = M.C;
    const int slab = blockIdx.y * 2 + w;
    if (slab * 128 >= C) return;                       // the warps of a CTA never synchronise with each other
    t -= M.first_tile;
    const int tiles_x = ceil_div(W, TX), tiles_y = ceil_div(H, TY);
    const int per_img = tiles_x * tiles_y;
    const int b = t / per_img;
    t -= b * per_img;
    const int tyi = t / tiles_x, txi = t - tyi * tiles_x;
    const int X0 = txi * TX, Y0 = tyi * TY;
    const int coff = slab * 128 + lane * 4;
    WarpSmemT<TY, TX> &ws = smem[w];
    float *tile_lane = ws.tile + lane * 4;
    const unsigned lt = (1u << lane) - 1u;
    bool dirty = false;                                // tile zeroed lazily, on the first queued sample

    for (int si = M.set_begin; si < M.set_end; ++si) {
        const TSet &S = P.s[si];
        const unsigned first = S.range[2 * b], last = ~S.range[2 * b + 1];
        if (first > last) continue;                    // no box of this set lives in image b
        const int ph = S.ph, pw = S.pw;
        const float *G1 = S.grads + coff, *G2 = S.grads2 ? S.grads2 + coff : nullptr, *COLL = S.coll + coff;
        int qn = 0, nl = 0;
        int4 nxt = __ldg(S.rec + min((first & ~31u) + lane, last));
#if FI_SCAN_DEPTH >= 3
        int4 nxt2 = __ldg(S.rec + min((first & ~31u) + 32 + lane, last));
        int4 nxt3 = __ldg(S.rec + min((first & ~31u) + 64 + lane, last));
#endif
        for (unsigned base = first & ~31u; base <= last; base += 32) {
            // ---- scan: ordered compaction of the boxes whose footprint overlaps the tile (the next 3 chunks' records in flight)
            const int4 rec = nxt;
            const bool last_chunk = base + 32 > last;
#if FI_SCAN_DEPTH >= 3
            nxt = nxt2;
            nxt2 = nxt3;
            if (base + 96 <= last) nxt3 = __ldg(S.rec + min(base + 96 + lane, last));
#else
            if (!last_chunk) nxt = __ldg(S.rec + min(base + 32 + lane, last));
#endif
            const unsigned r = base + lane;
            const int ymin = (short)(rec.x & 0xffff), ymax = (short)(rec.x >> 16);
            const int xmin = (short)(rec.y & 0xffff), xmax = (short)(rec.y >> 16);
            const bool hit = r >= first && r <= last && rec.z == b && ymin <= Y0 + TY - 1 && ymax >= Y0 && xmin <= X0 + TX - 1 && xmax >= X0;
            const unsigned hm = __ballot_sync(0xffffffffu, hit);
            if (hit) ws.list[nl + __popc(hm & lt)] = (int)r;
            nl += __popc(hm);
            __syncwarp();
            // ---- expand: a batch of up to 32 hit boxes, lane i <-> hit i
            while (nl >= 32 || (last_chunk && nl > 0)) {
                const int nb = min(nl, 32);
                const bool mine = lane < nb;
                const int rr = mine ? ws.list[lane] : 0;
                float4 geom = make_float4(0.f, 0.f, 0.f, 0.f);
                int4 hr = make_int4(0, 0, 0, 0);
                int grow = 0;
                if (mine) {
                    geom = __ldg(S.geom + rr);
                    hr = __ldg(S.rec + rr);
                    grow = S.src_row ? __ldg(S.src_row + rr) : rr;
                }
                __syncwarp();
                if (nl > 32) ws.list[lane] = ws.list[32 + lane];      // keep the overflow for the next batch
                nl -= nb;
                __syncwarp();
                // crop rows / columns of MY box whose taps touch the tile (positions are monotone in k: contiguous ranges)
                unsigned ymask = 0, xmask = 0;
                for (int k = 0; k < ph; ++k) {
                    const Tap tp = geom_tap(geom.x, geom.y, k, H);
                    if ((tp.lo >= Y0 && tp.lo < Y0 + TY) || (tp.hi >= Y0 && tp.hi < Y0 + TY)) ymask |= 1u << k;
                }
                for (int k = 0; k < pw; ++k) {
                    const Tap tp = geom_tap(geom.z, geom.w, k, W);
                    if ((tp.lo >= X0 && tp.lo < X0 + TX) || (tp.hi >= X0 && tp.hi < X0 + TX)) xmask |= 1u << k;
                }
                const bool deg = (!EXACT) && hr.w != 0;
                int iy0 = 0, ix0 = 0, nx = 1, cnt = 0;
                if (mine && ymask && xmask) {
                    iy0 = __ffs(ymask) - 1; ix0 = __ffs(xmask) - 1;
                    nx = 32 - __clz(xmask) - ix0;
                    cnt = deg ? 4 : (32 - __clz(ymask) - iy0) * nx;       // <= 256 samples of this box may touch the tile
                }
                const unsigned inv = (65536u + nx - 1) / nx;              // s / nx == (s * inv) >> 16 for s < 256, nx <= 16
                const int rect = iy0 | (ix0 << 4) | (nx << 8) | (deg ? 1 << 16 : 0);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                // ---- emit: lane l describes sample g0 + l of the batch's (box, crop row, crop column)-ordered stream
                for (int g0 = 0; g0 < total;) {
                    const int room = min(min(32, total - g0), kQ - qn);
                    const int g = g0 + lane;
                    int j = 0;
#pragma unroll
                    for (int st = 16; st >= 1; st >>= 1) {
                        const int probe = __shfl_sync(0xffffffffu, incl, min(j + st - 1, 31));
                        if (probe <= g) j += st;
                    }
                    j = min(j, 31);
                    const int s = g - (__shfl_sync(0xffffffffu, incl, j) - __shfl_sync(0xffffffffu, cnt, j));
                    const float gby = __shfl_sync(0xffffffffu, geom.x, j), gsy = __shfl_sync(0xffffffffu, geom.y, j);
                    const float gbx = __shfl_sync(0xffffffffu, geom.z, j), gsx = __shfl_sync(0xffffffffu, geom.w, j);
                    const int jrect = __shfl_sync(0xffffffffu, rect, j);
                    const unsigned jinv = __shfl_sync(0xffffffffu, inv, j);
                    const int jgrow = __shfl_sync(0xffffffffu, grow, j), jrr = __shfl_sync(0xffffffffu, rr, j);
                    const int jry = __shfl_sync(0xffffffffu, hr.x, j), jrx = __shfl_sync(0xffffffffu, hr.y, j);
                    uint4 q = make_uint4(0u, 0u, 0u, 0u);
                    float4 wts = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lane < room) {
                        if (jrect & (1 << 16)) {
                            // degenerate box: its (<= 4) corner sums come from the collapse kernel as unit-weight samples
                            const int bymin = (short)(jry & 0xffff), bymax = (short)(jry >> 16), bxmin = (short)(jrx & 0xffff), bxmax = (short)(jrx >> 16);
                            const int cy = (s & 2) ? bymax : bymin, cx = (s & 1) ? bxmax : bxmin;
                            const bool use = !((s & 2) && bymax == bymin) && !((s & 1) && bxmax == bxmin) && cy >= Y0 && cy < Y0 + TY && cx >= X0 &&
                                             cx < X0 + TX;
                            const int px = (cy - Y0) * TX + (cx - X0);
                            const int pk = use ? (1 | kCollapsed | (px * 512)) : kCollapsed;
                            q = make_uint4((unsigned)(jrr * 4 + s), 0u, (unsigned)pk, 0u);
                            wts = make_float4(1.f, 0.f, 0.f, 0.f);
                        } else {
                            const int jnx = (jrect >> 8) & 0xff;
                            const int ii = (int)(((unsigned)s * jinv) >> 16);
                            const int iy = (jrect & 15) + ii, ix = ((jrect >> 4) & 15) + (s - ii * jnx);
                            const Tap ty = geom_tap(gby, gsy, iy, H), tx = geom_tap(gbx, gsx, ix, W);
                            // a tap that coincides with its partner (integer sample position) carries weight 0: dropped
                            const bool top = ty.lo >= Y0 && ty.lo < Y0 + TY, bot = ty.hi >= Y0 && ty.hi < Y0 + TY && ty.hi != ty.lo;
                            const bool lef = tx.lo >= X0 && tx.lo < X0 + TX, rig = tx.hi >= X0 && tx.hi < X0 + TX && tx.hi != tx.lo;
                            const int f = (top && lef ? 1 : 0) | (top && rig ? 2 : 0) | (bot && lef ? 4 : 0) | (bot && rig ? 8 : 0);
                            const int px = (ty.lo - Y0) * TX + (tx.lo - X0);                              // TL pixel; may be negative
                            const int pk = f ? (f | (px * 512)) : 0;
                            q = make_uint4((unsigned)((jgrow * ph + iy) * pw + ix), (unsigned)((jrr * ph + iy) * pw + ix), (unsigned)pk, 0u);
                            const float wy0 = __fsub_rn(1.f, ty.frac), wx0 = __fsub_rn(1.f, tx.frac);      // crop_and_resize.c:241-247
                            wts = EXACT ? make_float4(wy0, ty.frac, wx0, tx.frac)
                                        : make_float4(wy0 * wx0, wy0 * tx.frac, ty.frac * wx0, ty.frac * tx.frac);
                        }
                    }
                    if (lane < room) {
                        ws.qa[qn + lane] = q;
                        ws.qw[qn + lane] = wts;
                    }
                    qn += room;
                    g0 += room;
                    if (qn == kQ) {                                    // full queue: 64 samples per drain
                        drain_any<EXACT, TY, TX>(ws.qa, ws.qw, qn, G1, G2, COLL, C, tile_lane, !dirty);
                        dirty = true;
                        qn = 0;
                    }
                }
            }
        }
        if (qn > 0) {                                   // set boundary: the queue holds rows of one set only
            drain_any<EXACT, TY, TX>(ws.qa, ws.qw, qn, G1, G2, COLL, C, tile_lane, !dirty);
            dirty = true;
        }
    }
    // ---- store the tile (or zeros), once
    __syncwarp();
    float *dst0 = M.gimg + (((long)b * H + Y0) * (long)W + X0) * C + coff;
    const long row_stride = (long)W * C;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (Y0 + TY <= H && X0 + TX <= W && !P.accumulate) {            // interior tile: straight-line stores
#pragma unroll
        for (int yy = 0; yy < TY; ++yy) {
#pragma unroll
            for (int xx = 0; xx < TX; ++xx) {
                const float4 v = dirty ? *reinterpret_cast<const float4 *>(tile_lane + (yy * TX + xx) * 128) : zero4;
                __stcs(reinterpret_cast<float4 *>(dst0 + yy * row_stride + xx * C), v);
            }
        }
    } else {
        for (int yy = 0; yy < TY && Y0 + yy < H; ++yy) {
            for (int xx = 0; xx < TX && X0 + xx < W; ++xx) {
                float *dst = dst0 + yy * row_stride + xx * C;
                float4 v = dirty ? *reinterpret_cast<const float4 *>(tile_lane + (yy * TX + xx) * 128) : zero4;
                if (P.accumulate) v = add_rn4(*reinterpret_cast<const float4 *>(dst), v);
                __stcs(reinterpret_cast<float4 *>(dst), v);
            }
        }
    }
}


// =====================================================================================================================
// Two-kernel form (default).  The single kernel above exposes ~12 dependent memory round trips per tile (ranges, box
// records chunk by chunk, box geometry, first gradient loads, ... per crop set) while only 12 warps fit next to their
// 16 KB accumulators: ncu shows 16 % occupancy, 0.34 IPC per scheduler, every stall a scoreboard wait.  Split in two:
//   bin_enumerate_kernel   one warp per TILE (not per slab), no accumulator, ~2 KB shared memory per warp -> many resident
//                          warps hide those round trips; scan and expansion as above, but the described samples go to a
//                          per-tile list in global memory (64-entry chunks handed out by one atomic, chained per tile);
//   bin_accumulate_kernel  one warp per (tile, slab) streams its list: chunk -> shared memory, gradient rows requested
//                          into L2 a window ahead, 8 register loads in flight, shared-memory adds, one store of the tile.
// Two-source sets (mask head + critic gradient of the same crop) queue TWO entries per sample: the first only loads
// (kDefer2) and is added to the second's gradient before the taps are applied -- (g1 + g2) first, like autograd's
// accumulation in the reference.  Every set's segment is padded to a multiple of 8 entries so pairs never straddle a group.
// pk2: tap flags (bits 0-3), kDefer2 (bit 4), source index 3 * set + {grads, grads2, collapse rows} (bits 5-10),
//      TL pixel index inside the tile, signed (bits 11..).

struct __align__(16) EnumSmem {
    float4 qw[kChunk];
    uint2 qa[kChunk];
    int list[64];
};

// 128 threads: warp w of block b owns tile 4 * b + w
template <bool EXACT, int TY, int TX>
__global__ void __launch_bounds__(128, 8) bin_enumerate_kernel(const TParams P) {
    // (Tried in round 2: a CTA-level pre-filter of the boxes against the bounding box of the CTA's 4 tiles, to save the per-tile scan
    // of all boxes of the image -- a third of the instructions.  No gain, 137 vs 128 us: the kernel is bound by the ~8 dependent
    // global round trips per (tile, set) at 24 resident warps per SM, not by its instruction count.)
    __shared__ EnumSmem smem[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if ((int)blockIdx.x * 4 + w >= P.bin.total_tiles) return;      // warps never synchronise with each other
    const int tile_id = P.bin.total_tiles - 1 - ((int)blockIdx.x * 4 + w);   // heaviest (coarsest-level) tiles first, see bin_accumulate_kernel
    int t = tile_id, mi = 0;
    while (mi + 1 < P.nmaps && t >= P.m[mi + 1].first_tile) ++mi;
    const TMap &M = P.m[mi];
    const int H = M.H, W = M.W;
    t -= M.first_tile;
    const int tiles_x = ceil_div(W, TX), tiles_y = ceil_div(H, TY);
    const int per_img = tiles_x * tiles_y;
    const int b = t / per_img;
    t -= b * per_img;
    const int tyi = t / tiles_x, txi = t - tyi * tiles_x;
    const int X0 = txi * TX, Y0 = tyi * TY;
    EnumSmem &ws = smem[w];
    const unsigned lt = (1u << lane) - 1u;
    int qn = 0, qtot = 0, prev_chunk = -1;
    int stored = 0;
    bool overflow = false;
    unsigned last_row = 0;
    int last_src = 0;

    auto flush = [&]() {                                // hand the queued entries (<= 64) to the tile's list
        int c = 0;
        if (lane == 0) {
            c = atomicAdd(P.bin.cursor, 1);
            if (c < P.bin.pool) {
                if (prev_chunk < 0) P.bin.tile_head[tile_id] = c;
                else P.bin.chunk_next[prev_chunk] = c;
            } else {
                atomicExch(P.bin.cursor + 1, 1);        // overflow (only possible with a caller's own, too small, entry bound)
            }
        }
        c = __shfl_sync(0xffffffffu, c, 0);
        __syncwarp();
        if (c < P.bin.pool) {                           // the default pool is sized for 4 tiles per sample: always true then
            P.bin.qa[(size_t)c * kChunk + lane] = ws.qa[lane];
            P.bin.qa[(size_t)c * kChunk + 32 + lane] = ws.qa[32 + lane];
            P.bin.qw[(size_t)c * kChunk + lane] = ws.qw[lane];
            P.bin.qw[(size_t)c * kChunk + 32 + lane] = ws.qw[32 + lane];
            prev_chunk = c;
            stored += qn;
        } else {
            overflow = true;                            // the tile's list ends at the last chunk that fitted
        }
        qn = 0;
        __syncwarp();
    };

    auto pad_entries = [&](int pad) {                   // tap-less entries re-reading the last row (a valid address); pad <= 7
        __syncwarp();
        if (lane < pad) {                               // qn is even / a multiple of 8 away from 64 here: the pad always fits
            ws.qa[qn + lane] = make_uint2(last_row, (unsigned)last_src);
            ws.qw[qn + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        qn += pad;
        qtot += pad;
        __syncwarp();
        if (qn == kChunk) flush();
    };

    for (int si = M.set_begin; si < M.set_end; ++si) {
        const TSet &S = P.s[si];
        const unsigned first = S.range[2 * b], last = ~S.range[2 * b + 1];
        if (first > last) continue;                    // no box of this set lives in image b
        const int ph = S.ph, pw = S.pw;
        const bool dual = S.grads2 != nullptr;
        int nl = 0;
        int4 nxt = __ldg(S.rec + min((first & ~31u) + lane, last));
        for (unsigned base = first & ~31u; base <= last; base += 32) {
            // ---- scan: ordered compaction of the boxes whose footprint overlaps the tile (next chunk's records in flight)
            const int4 rec = nxt;
            const bool last_chunk = base + 32 > last;
            if (!last_chunk) nxt = __ldg(S.rec + min(base + 32 + lane, last));
            const unsigned r = base + lane;
            const int ymin = (short)(rec.x & 0xffff), ymax = (short)(rec.x >> 16);
            const int xmin = (short)(rec.y & 0xffff), xmax = (short)(rec.y >> 16);
            const bool hit = r >= first && r <= last && rec.z == b && ymin <= Y0 + TY - 1 && ymax >= Y0 && xmin <= X0 + TX - 1 && xmax >= X0;
            const unsigned hm = __ballot_sync(0xffffffffu, hit);
            if (hit) ws.list[nl + __popc(hm & lt)] = (int)r;
            nl += __popc(hm);
            __syncwarp();
            // ---- expand: a batch of up to 32 hit boxes, lane i <-> hit i
            while (nl >= 32 || (last_chunk && nl > 0)) {
                const int nb = min(nl, 32);
                const bool mine = lane < nb;
                const int rr = mine ? ws.list[lane] : 0;
                float4 geom = make_float4(0.f, 0.f, 0.f, 0.f);
                int4 hr = make_int4(0, 0, 0, 0);
                int grow = 0;
                if (mine) {
                    geom = __ldg(S.geom + rr);
                    hr = __ldg(S.rec + rr);
                    grow = S.src_row ? __ldg(S.src_row + rr) : rr;
                }
                __syncwarp();
                if (nl > 32) ws.list[lane] = ws.list[32 + lane];      // keep the overflow for the next batch
                nl -= nb;
                __syncwarp();
                // crop rows / columns of MY box whose taps touch the tile (positions are monotone in k: contiguous ranges)
                // floor(pos) or ceil(pos) in [T0, T0 + T)  <=>  T0 - 1 < pos < T0 + T  (pos inside the map): the same positions as
                // geom_tap forms, compared as floats instead of being floored / ceiled
                unsigned ymask = 0, xmask = 0;
                {
                    const float lo_y = (float)(Y0 - 1), hi_y = (float)(Y0 + TY), lo_x = (float)(X0 - 1), hi_x = (float)(X0 + TX);
                    const float top_y = (float)(H - 1), top_x = (float)(W - 1);
                    for (int k = 0; k < ph; ++k) {
                        const float pos = __fadd_rn(geom.x, __fmul_rn((float)k, geom.y));
                        if (pos > lo_y && pos < hi_y && !(pos < 0.f || pos > top_y)) ymask |= 1u << k;
                    }
                    for (int k = 0; k < pw; ++k) {
                        const float pos = __fadd_rn(geom.z, __fmul_rn((float)k, geom.w));
                        if (pos > lo_x && pos < hi_x && !(pos < 0.f || pos > top_x)) xmask |= 1u << k;
                    }
                }
                const bool deg = (!EXACT) && hr.w != 0;
                int iy0 = 0, ix0 = 0, nx = 1, cnt = 0;
                if (mine && ymask && xmask) {
                    iy0 = __ffs(ymask) - 1; ix0 = __ffs(xmask) - 1;
                    nx = 32 - __clz(xmask) - ix0;
                    cnt = deg ? 4 : (32 - __clz(ymask) - iy0) * nx * (dual ? 2 : 1);      // <= 512 entries
                }
                const unsigned inv = (65536u + nx - 1) / nx;              // s / nx == (s * inv) >> 16 for s < 256, nx <= 16
                const int rect = iy0 | (ix0 << 4) | (nx << 8) | (deg ? 1 << 16 : 0);
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                // ---- emit: lane l describes entry g0 + l of the batch's (box, crop row, crop column)-ordered stream
                for (int g0 = 0; g0 < total;) {
                    const int room = min(min(32, total - g0), kChunk - qn);
                    const int g = g0 + lane;
                    int j = 0;
#pragma unroll
                    for (int st = 16; st >= 1; st >>= 1) {
                        const int probe = __shfl_sync(0xffffffffu, incl, min(j + st - 1, 31));
                        if (probe <= g) j += st;
                    }
                    j = min(j, 31);
                    const int s = g - (__shfl_sync(0xffffffffu, incl, j) - __shfl_sync(0xffffffffu, cnt, j));
                    const float gby = __shfl_sync(0xffffffffu, geom.x, j), gsy = __shfl_sync(0xffffffffu, geom.y, j);
                    const float gbx = __shfl_sync(0xffffffffu, geom.z, j), gsx = __shfl_sync(0xffffffffu, geom.w, j);
                    const int jrect = __shfl_sync(0xffffffffu, rect, j);
                    const unsigned jinv = __shfl_sync(0xffffffffu, inv, j);
                    const int jgrow = __shfl_sync(0xffffffffu, grow, j), jrr = __shfl_sync(0xffffffffu, rr, j);
                    const int jry = __shfl_sync(0xffffffffu, hr.x, j), jrx = __shfl_sync(0xffffffffu, hr.y, j);
                    uint2 q = make_uint2(0u, 0u);
                    float4 wts = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lane < room) {
                        if (jrect & (1 << 16)) {
                            // degenerate box: its (<= 4) corner sums come from the collapse kernel as unit-weight samples
                            const int bymin = (short)(jry & 0xffff), bymax = (short)(jry >> 16), bxmin = (short)(jrx & 0xffff), bxmax = (short)(jrx >> 16);
                            const int cy = (s & 2) ? bymax : bymin, cx = (s & 1) ? bxmax : bxmin;
                            const bool use = !((s & 2) && bymax == bymin) && !((s & 1) && bxmax == bxmin) && cy >= Y0 && cy < Y0 + TY && cx >= X0 &&
                                             cx < X0 + TX;
                            const int px = use ? (cy - Y0) * TX + (cx - X0) : 0;
                            q = make_uint2((unsigned)(jrr * 4 + s), (unsigned)((use ? 1 : 0) | ((3 * si + 2) << 5) | (px * 2048)));
                            wts = EXACT ? make_float4(1.f, 0.f, 1.f, 0.f) : make_float4(1.f, 0.f, 0.f, 0.f);
                        } else {
                            const int smp = dual ? (s >> 1) : s;
                            const bool defer = dual && !(s & 1);           // first of the pair: the scattered gradient, load only
                            const int jnx = (jrect >> 8) & 0xff;
                            const int ii = (int)(((unsigned)smp * jinv) >> 16);
                            const int iy = (jrect & 15) + ii, ix = ((jrect >> 4) & 15) + (smp - ii * jnx);
                            if (defer) {
                                q = make_uint2((unsigned)((jgrow * ph + iy) * pw + ix), (unsigned)(kDefer2 | ((3 * si) << 5)));
                            } else {
                                const Tap ty = geom_tap(gby, gsy, iy, H), tx = geom_tap(gbx, gsx, ix, W);
                                // a tap that coincides with its partner (integer sample position) carries weight 0: dropped
                                const bool top = ty.lo >= Y0 && ty.lo < Y0 + TY, bot = ty.hi >= Y0 && ty.hi < Y0 + TY && ty.hi != ty.lo;
                                const bool lef = tx.lo >= X0 && tx.lo < X0 + TX, rig = tx.hi >= X0 && tx.hi < X0 + TX && tx.hi != tx.lo;
                                const int f = (top && lef ? 1 : 0) | (top && rig ? 2 : 0) | (bot && lef ? 4 : 0) | (bot && rig ? 8 : 0);
                                const int px = f ? (ty.lo - Y0) * TX + (tx.lo - X0) : 0;                  // TL pixel; may be negative
                                const unsigned row = dual ? (unsigned)((jrr * ph + iy) * pw + ix) : (unsigned)((jgrow * ph + iy) * pw + ix);
                                q = make_uint2(row, (unsigned)(f | ((3 * si + (dual ? 1 : 0)) << 5) | (px * 2048)));
                                const float wy0 = __fsub_rn(1.f, ty.frac), wx0 = __fsub_rn(1.f, tx.frac);  // crop_and_resize.c:241-247
                                wts = EXACT ? make_float4(wy0, ty.frac, wx0, tx.frac)
                                            : make_float4(wy0 * wx0, wy0 * tx.frac, ty.frac * wx0, ty.frac * tx.frac);
                            }
                        }
                        ws.qa[qn + lane] = q;
                        ws.qw[qn + lane] = wts;
                    }
                    last_row = __shfl_sync(0xffffffffu, q.x, room - 1);
                    last_src = __shfl_sync(0xffffffffu, (int)q.y, room - 1) & (63 << 5);
                    qn += room;
                    qtot += room;
                    g0 += room;
                    if (qn == kChunk) flush();
                }
            }
        }
        // ---- pairs of a two-source set must start on an even slot: pad the set's segment to an even length
        if (qtot & 1) pad_entries(1);
    }
    if (qtot & 7) pad_entries(8 - (qtot & 7));          // whole groups of 8 per tile
    __syncwarp();
    if (qn > 0) flush();
    if (lane == 0) P.bin.tile_n[tile_id] = overflow ? stored : qtot;
}

// grid (total tiles, ceil(slabs / 2)), 64 threads: warp = (tile, 128-channel slab)
template <bool EXACT, int TY, int TX, int U, int MINB>
__global__ void __launch_bounds__(64, MINB) bin_accumulate_kernel(const TParams P) {
    // separate objects on purpose: the compiler then knows that the (runtime-offset) tile stores cannot alias the queue, and
    // hoists the queue reads of a whole group above the read-modify-write chain of the taps
    __shared__ __align__(16) float tile_s[2][TY * TX * 128];
    __shared__ __align__(16) float4 qw_s[2][kChunk];
    __shared__ __align__(16) uint2 qa_s[2][kChunk];
    __shared__ const float *srcs[3 * kMaxSets];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < 3 * kMaxSets) {
        const TSet &S = P.s[min((int)threadIdx.x / 3, P.nsets - 1)];
        const int which = threadIdx.x % 3;
        srcs[threadIdx.x] = which == 0 ? S.grads : which == 1 ? S.grads2 : S.coll;
    }
    __syncthreads();                                   // the only CTA-wide synchronisation
    // Heaviest tiles first: callers list the maps from the finest pyramid level to the coarsest, and a tile of a coarse map
    // collects ~10x the samples of a fine one (C2: ~700 queue entries per P5 tile against ~60 per P2 tile).  In launch order
    // those few hundred long tiles would run last, a handful of warps per SM; reversed they overlap with the short ones.
    const int tile_id = P.bin.total_tiles - 1 - (int)blockIdx.x;
    int t = tile_id, mi = 0;
    while (mi + 1 < P.nmaps && t >= P.m[mi + 1].first_tile) ++mi;
    const TMap &M = P.m[mi];
    const int H = M.H, W = M.W, C = M.C;
    const int slab = blockIdx.y * 2 + w;
    if (slab * 128 >= C) return;
    t -= M.first_tile;
    const int tiles_x = ceil_div(W, TX), tiles_y = ceil_div(H, TY);
    const int per_img = tiles_x * tiles_y;
    const int b = t / per_img;
    t -= b * per_img;
    const int tyi = t / tiles_x, txi = t - tyi * tiles_x;
    const int X0 = txi * TX, Y0 = tyi * TY;
    const int coff = slab * 128 + lane * 4;
    float *__restrict__ tile_lane = tile_s[w] + lane * 4;
    float4 *__restrict__ qw = qw_s[w];
    uint2 *__restrict__ qa = qa_s[w];
    const int n = P.bin.tile_n[tile_id];
    int c = n > 0 ? P.bin.tile_head[tile_id] : 0;
    const bool dirty = n > 0;
    if (dirty) {
        // first chunk on its way while the accumulator is cleared
        uint2 ea0 = P.bin.qa[(size_t)c * kChunk + lane], ea1 = P.bin.qa[(size_t)c * kChunk + 32 + lane];
        float4 ew0 = P.bin.qw[(size_t)c * kChunk + lane], ew1 = P.bin.qw[(size_t)c * kChunk + 32 + lane];
        int cn = n > kChunk ? P.bin.chunk_next[c] : 0;
#pragma unroll
        for (int p = 0; p < TY * TX; ++p) *reinterpret_cast<float4 *>(tile_lane + p * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e0 = 0; e0 < n; e0 += kChunk) {
            const int cnt = min(kChunk, n - e0);       // a multiple of 8
            __syncwarp();
            qa[lane] = ea0; qa[32 + lane] = ea1;
            qw[lane] = ew0; qw[32 + lane] = ew1;
            __syncwarp();
            if (e0 + kChunk < n) {                      // next chunk's entries in flight while this one is consumed
                c = cn;
                ea0 = P.bin.qa[(size_t)c * kChunk + lane]; ea1 = P.bin.qa[(size_t)c * kChunk + 32 + lane];
                ew0 = P.bin.qw[(size_t)c * kChunk + lane]; ew1 = P.bin.qw[(size_t)c * kChunk + 32 + lane];
                if (e0 + 2 * kChunk < n) cn = P.bin.chunk_next[c];
            }
            // ---- consume the chunk: groups of 8, loads of group i+1 in flight while group i is added
            float4 a0[U], a1[U];
            auto prefetch8 = [&](int q0) {              // lane -> (entry q0 + lane / 4, 128-byte line lane % 4 of the slab's 512-byte run)
                const int idx = q0 + (lane >> 2);
                if (idx < cnt) {
                    const uint2 e = qa[idx];
                    prefetch_l2(srcs[(e.y >> 5) & 63] + slab * 128 + (lane & 3) * 32 + (size_t)e.x * C);
                }
            };
            auto load = [&](int q0, float4(&a)[U]) {
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const uint2 e = qa[q0 + k];
                    a[k] = __ldcg(reinterpret_cast<const float4 *>(srcs[(e.y >> 5) & 63] + coff + (size_t)e.x * C));
                }
            };
            auto process = [&](int q0, float4(&a)[U]) {
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int pk = (int)qa[q0 + k].y;
                    if (pk & 15) {                                                       // tap-less entries (first of a pair, padding) only load
                        const float4 wts = qw[q0 + k];
                        float4 g = a[k];
                        if (k & 1) {
                            if (qa[q0 + k - 1].y & kDefer2) g = add_rn4(g, a[k - 1]);    // (g1 + g2) first
                        }
                        const int pk1 = (pk & 15) | ((pk >> 11) << 9);                   // flags | byte offset of the TL pixel
                        apply_sample<EXACT, TX>(tile_lane, pk1, wts, g);
                    }
                }
            };
#pragma unroll
            for (int q0 = U; q0 < U + 32; q0 += 8) prefetch8(q0);
            load(0, a0);
            for (int q0 = 0; q0 < cnt; q0 += 2 * U) {             // cnt is a multiple of 8, U is 4 or 8
                const bool more = q0 + U < cnt;
                prefetch8(q0 + U + 32);
                if (more) load(q0 + U, a1);
                process(q0, a0);
                if (more) {
                    if (U == 8) prefetch8(q0 + U + 40);
                    if (q0 + 2 * U < cnt) load(q0 + 2 * U, a0);
                    process(q0 + U, a1);
                }
            }
        }
    }
    // ---- store the tile (or zeros), once
    __syncwarp();
    float *dst0 = M.gimg + (((long)b * H + Y0) * (long)W + X0) * C + coff;
    const long row_stride = (long)W * C;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (Y0 + TY <= H && X0 + TX <= W && !P.accumulate) {            // interior tile: straight-line stores
#pragma unroll
        for (int yy = 0; yy < TY; ++yy) {
#pragma unroll
            for (int xx = 0; xx < TX; ++xx) {
                const float4 v = dirty ? *reinterpret_cast<const float4 *>(tile_lane + (yy * TX + xx) * 128) : zero4;
                __stcs(reinterpret_cast<float4 *>(dst0 + yy * row_stride + xx * C), v);
            }
        }
    } else {
        for (int yy = 0; yy < TY && Y0 + yy < H; ++yy) {
            for (int xx = 0; xx < TX && X0 + xx < W; ++xx) {
                float *dst = dst0 + yy * row_stride + xx * C;
                float4 v = dirty ? *reinterpret_cast<const float4 *>(tile_lane + (yy * TX + xx) * 128) : zero4;
                if (P.accumulate) v = add_rn4(*reinterpret_cast<const float4 *>(dst), v);
                __stcs(reinterpret_cast<float4 *>(dst), v);
            }
        }
    }
}


}  // namespace tile
}  // namespace fi

using namespace fi;
using namespace fi::tile;

namespace fi { namespace tile { int pix_accumulate(const TParams &P, int exact, cudaStream_t stream); } }   // roi_align_bwd_pix.cu

// =====================================================================================================================
// Host side.  One backward = PLAN (tile_prep + bin_enumerate: needs the boxes only, so it can run at forward time, on another
// stream) + RUN (tile_collapse + accumulate: needs the gradients).  All scratch memory is ONE caller-provided block
// (fi_crop_sets_backward_workspace tells its size) -- the Python layer takes it from torch's caching allocator, which also
// makes the whole thing CUDA-graph capturable; the one-shot entry fi_crop_sets_backward uses a grow-only block per
// (device, stream) instead and refuses to grow it while that stream is being captured.
// =====================================================================================================================
namespace {

struct PlanData {                      // what fi_bwd_plan holds
    unsigned magic;
    int exact, form, nsets_in, binned;
    size_t bytes;
    fi_bwd_set in[kMaxSets];           // the caller's sets at plan time, for the consistency check of run
    int set_of_in[kMaxSets];           // caller's set i -> index in P.s
    TParams P;
};
static_assert(sizeof(PlanData) <= sizeof(fi_bwd_plan), "fi_bwd_plan too small");
constexpr unsigned kPlanMagic = 0xf1b200a2u;

size_t up16(size_t v) { return (v + 15) / 16 * 16; }

// Fills P (maps, sets, tile counts) and binds every scratch array to ws (ws == nullptr: sizes only).  Returns the bytes
// needed, 0 when the sets do not qualify (caller falls back to the reduction kernels), or (size_t)-1 on invalid input.
size_t build(const fi_bwd_set *sets, int num_sets, int exact, long max_entries, char *ws, PlanData &D) {
    TParams &P = D.P;
    if (num_sets < 1 || num_sets > kMaxSets) return 0;
    const int form = option(FI_OPT_BWD_FORM), shape = option(FI_OPT_TILE_SHAPE);
    // tile shape: 4 rows x 8 pixels (a tile row is 8 KB contiguous in NHWC, C = 256); 4x4 / 2x8 only exist for the
    // shared-memory forms (measurements)
    const int TYs = (form != 0 && shape == 2) ? 2 : 4, TXs = (form != 0 && shape == 1) ? 4 : 8;
    D.form = form; D.exact = exact ? 1 : 0; D.nsets_in = num_sets; D.binned = form != 2;
    P.nsets = 0; P.nmaps = 0; P.accumulate = 0; P.collapse = exact ? 0 : 1;
    // group the sets by map, maps in order of first appearance
    for (int i = 0; i < num_sets; ++i) {
        const fi_bwd_set &h = sets[i];
        if (h.depth % 128 != 0 || h.image_height > 32767 || h.image_width > 32767 || h.crop_height > kMaxCrop || h.crop_width > kMaxCrop ||
            h.crop_height < 1 || h.crop_width < 1) return 0;
        if (((uintptr_t)h.grads_image % 16) || ((uintptr_t)h.grads % 16) || ((uintptr_t)h.grads2 % 16) || ((uintptr_t)h.boxes % 16)) return 0;
        if ((long)h.num_boxes * h.crop_height * h.crop_width >= (1L << 29) || h.num_boxes >= (1 << 24)) return 0;
        bool seen = false;
        for (int q = 0; q < i; ++q) seen = seen || (sets[q].grads_image == h.grads_image);
        if (seen) continue;
        if (P.nmaps == kMaxMaps) return 0;
        TMap &M = P.m[P.nmaps];
        M.gimg = h.grads_image; M.B = h.batch; M.H = h.image_height; M.W = h.image_width; M.C = h.depth;
        M.tiles_x = ceil_div(M.W, TXs); M.tiles_y = ceil_div(M.H, TYs);
        M.set_begin = P.nsets;
        for (int q = i; q < num_sets; ++q) {
            const fi_bwd_set &g = sets[q];
            if (g.grads_image != h.grads_image) continue;
            if (g.batch != h.batch || g.image_height != h.image_height || g.image_width != h.image_width || g.depth != h.depth) {
                set_error(FI_ERR_INVALID, "crop backward: sets %d and %d name the same map with different shapes", i, q);
                return (size_t)-1;
            }
            D.set_of_in[q] = P.nsets;
            TSet &S = P.s[P.nsets++];
            S.grads = g.grads; S.grads2 = g.grads2; S.boxes = reinterpret_cast<const float4 *>(g.boxes); S.box_ind = g.box_ind; S.src_row = g.src_row;
            S.R = g.num_boxes; S.R_dev = g.num_boxes_dev; S.ph = g.crop_height; S.pw = g.crop_width; S.map = P.nmaps;
        }
        M.set_end = P.nsets;
        ++P.nmaps;
    }
    long tiles = 0;
    long total_R = 0;
    for (int m = 0; m < P.nmaps; ++m) {
        TMap &M = P.m[m];
        if (tiles + (long)M.B * M.tiles_x * M.tiles_y >= (1L << 31)) return 0;
        M.first_tile = (int)tiles;
        tiles += (long)M.B * M.tiles_x * M.tiles_y;
    }
    // per-tile sample lists in 64-entry chunks.  A sample's taps lie in at most 2x2 tiles, two-source sets queue two entries
    // per sample; + one partly filled chunk per tile, + one more for the (<= 7 + one per set) tap-less padding entries.
    // max_entries > 0: the caller's own bound on the entries (it may know that the sets partition the boxes).
    long entries_bound = 0;
    for (int i = 0; i < P.nsets; ++i) {
        entries_bound += 4L * P.s[i].R * P.s[i].ph * P.s[i].pw * (P.s[i].grads2 ? 2 : 1) + 8L * 4 * P.s[i].R;
        total_R += P.s[i].R;
    }
    if (max_entries > 0 && max_entries < entries_bound) entries_bound = max_entries;
    const long pool = entries_bound / kChunk + 2 * tiles + 8;
    if (pool >= (1L << 31) / kChunk) D.binned = 0;
    // ---- layout: [per-set image ranges (memset 0xFF)] [degenerate count + list | chunk cursor, overflow flag, 32 tile counters
    //      (memset 0)] [records, geometry] [collapse rows] [lists]
    char *p = ws;
    size_t bytes = 0;
    auto take = [&](size_t n) { char *q = ws ? p : nullptr; if (ws) p += n; bytes += n; return q; };
    for (int i = 0; i < P.nsets; ++i) P.s[i].range = reinterpret_cast<unsigned *>(take(up16((size_t)P.m[P.s[i].map].B * 2 * sizeof(unsigned))));
    P.range_bytes = bytes;
    P.deg_list = reinterpret_cast<int *>(take(up16((size_t)(1 + total_R) * sizeof(int))));
    P.bin.cursor = reinterpret_cast<int *>(take(16));               // [0] chunks handed out, [1] overflow flag
    P.bin.work = reinterpret_cast<int *>(take(128));
    for (int i = 0; i < P.nsets; ++i) {
        P.s[i].rec = reinterpret_cast<int4 *>(take(up16((size_t)P.s[i].R * sizeof(int4))));
        P.s[i].geom = reinterpret_cast<float4 *>(take(up16((size_t)P.s[i].R * sizeof(float4))));
    }
    for (int i = 0; i < P.nsets; ++i)                                // only the rows of degenerate boxes are ever touched
        P.s[i].coll = reinterpret_cast<float *>(take(P.collapse ? (size_t)P.s[i].R * 4 * P.m[P.s[i].map].C * sizeof(float) : 0));
    P.bin.pool = (int)pool;
    P.bin.total_tiles = (int)tiles;
    P.bin.pix_group = 1 + option(FI_OPT_PIX_GROUP);
    P.bin.tile_head = nullptr; P.bin.tile_n = nullptr; P.bin.chunk_next = nullptr; P.bin.qw = nullptr; P.bin.qa = nullptr;
    if (D.binned) {
        P.bin.tile_head = reinterpret_cast<int *>(take(up16((size_t)tiles * sizeof(int))));
        P.bin.tile_n = reinterpret_cast<int *>(take(up16((size_t)tiles * sizeof(int))));
        P.bin.chunk_next = reinterpret_cast<int *>(take(up16((size_t)pool * sizeof(int))));
        P.bin.qw = reinterpret_cast<float4 *>(take((size_t)pool * kChunk * sizeof(float4)));
        P.bin.qa = reinterpret_cast<uint2 *>(take((size_t)pool * kChunk * sizeof(uint2)));
    }
    D.bytes = bytes ? bytes : 16;
    D.magic = kPlanMagic;
    for (int i = 0; i < num_sets; ++i) D.in[i] = sets[i];
    return D.bytes;
}

int max_slabs_of(const TParams &P) {
    int m = 1;
    for (int i = 0; i < P.nmaps; ++i) m = m > P.m[i].C / 128 ? m : P.m[i].C / 128;
    return m;
}

// prep (+ enumerate): everything that depends on the boxes only
int launch_plan(const PlanData &D, cudaStream_t stream) {
    const TParams &P = D.P;
    char *ws = reinterpret_cast<char *>(P.s[0].range);
    cudaError_t e = cudaSuccess;
    if (P.range_bytes) e = cudaMemsetAsync(ws, 0xFF, P.range_bytes, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.deg_list, 0, sizeof(int), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.bin.cursor, 0, 16 + 128, stream);
    if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "crop backward: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    int max_R = 0;
    for (int i = 0; i < P.nsets; ++i) max_R = max_R > P.s[i].R ? max_R : P.s[i].R;
    int rc = ok();
    if (max_R > 0) {
        tile_prep_kernel<<<dim3(ceil_div(max_R, 128), P.nsets), 128, 0, stream>>>(P);
        rc = check_launch("crop backward[prep]");
    }
    const int tiles = P.bin.total_tiles;
    if (rc == FI_OK && tiles > 0 && D.binned) {
        const int egrid = (tiles + 3) / 4;
        const int shape = option(FI_OPT_TILE_SHAPE);
        const bool t48 = D.form == 0 || shape == 0;
        if (t48) { if (D.exact) bin_enumerate_kernel<true, 4, 8><<<egrid, 128, 0, stream>>>(P); else bin_enumerate_kernel<false, 4, 8><<<egrid, 128, 0, stream>>>(P); }
        else if (shape == 1) { if (D.exact) bin_enumerate_kernel<true, 4, 4><<<egrid, 128, 0, stream>>>(P); else bin_enumerate_kernel<false, 4, 4><<<egrid, 128, 0, stream>>>(P); }
        else { if (D.exact) bin_enumerate_kernel<true, 2, 8><<<egrid, 128, 0, stream>>>(P); else bin_enumerate_kernel<false, 2, 8><<<egrid, 128, 0, stream>>>(P); }
        rc = check_launch("crop backward[enumerate]");
    }
    return rc;
}

// collapse + accumulate: everything that needs the gradients
int launch_run(const PlanData &D, const TParams &P, cudaStream_t stream) {
    int rc = ok();
    long total_R = 0;
    for (int i = 0; i < P.nsets; ++i) total_R += P.s[i].R;
    if (total_R > 0 && P.collapse) {
        const int grid = (int)(total_R < 4L * kNumSMs ? total_R : 4L * kNumSMs);
        tile_collapse_kernel<<<grid, 256, 0, stream>>>(P);
        rc = check_launch("crop backward[collapse]");
    }
    const int tiles = P.bin.total_tiles;
    if (rc != FI_OK || tiles <= 0) return rc;
    if (D.binned && D.form == 0) {                               // tile counters of the persistent kernel: a plan may be run more than once
        cudaError_t e = cudaMemsetAsync(P.bin.work, 0, 128, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "crop backward: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    }
    const int exact = D.exact, shape = option(FI_OPT_TILE_SHAPE);
    const dim3 grid((unsigned)tiles, ceil_div(max_slabs_of(P), 2));
    if (D.binned) {
        if (D.form == 0) {                                       // default: bulk-copy staged, register-accumulating kernel
            rc = pix_accumulate(P, exact, stream);
            if (rc != FI_ERR_UNSUPPORTED) return rc;             // (mixed channel counts: the shared-memory kernel takes it, lists are ready)
        }
        const bool t48 = D.form == 0 || shape == 0;
        if (t48) { if (exact) bin_accumulate_kernel<true, 4, 8, 8, 6><<<grid, 64, 0, stream>>>(P); else bin_accumulate_kernel<false, 4, 8, 8, 6><<<grid, 64, 0, stream>>>(P); }
        else if (shape == 1) { if (exact) bin_accumulate_kernel<true, 4, 4, 4, 10><<<grid, 64, 0, stream>>>(P); else bin_accumulate_kernel<false, 4, 4, 4, 10><<<grid, 64, 0, stream>>>(P); }
        else { if (exact) bin_accumulate_kernel<true, 2, 8, 4, 10><<<grid, 64, 0, stream>>>(P); else bin_accumulate_kernel<false, 2, 8, 4, 10><<<grid, 64, 0, stream>>>(P); }
        return check_launch("crop backward[accumulate]");
    }
    if (shape == 0) { if (exact) bwd_smem_tile_kernel<true, 4, 8, FI_TILE_MINB><<<grid, 64, 0, stream>>>(P); else bwd_smem_tile_kernel<false, 4, 8, FI_TILE_MINB><<<grid, 64, 0, stream>>>(P); }
    else if (shape == 1) { if (exact) bwd_smem_tile_kernel<true, 4, 4, 10><<<grid, 64, 0, stream>>>(P); else bwd_smem_tile_kernel<false, 4, 4, 10><<<grid, 64, 0, stream>>>(P); }
    else { if (exact) bwd_smem_tile_kernel<true, 2, 8, 10><<<grid, 64, 0, stream>>>(P); else bwd_smem_tile_kernel<false, 2, 8, 10><<<grid, 64, 0, stream>>>(P); }
    return check_launch("crop backward[tile]");
}

// Grow-only scratch memory of the one-shot entry, one block per (device, stream): every use is ordered on that stream, so
// consecutive calls can share it without synchronisation.  Growing frees and re-allocates (a synchronisation): refused while
// the stream is being captured -- a captured graph would keep the old pointer -- use the plan / run entries with a
// caller-owned workspace there.
struct WsSlot { int dev; cudaStream_t stream; char *ptr; size_t cap; };
WsSlot g_ws[32];
int g_nws = 0;
std::mutex g_ws_mu;

char *workspace(size_t bytes, cudaStream_t stream) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_ws_mu);
    WsSlot *slot = nullptr;
    for (int i = 0; i < g_nws; ++i) if (g_ws[i].dev == dev && g_ws[i].stream == stream) slot = &g_ws[i];
    if (slot && slot->cap >= bytes) return slot->ptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        set_error(FI_ERR_UNSUPPORTED, "crop backward: the internal workspace cannot grow during stream capture; use fi_crop_sets_backward_plan / _run with "
                                      "a caller-owned workspace");
        return nullptr;
    }
    if (!slot) {
        if (g_nws == 32) {                             // recycle the oldest entry
            cudaStreamSynchronize(g_ws[0].stream);
            cudaFree(g_ws[0].ptr);
            for (int i = 1; i < g_nws; ++i) g_ws[i - 1] = g_ws[i];
            --g_nws;
        }
        slot = &g_ws[g_nws++];
        slot->dev = dev; slot->stream = stream; slot->ptr = nullptr; slot->cap = 0;
    } else {
        cudaStreamSynchronize(stream);                 // the old block may still be in use by enqueued work
        cudaFree(slot->ptr);
        slot->ptr = nullptr; slot->cap = 0;
    }
    const size_t want = bytes + bytes / 2 + (1 << 20);
    cudaError_t e = cudaMalloc((void **)&slot->ptr, want);
    if (e != cudaSuccess) {
        slot->ptr = nullptr;
        set_error(FI_ERR_CUDA, "crop backward: workspace (%zu B): %s (torch's caching allocator may hold the memory: pass a torch-allocated workspace "
                               "to fi_crop_sets_backward_plan instead)", want, cudaGetErrorString(e));
        return nullptr;
    }
    slot->cap = want;
    return slot->ptr;
}

}  // namespace

namespace fi { char *tile_workspace(size_t bytes, cudaStream_t stream) { return workspace(bytes, stream); } }   // for roi_align_nchw_bwd.cu

namespace {

bool same_geometry(const fi_bwd_set &a, const fi_bwd_set &b) {
    return a.boxes == b.boxes && a.box_ind == b.box_ind && a.src_row == b.src_row && a.batch == b.batch &&
           a.image_height == b.image_height && a.image_width == b.image_width && a.depth == b.depth && a.num_boxes == b.num_boxes &&
           a.num_boxes_dev == b.num_boxes_dev && a.crop_height == b.crop_height && a.crop_width == b.crop_width && (a.grads2 != nullptr) == (b.grads2 != nullptr);
}

}  // namespace

FI_API size_t fi_crop_sets_backward_workspace(const fi_bwd_set *sets, int num_sets, int exact, long max_entries) {
    if (!sets || option(FI_OPT_BWD_FORM) == 3) return 0;
    PlanData D;
    const size_t b = build(sets, num_sets, exact, max_entries, nullptr, D);
    return b == (size_t)-1 ? 0 : b;
}

FI_API int fi_crop_sets_backward_plan(const fi_bwd_set *sets, int num_sets, int exact, long max_entries, void *workspace_ptr, size_t workspace_bytes,
                                      fi_bwd_plan *plan, cudaStream_t stream) {
    FI_REQUIRE(sets && plan && workspace_ptr, "fi_crop_sets_backward_plan: null pointer");
    FI_REQUIRE(((uintptr_t)workspace_ptr % 16) == 0, "fi_crop_sets_backward_plan: the workspace must be 16-byte aligned");
    if (option(FI_OPT_BWD_FORM) == 3) { set_error(FI_ERR_UNSUPPORTED, "crop backward: the reduction form has no plan"); return FI_ERR_UNSUPPORTED; }
    PlanData &D = *reinterpret_cast<PlanData *>(plan);
    D.magic = 0;
    const size_t need = build(sets, num_sets, exact, max_entries, static_cast<char *>(workspace_ptr), D);
    if (need == (size_t)-1) return FI_ERR_INVALID;
    if (need == 0) { D.magic = 0; set_error(FI_ERR_UNSUPPORTED, "crop backward: these sets need the reduction kernels (depth %% 128, crops <= 16, <= 8 maps, 16-byte alignment)"); return FI_ERR_UNSUPPORTED; }
    if (need > workspace_bytes) { D.magic = 0; set_error(FI_ERR_INVALID, "fi_crop_sets_backward_plan: workspace of %zu bytes, %zu needed", workspace_bytes, need); return FI_ERR_INVALID; }
    return launch_plan(D, stream);
}

// `sets` carries the gradients (grads, grads2, grads_image may differ from plan time); boxes / sizes / two-source pattern must
// be the plan's.
FI_API int fi_crop_sets_backward_run(const fi_bwd_plan *plan, const fi_bwd_set *sets, int num_sets, int zero_first, cudaStream_t stream) {
    FI_REQUIRE(plan && sets, "fi_crop_sets_backward_run: null pointer");
    const PlanData &D = *reinterpret_cast<const PlanData *>(plan);
    FI_REQUIRE(D.magic == kPlanMagic, "fi_crop_sets_backward_run: not a plan (fi_crop_sets_backward_plan failed or was not called)");
    FI_REQUIRE(num_sets == D.nsets_in, "fi_crop_sets_backward_run: %d sets, the plan was made for %d", num_sets, D.nsets_in);
    TParams P = D.P;
    P.accumulate = zero_first ? 0 : 1;
    for (int i = 0; i < num_sets; ++i) {
        FI_REQUIRE(same_geometry(sets[i], D.in[i]), "fi_crop_sets_backward_run: set %d differs from the planned one (boxes, sizes or two-source pattern)", i);
        FI_REQUIRE(sets[i].num_boxes == 0 || sets[i].grads, "fi_crop_sets_backward_run: set %d has no gradient", i);
        FI_REQUIRE(((uintptr_t)sets[i].grads % 16) == 0 && ((uintptr_t)sets[i].grads2 % 16) == 0, "fi_crop_sets_backward_run: unaligned gradient in set %d", i);
        TSet &S = P.s[D.set_of_in[i]];
        S.grads = sets[i].grads; S.grads2 = sets[i].grads2;
        // the maps may be named only now (plan time: any distinct 16-byte aligned placeholders); sets that shared one still must
        TMap &M = P.m[S.map];
        bool first = true;
        for (int q = 0; q < i; ++q) first = first && P.s[D.set_of_in[q]].map != S.map;
        FI_REQUIRE(sets[i].grads_image && ((uintptr_t)sets[i].grads_image % 16) == 0, "fi_crop_sets_backward_run: bad grads_image in set %d", i);
        FI_REQUIRE(first || M.gimg == sets[i].grads_image, "fi_crop_sets_backward_run: set %d no longer shares its map with the sets it was planned with", i);
        M.gimg = sets[i].grads_image;
    }
    for (int i = 0; i < num_sets; ++i)
        for (int q = 0; q < i; ++q)
            FI_REQUIRE((sets[i].grads_image == sets[q].grads_image) == (P.s[D.set_of_in[i]].map == P.s[D.set_of_in[q]].map),
                       "fi_crop_sets_backward_run: sets %d and %d: map sharing differs from the plan", q, i);
    return launch_run(D, P, stream);
}

// 1 when the sample lists overflowed their pool (cannot happen with the default bound; a caller's max_entries may be too
// small), 0 when not, negative on error.  Synchronises the stream.
FI_API int fi_crop_sets_backward_overflow(const fi_bwd_plan *plan, cudaStream_t stream) {
    FI_REQUIRE(plan, "fi_crop_sets_backward_overflow: null pointer");
    const PlanData &D = *reinterpret_cast<const PlanData *>(plan);
    FI_REQUIRE(D.magic == kPlanMagic, "fi_crop_sets_backward_overflow: not a plan");
    int flag = 0;
    cudaError_t e = cudaMemcpyAsync(&flag, D.P.bin.cursor + 1, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_crop_sets_backward_overflow: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    return flag ? 1 : 0;
}

// One-shot form on the internal workspace.  Returns FI_ERR_UNSUPPORTED (nothing touched) when the sets do not qualify so that
// the caller can use the reduction kernels.  exact != 0: arithmetic and order of crop_and_resize.c:190-250 (bit-identical for
// one set per map).
int fi_tile_backward(const fi_bwd_set *sets, int num_sets, int accumulate, int exact, cudaStream_t stream) {
    PlanData D;
    const size_t need = build(sets, num_sets, exact, 0, nullptr, D);
    if (need == (size_t)-1) return FI_ERR_INVALID;
    if (need == 0) return FI_ERR_UNSUPPORTED;
    char *ws = workspace(need, stream);
    if (!ws) return fi_last_status();
    build(sets, num_sets, exact, 0, ws, D);
    int rc = launch_plan(D, stream);
    if (rc != FI_OK) return rc;
    TParams P = D.P;
    P.accumulate = accumulate ? 1 : 0;
    return launch_run(D, P, stream);
}
