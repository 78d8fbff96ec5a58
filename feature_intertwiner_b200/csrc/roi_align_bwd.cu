// RoIAlign backward as a GATHER: every pixel of the dense gradient map is computed by exactly one warp and
// written exactly once.  This is the DETERMINISTIC mode of the backward (fi_set_deterministic(1) or the
// `deterministic` argument of fi_crop_and_resize_backward_multi); the default mode is the reduction kernel of
// roi_align.cu, which is faster today on heavily overlapping RoIs ("backward: two formulations").
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
    const int *range;                  // [B,2] first / last box index whose box_ind == b (first > last: none)
    int R, ph, pw;
};

constexpr int kMaxSets = 4;
struct SetsDev {
    SetDev s[kMaxSets];
    int n;
};

// ---- prep: one warp per box: tap table + footprint bounds --------------------------------------------
__global__ void __launch_bounds__(256) bwd_prep_kernel(const float *__restrict__ boxes, const int *__restrict__ box_ind, int R, int B, int H,
                                                      int W, int ph, int pw, TapEntry *__restrict__ taps, int4 *__restrict__ bounds, int *__restrict__ range) {
    const int lane = threadIdx.x & 31;
    const int r = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= R) return;
    const int b = box_ind[r];
    const bool bad = (b < 0 || b >= B);
    if (!bad && lane == 0) { atomicMin(range + 2 * b, r); atomicMax(range + 2 * b + 1, r); }
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

__global__ void range_init_kernel(int *range, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { range[2 * b] = 0x7fffffff; range[2 * b + 1] = -1; }
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
//
// A warp's work is an ORDERED stream of (sample -> this pixel row) contributions.  Walking it naively
// (find a sample, load it, add it) leaves one 512 B load in flight per warp; the first version of this kernel
// did that and ran 6x slower than the reduction kernels.  So the stream is split in two stages per warp:
//   enumerate  lanes expand (crop rows touching y) x (crop columns touching the tile) of a box in parallel --
//              candidate c -> (i, j) by find-nth-set-bit on the two ballot masks -- and append 32-byte
//              descriptors (pointers, lerp weights, target columns) to a per-warp queue in shared memory;
//   drain      the queue is consumed in order, 8 gradient loads in flight, then the ordered adds.
struct __align__(16) Desc {
    const float *g1, *g2;          // crop gradient(s) of the sample (slab offset not yet applied); g2 may be null
    float fy, fx;
    short xlo, xhi;                // target columns relative to the tile (may lie outside 0..7: skipped)
    short top, bot;                // does the sample's top / bottom tap row equal this warp's pixel row?
};
constexpr int kQueue = 64;

__device__ __forceinline__ void drain_queue(const Desc *queue, int qn, int coff, float4 (&acc)[kTile]) {
    for (int t0 = 0; t0 < qn; t0 += 8) {
        float4 gv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (t0 + q < qn) {
                const Desc &d = queue[t0 + q];
                gv[q] = __ldg(reinterpret_cast<const float4 *>(d.g1 + coff));
                if (d.g2) gv[q] = add_rn4(gv[q], __ldg(reinterpret_cast<const float4 *>(d.g2 + coff)));
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (t0 + q < qn) {
                const Desc &d = queue[t0 + q];
                const float fy = d.fy, fx = d.fx;
                const float wx_lo = __fsub_rn(1.f, fx);
                if (d.top) {                                                      // TL, TR  (crop_and_resize.c:241-243)
                    const float4 dtop = mul_rn4(__fsub_rn(1.f, fy), gv[q]);
                    acc_add(acc, d.xlo, mul_rn4(wx_lo, dtop));
                    acc_add(acc, d.xhi, mul_rn4(fx, dtop));
                }
                if (d.bot) {                                                      // BL, BR  (:245-247)
                    const float4 dbot = mul_rn4(fy, gv[q]);
                    acc_add(acc, d.xlo, mul_rn4(wx_lo, dbot));
                    acc_add(acc, d.xhi, mul_rn4(fx, dbot));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kTile * 32) bwd_tile_kernel(const SetsDev sets, int B, int H, int W, int C, int slabs, int accumulate,
                                                             float *__restrict__ gimg) {
    __shared__ int list[256];
    __shared__ int warp_hits[kTile];
    __shared__ int nlist;
    __shared__ Desc queues[kTile][kQueue];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.z / slabs, slab = blockIdx.z - b * slabs;
    const int X0 = blockIdx.x * kTile, Y0 = blockIdx.y * kTile;
    const int y = Y0 + w;
    const int coff = slab * 128 + lane * 4;
    Desc *queue = queues[w];
    int qn = 0;
    float4 acc[kTile];
#pragma unroll
    for (int x = 0; x < kTile; ++x) acc[x] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int si = 0; si < sets.n; ++si) {
        const SetDev &S = sets.s[si];
        const int r_begin = S.range[2 * b] & ~255, r_end = S.range[2 * b + 1];   // boxes of image b live in [r_begin, r_end]
        for (int base = r_begin; base <= r_end; base += 256) {
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
                    const int nx = __popc(xmask), ncand = __popc(ymask) * nx;
                    for (int c0 = 0; c0 < ncand; c0 += 32) {
                        if (qn + 32 > kQueue) { __syncwarp(); drain_queue(queue, qn, coff, acc); qn = 0; __syncwarp(); }
                        const int c = c0 + lane;
                        const bool valid = c < ncand;
                        const int ii = valid ? c / nx : 0, jj = valid ? c - ii * nx : 0;
                        const int i = __fns(ymask, 0, ii + 1), j = __fns(xmask, 0, jj + 1);   // (i major, j minor): the CPU loop order
                        const int ylo = __shfl_sync(0xffffffffu, (int)e.lo, i), yhi = __shfl_sync(0xffffffffu, (int)e.hi, i);
                        const float fy = __shfl_sync(0xffffffffu, e.frac, i);
                        const int xlo = __shfl_sync(0xffffffffu, (int)e.lo, 16 + j), xhi = __shfl_sync(0xffffffffu, (int)e.hi, 16 + j);
                        const float fx = __shfl_sync(0xffffffffu, e.frac, 16 + j);
                        if (valid) {
                            Desc d;
                            d.g1 = S.grads + (((grow * S.ph + i) * (long)S.pw) + j) * C;
                            d.g2 = S.grads2 ? S.grads2 + ((((long)rr * S.ph + i) * (long)S.pw) + j) * C : nullptr;
                            d.fy = fy; d.fx = fx;
                            d.xlo = (short)(xlo - X0); d.xhi = (short)(xhi - X0);
                            d.top = (ylo == y); d.bot = (yhi == y);
                            queue[qn + (c - c0)] = d;
                        }
                        qn += min(32, ncand - c0);
                    }
                }
            }
            __syncthreads();           // the list is rebuilt by the next chunk
        }
    }
    __syncwarp();
    drain_queue(queue, qn, coff, acc);
    if (y < H) {
        float *dst = gimg + (((long)b * H + y) * (long)W + X0) * C + coff;
#pragma unroll
        for (int x = 0; x < kTile; ++x) {
            if (X0 + x < W) {
                float4 v = acc[x];
                if (accumulate) v = add_rn4(*reinterpret_cast<const float4 *>(dst + (long)x * C), v);
                __stcs(reinterpret_cast<float4 *>(dst + (long)x * C), v);
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
        bytes += (size_t)s.num_boxes * (32 * sizeof(TapEntry) + sizeof(int4)) + (size_t)B * 2 * sizeof(int) + 16;
    }
    char *ws = nullptr;
    if (bytes) {
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
        d.range = reinterpret_cast<const int *>(p); p += ((size_t)B * 2 * sizeof(int) + 15) / 16 * 16;
        range_init_kernel<<<ceil_div(B, 128), 128, 0, stream>>>(const_cast<int *>(d.range), B);
        bwd_prep_kernel<<<ceil_div(s.num_boxes, 8), 256, 0, stream>>>(s.boxes, s.box_ind, s.num_boxes, B, H, W, s.crop_height, s.crop_width,
                                                                     const_cast<TapEntry *>(d.taps), const_cast<int4 *>(d.bounds), const_cast<int *>(d.range));
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

int fi_scatter_backward_nhwc(const float *grads, const float *grads2, const float *boxes, const int *box_ind, const int *src_row, int R, int B,
                             int H, int W, int ph, int pw, int C, float *gimg, cudaStream_t stream);   // roi_align.cu

int fi_tile_backward(const fi_bwd_set *sets, int num_sets, int accumulate, int exact, cudaStream_t stream);   // roi_align_bwd_tile.cu

static int g_deterministic = 0;
FI_API int fi_set_deterministic(int on) { const int old = g_deterministic; g_deterministic = on ? 1 : 0; return old; }
FI_API int fi_get_deterministic(void) { return g_deterministic; }

FI_API int fi_crop_and_resize_backward_multi(const fi_crop_set *sets, int num_sets, int batch, int image_height, int image_width, int depth,
                                             float *grads_image, int accumulate, int deterministic, cudaStream_t stream) {
    FI_REQUIRE(sets && num_sets >= 1 && batch > 0 && image_height > 0 && image_width > 0 && depth > 0 && grads_image, "fi_crop_and_resize_backward_multi: bad arguments");
    for (int i = 0; i < num_sets; ++i) {
        const fi_crop_set &s = sets[i];
        FI_REQUIRE(s.num_boxes >= 0 && s.crop_height > 0 && s.crop_width > 0, "fi_crop_and_resize_backward_multi: bad set %d", i);
        FI_REQUIRE(s.num_boxes == 0 || (s.grads && s.boxes && s.box_ind), "fi_crop_and_resize_backward_multi: null pointer in set %d", i);
    }
    fi_bwd_set tmp[12];
    const bool fits = num_sets <= 12 && depth % 128 == 0;
    for (int i = 0; fits && i < num_sets; ++i) {
        tmp[i].grads_image = grads_image; tmp[i].grads = sets[i].grads; tmp[i].grads2 = sets[i].grads2; tmp[i].boxes = sets[i].boxes;
        tmp[i].box_ind = sets[i].box_ind; tmp[i].src_row = sets[i].src_row; tmp[i].batch = batch; tmp[i].image_height = image_height;
        tmp[i].image_width = image_width; tmp[i].depth = depth; tmp[i].num_boxes = sets[i].num_boxes;
        tmp[i].crop_height = sets[i].crop_height; tmp[i].crop_width = sets[i].crop_width;
    }
    if (deterministic) {
        // exact tile-owner kernel (roi_align_bwd_tile.cu); FI_BWD=gather keeps the older register-accumulator gather of this file
        const char *mode = getenv("FI_BWD");
        int rc = FI_ERR_UNSUPPORTED;
        if (mode && mode[0] == 'g') rc = gather_backward(sets, num_sets, batch, image_height, image_width, depth, grads_image, accumulate, stream);
        else if (fits) rc = fi_tile_backward(tmp, num_sets, accumulate, 1, stream);
        if (rc != FI_ERR_UNSUPPORTED) return rc;
        set_error(FI_ERR_UNSUPPORTED, "deterministic RoIAlign backward needs depth %% 128 == 0, crops <= 16x16, 16-byte aligned NHWC tensors");
        return FI_ERR_UNSUPPORTED;
    }
    // tile-owner kernel / reductions when the shape qualifies (fi_crop_sets_backward picks) ...
    if (fits) {
        const int rc = fi_crop_sets_backward(tmp, num_sets, accumulate ? 0 : 1, stream);
        if (rc != FI_ERR_UNSUPPORTED) return rc;
    }
    // ... else one zero-fill for all sets, then vector reductions set by set
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(grads_image, 0, sizeof(float) * (size_t)batch * depth * image_height * image_width, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_crop_and_resize_backward_multi: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    }
    for (int i = 0; i < num_sets; ++i) {
        const fi_crop_set &s = sets[i];
        const int rc = fi_scatter_backward_nhwc(s.grads, s.grads2, s.boxes, s.box_ind, s.src_row, s.num_boxes, batch, image_height, image_width,
                                                s.crop_height, s.crop_width, depth, grads_image, stream);
        if (rc) return rc;
    }
    return ok();
}
