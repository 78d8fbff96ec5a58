// Greedy IoU NMS, batched over images, with the reduce ON the device.
//
// The reference (lib/nms/src/nms_cuda.c:17-67) builds the 64x64-tile suppression bitmask on the GPU, copies
// all N*ceil(N/64) words to the host (4.5 MB for N=6000, a blocking D2H on the legacy stream) and walks it
// serially on the CPU, once per image.  Here: one launch builds the upper-triangular mask of every image, a
// second launch (one CTA per image) performs the greedy sweep 64 boxes at a time (nms_reduce_kernel below).
// Rule: suppress when IoU > thresh (nms_kernel.cu:63), +1 pixel convention (nms_kernel.cu:16-24).
#include "fi_common.cuh"

namespace fi {

constexpr int kNmsTile = 64;
constexpr int kNmsMaxWordsPerLane = 8;      // n <= 32 * 8 * 64 = 16384 boxes per image

// grid (col tiles, row tiles, images), 64 threads.  `full` = also fill the lower triangle like the reference.
__global__ void __launch_bounds__(kNmsTile) nms_mask_kernel(const float *__restrict__ boxes, int n, float thresh,
                                                           unsigned long long *__restrict__ mask, int full) {
    const int row_t = blockIdx.y, col_t = blockIdx.x, img = blockIdx.z;
    if (!full && col_t < row_t) return;
    const int col_blocks = ceil_div(n, kNmsTile);
    const float *bx = boxes + (long)img * n * 5;
    unsigned long long *mk = mask + (long)img * n * col_blocks;
    const int row_size = min(n - row_t * kNmsTile, kNmsTile), col_size = min(n - col_t * kNmsTile, kNmsTile);
    __shared__ float tile[kNmsTile * 5];          // x1, y1, x2, y2, area(+1 convention) of the column boxes
    if ((int)threadIdx.x < col_size) {
        const float *b = bx + (long)(col_t * kNmsTile + threadIdx.x) * 5;
        const float x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
        tile[threadIdx.x * 5 + 0] = x1; tile[threadIdx.x * 5 + 1] = y1; tile[threadIdx.x * 5 + 2] = x2; tile[threadIdx.x * 5 + 3] = y2;
        tile[threadIdx.x * 5 + 4] = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
    }
    __syncthreads();
    if ((int)threadIdx.x < row_size) {
        const int i = row_t * kNmsTile + threadIdx.x;
        float me[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) me[k] = bx[(long)i * 5 + k];
        const float sa = __fmul_rn(__fadd_rn(__fsub_rn(me[2], me[0]), 1.f), __fadd_rn(__fsub_rn(me[3], me[1]), 1.f));
        unsigned long long bits = 0;
        const int start = (row_t == col_t) ? threadIdx.x + 1 : 0;
        // the reference's devIoU operations in its order (nms_kernel.cu:16-24), areas hoisted; pairs that do not intersect have IoU = +0 and are
        // skipped without the division when that cannot exceed the threshold
        const bool skip_empty = thresh >= 0.f;
        for (int j = start; j < col_size; ++j) {
            const float *o = tile + j * 5;
            const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(me[2], o[2]), fmaxf(me[0], o[0])), 1.f), 0.f);
            const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(me[3], o[3]), fmaxf(me[1], o[1])), 1.f), 0.f);
            const float inter = __fmul_rn(w, h);
            if (skip_empty && inter == 0.f) continue;             // IoU = +-0, or NaN for 0 / 0: never > thresh (a NaN inter falls through)
            if (__fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, o[4]), inter)) > thresh) bits |= 1ULL << j;
        }
        mk[(long)i * col_blocks + col_t] = bits;
    }
}

// One CTA (32 warps) per image.  Per block of 64 boxes:
//   A. warp 0 resolves suppression INSIDE the block from the diagonal tile (serial over bits, words passed by shuffle);
//   meanwhile EVERY warp already holds the mask rows of "its" two boxes of the block (box w and w + 32) to the right of the
//   diagonal -- requested one block AHEAD, before it is known which boxes survive, so the L2 round trip hides behind the
//   previous block's work --
//   B. the warps whose boxes survived OR their rows into the running `removed` bitmap in shared memory.
// The sweep stops as soon as `max_keep` survivors are known (the callers keep the first proposal_count / DET_MAX_INSTANCES
// of the list, lib/layers.py:118-121: the rest of the sweep cannot change those).
// (Round 1 did A and B in one warp, one survivor's row after the other: 2.5 ms for 4 x 6000 boxes, every survivor a dependent
// L2 round trip; loading the rows at the top of their own block: 0.35 ms; this form: see profiles/.)
constexpr int kReduceWarps = 32;
template <int WPL>     // mask words per lane: ceil(ceil(n / 64) / 32)
__global__ void __launch_bounds__(kReduceWarps * 32) nms_reduce_kernel(const unsigned long long *__restrict__ mask, int n, int max_keep,
                                                                      int *__restrict__ keep, int *__restrict__ num_keep) {
    __shared__ unsigned long long removed[32 * kNmsMaxWordsPerLane];
    __shared__ unsigned long long kept_s;
    __shared__ int kept_total_s;
    const int img = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int col_blocks = ceil_div(n, kNmsTile);
    const unsigned long long *mk = mask + (long)img * n * col_blocks;
    int *kp = keep + (long)img * n;
    for (int w = threadIdx.x; w < col_blocks; w += blockDim.x) removed[w] = 0;
    int kept_total = 0;                                            // tracked by warp 0
    // rows of boxes `wid` and `wid + 32` of block b, words b + 1 .. col_blocks - 1 (lane w owns w, w + 32, ...)
    auto load_rows = [&](int b, unsigned long long (&a0)[WPL], unsigned long long (&a1)[WPL]) {
        const int base = b * kNmsTile, size = min(n - base, kNmsTile);
#pragma unroll
        for (int q = 0; q < WPL; ++q) {
            const int w = b + 1 + lane + 32 * q;
            a0[q] = (wid < size && w < col_blocks) ? mk[(long)(base + wid) * col_blocks + w] : 0ULL;
            a1[q] = (wid + 32 < size && w < col_blocks) ? mk[(long)(base + wid + 32) * col_blocks + w] : 0ULL;
        }
    };
    unsigned long long r0[WPL], r1[WPL], n0[WPL], n1[WPL];
    load_rows(0, r0, r1);
    __syncthreads();
    for (int blk = 0; blk < col_blocks; ++blk) {
        const int base = blk * kNmsTile;
        const int size = min(n - base, kNmsTile);
        load_rows(blk + 1, n0, n1);                                // (all zeros past the last block)
        if (wid == 0) {
            // A: diagonal words of this block: lane holds boxes `lane` and `lane + 32`
            unsigned long long d0 = 0, d1 = 0;
            if (lane < size) d0 = mk[(long)(base + lane) * col_blocks + blk];
            if (lane + 32 < size) d1 = mk[(long)(base + lane + 32) * col_blocks + blk];
            unsigned long long cur = removed[blk];
            unsigned long long kept_bits = 0;
            for (int b = 0; b < size; ++b) {
                const unsigned long long row = __shfl_sync(0xffffffffu, b < 32 ? d0 : d1, b & 31);
                if (!((cur >> b) & 1ULL)) { kept_bits |= 1ULL << b; cur |= row; }
            }
            // emit kept indices in order
            const unsigned long long lo = kept_bits & 0xffffffffULL, hi = kept_bits >> 32;
            const int nlo = __popcll(lo);
            if ((lo >> lane) & 1ULL) kp[kept_total + __popcll(lo & ((1ULL << lane) - 1ULL))] = base + lane;
            if ((hi >> lane) & 1ULL) kp[kept_total + nlo + __popcll(hi & ((1ULL << lane) - 1ULL))] = base + 32 + lane;
            kept_total += __popcll(kept_bits);
            if (lane == 0) { kept_s = kept_bits; kept_total_s = kept_total; }
        }
        __syncthreads();
        // B: the survivors' rows go into the bitmap
        const unsigned long long kept_bits = kept_s;
        const bool enough = kept_total_s >= max_keep;              // read between the barriers: warp 0 rewrites it only after the next one
        const bool k0 = (kept_bits >> wid) & 1ULL, k1 = (kept_bits >> (wid + 32)) & 1ULL;
#pragma unroll
        for (int q = 0; q < WPL; ++q) {
            const int w = blk + 1 + lane + 32 * q;
            const unsigned long long v = (k0 ? r0[q] : 0ULL) | (k1 ? r1[q] : 0ULL);
            if (v != 0ULL && w < col_blocks) atomicOr(&removed[w], v);
        }
        __syncthreads();
        if (enough) break;
#pragma unroll
        for (int q = 0; q < WPL; ++q) { r0[q] = n0[q]; r1[q] = n1[q]; }
    }
    if (wid == 0) {
        for (int i = kept_total + lane; i < n; i += 32) kp[i] = -1;
        if (lane == 0) num_keep[img] = kept_total;
    }
}

static int launch_nms_reduce(const unsigned long long *mask, int n_images, int n, int max_keep, int *keep, int *num_keep, cudaStream_t stream) {
    const int wpl = ceil_div(ceil_div(n, kNmsTile), 32);
    const dim3 g((unsigned)n_images), b(kReduceWarps * 32);
    switch (wpl) {
        case 1: nms_reduce_kernel<1><<<g, b, 0, stream>>>(mask, n, max_keep, keep, num_keep); break;
        case 2: nms_reduce_kernel<2><<<g, b, 0, stream>>>(mask, n, max_keep, keep, num_keep); break;
        case 3: nms_reduce_kernel<3><<<g, b, 0, stream>>>(mask, n, max_keep, keep, num_keep); break;
        case 4: nms_reduce_kernel<4><<<g, b, 0, stream>>>(mask, n, max_keep, keep, num_keep); break;
        default: nms_reduce_kernel<kNmsMaxWordsPerLane><<<g, b, 0, stream>>>(mask, n, max_keep, keep, num_keep); break;
    }
    return check_launch("fi_nms_batched[reduce]");
}

}  // namespace fi

using namespace fi;

FI_API void _nms(int boxes_num, float *boxes_dev, unsigned long long *mask_dev, float nms_overlap_thresh) {
    if (boxes_num <= 0) { ok(); return; }
    dim3 grid(ceil_div(boxes_num, kNmsTile), ceil_div(boxes_num, kNmsTile), 1);
    nms_mask_kernel<<<grid, kNmsTile>>>(boxes_dev, boxes_num, nms_overlap_thresh, mask_dev, /*full=*/1);   // legacy stream, like nms_kernel.cu:79
    check_launch("_nms");
}

FI_API int fi_nms_batched_topk(const float *boxes, int n_images, int n, float thresh, int max_keep, unsigned long long *mask, int *keep,
                               int *num_keep, cudaStream_t stream) {
    FI_REQUIRE(n_images >= 0 && n >= 0 && n <= 32 * kNmsMaxWordsPerLane * kNmsTile, "fi_nms_batched: n=%d outside [0,%d]", n,
               32 * kNmsMaxWordsPerLane * kNmsTile);
    FI_REQUIRE(max_keep >= 1, "fi_nms_batched_topk: max_keep=%d", max_keep);
    if (n_images == 0) return ok();
    FI_REQUIRE(num_keep && (n == 0 || (boxes && mask && keep)), "fi_nms_batched: null pointer");
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_keep, 0, sizeof(int) * n_images, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_nms_batched: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
        return ok();
    }
    FI_REQUIRE(n_images <= 65535, "fi_nms_batched: too many images");
    const int t = ceil_div(n, kNmsTile);
    nms_mask_kernel<<<dim3(t, t, n_images), kNmsTile, 0, stream>>>(boxes, n, thresh, mask, /*full=*/0);
    if (int e = check_launch("fi_nms_batched[mask]")) return e;
    return launch_nms_reduce(mask, n_images, n, max_keep, keep, num_keep, stream);
}

FI_API int fi_nms_batched(const float *boxes, int n_images, int n, float thresh, unsigned long long *mask, int *keep, int *num_keep,
                          cudaStream_t stream) {
    return fi_nms_batched_topk(boxes, n_images, n, thresh, n > 0 ? n : 1, mask, keep, num_keep, stream);
}
