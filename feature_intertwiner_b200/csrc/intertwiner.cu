// The intertwiner's bookkeeping kernels: RoI -> pyramid level, reliable / less-reliable split,
// per-class segment mean, historical buffer update.  In the reference each of these is a chain of tiny
// torch ops with host round trips (`.any()`, `nonzero`, python loops over classes: lib/sub_module.py:397-493,
// 664-684; lib/model.py:148-166); here each is ONE launch and nothing comes back to the host.
#include "fi_common.cuh"

namespace fi {

// ------------------------------------------------------------------------------------------------
// Level rule.  lib/sub_module.py:397-410; log2(x) = log(x)/log(2) in fp32 (tools/utils.py:50-55).
// Same libdevice logf/sqrtf and IEEE division torch's own CUDA kernels use, no contraction possible.
// ------------------------------------------------------------------------------------------------
__global__ void roi_level_kernel(const float *__restrict__ rois, int n, float denom, int *__restrict__ level) {
    const float kLn2 = 0.693147182464599609375f;   // fp32 log(2.0f): torch.log(FloatTensor([2.0]))
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 r = __ldg(reinterpret_cast<const float4 *>(rois) + i);   // (y1,x1,y2,x2)
        const float h = __fsub_rn(r.z, r.x), w = __fsub_rn(r.w, r.y);
        const float area = __fmul_rn(w, h);
        const float v = __fadd_rn(4.f, __fdiv_rn(logf(__fdiv_rn(sqrtf(area), denom)), kLn2));
        // .round() is half-to-even; .int() of -inf (zero-padded RoI) / NaN saturates, then clamp(2,5)
        int l = __float2int_rz(rintf(v));
        level[i] = min(max(l, 2), 5);
    }
}

// ------------------------------------------------------------------------------------------------
// Split.  One 1024-thread block; thread t owns the contiguous slice [t*per, (t+1)*per) of the flat RoI
// array, so concatenating the threads' outputs in thread order reproduces torch.nonzero's row-major
// order.  Eight lists at once: small(l) = {level == l}, big(l) = {level > l}, l = 2..5.
// ------------------------------------------------------------------------------------------------
constexpr int kSplitThreads = 1024;
constexpr int kLists = 8;

// Optional gather outputs (all [4, n, ...], same order as the index lists) so the host does not have to run
// rois[idx], idx // R, gt[idx] as separate torch ops per level:
struct SplitGather {
    const int *order;          // optional visiting order (a permutation of 0..n-1): lists come out in THIS order
    const float4 *rois;        // [n] (y1,x1,y2,x2)
    const int *gt;             // [n] class ids or null
    int rois_per_image;
    float4 *small_boxes, *big_boxes;
    int *small_ind, *big_ind, *small_gt, *big_gt;
    int *img_cnt;              // optional [8, num_images]: members of list k that belong to image b (lists 0-3 small, 4-7 big)
    int num_images;
};

__global__ void __launch_bounds__(kSplitThreads) split_levels_kernel(const int *__restrict__ level, int n, int *__restrict__ small_idx,
                                                                    int *__restrict__ small_cnt, int *__restrict__ big_idx,
                                                                    int *__restrict__ big_cnt, int *__restrict__ slot, const SplitGather G) {
    __shared__ int warp_tot[kLists][kSplitThreads / 32];
    __shared__ int list_tot[kLists];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (G.img_cnt) for (int q = t; q < kLists * G.num_images; q += kSplitThreads) G.img_cnt[q] = 0;   // ordered before the atomics by the barriers below
    const int per = (n + kSplitThreads - 1) / kSplitThreads;
    const int lo = min(t * per, n), hi = min(lo + per, n);
    int cnt[kLists];
#pragma unroll
    for (int k = 0; k < kLists; ++k) cnt[k] = 0;
    for (int p = lo; p < hi; ++p) {
        const int l = level[G.order ? G.order[p] : p];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            cnt[k] += (l == 2 + k);
            cnt[4 + k] += (l > 2 + k);
        }
    }
    int base[kLists];
#pragma unroll
    for (int k = 0; k < kLists; ++k) {
        int v = cnt[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += u;
        }
        if (lane == 31) warp_tot[k][wid] = v;
        base[k] = v - cnt[k];                      // exclusive prefix inside the warp
    }
    __syncthreads();
    if (wid < kLists) {                            // warp k scans the 32 warp totals of list k
        int v = warp_tot[wid][lane];
        const int own = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += u;
        }
        warp_tot[wid][lane] = v - own;
        if (lane == 31) list_tot[wid] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kLists; ++k) base[k] += warp_tot[k][wid];
    for (int p = lo; p < hi; ++p) {
        const int i = G.order ? G.order[p] : p;
        const int l = level[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (l == 2 + k) {
                const int o = k * n + base[k];
                small_idx[o] = i; slot[i] = base[k]; ++base[k];
                if (G.img_cnt) atomicAdd(&G.img_cnt[k * G.num_images + i / G.rois_per_image], 1);
                if (G.rois) { G.small_boxes[o] = G.rois[i]; G.small_ind[o] = i / G.rois_per_image; if (G.gt) G.small_gt[o] = G.gt[i]; }
            }
            if (l > 2 + k) {
                const int o = k * n + base[4 + k];
                big_idx[o] = i; ++base[4 + k];
                if (G.img_cnt) atomicAdd(&G.img_cnt[(4 + k) * G.num_images + i / G.rois_per_image], 1);
                if (G.rois) { G.big_boxes[o] = G.rois[i]; G.big_ind[o] = i / G.rois_per_image; if (G.gt) G.big_gt[o] = G.gt[i]; }
            }
        }
    }
    if (t < 4) { small_cnt[t] = list_tot[t]; big_cnt[t] = list_tot[4 + t]; }
}

// ------------------------------------------------------------------------------------------------
// Segment mean.  Block (class c, 128-feature chunk): stages gt in shared memory 1024 ids at a time and
// accumulates the rows of class c in ascending row order (deterministic).  mean is [F, ncls].
// ------------------------------------------------------------------------------------------------
constexpr int kSegThreads = 128;
constexpr int kSegStage = 1024;

// Up to kSegSets (gt, feat) lists in one launch (blockIdx.z): Dev.forward takes the class means of 3 reliable + 3 less-reliable
// lists per pass, each a launch-latency sized problem.
constexpr int kSegSets = 8;
struct SegSet {
    const int *gt, *k_dev;
    const float *feat, *gmean, *cntp;
    float *mean, *cnt, *gfeat;
    int k;
};
struct SegSets { SegSet s[kSegSets]; };

__global__ void __launch_bounds__(kSegThreads) segment_mean_fwd_kernel(const SegSets sets, int F, int ncls) {
    __shared__ int rows[kSegStage];
    __shared__ int nrows;
    const SegSet &S = sets.s[blockIdx.z];
    const int *__restrict__ gt = S.gt;
    const float *__restrict__ feat = S.feat;
    float *__restrict__ mean = S.mean, *__restrict__ cnt = S.cnt;
    int k = S.k;
    if (S.k_dev) k = min(k, *S.k_dev);             // list length kept on the device: k is the capacity
    const int c = blockIdx.x;
    const int f = blockIdx.y * kSegThreads + threadIdx.x;
    float acc = 0.f;
    int total = 0;
    if (c != 0) {                                  // background never contributes (sub_module.py:677-678)
        for (int s0 = 0; s0 < k; s0 += kSegStage) {
            if (threadIdx.x == 0) nrows = 0;
            __syncthreads();
            // ordered compaction of this stage's matching rows: one warp-ballot pass per 128 ids
            for (int off = 0; off < kSegStage; off += kSegThreads) {
                const int i = s0 + off + threadIdx.x;
                const bool hit = (i < k) && (gt[i] == c);
                __shared__ int warp_hits[kSegThreads / 32];
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if ((threadIdx.x & 31) == 0) warp_hits[threadIdx.x >> 5] = __popc(m);
                __syncthreads();
                int before = nrows;
                for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += warp_hits[w];
                if (hit) rows[before + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = i;
                __syncthreads();
                if (threadIdx.x == 0) {
                    int s = 0;
                    for (int w = 0; w < kSegThreads / 32; ++w) s += warp_hits[w];
                    nrows += s;
                }
                __syncthreads();
            }
            const int nr = nrows;
            total += nr;
            if (f < F) {
                int q = 0;
                for (; q + 4 <= nr; q += 4) {      // 4 independent loads in flight
                    const float v0 = __ldg(feat + (long)rows[q] * F + f), v1 = __ldg(feat + (long)rows[q + 1] * F + f);
                    const float v2 = __ldg(feat + (long)rows[q + 2] * F + f), v3 = __ldg(feat + (long)rows[q + 3] * F + f);
                    acc += v0; acc += v1; acc += v2; acc += v3;
                }
                for (; q < nr; ++q) acc += __ldg(feat + (long)rows[q] * F + f);
            }
            __syncthreads();
        }
    }
    if (f < F) mean[(long)f * ncls + c] = total > 0 ? __fdiv_rn(acc, (float)total) : 0.f;
    if (blockIdx.y == 0 && threadIdx.x == 0) cnt[c] = (float)total;
}

__global__ void segment_mean_bwd_kernel(const SegSets sets, int F, int ncls) {
    const SegSet &S = sets.s[blockIdx.y];
    const int *__restrict__ gt = S.gt;
    const float *__restrict__ gmean = S.gmean, *__restrict__ cnt = S.cntp;
    float *__restrict__ gfeat = S.gfeat;
    const int k = S.k;
    const long total = (long)k * F;
    const int live = S.k_dev ? min(k, *S.k_dev) : k;   // rows past the device-side length get a zero gradient
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int i = (int)(e / F), f = (int)(e - (long)i * F);
        const int c = i < live ? gt[i] : 0;
        float v = 0.f;
        if (c > 0 && c < ncls) v = __fdiv_rn(__ldg(gmean + (long)f * ncls + c), cnt[c]);
        gfeat[e] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Buffer update.  lib/model.py:148-166.  EPS = 1e-20 as in the reference.  B == 1: count-weighted running
// mean over all history.  B > 1: the reference shifts the whole [B,F,ncls] buffer left by one slot every
// iteration; here the buffer is a ring (the caller passes the slot to overwrite) and the weighted mean is
// accumulated oldest -> newest, the order the reference's torch.sum(dim=0) walks its shifted copy.
// ------------------------------------------------------------------------------------------------
__global__ void buffer_update_kernel(const float *__restrict__ big_sum, const float *__restrict__ big_n, int B, int slot, int F, int ncls,
                                     float *__restrict__ buffer, float *__restrict__ buffer_cnt, float *__restrict__ final_big) {
    const float kEps = 1e-20f;
    const int total = F * ncls;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int c = e % ncls;
        const float n = big_n[c];
        const float m = __fdiv_rn(big_sum[e], __fadd_rn(n, kEps));      // _merge_feat_vec (model.py:217-224)
        if (B == 1) {
            const float bc = buffer_cnt[c];
            const float s = __fadd_rn(__fmul_rn(buffer[e], bc), __fmul_rn(m, n));
            const float v = __fdiv_rn(s, __fadd_rn(__fadd_rn(bc, n), kEps));
            buffer[e] = v;
            final_big[e] = v;
        } else {
            buffer[(long)slot * total + e] = m;
            float s = 0.f, cn = 0.f;
            for (int q = 1; q <= B; ++q) {
                const int b = (slot + q) % B;                         // oldest first, `slot` (newest) last
                const float bc = (b == slot) ? n : buffer_cnt[b * ncls + c];
                s = __fadd_rn(s, __fmul_rn(buffer[(long)b * total + e], bc));
                cn = __fadd_rn(cn, bc);
            }
            final_big[e] = __fdiv_rn(s, __fadd_rn(cn, kEps));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Visiting order of the RoIs (intertwiner.py::spatial_order): every image walked coarse tile by coarse tile (grid x grid, by box
// centre, boustrophedon), stable.  The rank of an element is counted directly (R <= 4096 per image, keys in shared memory) -- replaces ~10 pointwise launches and a radix sort.
// ------------------------------------------------------------------------------------------------
constexpr int kOrderMaxR = 4096;
constexpr int kOrderPerCta = 32;      // RoIs ranked by one CTA (8 warps x 4): grid = (images, ceil(R / 32))
__global__ void __launch_bounds__(256) spatial_order_kernel(const float4 *__restrict__ rois, int R, int grid, int *__restrict__ order) {
    __shared__ short keys[kOrderMaxR];
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float g2 = __fmul_rn(0.5f, (float)grid), top = (float)(grid - 1);
    for (int i = threadIdx.x; i < R; i += blockDim.x) {             // every CTA of the image forms all keys (R float4 loads from L2)
        const float4 r = __ldg(rois + (long)b * R + i);               // y1, x1, y2, x2
        const float cy = floorf(fminf(fmaxf(__fmul_rn(__fadd_rn(r.x, r.z), g2), 0.f), top));
        const float cx = floorf(fminf(fmaxf(__fmul_rn(__fadd_rn(r.y, r.w), g2), 0.f), top));
        const int iy = (int)cy, ix = (int)cx;
        keys[i] = (short)(iy * grid + ((iy & 1) ? grid - 1 - ix : ix));
    }
    __syncthreads();
    // warp per RoI, lanes stride over the others (one CTA per image walked R^2 / 256 comparisons per thread: 53 us at R = 1000,
    // 0.2 ms at R = 2000 on 4 of 148 SMs)
    for (int q = 0; q < kOrderPerCta / 8; ++q) {
        const int i = blockIdx.y * kOrderPerCta + wid * (kOrderPerCta / 8) + q;
        if (i >= R) break;
        const int k = keys[i];
        int rank = 0;
        for (int j = lane; j < R; j += 32) {
            const int kj = keys[j];
            rank += (kj < k) || (kj == k && j < i);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, d);
        if (lane == 0) order[(long)b * R + rank] = b * R + i;
    }
}

// counts are updated after every feature element has read the old value
__global__ void buffer_cnt_kernel(const float *__restrict__ big_n, int B, int slot, int ncls, float *__restrict__ buffer_cnt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncls) return;
    if (B == 1) buffer_cnt[c] = __fadd_rn(buffer_cnt[c], big_n[c]);
    else buffer_cnt[slot * ncls + c] = big_n[c];
}

}  // namespace fi

using namespace fi;

FI_API int fi_roi_level(const float *rois, int n, float image_area, float base, int *level, cudaStream_t stream) {
    FI_REQUIRE(n >= 0 && image_area > 0.f && base > 0.f, "fi_roi_level: bad arguments");
    if (n == 0) return ok();
    FI_REQUIRE(rois && level && ((uintptr_t)rois % 16 == 0), "fi_roi_level: rois must be a 16-byte aligned device pointer");
    const float denom = base / sqrtf(image_area);                     // base / torch.sqrt(image_area), fp32
    int grid = ceil_div(n, 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    roi_level_kernel<<<grid, 256, 0, stream>>>(rois, n, denom, level);
    return check_launch("fi_roi_level");
}

FI_API int fi_split_levels(const int *level, int n, int *small_idx, int *small_cnt, int *big_idx, int *big_cnt, int *slot,
                           cudaStream_t stream) {
    FI_REQUIRE(n >= 0 && n <= 65536, "fi_split_levels: n=%d outside [0,65536]", n);
    FI_REQUIRE(small_cnt && big_cnt && (n == 0 || (level && small_idx && big_idx && slot)), "fi_split_levels: null pointer");
    SplitGather G = {};
    split_levels_kernel<<<1, kSplitThreads, 0, stream>>>(level, n, small_idx, small_cnt, big_idx, big_cnt, slot, G);
    return check_launch("fi_split_levels");
}

FI_API int fi_split_levels_gather(const int *level, const float *rois, const int *gt, const int *order, int n, int rois_per_image, int *small_idx, int *small_cnt,
                                  int *big_idx, int *big_cnt, int *slot, float *small_boxes, int *small_ind, int *small_gt, float *big_boxes,
                                  int *big_ind, int *big_gt, int *img_cnt, cudaStream_t stream) {
    FI_REQUIRE(n >= 0 && n <= 65536 && rois_per_image > 0, "fi_split_levels_gather: n=%d outside [0,65536] or bad rois_per_image", n);
    FI_REQUIRE(small_cnt && big_cnt && (n == 0 || (level && rois && small_idx && big_idx && slot && small_boxes && small_ind && big_boxes && big_ind)),
               "fi_split_levels_gather: null pointer");
    FI_REQUIRE(gt == nullptr || (small_gt && big_gt), "fi_split_levels_gather: gt given without small_gt / big_gt");
    FI_REQUIRE(((uintptr_t)rois % 16 == 0) && ((uintptr_t)small_boxes % 16 == 0) && ((uintptr_t)big_boxes % 16 == 0), "fi_split_levels_gather: box arrays must be 16-byte aligned");
    SplitGather G;
    G.order = order;
    G.rois = reinterpret_cast<const float4 *>(rois); G.gt = gt; G.rois_per_image = rois_per_image;
    G.small_boxes = reinterpret_cast<float4 *>(small_boxes); G.big_boxes = reinterpret_cast<float4 *>(big_boxes);
    G.small_ind = small_ind; G.big_ind = big_ind; G.small_gt = small_gt; G.big_gt = big_gt;
    G.img_cnt = img_cnt; G.num_images = (n + rois_per_image - 1) / rois_per_image;
    split_levels_kernel<<<1, kSplitThreads, 0, stream>>>(level, n, small_idx, small_cnt, big_idx, big_cnt, slot, G);
    return check_launch("fi_split_levels_gather");
}

FI_API int fi_segment_mean_forward_batch(const fi_seg_set *sets, int num_sets, int F, int ncls, cudaStream_t stream) {
    FI_REQUIRE(sets && num_sets >= 1 && num_sets <= kSegSets && F > 0 && ncls > 0 && ncls <= 1024, "fi_segment_mean_forward: 1..%d lists, F > 0, ncls in [1,1024]", kSegSets);
    SegSets dev = {};
    for (int i = 0; i < num_sets; ++i) {
        const fi_seg_set &h = sets[i];
        FI_REQUIRE(h.k >= 0 && h.mean && h.cnt && (h.k == 0 || (h.gt && h.feat)), "fi_segment_mean_forward: bad list %d", i);
        dev.s[i].gt = h.gt; dev.s[i].k_dev = h.k_dev; dev.s[i].feat = h.feat; dev.s[i].mean = h.mean; dev.s[i].cnt = h.cnt; dev.s[i].k = h.k;
    }
    dim3 grid(ncls, ceil_div(F, kSegThreads), num_sets);
    segment_mean_fwd_kernel<<<grid, kSegThreads, 0, stream>>>(dev, F, ncls);
    return check_launch("fi_segment_mean_forward");
}
FI_API int fi_segment_mean_forward_n(const int *gt, const float *feat, int k, const int *k_dev, int F, int ncls, float *mean, float *cnt,
                                     cudaStream_t stream) {
    fi_seg_set one = {gt, feat, k, k_dev, mean, cnt, nullptr, nullptr};
    return fi_segment_mean_forward_batch(&one, 1, F, ncls, stream);
}
FI_API int fi_segment_mean_forward(const int *gt, const float *feat, int k, int F, int ncls, float *mean, float *cnt, cudaStream_t stream) {
    return fi_segment_mean_forward_n(gt, feat, k, nullptr, F, ncls, mean, cnt, stream);
}

/* backward of every list: grad_feat rows from grad_mean (h.grad_mean) and the forward's counts (h.cnt) */
FI_API int fi_segment_mean_backward_batch(const fi_seg_set *sets, int num_sets, int F, int ncls, cudaStream_t stream) {
    FI_REQUIRE(sets && num_sets >= 1 && num_sets <= kSegSets && F > 0 && ncls > 0, "fi_segment_mean_backward: 1..%d lists", kSegSets);
    SegSets dev = {};
    int max_k = 0;
    for (int i = 0; i < num_sets; ++i) {
        const fi_seg_set &h = sets[i];
        FI_REQUIRE(h.k >= 0 && (h.k == 0 || (h.gt && h.grad_mean && h.cnt && h.grad_feat)), "fi_segment_mean_backward: bad list %d", i);
        dev.s[i].gt = h.gt; dev.s[i].k_dev = h.k_dev; dev.s[i].gmean = h.grad_mean; dev.s[i].cntp = h.cnt; dev.s[i].gfeat = h.grad_feat; dev.s[i].k = h.k;
        max_k = max_k > h.k ? max_k : h.k;
    }
    if (max_k == 0) return ok();
    long grid = ((long)max_k * F + 255) / 256;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    segment_mean_bwd_kernel<<<dim3((unsigned)grid, num_sets), 256, 0, stream>>>(dev, F, ncls);
    return check_launch("fi_segment_mean_backward");
}
FI_API int fi_segment_mean_backward_n(const int *gt, const float *grad_mean, const float *cnt, int k, const int *k_dev, int F, int ncls,
                                      float *grad_feat, cudaStream_t stream) {
    FI_REQUIRE(k >= 0 && F > 0 && ncls > 0, "fi_segment_mean_backward: bad arguments");
    if (k == 0) return ok();
    fi_seg_set one = {gt, nullptr, k, k_dev, nullptr, const_cast<float *>(cnt), grad_mean, grad_feat};
    return fi_segment_mean_backward_batch(&one, 1, F, ncls, stream);
}
FI_API int fi_segment_mean_backward(const int *gt, const float *grad_mean, const float *cnt, int k, int F, int ncls, float *grad_feat,
                                    cudaStream_t stream) {
    return fi_segment_mean_backward_n(gt, grad_mean, cnt, k, nullptr, F, ncls, grad_feat, stream);
}

FI_API int fi_buffer_update(const float *big_sum, const float *big_n, int B, int slot, int F, int ncls, float *buffer,
                            float *buffer_cnt, float *final_big, cudaStream_t stream) {
    FI_REQUIRE(B >= 1 && F > 0 && ncls > 0 && slot >= 0 && slot < B, "fi_buffer_update: bad arguments");
    FI_REQUIRE(big_sum && big_n && buffer && buffer_cnt && final_big, "fi_buffer_update: null pointer");
    int grid = ceil_div(F * ncls, 256);
    if (grid > kNumSMs * 4) grid = kNumSMs * 4;
    buffer_update_kernel<<<grid, 256, 0, stream>>>(big_sum, big_n, B, slot, F, ncls, buffer, buffer_cnt, final_big);
    if (int e = check_launch("fi_buffer_update")) return e;
    buffer_cnt_kernel<<<ceil_div(ncls, 128), 128, 0, stream>>>(big_n, B, slot, ncls, buffer_cnt);
    return check_launch("fi_buffer_update[cnt]");
}

FI_API int fi_spatial_order(const float *rois, int batch, int rois_per_image, int grid, int *order, cudaStream_t stream) {
    FI_REQUIRE(batch >= 0 && rois_per_image >= 0 && grid >= 1 && grid <= 128, "fi_spatial_order: bad sizes");
    if (batch == 0 || rois_per_image == 0) return ok();
    if (rois_per_image > kOrderMaxR) { set_error(FI_ERR_UNSUPPORTED, "fi_spatial_order: more than %d RoIs per image", kOrderMaxR); return FI_ERR_UNSUPPORTED; }
    FI_REQUIRE(rois && order && ((uintptr_t)rois % 16) == 0, "fi_spatial_order: rois must be a 16-byte aligned device pointer");
    spatial_order_kernel<<<dim3((unsigned)batch, (unsigned)ceil_div(rois_per_image, kOrderPerCta)), 256, 0, stream>>>(
        reinterpret_cast<const float4 *>(rois), rois_per_image, grid, order);
    return check_launch("fi_spatial_order");
}
