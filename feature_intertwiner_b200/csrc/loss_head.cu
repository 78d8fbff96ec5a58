// The pointwise steps of the class-level intertwiner loss head (lib/model.py:143-224 + lib/OT_module.py:67-102, the 1-D branch)
// as a handful of launches.
//
// Written in torch the head is ~80 kernels of a few microseconds each -- products, sums, transposes, slices, contiguous
// copies, concatenations, masks, and the same again backwards -- 0.36 ms per iteration at 80 classes x 1024 features,
// almost all of it launch-to-launch latency on a dependent chain (profiles/r02_launches_bench_v2_summary.json).  The dense
// products (80 x 1024 x 1024, 160 x 1024 x 256 and their gradients) stay library GEMMs; everything between them is here:
//   merge_stats      _merge_feat_vec numerators / denominators of both sets, straight into the buffer the all-reduce works on
//   ot_head_prep     final_small = sum / (n + EPS), comparison mask, and the [class, feature] transposes the critic wants
//   ot_head_combine  2 W(x^,y) - W(x^,x^) - W(y,y), masked                                   (OT_module.py:78-80)
//   ot_head_dcritic  gradient of that combination through the three Sinkhorn problems and the critic's ReLU
//   relu_mask        ReLU backward in place;  col_sum: the bias gradients
//   ot_head_dsum / merge_stats_bwd   back through the division and the transpose; through the all-reduce scale and the count weighting
//   centre_tap_embed gradient of W[:, :, 1] as the full [out, in, 3] Conv1d weight gradient (zeros elsewhere)
#include "fi_common.cuh"

namespace fi {

constexpr float kHeadEps = 1e-20f;          // EPS of lib/model.py

// packed = [ big_sum F*ncls | big_n ncls | small_sum F*ncls | small_n ncls ];  feat [GS,F,ncls], cnt [GS,ncls]
__global__ void merge_stats_kernel(const float *__restrict__ bf, const float *__restrict__ bc, const float *__restrict__ sf,
                                   const float *__restrict__ sc, int GS, int F, int ncls, float *__restrict__ packed) {
    const int total = F * ncls;
    float *bs = packed, *bn = packed + total, *ss = bn + ncls, *sn = ss + total;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int c = e % ncls;
        float a = 0.f, b = 0.f;
        for (int g = 0; g < GS; ++g) {                                        // (feat * cnt).sum over (gpu, scale), model.py:219-222
            a = __fadd_rn(a, __fmul_rn(bf[(long)g * total + e], bc[g * ncls + c]));
            b = __fadd_rn(b, __fmul_rn(sf[(long)g * total + e], sc[g * ncls + c]));
        }
        bs[e] = a;
        ss[e] = b;
        if (e < ncls) {
            float na = 0.f, nb = 0.f;
            for (int g = 0; g < GS; ++g) { na = __fadd_rn(na, bc[g * ncls + e]); nb = __fadd_rn(nb, sc[g * ncls + e]); }
            bn[e] = na;
            sn[e] = nb;
        }
    }
}

// X[c-1, f] = small_sum[f, c] / (small_n[c] + EPS), Y[c-1, f] = final_big[f, c]  (c = 1 .. ncls-1: background excluded, model.py:178)
// mask[c-1] = small_n[c] > 0 and the class is in the buffer (model.py:179-186)
__global__ void ot_head_prep_kernel(const float *__restrict__ final_big, const float *__restrict__ ss, const float *__restrict__ sn,
                                    const float *__restrict__ buffer_cnt, int B, int F, int ncls, float *__restrict__ X, float *__restrict__ Y,
                                    float *__restrict__ mask) {
    const int n = ncls - 1;
    const long total = (long)n * F;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int p = (int)(e / F), f = (int)(e - (long)p * F), c = p + 1;
        X[e] = __fdiv_rn(ss[(long)f * ncls + c], __fadd_rn(sn[c], kHeadEps));
        Y[e] = final_big[(long)f * ncls + c];
        if (f == 0) {
            float in_buf = 0.f;
            for (int b = 0; b < B; ++b) in_buf = __fadd_rn(in_buf, buffer_cnt[b * ncls + c]);
            mask[p] = (sn[c] > 0.f && in_buf > 0.f) ? 1.f : 0.f;
        }
    }
}

__global__ void ot_head_combine_kernel(const float *__restrict__ w, const float *__restrict__ mask, int n, float *__restrict__ loss) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) loss[p] = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(2.f, w[p]), w[n + p]), w[2 * n + p]), mask[p]);
}

// Sinkhorn problems [0,n): (cx, cy); [n,2n): (cx, cx); [2n,3n): (cy, cy).  gx / gy [3n, N]: their gradients for unit upstream.
// dC[0:n] = d loss / d cx, dC[n:2n] = d loss / d cy, through the ReLU that produced Cc = [cx; cy].
__global__ void ot_head_dcritic_kernel(const float *__restrict__ gx, const float *__restrict__ gy, const float *__restrict__ g,
                                       const float *__restrict__ mask, const float *__restrict__ Cc, int n, int N, float *__restrict__ dC) {
    const long total = (long)n * N;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int p = (int)(e / N);
        const float gm = __fmul_rn(g[p], mask[p]);
        const float two = __fmul_rn(2.f, gm), neg = -gm;
        const float dx = __fadd_rn(__fadd_rn(__fmul_rn(gx[e], two), __fmul_rn(gx[total + e], neg)), __fmul_rn(gy[total + e], neg));
        const float dy = __fadd_rn(__fadd_rn(__fmul_rn(gy[e], two), __fmul_rn(gx[2 * total + e], neg)), __fmul_rn(gy[2 * total + e], neg));
        dC[e] = Cc[e] > 0.f ? dx : 0.f;
        dC[total + e] = Cc[total + e] > 0.f ? dy : 0.f;
    }
}

__global__ void relu_mask_kernel(float *__restrict__ d, const float *__restrict__ h, long count) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x)
        if (!(h[e] > 0.f)) d[e] = 0.f;
}

// d small_sum[f, c] = dX[c-1, f] / (small_n[c] + EPS)   (0 for the background column): back through the division and the transpose
__global__ void ot_head_dsum_kernel(const float *__restrict__ dX, const float *__restrict__ sn, int F, int ncls, float *__restrict__ dss) {
    const int total = F * ncls;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int f = e / ncls, c = e - f * ncls;
        dss[e] = c > 0 ? __fdiv_rn(dX[(long)(c - 1) * F + f], __fadd_rn(sn[c], kHeadEps)) : 0.f;
    }
}

// d small_feat[g, f, c] = d small_sum[f, c] * scale * small_cnt[g, c]: back through the all-reduce (scale = world size when the
// caller compensates DDP's gradient averaging, dist.py::_AllReduceSum) and the count weighting
__global__ void merge_stats_bwd_kernel(const float *__restrict__ dss, const float *__restrict__ sc, int GS, int F, int ncls, float scale,
                                       float *__restrict__ dsf) {
    const int total = F * ncls;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int c = e % ncls;
        const float d = __fmul_rn(dss[e], scale);
        for (int g = 0; g < GS; ++g) dsf[(long)g * total + e] = __fmul_rn(d, sc[g * ncls + c]);
    }
}

// dst[c] = sum over rows of src[r, c] (bias gradients: a few hundred rows at most).  CTA = 32 columns x 8 row groups: warp w sums
// rows w, w + 8, ... (coalesced across its 32 columns, 4 loads in flight), the 8 partial sums meet in shared memory in a fixed order.
__global__ void __launch_bounds__(256) col_sum_kernel(const float *__restrict__ src, int rows, int cols, float *__restrict__ dst) {
    __shared__ float part[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < cols) {
        int r = w;
        for (; r + 24 < rows; r += 32) {
            a0 += src[(long)r * cols + c]; a1 += src[(long)(r + 8) * cols + c];
            a2 += src[(long)(r + 16) * cols + c]; a3 += src[(long)(r + 24) * cols + c];
        }
        for (; r < rows; r += 8) a0 += src[(long)r * cols + c];
    }
    part[w][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (w == 0 && c < cols) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += part[q][lane];
        dst[c] = s;
    }
}

__global__ void centre_tap_embed_kernel(const float *__restrict__ w1, long count, float *__restrict__ full) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x) {
        full[3 * e + 0] = 0.f;
        full[3 * e + 1] = w1[e];
        full[3 * e + 2] = 0.f;
    }
}

static int grid_1d(long n, int block) {
    long g = (n + block - 1) / block;
    const long cap = (long)kNumSMs * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace fi

using namespace fi;

FI_API int fi_merge_stats(const float *big_feat, const float *big_cnt, const float *small_feat, const float *small_cnt, int GS, int F, int ncls,
                          float *packed, cudaStream_t stream) {
    FI_REQUIRE(GS >= 1 && F > 0 && ncls > 0 && big_feat && big_cnt && small_feat && small_cnt && packed, "fi_merge_stats: bad arguments");
    merge_stats_kernel<<<grid_1d((long)F * ncls, 256), 256, 0, stream>>>(big_feat, big_cnt, small_feat, small_cnt, GS, F, ncls, packed);
    return check_launch("fi_merge_stats");
}

FI_API int fi_ot_head_prep(const float *final_big, const float *small_sum, const float *small_n, const float *buffer_cnt, int B, int F, int ncls,
                           float *X, float *Y, float *mask, cudaStream_t stream) {
    FI_REQUIRE(B >= 1 && F > 0 && ncls > 1 && final_big && small_sum && small_n && buffer_cnt && X && Y && mask, "fi_ot_head_prep: bad arguments");
    ot_head_prep_kernel<<<grid_1d((long)(ncls - 1) * F, 256), 256, 0, stream>>>(final_big, small_sum, small_n, buffer_cnt, B, F, ncls, X, Y, mask);
    return check_launch("fi_ot_head_prep");
}

FI_API int fi_ot_head_combine(const float *w, const float *mask, int n, float *loss, cudaStream_t stream) {
    FI_REQUIRE(n > 0 && w && mask && loss, "fi_ot_head_combine: bad arguments");
    ot_head_combine_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(w, mask, n, loss);
    return check_launch("fi_ot_head_combine");
}

FI_API int fi_ot_head_dcritic(const float *gx, const float *gy, const float *g, const float *mask, const float *Cc, int n, int N, float *dC,
                              cudaStream_t stream) {
    FI_REQUIRE(n > 0 && N > 0 && gx && gy && g && mask && Cc && dC, "fi_ot_head_dcritic: bad arguments");
    ot_head_dcritic_kernel<<<grid_1d((long)n * N, 256), 256, 0, stream>>>(gx, gy, g, mask, Cc, n, N, dC);
    return check_launch("fi_ot_head_dcritic");
}

FI_API int fi_relu_mask(float *d, const float *h, long count, cudaStream_t stream) {
    FI_REQUIRE(count >= 0 && (count == 0 || (d && h)), "fi_relu_mask: bad arguments");
    if (count == 0) return ok();
    relu_mask_kernel<<<grid_1d(count, 256), 256, 0, stream>>>(d, h, count);
    return check_launch("fi_relu_mask");
}

FI_API int fi_ot_head_dsum(const float *dX, const float *small_n, int F, int ncls, float *d_small_sum, cudaStream_t stream) {
    FI_REQUIRE(F > 0 && ncls > 1 && dX && small_n && d_small_sum, "fi_ot_head_dsum: bad arguments");
    ot_head_dsum_kernel<<<grid_1d((long)F * ncls, 256), 256, 0, stream>>>(dX, small_n, F, ncls, d_small_sum);
    return check_launch("fi_ot_head_dsum");
}

FI_API int fi_merge_stats_backward(const float *d_small_sum, const float *small_cnt, int GS, int F, int ncls, float scale, float *d_small_feat,
                                   cudaStream_t stream) {
    FI_REQUIRE(GS >= 1 && F > 0 && ncls > 0 && d_small_sum && small_cnt && d_small_feat, "fi_merge_stats_backward: bad arguments");
    merge_stats_bwd_kernel<<<grid_1d((long)F * ncls, 256), 256, 0, stream>>>(d_small_sum, small_cnt, GS, F, ncls, scale, d_small_feat);
    return check_launch("fi_merge_stats_backward");
}

FI_API int fi_col_sum(const float *src, int rows, int cols, float *dst, cudaStream_t stream) {
    FI_REQUIRE(rows >= 0 && cols > 0 && src && dst, "fi_col_sum: bad arguments");
    col_sum_kernel<<<ceil_div(cols, 32), 256, 0, stream>>>(src, rows, cols, dst);
    return check_launch("fi_col_sum");
}

FI_API int fi_centre_tap_embed(const float *w1, long count, float *full, cudaStream_t stream) {
    FI_REQUIRE(count >= 0 && (count == 0 || (w1 && full)), "fi_centre_tap_embed: bad arguments");
    if (count == 0) return ok();
    centre_tap_embed_kernel<<<grid_1d(count, 256), 256, 0, stream>>>(w1, count, full);
    return check_launch("fi_centre_tap_embed");
}
