// Proposal layer front half (lib/layers.py:87-122): gather the top-scoring anchors, apply the box deltas
// (tools/box_utils.py:7-29), clip to the image window (tools/box_utils.py:32-45) -- one thread per (image, proposal),
// writing the boxes both as (y1,x1,y2,x2) for the caller and as (x1,y1,x2,y2,score) rows for fi_nms_batched, so that the
// reference's per-image index_select loops, five elementwise launches and the torch.cat for NMS become one launch.
// Arithmetic is the reference's op sequence with explicitly rounded fp32 operations (each torch op there rounds once).
#include "fi_common.cuh"

namespace fi {

__global__ void __launch_bounds__(256) proposal_decode_kernel(const float *__restrict__ deltas, const float *__restrict__ anchors,
                                                             const long long *__restrict__ order, const float *__restrict__ scores, int bs, int A,
                                                             int K, float s0, float s1, float s2, float s3, float win_h, float win_w,
                                                             float *__restrict__ boxes, float *__restrict__ dets) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= (long)bs * K) return;
    const int b = (int)(i / K);
    const long long a = order[i];
    if (a < 0 || a >= A) {                                  // cannot happen for a sort order; keep memory safe anyway
#pragma unroll
        for (int k = 0; k < 4; ++k) boxes[i * 4 + k] = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) dets[i * 5 + k] = 0.f;
        return;
    }
    const float4 an = *reinterpret_cast<const float4 *>(anchors + a * 4);                    // y1, x1, y2, x2
    const float4 dl = *reinterpret_cast<const float4 *>(deltas + ((long)b * A + a) * 4);      // dy, dx, log dh, log dw
    const float dy = __fmul_rn(dl.x, s0), dx = __fmul_rn(dl.y, s1), dh = __fmul_rn(dl.z, s2), dw = __fmul_rn(dl.w, s3);   // layers.py:96
    float height = __fsub_rn(an.z, an.x), width = __fsub_rn(an.w, an.y);                                                // box_utils.py:14-15
    float cy = __fadd_rn(an.x, __fmul_rn(0.5f, height)), cx = __fadd_rn(an.y, __fmul_rn(0.5f, width));                   // :16-17
    cy = __fadd_rn(cy, __fmul_rn(dy, height));                                                                           // :19
    cx = __fadd_rn(cx, __fmul_rn(dx, width));                                                                            // :20
    height = __fmul_rn(height, expf(dh));                                                                                // :21
    width = __fmul_rn(width, expf(dw));                                                                                  // :22
    float y1 = __fsub_rn(cy, __fmul_rn(0.5f, height)), x1 = __fsub_rn(cx, __fmul_rn(0.5f, width));                       // :24-25
    float y2 = __fadd_rn(y1, height), x2 = __fadd_rn(x1, width);                                                         // :26-27
    y1 = fminf(fmaxf(y1, 0.f), win_h); x1 = fminf(fmaxf(x1, 0.f), win_w);                                                // box_utils.py:39-44
    y2 = fminf(fmaxf(y2, 0.f), win_h); x2 = fminf(fmaxf(x2, 0.f), win_w);
    *reinterpret_cast<float4 *>(boxes + i * 4) = make_float4(y1, x1, y2, x2);
    float *d = dets + i * 5;
    d[0] = x1; d[1] = y1; d[2] = x2; d[3] = y2; d[4] = scores[i];                                                       // pth_nms.py:28-33
}

// Back half (lib/layers.py:128-139 + lib/nms/nms_wrapper.py:24-33): every image is truncated to the smallest keep count of
// the batch (and to proposal_count), the kept boxes are gathered and normalised; rows past that count are zero -- what
// the reference's later layers pad with anyway (lib/layers.py:413,427).  The count itself stays on the device.
__global__ void __launch_bounds__(256) proposal_gather_kernel(const float *__restrict__ boxes, const int *__restrict__ keep, const int *__restrict__ num,
                                                             int bs, int K, int count, float norm_h, float norm_w, float *__restrict__ out,
                                                             int *__restrict__ m_out) {
    int m = count;
    for (int b = 0; b < bs; ++b) m = min(m, num[b]);        // bs is a handful of images: every thread reads them (L1 broadcast)
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i == 0 && m_out) *m_out = m;
    if (i >= (long)bs * count) return;
    const int b = (int)(i / count), j = (int)(i - (long)b * count);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < m) {
        const int k = keep[(long)b * K + j];
        if (k >= 0 && k < K) {
            const float4 bx = *reinterpret_cast<const float4 *>(boxes + ((long)b * K + k) * 4);
            v = make_float4(__fdiv_rn(bx.x, norm_h), __fdiv_rn(bx.y, norm_w), __fdiv_rn(bx.z, norm_h), __fdiv_rn(bx.w, norm_w));   // boxes_keep / norm
        }
    }
    *reinterpret_cast<float4 *>(out + i * 4) = v;
}

}  // namespace fi

using namespace fi;

FI_API int fi_proposal_gather(const float *boxes, const int *keep, const int *num_keep, int batch, int num_proposals, int proposal_count,
                              float window_height, float window_width, float *rois, int *num_rois, cudaStream_t stream) {
    FI_REQUIRE(batch >= 0 && num_proposals >= 0 && proposal_count >= 0 && window_height > 0.f && window_width > 0.f, "fi_proposal_gather: bad sizes");
    if (batch == 0 || proposal_count == 0) return ok();
    FI_REQUIRE(boxes && keep && num_keep && rois, "fi_proposal_gather: null pointer");
    FI_REQUIRE(((uintptr_t)boxes % 16) == 0 && ((uintptr_t)rois % 16) == 0, "fi_proposal_gather: 16-byte aligned tensors");
    const long total = (long)batch * proposal_count;
    proposal_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(boxes, keep, num_keep, batch, num_proposals, proposal_count,
                                                                               window_height, window_width, rois, num_rois);
    return check_launch("fi_proposal_gather");
}

FI_API int fi_proposal_decode(const float *deltas, const float *anchors, const long long *order, const float *scores_sorted, int batch,
                              int num_anchors, int num_proposals, const float *std_dev4, float window_height, float window_width,
                              float *boxes, float *dets_xyxys, cudaStream_t stream) {
    FI_REQUIRE(batch >= 0 && num_anchors > 0 && num_proposals >= 0 && num_proposals <= num_anchors, "fi_proposal_decode: bad sizes bs=%d A=%d K=%d",
               batch, num_anchors, num_proposals);
    FI_REQUIRE(std_dev4, "fi_proposal_decode: std_dev4 (4 host floats) is required");
    if (batch == 0 || num_proposals == 0) return ok();
    FI_REQUIRE(deltas && anchors && order && scores_sorted && boxes && dets_xyxys, "fi_proposal_decode: null pointer");
    FI_REQUIRE(((uintptr_t)deltas % 16) == 0 && ((uintptr_t)anchors % 16) == 0 && ((uintptr_t)boxes % 16) == 0, "fi_proposal_decode: 16-byte aligned tensors");
    const long total = (long)batch * num_proposals;
    proposal_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(deltas, anchors, order, scores_sorted, batch, num_anchors, num_proposals,
                                                                               std_dev4[0], std_dev4[1], std_dev4[2], std_dev4[3], window_height,
                                                                               window_width, boxes, dets_xyxys);
    return check_launch("fi_proposal_decode");
}
