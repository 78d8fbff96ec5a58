// The two remaining consumers of the RoIAlign / NMS kernels in the detector:
//
//  * mask targets of the positive RoIs (lib/layers.py:296-323): the reference gathers one GT mask per positive RoI
//    (gt_masks[assignment]), rewrites the RoI into the mask's own frame (USE_MINI_MASK), calls CropAndResizeFunction with
//    C = 1 and one "image" per box, and rounds.  Here: one launch, thread per target pixel, straight from gt_masks (no
//    gathered copy), same fp32 operations in the same order (each torch op there rounds once), same bilinear as the forward
//    kernels (fi_common.cuh::axis_sample), torch.round = half-to-even.
//  * the decode half of detection_layer (lib/layers.py:738-770): per RoI arg-max class, its class-specific deltas * std_dev,
//    apply_box_deltas (tools/box_utils.py:7-29), scale to pixels, clip to the image window, round, keep flag -- one launch
//    instead of ~25 pointwise ones.  The per-class NMS itself is the batched kernel of nms.cu on class-offset boxes (nms.py).
#include "fi_common.cuh"

namespace fi {

__global__ void __launch_bounds__(256) mask_target_kernel(const float *__restrict__ gt_masks, const float *__restrict__ pos_rois,
                                                         const float *__restrict__ gt_boxes, const int *__restrict__ assign, int n, int G, int mh,
                                                         int mw, int MH, int MW, int mini, float *__restrict__ out) {
    const long total = (long)n * MH * MW;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int ox = (int)(e % MW);
        const int oy = (int)((e / MW) % MH);
        const int i = (int)(e / ((long)MW * MH));
        const int a = assign[i];
        float v = 0.f;
        if (a >= 0 && a < G) {
            const float4 r = __ldg(reinterpret_cast<const float4 *>(pos_rois) + i);          // y1, x1, y2, x2
            float y1 = r.x, x1 = r.y, y2 = r.z, x2 = r.w;
            if (mini) {                                                                        // layers.py:304-313
                const float4 gb = __ldg(reinterpret_cast<const float4 *>(gt_boxes) + a);
                const float gt_h = __fsub_rn(gb.z, gb.x), gt_w = __fsub_rn(gb.w, gb.y);
                y1 = __fdiv_rn(__fsub_rn(y1, gb.x), gt_h); x1 = __fdiv_rn(__fsub_rn(x1, gb.y), gt_w);
                y2 = __fdiv_rn(__fsub_rn(y2, gb.x), gt_h); x2 = __fdiv_rn(__fsub_rn(x2, gb.y), gt_w);
            }
            const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, mh, MH), oy, mh, MH);
            const AxisTap tx = axis_sample(x1, x2, axis_step(x1, x2, mw, MW), ox, mw, MW);
            if (ty.inside && tx.inside) {                                                      // extrapolation value 0 (crop_and_resize.py:16)
                const float *m = gt_masks + (long)a * mh * mw;
                const float top = lerp_rn(__ldg(m + (long)ty.lo * mw + tx.lo), __ldg(m + (long)ty.lo * mw + tx.hi), tx.frac);
                const float bot = lerp_rn(__ldg(m + (long)ty.hi * mw + tx.lo), __ldg(m + (long)ty.hi * mw + tx.hi), tx.frac);
                v = lerp_rn(top, bot, ty.frac);
            }
        }
        out[e] = rintf(v);                                                                     // torch.round, layers.py:323
    }
}

// thread per RoI: probs[n, ncls], deltas[n, ncls, 4], rois[n, 4] normalised, windows[bs, 4] pixels
__global__ void __launch_bounds__(128) detection_decode_kernel(const float *__restrict__ rois, const float *__restrict__ probs,
                                                              const float *__restrict__ deltas, const float *__restrict__ windows, int n,
                                                              int per_image, int ncls, float s0, float s1, float s2, float s3, float img_h, float img_w,
                                                              float min_conf, float *__restrict__ boxes, float *__restrict__ scores,
                                                              int *__restrict__ class_ids, int *__restrict__ keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = probs + (long)i * ncls;
    float best = p[0];
    int cid = 0;
    for (int c = 1; c < ncls; ++c) {                        // torch.max(probs, dim=1): first maximum
        const float v = p[c];
        if (v > best) { best = v; cid = c; }
    }
    const float4 d = *reinterpret_cast<const float4 *>(deltas + ((long)i * ncls + cid) * 4);
    const float4 r = *reinterpret_cast<const float4 *>(rois + (long)i * 4);
    const float dy = __fmul_rn(d.x, s0), dx = __fmul_rn(d.y, s1), dh = __fmul_rn(d.z, s2), dw = __fmul_rn(d.w, s3);      // layers.py:749
    float height = __fsub_rn(r.z, r.x), width = __fsub_rn(r.w, r.y);                                                    // box_utils.py:14-15
    float cy = __fadd_rn(r.x, __fmul_rn(0.5f, height)), cx = __fadd_rn(r.y, __fmul_rn(0.5f, width));
    cy = __fadd_rn(cy, __fmul_rn(dy, height));
    cx = __fadd_rn(cx, __fmul_rn(dx, width));
    height = __fmul_rn(height, expf(dh));
    width = __fmul_rn(width, expf(dw));
    float y1 = __fsub_rn(cy, __fmul_rn(0.5f, height)), x1 = __fsub_rn(cx, __fmul_rn(0.5f, width));
    float y2 = __fadd_rn(y1, height), x2 = __fadd_rn(x1, width);
    y1 = __fmul_rn(y1, img_h); x1 = __fmul_rn(x1, img_w); y2 = __fmul_rn(y2, img_h); x2 = __fmul_rn(x2, img_w);        // layers.py:758
    const float4 w = *reinterpret_cast<const float4 *>(windows + (long)(i / per_image) * 4);                            // box_utils.py:46-59
    y1 = fminf(fmaxf(y1, w.x), w.z); x1 = fminf(fmaxf(x1, w.y), w.w);
    y2 = fminf(fmaxf(y2, w.x), w.z); x2 = fminf(fmaxf(x2, w.y), w.w);
    y1 = rintf(y1); x1 = rintf(x1); y2 = rintf(y2); x2 = rintf(x2);                                                    // layers.py:762
    const float area = __fmul_rn(__fsub_rn(y1, y2), __fsub_rn(x1, x2));                                                 // :765
    *reinterpret_cast<float4 *>(boxes + (long)i * 4) = make_float4(y1, x1, y2, x2);
    scores[i] = best;
    class_ids[i] = cid;
    keep[i] = (cid > 0 && best >= min_conf && area > 0.f) ? 1 : 0;                                                      // :766
}

}  // namespace fi

using namespace fi;

FI_API int fi_mask_targets(const float *gt_masks, const float *pos_rois, const float *gt_boxes, const int *assignment, int num_rois, int num_gt,
                           int mask_height, int mask_width, int target_height, int target_width, int use_mini_mask, float *targets,
                           cudaStream_t stream) {
    FI_REQUIRE(num_rois >= 0 && num_gt >= 0 && mask_height > 0 && mask_width > 0 && target_height > 0 && target_width > 0, "fi_mask_targets: bad sizes");
    if (num_rois == 0) return ok();
    FI_REQUIRE(gt_masks && pos_rois && assignment && targets && (!use_mini_mask || gt_boxes), "fi_mask_targets: null pointer");
    FI_REQUIRE(((uintptr_t)pos_rois % 16) == 0 && ((uintptr_t)gt_boxes % 16) == 0, "fi_mask_targets: box arrays must be 16-byte aligned");
    const long total = (long)num_rois * target_height * target_width;
    long grid = (total + 255) / 256;
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    mask_target_kernel<<<(int)grid, 256, 0, stream>>>(gt_masks, pos_rois, gt_boxes, assignment, num_rois, num_gt, mask_height, mask_width, target_height,
                                                     target_width, use_mini_mask ? 1 : 0, targets);
    return check_launch("fi_mask_targets");
}

FI_API int fi_detection_decode(const float *rois, const float *probs, const float *deltas, const float *windows, int batch, int rois_per_image,
                               int num_classes, const float *std_dev4, float image_height, float image_width, float min_confidence, float *boxes,
                               float *scores, int *class_ids, int *keep, cudaStream_t stream) {
    FI_REQUIRE(batch >= 0 && rois_per_image >= 0 && num_classes >= 1 && std_dev4, "fi_detection_decode: bad arguments");
    const int n = batch * rois_per_image;
    if (n == 0) return ok();
    FI_REQUIRE(rois && probs && deltas && windows && boxes && scores && class_ids && keep, "fi_detection_decode: null pointer");
    FI_REQUIRE(((uintptr_t)rois % 16) == 0 && ((uintptr_t)deltas % 16) == 0 && ((uintptr_t)windows % 16) == 0 && ((uintptr_t)boxes % 16) == 0,
               "fi_detection_decode: 16-byte aligned tensors");
    detection_decode_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(rois, probs, deltas, windows, n, rois_per_image, num_classes, std_dev4[0], std_dev4[1],
                                                                 std_dev4[2], std_dev4[3], image_height, image_width, min_confidence, boxes, scores,
                                                                 class_ids, keep);
    return check_launch("fi_detection_decode");
}
