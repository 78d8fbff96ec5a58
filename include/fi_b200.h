/* fi_b200.h -- C ABI of libfi_b200.so, the B200 (sm_100a) implementation of the Feature Intertwiner
 * hot path.  Plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller;
 * every entry point enqueues on the given stream and returns without synchronising.
 *
 * The reference's drop-in boundary is the `extern "C"` launcher layer underneath its (long dead)
 * torch.utils.ffi modules -- SURVEY.md section 8(b).  Section 1 below re-exports those launchers with
 * their exact names and signatures; section 2 onwards are the B200-native entry points the Python
 * operator layer (feature_intertwiner_b200/*.py) binds.  Citations are relative to the reference tree.
 *
 * Error convention: the reference's launchers print and exit(-1) on a launch failure
 * (crop_and_resize_kernel.cu:186-191).  Here nothing ever exits: the fi_* entry points return FI_OK or
 * a negative status and fi_last_error() describes it; the void reference-named launchers record the
 * status where fi_last_status() can read it.
 */
#ifndef FI_B200_H
#define FI_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if !defined(__DRIVER_TYPES_H__) && !defined(__CUDA_RUNTIME_H__)
typedef struct CUstream_st *cudaStream_t;
#endif

#define FI_OK 0
#define FI_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, unsupported shape) */
#define FI_ERR_CUDA (-2)        /* a CUDA call or kernel launch failed                             */
#define FI_ERR_UNSUPPORTED (-3) /* valid request this build has no kernel for                      */

#define FI_LAYOUT_NCHW 0 /* [N,C,H,W] contiguous -- the reference's layout                           */
#define FI_LAYOUT_NHWC 1 /* torch.channels_last: same logical tensor, C innermost in memory        */

int fi_abi_version(void);
const char *fi_last_error(void); /* thread-local, never NULL */
int fi_last_status(void);        /* status of the last call made by this thread */
unsigned long long fi_kernel_launches(void); /* kernels this library has launched in this process */

/* Process-wide kernel-selection switches (experiments, A/B measurements, tests).  Each option is initialised ONCE, on
 * first use, from its environment variable and can be changed at run time; no entry point calls getenv() per launch.
 * fi_set_option returns the previous value (FI_ERR_INVALID for an unknown option / value). */
#define FI_OPT_BWD_FORM 0         /* NHWC RoIAlign backward: 0 bulk-copy staged register accumulation (default), 1 shared-memory
                                     accumulate kernel [FI_BWD_ACC=smem], 2 fused single tile kernel [FI_BWD_TILE=fused],
                                     3 vector reductions [FI_BWD=red] */
#define FI_OPT_TILE_SHAPE 1       /* tile of forms 1 and 2: 0 = 4x8 (default), 1 = 4x4 [FI_TILE=4x4], 2 = 2x8 [FI_TILE=2x8] */
#define FI_OPT_FWD_FORM 2         /* NHWC RoIAlign forward: 0 = default (the lean unit, both 128-channel slabs per warp, 2 x 8 warps / SM);
                                     1 = round-1 unit (one slab per warp, scalar lerps); 2..6 = lean shapes kept for A/B runs
                                     (csrc/roi_align.cu::kLeanShapes).  All forms give identical bits. [FI_FWD_FORM] */
#define FI_OPT_SINKHORN_GENERIC 3 /* D = 1 problems: 0 = solved by classes, dense kernels only for what that leaves (default);
                                     1 = dense kernels only (K in registers for N = 256); 2 = generic shared-memory kernel only
                                     [FI_SINKHORN_GENERIC=1|2] */
#define FI_OPT_PIX_CFG 4          /* form 0, ring shape (slots per batch x batches, CTAs per SM): 0 = 32x2,3 (default); 1 = 32x3,2;
                                     2 = 16x6,2; 3 = 16x4,3 [FI_PIX_CFG] */
#define FI_OPT_PIX_GROUP 5        /* form 0, tiles per work ticket minus 1: 0..7 [FI_PIX_GROUP] */
#define FI_OPT_FWD_CHUNK 6        /* level-batched lean forward: units per block chunk; 0 = default (warps per block: consecutive
                                     units go to consecutive warps of the grid), 1..6 = 8, 16, 32, 64, 128, 256 (slower, kept for A/B) [FI_FWD_CHUNK] */
#define FI_OPT_FWD_PAIR 7         /* level-batched lean forward: 0 / 2 = walk two sets over the same map and boxes box by box (default), 1 = off [FI_FWD_PAIR] */
#define FI_OPT_FWD_SCHED 8        /* level-batched lean forward: 0 / 2 = tickets (every warp draws its next units from a counter in global
                                     memory: faster SMs take more; FI_OPT_FWD_CHUNK then selects the units per draw, 0 = default (1),
                                     1..6 = 1, 2, 3, 4, 7, 14), 1 = static chunks [FI_FWD_SCHED] */
#define FI_OPT_COUNT 9
int fi_set_option(int option, int value);
int fi_get_option(int option);

/* ---------------------------------------------------------------------------------------------
 * 1. Reference-named launchers (exact signatures).
 * ------------------------------------------------------------------------------------------- */

/* lib/roi_align/src/cuda/crop_and_resize_kernel.h:8-12.  image[batch,depth,H,W] NCHW, boxes[num_boxes,4]
 * normalised (y1,x1,y2,x2), box_ind[num_boxes] -> crops[num_boxes,depth,crop_h,crop_w] NCHW.
 * Unlike the reference the caller need NOT zero `crops` first (crop_and_resize_gpu.c:25): rows whose
 * box_ind is out of range are written as zeros by the kernel itself. */
void CropAndResizeLaucher(const float *image_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes,
                          int batch, int image_height, int image_width, int crop_height, int crop_width, int depth,
                          float extrapolation_value, float *crops_ptr, cudaStream_t stream);

/* lib/roi_align/src/cuda/crop_and_resize_kernel.h:14-18.  ACCUMULATES into grads_image exactly like the
 * reference kernel; the caller zeroes it first (crop_and_resize_gpu.c:57). */
void CropAndResizeBackpropImageLaucher(const float *grads_ptr, const float *boxes_ptr, const int *box_ind_ptr,
                                       int num_boxes, int batch, int image_height, int image_width, int crop_height,
                                       int crop_width, int depth, float *grads_image_ptr, cudaStream_t stream);

/* lib/roi_pooling/src/roi_pooling_kernel.h:8-18.  bottom[B,C,H,W] NCHW, rois[num_rois,5]=(b,x1,y1,x2,y2) px.
 * Return 1 on success (the reference's convention), 0 on failure. */
int ROIPoolForwardLaucher(const float *bottom_data, const float spatial_scale, const int num_rois, const int height,
                          const int width, const int channels, const int pooled_height, const int pooled_width,
                          const float *bottom_rois, float *top_data, int *argmax_data, cudaStream_t stream);
int ROIPoolBackwardLaucher(const float *top_diff, const float spatial_scale, const int batch_size, const int num_rois,
                           const int height, const int width, const int channels, const int pooled_height,
                           const int pooled_width, const float *bottom_rois, float *bottom_diff,
                           const int *argmax_data, cudaStream_t stream);

/* lib/nms/src/cuda/nms_kernel.h:11-12.  boxes_dev[boxes_num,5]=(x1,y1,x2,y2,score) sorted by score;
 * mask_dev[boxes_num, ceil(boxes_num/64)] bit j of word (i,w) set <=> IoU(i, 64w+j) > thresh, j > i.
 * Runs on the legacy default stream like the reference (nms_kernel.cu:79). */
void _nms(int boxes_num, float *boxes_dev, unsigned long long *mask_dev, float nms_overlap_thresh);

/* ---------------------------------------------------------------------------------------------
 * 2. RoIAlign (TF crop_and_resize semantics), B200-native.
 * ------------------------------------------------------------------------------------------- */

/* Forward.  Replaces crop_and_resize_gpu_forward (lib/roi_align/src/crop_and_resize_gpu.c:7-37).
 *   image         [batch,depth,H,W] in `image_layout`
 *   boxes         [num_boxes,4] (y1,x1,y2,x2) normalised; box_ind [num_boxes] int32
 *   dst_row       NULL, or [num_boxes] int32: crop r is written to row dst_row[r] of `crops` (fuses the
 *                 scatter-back of lib/sub_module.py:645-662 into the op); rows are never zero-filled here
 *   crops         [>=num_boxes,depth,crop_h,crop_w] in `crops_layout`
 * image_layout and crops_layout must be equal (NCHW/NCHW or NHWC/NHWC); NHWC needs depth % 4 == 0. */
int fi_crop_and_resize_forward(const float *image, int image_layout, const float *boxes, const int *box_ind,
                               const int *dst_row, int num_boxes, int batch, int image_height, int image_width,
                               int crop_height, int crop_width, int depth, float extrapolation_value, float *crops,
                               int crops_layout, cudaStream_t stream);

/* Forward with two destinations (NHWC, depth % 128 == 0): crop r goes to row dst_row[r] of `crops` AND to row r of
 * `crops_compact`.  Dev.forward needs both for the 14x14 crops of levels 2-4: the scattered copy feeds the mask head in
 * (image, roi) order (lib/sub_module.py:645-662), the compact one feeds the critic (:582). */
int fi_crop_and_resize_forward_dual(const float *image, const float *boxes, const int *box_ind, const int *dst_row,
                                    int num_boxes, int batch, int image_height, int image_width, int crop_height,
                                    int crop_width, int depth, float extrapolation_value, float *crops,
                                    float *crops_compact, cudaStream_t stream);

/* Backward.  Replaces crop_and_resize_gpu_backward (crop_and_resize_gpu.c:40-69).
 *   grads        [>=num_boxes,depth,crop_h,crop_w] in `grads_layout`; src_row as dst_row above
 *   grads_image  [batch,depth,H,W] in `image_layout`; zero-filled first unless accumulate != 0 */
int fi_crop_and_resize_backward(const float *grads, int grads_layout, const float *boxes, const int *box_ind,
                                const int *src_row, int num_boxes, int batch, int image_height, int image_width,
                                int crop_height, int crop_width, int depth, float *grads_image, int image_layout,
                                int accumulate, cudaStream_t stream);

/* Several crop sets that were taken from the SAME feature map (e.g. the 7x7 and the 14x14 crops of one level,
 * lib/sub_module.py:555-572), back-propagated in ONE pass that writes grads_image exactly once.  NHWC only.
 *   grads    [rows,crop_h,crop_w,depth]: gradient of crop r is row src_row[r] (row r when src_row is NULL)
 *   grads2   NULL, or a second gradient for the same crops in compact row order (row r), added to the first
 * Both modes use the tile-owner kernels (crop_h, crop_w <= 16, depth % 128 == 0; csrc/roi_align_bwd_tile.cu): every
 * 4x8-pixel tile is accumulated in shared memory by one warp per 128-channel slab, in the order of the reference's serial
 * CPU loop, and the map is written exactly once (no zero fill, no atomics) -- run-to-run identical either way.
 * deterministic != 0: un-fused fp32 arithmetic as well, bit-identical to crop_and_resize.c:157-252 for a single set.
 * deterministic == 0: packed FMAs with pre-multiplied weights; zero-padded / sub-pixel boxes are pre-reduced in parallel.
 * Other shapes (or FI_BWD=red): one zero fill, then 128-bit vector reductions set by set. */
typedef struct fi_crop_set {
    const float *grads;
    const float *grads2;
    const float *boxes;   /* [num_boxes,4] */
    const int *box_ind;   /* [num_boxes]   */
    const int *src_row;   /* [num_boxes] or NULL */
    int num_boxes, crop_height, crop_width;
} fi_crop_set;
int fi_crop_and_resize_backward_multi(const fi_crop_set *sets, int num_sets, int batch, int image_height, int image_width,
                                      int depth, float *grads_image, int accumulate, int deterministic, cudaStream_t stream);

/* Process-wide arithmetic of the NHWC backward: 0 = default (packed FMAs), 1 = exact (bit-identical to the reference's
 * CPU loop, see above).  Returns the old value. */
int fi_set_deterministic(int on);
int fi_get_deterministic(void);

/* Level-batched RoIAlign: every crop set of one Dev.forward pass (lib/sub_module.py:429-600: up to 3 "big" sets on the raw
 * maps and 4 x 2 "small" sets on the made-up maps, each with its own map, boxes and crop size) in ONE launch, forward and
 * backward.  NHWC, depth % 128 == 0.  At most 12 sets. */
typedef struct fi_fwd_set {
    const float *image;      /* [batch,H,W,depth] */
    const float *boxes;      /* [num_boxes,4] */
    const int *box_ind;      /* [num_boxes] */
    const int *dst_row;      /* [num_boxes] or NULL (row r) */
    float *crops;            /* rows dst_row[r] */
    float *crops_compact;    /* NULL, or a second copy at row r */
    int batch, image_height, image_width, depth, num_boxes, crop_height, crop_width;
    float extrapolation_value;
    const int *num_boxes_dev; /* NULL, or a device int: the actual number of boxes (<= num_boxes = capacity); rows past it are not written */
} fi_fwd_set;
int fi_crop_sets_forward(const fi_fwd_set *sets, int num_sets, cudaStream_t stream);

typedef struct fi_bwd_set {
    float *grads_image;      /* [batch,H,W,depth]; several sets may name the same map */
    const float *grads;      /* rows src_row[r] (row r when NULL) */
    const float *grads2;     /* NULL, or a second gradient at row r, added to the first */
    const float *boxes;
    const int *box_ind;
    const int *src_row;
    int batch, image_height, image_width, depth, num_boxes, crop_height, crop_width;
    const int *num_boxes_dev; /* NULL, or a device int: the actual number of boxes (<= num_boxes, which then is the capacity
                                 of boxes / box_ind / src_row).  Lets a caller keep list lengths on the device (no host sync,
                                 fixed shapes for CUDA graphs).  Tile-owner forms only. */
} fi_bwd_set;
/* zero_first != 0: the maps are overwritten (the tile-owner kernels write every pixel once; the reduction fallback
 * zero-fills each distinct grads_image first); zero_first == 0: the sums are added onto the existing contents.
 * Scratch memory: a grow-only block per (device, stream) inside the library; it cannot grow while `stream` is being
 * captured -- use the plan / run pair below there. */
int fi_crop_sets_backward(const fi_bwd_set *sets, int num_sets, int zero_first, cudaStream_t stream);

/* The same backward in two steps on caller-owned scratch memory:
 *   plan  tile_prep + bin_enumerate: per-tile sample lists.  Needs boxes / box_ind / src_row / sizes and WHETHER grads2 will
 *         be given (any non-NULL value), not the gradients: it can run at forward time, on another stream.
 *   run   tile_collapse + accumulate: reads the gradients named by `sets` (grads, grads2, grads_image may differ from plan
 *         time; everything else must match the plan, FI_ERR_INVALID otherwise) and writes every map pixel once.
 * `exact` as fi_set_deterministic.  max_entries > 0: caller's bound on the number of list entries (4 per sample of a
 * one-source set, 8 of a two-source set, + 32 per box) when it knows better than the per-set capacities, e.g. because the
 * sets partition the boxes; fi_crop_sets_backward_overflow tells (with a stream synchronisation) whether it was too small.
 * fi_crop_sets_backward_workspace returns 0 when the sets need the reduction kernels (plan then returns FI_ERR_UNSUPPORTED). */
typedef struct fi_bwd_plan { unsigned long long opaque[640]; } fi_bwd_plan;
size_t fi_crop_sets_backward_workspace(const fi_bwd_set *sets, int num_sets, int exact, long max_entries);
int fi_crop_sets_backward_plan(const fi_bwd_set *sets, int num_sets, int exact, long max_entries, void *workspace,
                               size_t workspace_bytes, fi_bwd_plan *plan, cudaStream_t stream);
int fi_crop_sets_backward_run(const fi_bwd_plan *plan, const fi_bwd_set *sets, int num_sets, int zero_first, cudaStream_t stream);
int fi_crop_sets_backward_overflow(const fi_bwd_plan *plan, cudaStream_t stream);

/* Integer taps of every sample, taps[num_boxes,crop_h,crop_w,5] = (y_lo,y_hi,x_lo,x_hi,inside): the "RoI
 * indices" the parity bar requires bit-exact (crop_and_resize_kernel.cu:40-70). */
int fi_crop_taps(const float *boxes, int num_boxes, int image_height, int image_width, int crop_height,
                 int crop_width, int *taps, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 3. RoI -> pyramid level, reliable / less-reliable split (lib/sub_module.py:397-448,474-493,541-548).
 * ------------------------------------------------------------------------------------------- */

/* level[n] = clamp(round(4 + log2(sqrt(w*h) / (base / sqrt(image_area)))), 2, 5); zero-padded RoIs -> 2. */
int fi_roi_level(const float *rois, int n, float image_area, float base, int *level, cudaStream_t stream);

/* One launch builds every per-level index list of Dev.forward in torch.nonzero (row-major) order:
 *   small_idx [4,n]  flat RoI ids with level == 2+l          small_cnt [4]
 *   big_idx   [4,n]  flat RoI ids with level  > 2+l          big_cnt   [4]
 *   slot      [n]    position of RoI i inside its own level's small list (its row in that level's crop)
 * n <= 65536. */
int fi_split_levels(const int *level, int n, int *small_idx, int *small_cnt, int *big_idx, int *big_cnt, int *slot,
                    cudaStream_t stream);

/* Same launch, additionally emitting what Dev.forward gathers per level (lib/sub_module.py:489-493,541-548):
 * small_boxes/big_boxes [4,n,4] = rois[idx], small_ind/big_ind [4,n] = idx / rois_per_image (the image a RoI belongs to),
 * small_gt/big_gt [4,n] = gt[idx] (only when gt != NULL).  `order` (NULL or a permutation of 0..n-1) is the order in which RoIs
 * are visited: with a spatially sorted order the lists -- and so the RoIAlign work -- walk each image coherently, which is what
 * keeps overlapping RoIs' taps in L2; NULL gives torch.nonzero order.  img_cnt (NULL or [8, ceil(n / rois_per_image)]): how many
 * members of list k (0-3 small, 4-7 big) belong to image b -- with an image-major visiting order these are the per-image extents
 * of every list, which fi_crop_sets_backward_by_image needs. */
int fi_split_levels_gather(const int *level, const float *rois, const int *gt, const int *order, int n, int rois_per_image, int *small_idx,
                           int *small_cnt, int *big_idx, int *big_cnt, int *slot, float *small_boxes, int *small_ind,
                           int *small_gt, float *big_boxes, int *big_ind, int *big_gt, int *img_cnt, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 4. Per-class statistics (lib/sub_module.py:664-684 _assign_feat2cls).
 * ------------------------------------------------------------------------------------------- */

/* gt[k] int32 class ids, feat[k,F] -> mean[F,ncls] (class-minor like the reference), cnt[ncls];
 * class 0 (background) and ids outside [0,ncls) contribute nothing.  ncls <= 1024.  Deterministic. */
int fi_segment_mean_forward(const int *gt, const float *feat, int k, int F, int ncls, float *mean, float *cnt,
                            cudaStream_t stream);
/* grad_feat[i,:] = grad_mean[:,gt_i] / cnt[gt_i]  (0 for background rows) */
int fi_segment_mean_backward(const int *gt, const float *grad_mean, const float *cnt, int k, int F, int ncls,
                             float *grad_feat, cudaStream_t stream);
/* Several lists in one launch (Dev.forward takes 6 per pass).  Forward reads gt / feat / k / k_dev and writes mean / cnt;
 * backward reads gt / k / k_dev / cnt / grad_mean and writes grad_feat.  At most 8 lists. */
typedef struct fi_seg_set {
    const int *gt;
    const float *feat;
    int k;
    const int *k_dev;
    float *mean;
    float *cnt;
    const float *grad_mean;
    float *grad_feat;
} fi_seg_set;
int fi_segment_mean_forward_batch(const fi_seg_set *sets, int num_sets, int F, int ncls, cudaStream_t stream);
int fi_segment_mean_backward_batch(const fi_seg_set *sets, int num_sets, int F, int ncls, cudaStream_t stream);

/* Visiting order of rois[batch, R, 4] that walks every image coarse tile by coarse tile (grid x grid by box centre, boustrophedon,
 * stable): order[batch * R] int32, a permutation of the flat RoI ids, image-major.  R <= 4096. */
int fi_spatial_order(const float *rois, int batch, int rois_per_image, int grid, int *order, cudaStream_t stream);

/* The same with the list length on the device: k is the capacity of gt / feat, *k_dev (NULL: k) the rows in use;
 * backward writes zeros into the rows past it. */
int fi_segment_mean_forward_n(const int *gt, const float *feat, int k, const int *k_dev, int F, int ncls, float *mean, float *cnt,
                              cudaStream_t stream);
int fi_segment_mean_backward_n(const int *gt, const float *grad_mean, const float *cnt, int k, const int *k_dev, int F, int ncls,
                               float *grad_feat, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 5. Sinkhorn optimal-transport loss (lib/OT_module.py:104-135), batched.
 * ------------------------------------------------------------------------------------------- */

/* n_problems independent problems; x,y [n_problems,N,D] (rows = critic channels, cols = positions).
 * loss[p] = <P,C>, C = 1 - xh yh^T, K = exp(-inv_eps C), L iterations, P = diag(a) K diag(b).
 * grad_x / grad_y (NULL or [n_problems,N,D]): d loss / d x, y with P held constant (no_bp_P_L=True,
 * OT_module.py:130-131) through the row normalisation.  N <= 256; any D >= 1. */
int fi_sinkhorn(const float *x, const float *y, int n_problems, int N, int D, float inv_eps, int L, float *loss,
                float *grad_x, float *grad_y, cudaStream_t stream);
/* The same with caller-owned scratch memory for large D (FPN-level loss, N = 64, D up to 4096): the two D-long parts -- the cost
 * matrix x^ y^T and the N x D gradient products -- are spread over (problem, slice of D) CTAs instead of one CTA per problem.
 * fi_sinkhorn_workspace: bytes needed, 0 when the shape keeps everything inside the solver kernel (then workspace may be NULL;
 * without a workspace a large-D call still works, one CTA per problem).  `want_grad` does not change the size. */
size_t fi_sinkhorn_workspace(int n_problems, int N, int D, int want_grad);
int fi_sinkhorn_ws(const float *x, const float *y, int n_problems, int N, int D, float inv_eps, int L, float *loss, float *grad_x,
                   float *grad_y, void *workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 6. Buffer update + class match (lib/model.py:143-224, B == 1 running mean and B > 1 FIFO).
 * ------------------------------------------------------------------------------------------- */

/* Inputs are the UN-normalised sums an all-reduce produces: big_sum[F,ncls] = sum_{gpu,scale} feat*cnt,
 * big_n[ncls] = sum cnt.  Updates buffer[B,F,ncls] / buffer_cnt[B,ncls] in place and writes
 * final_big[F,ncls] (the count-weighted mean over the B slots).  B == 1: running mean over all history
 * (model.py:153-158), slot must be 0.  B > 1: the buffer is a RING -- `slot` is the position this call
 * overwrites (the reference instead shifts the whole buffer left each iteration, model.py:161-164). */
int fi_buffer_update(const float *big_sum, const float *big_n, int B, int slot, int F, int ncls, float *buffer,
                     float *buffer_cnt, float *final_big, cudaStream_t stream);

/* The pointwise steps of the class-level loss head between its library GEMMs, one launch each (csrc/loss_head.cu); fp32, every
 * tensor contiguous.  n = ncls - 1 foreground classes.
 *   fi_merge_stats           _merge_feat_vec (model.py:217-224): feat[GS,F,ncls], cnt[GS,ncls] of the reliable and the less-reliable
 *                            set -> packed = [big_sum F*ncls | big_n ncls | small_sum F*ncls | small_n ncls], the all-reduce operand
 *   fi_merge_stats_backward  d small_feat[GS,F,ncls] = d small_sum[F,ncls] * scale * small_cnt
 *   fi_ot_head_prep          X[n,F] = (small_sum / (small_n + EPS))^T, Y[n,F] = final_big^T without the background class,
 *                            mask[n] = class seen in this batch and present in the buffer (model.py:176-190)
 *   fi_ot_head_combine       loss[n] = (2 w[0:n] - w[n:2n] - w[2n:3n]) * mask          (OT_module.py:78-80)
 *   fi_ot_head_dcritic       the gradient of that combination through the Sinkhorn problems (cx,cy), (cx,cx), (cy,cy) -- gx, gy [3n,N]
 *                            as fi_sinkhorn returns them, g[n] upstream -- and through the ReLU of Cc = [cx; cy] -> dC[2n,N]
 *   fi_relu_mask             d[i] = 0 where h[i] <= 0 (ReLU backward, in place)
 *   fi_col_sum               dst[cols] = column sums of src[rows,cols] (bias gradients)
 *   fi_ot_head_dsum          d small_sum[F,ncls] from dX[n,F] (the division and the transpose backwards; background column 0)
 *   fi_centre_tap_embed      full[count,3] = (0, w1[count], 0): the gradient of Conv1d weight[:, :, 1] as the whole weight's */
int fi_merge_stats(const float *big_feat, const float *big_cnt, const float *small_feat, const float *small_cnt, int GS, int F, int ncls,
                   float *packed, cudaStream_t stream);
int fi_merge_stats_backward(const float *d_small_sum, const float *small_cnt, int GS, int F, int ncls, float scale, float *d_small_feat,
                            cudaStream_t stream);
int fi_ot_head_prep(const float *final_big, const float *small_sum, const float *small_n, const float *buffer_cnt, int B, int F, int ncls,
                    float *X, float *Y, float *mask, cudaStream_t stream);
int fi_ot_head_combine(const float *w, const float *mask, int n, float *loss, cudaStream_t stream);
int fi_ot_head_dcritic(const float *gx, const float *gy, const float *g, const float *mask, const float *Cc, int n, int N, float *dC,
                       cudaStream_t stream);
int fi_relu_mask(float *d, const float *h, long count, cudaStream_t stream);
int fi_col_sum(const float *src, int rows, int cols, float *dst, cudaStream_t stream);
int fi_ot_head_dsum(const float *dX, const float *small_n, int F, int ncls, float *d_small_sum, cudaStream_t stream);
int fi_centre_tap_embed(const float *w1, long count, float *full, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 7. NMS with the reduce on the device (lib/nms/src/nms_cuda.c:17-67, no host round trip).
 * ------------------------------------------------------------------------------------------- */

/* boxes[n_images, n, 5] = (x1,y1,x2,y2,score), each image sorted by descending score.
 * mask: workspace [n_images, n, ceil(n/64)] u64.  keep[n_images, n] int32 (kept indices first, rest -1),
 * num_keep[n_images].  Suppress when IoU > thresh (the reference's GPU rule, nms_kernel.cu:63). */
int fi_nms_batched(const float *boxes, int n_images, int n, float thresh, unsigned long long *mask, int *keep,
                   int *num_keep, cudaStream_t stream);
/* The same for callers that use only the first max_keep survivors of every image (proposal_count, lib/layers.py:118-121;
 * DET_MAX_INSTANCES, :790-795): the sweep of an image stops once it knows that many.  keep[] holds at least
 * min(max_keep, all survivors) entries, identical to the head of fi_nms_batched's list; num_keep[i] >= max_keep means "cut". */
int fi_nms_batched_topk(const float *boxes, int n_images, int n, float thresh, int max_keep, unsigned long long *mask, int *keep,
                        int *num_keep, cudaStream_t stream);

/* Front half of proposal_layer (lib/layers.py:87-122 + tools/box_utils.py:7-45) in one launch: for proposal k of image b,
 * a = order[b,k] (descending-score order, int64 as torch.sort returns it); box = clip(apply_box_deltas(anchors[a],
 * deltas[b,a] * std_dev), [0,0,window_height,window_width]).  Writes boxes[batch,K,4] = (y1,x1,y2,x2) and
 * dets_xyxys[batch,K,5] = (x1,y1,x2,y2,scores_sorted[b,k]) -- the input layout of fi_nms_batched.  std_dev4: 4 HOST floats. */
int fi_proposal_decode(const float *deltas, const float *anchors, const long long *order, const float *scores_sorted, int batch,
                       int num_anchors, int num_proposals, const float *std_dev4, float window_height, float window_width,
                       float *boxes, float *dets_xyxys, cudaStream_t stream);

/* Back half (lib/layers.py:128-139, lib/nms/nms_wrapper.py:24-33) without a host read: m = min(proposal_count, min_b num_keep[b]);
 * rois[batch, proposal_count, 4] = boxes[b, keep[b,j]] / (H,W,H,W) for j < m, zeros after (the reference zero-pads RoI lists,
 * lib/layers.py:413,427); *num_rois = m (device int, may be NULL).  boxes[batch,K,4], keep[batch,K], num_keep[batch] as
 * fi_proposal_decode / fi_nms_batched leave them. */
int fi_proposal_gather(const float *boxes, const int *keep, const int *num_keep, int batch, int num_proposals, int proposal_count,
                       float window_height, float window_width, float *rois, int *num_rois, cudaStream_t stream);

/* Mask targets of the positive RoIs (lib/layers.py:296-323) in one launch: targets[i] = round(crop_and_resize(gt_masks[assignment[i]],
 * box_i, target_h x target_w)), box_i = pos_rois[i] rewritten into the frame of gt_boxes[assignment[i]] when use_mini_mask
 * (lib/layers.py:304-313), else pos_rois[i].  gt_masks[num_gt, mask_h, mask_w] fp32, pos_rois[num_rois,4], gt_boxes[num_gt,4],
 * assignment[num_rois] int32 -> targets[num_rois, target_h, target_w]. */
int fi_mask_targets(const float *gt_masks, const float *pos_rois, const float *gt_boxes, const int *assignment, int num_rois, int num_gt,
                    int mask_height, int mask_width, int target_height, int target_width, int use_mini_mask, float *targets,
                    cudaStream_t stream);

/* Decode half of detection_layer (lib/layers.py:738-766) in one launch, one RoI per thread: arg-max class, its deltas * std_dev
 * (4 HOST floats), apply_box_deltas, scale to pixels, clip to windows[batch,4], round, keep = class > 0 && score >= min_confidence
 * && area > 0.  rois[batch*R,4] normalised, probs[batch*R,ncls], deltas[batch*R,ncls,4] -> boxes[batch*R,4] (y1,x1,y2,x2) pixels,
 * scores, class_ids, keep. */
int fi_detection_decode(const float *rois, const float *probs, const float *deltas, const float *windows, int batch, int rois_per_image,
                        int num_classes, const float *std_dev4, float image_height, float image_width, float min_confidence, float *boxes,
                        float *scores, int *class_ids, int *keep, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 8. RoIPool with an argmax-scatter backward (lib/roi_pooling/src/roi_pooling_cuda.c:7-88).
 * ------------------------------------------------------------------------------------------- */
int fi_roi_pool_forward(const float *bottom, float spatial_scale, int batch, int num_rois, int height, int width,
                        int channels, int pooled_h, int pooled_w, const float *rois, float *top, int *argmax,
                        cudaStream_t stream);
/* bottom_diff is zero-filled here, then receives exactly what ROIPoolBackward (roi_pooling_kernel.cu:128-203)
 * computes, at O(outputs) instead of O(inputs x rois) work. */
int fi_roi_pool_backward(const float *top_diff, float spatial_scale, int batch, int num_rois, int height, int width,
                         int channels, int pooled_h, int pooled_w, const float *rois, float *bottom_diff,
                         const int *argmax, cudaStream_t stream);

/* The same on torch.channels_last tensors (bottom[B,H,W,C], top / argmax[R,ph,pw,C], channels % 128 == 0): warp per (RoI, bin,
 * 128-channel slab), coalesced 128-bit loads.  argmax holds the reference's flat NCHW offsets (layout-independent): values and
 * indices equal those of ROIPoolForwardLaucher on the same data. */
int fi_roi_pool_forward_nhwc(const float *bottom, float spatial_scale, int batch, int num_rois, int height, int width, int channels,
                             int pooled_h, int pooled_w, const float *rois, float *top, int *argmax, cudaStream_t stream);
int fi_roi_pool_backward_nhwc(const float *top_diff, float spatial_scale, int batch, int num_rois, int height, int width, int channels,
                              int pooled_h, int pooled_w, const float *rois, float *bottom_diff, const int *argmax, cudaStream_t stream);

/* ---- the exchange step over NVLink / NVSwitch peer memory (one process per GPU) -------------------------------------------
 * Replaces the gather of nn.DataParallel + _merge_feat_vec's reduction over the gpu axis (lib/model.py:217-224, 394-402;
 * tools/utils.py:645-654): an all-reduce(SUM) of fp32 values as ONE kernel per rank that reads every peer's copy through
 * peer-mapped memory and adds in rank order (bit-identical totals on every rank).  Capturable in a CUDA graph: the call counter
 * lives on the device.  Set-up: every rank allocates a region, exports its handle, the host side exchanges the handles
 * (any transport: torch.distributed, MPI, a file) and imports the peers' ones; `regions[r]` = rank r's region as mapped in THIS
 * process (regions[rank] = the pointer fi_peer_region_alloc returned).  All ranks must make the same sequence of calls on one
 * set of regions; `n` may vary per call up to the capacity.  A peer that does not arrive within 4 s sets the error word read by
 * fi_peer_error instead of hanging the device. */
#define FI_PEER_MAX_WORLD 16
#define FI_PEER_HANDLE_BYTES 64
size_t fi_peer_region_bytes(size_t capacity_floats);
int fi_peer_region_alloc(size_t capacity_floats, void **region);
int fi_peer_region_free(void *region);
int fi_peer_region_export(void *region, unsigned char handle[FI_PEER_HANDLE_BYTES]);
int fi_peer_region_import(const unsigned char handle[FI_PEER_HANDLE_BYTES], void **region);
int fi_peer_region_release(void *imported);
int fi_peer_error(void *own_region, int *error);
int fi_peer_allreduce_sum(const float *in, float *out, size_t n, int rank, int world, void *const *regions, size_t capacity_floats,
                          cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FI_B200_H */
