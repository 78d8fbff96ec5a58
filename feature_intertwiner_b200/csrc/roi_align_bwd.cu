// RoIAlign backward, multi-set entry for ONE dense gradient map (fi_crop_and_resize_backward_multi) and the process-wide
// "deterministic" switch.  The kernels live in roi_align_bwd_tile.cu (tile-owner form: shared-memory accumulation, every map
// pixel written once; its EXACT mode is the deterministic one -- order and un-fused arithmetic of crop_and_resize.c:190-250)
// and roi_align.cu (vector reductions: the reference's scatter formulation, kept as fallback / FI_BWD=red).
//
// History (profiles/): an earlier deterministic mode kept the accumulators in registers (warp = pixel row of an
// 8x8 tile) and searched, per warp and box, for the samples touching its row: 560-880 warp instructions per sample, 14.5 ms on
// C2.  The tile-owner kernels enumerate samples lane-parallel instead and replaced it.
#include <stdlib.h>

#include "fi_common.cuh"

using namespace fi;

int fi_scatter_backward_nhwc(const float *grads, const float *grads2, const float *boxes, const int *box_ind, const int *src_row, int R, int B,
                             int H, int W, int ph, int pw, int C, float *gimg, cudaStream_t stream);   // roi_align.cu

int fi_tile_backward(const fi_bwd_set *sets, int num_sets, int accumulate, int exact, cudaStream_t stream);   // roi_align_bwd_tile.cu

namespace fi { extern int g_deterministic_init; }   // abi.cu: FI_BWD=exact in the environment at first use
static int g_deterministic = -1;                     // -1: not set yet, take the environment's word
FI_API int fi_get_deterministic(void) {
    if (g_deterministic < 0) { fi::option(FI_OPT_BWD_FORM); g_deterministic = fi::g_deterministic_init; }
    return g_deterministic;
}
FI_API int fi_set_deterministic(int on) { const int old = fi_get_deterministic(); g_deterministic = on ? 1 : 0; return old; }

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
        tmp[i].crop_height = sets[i].crop_height; tmp[i].crop_width = sets[i].crop_width; tmp[i].num_boxes_dev = nullptr;
    }
    if (deterministic) {
        const int rc = fits ? fi_tile_backward(tmp, num_sets, accumulate, 1, stream) : FI_ERR_UNSUPPORTED;
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
