// NCHW RoIAlign backward (the reference's layout: CropAndResizeBackpropImageLaucher, crop_and_resize_kernel.cu:84-165) through the
// NHWC tile-owner kernels.
//
// In NCHW the 256 channel values of a tap are H*W*4 bytes apart: the scatter formulation is one scalar atomic per (channel, tap)
// (crop_bwd_nchw_kernel, roi_align.cu) and was measured SLOWER than the reference's own kernel recompiled for sm_100a on the large
// maps (C2 level 2, 14x14: 2.42 ms vs 1.85 ms, profiles/r01_microbench_c2_v4.json).  Here instead
//   1. the crop gradients [R,C,P*P] are transposed to [R,P*P,C] through shared memory (src_row applied on the way),
//   2. the tile-owner backward (roi_align_bwd_tile.cu / roi_align_bwd_pix.cu: no atomics, every pixel written once) produces the
//      dense map in NHWC,
//   3. the map is transposed back to [B,C,H*W], ADDING onto the caller's contents (the reference launcher accumulates).
// Two extra streaming passes over the gradients and the map (fully coalesced both ways) instead of 4 scalar atomics per element.
#include "fi_common.cuh"

namespace fi {

// in[n][a][b] -> out[n][b][a]; grid (ceil(Bd / 32), ceil(A / 32), N), 32 x 8 threads.  row_of: optional source batch index per n.
template <bool ACC>
__global__ void __launch_bounds__(256) transpose_batched_kernel(const float *__restrict__ in, float *__restrict__ out, int A, int Bd,
                                                               const int *__restrict__ row_of) {
    __shared__ float tile[32][33];
    const long n_in = row_of ? (long)row_of[blockIdx.z] : (long)blockIdx.z;
    const float *src = in + n_in * (long)A * Bd;
    float *dst = out + (long)blockIdx.z * A * Bd;
    const int b0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int a = a0 + ty + k, b = b0 + tx;
        if (a < A && b < Bd) tile[ty + k][tx] = __ldcs(src + (long)a * Bd + b);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int b = b0 + ty + k, a = a0 + tx;
        if (a < A && b < Bd) {
            float *p = dst + (long)b * A + a;
            const float v = tile[tx][ty + k];
            if (ACC) *p = __fadd_rn(*p, v);
            else __stcs(p, v);
        }
    }
}

}  // namespace fi

using namespace fi;

namespace fi { char *tile_workspace(size_t bytes, cudaStream_t stream); }     // roi_align_bwd_tile.cu: grow-only block per (device, stream)

// Returns FI_ERR_UNSUPPORTED (nothing touched) when the shape does not qualify.
int fi_nchw_backward_via_nhwc(const float *grads, const float *boxes, const int *box_ind, const int *src_row, int R, int B, int H, int W, int ph,
                              int pw, int C, float *gimg, int accumulate, cudaStream_t stream) {
    if (R <= 0 || C % 128 != 0 || ph > 16 || pw > 16 || option(FI_OPT_BWD_FORM) == 3) return FI_ERR_UNSUPPORTED;
    if (((uintptr_t)boxes % 16) != 0 || (long)R * ph * pw >= (1L << 29) || R >= 65536 || B >= 65536) return FI_ERR_UNSUPPORTED;
    const size_t g_bytes = ((size_t)R * C * ph * pw * sizeof(float) + 255) / 256 * 256;
    const size_t m_bytes = ((size_t)B * C * H * W * sizeof(float) + 255) / 256 * 256;
    fi_bwd_set set;
    set.grads_image = reinterpret_cast<float *>(16); set.grads = reinterpret_cast<const float *>(16); set.grads2 = nullptr;
    set.boxes = boxes; set.box_ind = box_ind; set.src_row = nullptr;
    set.batch = B; set.image_height = H; set.image_width = W; set.depth = C; set.num_boxes = R; set.crop_height = ph; set.crop_width = pw;
    set.num_boxes_dev = nullptr;
    const int exact = fi_get_deterministic();
    const size_t t_bytes = fi_crop_sets_backward_workspace(&set, 1, exact, 0);
    if (t_bytes == 0) return FI_ERR_UNSUPPORTED;
    char *ws = tile_workspace(g_bytes + m_bytes + t_bytes, stream);
    if (!ws) return fi_last_status();
    float *g_nhwc = reinterpret_cast<float *>(ws), *m_nhwc = reinterpret_cast<float *>(ws + g_bytes);
    // 1. [R, C, P*P] -> [R, P*P, C]
    const int pp = ph * pw;
    transpose_batched_kernel<false><<<dim3(ceil_div(pp, 32), ceil_div(C, 32), R), 256, 0, stream>>>(grads, g_nhwc, C, pp, src_row);
    if (int e = check_launch("fi_crop_and_resize_backward[nchw: gradients to nhwc]")) return e;
    // 2. tile-owner backward in NHWC
    set.grads_image = m_nhwc; set.grads = g_nhwc;
    fi_bwd_plan plan;
    if (int e = fi_crop_sets_backward_plan(&set, 1, exact, 0, ws + g_bytes + m_bytes, t_bytes, &plan, stream)) return e;
    if (int e = fi_crop_sets_backward_run(&plan, &set, 1, 1, stream)) return e;
    // 3. [B, H*W, C] -> [B, C, H*W], onto the caller's map
    const dim3 grid(ceil_div(C, 32), ceil_div(H * W, 32), B);
    if (accumulate) transpose_batched_kernel<true><<<grid, 256, 0, stream>>>(m_nhwc, gimg, H * W, C, nullptr);
    else transpose_batched_kernel<false><<<grid, 256, 0, stream>>>(m_nhwc, gimg, H * W, C, nullptr);
    return check_launch("fi_crop_and_resize_backward[nchw: map from nhwc]");
}
