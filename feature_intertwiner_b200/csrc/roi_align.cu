// RoIAlign (TF crop_and_resize, one bilinear sample per bin) for B200 / sm_100a.
//
// Replaces lib/roi_align/src/cuda/crop_and_resize_kernel.cu (thread-per-output NCHW gather, 4 scalar
// uncoalesced loads per element, 4 scalar atomics per element in backward, full memsets of both outputs).
//
// Layouts.  The op is a gather of C-vectors: each sample reads 4 feature pixels and every channel of a
// pixel shares the same coordinates.  In NCHW those C values are H*W*4 bytes apart; in NHWC
// (torch.channels_last -- what cuDNN produces natively on this part) they are ONE contiguous C*4-byte run
// (1 KB for FPN's C=256), so a tap is a fully coalesced 128-bit-per-lane read and a backward tap is a
// coalesced vector reduction.  NHWC is therefore the native layout here; NCHW is kept for drop-in parity
// with callers that hand over contiguous NCHW tensors.
//
// Coordinate math is shared with fi_common.cuh::axis_sample (bit-exact tap indices vs the reference).
#include <stdlib.h>

#include "fi_common.cuh"
#include "roi_align_units.cuh"
#include "roi_align_fwd_lean.cuh"

namespace fi {

// ------------------------------------------------------------------------------------------------
// Tap table (test / verification entry point).
// ------------------------------------------------------------------------------------------------
__global__ void crop_taps_kernel(const float *__restrict__ boxes, int R, int H, int W, int ph, int pw, int *__restrict__ taps) {
    const long total = (long)R * ph * pw;
    for (long s = blockIdx.x * (long)blockDim.x + threadIdx.x; s < total; s += (long)gridDim.x * blockDim.x) {
        const int r = (int)(s / (ph * pw));
        const int rem = (int)(s - (long)r * ph * pw);
        const int i = rem / pw, j = rem - i * pw;
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        const AxisTap tx = axis_sample(x1, x2, axis_step(x1, x2, W, pw), j, W, pw);
        int *t = taps + s * 5;
        t[0] = ty.lo; t[1] = ty.hi; t[2] = tx.lo; t[3] = tx.hi; t[4] = (ty.inside && tx.inside) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// NHWC kernels: one WARP per crop row (box r, row i).
//
// A thread-per-output-vector mapping re-derives the sample coordinates (two IEEE divisions, floor/ceil,
// index div/mod) in every thread and is issue-bound at ~1/4 of HBM speed (measured, profiles/r01).  Here
// the y tap is computed once per row, lane j computes the x tap of sample j, and the row is walked with
// the taps handed around by warp shuffle: per sample a lane only forms 4 addresses, moves VPL float4 per
// tap and does the lerps.  Loads are coalesced 512 B per tap row (C=256: two float4 per lane).
//
// ------------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) crop_fwd_nhwc_kernel(const FwdSet S, long nunits) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long u = warp; u < nunits; u += nwarps) fwd_unit<U>(S, u, lane);
}

// Every crop set of a Dev.forward pass (up to 3 "big" + 4x2 "small") in ONE launch: the small levels (a few hundred
// boxes on a 26x42 map) are far too small to fill 148 SMs on their own.
constexpr int kMaxFwdSets = 12;
struct FwdSets {
    FwdSet s[kMaxFwdSets];
    long first_unit[kMaxFwdSets + 1];
    int n;
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32) crop_fwd_nhwc_sets_kernel(const FwdSets sets) {
    // unit -> set: prefix of the sets' unit counts.  With list lengths on the device the prefix is formed here from the LIVE
    // lengths: walking the capacity ranges instead leaves the live units of every set as a prefix of its own range, and their
    // round-robin assignment to the (persistent) warps then differs by a unit or two per set -- measured +17 % on C2.
    // Tried on this kernel in round 2 and not kept (profiles/r02_fwd_variants.json): taps staged by cp.async.bulk into a per-warp
    // double buffer, lerps from shared memory (bit-identical; 0.96 ms with 7 warps per SM, 1.36 ms with 4: too few warps fit next
    // to their buffers to cover the loaded DRAM latency) and an L2 prefetch of the warp's next unit (1.00 ms) -- against 0.85 ms.
    __shared__ long first[kMaxFwdSets + 1];
    if (threadIdx.x == 0) {
        long acc = 0;
        for (int k = 0; k < sets.n; ++k) {
            const FwdSet &S = sets.s[k];
            const int R = S.R_dev ? max(0, min(*S.R_dev, S.R)) : S.R;
            first[k] = acc;
            acc += (long)R * S.ph * S.slabs;
        }
        first[sets.n] = acc;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const long total = first[sets.n];
    for (long u = warp; u < total; u += nwarps) {
        int k = 0;
        while (u >= first[k + 1]) ++k;
        fwd_unit<4>(sets.s[k], u - first[k], lane);
    }
}

// Ticket counters of the lean kernels (see "TICKETS" below).
constexpr int kTicketSlots = 256;
__device__ unsigned g_fwd_tickets[2 * kTicketSlots];      // per slot: next unit, finished blocks

__device__ __forceinline__ unsigned draw_ticket(unsigned *next, unsigned grab, int lane) {
    unsigned u = 0;
    if (lane == 0) u = atomicAdd(next, grab);
    return __shfl_sync(0xffffffffu, u, 0);
}
__device__ __forceinline__ void retire_tickets(unsigned *next, unsigned *done) {      // every block, after its last draw
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            *next = 0u;
            *done = 0u;
            __threadfence();
        }
    }
}

// Lean formulation (roi_align_fwd_lean.cuh): WARPS warps per block, MINB blocks per SM, persistent over the units; schedule as in
// the level-batched kernel below (tickets: `chunk` = units per draw, `slot` = counter pair; else static chunks).
template <int VPL, int U, int WARPS, int MINB, bool TICKETS>
__global__ void __launch_bounds__(WARPS * 32, MINB) crop_fwd_nhwc_lean_kernel(const FwdSet S, unsigned nunits, unsigned chunk, int slot, u64 nz) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (TICKETS) {
        unsigned *next = g_fwd_tickets + 2 * slot, *done = next + 1;
        unsigned u = draw_ticket(next, chunk, lane);
        while (u < nunits) {
            const unsigned un = draw_ticket(next, chunk, lane);
            const unsigned ue = min(u + chunk, nunits);
            for (; u < ue; ++u) fwd_unit_lean<VPL, U>(S, u, lane, nz);
            u = un;
        }
        retire_tickets(next, done);
    } else {
        for (unsigned c0 = blockIdx.x * chunk; c0 < nunits; c0 += gridDim.x * chunk) {
            const unsigned c1 = min(c0 + chunk, nunits);
            for (unsigned u = c0 + wib; u < c1; u += WARPS) fwd_unit_lean<VPL, U>(S, u, lane, nz);
        }
    }
}

// Level-batched launch.  Schedule (measured, profiles/r02_fwd_ab_c2_v2.json):
//   * PAIRS: Dev crops every small box twice from the same made-up map (7x7 for the box head, 14x14 for the mask head /
//     critic).  Two consecutive sets with the same map, boxes and list are walked box by box -- the 7 rows of the 7x7 crop, then
//     the 14 rows of the 14x14 crop -- so the second crop finds the box's pixels in L1 / L2 instead of reading them from DRAM a
//     second time (as separate unit ranges the two passes over a 760 MB map are far apart in time): 0.736 -> 0.704 ms on C2.
//   * CHUNKS: the unit sequence is cut into chunks of `chunk` units; chunk c goes to block c % gridDim.x, whose warps stride
//     through it.  chunk = warps per block (the default) deals consecutive units to consecutive warps of the whole grid: at any
//     moment the machine works on ONE window of ~2400 consecutive crop rows (~110 spatially sorted boxes of one image), which
//     is what L2 and the DRAM pages like.  Larger chunks (a block keeps a box and its neighbours to itself) were tried for L1
//     reuse and are monotonically WORSE -- 0.83 ms at 16, 1.00 ms at 256: the window spreads over the whole map.
//   * TICKETS: ncu on the static schedule shows the SMs' active cycles spread from 1.07 M to 1.29 M around a mean of 1.13 M
//     (profiles/r02_ncu_fwd_lean_v2_default.txt): equal SHARES are not equal TIMES -- GPCs hold different numbers of SMs behind the
//     same port into L2, half of L2 is on the other die -- and the launch lasts as long as its slowest SM.  With tickets every
//     warp draws its next `grab` consecutive units from one counter in global memory (the draw for the NEXT units is issued
//     before the current ones are processed, so its round trip is hidden); units are still handed out in order, i.e. the
//     machine-wide window stays contiguous.  C2: 0.703 ms (static) -> 0.587 ms with one unit per draw, SM active cycles 98.7 % of
//     elapsed; more units per draw (rows i, i + 1 of a box to the same warp, for L1) only coarsen the balance: 0.647 ms at 2,
//     0.75 at 4, 0.98 at 14 (profiles/r02_fwd_ab_c2_v4_tickets.json).  The counter resets itself: the last block to finish zeroes
//     it for the next launch that is given the same slot (256 slots, handed out round-robin per launch; a launch baked into a CUDA
//     graph keeps its slot, replays of one graph being ordered).
struct FwdPlan {
    unsigned pair_mask;           // bit k: sets k and k + 1 are interleaved by box (k + 1 is then skipped)
    int chunk;                    // static schedule: units per block chunk
    int grab;                     // ticket schedule: units per draw; 0 = static schedule
    int slot;                     // ticket schedule: which counter pair of g_fwd_tickets
};

template <int VPL, int U, int WARPS, int MINB, bool TICKETS>
__global__ void __launch_bounds__(WARPS * 32, MINB) crop_fwd_nhwc_sets_lean_kernel(const FwdSets sets, const FwdPlan plan, u64 nz) {
    // groups = single sets or pairs; prefix of their LIVE unit counts (see crop_fwd_nhwc_sets_kernel)
    __shared__ unsigned gfirst[kMaxFwdSets + 1];
    __shared__ int gset[kMaxFwdSets], ngroups;
    if (threadIdx.x == 0) {
        unsigned acc = 0;
        int g = 0;
        for (int k = 0; k < sets.n; ++k) {
            const FwdSet &S = sets.s[k];
            const int R = S.R_dev ? max(0, min(*S.R_dev, S.R)) : S.R;
            gfirst[g] = acc;
            gset[g] = k;
            unsigned rows = (unsigned)S.ph;
            if ((plan.pair_mask >> k) & 1u) rows += (unsigned)sets.s[++k].ph;
            acc += (unsigned)R * rows * (unsigned)S.slabs;
            ++g;
        }
        gfirst[g] = acc;
        ngroups = g;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned total = gfirst[ngroups];
    auto unit = [&](unsigned u) {                                  // unit -> (set, box, crop row, slab); ONE inlined copy of the unit
        int g = 0;
        while (u >= gfirst[g + 1]) ++g;
        int k = gset[g];
        const unsigned v = u - gfirst[g];
        const unsigned slabs = (unsigned)sets.s[k].slabs, ua = (unsigned)sets.s[k].ph * slabs;
        unsigned per = ua;
        if ((plan.pair_mask >> k) & 1u) per += (unsigned)sets.s[k + 1].ph * slabs;
        const unsigned r = v / per;
        unsigned w = v - r * per;
        if (w >= ua) { w -= ua; ++k; }                             // second set of a pair
        const unsigned i = w / slabs;
        fwd_unit_lean<VPL, U>(sets.s[k], (int)r, (int)i, (int)(w - i * slabs), lane, nz);
    };
    if (TICKETS) {
        unsigned *next = g_fwd_tickets + 2 * plan.slot, *done = next + 1;
        const unsigned grab = (unsigned)plan.grab;
        unsigned u = draw_ticket(next, grab, lane);
        while (u < total) {
            const unsigned un = draw_ticket(next, grab, lane);     // in flight while this draw's units are processed
            const unsigned ue = min(u + grab, total);
            for (; u < ue; ++u) unit(u);
            u = un;
        }
        retire_tickets(next, done);
    } else {
        const unsigned chunk = (unsigned)plan.chunk;
        for (unsigned c0 = blockIdx.x * chunk; c0 < total; c0 += gridDim.x * chunk) {
            const unsigned c1 = min(c0 + chunk, total);
            for (unsigned u = c0 + wib; u < c1; u += WARPS) unit(u);
        }
    }
}

// Scalar-channel variant for C % 4 != 0 or unaligned pointers (same structure, one float per lane).
__global__ void __launch_bounds__(kWarpsPerBlock * 32) crop_fwd_nhwc_scalar_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                                                                                  const int *__restrict__ box_ind, const int *__restrict__ dst_row,
                                                                                  long nunits, int B, int H, int W, int ph, int pw, int C, float extrap,
                                                                                  float *__restrict__ crops) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long u = warp; u < nunits; u += nwarps) {
        const int r = (int)(u / ph), i = (int)(u - (long)r * ph);
        const int b = box_ind[r];
        const long orow = dst_row ? (long)dst_row[r] : (long)r;
        float *out = crops + ((orow * ph + i) * (long)pw) * C;
        const bool bad = (b < 0 || b >= B);
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        if (bad || !ty.inside) {
            const float v = bad ? 0.f : extrap;
            for (int e = lane; e < pw * C; e += 32) out[e] = v;
            continue;
        }
        const float sx = axis_step(x1, x2, W, pw);
        const float *rowT = image + ((long)b * H + ty.lo) * (long)W * C;
        const float *rowB = image + ((long)b * H + ty.hi) * (long)W * C;
        AxisTap mine;
        for (int j = 0; j < pw; ++j) {
            if ((j & 31) == 0) mine = axis_sample(x1, x2, sx, j + lane, W, pw);
            const AxisTap tx = shfl_tap(mine, j & 31);
            float *o = out + (long)j * C;
            for (int c = lane; c < C; c += 32) {
                float v = extrap;
                if (tx.inside) {
                    const float top = lerp_rn(__ldg(rowT + (long)tx.lo * C + c), __ldg(rowT + (long)tx.hi * C + c), tx.frac);
                    const float bot = lerp_rn(__ldg(rowB + (long)tx.lo * C + c), __ldg(rowB + (long)tx.hi * C + c), tx.frac);
                    v = lerp_rn(top, bot, ty.frac);
                }
                o[c] = v;
            }
        }
    }
}

template <int U>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) crop_bwd_nhwc_kernel(const BwdSet S, long nunits) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long u = warp; u < nunits; u += nwarps) bwd_unit<U>(S, u, lane);
}

constexpr int kMaxBwdSets = 12;
struct BwdSets {
    BwdSet s[kMaxBwdSets];
    long first_unit[kMaxBwdSets + 1];
    int n;
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32) crop_bwd_nhwc_sets_kernel(const BwdSets sets) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const long total = sets.first_unit[sets.n];
    for (long u = warp; u < total; u += nwarps) {
        int k = 0;
        while (u >= sets.first_unit[k + 1]) ++k;
        bwd_unit<4>(sets.s[k], u - sets.first_unit[k], lane);
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) crop_bwd_nhwc_scalar_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                                                                                  const int *__restrict__ box_ind, const int *__restrict__ src_row,
                                                                                  long nunits, int B, int H, int W, int ph, int pw, int C,
                                                                                  float *__restrict__ gimg) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long u = warp; u < nunits; u += nwarps) {
        const int r = (int)(u / ph), i = (int)(u - (long)r * ph);
        const int b = box_ind[r];
        if (b < 0 || b >= B) continue;
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        if (!ty.inside) continue;
        const float sx = axis_step(x1, x2, W, pw);
        const float wy_hi = ty.frac, wy_lo = __fsub_rn(1.f, ty.frac);
        const long grow = src_row ? (long)src_row[r] : (long)r;
        const float *g = grads + ((grow * ph + i) * (long)pw) * C;
        float *rowT = gimg + ((long)b * H + ty.lo) * (long)W * C;
        float *rowB = gimg + ((long)b * H + ty.hi) * (long)W * C;
        AxisTap mine;
        for (int j = 0; j < pw; ++j) {
            if ((j & 31) == 0) mine = axis_sample(x1, x2, sx, j + lane, W, pw);
            const AxisTap tx = shfl_tap(mine, j & 31);
            if (!tx.inside) continue;
            const float wx_hi = tx.frac, wx_lo = __fsub_rn(1.f, tx.frac);
            for (int c = lane; c < C; c += 32) {
                const float gv = g[(long)j * C + c];
                const float dtop = __fmul_rn(wy_lo, gv), dbot = __fmul_rn(wy_hi, gv);
                atomicAdd(rowT + (long)tx.lo * C + c, __fmul_rn(wx_lo, dtop));
                atomicAdd(rowT + (long)tx.hi * C + c, __fmul_rn(wx_hi, dtop));
                atomicAdd(rowB + (long)tx.lo * C + c, __fmul_rn(wx_lo, dbot));
                atomicAdd(rowB + (long)tx.hi * C + c, __fmul_rn(wx_hi, dbot));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NCHW forward / backward (drop-in layout).  One block per (box, channel chunk): the P*P tap table of
// the box is computed once into shared memory, then threads sweep (channel, i, j) with j fastest so the
// crop is written fully coalesced.  Reads are 4 L1-cached scalar loads per output.
// ------------------------------------------------------------------------------------------------
constexpr int kNchwMaxTaps = 64;       // per axis, held in shared memory
constexpr int kNchwChunk = 32;         // channels per block

struct SmemTaps {
    int ylo[kNchwMaxTaps], yhi[kNchwMaxTaps], xlo[kNchwMaxTaps], xhi[kNchwMaxTaps];
    float yf[kNchwMaxTaps], xf[kNchwMaxTaps];
    unsigned char yin[kNchwMaxTaps], xin[kNchwMaxTaps];
};

__device__ __forceinline__ void fill_taps(SmemTaps &t, const float *boxes, int r, int H, int W, int ph, int pw) {
    const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
    const float sy = axis_step(y1, y2, H, ph), sx = axis_step(x1, x2, W, pw);
    for (int k = threadIdx.x; k < ph; k += blockDim.x) {
        const AxisTap a = axis_sample(y1, y2, sy, k, H, ph);
        t.ylo[k] = a.lo; t.yhi[k] = a.hi; t.yf[k] = a.frac; t.yin[k] = a.inside;
    }
    for (int k = threadIdx.x; k < pw; k += blockDim.x) {
        const AxisTap a = axis_sample(x1, x2, sx, k, W, pw);
        t.xlo[k] = a.lo; t.xhi[k] = a.hi; t.xf[k] = a.frac; t.xin[k] = a.inside;
    }
}

__global__ void __launch_bounds__(256) crop_fwd_nchw_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                                                           const int *__restrict__ box_ind, const int *__restrict__ dst_row,
                                                           int B, int H, int W, int ph, int pw, int C, float extrap,
                                                           float *__restrict__ crops) {
    __shared__ SmemTaps t;
    const int r = blockIdx.x;
    const int c0 = blockIdx.y * kNchwChunk;
    const int nc = min(kNchwChunk, C - c0);
    const int pp = ph * pw;
    const int b = box_ind[r];
    const long orow = dst_row ? (long)dst_row[r] : (long)r;
    float *out = crops + (orow * C + c0) * (long)pp;
    if (b < 0 || b >= B) {
        for (int e = threadIdx.x; e < nc * pp; e += blockDim.x) out[e] = 0.f;
        return;
    }
    fill_taps(t, boxes, r, H, W, ph, pw);
    __syncthreads();
    const float *img = image + ((long)b * C + c0) * H * W;
    for (int e = threadIdx.x; e < nc * pp; e += blockDim.x) {
        const int c = e / pp;
        const int s = e - c * pp;
        const int i = s / pw, j = s - i * pw;
        float v = extrap;
        if (t.yin[i] && t.xin[j]) {
            const float *p = img + (long)c * H * W;
            const float *rt = p + (long)t.ylo[i] * W, *rb = p + (long)t.yhi[i] * W;
            const float top = lerp_rn(__ldg(rt + t.xlo[j]), __ldg(rt + t.xhi[j]), t.xf[j]);
            const float bot = lerp_rn(__ldg(rb + t.xlo[j]), __ldg(rb + t.xhi[j]), t.xf[j]);
            v = lerp_rn(top, bot, t.yf[i]);
        }
        __stcs(out + e, v);
    }
}

__global__ void __launch_bounds__(256) crop_bwd_nchw_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                                                           const int *__restrict__ box_ind, const int *__restrict__ src_row,
                                                           int B, int H, int W, int ph, int pw, int C, float *__restrict__ gimg) {
    __shared__ SmemTaps t;
    const int r = blockIdx.x;
    const int c0 = blockIdx.y * kNchwChunk;
    const int nc = min(kNchwChunk, C - c0);
    const int pp = ph * pw;
    const int b = box_ind[r];
    if (b < 0 || b >= B) return;
    fill_taps(t, boxes, r, H, W, ph, pw);
    __syncthreads();
    const long grow = src_row ? (long)src_row[r] : (long)r;
    const float *g = grads + (grow * C + c0) * (long)pp;
    float *img = gimg + ((long)b * C + c0) * H * W;
    for (int e = threadIdx.x; e < nc * pp; e += blockDim.x) {
        const int c = e / pp;
        const int s = e - c * pp;
        const int i = s / pw, j = s - i * pw;
        if (!(t.yin[i] && t.xin[j])) continue;
        const float gv = __ldcs(g + e);
        float *p = img + (long)c * H * W;
        float *rt = p + (long)t.ylo[i] * W, *rb = p + (long)t.yhi[i] * W;
        const float dtop = __fmul_rn(__fsub_rn(1.f, t.yf[i]), gv), dbot = __fmul_rn(t.yf[i], gv);
        const float wl = __fsub_rn(1.f, t.xf[j]), wh = t.xf[j];
        atomicAdd(rt + t.xlo[j], __fmul_rn(wl, dtop));
        atomicAdd(rt + t.xhi[j], __fmul_rn(wh, dtop));
        atomicAdd(rb + t.xlo[j], __fmul_rn(wl, dbot));
        atomicAdd(rb + t.xhi[j], __fmul_rn(wh, dbot));
    }
}

// Generic fallback for crops wider than the shared tap table: thread per output element.
__global__ void crop_fwd_nchw_generic_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                                             const int *__restrict__ box_ind, const int *__restrict__ dst_row, long total,
                                             int B, int H, int W, int ph, int pw, int C, float extrap, float *__restrict__ crops) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long idx = e;
        const int j = (int)(idx % pw); idx /= pw;
        const int i = (int)(idx % ph); idx /= ph;
        const int c = (int)(idx % C);
        const int r = (int)(idx / C);
        const long orow = dst_row ? (long)dst_row[r] : (long)r;
        float *o = crops + ((orow * C + c) * ph + i) * (long)pw + j;
        const int b = box_ind[r];
        if (b < 0 || b >= B) { *o = 0.f; continue; }
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        const AxisTap tx = axis_sample(x1, x2, axis_step(x1, x2, W, pw), j, W, pw);
        float v = extrap;
        if (ty.inside && tx.inside) {
            const float *p = image + ((long)b * C + c) * H * W;
            const float top = lerp_rn(__ldg(p + (long)ty.lo * W + tx.lo), __ldg(p + (long)ty.lo * W + tx.hi), tx.frac);
            const float bot = lerp_rn(__ldg(p + (long)ty.hi * W + tx.lo), __ldg(p + (long)ty.hi * W + tx.hi), tx.frac);
            v = lerp_rn(top, bot, ty.frac);
        }
        *o = v;
    }
}

__global__ void crop_bwd_nchw_generic_kernel(const float *__restrict__ grads, const float *__restrict__ boxes,
                                             const int *__restrict__ box_ind, const int *__restrict__ src_row, long total,
                                             int B, int H, int W, int ph, int pw, int C, float *__restrict__ gimg) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long idx = e;
        const int j = (int)(idx % pw); idx /= pw;
        const int i = (int)(idx % ph); idx /= ph;
        const int c = (int)(idx % C);
        const int r = (int)(idx / C);
        const int b = box_ind[r];
        if (b < 0 || b >= B) continue;
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1], y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const AxisTap ty = axis_sample(y1, y2, axis_step(y1, y2, H, ph), i, H, ph);
        const AxisTap tx = axis_sample(x1, x2, axis_step(x1, x2, W, pw), j, W, pw);
        if (!(ty.inside && tx.inside)) continue;
        const long grow = src_row ? (long)src_row[r] : (long)r;
        const float gv = grads[((grow * C + c) * ph + i) * (long)pw + j];
        float *p = gimg + ((long)b * C + c) * H * W;
        const float dtop = __fmul_rn(__fsub_rn(1.f, ty.frac), gv), dbot = __fmul_rn(ty.frac, gv);
        const float wl = __fsub_rn(1.f, tx.frac), wh = tx.frac;
        atomicAdd(p + (long)ty.lo * W + tx.lo, __fmul_rn(wl, dtop));
        atomicAdd(p + (long)ty.lo * W + tx.hi, __fmul_rn(wh, dtop));
        atomicAdd(p + (long)ty.hi * W + tx.lo, __fmul_rn(wl, dbot));
        atomicAdd(p + (long)ty.hi * W + tx.hi, __fmul_rn(wh, dbot));
    }
}

static int grid_for(long work_items, int per_block, int blocks_per_sm) {
    long need = (work_items + per_block - 1) / per_block;
    long cap = (long)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static int check_common(const void *a, const void *boxes, const void *box_ind, const void *out, int R, int B, int H, int W,
                        int ph, int pw, int C) {
    FI_REQUIRE(R >= 0 && B > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && C > 0, "crop_and_resize: bad sizes R=%d B=%d H=%d W=%d ph=%d pw=%d C=%d", R, B, H, W, ph, pw, C);
    FI_REQUIRE(R == 0 || (a && boxes && box_ind && out), "crop_and_resize: null pointer");
    return FI_OK;
}

}  // namespace fi

using namespace fi;

// NHWC scatter backward (reduction kernels); gimg must already hold the values to accumulate onto.
int fi_scatter_backward_nhwc(const float *grads, const float *grads2, const float *boxes, const int *box_ind, const int *src_row, int R, int B,
                             int H, int W, int ph, int pw, int C, float *gimg, cudaStream_t stream) {
    if (R == 0) return ok();
    const bool vec = (C % 128 == 0) && ((uintptr_t)gimg % 16 == 0) && ((uintptr_t)grads % 16 == 0) && ((uintptr_t)grads2 % 16 == 0);
    if (vec) {
        BwdSet S;
        S.grads = grads; S.grads2 = grads2; S.boxes = boxes; S.box_ind = box_ind; S.src_row = src_row; S.gimg = gimg;
        S.B = B; S.H = H; S.W = W; S.C = C; S.ph = ph; S.pw = pw; S.slabs = C / 128;
        const long nunits = (long)R * ph * S.slabs;
        crop_bwd_nhwc_kernel<4><<<grid_for(nunits, kWarpsPerBlock, 8), kWarpsPerBlock * 32, 0, stream>>>(S, nunits);
        return check_launch("fi_crop_and_resize_backward[nhwc scatter]");
    }
    const long nunits = (long)R * ph;
    const int grid = grid_for(nunits, kWarpsPerBlock, 8);
    crop_bwd_nhwc_scalar_kernel<<<grid, kWarpsPerBlock * 32, 0, stream>>>(grads, boxes, box_ind, src_row, nunits, B, H, W, ph, pw, C, gimg);
    if (int e = check_launch("fi_crop_and_resize_backward[nhwc scatter]")) return e;
    if (grads2) {
        crop_bwd_nhwc_scalar_kernel<<<grid, kWarpsPerBlock * 32, 0, stream>>>(grads2, boxes, box_ind, nullptr, nunits, B, H, W, ph, pw, C, gimg);
        return check_launch("fi_crop_and_resize_backward[nhwc scatter 2]");
    }
    return ok();
}

FI_API int fi_crop_taps(const float *boxes, int num_boxes, int H, int W, int ph, int pw, int *taps, cudaStream_t stream) {
    FI_REQUIRE(num_boxes >= 0 && H > 0 && W > 0 && ph > 0 && pw > 0, "fi_crop_taps: bad sizes");
    if (num_boxes == 0) return ok();
    FI_REQUIRE(boxes && taps, "fi_crop_taps: null pointer");
    const long total = (long)num_boxes * ph * pw;
    crop_taps_kernel<<<grid_for(total, 256, 8), 256, 0, stream>>>(boxes, num_boxes, H, W, ph, pw, taps);
    return check_launch("fi_crop_taps");
}

int fi_nchw_backward_via_nhwc(const float *grads, const float *boxes, const int *box_ind, const int *src_row, int R, int B, int H, int W, int ph,
                              int pw, int C, float *gimg, int accumulate, cudaStream_t stream);   // roi_align_nchw_bwd.cu

// ---- lean forward launchers (FI_OPT_FWD_FORM) -------------------------------------------------------------------------------
// form 0 = default (kLeanDefault), 1 = round-1 unit (fwd_unit), 2.. = the other lean shapes kept for A/B runs.
struct LeanShape { int vpl, u, warps, minb; };
static const LeanShape kLeanShapes[] = {
    {2, 2, 4, 5},   // form 2: both slabs per warp, 16 tap loads in flight, 20 warps / SM in blocks of 4 (96 registers)
    {2, 2, 8, 2},   // form 3: 16 warps / SM in blocks of 8 (112 registers, nothing spilled) -- the default
    {2, 2, 16, 1},  // form 4: one block of 16 warps / SM
    {1, 4, 8, 2},   // form 5: one slab per warp, 16 loads in flight (also what depth % 256 != 0 falls back to)
    {2, 2, 4, 4},   // form 6: blocks of 4 warps without the 96-register cap
};
constexpr int kLeanDefault = 3;
constexpr int kLeanGrabDefault = 1;       // ticket schedule: units per draw
constexpr int kLeanChunkDefault = 0;      // 0: one unit per warp per pass (chunk = warps per block)

static int lean_shape_index() {           // -1: round-1 unit
    int form = option(FI_OPT_FWD_FORM);
    if (form == 0) form = kLeanDefault;
    if (form == 1) return -1;
    return form - 2;
}

static int lean_chunk(int warps) {         // units per block chunk: FI_OPT_FWD_CHUNK 1..6 -> 8, 16, 32, 64, 128, 256
    const int copt = option(FI_OPT_FWD_CHUNK);
    const int chunk = copt ? (4 << copt) : kLeanChunkDefault;
    return chunk < warps ? warps : chunk;
}

// ticket schedule (default): units per draw (FI_OPT_FWD_CHUNK 1..6 -> 1, 2, 3, 4, 7, 14) and the counter slot of this launch; 0: static
static int lean_grab(int *slot) {
    if (option(FI_OPT_FWD_SCHED) == 1) return 0;
    static const int kGrab[7] = {kLeanGrabDefault, 1, 2, 3, 4, 7, 14};
    static unsigned next_slot = 0;
    *slot = (int)(__atomic_fetch_add(&next_slot, 1u, __ATOMIC_RELAXED) % kTicketSlots);
    return kGrab[option(FI_OPT_FWD_CHUNK)];
}

static bool lean_ok(const void *image, const void *boxes, const void *crops, const void *crops2, int W, int C, int vpl, long units) {
    return (C % (128 * vpl) == 0) && ((long)W * C * 4 <= (1L << 30)) && units < (1L << 31) && ((uintptr_t)image % 16 == 0) &&
           ((uintptr_t)boxes % 16 == 0) && ((uintptr_t)crops % 16 == 0) && ((uintptr_t)crops2 % 16 == 0);
}

template <int VPL, int U, int WARPS, int MINB>
static void launch_lean_one(const FwdSet &S, unsigned nunits, cudaStream_t stream) {
    int slot = 0;
    const int grab = lean_grab(&slot);
    if (grab > 0) {
        const unsigned per_block = (unsigned)grab * WARPS;
        crop_fwd_nhwc_lean_kernel<VPL, U, WARPS, MINB, true><<<grid_for((nunits + per_block - 1) / per_block, 1, MINB), WARPS * 32, 0, stream>>>(
            S, nunits, (unsigned)grab, slot, kNegZeroPair);
    } else {
        const unsigned chunk = (unsigned)lean_chunk(WARPS);
        crop_fwd_nhwc_lean_kernel<VPL, U, WARPS, MINB, false><<<grid_for((nunits + chunk - 1) / chunk, 1, MINB), WARPS * 32, 0, stream>>>(
            S, nunits, chunk, 0, kNegZeroPair);
    }
}
template <int VPL, int U, int WARPS, int MINB>
static void launch_lean_sets(const FwdSets &sets, const FwdPlan &plan, long units, cudaStream_t stream) {
    if (plan.grab > 0) {
        const long per_block = (long)plan.grab * WARPS;
        crop_fwd_nhwc_sets_lean_kernel<VPL, U, WARPS, MINB, true><<<grid_for((units + per_block - 1) / per_block, 1, MINB), WARPS * 32, 0, stream>>>(sets, plan, kNegZeroPair);
    } else {
        crop_fwd_nhwc_sets_lean_kernel<VPL, U, WARPS, MINB, false><<<grid_for((units + plan.chunk - 1) / plan.chunk, 1, MINB), WARPS * 32, 0, stream>>>(sets, plan, kNegZeroPair);
    }
}
#define FI_LEAN_DISPATCH(idx, CALL)                  \
    switch (idx) {                                   \
        case 0: CALL(2, 2, 4, 5); break;             \
        case 1: CALL(2, 2, 8, 2); break;             \
        case 2: CALL(2, 2, 16, 1); break;            \
        case 3: CALL(1, 4, 8, 2); break;             \
        default: CALL(2, 2, 4, 4); break;            \
    }

static int forward_impl(const float *image, int image_layout, const float *boxes, const int *box_ind, const int *dst_row, int R, int B, int H,
                        int W, int ph, int pw, int C, float extrap, float *crops, int crops_layout, float *crops2, cudaStream_t stream) {
    if (int e = check_common(image, boxes, box_ind, crops, R, B, H, W, ph, pw, C)) return e;
    if (R == 0) return ok();
    if (image_layout != crops_layout) {
        set_error(FI_ERR_UNSUPPORTED, "fi_crop_and_resize_forward: mixed layouts (image %d, crops %d)", image_layout, crops_layout);
        return FI_ERR_UNSUPPORTED;
    }
    if (image_layout == FI_LAYOUT_NHWC) {
        const bool vec = (C % 128 == 0) && (W <= 32768) && ((uintptr_t)image % 16 == 0) && ((uintptr_t)crops % 16 == 0) && ((uintptr_t)crops2 % 16 == 0);
        if (vec) {
            FwdSet S;
            S.image = image; S.boxes = boxes; S.box_ind = box_ind; S.dst_row = dst_row; S.R_dev = nullptr; S.crops = crops; S.crops2 = crops2;
            S.B = B; S.H = H; S.W = W; S.C = C; S.ph = ph; S.pw = pw; S.slabs = C / 128; S.R = R; S.extrap = extrap;
            int li = lean_shape_index();
            if (li >= 0 && !lean_ok(image, boxes, crops, crops2, W, C, kLeanShapes[li].vpl, (long)R * ph * S.slabs)) li = (kLeanShapes[li].vpl == 2) ? 3 : -1;   // one slab per warp
            if (li >= 0 && lean_ok(image, boxes, crops, crops2, W, C, kLeanShapes[li].vpl, (long)R * ph * S.slabs)) {
                S.slabs = C / (128 * kLeanShapes[li].vpl);
                const unsigned nu = (unsigned)((long)R * ph * S.slabs);
#define FI_CALL_ONE(V, UU, WW, MB) launch_lean_one<V, UU, WW, MB>(S, nu, stream)
                FI_LEAN_DISPATCH(li, FI_CALL_ONE)
#undef FI_CALL_ONE
                return check_launch("fi_crop_and_resize_forward[nhwc lean]");
            }
            const long nunits = (long)R * ph * S.slabs;  // one warp per (crop row, 128-channel slab)
            const int grid = grid_for(nunits, kWarpsPerBlock, 8);
            if (pw % 4 == 0 || pw > 12)
                crop_fwd_nhwc_kernel<4><<<grid, kWarpsPerBlock * 32, 0, stream>>>(S, nunits);
            else
                crop_fwd_nhwc_kernel<2><<<grid, kWarpsPerBlock * 32, 0, stream>>>(S, nunits);
        } else {
            if (crops2) { set_error(FI_ERR_UNSUPPORTED, "fi_crop_and_resize_forward_dual needs depth %% 128 == 0 and 16-byte aligned tensors"); return FI_ERR_UNSUPPORTED; }
            const long nunits = (long)R * ph;
            crop_fwd_nhwc_scalar_kernel<<<grid_for(nunits, kWarpsPerBlock, 8), kWarpsPerBlock * 32, 0, stream>>>(image, boxes, box_ind, dst_row, nunits, B, H, W, ph, pw, C, extrap, crops);
        }
        return check_launch("fi_crop_and_resize_forward[nhwc]");
    }
    if (image_layout == FI_LAYOUT_NCHW) {
        if (crops2) { set_error(FI_ERR_UNSUPPORTED, "fi_crop_and_resize_forward_dual is NHWC only"); return FI_ERR_UNSUPPORTED; }
        if (ph <= kNchwMaxTaps && pw <= kNchwMaxTaps) {
            dim3 grid(R, ceil_div(C, kNchwChunk));
            crop_fwd_nchw_kernel<<<grid, 256, 0, stream>>>(image, boxes, box_ind, dst_row, B, H, W, ph, pw, C, extrap, crops);
        } else {
            const long total = (long)R * C * ph * pw;
            crop_fwd_nchw_generic_kernel<<<grid_for(total, 256, 8), 256, 0, stream>>>(image, boxes, box_ind, dst_row, total, B, H, W,
                                                                                     ph, pw, C, extrap, crops);
        }
        return check_launch("fi_crop_and_resize_forward[nchw]");
    }
    set_error(FI_ERR_INVALID, "fi_crop_and_resize_forward: unknown layout %d", image_layout);
    return FI_ERR_INVALID;
}

FI_API int fi_crop_and_resize_forward(const float *image, int image_layout, const float *boxes, const int *box_ind,
                                      const int *dst_row, int R, int B, int H, int W, int ph, int pw, int C, float extrap,
                                      float *crops, int crops_layout, cudaStream_t stream) {
    return forward_impl(image, image_layout, boxes, box_ind, dst_row, R, B, H, W, ph, pw, C, extrap, crops, crops_layout, nullptr, stream);
}

FI_API int fi_crop_and_resize_forward_dual(const float *image, const float *boxes, const int *box_ind, const int *dst_row, int R, int B, int H,
                                           int W, int ph, int pw, int C, float extrap, float *crops, float *crops_compact, cudaStream_t stream) {
    FI_REQUIRE(R == 0 || (dst_row && crops_compact), "fi_crop_and_resize_forward_dual: dst_row and crops_compact are required");
    return forward_impl(image, FI_LAYOUT_NHWC, boxes, box_ind, dst_row, R, B, H, W, ph, pw, C, extrap, crops, FI_LAYOUT_NHWC, crops_compact, stream);
}

FI_API int fi_crop_and_resize_backward(const float *grads, int grads_layout, const float *boxes, const int *box_ind,
                                       const int *src_row, int R, int B, int H, int W, int ph, int pw, int C,
                                       float *gimg, int image_layout, int accumulate, cudaStream_t stream) {
    FI_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && gimg, "fi_crop_and_resize_backward: bad image");
    if (image_layout == FI_LAYOUT_NHWC && grads_layout == FI_LAYOUT_NHWC) {
        // reduction kernels, or the deterministic write-once gather when fi_set_deterministic(1) -- roi_align_bwd.cu
        fi_crop_set one;
        one.grads = grads; one.grads2 = nullptr; one.boxes = boxes; one.box_ind = box_ind; one.src_row = src_row;
        one.num_boxes = R; one.crop_height = ph; one.crop_width = pw;
        FI_REQUIRE(R >= 0 && ph > 0 && pw > 0, "fi_crop_and_resize_backward: bad sizes");
        return fi_crop_and_resize_backward_multi(&one, 1, B, H, W, C, gimg, accumulate, fi_get_deterministic(), stream);
    }
    if (int e = check_common(grads, boxes, box_ind, gimg, R, B, H, W, ph, pw, C)) return e;
    if (image_layout == FI_LAYOUT_NCHW && grads_layout == FI_LAYOUT_NCHW && R > 0) {
        // transposed through the NHWC tile-owner kernels when the shape qualifies (roi_align_nchw_bwd.cu): no zero fill, no atomics
        const int rc = fi_nchw_backward_via_nhwc(grads, boxes, box_ind, src_row, R, B, H, W, ph, pw, C, gimg, accumulate, stream);
        if (rc != FI_ERR_UNSUPPORTED) return rc;
    }
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(gimg, 0, sizeof(float) * (size_t)B * C * H * W, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_crop_and_resize_backward: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    }
    if (R == 0) return ok();
    if (image_layout != grads_layout) {
        set_error(FI_ERR_UNSUPPORTED, "fi_crop_and_resize_backward: mixed layouts (grads %d, image %d)", grads_layout, image_layout);
        return FI_ERR_UNSUPPORTED;
    }
    if (image_layout == FI_LAYOUT_NHWC) return fi_scatter_backward_nhwc(grads, nullptr, boxes, box_ind, src_row, R, B, H, W, ph, pw, C, gimg, stream);
    if (image_layout == FI_LAYOUT_NCHW) {
        if (ph <= kNchwMaxTaps && pw <= kNchwMaxTaps) {
            dim3 grid(R, ceil_div(C, kNchwChunk));
            crop_bwd_nchw_kernel<<<grid, 256, 0, stream>>>(grads, boxes, box_ind, src_row, B, H, W, ph, pw, C, gimg);
        } else {
            const long total = (long)R * C * ph * pw;
            crop_bwd_nchw_generic_kernel<<<grid_for(total, 256, 8), 256, 0, stream>>>(grads, boxes, box_ind, src_row, total, B, H, W,
                                                                                     ph, pw, C, gimg);
        }
        return check_launch("fi_crop_and_resize_backward[nchw]");
    }
    set_error(FI_ERR_INVALID, "fi_crop_and_resize_backward: unknown layout %d", image_layout);
    return FI_ERR_INVALID;
}

FI_API int fi_crop_sets_forward(const fi_fwd_set *sets, int num_sets, cudaStream_t stream) {
    FI_REQUIRE(sets && num_sets >= 1 && num_sets <= kMaxFwdSets, "fi_crop_sets_forward: 1..%d sets", kMaxFwdSets);
    FwdSets dev;
    dev.n = 0;
    long units = 0;
    for (int i = 0; i < num_sets; ++i) {
        const fi_fwd_set &h = sets[i];
        if (int e = check_common(h.image, h.boxes, h.box_ind, h.crops, h.num_boxes, h.batch, h.image_height, h.image_width, h.crop_height,
                                 h.crop_width, h.depth)) return e;
        const bool vec = (h.depth % 128 == 0) && (h.image_width <= 32768) && ((uintptr_t)h.image % 16 == 0) && ((uintptr_t)h.crops % 16 == 0) &&
                         ((uintptr_t)h.crops_compact % 16 == 0);
        if (!vec) { set_error(FI_ERR_UNSUPPORTED, "fi_crop_sets_forward: set %d needs NHWC, depth %% 128 == 0, 16-byte aligned tensors", i); return FI_ERR_UNSUPPORTED; }
        if (h.num_boxes == 0) continue;
        FwdSet &S = dev.s[dev.n];
        S.image = h.image; S.boxes = h.boxes; S.box_ind = h.box_ind; S.dst_row = h.dst_row; S.R_dev = h.num_boxes_dev; S.crops = h.crops; S.crops2 = h.crops_compact;
        S.B = h.batch; S.H = h.image_height; S.W = h.image_width; S.C = h.depth; S.ph = h.crop_height; S.pw = h.crop_width;
        S.slabs = h.depth / 128; S.R = h.num_boxes; S.extrap = h.extrapolation_value;
        dev.first_unit[dev.n] = units;
        units += (long)h.num_boxes * h.crop_height * S.slabs;
        ++dev.n;
    }
    dev.first_unit[dev.n] = units;
    if (units == 0) return ok();
    {   // lean formulation when every set qualifies for the chosen shape (else: a one-slab lean shape, else the round-1 unit)
        int li = lean_shape_index();
        for (int pass = 0; pass < 2 && li >= 0; ++pass) {
            bool all = true;
            long lu = 0;
            for (int k = 0; k < dev.n; ++k) {
                const FwdSet &S = dev.s[k];
                lu += (long)S.R * S.ph * (S.C / (128 * kLeanShapes[li].vpl));
                all = all && lean_ok(S.image, S.boxes, S.crops, S.crops2, S.W, S.C, kLeanShapes[li].vpl, 0);
            }
            all = all && lu < (1L << 31);
            if (all) {
                for (int k = 0; k < dev.n; ++k) dev.s[k].slabs = dev.s[k].C / (128 * kLeanShapes[li].vpl);
                FwdPlan plan;
                const int popt = option(FI_OPT_FWD_PAIR);
                plan.chunk = lean_chunk(kLeanShapes[li].warps);
                plan.slot = 0;
                plan.grab = lean_grab(&plan.slot);
                plan.pair_mask = 0;
                if (popt != 1)
                    for (int k = 0; k + 1 < dev.n; ++k) {
                        const FwdSet &A = dev.s[k], &Bs = dev.s[k + 1];
                        if (A.image == Bs.image && A.boxes == Bs.boxes && A.box_ind == Bs.box_ind && A.R == Bs.R && A.R_dev == Bs.R_dev &&
                            A.C == Bs.C && A.B == Bs.B && A.H == Bs.H && A.W == Bs.W) {
                            plan.pair_mask |= 1u << k;
                            ++k;
                        }
                    }
#define FI_CALL_SETS(V, UU, WW, MB) launch_lean_sets<V, UU, WW, MB>(dev, plan, lu, stream)
                FI_LEAN_DISPATCH(li, FI_CALL_SETS)
#undef FI_CALL_SETS
                return check_launch("fi_crop_sets_forward[lean]");
            }
            li = (kLeanShapes[li].vpl == 2) ? 3 : -1;      // retry with one slab per warp
        }
    }
    crop_fwd_nhwc_sets_kernel<<<grid_for(units, kWarpsPerBlock, 8), kWarpsPerBlock * 32, 0, stream>>>(dev);
    return check_launch("fi_crop_sets_forward");
}

int fi_tile_backward(const fi_bwd_set *sets, int num_sets, int accumulate, int exact, cudaStream_t stream);   // roi_align_bwd_tile.cu

FI_API int fi_crop_sets_backward(const fi_bwd_set *sets, int num_sets, int zero_first, cudaStream_t stream) {
    FI_REQUIRE(sets && num_sets >= 1 && num_sets <= kMaxBwdSets, "fi_crop_sets_backward: 1..%d sets", kMaxBwdSets);
    for (int i = 0; i < num_sets; ++i) {
        const fi_bwd_set &h = sets[i];
        FI_REQUIRE(h.grads_image && h.batch > 0 && h.image_height > 0 && h.image_width > 0 && h.depth > 0, "fi_crop_sets_backward: bad map in set %d", i);
        if (int e = check_common(h.grads, h.boxes, h.box_ind, h.grads_image, h.num_boxes, h.batch, h.image_height, h.image_width, h.crop_height,
                                 h.crop_width, h.depth)) return e;
    }
    {   // Formulation.  Default: tile-owner kernels (roi_align_bwd_tile.cu) -- shared-memory accumulation,
        // every map pixel written once, no zero fill, no atomics; exact (bit-identical to crop_and_resize.c) when
        // fi_set_deterministic(1).  fi_set_option(FI_OPT_BWD_FORM, 3): the vector reductions below (also the fallback for shapes
        // the tile kernels do not take: crops wider than 16, more than 8 maps).
        if (option(FI_OPT_BWD_FORM) != 3) {
            const int exact = fi_get_deterministic();
            const int rc = fi_tile_backward(sets, num_sets, zero_first ? 0 : 1, exact, stream);
            if (rc != FI_ERR_UNSUPPORTED) return rc;
        }
    }
    BwdSets dev;
    dev.n = 0;
    long units = 0;
    for (int i = 0; i < num_sets; ++i) {
        const fi_bwd_set &h = sets[i];
        FI_REQUIRE(h.grads_image && h.batch > 0 && h.image_height > 0 && h.image_width > 0 && h.depth > 0, "fi_crop_sets_backward: bad map in set %d", i);
        if (int e = check_common(h.grads, h.boxes, h.box_ind, h.grads_image, h.num_boxes, h.batch, h.image_height, h.image_width, h.crop_height,
                                 h.crop_width, h.depth)) return e;
        const bool vec = (h.depth % 128 == 0) && ((uintptr_t)h.grads_image % 16 == 0) && ((uintptr_t)h.grads % 16 == 0) && ((uintptr_t)h.grads2 % 16 == 0);
        if (!vec) { set_error(FI_ERR_UNSUPPORTED, "fi_crop_sets_backward: set %d needs NHWC, depth %% 128 == 0, 16-byte aligned tensors", i); return FI_ERR_UNSUPPORTED; }
        if (h.num_boxes_dev) { set_error(FI_ERR_UNSUPPORTED, "fi_crop_sets_backward: device-side box counts need the tile-owner kernels (set %d)", i); return FI_ERR_UNSUPPORTED; }
        if (zero_first) {                       // each distinct map once
            bool seen = false;
            for (int q = 0; q < i; ++q) seen = seen || (sets[q].grads_image == h.grads_image);
            if (!seen) {
                cudaError_t e = cudaMemsetAsync(h.grads_image, 0, sizeof(float) * (size_t)h.batch * h.depth * h.image_height * h.image_width, stream);
                if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_crop_sets_backward: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
            }
        }
        if (h.num_boxes == 0) continue;
        BwdSet &S = dev.s[dev.n];
        S.grads = h.grads; S.grads2 = h.grads2; S.boxes = h.boxes; S.box_ind = h.box_ind; S.src_row = h.src_row; S.gimg = h.grads_image;
        S.B = h.batch; S.H = h.image_height; S.W = h.image_width; S.C = h.depth; S.ph = h.crop_height; S.pw = h.crop_width; S.slabs = h.depth / 128;
        dev.first_unit[dev.n] = units;
        units += (long)h.num_boxes * h.crop_height * S.slabs;
        ++dev.n;
    }
    dev.first_unit[dev.n] = units;
    if (units == 0) return ok();
    crop_bwd_nhwc_sets_kernel<<<grid_for(units, kWarpsPerBlock, 8), kWarpsPerBlock * 32, 0, stream>>>(dev);
    return check_launch("fi_crop_sets_backward");
}

// ---- reference-named launchers (lib/roi_align/src/cuda/crop_and_resize_kernel.h:8-18) ------------
FI_API void CropAndResizeLaucher(const float *image_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes, int batch,
                                 int image_height, int image_width, int crop_height, int crop_width, int depth,
                                 float extrapolation_value, float *crops_ptr, cudaStream_t stream) {
    fi_crop_and_resize_forward(image_ptr, FI_LAYOUT_NCHW, boxes_ptr, box_ind_ptr, nullptr, num_boxes, batch, image_height,
                               image_width, crop_height, crop_width, depth, extrapolation_value, crops_ptr, FI_LAYOUT_NCHW, stream);
}

FI_API void CropAndResizeBackpropImageLaucher(const float *grads_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes,
                                              int batch, int image_height, int image_width, int crop_height, int crop_width,
                                              int depth, float *grads_image_ptr, cudaStream_t stream) {
    fi_crop_and_resize_backward(grads_ptr, FI_LAYOUT_NCHW, boxes_ptr, box_ind_ptr, nullptr, num_boxes, batch, image_height,
                                image_width, crop_height, crop_width, depth, grads_image_ptr, FI_LAYOUT_NCHW, /*accumulate=*/1, stream);
}
