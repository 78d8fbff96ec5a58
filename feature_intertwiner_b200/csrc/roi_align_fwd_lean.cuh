// NHWC RoIAlign forward, second formulation ("lean"): the same warp-per-crop-row walk as roi_align_units.cuh::fwd_unit with the
// per-sample instruction count more than halved (445 M -> 193 M warp instructions on C2).
//
// ncu on fwd_unit (profiles/r01_ncu_prof_fwd_sets_v2.txt): 445 M warp instructions for 3.6 M (crop sample, 128-channel slab)
// pairs = ~94 instructions per 4 tap loads, issue slots 51 % busy at 16 warps per SM -- the kernel is as much issue-bound as it
// is latency-bound.  Where the instructions went, and what this unit does instead:
//   * 36 scalar FADD/FMUL/FADD per float4 sample (the reference's un-fused `top + (bottom - top) * w`, crop_and_resize.c:102-106)
//     -> 18 packed two-float instructions (FADD2 / FFMA2 / FADD2).  ptxas CONTRACTS `mul.rn.f32x2` + `add.rn.f32x2` into one FFMA2
//     even though both carry an explicit rounding mode (checked in SASS; --fmad=false does not change it), which would break
//     bit-exactness.  The product is therefore formed as `fma.rn.f32x2 t, d, w, -0.0` with the -0.0 pair arriving as a KERNEL
//     PARAMETER (a value ptxas cannot fold): d*w + (-0) rounds exactly like d*w, signed zeros included, and the following packed
//     add has nothing left to fuse with.
//   * 64-bit multiplies for every tap address and output address -> 32-bit element offsets formed once per lane (lo * C, with the
//     `hi != lo` and `inside` flags in the top bits of the same word, so a sample still costs two shuffles), one 64-bit add per
//     tap address on opaque row base addresses, running output pointers.
//   * a warp takes BOTH 128-channel slabs of a 256-channel pixel (VPL = 2): shuffles, unpacking, address formation and loop
//     control are paid once per 1 KB tap instead of once per 512 B.
//   * 64-bit div/mod chains per unit -> 32-bit.
// Results are bit-identical to fwd_unit (tests/test_roi_align_gpu.py::test_forward_forms_bit_identical).
#pragma once
#include "roi_align_units.cuh"

namespace fi {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

// a + (b - a) * w on two floats at once, every step rounded on its own (see the header: nz must be an opaque (-0.0, -0.0))
__device__ __forceinline__ u64 lerp2_rn(u64 a, u64 b, u64 w, u64 nz) {
    u64 d;
    asm("{\n\t.reg .b64 t;\n\tsub.rn.f32x2 t, %2, %1;\n\tfma.rn.f32x2 t, t, %3, %4;\n\tadd.rn.f32x2 %0, %1, t;\n\t}"
        : "=l"(d) : "l"(a), "l"(b), "l"(w), "l"(nz));
    return d;
}
__device__ __forceinline__ ulonglong2 lerp4_rn(ulonglong2 a, ulonglong2 b, u64 w, u64 nz) {
    return make_ulonglong2(lerp2_rn(a.x, b.x, w, nz), lerp2_rn(a.y, b.y, w, nz));
}
__device__ __forceinline__ ulonglong2 ldg_v4(const float *p) { return __ldg(reinterpret_cast<const ulonglong2 *>(p)); }
__device__ __forceinline__ void st_stream_v4(float *p, ulonglong2 v) { __stcs(reinterpret_cast<ulonglong2 *>(p), v); }

constexpr u64 kNegZeroPair = 0x8000000080000000ull;

// unit = (box r, crop row i, slab of 128 * VPL channels).  W * C * 4 <= 2^30.
template <int VPL, int U>
__device__ __forceinline__ void fwd_unit_lean(const FwdSet &S, int r, int i, int slab, int lane, u64 nz) {
    const float *__restrict__ boxes = S.boxes;
    const int B = S.B, H = S.H, W = S.W, ph = S.ph, pw = S.pw, C = S.C;
    const int coff = lane * 4 + slab * (128 * VPL);
    const int b = __ldg(S.box_ind + r);
    const long orow = S.dst_row ? (long)__ldg(S.dst_row + r) : (long)r;
    const long row_elems = (long)pw * C;
    float *o = S.crops + (orow * ph + i) * row_elems + coff;
    float *o2 = S.crops2 ? S.crops2 + ((long)r * ph + i) * row_elems + coff : nullptr;   // optional compact copy (row r)
    const bool bad = (b < 0 || b >= B);     // reference leaves such rows at their zero fill (crop_and_resize_kernel.cu:34-38)
    const float4 bx = __ldg(reinterpret_cast<const float4 *>(boxes) + r);               // y1, x1, y2, x2
    const AxisTap ty = axis_sample(bx.x, bx.z, axis_step(bx.x, bx.z, H, ph), i, H, ph);
    const u64 ex2 = pack2(S.extrap, S.extrap);
    if (bad || !ty.inside) {
        const ulonglong2 v = bad ? make_ulonglong2(0ull, 0ull) : make_ulonglong2(ex2, ex2);
        for (int j = 0; j < pw; ++j) {
#pragma unroll
            for (int s = 0; s < VPL; ++s) {
                st_stream_v4(o + s * 128, v);
                if (o2) st_stream_v4(o2 + s * 128, v);
            }
            o += C;
            if (o2) o2 += C;
        }
        return;
    }
    const float sx = axis_step(bx.y, bx.w, W, pw);
    // the two image rows as opaque byte addresses: left to itself the compiler keeps 64-bit ELEMENT indices and re-forms every tap
    // address from the image base (IADD3 + IMAD.X + LEA + LEA.HI.X per tap)
    u64 rowT = (u64)(S.image + ((long)b * H + ty.lo) * (long)W * C + coff);
    u64 rowB = (u64)(S.image + ((long)b * H + ty.hi) * (long)W * C + coff);
    asm volatile("" : "+l"(rowT), "+l"(rowB));
    const unsigned Cb = (unsigned)C * 4u;                        // bytes per pixel
    const u64 wy = pack2(ty.frac, ty.frac);
    for (int jb = 0; jb < pw; jb += 32) {                        // lane l computes the x tap of sample jb + l
        const AxisTap mine = axis_sample(bx.y, bx.w, sx, jb + lane, W, pw);
        const int lo_c = min(max(mine.lo, 0), W - 1), hi_c = min(max(mine.hi, 0), W - 1);
        // byte offset of the lo pixel | inside << 30 | (hi != lo) << 31
        const int packed = (int)((unsigned)lo_c * Cb) | (mine.inside ? (1 << 30) : 0) | (hi_c != lo_c ? (int)0x80000000 : 0);
        const int jn = min(32, pw - jb);
        for (int j0 = 0; j0 < jn; j0 += U) {
            ulonglong2 tl[U][VPL], tr[U][VPL], bl[U][VPL], br[U][VPL];
            int pk[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int j = min(j0 + k, jn - 1);               // tail lanes re-read the last sample (discarded)
                pk[k] = __shfl_sync(0xffffffffu, packed, j);
                const unsigned ol = (unsigned)pk[k] & 0x3fffffffu;
                const unsigned oh = ol - (unsigned)(pk[k] >> 31) * Cb;              // + Cb when hi != lo
                const float *pTl = (const float *)(rowT + ol), *pTh = (const float *)(rowT + oh);
                const float *pBl = (const float *)(rowB + ol), *pBh = (const float *)(rowB + oh);
#pragma unroll
                for (int s = 0; s < VPL; ++s) {
                    tl[k][s] = ldg_v4(pTl + s * 128); tr[k][s] = ldg_v4(pTh + s * 128);
                    bl[k][s] = ldg_v4(pBl + s * 128); br[k][s] = ldg_v4(pBh + s * 128);
                }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int j = j0 + k;
                const float fx = __shfl_sync(0xffffffffu, mine.frac, min(j, jn - 1));
                if (j < jn) {
                    const u64 wx = pack2(fx, fx);
                    if (pk[k] & (1 << 30)) {                     // inside (warp-uniform: the flag came from lane j)
#pragma unroll
                        for (int s = 0; s < VPL; ++s) {
                            const ulonglong2 top = lerp4_rn(tl[k][s], tr[k][s], wx, nz);        // crop_and_resize.c:102
                            const ulonglong2 bot = lerp4_rn(bl[k][s], br[k][s], wx, nz);        // :103-104
                            const ulonglong2 v = lerp4_rn(top, bot, wy, nz);                    // :106
                            st_stream_v4(o + s * 128, v);
                            if (o2) st_stream_v4(o2 + s * 128, v);
                        }
                    } else {
                        const ulonglong2 v = make_ulonglong2(ex2, ex2);
#pragma unroll
                        for (int s = 0; s < VPL; ++s) {
                            st_stream_v4(o + s * 128, v);
                            if (o2) st_stream_v4(o2 + s * 128, v);
                        }
                    }
                    o += C;
                    if (o2) o2 += C;
                }
            }
        }
    }
}

// local unit index of a set -> (r, i, slab).  S.slabs = C / (128 * VPL); units < 2^31.
template <int VPL, int U>
__device__ __forceinline__ void fwd_unit_lean(const FwdSet &S, unsigned u, int lane, u64 nz) {
    unsigned q = u;
    int slab = 0;
    if (S.slabs > 1) {
        q = u / (unsigned)S.slabs;
        slab = (int)(u - q * (unsigned)S.slabs);
    }
    const int r = (int)(q / (unsigned)S.ph);
    fwd_unit_lean<VPL, U>(S, r, (int)(q - (unsigned)r * (unsigned)S.ph), slab, lane, nz);
}

}  // namespace fi
