/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the Feature Intertwiner hot path.
 *
 * A plain-C restatement of the reference's arithmetic, written from the semantics in
 * SURVEY.md Appendix A.  Nothing in the product package may link, import or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Pinning: the reference ships no tests or golden vectors.  This oracle is
 * pinned instead against the reference's OWN sources compiled unmodified into oracle/_ref/
 * (crop_and_resize.c, nms.c; see oracle/Makefile) and against lib/OT_module.py imported as-is
 * in the build container (tests/golden/make_golden.py) -- tests/test_oracle_pins.py.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off  (no FMA contraction: tap indices must come out of
 * un-fused fp32 mul/add exactly as the reference's -std=c99 build produces them).
 *
 * All citations are relative to /root/reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FI_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * crop_and_resize sampling geometry
 * follows lib/roi_align/src/crop_and_resize.c:44-96 (== cuda/crop_and_resize_kernel.cu:40-70)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int lo, hi;   /* floor / ceil pixel index          */
    float frac;   /* lerp weight of the `hi` pixel     */
    int inside;   /* 0 -> extrapolation_value          */
} fi_axis_tap;

/* spacing between consecutive samples along one axis (crop_and_resize.c:44-50) */
static float axis_step(float c1, float c2, int extent, int crop) {
    if (crop > 1) return (c2 - c1) * (extent - 1) / (crop - 1);
    return 0;
}

/* the k-th sample along one axis (crop_and_resize.c:54-56,58,72-74) */
static fi_axis_tap axis_sample(float c1, float c2, float step, int k, int extent, int crop) {
    fi_axis_tap t;
    float pos;
    if (crop > 1) pos = c1 * (extent - 1) + k * step;
    else pos = 0.5 * (c1 + c2) * (extent - 1);   /* double arithmetic, narrowed on assignment */
    t.inside = !(pos < 0 || pos > extent - 1);
    t.lo = (int)floorf(pos);
    t.hi = (int)ceilf(pos);
    t.frac = pos - t.lo;
    return t;
}

/* Forward: image[B,C,H,W] (NCHW), boxes[R,4]=(y1,x1,y2,x2) normalised, box_ind[R] ->
 * crops[R,C,ph,pw].  crop_and_resize.c:6-112.  A box whose box_ind is out of range leaves its
 * crop at 0 (the GPU behaviour, crop_and_resize_kernel.cu:34-38; the CPU file aborts instead,
 * crop_and_resize.c:39-42) and is counted in the return value. */
FI_EXPORT int fi_oracle_crop_and_resize_fwd(const float *image, int B, int C, int H, int W,
                                            const float *boxes, const int *box_ind, int R,
                                            int ph, int pw, float extrapolation, float *crops) {
    int bad = 0;
    const size_t plane = (size_t)H * W;
    memset(crops, 0, sizeof(float) * (size_t)R * C * ph * pw);
    for (int r = 0; r < R; ++r) {
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1];
        const float y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const int b = box_ind[r];
        if (b < 0 || b >= B) { ++bad; continue; }
        const float sy = axis_step(y1, y2, H, ph);
        const float sx = axis_step(x1, x2, W, pw);
        float *out = crops + (size_t)r * C * ph * pw;
        const float *img = image + (size_t)b * C * plane;
        for (int i = 0; i < ph; ++i) {
            const fi_axis_tap ty = axis_sample(y1, y2, sy, i, H, ph);
            for (int j = 0; j < pw; ++j) {
                const fi_axis_tap tx = axis_sample(x1, x2, sx, j, W, pw);
                for (int c = 0; c < C; ++c) {
                    float v = extrapolation;
                    if (ty.inside && tx.inside) {
                        const float *p = img + (size_t)c * plane;
                        const float tl = p[(size_t)ty.lo * W + tx.lo], tr = p[(size_t)ty.lo * W + tx.hi];
                        const float bl = p[(size_t)ty.hi * W + tx.lo], br = p[(size_t)ty.hi * W + tx.hi];
                        const float top = tl + (tr - tl) * tx.frac;          /* crop_and_resize.c:102 */
                        const float bot = bl + (br - bl) * tx.frac;          /* :103-104 */
                        v = top + (bot - top) * ty.frac;                     /* :106 */
                    }
                    out[((size_t)c * ph + i) * pw + j] = v;
                }
            }
        }
    }
    return bad;
}

/* Backward: grads[R,C,ph,pw] -> grad_image[B,C,H,W] (zeroed here, crop_and_resize.c:183).
 * crop_and_resize.c:157-252: serial scatter in (box, y, x, channel) order. */
FI_EXPORT int fi_oracle_crop_and_resize_bwd(const float *grads, int B, int C, int H, int W,
                                            const float *boxes, const int *box_ind, int R,
                                            int ph, int pw, float *grad_image) {
    int bad = 0;
    const size_t plane = (size_t)H * W;
    memset(grad_image, 0, sizeof(float) * (size_t)B * C * plane);
    for (int r = 0; r < R; ++r) {
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1];
        const float y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const int b = box_ind[r];
        if (b < 0 || b >= B) { ++bad; continue; }
        const float sy = axis_step(y1, y2, H, ph);
        const float sx = axis_step(x1, x2, W, pw);
        const float *g = grads + (size_t)r * C * ph * pw;
        float *img = grad_image + (size_t)b * C * plane;
        for (int i = 0; i < ph; ++i) {
            const fi_axis_tap ty = axis_sample(y1, y2, sy, i, H, ph);
            if (!ty.inside) continue;
            for (int j = 0; j < pw; ++j) {
                const fi_axis_tap tx = axis_sample(x1, x2, sx, j, W, pw);
                if (!tx.inside) continue;
                for (int c = 0; c < C; ++c) {
                    float *p = img + (size_t)c * plane;
                    const float gv = g[((size_t)c * ph + i) * pw + j];
                    const float dtop = (1 - ty.frac) * gv;                   /* :241 */
                    p[(size_t)ty.lo * W + tx.lo] += (1 - tx.frac) * dtop;    /* :242 */
                    p[(size_t)ty.lo * W + tx.hi] += tx.frac * dtop;          /* :243 */
                    const float dbot = ty.frac * gv;                         /* :245 */
                    p[(size_t)ty.hi * W + tx.lo] += (1 - tx.frac) * dbot;    /* :246 */
                    p[(size_t)ty.hi * W + tx.hi] += tx.frac * dbot;          /* :247 */
                }
            }
        }
    }
    return bad;
}

/* Integer taps of every sample (the "RoI indices" that must match bit-exactly):
 * taps[R,ph,pw,5] = (y_lo, y_hi, x_lo, x_hi, inside).  Same geometry as above. */
FI_EXPORT void fi_oracle_crop_taps(int H, int W, const float *boxes, int R, int ph, int pw, int *taps) {
    for (int r = 0; r < R; ++r) {
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1];
        const float y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const float sy = axis_step(y1, y2, H, ph);
        const float sx = axis_step(x1, x2, W, pw);
        for (int i = 0; i < ph; ++i) {
            const fi_axis_tap ty = axis_sample(y1, y2, sy, i, H, ph);
            for (int j = 0; j < pw; ++j) {
                const fi_axis_tap tx = axis_sample(x1, x2, sx, j, W, pw);
                int *t = taps + (((size_t)r * ph + i) * pw + j) * 5;
                t[0] = ty.lo; t[1] = ty.hi; t[2] = tx.lo; t[3] = tx.hi;
                t[4] = ty.inside && tx.inside;
            }
        }
    }
}

/* Number of distinct (b,y,x) feature pixels read by a forward call (U in SURVEY.md 8(d)):
 * the algorithmic read volume is 4*C*U bytes.  `scratch` is B*H*W bytes. */
FI_EXPORT long fi_oracle_crop_unique_pixels(int B, int H, int W, const float *boxes, const int *box_ind,
                                            int R, int ph, int pw, unsigned char *scratch) {
    long u = 0;
    memset(scratch, 0, (size_t)B * H * W);
    for (int r = 0; r < R; ++r) {
        const float y1 = boxes[4 * r + 0], x1 = boxes[4 * r + 1];
        const float y2 = boxes[4 * r + 2], x2 = boxes[4 * r + 3];
        const int b = box_ind[r];
        if (b < 0 || b >= B) continue;
        const float sy = axis_step(y1, y2, H, ph);
        const float sx = axis_step(x1, x2, W, pw);
        unsigned char *m = scratch + (size_t)b * H * W;
        for (int i = 0; i < ph; ++i) {
            const fi_axis_tap ty = axis_sample(y1, y2, sy, i, H, ph);
            if (!ty.inside) continue;
            for (int j = 0; j < pw; ++j) {
                const fi_axis_tap tx = axis_sample(x1, x2, sx, j, W, pw);
                if (!tx.inside) continue;
                const size_t idx[4] = {(size_t)ty.lo * W + tx.lo, (size_t)ty.lo * W + tx.hi,
                                       (size_t)ty.hi * W + tx.lo, (size_t)ty.hi * W + tx.hi};
                for (int k = 0; k < 4; ++k) if (!m[idx[k]]) { m[idx[k]] = 1; ++u; }
            }
        }
    }
    return u;
}

/* ------------------------------------------------------------------------------------------
 * FPN level rule -- lib/sub_module.py:397-410 with log2 = tools/utils.py:50-55.
 * rois[n,4] normalised (y1,x1,y2,x2) -> level[n] in 2..5; pre_round[n] (optional) receives the
 * value before rounding so tests can tell a genuine mismatch from a .5 tie.
 * torch.round is round-half-to-even == rintf in the default rounding mode.
 * ------------------------------------------------------------------------------------------ */
FI_EXPORT void fi_oracle_roi_level(const float *rois, int n, float image_area, float base,
                                   int *level, float *pre_round) {
    const float ln2 = logf(2.0f);
    const float denom = base / sqrtf(image_area);
    for (int i = 0; i < n; ++i) {
        const float h = rois[4 * i + 2] - rois[4 * i + 0];
        const float w = rois[4 * i + 3] - rois[4 * i + 1];
        const float area = w * h;
        const float v = 4 + logf(sqrtf(area) / denom) / ln2;
        if (pre_round) pre_round[i] = v;
        const float rv = rintf(v);
        int l;
        /* float->int of -inf/NaN is what .int() does on the reference's device; zero-padded
         * RoIs give log(0) = -inf which must land on level 2. */
        if (!(rv >= 2.0f)) l = 2; else if (rv > 5.0f) l = 5; else l = (int)rv;
        level[i] = l;
    }
}

/* ------------------------------------------------------------------------------------------
 * Per-class segment mean -- lib/sub_module.py:664-684.
 * gt[k] int class ids, f[k,F] -> feat[F,ncls] (class-minor, like the reference), cnt[ncls].
 * Background (0) skipped.
 * ------------------------------------------------------------------------------------------ */
FI_EXPORT void fi_oracle_segment_mean(const int *gt, const float *f, int k, int F, int ncls,
                                      float *feat, float *cnt) {
    double *acc = (double *)calloc((size_t)F * ncls, sizeof(double));
    memset(cnt, 0, sizeof(float) * ncls);
    memset(feat, 0, sizeof(float) * (size_t)F * ncls);
    for (int i = 0; i < k; ++i) {
        const int c = gt[i];
        if (c <= 0 || c >= ncls) continue;
        cnt[c] += 1.0f;
        for (int j = 0; j < F; ++j) acc[(size_t)j * ncls + c] += f[(size_t)i * F + j];
    }
    for (int c = 1; c < ncls; ++c)
        if (cnt[c] > 0)
            for (int j = 0; j < F; ++j) feat[(size_t)j * ncls + c] = (float)(acc[(size_t)j * ncls + c] / cnt[c]);
    free(acc);
}

/* ------------------------------------------------------------------------------------------
 * Sinkhorn -- lib/OT_module.py:104-135 (cosine cost).  One problem: x[N,D], y[N,D].
 * Returns <P,C>.  P_out (N*N, optional) receives the transport plan, gx/gy (N*D, optional) the
 * gradient of the loss w.r.t. the UN-normalised x, y with P treated as a constant
 * (no_bp_P_L=True, OT_module.py:130-131) and the normalisation done out of place.
 * `wide`: 0 = fp32 accumulation in index order, 1 = double accumulation (the "exact" value
 * against which both the reference's and the kernel's rounding are judged).
 * ------------------------------------------------------------------------------------------ */
FI_EXPORT double fi_oracle_sinkhorn(const float *x, const float *y, int N, int D, float inv_eps, int L,
                                    int wide, float *P_out, float *gx, float *gy) {
    const float EPS = 1e-20f;
    float *xh = (float *)malloc(sizeof(float) * N * D), *yh = (float *)malloc(sizeof(float) * N * D);
    float *nx = (float *)malloc(sizeof(float) * N), *ny = (float *)malloc(sizeof(float) * N);
    float *Cm = (float *)malloc(sizeof(float) * N * N), *K = (float *)malloc(sizeof(float) * N * N);
    float *a = (float *)malloc(sizeof(float) * N), *b = (float *)malloc(sizeof(float) * N);
    for (int i = 0; i < N; ++i) {                                   /* OT_module.py:111-112 */
        double sx = 0, sy = 0;
        float fx = 0, fy = 0;
        for (int d = 0; d < D; ++d) {
            const float u = x[i * D + d], v = y[i * D + d];
            sx += (double)u * u; sy += (double)v * v; fx += u * u; fy += v * v;
        }
        nx[i] = wide ? (float)sqrt(sx) : sqrtf(fx);
        ny[i] = wide ? (float)sqrt(sy) : sqrtf(fy);
        for (int d = 0; d < D; ++d) { xh[i * D + d] = x[i * D + d] / (nx[i] + EPS); yh[i * D + d] = y[i * D + d] / (ny[i] + EPS); }
    }
    for (int i = 0; i < N; ++i)                                     /* :113, :116 */
        for (int j = 0; j < N; ++j) {
            double s = 0; float fs = 0;
            for (int d = 0; d < D; ++d) { s += (double)xh[i * D + d] * yh[j * D + d]; fs += xh[i * D + d] * yh[j * D + d]; }
            Cm[i * N + j] = 1 - (wide ? (float)s : fs);
            K[i * N + j] = expf(-inv_eps * Cm[i * N + j]);
        }
    const float c0 = 1.0f / N;                                      /* :118-119 */
    for (int j = 0; j < N; ++j) b[j] = c0;
    for (int i = 0; i < N; ++i) a[i] = c0;
    for (int it = 0; it < L; ++it) {                                /* :120-122 */
        for (int i = 0; i < N; ++i) {
            double s = 0; float fs = 0;
            for (int j = 0; j < N; ++j) { s += (double)K[i * N + j] * b[j]; fs += K[i * N + j] * b[j]; }
            a[i] = c0 / ((wide ? (float)s : fs) + EPS);
        }
        for (int j = 0; j < N; ++j) {
            double s = 0; float fs = 0;
            for (int i = 0; i < N; ++i) { s += (double)K[i * N + j] * a[i]; fs += K[i * N + j] * a[i]; }
            b[j] = c0 / ((wide ? (float)s : fs) + EPS);
        }
    }
    double loss = 0; float floss = 0;                               /* :129-134 */
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            const float p = a[i] * K[i * N + j] * b[j];
            if (P_out) P_out[i * N + j] = p;
            loss += (double)p * Cm[i * N + j]; floss += p * Cm[i * N + j];
        }
    if (gx || gy) {
        /* dL/dC = P  =>  dL/dxh_i = -sum_j P_ij yh_j ; dL/dyh_j = -sum_i P_ij xh_i ;
         * xh = x / (|x| + EPS)  =>  dL/dx = (g - xh * <xh,g> * |x|/(|x|+EPS)) / (|x|+EPS) */
        double *g = (double *)malloc(sizeof(double) * D);
        for (int pass = 0; pass < 2; ++pass) {
            float *out = pass == 0 ? gx : gy;
            if (!out) continue;
            const float *self = pass == 0 ? xh : yh, *other = pass == 0 ? yh : xh, *nrm = pass == 0 ? nx : ny;
            for (int i = 0; i < N; ++i) {
                for (int d = 0; d < D; ++d) g[d] = 0;
                for (int j = 0; j < N; ++j) {
                    const double p = pass == 0 ? (double)a[i] * K[i * N + j] * b[j] : (double)a[j] * K[j * N + i] * b[i];
                    for (int d = 0; d < D; ++d) g[d] -= p * other[j * D + d];
                }
                double dot = 0;
                for (int d = 0; d < D; ++d) dot += g[d] * self[i * D + d];
                const double den = (double)nrm[i] + EPS;
                for (int d = 0; d < D; ++d)
                    out[i * D + d] = (float)((g[d] - self[i * D + d] * dot * (nrm[i] / den)) / den);
            }
        }
        free(g);
    }
    free(xh); free(yh); free(nx); free(ny); free(Cm); free(K); free(a); free(b);
    return wide ? loss : (double)floss;
}

/* ------------------------------------------------------------------------------------------
 * NMS.  boxes[n,5] = (x1,y1,x2,y2,score) already sorted by descending score, +1 pixel
 * convention.  `strict`: 1 -> suppress when IoU >  thr (GPU, lib/nms/src/cuda/nms_kernel.cu:63
 *                                    + host reduce lib/nms/src/nms_cuda.c:43-58)
 *                        0 -> suppress when IoU >= thr (CPU, lib/nms/src/nms.c:59)
 * IoU is evaluated in the operand order each variant uses.  Returns the keep count.
 * ------------------------------------------------------------------------------------------ */
static float iou_gpu_order(const float *a, const float *b) {     /* nms_kernel.cu:16-24 */
    const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    const float w = fmaxf(right - left + 1, 0.f), h = fmaxf(bottom - top + 1, 0.f);
    const float inter = w * h;
    const float sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
    const float sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
    return inter / (sa + sb - inter);
}

FI_EXPORT int fi_oracle_nms(const float *boxes, int n, float thr, int strict, int *keep) {
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[kept++] = i;
        const float *bi = boxes + 5 * i;
        const float iarea = (bi[2] - bi[0] + 1) * (bi[3] - bi[1] + 1);
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            const float *bj = boxes + 5 * j;
            if (strict) {
                if (iou_gpu_order(bi, bj) > thr) dead[j] = 1;
            } else {                                               /* nms.c:49-60 */
                const float xx1 = fmaxf(bi[0], bj[0]), yy1 = fmaxf(bi[1], bj[1]);
                const float xx2 = fminf(bi[2], bj[2]), yy2 = fminf(bi[3], bj[3]);
                const float w = fmaxf(0.0, xx2 - xx1 + 1), h = fmaxf(0.0, yy2 - yy1 + 1);
                const float inter = w * h;
                const float jarea = (bj[2] - bj[0] + 1) * (bj[3] - bj[1] + 1);
                if (inter / (iarea + jarea - inter) >= thr) dead[j] = 1;
            }
        }
    }
    free(dead);
    return kept;
}

/* ------------------------------------------------------------------------------------------
 * RoIPool -- lib/roi_pooling/src/roi_pooling_kernel.cu:24-93 (forward) and :128-203 (backward,
 * expressed as the equivalent argmax scatter: every pooled cell adds its gradient to the input
 * element its argmax names).  features[B,C,H,W] NCHW, rois[R,5]=(b,x1,y1,x2,y2) in pixels.
 * round() is C round (half away from zero), as in the kernel.
 * ------------------------------------------------------------------------------------------ */
FI_EXPORT void fi_oracle_roi_pool_fwd(const float *feat, int B, int C, int H, int W, const float *rois, int R,
                                      int ph, int pw, float scale, float *top, int *argmax) {
    (void)B;
    for (int n = 0; n < R; ++n) {
        const float *roi = rois + 5 * n;
        const int b = (int)roi[0];
        const int sw = (int)round(roi[1] * scale), sh = (int)round(roi[2] * scale);
        const int ew = (int)round(roi[3] * scale), eh = (int)round(roi[4] * scale);
        const int rw = (int)fmaxf(ew - sw + 1, 1), rh = (int)fmaxf(eh - sh + 1, 1);
        const float bh = (float)rh / (float)ph, bw = (float)rw / (float)pw;
        for (int c = 0; c < C; ++c) {
            const size_t base = ((size_t)b * C + c) * H * W;
            for (int i = 0; i < ph; ++i)
                for (int j = 0; j < pw; ++j) {
                    int h0 = (int)floor((float)i * bh), w0 = (int)floor((float)j * bw);
                    int h1 = (int)ceil((float)(i + 1) * bh), w1 = (int)ceil((float)(j + 1) * bw);
                    h0 = (int)fminf(fmaxf(h0 + sh, 0), H); h1 = (int)fminf(fmaxf(h1 + sh, 0), H);
                    w0 = (int)fminf(fmaxf(w0 + sw, 0), W); w1 = (int)fminf(fmaxf(w1 + sw, 0), W);
                    const int empty = (h1 <= h0) || (w1 <= w0);
                    float best = empty ? 0 : -3.402823466e+38F;
                    int where = -1;
                    for (int h = h0; h < h1; ++h)
                        for (int w = w0; w < w1; ++w) {
                            const float v = feat[base + (size_t)h * W + w];
                            if (v > best) { best = v; where = (int)(base + (size_t)h * W + w); }
                        }
                    const size_t o = (((size_t)n * C + c) * ph + i) * pw + j;
                    top[o] = best;
                    if (argmax) argmax[o] = where;
                }
        }
    }
}

/* Backward.  The reference walks every INPUT element and every RoI (roi_pooling_kernel.cu:147-200)
 * and adds top_diff where argmax == that element, but only if the element also passes the kernel's
 * own feasibility tests: same image (:153-156), inside the rounded RoI rectangle (:164-168 -- a
 * malformed RoI with end < start therefore back-propagates nothing) and the pooled cell inside
 * [phstart,phend) x [pwstart,pwend) (:185-193).  Written here as the equivalent scatter over
 * pooled cells with those tests applied to the argmax element. */
FI_EXPORT void fi_oracle_roi_pool_bwd(const float *top_diff, const int *argmax, int B, int C, int H, int W,
                                      const float *rois, int R, int ph, int pw, float scale,
                                      float *bottom_diff) {
    memset(bottom_diff, 0, sizeof(float) * (size_t)B * C * H * W);
    for (int n = 0; n < R; ++n) {
        const float *roi = rois + 5 * n;
        const int b = (int)roi[0];
        const int sw = (int)round(roi[1] * scale), sh = (int)round(roi[2] * scale);
        const int ew = (int)round(roi[3] * scale), eh = (int)round(roi[4] * scale);
        const int rw = (int)fmaxf(ew - sw + 1, 1), rh = (int)fmaxf(eh - sh + 1, 1);
        const float bh = (float)rh / (float)ph, bw = (float)rw / (float)pw;
        for (int c = 0; c < C; ++c)
            for (int i = 0; i < ph; ++i)
                for (int j = 0; j < pw; ++j) {
                    const size_t o = (((size_t)n * C + c) * ph + i) * pw + j;
                    const int idx = argmax[o];
                    if (idx < 0) continue;
                    const int w = idx % W, h = (idx / W) % H, cc = (idx / (W * H)) % C, nn = idx / (W * H * C);
                    if (nn != b || cc != c) continue;
                    if (!(w >= sw && w <= ew && h >= sh && h <= eh)) continue;
                    int p0 = (int)floor((float)(h - sh) / bh), p1 = (int)ceil((float)(h - sh + 1) / bh);
                    int q0 = (int)floor((float)(w - sw) / bw), q1 = (int)ceil((float)(w - sw + 1) / bw);
                    p0 = (int)fminf(fmaxf(p0, 0), ph); p1 = (int)fminf(fmaxf(p1, 0), ph);
                    q0 = (int)fminf(fmaxf(q0, 0), pw); q1 = (int)fminf(fmaxf(q1, 0), pw);
                    if (i < p0 || i >= p1 || j < q0 || j >= q1) continue;
                    bottom_diff[idx] += top_diff[o];
                }
    }
}
