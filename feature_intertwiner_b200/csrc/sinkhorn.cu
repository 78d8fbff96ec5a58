// Batched Sinkhorn optimal-transport loss for B200 / sm_100a.
//
// Replaces the inner loop of lib/OT_module.py:104-135, which per problem launches ~4L+10 tiny cuBLAS /
// pointwise kernels and streams the N x N kernel matrix K through L2/HBM 2L+4 times.  Here one thread-block
// CLUSTER owns one problem and K never leaves the chip:
//
//   * rows of K are split across the CTAs of the cluster (1 CTA when K fits one SM's shared memory,
//     a CTA pair for N = 256 where K alone is 256 KB); every CTA keeps all N columns of its rows;
//   * a-update (K b) is local; b-update (K^T a) produces per-CTA partial column sums that are exchanged
//     through distributed shared memory, one cluster barrier per iteration (double-buffered partials);
//   * the loss <P,C> and, because P is a constant for autograd (no_bp_P_L=True, OT_module.py:130-131),
//     the analytic gradients dL/dx, dL/dy are produced by the same launch.
//
// The bound is shared-memory bandwidth (2 N^2 words per iteration), not HBM and not tensor cores: the
// iteration is a pair of mat-VECs.  The cost "GEMM" x^ y^T (dense only when D is large: N=64, D>=256 in the
// FPN-level loss) is done in fp32 FMA with 64x64 register tiles -- TF32/bf16 tensor cores would break the
// 1e-4 loss tolerance through the three-term cancellation 2W(x,y)-W(x,x)-W(y,y).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "fi_common.cuh"

namespace cg = cooperative_groups;

namespace fi {

constexpr int kSinkThreads = 256;
constexpr int kTile = 64;     // C/K tile edge
constexpr int kDT = 16;       // feature-dimension chunk staged per step
constexpr float kEps = 1e-20f;

struct SinkParams {
    const float *x, *y;        // [P, N, D]
    float *loss;               // [P]
    float *gx, *gy;            // [P, N, D] or null
    int N, D, L;
    float inv_eps;
    int NR;                    // rows of K per CTA
    int pitch;                 // row pitch of K/C in shared memory (floats)
    int TPR;                   // threads cooperating on one row in the a-update
    int TPC;                   // threads cooperating on one column in the b-update
    int keepC;                 // C kept in shared memory next to K (else recomputed for the loss)
    int masked;                // only problems whose loss slot holds NaN are solved (the others were done by the D = 1 class kernel)
    float *pbuf;               // null, or [P, N*N + 2N]: the plan P and the row norms go here and the gradient is left to
                               // sinkhorn_grad_raw_kernel / sinkhorn_grad_chain_kernel (large D: one CTA per problem cannot feed it)
    const float *nbuf;         // null, or the same [P, N*N + 2N] block with the row norms ALREADY in its tail (sinkhorn_norm_kernel)
    const float *gram;         // null, or [P, S, N, N]: x^ y^T summed over S slices of D by sinkhorn_gram_split_kernel
    int S;
};

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }

// Shared-memory carve-up (floats), identical on host and device.
struct SinkSmem {
    int K, C, a, b, part, colp, denx, deny, xs, ys, red, total;
    __host__ __device__ SinkSmem(int N, int NR, int pitch, int TPC, int keepC) {
        int o = 0;
        K = o; o += NR * pitch;
        C = o; o += keepC ? NR * pitch : 0;
        o = round4(o);
        a = o; o += round4(NR);
        b = o; o += round4(N);
        part = o; o += 2 * round4(N);
        colp = o; o += TPC * round4(N);
        denx = o; o += round4(NR);
        deny = o; o += round4(N);
        xs = o; o += kDT * kTile;
        ys = o; o += kDT * kTile;
        red = o; o += 32;
        total = o;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// One 64x64 tile of  S = x^_rows . y^_cols^T  accumulated over D in kDT chunks (fp32 FMA).
// x rows are this CTA's local rows; operands are normalised on load with a true division, exactly as the
// reference normalises before its mm (OT_module.py:111-113).
__device__ __forceinline__ void gram_tile(const float *__restrict__ xp, const float *__restrict__ yp, int N, int D, int row0g, int nrows,
                                          int ti, int tj, const float *denx, const float *deny, float *xs, float *ys, float acc[4][4],
                                          int d_begin = 0, int d_end = -1) {
    if (d_end < 0) d_end = D;
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
    // software pipeline: the global loads of chunk k+1 are in flight while chunk k is multiplied (one CTA walks its share of D
    // alone -- without this every chunk paid a full DRAM round trip before its FMAs could start)
    constexpr int kPer = (kTile * kDT) / kSinkThreads;
    float xr[kPer], yr[kPer];
    auto fetch = [&](int d0) {
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int e = t + k * kSinkThreads;
            const int dd = e % kDT, r = e / kDT;
            const int d = d0 + dd;
            const int il = ti * kTile + r;          // local x row
            const int j = tj * kTile + r;           // y row (= column of K)
            xr[k] = 0.f; yr[k] = 0.f;
            if (d < d_end) {
                if (il < nrows) xr[k] = __ldg(xp + (long)(row0g + il) * D + d);
                if (j < N) yr[k] = __ldg(yp + (long)j * D + d);
            }
        }
    };
    if (d_begin < d_end) fetch(d_begin);
    for (int d0 = d_begin; d0 < d_end; d0 += kDT) {
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int e = t + k * kSinkThreads;
            const int dd = e % kDT, r = e / kDT;
            const int il = ti * kTile + r, j = tj * kTile + r;
            xs[dd * kTile + r] = il < nrows ? __fdiv_rn(xr[k], denx[il]) : 0.f;
            ys[dd * kTile + r] = j < N ? __fdiv_rn(yr[k], deny[j]) : 0.f;
        }
        __syncthreads();
        if (d0 + kDT < d_end) fetch(d0 + kDT);
#pragma unroll
        for (int dd = 0; dd < kDT; ++dd) {
            const float4 a4 = *reinterpret_cast<const float4 *>(xs + dd * kTile + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4 *>(ys + dd * kTile + tx * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
        __syncthreads();
    }
}

template <int CS>
__global__ void __launch_bounds__(kSinkThreads) sinkhorn_kernel(const SinkParams p) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = CS > 1 ? (int)cluster.block_rank() : 0;
    const int prob = blockIdx.x / CS;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int N = p.N, D = p.D, NR = p.NR, pitch = p.pitch;
    const SinkSmem L(N, NR, pitch, p.TPC, p.keepC);
    float *Ks = smem + L.K, *Cs = smem + L.C, *a_s = smem + L.a, *b_s = smem + L.b, *part = smem + L.part;
    float *colp = smem + L.colp, *denx = smem + L.denx, *deny = smem + L.deny, *xs = smem + L.xs, *ys = smem + L.ys;
    float *red = smem + L.red;
    const int Np = round4(N);

    if (p.masked) {                                               // the whole cluster reads the same flag: all or none return
        const float flag = p.loss[prob];
        if (!isnan(flag)) return;
        if (CS > 1) {
            cluster.sync();                                       // every CTA has seen the flag before it is cleared
            if (rank == 0 && t == 0) p.loss[prob] = 0.f;          // the CTAs add their parts onto it at the end
        }
    }
    const int row0g = rank * NR;                                  // first global row owned by this CTA
    const int nrows = max(0, min(NR, N - row0g));                 // valid local rows
    const float *xp = p.x + (long)prob * N * D;
    const float *yp = p.y + (long)prob * N * D;

    // ---- phase 0: row norms (warp per row; OT_module.py:111-112), stored as (norm + EPS) -------------
    if (p.nbuf != nullptr) {                                      // large D (CS == 1): done by sinkhorn_norm_kernel, same bits
        const float *w = p.nbuf + (long)prob * ((long)N * N + 2 * N) + (long)N * N;
        for (int e = t; e < N; e += kSinkThreads) { denx[e] = w[e]; deny[e] = w[N + e]; }
    } else
    for (int q = wid; q < nrows + N; q += kSinkThreads / 32) {
        const float *src = q < nrows ? xp + (long)(row0g + q) * D : yp + (long)(q - nrows) * D;
        float s = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = __ldg(src + d); s = fmaf(v, v, s); }
        s = warp_sum(s);
        if (lane == 0) {
            const float den = __fadd_rn(sqrtf(s), kEps);
            if (q < nrows) denx[q] = den; else deny[q - nrows] = den;
        }
    }
    __syncthreads();

    // ---- phase 1: C = 1 - x^ y^T,  K = exp(-C / eps)   (OT_module.py:113,116) ------------------------
    const int tx = t & 15, ty = t >> 4;
    if (p.gram != nullptr) {                                      // the D-slices of sinkhorn_gram_split_kernel, added in slice order
        const float *g = p.gram + (long)prob * p.S * N * N;
        for (int e = t; e < N * N; e += kSinkThreads) {
            float acc = g[e];
            for (int sl = 1; sl < p.S; ++sl) acc = __fadd_rn(acc, g[(long)sl * N * N + e]);
            const int il = e / N, j = e - il * N;
            const float c = __fsub_rn(1.f, acc);
            Ks[il * pitch + j] = expf(-p.inv_eps * c);
            Cs[il * pitch + j] = c;                                // keepC is a precondition of this form (host)
        }
    } else
    for (int ti = 0; ti * kTile < nrows; ++ti)
        for (int tj = 0; tj * kTile < N; ++tj) {
            float acc[4][4];
            gram_tile(xp, yp, N, D, row0g, nrows, ti, tj, denx, deny, xs, ys, acc);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int il = ti * kTile + ty * 4 + u, j = tj * kTile + tx * 4 + v;
                    if (il < nrows && j < N) {
                        const float c = __fsub_rn(1.f, acc[u][v]);
                        Ks[il * pitch + j] = expf(-p.inv_eps * c);
                        if (p.keepC) Cs[il * pitch + j] = c;
                    }
                }
        }
    const float c0 = 1.0f / (float)N;                              // OT_module.py:118-119
    for (int j = t; j < N; j += kSinkThreads) b_s[j] = c0;
    __syncthreads();

    // ---- phase 2: L Sinkhorn iterations (OT_module.py:120-122) ---------------------------------------
    const int TPR = p.TPR, TPC = p.TPC;
    const int r_row = t / TPR, r_part = t - r_row * TPR;           // a-update role
    const int cthreads = kSinkThreads / TPC;
    const int c_col = t % cthreads, c_part = t / cthreads;         // b-update role
    for (int it = 0; it < p.L; ++it) {
        // a_i = c0 / (sum_j K_ij b_j + EPS), local rows
        {
            float s = 0.f;
            if (r_row < nrows) {
                const float *kr = Ks + r_row * pitch;
#pragma unroll 8
                for (int j = r_part; j < N; j += TPR) s = fmaf(kr[j], b_s[j], s);
            }
            for (int d = TPR >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            if (r_row < nrows && r_part == 0) a_s[r_row] = __fdiv_rn(c0, __fadd_rn(s, kEps));
        }
        __syncthreads();
        // partial_j = sum_{local i} K_ij a_i
        {
            float s = 0.f;
            if (c_col < N) {
                const float *kc = Ks + c_col;
#pragma unroll 8
                for (int il = c_part; il < nrows; il += TPC) s = fmaf(kc[il * pitch], a_s[il], s);
            }
            if (TPC > 1) {
                if (c_col < N) colp[c_part * Np + c_col] = s;
                __syncthreads();
                if (c_part == 0 && c_col < N) {
                    for (int q = 1; q < TPC; ++q) s += colp[q * Np + c_col];
                }
            }
            float *mine = part + (it & 1) * Np;
            if (CS > 1) {
                if (c_part == 0 && c_col < N) mine[c_col] = s;
                cluster.sync();                                    // partials of every CTA are visible
                if (c_part == 0 && c_col < N) {
                    float tot = 0.f;
#pragma unroll
                    for (int r = 0; r < CS; ++r) tot += cluster.map_shared_rank(mine, r)[c_col];   // rank order: same bits everywhere
                    b_s[c_col] = __fdiv_rn(c0, __fadd_rn(tot, kEps));
                }
            } else {
                if (c_part == 0 && c_col < N) b_s[c_col] = __fdiv_rn(c0, __fadd_rn(s, kEps));
            }
        }
        __syncthreads();
    }

    // ---- phase 3: P = a K b^T (in place over K), loss = <P, C>   (OT_module.py:129-134) -------------
    float lsum = 0.f;
    if (p.keepC) {
        if (r_row < nrows) {
            const float ai = a_s[r_row];
            for (int j = r_part; j < N; j += TPR) {
                const float pij = __fmul_rn(__fmul_rn(ai, Ks[r_row * pitch + j]), b_s[j]);
                Ks[r_row * pitch + j] = pij;
                lsum = fmaf(pij, Cs[r_row * pitch + j], lsum);
            }
        }
    } else {
        for (int ti = 0; ti * kTile < nrows; ++ti)
            for (int tj = 0; tj * kTile < N; ++tj) {
                float acc[4][4];
                gram_tile(xp, yp, N, D, row0g, nrows, ti, tj, denx, deny, xs, ys, acc);
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int il = ti * kTile + ty * 4 + u, j = tj * kTile + tx * 4 + v;
                        if (il < nrows && j < N) {
                            const float pij = __fmul_rn(__fmul_rn(a_s[il], Ks[il * pitch + j]), b_s[j]);
                            Ks[il * pitch + j] = pij;
                            lsum = fmaf(pij, __fsub_rn(1.f, acc[u][v]), lsum);
                        }
                    }
            }
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[wid] = lsum;
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
        for (int w = 0; w < kSinkThreads / 32; ++w) s += red[w];
        if (CS > 1) atomicAdd(p.loss + prob, s);                   // two addends: order-independent
        else p.loss[prob] = s;
    }

    // ---- phase 4: gradients with P constant ----------------------------------------------------------
    //   dL/dx^_i = -sum_j P_ij y^_j ,  dL/dy^_j = -sum_i P_ij x^_i ,
    //   x^ = x / (|x| + EPS)  =>  dL/dx = (g - x^ <x^,g> |x|/(|x|+EPS)) / (|x|+EPS)
    if (p.gx != nullptr && p.pbuf != nullptr) {              // CS == 1 here (host): the whole plan is in this CTA's shared memory
        __syncthreads();
        float *w = p.pbuf + (long)prob * ((long)N * N + 2 * N);
        for (int e = t; e < N * N; e += kSinkThreads) w[e] = Ks[(e / N) * pitch + (e % N)];
        for (int e = t; e < N; e += kSinkThreads) { w[N * N + e] = denx[e]; w[N * N + N + e] = deny[e]; }
    } else if (p.gx != nullptr) {
        float *gxp = p.gx + (long)prob * N * D, *gyp = p.gy + (long)prob * N * D;
        for (long e = t; e < (long)nrows * D; e += kSinkThreads) {
            const int il = (int)(e / D), d = (int)(e - (long)il * D);
            float s = 0.f;
            for (int j = 0; j < N; ++j) s = fmaf(Ks[il * pitch + j], __fdiv_rn(__ldg(yp + (long)j * D + d), deny[j]), s);
            gxp[(long)(row0g + il) * D + d] = -s;
        }
        for (long e = t; e < (long)N * D; e += kSinkThreads) {
            const int j = (int)(e / D), d = (int)(e - (long)j * D);
            float s = 0.f;
            for (int il = 0; il < nrows; ++il) s = fmaf(Ks[il * pitch + j], __fdiv_rn(__ldg(xp + (long)(row0g + il) * D + d), denx[il]), s);
            if (CS > 1) atomicAdd(gyp + e, -s);
            else gyp[e] = -s;
        }
        __threadfence();
        if (CS > 1) cluster.sync(); else __syncthreads();
        // chain through the row normalisation; this CTA finalises the rows it owns (x and y alike)
        for (int q = wid; q < 2 * nrows; q += kSinkThreads / 32) {
            const bool isx = q < nrows;
            const int il = isx ? q : q - nrows;
            const int i = row0g + il;
            const float *src = (isx ? xp : yp) + (long)i * D;
            float *g = (isx ? gxp : gyp) + (long)i * D;
            const float den = isx ? denx[il] : deny[i];
            float dot = 0.f;
            for (int d = lane; d < D; d += 32) dot = fmaf(__ldcg(g + d), __fdiv_rn(__ldg(src + d), den), dot);
            dot = warp_sum(dot);
            const float shrink = __fdiv_rn(__fsub_rn(den, kEps), den);   // |x| / (|x| + EPS)
            for (int d = lane; d < D; d += 32) {
                const float xh = __fdiv_rn(__ldg(src + d), den);
                g[d] = __fdiv_rn(__fsub_rn(__ldcg(g + d), __fmul_rn(__fmul_rn(xh, dot), shrink)), den);
            }
        }
    }
    if (CS > 1) cluster.sync();   // no CTA may exit while a peer can still read its shared memory
}

// ------------------------------------------------------------------------------------------------
// Cost matrix for large D (the FPN-level loss: N = 64, D = (s/4)^2 up to 4096, lib/sub_module.py:179-213).  Inside the solver
// kernel ONE CTA per problem walks all of D for its 64x64 tile: 24 CTAs on 148 SMs, 0.7 ms for 24 problems at D = 4096 where
// the 48 MB of operands are a 10 us read.  So for D >= 128 the two D-long passes leave the solver:
//   sinkhorn_norm_kernel        warp per row: |row| + EPS into the workspace (the solver's own summation order: same bits);
//   sinkhorn_gram_split_kernel  CTA = (problem, slice of D, 64x64 tile): the slice's share of x^ y^T (fp32 FMA, operands
//                               normalised on load as the reference does before its mm, OT_module.py:111-113);
// and the solver adds the S slices in order.  Deterministic; differs from the one-CTA form only in the summation order over D.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sinkhorn_norm_kernel(const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ nbuf, int P,
                                                           int N, int D) {
    const int lane = threadIdx.x & 31;
    const long row = blockIdx.x * 8L + (threadIdx.x >> 5);
    if (row >= 2L * P * N) return;
    const int prob = (int)(row / (2 * N)), q = (int)(row - (long)prob * 2 * N);
    const float *src = (q < N ? x : y) + ((long)prob * N + (q < N ? q : q - N)) * D;
    float s = 0.f;
#pragma unroll 8
    for (int d = lane; d < D; d += 32) { const float v = __ldg(src + d); s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) nbuf[(long)prob * ((long)N * N + 2 * N) + (long)N * N + q] = __fadd_rn(sqrtf(s), kEps);
}

__global__ void __launch_bounds__(kSinkThreads) sinkhorn_gram_split_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                                          const float *__restrict__ nbuf, float *__restrict__ gram, int N, int D,
                                                                          int S, int Dc) {
    __shared__ __align__(16) float xs[kDT * kTile], ys[kDT * kTile];
    const int prob = blockIdx.x, sl = blockIdx.y;
    const int tiles = (N + kTile - 1) / kTile, ti = blockIdx.z / tiles, tj = blockIdx.z % tiles;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const float *den = nbuf + (long)prob * ((long)N * N + 2 * N) + (long)N * N;
    float acc[4][4];
    const int d0 = sl * Dc, d1 = min(D, d0 + Dc);
    gram_tile(x + (long)prob * N * D, y + (long)prob * N * D, N, D, 0, N, ti, tj, den, den + N, xs, ys, acc, d0, d1);
    float *g = gram + ((long)prob * S + sl) * N * N;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int il = ti * kTile + ty * 4 + u, j = tj * kTile + tx * 4;
        if (il < N && j < N) *reinterpret_cast<float4 *>(g + (long)il * N + j) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);   // N % 4 == 0
    }
}

// ------------------------------------------------------------------------------------------------
// Gradient for large D (the FPN-level loss: N = 64, D = (s/4)^2 up to 4096, lib/sub_module.py:179-213).  Inside the solver
// kernel one CTA per problem walks N*D outputs with an N-term sum each -- 24 CTAs on 148 SMs, 28 ms for 24 problems at
// D = 4096.  The plan P (N x N) is tiny, so it is handed over through global memory and the two products
//     g~x = -P y^   [N x D],      g~y = -P^T x^   [N x D]
// are spread over (problem, 64-column chunk of D) CTAs: P in shared memory (broadcast reads, float4 along the summed index),
// thread = (column d, quarter of the rows), the same ascending summation order as the in-kernel form (identical bits).
// A second launch (warp per row) applies the chain through the row normalisation, which needs the full-row dot product.
// ------------------------------------------------------------------------------------------------
// The two products for one column d and PER rows / columns starting at r0 (compile-time PER: a run-time bound on the unrolled
// accumulators is predication, and predicated-off FFMAs still issue -- ncu: twice the useful FFMA count at N = 64).
template <int PER, bool RAGGED>
__device__ __forceinline__ void grad_raw_products(const float *Ps, const float *xt, const float *yt, float *gxp, float *gyp, int N, int D, int r0,
                                                  int c) {
    float acc[PER];
    // g~x[i][d] = -sum_j P[i][j] y^[j][d]
#pragma unroll
    for (int r = 0; r < PER; ++r) acc[r] = 0.f;
    for (int j = 0; j < N; j += 4) {
        float yv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) yv[u] = yt[(j + u) * 64 + c];
#pragma unroll
        for (int r = 0; r < PER; ++r) {
            if (!RAGGED || r0 + r < N) {
                const float4 pv = *reinterpret_cast<const float4 *>(Ps + (r0 + r) * N + j);
                acc[r] = fmaf(pv.x, yv[0], acc[r]); acc[r] = fmaf(pv.y, yv[1], acc[r]);
                acc[r] = fmaf(pv.z, yv[2], acc[r]); acc[r] = fmaf(pv.w, yv[3], acc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < PER; ++r)
        if (!RAGGED || r0 + r < N) gxp[(long)(r0 + r) * D] = -acc[r];
    // g~y[j][d] = -sum_i P[i][j] x^[i][d]
#pragma unroll
    for (int r = 0; r < PER; ++r) acc[r] = 0.f;
    for (int i = 0; i < N; ++i) {
        const float xv = xt[i * 64 + c];
#pragma unroll
        for (int c4 = 0; c4 < PER / 4; ++c4) {
            if (!RAGGED || r0 + 4 * c4 < N) {                              // N % 4 == 0: a float4 of columns is all in or all out
                const float4 pv = *reinterpret_cast<const float4 *>(Ps + i * N + r0 + 4 * c4);
                acc[4 * c4 + 0] = fmaf(pv.x, xv, acc[4 * c4 + 0]); acc[4 * c4 + 1] = fmaf(pv.y, xv, acc[4 * c4 + 1]);
                acc[4 * c4 + 2] = fmaf(pv.z, xv, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(pv.w, xv, acc[4 * c4 + 3]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < PER; ++r)
        if (!RAGGED || r0 + r < N) gyp[(long)(r0 + r) * D] = -acc[r];
}

template <int PER>
__global__ void __launch_bounds__(256) sinkhorn_grad_raw_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ pbuf,
                                                               float *__restrict__ gx, float *__restrict__ gy, int N, int D) {
    extern __shared__ __align__(16) float sm[];
    float *Ps = sm, *denx = sm + N * N, *deny = denx + N;
    float *xt = deny + N, *yt = xt + N * 64;                // the CTA's 64 columns of x^ and y^ (N x 64 each), normalised once
    const int prob = blockIdx.y, chunk = blockIdx.x, t = threadIdx.x;   // chunk fastest: co-running CTAs read neighbouring 256 B pieces of the same rows
    const float *w = pbuf + (long)prob * ((long)N * N + 2 * N);
    for (int e = t; e < N * N + 2 * N; e += 256) sm[e] = w[e];
    __syncthreads();
    {
        // all loads of the tile in flight at once: the loops below used to fetch one operand per iteration straight from
        // global memory -- 64 dependent DRAM round trips per CTA, 0.5 ms for 24 problems at D = 4096
        const float *xb = x + (long)prob * N * D + chunk * 64, *yb = y + (long)prob * N * D + chunk * 64;
        const int c = t & 63, dcol = chunk * 64 + c;
#pragma unroll 8
        for (int r = t >> 6; r < N; r += 4) {
            float xv = 0.f, yv = 0.f;
            if (dcol < D) { xv = __ldg(xb + (long)r * D + c); yv = __ldg(yb + (long)r * D + c); }
            xt[r * 64 + c] = __fdiv_rn(xv, denx[r]);
            yt[r * 64 + c] = __fdiv_rn(yv, deny[r]);
        }
    }
    __syncthreads();
    const int c = t & 63, d = chunk * 64 + c;
    const int r0 = (t >> 6) * PER;                          // quarter of the rows (gx) / columns (gy): PER = 4 * ceil(N / 16) each
    if (d >= D || r0 >= N) return;
    float *gxp = gx + (long)prob * N * D + d, *gyp = gy + (long)prob * N * D + d;
    if (r0 + PER <= N) grad_raw_products<PER, false>(Ps, xt, yt, gxp, gyp, N, D, r0, c);
    else grad_raw_products<PER, true>(Ps, xt, yt, gxp, gyp, N, D, r0, c);       // ragged last quarter (N % 16 != 0), warp-uniform
}

// warp per row (x rows then y rows of every problem): dL/dx = (g - x^ <x^, g> |x| / (|x| + EPS)) / (|x| + EPS)
__global__ void __launch_bounds__(256) sinkhorn_grad_chain_kernel(const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ pbuf,
                                                                 float *__restrict__ gx, float *__restrict__ gy, int P, int N, int D) {
    const int lane = threadIdx.x & 31;
    const long row = blockIdx.x * 8L + (threadIdx.x >> 5);
    if (row >= 2L * P * N) return;
    const int prob = (int)(row / (2 * N)), q = (int)(row - (long)prob * 2 * N);
    const bool isx = q < N;
    const int i = isx ? q : q - N;
    const float *src = (isx ? x : y) + ((long)prob * N + i) * D;
    float *g = (isx ? gx : gy) + ((long)prob * N + i) * D;
    const float den = pbuf[(long)prob * ((long)N * N + 2 * N) + (long)N * N + (isx ? 0 : N) + i];
    float dot = 0.f;
#pragma unroll 8
    for (int d = lane; d < D; d += 32) dot = fmaf(g[d], __fdiv_rn(__ldg(src + d), den), dot);
    dot = warp_sum(dot);
    const float shrink = __fdiv_rn(__fsub_rn(den, kEps), den);   // |x| / (|x| + EPS)
#pragma unroll 8
    for (int d = lane; d < D; d += 32) {
        const float xh = __fdiv_rn(__ldg(src + d), den);
        g[d] = __fdiv_rn(__fsub_rn(g[d], __fmul_rn(__fmul_rn(xh, dot), shrink)), den);
    }
}

// The same chain with one CTA per row and the row kept in registers (D <= 256 * kChainCache): one pass over memory, 16 loads in
// flight per thread.  The warp-per-row form above has 0.4 waves of warps each waiting on its own DRAM round trips (ncu: 50
// stalled-on-long-scoreboard warps per issue, 190 us for 24 problems at D = 4096).
constexpr int kChainCache = 16;
__global__ void __launch_bounds__(256) sinkhorn_grad_chain_rows_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                                      const float *__restrict__ pbuf, float *__restrict__ gx, float *__restrict__ gy,
                                                                      int P, int N, int D) {
    __shared__ float red[8];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const long row = blockIdx.x;
    const int prob = (int)(row / (2 * N)), q = (int)(row - (long)prob * 2 * N);
    const bool isx = q < N;
    const int i = isx ? q : q - N;
    const float *src = (isx ? x : y) + ((long)prob * N + i) * D;
    float *g = (isx ? gx : gy) + ((long)prob * N + i) * D;
    const float den = pbuf[(long)prob * ((long)N * N + 2 * N) + (long)N * N + (isx ? 0 : N) + i];
    float xh[kChainCache], gv[kChainCache];
#pragma unroll
    for (int k = 0; k < kChainCache; ++k) {
        const int d = t + k * 256;
        xh[k] = d < D ? __ldg(src + d) : 0.f;
        gv[k] = d < D ? g[d] : 0.f;
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < kChainCache; ++k) { xh[k] = __fdiv_rn(xh[k], den); dot = fmaf(gv[k], xh[k], dot); }
    dot = warp_sum(dot);
    if (lane == 0) red[wid] = dot;
    __syncthreads();
    dot = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) dot = __fadd_rn(dot, red[w]);
    const float shrink = __fdiv_rn(__fsub_rn(den, kEps), den);   // |x| / (|x| + EPS)
#pragma unroll
    for (int k = 0; k < kChainCache; ++k) {
        const int d = t + k * 256;
        if (d < D) g[d] = __fdiv_rn(__fsub_rn(gv[k], __fmul_rn(__fmul_rn(xh[k], dot), shrink)), den);
    }
}

// ------------------------------------------------------------------------------------------------
// N = 256, D = 1 -- the RoI-level intertwiner loss (critic output [n,256,1], lib/OT_module.py:95-101) --
// with K resident in REGISTERS.
//
// The generic kernel above streams K from shared memory twice per iteration and is bound by shared-memory
// bandwidth (2*N^2 words / 128 B/clk).  For this shape a CTA pair owns a problem, each CTA 128 rows, and
// every thread keeps a 4-row x 32-column tile of K (128 registers): an iteration is 256 FMAs per thread
// plus two small cross-lane reductions, and the only traffic is the b / partial-sum vectors:
//   lane = cg | rl << 3 : cg = column group (8), rl = row lane (4); warp w, row group rg = 4w + rl
//   rows    4 rg .. 4 rg + 3
//   columns 32 k + 4 cg + e, k = 0..7, e = 0..3 (a warp's 8 column groups read 128 contiguous bytes of b)
//   K b     per-thread partial over its 32 columns, xor-shuffle over the 3 cg bits
//   K^T a   per-thread partial over its 4 rows for 32 columns, reduce-scatter butterfly over the 2 rl
//           bits (24 shuffles, 8 column sums left per lane), 8 warps combined through shared memory, the
//           two CTAs through distributed shared memory -- one cluster barrier per iteration.
// ------------------------------------------------------------------------------------------------
struct Sink256Smem {
    float xh[128], yh[256], b[256], denx[128], deny[256];
    float colp[8][256];
    float part[2][256];
    float red[8];
};

// reduce 32 per-thread column partials over the row lanes of the warp and over the 8 warps; returns the CTA
// total of column `threadIdx.x`.  Ends with a __syncthreads-protected read, so `colp` may be reused afterwards
// only after the caller's next barrier.
__device__ __forceinline__ float cta_column_sum(float (&cp)[32], Sink256Smem &sm, int lane, int w, int cg) {
    const bool hi = (lane >> 3) & 1, hi2 = (lane >> 4) & 1;
    float q[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float keep = hi ? cp[16 + j] : cp[j], send = hi ? cp[j] : cp[16 + j];
        q[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = hi2 ? q[8 + j] : q[j], send = hi2 ? q[j] : q[8 + j];
        z[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    const int k0 = (hi ? 4 : 0) + (hi2 ? 2 : 0);          // the two float4 column chunks this lane now owns
    *reinterpret_cast<float4 *>(&sm.colp[w][32 * k0 + 4 * cg]) = make_float4(z[0], z[1], z[2], z[3]);
    *reinterpret_cast<float4 *>(&sm.colp[w][32 * (k0 + 1) + 4 * cg]) = make_float4(z[4], z[5], z[6], z[7]);
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int q8 = 0; q8 < 8; ++q8) tot += sm.colp[q8][threadIdx.x];
    return tot;
}

__global__ void __launch_bounds__(256, 1) sinkhorn_n256_d1_kernel(const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ loss,
                                                                 float *__restrict__ gx, float *__restrict__ gy, float inv_eps, int L, int masked) {
    __shared__ __align__(16) Sink256Smem sm;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int prob = blockIdx.x >> 1;
    if (masked) {                                          // solved by sinkhorn_d1_classes_kernel unless its loss slot holds NaN
        const float flag = loss[prob];
        if (!isnan(flag)) return;                          // both CTAs of the pair read the same flag
        cluster.sync();
        if (rank == 0 && threadIdx.x == 0) loss[prob] = 0.f;
    }
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int cgp = lane & 7, rl = lane >> 3;
    const int r0 = (w * 4 + rl) * 4;                       // first of this thread's 4 local rows
    const float *xp = x + (long)prob * 256, *yp = y + (long)prob * 256;
    const float c0 = 1.0f / 256.0f;

    // ---- normalise (OT_module.py:111-112); D == 1: |v| = sqrt(v*v)
    if (t < 128) {
        const float v = __ldg(xp + rank * 128 + t);
        const float den = __fadd_rn(sqrtf(__fmul_rn(v, v)), kEps);
        sm.denx[t] = den; sm.xh[t] = __fdiv_rn(v, den);
    }
    {
        const float v = __ldg(yp + t);
        const float den = __fadd_rn(sqrtf(__fmul_rn(v, v)), kEps);
        sm.deny[t] = den; sm.yh[t] = __fdiv_rn(v, den);
        sm.b[t] = c0;
    }
    __syncthreads();
    float xh[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xh[u] = sm.xh[r0 + u];

    // ---- K = exp(-(1 - xh yh) / eps) straight into registers (OT_module.py:113,116)
    float Kr[4][32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 y4 = *reinterpret_cast<const float4 *>(&sm.yh[32 * k + 4 * cgp]);
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int u = 0; u < 4; ++u) Kr[u][4 * k + e] = expf(-inv_eps * __fsub_rn(1.f, __fmul_rn(xh[u], yv[e])));
    }

    // ---- L Sinkhorn iterations (OT_module.py:120-122)
    float a[4] = {c0, c0, c0, c0};
    for (int it = 0; it < L; ++it) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 b4 = *reinterpret_cast<const float4 *>(&sm.b[32 * k + 4 * cgp]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s[u] = fmaf(Kr[u][4 * k + 0], b4.x, s[u]); s[u] = fmaf(Kr[u][4 * k + 1], b4.y, s[u]);
                s[u] = fmaf(Kr[u][4 * k + 2], b4.z, s[u]); s[u] = fmaf(Kr[u][4 * k + 3], b4.w, s[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s[u] += __shfl_xor_sync(0xffffffffu, s[u], 1);
            s[u] += __shfl_xor_sync(0xffffffffu, s[u], 2);
            s[u] += __shfl_xor_sync(0xffffffffu, s[u], 4);
            a[u] = __fdiv_rn(c0, __fadd_rn(s[u], kEps));
        }
        float cp[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) cp[j] = fmaf(Kr[3][j], a[3], fmaf(Kr[2][j], a[2], fmaf(Kr[1][j], a[1], Kr[0][j] * a[0])));
        const float mine = cta_column_sum(cp, sm, lane, w, cgp);
        float *slot = sm.part[it & 1];
        slot[t] = mine;
        cluster.sync();                                    // both CTAs' partial column sums are visible
        const float tot = cluster.map_shared_rank(slot, 0)[t] + cluster.map_shared_rank(slot, 1)[t];   // rank order: same bits on both CTAs
        sm.b[t] = __fdiv_rn(c0, __fadd_rn(tot, kEps));
        __syncthreads();
    }

    // ---- P = a K b^T, loss = <P, C>   (OT_module.py:129-134); K registers now hold P
    float lsum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 b4 = *reinterpret_cast<const float4 *>(&sm.b[32 * k + 4 * cgp]);
        const float4 y4 = *reinterpret_cast<const float4 *>(&sm.yh[32 * k + 4 * cgp]);
        const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float pij = __fmul_rn(__fmul_rn(a[u], Kr[u][4 * k + e]), bv[e]);
                Kr[u][4 * k + e] = pij;
                lsum = fmaf(pij, __fsub_rn(1.f, __fmul_rn(xh[u], yv[e])), lsum);
            }
    }
    lsum = warp_sum(lsum);
    if (lane == 0) sm.red[w] = lsum;
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += sm.red[q];
        atomicAdd(loss + prob, s);                         // two addends (one per CTA): order-independent
    }

    // ---- gradients with P constant (see the generic kernel); D == 1
    if (gx != nullptr) {
        float gxh[4] = {0.f, 0.f, 0.f, 0.f};
        float cp[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 y4 = *reinterpret_cast<const float4 *>(&sm.yh[32 * k + 4 * cgp]);
            const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float c = 0.f;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    gxh[u] = fmaf(Kr[u][4 * k + e], yv[e], gxh[u]);
                    c = fmaf(Kr[u][4 * k + e], xh[u], c);
                }
                cp[4 * k + e] = c;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            gxh[u] += __shfl_xor_sync(0xffffffffu, gxh[u], 1);
            gxh[u] += __shfl_xor_sync(0xffffffffu, gxh[u], 2);
            gxh[u] += __shfl_xor_sync(0xffffffffu, gxh[u], 4);
        }
        if (cgp == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float den = sm.denx[r0 + u], g = -gxh[u];
                const float shrink = __fdiv_rn(__fsub_rn(den, kEps), den);
                gx[(long)prob * 256 + rank * 128 + r0 + u] = __fdiv_rn(__fsub_rn(g, __fmul_rn(__fmul_rn(xh[u], __fmul_rn(g, xh[u])), shrink)), den);
            }
        }
        const float mine = cta_column_sum(cp, sm, lane, w, cgp);     // colp is free again: barrier at the end of the loss phase
        float *slot = sm.part[L & 1];
        cluster.sync();                                    // peer is past its last read of `part`
        slot[t] = mine;
        cluster.sync();
        if ((t >> 7) == rank) {                            // this CTA finalises the y rows it owns
            const float g = -(cluster.map_shared_rank(slot, 0)[t] + cluster.map_shared_rank(slot, 1)[t]);
            const float den = sm.deny[t], yh = sm.yh[t];
            const float shrink = __fdiv_rn(__fsub_rn(den, kEps), den);
            gy[(long)prob * 256 + t] = __fdiv_rn(__fsub_rn(g, __fmul_rn(__fmul_rn(yh, __fmul_rn(g, yh)), shrink)), den);
        }
    }
    cluster.sync();   // no CTA may exit while its peer can still read its shared memory
}

// ------------------------------------------------------------------------------------------------
// D = 1 by CLASSES.  With one feature per row the cosine normalisation x^ = x / (|x| + 1e-20) is a sign function: every
// entry with |x| >> 1e-20 becomes exactly -1, 0 or +1 in fp32 (the critic ends in a ReLU: 0 or 1).
// Then C_ij = 1 - x^_i y^_j and K_ij = exp(-C_ij / eps) take one value per (class of i, class of j), all rows of a class see
// the same sums, and the iteration a = c0 / (K b + EPS), b = c0 / (K^T a + EPS) closes on THREE a's and THREE b's weighted by
// the class sizes: O(L) work per problem instead of O(L N^2).  One warp per problem; lane l holds elements l, l + 32, ...
// A problem with any other value of x^ or y^ (|x| within a few orders of 1e-20) is left to the dense kernels: its loss slot
// is set to NaN, which those kernels take as their work mask.
// Values: the same K entries, the same update formulas; the sums over j run over classes instead of elements, i.e. another
// summation order of identical terms (differences ~1e-7 relative against the 1e-4 bar).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sinkhorn_d1_classes_kernel(const float *__restrict__ x, const float *__restrict__ y, float *__restrict__ loss,
                                                                 float *__restrict__ gx, float *__restrict__ gy, int P, int N, float inv_eps, int L) {
    const int lane = threadIdx.x & 31;
    const int prob = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (prob >= P) return;
    const float *xp = x + (long)prob * N, *yp = y + (long)prob * N;
    float xh[8], yh[8], dx[8], dy[8];
    int nx0 = 0, nx1 = 0, nx2 = 0, ny0 = 0, ny1 = 0, ny2 = 0;      // class sizes: x^ = -1, 0, +1
    bool clean = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane + 32 * k;
        xh[k] = 0.f; yh[k] = 0.f; dx[k] = 1.f; dy[k] = 1.f;
        if (i < N) {
            const float v = __ldg(xp + i), w = __ldg(yp + i);
            dx[k] = __fadd_rn(sqrtf(__fmul_rn(v, v)), kEps);     // OT_module.py:111-112 with D == 1
            dy[k] = __fadd_rn(sqrtf(__fmul_rn(w, w)), kEps);
            xh[k] = __fdiv_rn(v, dx[k]);
            yh[k] = __fdiv_rn(w, dy[k]);
            nx0 += xh[k] == -1.f; nx1 += xh[k] == 0.f; nx2 += xh[k] == 1.f;
            ny0 += yh[k] == -1.f; ny1 += yh[k] == 0.f; ny2 += yh[k] == 1.f;
            clean = clean && (xh[k] == -1.f || xh[k] == 0.f || xh[k] == 1.f) && (yh[k] == -1.f || yh[k] == 0.f || yh[k] == 1.f);
        }
    }
    clean = __all_sync(0xffffffffu, clean);
    if (!clean) {
        if (lane == 0) loss[prob] = __int_as_float(0x7fc00000);   // NaN: "still to do" for the dense kernels
        if (gy != nullptr)
            for (int i = lane; i < N; i += 32) gy[(long)prob * N + i] = 0.f;   // the CTA-pair kernels add their halves onto it
        return;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        nx0 += __shfl_xor_sync(0xffffffffu, nx0, d); nx1 += __shfl_xor_sync(0xffffffffu, nx1, d); nx2 += __shfl_xor_sync(0xffffffffu, nx2, d);
        ny0 += __shfl_xor_sync(0xffffffffu, ny0, d); ny1 += __shfl_xor_sync(0xffffffffu, ny1, d); ny2 += __shfl_xor_sync(0xffffffffu, ny2, d);
    }
    const float nxf[3] = {(float)nx0, (float)nx1, (float)nx2}, nyf[3] = {(float)ny0, (float)ny1, (float)ny2};
    float K[3][3], Cc[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Cc[r][c] = __fsub_rn(1.f, __fmul_rn((float)(r - 1), (float)(c - 1)));   // OT_module.py:113
            K[r][c] = expf(-inv_eps * Cc[r][c]);                                    // :116
        }
    const float c0 = 1.0f / (float)N;                              // :118-119
    float a[3] = {c0, c0, c0}, b[3] = {c0, c0, c0};
    for (int it = 0; it < L; ++it) {                               // :120-122, every lane the same scalars
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) sum = fmaf(__fmul_rn(nyf[c], K[r][c]), b[c], sum);
            a[r] = __fdiv_rn(c0, __fadd_rn(sum, kEps));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) sum = fmaf(__fmul_rn(nxf[r], K[r][c]), a[r], sum);
            b[c] = __fdiv_rn(c0, __fadd_rn(sum, kEps));
        }
    }
    // P = a K b^T, loss = <P, C>   (:129-134)
    float Pm[3][3], total = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Pm[r][c] = __fmul_rn(__fmul_rn(a[r], K[r][c]), b[c]);
            total = fmaf(__fmul_rn(__fmul_rn(nxf[r], nyf[c]), Pm[r][c]), Cc[r][c], total);
        }
    if (lane == 0) loss[prob] = total;
    if (gx == nullptr) return;
    // gradients with P constant: dL/dx^_i = -sum_j P_ij y^_j, then through the normalisation exactly as the dense kernels do
    float gxh[3], gyh[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) gxh[r] = -(__fmul_rn(nyf[2], Pm[r][2]) - __fmul_rn(nyf[0], Pm[r][0]));
#pragma unroll
    for (int c = 0; c < 3; ++c) gyh[c] = -(__fmul_rn(nxf[2], Pm[2][c]) - __fmul_rn(nxf[0], Pm[0][c]));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane + 32 * k;
        if (i < N) {
            {
                const float h = xh[k], g = h == -1.f ? gxh[0] : (h == 0.f ? gxh[1] : gxh[2]), den = dx[k];
                const float shrink = __fdiv_rn(__fsub_rn(den, kEps), den);
                gx[(long)prob * N + i] = __fdiv_rn(__fsub_rn(g, __fmul_rn(__fmul_rn(h, __fmul_rn(g, h)), shrink)), den);
            }
            {
                const float h = yh[k], g = h == -1.f ? gyh[0] : (h == 0.f ? gyh[1] : gyh[2]), den = dy[k];
                const float shrink = __fdiv_rn(__fsub_rn(den, kEps), den);
                gy[(long)prob * N + i] = __fdiv_rn(__fsub_rn(g, __fmul_rn(__fmul_rn(h, __fmul_rn(g, h)), shrink)), den);
            }
        }
    }
}

constexpr int kMaxSmemBytes = 227 * 1024;

static int pow2_floor(int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; }

static bool plan(int N, int CS, int keepC, SinkParams &p, size_t &bytes) {
    p.NR = (N + CS - 1) / CS;
    p.TPR = pow2_floor(kSinkThreads / p.NR > 0 ? kSinkThreads / p.NR : 1);
    if (p.TPR > 32) p.TPR = 32;
    p.TPC = pow2_floor(kSinkThreads / N > 0 ? kSinkThreads / N : 1);
    if (p.TPC > 8) p.TPC = 8;
    int pitch = N;
    while (pitch % 32 != p.TPR % 32) ++pitch;          // a-update: lane -> distinct bank
    p.pitch = pitch;
    p.keepC = keepC;
    SinkSmem L(N, p.NR, pitch, p.TPC, keepC);
    bytes = (size_t)L.total * sizeof(float);
    return bytes <= (size_t)kMaxSmemBytes && p.NR <= kSinkThreads;
}

}  // namespace fi

using namespace fi;

// Slices of D the cost matrix is spread over: enough CTAs for two waves of the 148 SMs, at least 64 columns each.
static int gram_slices(int n_problems, int N, int D) {
    const int tiles = ceil_div(N, kTile) * ceil_div(N, kTile);
    int S = ceil_div(2 * kNumSMs, n_problems * tiles);
    S = S < 1 ? 1 : (S > 32 ? 32 : S);
    if (S > D / 64) S = D / 64 > 0 ? D / 64 : 1;
    return S;
}
static int gram_slice_width(int D, int S) { return ceil_div(ceil_div(D, S), kDT) * kDT; }

// Scratch bytes of the large-D form (0: everything stays inside the solver kernel): per problem the plan (N x N) and its row
// norms (2N) -- what the split gradient reads -- and the S slices of the cost matrix.  `want_grad` no longer changes the size
// (the norms and the slices serve the forward pass too); the argument stays for ABI stability.
FI_API size_t fi_sinkhorn_workspace(int n_problems, int N, int D, int want_grad) {
    (void)want_grad;
    if (n_problems <= 0 || D < 128 || N < 4 || N > 128 || (N % 4) != 0) return 0;
    const size_t per = (size_t)N * N + 2 * (size_t)N;
    return (size_t)n_problems * (per + (size_t)gram_slices(n_problems, N, D) * N * N) * sizeof(float);
}

FI_API int fi_sinkhorn(const float *x, const float *y, int n_problems, int N, int D, float inv_eps, int L, float *loss,
                       float *grad_x, float *grad_y, cudaStream_t stream) {
    return fi_sinkhorn_ws(x, y, n_problems, N, D, inv_eps, L, loss, grad_x, grad_y, nullptr, 0, stream);
}

FI_API int fi_sinkhorn_ws(const float *x, const float *y, int n_problems, int N, int D, float inv_eps, int L, float *loss,
                          float *grad_x, float *grad_y, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    FI_REQUIRE(n_problems >= 0 && N >= 1 && N <= 256 && D >= 1 && L >= 1, "fi_sinkhorn: need N in [1,256], D >= 1, L >= 1 (N=%d D=%d L=%d)", N, D, L);
    if (n_problems == 0) return ok();
    FI_REQUIRE(x && y && loss, "fi_sinkhorn: null pointer");
    FI_REQUIRE((grad_x == nullptr) == (grad_y == nullptr), "fi_sinkhorn: grad_x and grad_y must both be given or both be NULL");
    // D = 1: solved by classes (O(L) per problem) unless an entry's normalised value is not exactly -1 / 0 / +1; those problems
    // keep NaN in their loss slot and the dense kernels below pick them up (masked).  FI_OPT_SINKHORN_GENERIC: 1 = dense only.
    const int dense_only = option(FI_OPT_SINKHORN_GENERIC);
    int masked = 0;
    if (D == 1 && dense_only == 0) {
        sinkhorn_d1_classes_kernel<<<ceil_div(n_problems, 8), 256, 0, stream>>>(x, y, loss, grad_x, grad_y, n_problems, N, inv_eps, L);
        if (int e = check_launch("fi_sinkhorn[d1 classes]")) return e;
        masked = 1;
    }
    if (N == 256 && D == 1 && dense_only != 2) {      // RoI-level loss: K in registers
        cudaError_t e = cudaSuccess;
        if (!masked) e = cudaMemsetAsync(loss, 0, sizeof(float) * n_problems, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_sinkhorn: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_problems * 2));
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, sinkhorn_n256_d1_kernel, x, y, loss, grad_x, grad_y, inv_eps, L, masked);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_sinkhorn: launch: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
        return check_launch("fi_sinkhorn[n256 d1]");
    }
    SinkParams p;
    p.x = x; p.y = y; p.loss = loss; p.gx = grad_x; p.gy = grad_y;
    p.N = N; p.D = D; p.L = L; p.inv_eps = inv_eps;
    size_t bytes = 0;
    int CS = 1;
    if (!plan(N, 1, 1, p, bytes) && !plan(N, 1, 0, p, bytes)) {
        CS = 2;
        if (!plan(N, 2, 1, p, bytes) && !plan(N, 2, 0, p, bytes)) {
            set_error(FI_ERR_UNSUPPORTED, "fi_sinkhorn: N=%d does not fit a CTA pair", N);
            return FI_ERR_UNSUPPORTED;
        }
    }
    p.masked = masked;
    p.pbuf = nullptr; p.nbuf = nullptr; p.gram = nullptr; p.S = 0;
    const size_t need_ws = fi_sinkhorn_workspace(n_problems, N, D, grad_x != nullptr);
    cudaError_t e;
    if (CS == 1 && p.keepC && need_ws > 0 && workspace != nullptr && workspace_bytes >= need_ws && ((uintptr_t)workspace % 16) == 0) {
        float *base = static_cast<float *>(workspace);
        float *gram = base + (size_t)n_problems * ((size_t)N * N + 2 * (size_t)N);
        const int S = gram_slices(n_problems, N, D), Dc = gram_slice_width(D, S);
        const long rows = 2L * n_problems * N;
        sinkhorn_norm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(x, y, base, n_problems, N, D);
        if (int rc = check_launch("fi_sinkhorn[norms]")) return rc;
        const int tiles = ceil_div(N, kTile) * ceil_div(N, kTile);
        sinkhorn_gram_split_kernel<<<dim3((unsigned)n_problems, (unsigned)S, (unsigned)tiles), kSinkThreads, 0, stream>>>(x, y, base, gram, N, D, S, Dc);
        if (int rc = check_launch("fi_sinkhorn[cost slices]")) return rc;
        p.nbuf = base; p.gram = gram; p.S = S;
        if (grad_x != nullptr) p.pbuf = base;
    }
    if (CS > 1 && !masked) {
        e = cudaMemsetAsync(loss, 0, sizeof(float) * n_problems, stream);
        if (e == cudaSuccess && grad_y) e = cudaMemsetAsync(grad_y, 0, sizeof(float) * (size_t)n_problems * N * D, stream);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_sinkhorn: memset: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    }
    auto kern = CS == 1 ? sinkhorn_kernel<1> : sinkhorn_kernel<2>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_sinkhorn: smem attr (%zu B): %s", bytes, cudaGetErrorString(e)); return FI_ERR_CUDA; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_problems * CS));
    cfg.blockDim = dim3(kSinkThreads);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_sinkhorn: launch: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    if (int rc = check_launch("fi_sinkhorn")) return rc;
    if (p.pbuf != nullptr) {                                 // large D: the gradient as two more launches over (problem, D chunk) / rows
        const size_t sm = ((size_t)N * N + 2 * (size_t)N + 2 * (size_t)N * 64) * sizeof(float);
        const int per = ((N + 15) / 16) * 4;
        void (*raw)(const float *, const float *, const float *, float *, float *, int, int) = nullptr;
        switch (per) {
            case 4: raw = sinkhorn_grad_raw_kernel<4>; break;
            case 8: raw = sinkhorn_grad_raw_kernel<8>; break;
            case 12: raw = sinkhorn_grad_raw_kernel<12>; break;
            case 16: raw = sinkhorn_grad_raw_kernel<16>; break;
            case 20: raw = sinkhorn_grad_raw_kernel<20>; break;
            case 24: raw = sinkhorn_grad_raw_kernel<24>; break;
            case 28: raw = sinkhorn_grad_raw_kernel<28>; break;
            default: raw = sinkhorn_grad_raw_kernel<32>; break;
        }
        e = cudaFuncSetAttribute(raw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "fi_sinkhorn: smem attr (%zu B): %s", sm, cudaGetErrorString(e)); return FI_ERR_CUDA; }
        raw<<<dim3((unsigned)ceil_div(D, 64), (unsigned)n_problems), 256, sm, stream>>>(x, y, p.pbuf, grad_x, grad_y, N, D);
        if (int rc = check_launch("fi_sinkhorn[gradient]")) return rc;
        const long rows = 2L * n_problems * N;
        if (D <= 256 * kChainCache) sinkhorn_grad_chain_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(x, y, p.pbuf, grad_x, grad_y, n_problems, N, D);
        else sinkhorn_grad_chain_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(x, y, p.pbuf, grad_x, grad_y, n_problems, N, D);
        return check_launch("fi_sinkhorn[gradient chain]");
    }
    return ok();
}
