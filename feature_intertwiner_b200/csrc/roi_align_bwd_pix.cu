// RoIAlign backward: the accumulate kernel of the tile-owner form (roi_align_bwd_tile.cu), "pixel-owner" variant.
//
// What it replaces and why.  bin_accumulate_kernel keeps a tile (4x8 pixels x 128 channels) in shared memory and applies
// every queued sample as up to four LDS.128 / FFMA2 / STS.128 read-modify-writes, with the gradient rows coming through
// the same LSU pipe as LDG.E.128 into registers: ~28 shared-memory / L1 wavefronts per 512 B of gradient, ncu
// l1tex 79 % busy at 62 % of DRAM peak, 144 registers, 16 % occupancy (profiles/r01_ncu_bwd_tile_v8.txt).  Here
//   * the gradient rows never pass through the LSU or a register: a PRODUCER warp walks the tile's sample list and
//     stages each row (CB channels = 512 B or 1 KB) with one cp.async.bulk into a shared-memory ring, completion
//     counted on an mbarrier (expect_tx) -- SASS: UBLKCP + SYNCS;
//   * the accumulators live in REGISTERS: 8 CONSUMER warps per tile, warp w owns the 2x2 pixel block (w >> 2, w & 3)
//     of the 4x8 tile for all CB channels (4 pixels x CB/128 float4 per lane).  Per batch of staged rows every lane
//     first decides for ONE entry whether any of its taps falls into the warp's block (and with which weights); the
//     warp then walks the hits (ballot order = list order): one LDS.128 per 512 B of gradient, FFMA2 into registers.
//     No shared-memory stores, no read-modify-write chain: ~13 wavefronts per 512 B of gradient;
//   * CTAs are persistent (3 per SM, a two-batch ring each) and take tiles by tickets from a work counter, heaviest
//     (coarsest map) first; the producer prefetches tile descriptors a ticket ahead and list entries a batch ahead, so
//     no per-tile round trip is exposed, and the sample list is read once per tile (not once per 128-channel slab);
//   * runs of consecutive gradient rows (neighbouring samples of a crop row) are fetched by ONE bulk copy.
// Measured on C2 (profiles/r02_ncu_pix_v3.txt): 0.72 ms, 4.34 GB of DRAM traffic at 6.05 TB/s = 92 % of the copy peak
// (bin_accumulate_kernel: 1.05 ms, 62 %).
// Every pixel is still summed by ONE warp in list order (box, crop row, crop column) with the same operations as
// bin_accumulate_kernel, so both modes give the same bits as before: default = packed FMAs with pre-multiplied
// weights; EXACT = the un-fused arithmetic and order of crop_and_resize.c:190-250 (bit-identical to the reference's
// serial CPU loop).  Replaces crop_and_resize_kernel.cu:84-165 (zero fill + 4 atomics per crop element).
#include "roi_align_bwd_tile.cuh"

namespace fi {
namespace tile {

constexpr int kPixWarps = 8;                       // consumer warps per CTA
constexpr int kPixThreads = (kPixWarps + 1) * 32;  // + the producer warp
constexpr int kPixGroupMax = 8;                    // tiles per ticket of the work counter: 1..8 (P.bin.pix_group)
constexpr unsigned kFullMask = 0xffffffffu;

template <int CB, int BS, int NB>
struct __align__(128) PixSmem {
    float data[NB][BS][CB];                        // staged gradient rows (bulk-copy destination)
    float4 qw[NB][BS];                             // entry weights
    uint2 qa[NB][BS];                              // entry (row, pk2); pk2 == 0 for unused slots
    int4 hdr[NB];                                  // (entries | flags << 8 | map << 16, image, Y0, X0); flags: 1 first, 2 last, 4 done,
                                                   // 8 split layout (two-source batch: load-only halves in slots 0-15, tapped halves in 16-31)
    unsigned long long full[NB], empty[NB];        // mbarriers: rows + metadata landed / all consumers done with the slot
    const float *srcs[3 * kMaxSets];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// global -> shared bulk copy (TMA engine, no registers, no LSU): completion is signalled on `bar` in bytes
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// grid (MINB CTAs per SM, C / CB), kPixThreads threads, dynamic shared memory = sizeof(PixSmem)
template <bool EXACT, int CB, int BS, int NB, int MINB>
__global__ void __launch_bounds__(kPixThreads, MINB) pix_accumulate_kernel(const TParams P) {
    constexpr int TY = 4, TX = 8;
    constexpr int NCH = CB / 128;
    static_assert(kChunk % BS == 0 && BS <= 32, "a batch never straddles a list chunk");
    extern __shared__ __align__(128) unsigned char pix_smem_raw[];
    PixSmem<CB, BS, NB> &S = *reinterpret_cast<PixSmem<CB, BS, NB> *>(pix_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cb = blockIdx.y;
    if (threadIdx.x < 3 * kMaxSets) {
        const TSet &T = P.s[min((int)threadIdx.x / 3, P.nsets - 1)];
        const int which = threadIdx.x % 3;
        S.srcs[threadIdx.x] = which == 0 ? T.grads : which == 1 ? T.grads2 : T.coll;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NB; ++i) {
            mbar_init(&S.full[i], 1);
            mbar_init(&S.empty[i], kPixWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int total = P.bin.total_tiles;

    if (warp == kPixWarps) {
        // ================================================ producer ================================================
        int *work = P.bin.work + cb;
        // group of kPixGroup consecutive tiles (heaviest-first order); lane i holds (entries, first chunk) of tile t0 + i
        int my_n = 0, my_head = 0, nx_n = 0, nx_head = 0;
        // A ticket stands for kPixGroup tiles that are gridDim.x apart in the heaviest-first order: the CTAs that hold
        // neighbouring tickets then work on neighbouring tiles at the same time, so the gradient rows two adjacent tiles share
        // are still in L2 when the second one asks (consecutive tiles per ticket: L2 hit rate 6 %, +0.4 GB of DRAM reads).
        const int gx = (int)gridDim.x;
        const int kPixGroup = P.bin.pix_group;
        auto tile_of = [&](int ticket, int i) { return (ticket / gx) * (kPixGroup * gx) + (ticket % gx) + i * gx; };
        auto load_info = [&](int base, int &n, int &head) {
            n = 0; head = 0;
            const int t = tile_of(base, lane & (kPixGroupMax - 1));
            if (lane < kPixGroup && t < total) {
                n = P.bin.tile_n[total - 1 - t];
                head = P.bin.tile_head[total - 1 - t];            // only meaningful when n > 0
            }
        };
        int t0 = 0, t0n_l0 = 0;
        if (lane == 0) t0 = atomicAdd(work, 1);
        if (lane == 0) t0n_l0 = atomicAdd(work, 1);               // the next ticket stays in lane 0 until it is needed
        t0 = __shfl_sync(kFullMask, t0, 0);
        load_info(t0, my_n, my_head);
        auto count_of = [&](int ticket) {                       // tiles of the ticket that exist (a prefix: tile_of grows with i)
            int c = 0;
            while (c < kPixGroup && tile_of(ticket, c) < total) ++c;
            return c;
        };
        int gcount = count_of(t0), ti = 0;
        bool need_info = true;

        // current batch descriptor (t < 0: end marker) and its list entries (loads in flight)
        int c_t = gcount > 0 ? tile_of(t0, 0) : -1, c_n = 0, c_e0 = 0, c_chunk = 0, nextc = 0;
        if (c_t >= 0) {
            c_n = __shfl_sync(kFullMask, my_n, 0);
            c_chunk = __shfl_sync(kFullMask, my_head, 0);
            if (c_n > kChunk) nextc = P.bin.chunk_next[c_chunk];
        }
        auto load_entries = [&](int t, int n, int e0, int chunk, uint2 &ea, float4 &ew) {
            ea = make_uint2(0u, 0u);
            ew = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && lane < BS && e0 + lane < n) {
                const size_t idx = (size_t)chunk * kChunk + (e0 & (kChunk - 1)) + lane;
                ea = P.bin.qa[idx];
                ew = P.bin.qw[idx];
            }
        };
        uint2 ea; float4 ew;
        load_entries(c_t, c_n, c_e0, c_chunk, ea, ew);
        // geometry of the current tile
        int mi = 0, tb = 0, Y0 = 0, X0 = 0, C = P.m[0].C;
        auto decode = [&](int t) {
            int id = total - 1 - t;
            mi = 0;
            while (mi + 1 < P.nmaps && id >= P.m[mi + 1].first_tile) ++mi;
            const TMap &M = P.m[mi];
            id -= M.first_tile;
            const int per_img = M.tiles_x * M.tiles_y;
            tb = id / per_img;
            id -= tb * per_img;
            const int tyi = id / M.tiles_x;
            Y0 = tyi * TY; X0 = (id - tyi * M.tiles_x) * TX; C = M.C;
        };
        if (c_t >= 0) decode(c_t);

        for (int it = 0;; ++it) {
            // ---- descriptor of the batch after this one; its entries are requested before this batch is issued
            int n_t = c_t, n_n = c_n, n_e0 = c_e0 + BS, n_chunk = c_chunk;
            bool new_tile = false;
            if (c_t >= 0) {
                if (n_e0 < c_n) {
                    if ((n_e0 & (kChunk - 1)) == 0) {
                        n_chunk = nextc;
                        if (n_e0 + kChunk < c_n) nextc = P.bin.chunk_next[n_chunk];
                    }
                } else {
                    new_tile = true;
                    if (++ti == gcount) {                         // next group
                        t0 = __shfl_sync(kFullMask, t0n_l0, 0);
                        if (need_info) load_info(t0, nx_n, nx_head);
                        my_n = nx_n; my_head = nx_head;
                        ti = 0;
                        gcount = count_of(t0);
                        need_info = true;
                        if (gcount > 0 && lane == 0) t0n_l0 = atomicAdd(work, 1);
                    }
                    if (gcount <= 0) n_t = -1;
                    else {
                        if (need_info && ti >= min(gcount - 1, kPixGroup / 2)) {      // the group after this one: descriptors on their way
                            load_info(__shfl_sync(kFullMask, t0n_l0, 0), nx_n, nx_head);
                            need_info = false;
                        }
                        n_t = tile_of(t0, ti);
                        n_n = __shfl_sync(kFullMask, my_n, ti);
                        n_chunk = __shfl_sync(kFullMask, my_head, ti);
                        n_e0 = 0;
                        if (n_n > kChunk) nextc = P.bin.chunk_next[n_chunk];
                    }
                }
            }
            uint2 ea2; float4 ew2;
            load_entries(n_t, n_n, n_e0, n_chunk, ea2, ew2);

            // ---- issue the current batch into ring slot it % NB
            const int slot = it % NB;
            if (it >= NB) mbar_wait(&S.empty[slot], ((it / NB) - 1) & 1);
            // Two-source batches (every even entry is the load-only first half of a pair) are laid out split: first halves in
            // slots 0-15, tapped halves in 16-31 -- the rows of each half are then consecutive in memory AND in the ring.
            const int cnt = c_t >= 0 ? max(0, min(BS, c_n - c_e0)) : 0;
            bool split = false;
            if (BS == 32) {
                const unsigned even_valid = __ballot_sync(kFullMask, lane < cnt && !(lane & 1));
                const unsigned even_defer = __ballot_sync(kFullMask, lane < cnt && !(lane & 1) && ((int)ea.y & kDefer2));
                split = cnt > 2 && even_valid == even_defer;
                if (split) {
                    const int from = ((lane & 15) << 1) | (lane >> 4);
                    ea.x = __shfl_sync(kFullMask, ea.x, from); ea.y = __shfl_sync(kFullMask, ea.y, from);
                    ew.x = __shfl_sync(kFullMask, ew.x, from); ew.y = __shfl_sync(kFullMask, ew.y, from);
                    ew.z = __shfl_sync(kFullMask, ew.z, from); ew.w = __shfl_sync(kFullMask, ew.w, from);
                }
            }
            const int pk = (int)ea.y;
            if (lane < BS) {
                S.qa[slot][lane] = ea;
                S.qw[slot][lane] = ew;
            }
            if (lane == 0) {
                const int flags = c_t < 0 ? 4 : ((c_e0 == 0 ? 1 : 0) | (c_e0 + BS >= c_n ? 2 : 0) | (split ? 8 : 0));
                S.hdr[slot] = make_int4(cnt | (flags << 8) | (mi << 16), tb, Y0, X0);
            }
            // Runs of consecutive gradient rows (neighbouring samples of one crop row) in consecutive slots go out as ONE bulk
            // copy: the copies are issued one at a time by this warp (UBLKCP takes uniform operands), 3-5x fewer of them.
            const bool needs = (pk & (15 | kDefer2)) != 0;        // tap-less padding entries are not fetched
            const int src = (pk >> 5) & 63;
            const unsigned row_p = __shfl_up_sync(kFullMask, ea.x, 1);
            const int src_p = __shfl_up_sync(kFullMask, src, 1);
            const bool needs_p = __shfl_up_sync(kFullMask, (int)needs, 1) != 0;
            const bool cont = needs && lane > 0 && needs_p && src_p == src && ea.x == row_p + 1u && C == CB;
            const unsigned m = __ballot_sync(kFullMask, needs);
            const unsigned breaks = ~__ballot_sync(kFullMask, cont);       // lanes that do not continue their left neighbour
            __syncwarp();
            if (lane == 0) {
                if (m) mbar_arrive_expect_tx(&S.full[slot], (unsigned)__popc(m) * CB * 4u);
                else mbar_arrive(&S.full[slot]);
            }
            if (needs && !cont) {
                const unsigned above = lane < 31 ? (breaks >> (lane + 1)) : 0u;
                const int run = above ? __ffs(above) : 32 - lane;
                bulk_g2s(&S.data[slot][lane][0], S.srcs[src] + (size_t)ea.x * C + cb * CB, (unsigned)run * CB * 4u, &S.full[slot]);
            }
            if (c_t < 0) break;
            c_t = n_t; c_n = n_n; c_e0 = n_e0; c_chunk = n_chunk; ea = ea2; ew = ew2;
            if (new_tile && c_t >= 0) decode(c_t);
        }
        return;
    }

    // ==================================================== consumers ====================================================
    const int by = (warp >> 2) * 2, bx = (warp & 3) * 2;          // my 2x2 pixel block inside the tile
    float4 acc[4][NCH];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc[p][c] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int it = 0;; ++it) {
        const int slot = it % NB;
        mbar_wait(&S.full[slot], (it / NB) & 1);
        const int4 h = S.hdr[slot];
        const int flags = (h.x >> 8) & 0xff;
        if (flags & 4) break;
        if (flags & 1) {
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int c = 0; c < NCH; ++c) acc[p][c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // ---- lane l: does entry l touch my block, and with which weights (taps of one sample are distinct pixels)
        uint2 qa = make_uint2(0u, 0u);
        float4 qw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < BS) { qa = S.qa[slot][lane]; qw = S.qw[slot][lane]; }
        const int pk = (int)qa.y;
        const int px = pk >> 11;                                  // TL pixel index inside the tile (signed)
        int mask = 0;
        float wa0 = 0.f, wa1 = 0.f, wa2 = 0.f, wa3 = 0.f;          // default: tap weight per block pixel; EXACT: y factor
        float wb0 = 0.f, wb1 = 0.f, wb2 = 0.f, wb3 = 0.f;          // EXACT: x factor
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int pidx = px + (k & 1) + (k >> 1) * TX;        // in [0, 32) whenever the tap's flag is set
            const bool hit = ((pk >> k) & 1) && (pidx >> 4) == (warp >> 2) && ((pidx >> 1) & 3) == (warp & 3);
            const int loc = ((pidx >> 3) & 1) * 2 + (pidx & 1);
            float wk, wx = 0.f;
            if (EXACT) {                                          // qw = (1 - fy, fy, 1 - fx, fx), crop_and_resize.c:241-247
                wk = (k < 2) ? qw.x : qw.y;
                wx = (k & 1) ? qw.w : qw.z;
            } else {
                wk = k == 0 ? qw.x : k == 1 ? qw.y : k == 2 ? qw.z : qw.w;
            }
            if (hit) {
                mask |= 1 << loc;
                if (loc == 0) { wa0 = wk; wb0 = wx; }
                if (loc == 1) { wa1 = wk; wb1 = wx; }
                if (loc == 2) { wa2 = wk; wb2 = wx; }
                if (loc == 3) { wa3 = wk; wb3 = wx; }
            }
        }
        // two-source samples: the entry before a tapped one may be its load-only first half (pairs start on even slots)
        const int poff = (flags & 8) ? 16 : 1;                    // split layout: the first half sits 16 slots below
        const int prev_pk = __shfl_sync(kFullMask, pk, (lane - poff) & 31);
        const int info = mask | ((lane >= poff && (prev_pk & kDefer2)) ? 16 : 0);
        unsigned hits = __ballot_sync(kFullMask, mask != 0);
        const float4 *rows = reinterpret_cast<const float4 *>(&S.data[slot][0][0]) + lane;
        while (hits) {
            const int e = __ffs(hits) - 1;
            hits &= hits - 1;
            const int m = __shfl_sync(kFullMask, info, e);
            float4 g[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) g[c] = rows[e * (CB / 4) + c * 32];
            if (m & 16) {                                         // (g1 + g2) first, like autograd's accumulation in the reference
#pragma unroll
                for (int c = 0; c < NCH; ++c) g[c] = add_rn4(g[c], rows[(e - poff) * (CB / 4) + c * 32]);
            }
            const float u0 = __shfl_sync(kFullMask, wa0, e), u1 = __shfl_sync(kFullMask, wa1, e);
            const float u2 = __shfl_sync(kFullMask, wa2, e), u3 = __shfl_sync(kFullMask, wa3, e);
            if (EXACT) {
                const float v0 = __shfl_sync(kFullMask, wb0, e), v1 = __shfl_sync(kFullMask, wb1, e);
                const float v2 = __shfl_sync(kFullMask, wb2, e), v3 = __shfl_sync(kFullMask, wb3, e);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (m & 1) acc[0][c] = add_rn4(acc[0][c], mul_rn4(v0, mul_rn4(u0, g[c])));
                    if (m & 2) acc[1][c] = add_rn4(acc[1][c], mul_rn4(v1, mul_rn4(u1, g[c])));
                    if (m & 4) acc[2][c] = add_rn4(acc[2][c], mul_rn4(v2, mul_rn4(u2, g[c])));
                    if (m & 8) acc[3][c] = add_rn4(acc[3][c], mul_rn4(v3, mul_rn4(u3, g[c])));
                }
            } else {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (m & 1) acc[0][c] = fma4(g[c], u0, acc[0][c]);
                    if (m & 2) acc[1][c] = fma4(g[c], u1, acc[1][c]);
                    if (m & 4) acc[2][c] = fma4(g[c], u2, acc[2][c]);
                    if (m & 8) acc[3][c] = fma4(g[c], u3, acc[3][c]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[slot]);               // the slot's rows and metadata are no longer needed by this warp
        if (flags & 2) {
            // ---- last batch of the tile: my four pixels go out, once
            const TMap &M = P.m[(h.x >> 16) & 0xff];
            const int H = M.H, W = M.W, C = M.C;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int y = h.z + by + (p >> 1), x = h.w + bx + (p & 1);
                if (y < H && x < W) {
                    float *dst = M.gimg + (((long)h.y * H + y) * (long)W + x) * C + cb * CB + lane * 4;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        float4 v = acc[p][c];
                        if (P.accumulate) v = add_rn4(*reinterpret_cast<const float4 *>(dst + c * 128), v);
                        __stcs(reinterpret_cast<float4 *>(dst + c * 128), v);
                    }
                }
            }
        }
    }
}

template <bool EXACT, int CB, int BS, int NB, int MINB>
static int launch_pix(const TParams &P, cudaStream_t stream) {
    const int ctas_per_sm = MINB;
    const size_t smem = sizeof(PixSmem<CB, BS, NB>);
    static bool configured_dev[64] = {};                          // the attribute is per function AND per device
    int dev = 0;
    cudaGetDevice(&dev);
    bool &configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(pix_accumulate_kernel<EXACT, CB, BS, NB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "crop backward[accumulate]: shared memory attribute: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
        configured = true;
    }
    const int chan_blocks = P.m[0].C / CB;
    const long want = P.bin.total_tiles;
    const int grid_x = (int)(want < (long)kNumSMs * ctas_per_sm ? want : (long)kNumSMs * ctas_per_sm);
    pix_accumulate_kernel<EXACT, CB, BS, NB, MINB><<<dim3(grid_x, chan_blocks), kPixThreads, smem, stream>>>(P);
    return check_launch("crop backward[accumulate]");
}

// All maps of a launch must share C (they do: one channel block index per CTA).  Returns FI_ERR_UNSUPPORTED otherwise.
int pix_accumulate(const TParams &P, int exact, cudaStream_t stream) {
    const int C = P.m[0].C;
    for (int m = 1; m < P.nmaps; ++m)
        if (P.m[m].C != C) return FI_ERR_UNSUPPORTED;
    if (C % 128 != 0 || C / 128 > 32) return FI_ERR_UNSUPPORTED;
    // Ring shape (measured on C2, profiles/r02_pix_sweep.json): three CTAs per SM with a two-batch ring beat two CTAs with
    // three batches (0.72 vs 0.86 ms, DRAM at 92 % vs 80 % of the copy peak): each CTA is latency-bound on its own ring, so
    // more independent rings per SM keep more bytes in flight; 16-slot batches lose to the per-batch work of the consumers.
    // (Tried for small maps -- a level-4 / level-5 call alone has fewer tiles than a few rounds of the persistent grid and lasts
    // as long as its heaviest tile: 128-channel work items on a 256-channel map, twice the items at half the bytes, were SLOWER,
    // 172 vs 125 us at level 5 and 131 vs 93 us at level 4 of C2: the per-entry work of the producer, not the bytes, is the tile's
    // critical path, and rows of half a pixel cannot be run-coalesced.  profiles/r02_small_map_bwd.txt)
    const int cfg = option(FI_OPT_PIX_CFG);
    if (C % 256 == 0) {
        if (cfg == 1) return exact ? launch_pix<true, 256, 32, 3, 2>(P, stream) : launch_pix<false, 256, 32, 3, 2>(P, stream);
        if (cfg == 2) return exact ? launch_pix<true, 256, 16, 6, 2>(P, stream) : launch_pix<false, 256, 16, 6, 2>(P, stream);
        if (cfg == 3) return exact ? launch_pix<true, 256, 16, 4, 3>(P, stream) : launch_pix<false, 256, 16, 4, 3>(P, stream);
        return exact ? launch_pix<true, 256, 32, 2, 3>(P, stream) : launch_pix<false, 256, 32, 2, 3>(P, stream);
    }
    return exact ? launch_pix<true, 128, 32, 3, 3>(P, stream) : launch_pix<false, 128, 32, 3, 3>(P, stream);
}

}  // namespace tile
}  // namespace fi
