// RoIAlign backward, L2-resident ("banded") reduction.
//
// ncu on the plain reduction kernel (profiles/r01_ncu_prof_bwd_sets.txt): 4.8 TB/s of DRAM traffic -- 73 % of the
// measured copy peak -- but 2.3x the algorithmic bytes: the dense maps are zero-filled up front (1.5 GB written),
// every line a reduction touches has left the 126 MB L2 by then and is fetched back (2.5 GB), and dirty lines are
// evicted more than once (2.85 GB).  The kernel is DRAM-bound on traffic it should not generate.
//
// RoIs come grouped by image (torch.nonzero order), and one image of the largest map is 71.5 MB.  So the work is
// ordered in BANDS: zero image b of a map, immediately reduce into it every crop of every set whose box lies in
// image b (the band is still in L2: reductions hit, nothing is fetched), move on; the band is written to DRAM once,
// when the next band pushes it out.  One persistent launch, no grid barrier: CTAs draw tickets from an ordered item
// list  Z(m,0) S(m,0) Z(m,1) S(m,1) ...  (Z = 256 KB zero-fill chunks, S = 8 warp-units of reduction work); an S item
// waits on its band's zero-fill counter.  Tickets are handed out in order and Z items never wait, so a waiting CTA is
// always waiting on CTAs that are already running: no deadlock for any grid size.
//
// A tiny single-CTA planner checks on the device that box_ind is non-decreasing per set and builds the item table;
// sets that are not sorted by image fall back to one band per map (zero everything, then reduce).
#include "roi_align_units.cuh"

namespace fi {

constexpr int kMaxBandSets = 12;
constexpr int kMaxBandImages = 64;
constexpr int kMaxEntries = 2 * kMaxBandSets * kMaxBandImages + 2 * kMaxBandSets;
constexpr long kZeroChunkFloats = 64 * 1024;          // 256 KB per zero item
constexpr int kUnitsPerItem = kWarpsPerBlock;         // one unit per warp

struct BandEntry {
    long first_ticket, count;
    float *zero_base;        // type 0: start of the region to clear
    long zero_floats;        //         its length
    int type;                // 0 = zero fill, 1 = reduce
    int set;                 // type 1: which crop set
    int r_lo, nbox;          //         boxes [r_lo, r_lo + nbox) of that set
    int counter;             // zero-fill completion counter of the band
    int need;                // type 1: value the counter must reach
    int pad0, pad1;
};

struct BandCtl {
    unsigned long long ticket;
    long total;
    int n_entries;
    int pad;
    int counters[kMaxBandSets * kMaxBandImages + kMaxBandSets];
    BandEntry e[kMaxEntries];
};

struct BandSets {
    BwdSet s[kMaxBandSets];
    int R[kMaxBandSets];
    int map_of[kMaxBandSets];       // index of the first set that names the same grads_image
    int n;
    int zero_first;
};

// ---- planner: one CTA ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) band_plan_kernel(const BandSets sets, BandCtl *__restrict__ ctl) {
    __shared__ int hist[kMaxBandSets][kMaxBandImages];
    __shared__ int sorted[kMaxBandSets];
    const int t = threadIdx.x;
    for (int k = t; k < kMaxBandSets * kMaxBandImages; k += blockDim.x) (&hist[0][0])[k] = 0;
    if (t < kMaxBandSets) sorted[t] = 1;
    __syncthreads();
    for (int s = 0; s < sets.n; ++s) {
        const int B = sets.s[s].B, R = sets.R[s];
        const int *bi = sets.s[s].box_ind;
        if (B > kMaxBandImages) { if (t == 0) sorted[s] = 0; continue; }
        for (int r = t; r < R; r += blockDim.x) {
            const int b = bi[r];
            if (b < 0 || b >= B) { sorted[s] = 0; continue; }            // such boxes are skipped by the reduction; keep it simple
            atomicAdd(&hist[s][b], 1);
            if (r > 0 && bi[r - 1] > b) sorted[s] = 0;
        }
    }
    __syncthreads();
    if (t != 0) return;
    int ne = 0, nc = 0;
    long ticket = 0;
    for (int m = 0; m < sets.n; ++m) {
        if (sets.map_of[m] != m) continue;                                 // visit every distinct map once
        const BwdSet &M = sets.s[m];
        const long img_floats = (long)M.H * M.W * M.C;
        bool banded = true;
        for (int s = 0; s < sets.n; ++s) if (sets.map_of[s] == m && !sorted[s]) banded = false;
        const int nbands = banded ? M.B : 1;
        int done_boxes[kMaxBandSets];
        for (int s = 0; s < sets.n; ++s) done_boxes[s] = 0;
        for (int band = 0; band < nbands; ++band) {
            const int counter = nc++;
            ctl->counters[counter] = 0;
            int need = 0;
            if (sets.zero_first) {
                BandEntry &z = ctl->e[ne++];
                z.type = 0; z.set = m; z.counter = counter;
                z.zero_base = M.gimg + (banded ? (long)band * img_floats : 0);
                z.zero_floats = banded ? img_floats : img_floats * M.B;
                z.first_ticket = ticket;
                z.count = (z.zero_floats + kZeroChunkFloats - 1) / kZeroChunkFloats;
                z.r_lo = z.nbox = z.need = 0;
                ticket += z.count;
                need = (int)z.count;
            }
            for (int s = 0; s < sets.n; ++s) {
                if (sets.map_of[s] != m) continue;
                const int nbox = banded ? hist[s][band] : sets.R[s];
                if (nbox == 0) continue;
                BandEntry &e = ctl->e[ne++];
                e.type = 1; e.set = s; e.counter = counter; e.need = need;
                e.r_lo = banded ? done_boxes[s] : 0;
                e.nbox = nbox;
                e.zero_base = nullptr; e.zero_floats = 0;
                const long units = (long)nbox * sets.s[s].ph * sets.s[s].slabs;
                e.first_ticket = ticket;
                e.count = (units + kUnitsPerItem - 1) / kUnitsPerItem;
                ticket += e.count;
                done_boxes[s] += nbox;
            }
        }
    }
    ctl->n_entries = ne;
    ctl->total = ticket;
    ctl->ticket = 0ULL;
}

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- persistent worker -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) band_run_kernel(const BandSets sets, BandCtl *__restrict__ ctl) {
    __shared__ long s_ticket;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long total = ctl->total;
    const int ne = ctl->n_entries;
    if (t == 0) s_ticket = (long)atomicAdd(&ctl->ticket, 1ULL);
    __syncthreads();
    long ticket = s_ticket;
    while (ticket < total) {
        __syncthreads();                                            // everyone has read s_ticket
        if (t == 0) s_ticket = (long)atomicAdd(&ctl->ticket, 1ULL);  // prefetch the next ticket while this item runs
        int lo = 0, hi = ne - 1;                                    // entry with first_ticket <= ticket < first_ticket + count
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ctl->e[mid].first_ticket <= ticket) lo = mid; else hi = mid - 1;
        }
        const BandEntry &e = ctl->e[lo];
        const long k = ticket - e.first_ticket;
        if (e.type == 0) {
            float4 *p = reinterpret_cast<float4 *>(e.zero_base + k * kZeroChunkFloats);
            const long n4 = min(kZeroChunkFloats, e.zero_floats - k * kZeroChunkFloats) >> 2;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (long q = t; q < n4; q += kWarpsPerBlock * 32) p[q] = z;                    // default policy: stays in L2
            __syncthreads();
            if (t == 0) { __threadfence(); atomicAdd(&ctl->counters[e.counter], 1); }
        } else {
            if (e.need > 0) {
                if (t == 0) while (ld_acquire(&ctl->counters[e.counter]) < e.need) __nanosleep(64);
                __syncthreads();
            }
            const BwdSet &S = sets.s[e.set];
            const long lu = k * kUnitsPerItem + w;                                          // local unit inside the band
            const long units = (long)e.nbox * S.ph * S.slabs;
            if (lu < units) {
                const int slab = (int)(lu % S.slabs);
                const long q = lu / S.slabs;
                const int r = e.r_lo + (int)(q / S.ph), i = (int)(q % S.ph);
                bwd_unit<4>(S, ((long)r * S.ph + i) * S.slabs + slab, lane);
            }
        }
        __syncthreads();
        ticket = s_ticket;
    }
}

}  // namespace fi

using namespace fi;

// Host entry used by fi_crop_sets_backward (roi_align.cu).  Returns FI_ERR_UNSUPPORTED when the shape does not qualify.
int fi_banded_backward(const fi_bwd_set *sets, int num_sets, int zero_first, cudaStream_t stream) {
    if (num_sets < 1 || num_sets > kMaxBandSets) return FI_ERR_UNSUPPORTED;
    BandSets dev;
    dev.n = 0;
    dev.zero_first = zero_first ? 1 : 0;
    for (int i = 0; i < num_sets; ++i) {
        const fi_bwd_set &h = sets[i];
        // empty sets still matter when their map must be cleared, so they are kept (R = 0)
        const bool vec = (h.depth % 128 == 0) && ((uintptr_t)h.grads_image % 16 == 0) && ((uintptr_t)h.grads % 16 == 0) && ((uintptr_t)h.grads2 % 16 == 0);
        if (!vec || ((long)h.image_height * h.image_width * h.depth) % 4 != 0) return FI_ERR_UNSUPPORTED;
        BwdSet &S = dev.s[dev.n];
        S.grads = h.grads; S.grads2 = h.grads2; S.boxes = h.boxes; S.box_ind = h.box_ind; S.src_row = h.src_row; S.gimg = h.grads_image;
        S.B = h.batch; S.H = h.image_height; S.W = h.image_width; S.C = h.depth; S.ph = h.crop_height; S.pw = h.crop_width; S.slabs = h.depth / 128;
        dev.R[dev.n] = h.num_boxes;
        dev.map_of[dev.n] = dev.n;
        for (int q = 0; q < dev.n; ++q)
            if (dev.s[q].gimg == S.gimg) { dev.map_of[dev.n] = dev.map_of[q]; break; }
        ++dev.n;
    }
    BandCtl *ctl = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&ctl, sizeof(BandCtl), stream);
    if (e != cudaSuccess) { set_error(FI_ERR_CUDA, "banded backward: workspace: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    band_plan_kernel<<<1, 1024, 0, stream>>>(dev, ctl);
    int rc = check_launch("banded backward[plan]");
    if (rc == FI_OK) {
        band_run_kernel<<<kNumSMs * 2, kWarpsPerBlock * 32, 0, stream>>>(dev, ctl);
        rc = check_launch("banded backward[run]");
    }
    cudaFreeAsync(ctl, stream);
    return rc;
}
