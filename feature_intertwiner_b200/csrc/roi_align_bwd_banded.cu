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
// when the next band pushes it out.  One persistent launch, no grid barrier: the warps of a chip-filling grid walk an
// ordered unit list  S(m,0) S(m,1) ...  (the usual (box,row,slab) reduction units) with a fixed stride; the first
// warps to reach a band clear it cooperatively (32 KB chunks drawn from a per-band counter -- no chunk has a static
// owner, so nobody waits on a warp that is busy elsewhere) and everyone waits only for the last chunks in flight.
//
// A tiny single-CTA planner checks on the device that box_ind is non-decreasing per set and builds the item table;
// sets that are not sorted by image fall back to one band per map (zero everything, then reduce).
#include "roi_align_units.cuh"

namespace fi {

constexpr int kMaxBandSets = 12;
constexpr int kMaxBandImages = 64;
constexpr int kMaxEntries = 2 * kMaxBandSets * kMaxBandImages + 2 * kMaxBandSets;
constexpr long kZeroChunkFloats = 8 * 1024;           // 32 KB per zero unit (one warp)
constexpr int kUnitsPerItem = 1;                      // list positions are warp-units

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
    int counters[2 * (kMaxBandSets * kMaxBandImages + kMaxBandSets)];
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
            const int counter = nc; nc += 2;                                     // [counter] = chunks handed out, [counter+1] = chunks done
            ctl->counters[counter] = 0; ctl->counters[counter + 1] = 0;
            float *zbase = M.gimg + (banded ? (long)band * img_floats : 0);
            const long zfloats = banded ? img_floats : img_floats * M.B;
            const int need = sets.zero_first ? (int)((zfloats + kZeroChunkFloats - 1) / kZeroChunkFloats) : 0;
            bool any = false;
            for (int s = 0; s < sets.n; ++s) {
                if (sets.map_of[s] != m) continue;
                const int nbox = banded ? hist[s][band] : sets.R[s];
                if (nbox == 0) continue;
                any = true;
                BandEntry &e = ctl->e[ne++];
                e.type = 1; e.set = s; e.counter = counter; e.need = need;
                e.r_lo = banded ? done_boxes[s] : 0;
                e.nbox = nbox;
                e.zero_base = zbase; e.zero_floats = zfloats;
                e.first_ticket = ticket;
                e.count = (long)nbox * sets.s[s].ph * sets.s[s].slabs;
                ticket += e.count;
                done_boxes[s] += nbox;
            }
            if (!any && need > 0) {                                              // nothing reduces into this band: one unit just clears it
                BandEntry &e = ctl->e[ne++];
                e.type = 0; e.set = m; e.counter = counter; e.need = need;
                e.r_lo = e.nbox = 0;
                e.zero_base = zbase; e.zero_floats = zfloats;
                e.first_ticket = ticket;
                e.count = 1;
                ticket += 1;
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
// Exactly as many CTAs as fit on the chip at once (cooperative launch: co-residency guaranteed), every WARP walks the
// ordered unit list with a fixed stride, no block-level synchronisation at all.  A reduce unit spins (lane 0, then
// __syncwarp) until its band's zero-fill counter is complete; zero units never wait, and since every warp is resident
// a waiting warp always waits on warps that are running.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) band_run_kernel(const BandSets sets, BandCtl *__restrict__ ctl) {
    const int lane = threadIdx.x & 31;
    const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nw = ((long)gridDim.x * blockDim.x) >> 5;
    const long total = ctl->total;
    const int ne = ctl->n_entries;
    int cur = 0;                                                    // entries are visited in increasing order: resume the search
    int ready_counter = -1;                                         // last band this warp has seen fully cleared
    for (long u = gw; u < total; u += nw) {
        while (cur + 1 < ne && ctl->e[cur + 1].first_ticket <= u) ++cur;
        const BandEntry &e = ctl->e[cur];
        if (e.need > 0 && ready_counter != e.counter) {
            // The band is cleared by whoever gets here first: grab 32 KB chunks until none are left, then wait for the
            // stragglers' chunks.  (No static owner of a chunk, so nobody waits on a warp that is busy elsewhere.)
            int *handed = &ctl->counters[e.counter], *done = &ctl->counters[e.counter + 1];
            if (ld_acquire(done) < e.need) {
                for (;;) {
                    int c = 0;
                    if (lane == 0) c = atomicAdd(handed, 1);
                    c = __shfl_sync(0xffffffffu, c, 0);
                    if (c >= e.need) break;
                    float4 *p = reinterpret_cast<float4 *>(e.zero_base + (long)c * kZeroChunkFloats);
                    const long n4 = min(kZeroChunkFloats, e.zero_floats - (long)c * kZeroChunkFloats) >> 2;
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (long q = lane; q < n4; q += 32) p[q] = z;                          // default policy: stays in L2
                    __syncwarp();
                    if (lane == 0) { __threadfence(); atomicAdd(done, 1); }
                }
                if (lane == 0) while (ld_acquire(done) < e.need) __nanosleep(32);
                __syncwarp();
            }
            ready_counter = e.counter;
        }
        if (e.type == 1) {
            const BwdSet &S = sets.s[e.set];
            bwd_unit<4>(S, (long)e.r_lo * S.ph * S.slabs + (u - e.first_ticket), lane);     // the band's units are contiguous in the set
        }
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
        static int blocks_per_sm = -1, num_sms = kNumSMs;
        if (blocks_per_sm < 0) {
            int devid = 0;
            cudaGetDevice(&devid);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, devid);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, band_run_kernel, kWarpsPerBlock * 32, 0) != cudaSuccess || blocks_per_sm < 1)
                blocks_per_sm = 1;
        }
        void *args[] = {(void *)&dev, (void *)&ctl};
        e = cudaLaunchCooperativeKernel((const void *)band_run_kernel, dim3(num_sms * blocks_per_sm), dim3(kWarpsPerBlock * 32), args, 0, stream);
        if (e != cudaSuccess) { cudaGetLastError(); cudaFreeAsync(ctl, stream); return FI_ERR_UNSUPPORTED; }   // caller falls back to the plain kernel
        rc = check_launch("banded backward[run]");
    }
    cudaFreeAsync(ctl, stream);
    return rc;
}
