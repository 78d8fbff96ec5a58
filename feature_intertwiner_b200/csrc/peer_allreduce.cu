// The one exchange step of the path when it is sharded by image batch, as ONE kernel over NVLink / NVSwitch peer memory.
//
// The reference runs nn.DataParallel: every replica returns its per-class feature sums and counts, gather concatenates them on
// GPU 0 and _merge_feat_vec reduces over (gpu, scale) (lib/model.py:217-224, 394-402; tools/utils.py:645-654).  One process
// per GPU replaces the gather by an all-reduce(SUM) of the packed sums (feature_intertwiner_b200/dist.py).  At 0.66 MB that
// collective is pure latency: NCCL's costs 0.12 ms at 8 ranks and cannot sit inside a CUDA graph capture next to autograd hooks
// (DESIGN.md section 6), which splits the step into three graphs with eager hops between them.  This kernel needs neither:
//
//   every rank owns a REGION in its own HBM that every peer maps (cudaIpc*):   header | slot 0 | slot 1
//   CTA b of rank r, call number e (kept on the device, so a graph replay advances it):
//     1. copies chunk b of the input into slot e&1 of its own region,
//     2. pushes "e" into flags[r][b] of EVERY peer's header (st.release.sys after __threadfence_system) and waits until its
//        own header shows >= e from every peer (ld.acquire.sys on LOCAL memory: the spin never crosses NVLink),
//     3. reads chunk b from slot e&1 of every region IN RANK ORDER 0..world-1 and adds them in that order -- every rank
//        computes bit-identical totals, which keeps the historical class buffers identical across ranks without a broadcast --
//     4. writes the total to the (local) output.
//   The slots alternate per call, so no second barrier is needed: a rank can only overwrite slot e&1 in call e+2, which it
//   enters after passing the barrier of call e+1, which every peer signals only after it finished reading in call e.
//   No CTA waits for another CTA of its own rank, so the grid needs no co-residency; the grid size and the chunk a CTA owns are
//   functions of the regions' capacity alone, hence equal on all ranks and for every call.  A peer that never arrives
//   (crashed rank) ends the spin after kSpinLimitNs with the region's error word set instead of hanging the GPU (fi_peer_error).
//
// One-shot (every rank reads world x n) is the right shape up to a few MB on NVSwitch: 8 x 0.66 MB = 5.3 MB per rank at
// >600 GB/s is < 10 us, the same order as the flag round trip.  The 15.7 MB OptTrans gradient uses the same kernel on a side
// stream under the RoIAlign backward (0.8 ms), where its 110 MB of NVLink reads per rank are hidden.
#include <stddef.h>
#include <string.h>

#include "fi_common.cuh"

namespace fi {

constexpr int kPeerMaxWorld = FI_PEER_MAX_WORLD;
constexpr int kPeerMaxCtas = 128;
constexpr int kPeerThreads = 512;
constexpr unsigned long long kSpinLimitNs = 4000000000ULL;     // 4 s

struct PeerHeader {
    unsigned flags[kPeerMaxWorld][kPeerMaxCtas];   // flags[r][b]: last call number rank r's CTA b has published (written by rank r)
    unsigned epoch[kPeerMaxCtas];                   // this rank's call counter, one copy per CTA (written by CTA b only)
    unsigned error;                                 // 1 = a wait timed out
    unsigned pad[127];
};
static_assert(sizeof(PeerHeader) % 256 == 0, "slots stay 256-byte aligned");

struct PeerArgs {
    char *region[kPeerMaxWorld];
    int rank, world;
    size_t n, cap;
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float4 add_in_order(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

__device__ __forceinline__ float *peer_slot(const PeerArgs &a, int r, unsigned e) {
    return reinterpret_cast<float *>(a.region[r] + sizeof(PeerHeader)) + (size_t)(e & 1u) * a.cap;
}

template <int WORLD>   // 0 = any world size (run-time loop)
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(const PeerArgs a, const float *in, float *out) {   // in == out is allowed
    const int b = blockIdx.x, t = threadIdx.x;
    const int world = WORLD ? WORLD : a.world;
    PeerHeader *mine = reinterpret_cast<PeerHeader *>(a.region[a.rank]);
    const unsigned e = mine->epoch[b] + 1u;
    // CTA b owns float4s [b * per, (b + 1) * per) of the CAPACITY whatever n is: a slot range is only ever written and read by the
    // CTAs with this index, the ones the flag round orders -- calls of different lengths may share one pair of regions
    const size_t n4 = a.n / 4, per = (a.cap / 4 + gridDim.x - 1) / gridDim.x;
    const size_t lo = (size_t)b * per < n4 ? (size_t)b * per : n4, hi = lo + per < n4 ? lo + per : n4;
    const bool tail = (size_t)b == n4 / per && (size_t)t < a.n - 4 * n4;       // the last n % 4 floats: one thread each, owner of float4 n4

    // 1. publish my chunk
    {
        float *slot = peer_slot(a, a.rank, e);
        const float4 *src = reinterpret_cast<const float4 *>(in);
        float4 *dst = reinterpret_cast<float4 *>(slot);
        for (size_t i = lo + t; i < hi; i += kPeerThreads) dst[i] = src[i];
        if (tail) slot[4 * n4 + t] = in[4 * n4 + t];
    }
    __threadfence_system();
    __syncthreads();
    // 2. flag round: thread r talks to rank r
    if (t < world) {
        st_release_sys(&reinterpret_cast<PeerHeader *>(a.region[t])->flags[a.rank][b], e);
        const unsigned *f = &mine->flags[t][b];
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(f) - e) < 0) {
            if ((++spins & 1023u) == 0) {
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kSpinLimitNs) { atomicExch(&mine->error, 1u); break; }
            }
        }
    }
    __syncthreads();
    // 3 + 4. totals in rank order
    if constexpr (WORLD > 0) {
        const float4 *slots[WORLD ? WORLD : 1];
#pragma unroll
        for (int r = 0; r < WORLD; ++r) slots[r] = reinterpret_cast<const float4 *>(peer_slot(a, r, e));
        for (size_t i = lo + t; i < hi; i += kPeerThreads) {
            float4 v[WORLD ? WORLD : 1];
#pragma unroll
            for (int r = 0; r < WORLD; ++r) v[r] = __ldcg(slots[r] + i);       // all loads in flight before the first add
            float4 acc = v[0];
#pragma unroll
            for (int r = 1; r < WORLD; ++r) acc = add_in_order(acc, v[r]);
            reinterpret_cast<float4 *>(out)[i] = acc;
        }
    } else {
        for (size_t i = lo + t; i < hi; i += kPeerThreads) {
            float4 acc = __ldcg(reinterpret_cast<const float4 *>(peer_slot(a, 0, e)) + i);
            for (int r = 1; r < world; ++r) acc = add_in_order(acc, __ldcg(reinterpret_cast<const float4 *>(peer_slot(a, r, e)) + i));
            reinterpret_cast<float4 *>(out)[i] = acc;
        }
    }
    if (tail) {
        float acc = __ldcg(peer_slot(a, 0, e) + 4 * n4 + t);
        for (int r = 1; r < world; ++r) acc = __fadd_rn(acc, __ldcg(peer_slot(a, r, e) + 4 * n4 + t));
        out[4 * n4 + t] = acc;
    }
    __syncthreads();
    if (t == 0) mine->epoch[b] = e;
}

static int peer_grid(size_t cap) {          // a function of the regions' capacity alone: equal on all ranks and for every call
    const size_t n4 = cap / 4;
    size_t g = (n4 + kPeerThreads - 1) / kPeerThreads;                        // one float4 per thread and rank, up to kPeerMaxCtas CTAs
    if (g < 1) g = 1;
    if (g > (size_t)kPeerMaxCtas) g = kPeerMaxCtas;
    return (int)g;
}

}  // namespace fi

FI_API size_t fi_peer_region_bytes(size_t capacity_floats) {
    const size_t cap = (capacity_floats + 63) / 64 * 64;
    return sizeof(fi::PeerHeader) + 2 * cap * sizeof(float);
}

FI_API int fi_peer_region_alloc(size_t capacity_floats, void **region) {
    FI_REQUIRE(region != nullptr && capacity_floats > 0, "fi_peer_region_alloc: region pointer and a capacity are required");
    void *p = nullptr;
    const size_t bytes = fi_peer_region_bytes(capacity_floats);
    // cudaMalloc, not a caller's pooled allocation: the IPC handle maps the WHOLE underlying allocation into every peer
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        fi::set_error(FI_ERR_CUDA, "fi_peer_region_alloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
        return FI_ERR_CUDA;
    }
    *region = p;
    return fi::ok();
}

FI_API int fi_peer_region_free(void *region) {
    if (region && cudaFree(region) != cudaSuccess) { cudaGetLastError(); fi::set_error(FI_ERR_CUDA, "fi_peer_region_free failed"); return FI_ERR_CUDA; }
    return fi::ok();
}

FI_API int fi_peer_region_export(void *region, unsigned char handle[FI_PEER_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == FI_PEER_HANDLE_BYTES, "handle size");
    FI_REQUIRE(region != nullptr && handle != nullptr, "fi_peer_region_export: null argument");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, region);
    if (e != cudaSuccess) { cudaGetLastError(); fi::set_error(FI_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    memcpy(handle, &h, sizeof(h));
    return fi::ok();
}

FI_API int fi_peer_region_import(const unsigned char handle[FI_PEER_HANDLE_BYTES], void **region) {
    FI_REQUIRE(region != nullptr && handle != nullptr, "fi_peer_region_import: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); fi::set_error(FI_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    *region = p;
    return fi::ok();
}

FI_API int fi_peer_region_release(void *imported) {
    if (imported && cudaIpcCloseMemHandle(imported) != cudaSuccess) { cudaGetLastError(); fi::set_error(FI_ERR_CUDA, "cudaIpcCloseMemHandle failed"); return FI_ERR_CUDA; }
    return fi::ok();
}

FI_API int fi_peer_error(void *own_region, int *error) {
    FI_REQUIRE(own_region != nullptr && error != nullptr, "fi_peer_error: null argument");
    unsigned v = 0;
    cudaError_t e = cudaMemcpy(&v, reinterpret_cast<char *>(own_region) + offsetof(fi::PeerHeader, error), sizeof(v), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { fi::set_error(FI_ERR_CUDA, "fi_peer_error: %s", cudaGetErrorString(e)); return FI_ERR_CUDA; }
    *error = (int)v;
    return fi::ok();
}

FI_API int fi_peer_allreduce_sum(const float *in, float *out, size_t n, int rank, int world, void *const *regions, size_t capacity_floats, cudaStream_t stream) {
    FI_REQUIRE(world >= 1 && world <= fi::kPeerMaxWorld && rank >= 0 && rank < world, "fi_peer_allreduce_sum: rank %d of %d (at most %d ranks)", rank,
               world, fi::kPeerMaxWorld);
    FI_REQUIRE(regions != nullptr && in != nullptr && out != nullptr, "fi_peer_allreduce_sum: null argument");
    FI_REQUIRE(n <= capacity_floats, "fi_peer_allreduce_sum: %zu floats exceed the regions' capacity %zu", n, capacity_floats);
    FI_REQUIRE(((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0), "fi_peer_allreduce_sum: in/out must be 16-byte aligned");
    if (n == 0) return fi::ok();
    fi::PeerArgs a;
    for (int r = 0; r < fi::kPeerMaxWorld; ++r) a.region[r] = r < world ? reinterpret_cast<char *>(regions[r]) : nullptr;
    for (int r = 0; r < world; ++r) FI_REQUIRE(a.region[r] != nullptr, "fi_peer_allreduce_sum: region of rank %d is null", r);
    a.rank = rank; a.world = world; a.n = n; a.cap = (capacity_floats + 63) / 64 * 64;
    const int grid = fi::peer_grid(a.cap);
    cudaStream_t s = stream;
    switch (world) {
        case 2: fi::peer_allreduce_kernel<2><<<grid, fi::kPeerThreads, 0, s>>>(a, in, out); break;
        case 4: fi::peer_allreduce_kernel<4><<<grid, fi::kPeerThreads, 0, s>>>(a, in, out); break;
        case 8: fi::peer_allreduce_kernel<8><<<grid, fi::kPeerThreads, 0, s>>>(a, in, out); break;
        default: fi::peer_allreduce_kernel<0><<<grid, fi::kPeerThreads, 0, s>>>(a, in, out); break;
    }
    return fi::check_launch("peer_allreduce_kernel");
}
