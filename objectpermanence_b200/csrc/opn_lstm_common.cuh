// Shared pieces of the persistent LSTM recurrence kernels (opn_lstm.cu: FP32 FMA math; opn_lstm_mma.cu:
// split-fp16 tensor-core math): launch parameters, the flag-in-data exchange primitives for both media
// (global-memory ring / thread-block-cluster shared memory), the polling gather with its time-out.
#pragma once
#include <cooperative_groups.h>

#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "opn_common.cuh"

namespace cg = cooperative_groups;

namespace opn {

constexpr int kThreads = 128;
constexpr int kGroup = 8;   // videos per batch group
constexpr int kUnits = 8;   // hidden units per CTA
constexpr long long kTimeoutCycles = 3000000000LL;  // ~1.5 s at 2 GHz

constexpr uint32_t kStatusPollTimeout = 1;

struct FwdParams {
    const float* xproj;   // [B,T,4H]
    const float* w_hh;    // [4H,H]
    float* hs;            // [B,T,H]
    float* gates;         // [B,T,4H] or null
    float* cells;         // [B,T,H] or null
    uint32_t* ring;       // [n_groups_total][2][8][H]   flagged copies of h_t
    unsigned int* status; // 4 words
    int B, T;
    int group_offset;  // first batch group handled by this launch
    int n_slices;      // H / 8
};

struct BwdParams {
    const float* w_hh;    // [4H,H]
    const float* gates;   // [B,T,4H]
    const float* cells;   // [B,T,H]
    const float* dh_out;  // [B,T,H]
    float* dgates;        // [B,T,4H]
    uint32_t* ring;       // [n_groups_total][2][H/U producers][8][H]  flagged partial products
    unsigned int* status;
    int B, T;
    int group_offset;
    int n_slices;
};

namespace {

// parity carried by the words of step t: slot t&1 is rewritten every 2 steps, so the bit
// alternates per rewrite; the first write (t = 0, 1) carries 1 to differ from the zeroed ring.
__device__ __forceinline__ uint32_t step_parity(int t) { return ((uint32_t)(t >> 1) & 1u) ^ 1u; }

__device__ __forceinline__ void st_flagged(uint32_t* p, float v, uint32_t parity) {
    const uint32_t bits = (__float_as_uint(v) & ~1u) | parity;
    asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(bits) : "memory");
}
__device__ __forceinline__ uint4 ld_flagged4(const uint32_t* p) {
    uint4 v;
    // ld.volatile streams at ~80 B/clk/SM from L2, ld.relaxed.gpu / ld.global.cg at ~40 and ld.acquire.gpu at ~3
    // (measured with tools/load_flavors.cu, profiles/r01_lstm_handoff.md); every 32-bit element is still a
    // single-copy-atomic access that is always served by L2.
    asm volatile("ld.volatile.global.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ bool ready4(const uint4& v, uint32_t parity) {
    return ((((v.x ^ parity) | (v.y ^ parity) | (v.z ^ parity) | (v.w ^ parity)) & 1u) == 0u);
}

// Time-out bookkeeping of the polling loops: called every 64 unsuccessful sweeps.  Returns true
// when the caller must give up (another CTA reported a failure, or this wait has expired).
__device__ __noinline__ bool poll_expired(long long t0, unsigned int* status, int t) {
    if (ld_relaxed(status) != 0) return true;
    if (clock64() - t0 > kTimeoutCycles) {
        if (atomicCAS(status, 0u, kStatusPollTimeout) == 0u) {
            status[1] = (unsigned int)t;
            status[2] = blockIdx.x;
            status[3] = threadIdx.x;
        }
        return true;
    }
    return false;
}

// ---- thread-block-cluster flavour of the exchange: peers push flagged words straight into this CTA's shared
// memory (st.shared::cluster through a mapa-translated address) and the CTA polls its OWN shared memory.
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_peer_v2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.relaxed.cluster.shared::cluster.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void st_peer_b32(uint32_t a, uint32_t x) {
    asm volatile("st.relaxed.cluster.shared::cluster.b32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}
__device__ __forceinline__ uint4 lds_flagged4(uint32_t a) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(a)
                 : "memory");
    return v;
}
__device__ __forceinline__ uint32_t flagged(float v, uint32_t parity) { return (__float_as_uint(v) & ~1u) | parity; }

__device__ __forceinline__ uint4 ld_word(const uint4*, const uint32_t* p) { return ld_flagged4(p); }
__device__ __forceinline__ uint4 ld_word(const uint4*, uint32_t smem_addr) { return lds_flagged4(smem_addr); }
__device__ __forceinline__ bool word_ready(const uint4& v, uint32_t parity) { return ready4(v, parity); }

// Fetch N flagged words (uint4 or uint32_t) whose addresses / validity are given by functors.
// Every load of a sweep is issued back to back (one L2 round trip when the data is already there);
// stale words are re-fetched in further sweeps.  (A variant that spun on a single canary word per
// thread and fetched the rest afterwards was measured slower on B200 -- 10.0 vs 3.8 us/step for the
// H=256 backward recurrence -- and was dropped; see profiles/r01_lstm_handoff.md.)
// Returns false on time-out / abort.
template <int N, typename Word, typename AddrFn, typename ValidFn>
__device__ __forceinline__ bool gather_flagged(Word (&v)[N], AddrFn addr, ValidFn valid, uint32_t par,
                                               unsigned int* status, int t) {
    bool pending = false;
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (valid(i)) v[i] = ld_word((const Word*)nullptr, addr(i));
#pragma unroll
    for (int i = 0; i < N; ++i)
        if (valid(i) && !word_ready(v[i], par)) pending = true;
    if (!pending) return true;

    const long long t0 = clock64();
    unsigned int sweeps = 0;
    for (;;) {
        pending = false;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (valid(i) && !word_ready(v[i], par)) v[i] = ld_word((const Word*)nullptr, addr(i));
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (valid(i) && !word_ready(v[i], par)) pending = true;
        if (!pending) return true;
        if ((++sweeps & 63u) == 0 && poll_expired(t0, status, t)) return false;
    }
}

// One stage of the transposing butterfly: the N live values (compact index) are halved; bit
// ABIT of the compact index is resolved by lane bit `mask`.
template <int N, int ABIT, int SZ>
__device__ __forceinline__ void butterfly_stage(float (&v)[SZ], bool hi, int mask) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const int lo = ((i >> ABIT) << (ABIT + 1)) | (i & ((1 << ABIT) - 1));
        const int up = lo | (1 << ABIT);
        const float send = hi ? v[lo] : v[up];
        const float keep = hi ? v[up] : v[lo];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
}

// ---- host-side launch helpers ---------------------------------------------------------------------------
template <typename Kernel>
int max_coresident(Kernel kernel, int threads, size_t smem, int* out) {
    int dev = 0, sms = 0, per_sm = 0;
    OPN_CUDA(cudaGetDevice(&dev));
    OPN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    OPN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OPN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    *out = sms * per_sm;
    return OPN_OK;
}

// Ring flavour: the n_slices CTAs of a batch group exchange through global memory, so they must be co-resident:
// cooperative launches of as many whole batch groups as fit on the device.
// cluster_size > 1: the CTAs are additionally grouped into thread-block clusters of that many consecutive slices
// (n_slices must be a multiple of it): cooperative launch + cluster dimension.
// group_begin / group_end: only the batch groups [group_begin, group_end) (default: all of them)
template <typename Kernel, typename Params>
int launch_ring(Kernel kernel, Params p, int threads, int n_slices, size_t smem, int64_t B, cudaStream_t stream,
                const char* what, int cluster_size = 1, int group_begin = 0, int group_end = -1) {
    // the capacity of the device for this kernel is queried once (the occupancy calls cost 0.1-1 ms of host time each)
    static std::mutex cache_mutex;
    static std::map<std::pair<const void*, int>, int> cap_cache;   // (kernel, device) -> co-resident CTAs
    int dev = 0;
    OPN_CUDA(cudaGetDevice(&dev));
    const std::pair<const void*, int> key((const void*)kernel, dev);
    int cap = -1;
    {
        std::lock_guard<std::mutex> lock(cache_mutex);
        auto it = cap_cache.find(key);
        if (it != cap_cache.end()) cap = it->second;
    }
    if (cap < 0) {
        int rc = max_coresident(kernel, threads, smem, &cap);
        if (rc != OPN_OK) return rc;
        if (cluster_size > 1) {
            // whole clusters must fit: ask the occupancy calculator for the cluster flavour
            cudaLaunchConfig_t probe = {};
            probe.gridDim = dim3((unsigned)n_slices);
            probe.blockDim = dim3((unsigned)threads);
            probe.dynamicSmemBytes = smem;
            cudaLaunchAttribute pa[1];
            pa[0].id = cudaLaunchAttributeClusterDimension;
            pa[0].val.clusterDim.x = (unsigned)cluster_size;
            pa[0].val.clusterDim.y = 1;
            pa[0].val.clusterDim.z = 1;
            probe.attrs = pa;
            probe.numAttrs = 1;
            int max_clusters = 0;
            OPN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &probe));
            if (max_clusters * cluster_size < cap) cap = max_clusters * cluster_size;
        }
        std::lock_guard<std::mutex> lock(cache_mutex);
        cap_cache[key] = cap;
    }
    int groups = (int)((B + kGroup - 1) / kGroup);
    if (group_end >= 0 && group_end < groups) groups = group_end;
    const int per_launch = cap / n_slices;
    if (per_launch < 1) {
        set_error("%s: device cannot co-schedule %d CTAs (capacity %d)", what, n_slices, cap);
        return OPN_ERR_UNSUPPORTED;
    }
    p.n_slices = n_slices;
    for (int g0 = group_begin; g0 < groups; g0 += per_launch) {
        const int ng = groups - g0 < per_launch ? groups - g0 : per_launch;
        p.group_offset = g0;
        if (cluster_size > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(n_slices * ng));
            cfg.blockDim = dim3((unsigned)threads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)cluster_size;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeCooperative;
            attr[1].val.cooperative = 1;
            cfg.attrs = attr;
            // Nsight Compute cannot replay cooperative cluster launches (the driver's own ncu pass of round 1 died on this
            // kernel).  The cooperative attribute only asks the driver to verify co-residency -- the grid never exceeds
            // the capacity computed above and every wait has a time-out -- so it is dropped when a profiler is injected
            // (CUDA_INJECTION64_PATH / NV_COMPUTE_PROFILER_PERFWORKS_DIR are set by ncu for its target), when
            // OPN_NO_COOP_CLUSTER=1 asks for it, or after a cooperative launch has been refused once.
            static int plain_cluster = -1;
            if (plain_cluster < 0) {
                const char* e = getenv("OPN_NO_COOP_CLUSTER");
                plain_cluster = ((e && e[0] == '1') || getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")) ? 1 : 0;
            }
            cfg.numAttrs = plain_cluster ? 1 : 2;
            cudaError_t le = cudaLaunchKernelEx(&cfg, kernel, p);
            if (le != cudaSuccess && !plain_cluster) {
                (void)cudaGetLastError();
                plain_cluster = 1;
                cfg.numAttrs = 1;
                le = cudaLaunchKernelEx(&cfg, kernel, p);
            }
            OPN_CUDA(le);
        } else {
            void* args[] = {(void*)&p};
            OPN_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(n_slices * ng), dim3(threads), args, smem,
                                                 stream));
        }
        count_launch();
    }
    return OPN_OK;
}

// Cluster flavour: one launch, one thread-block cluster of `cluster_size` CTAs per batch group.  Clusters are
// independent, so the grid may exceed the device (later clusters start as earlier ones retire).
// *launched = false (and OPN_OK) when this device cannot host such a cluster: the caller falls back to the ring.
template <typename Kernel, typename Params>
int launch_cluster(Kernel kernel, Params p, int threads, int cluster_size, size_t smem, int64_t B, cudaStream_t stream,
                   bool* launched) {
    *launched = false;
    if (cluster_size > 8)
        OPN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    OPN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int groups = (int)((B + kGroup - 1) / kGroup);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * cluster_size));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        return OPN_OK;
    }
    p.n_slices = cluster_size;
    p.group_offset = 0;
    OPN_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
    count_launch();
    *launched = true;
    return OPN_OK;
}

}  // namespace
}  // namespace opn
