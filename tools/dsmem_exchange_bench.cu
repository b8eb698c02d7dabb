// Micro-benchmark of a per-step all-gather among the CTAs of one thread-block cluster through
// distributed shared memory (DSMEM), with the ready bit carried in the data (LSB of each word =
// step parity), as a replacement for the L2 ring of the persistent LSTM kernels.
//
// Each of CS CTAs (256 threads) owns 16 units x 8 videos = 128 words per step and pushes them into
// the operand buffer [8 videos][H = 16*CS] of every CTA of the cluster (st.shared::cluster); every
// CTA polls its OWN shared memory until all 8*H words carry the step parity, then barriers.
//   store variants: 0 = v4 (values staged through local smem + __syncthreads, 2 x 16 B per thread)
//                   1 = v2 (4 x 8 B per thread, no staging barrier)
//                   2 = scalar (128 threads x CS x 4 B)
//                   3 = v4 staged + mbarrier signalling (remote arrive.release.cluster per warp) instead of flags
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_exchange_bench tools/dsmem_exchange_bench.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t a, uint4 v) {
    asm volatile("st.relaxed.cluster.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.relaxed.cluster.shared::cluster.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void st_cluster_b32(uint32_t a, uint32_t x) {
    asm volatile("st.relaxed.cluster.shared::cluster.b32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}
__device__ __forceinline__ uint4 lds_volatile_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ bool ready4(const uint4& v, uint32_t par) {
    return ((((v.x ^ par) | (v.y ^ par) | (v.z ^ par) | (v.w ^ par)) & 1u) == 0u);
}

template <int CS>
__global__ void __launch_bounds__(256, 1) dsmem_kernel(int steps, int variant, int compute_cycles, long long* out,
                                                        unsigned int* fail, uint32_t* sink) {
    constexpr int H = 16 * CS;
    constexpr int NV = (8 * H / 4 + 255) / 256;  // float4 per thread to poll
    __shared__ __align__(16) uint32_t buf[2][8 * H];
    __shared__ __align__(16) uint32_t stage[8 * 16];
    __shared__ __align__(8) uint64_t bars[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster.block_rank();
    for (int i = tid; i < 2 * 8 * H; i += 256) (&buf[0][0])[i] = 0;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(CS) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();
    long long t_begin = 0;
    uint32_t acc = 0;
    unsigned int failed = 0;
    for (int s = 0; s < steps && !failed; ++s) {
        if (s == 16 && tid == 0) t_begin = clock64();
        const uint32_t par = ((uint32_t)(s >> 1) & 1u) ^ 1u;
        const int slot = s & 1;
        const uint32_t val = ((uint32_t)(s * 2654435761u + tid) & ~1u) | par;
        const uint32_t base = smem_u32(&buf[slot][0]);
        if (variant == 0 || variant == 3) {
            if (tid < 128) stage[tid] = val;  // [b][16 units]
            __syncthreads();
            // 32 float4 per destination (8 videos x 4 quads), CS destinations: CS*32 stores over 256 threads
#pragma unroll
            for (int i = 0; i < CS * 32 / 256; ++i) {
                const int idx = tid + 256 * i;
                const int d = idx >> 5, v = idx & 31, b = v >> 2, q = v & 3;
                const uint4 x = *reinterpret_cast<const uint4*>(&stage[b * 16 + 4 * q]);
                st_cluster_v4(mapa(base + (b * H + 16 * rank + 4 * q) * 4, d), x);
            }
            if (variant == 3) {
                __syncwarp();
                // warp w stored to destinations { (w*32 + l + 256 i) >> 5 } = w + 8 i: lane i arrives there
                if (lane < CS * 32 / 256) {
                    const uint32_t rb = mapa(smem_u32(&bars[slot]), warp + 8 * lane);
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rb) : "memory");
                }
            }
        } else if (variant == 1) {
            // thread: video b = (lane>>1)&7 ... emulate: each thread pushes 8 B (2 units) of one video to CS/4 destinations... use 4 dests
            const int b = (lane >> 2) & 7, up = warp;  // unit pair index 0..7 (2 units each)
#pragma unroll
            for (int i = 0; i < CS / 4; ++i) {
                const int d = (lane & 3) * (CS / 4) + i;
                st_cluster_v2(mapa(base + (b * H + 16 * rank + 2 * up) * 4, d), val, val);
            }
        } else {
            if (tid < 128) {
                const int b = tid >> 4, u = tid & 15;
#pragma unroll 4
                for (int d = 0; d < CS; ++d) st_cluster_b32(mapa(base + (b * H + 16 * rank + u) * 4, d), val);
            }
        }
        // ---- wait for the whole buffer
        if (variant == 3) {
            uint32_t ok = 0;
            const uint32_t ph = (uint32_t)(s >> 1) & 1u;
            long long t0 = clock64();
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bars[slot])), "r"(ph) : "memory");
                if (!ok && clock64() - t0 > 2000000000LL) { atomicExch(fail, 1u + s); failed = 1; break; }
            }
            acc += buf[slot][tid];
        } else {
            uint4 v[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = lds_volatile_v4(base + (tid + 256 * i) * 16);
            bool pending = false;
#pragma unroll
            for (int i = 0; i < NV; ++i) if ((tid + 256 * i) * 4 < 8 * H && !ready4(v[i], par)) pending = true;
            long long t0 = clock64();
            while (pending) {
                pending = false;
#pragma unroll
                for (int i = 0; i < NV; ++i) if ((tid + 256 * i) * 4 < 8 * H && !ready4(v[i], par)) v[i] = lds_volatile_v4(base + (tid + 256 * i) * 16);
#pragma unroll
                for (int i = 0; i < NV; ++i) if ((tid + 256 * i) * 4 < 8 * H && !ready4(v[i], par)) pending = true;
                if (pending && clock64() - t0 > 2000000000LL) { atomicExch(fail, 1u + s); failed = 1; break; }
            }
            acc += v[0].x;
        }
        failed = __syncthreads_or(failed);
        if (compute_cycles > 0) { long long c0 = clock64(); while (clock64() - c0 < compute_cycles) {} }
    }
    if (tid == 0) out[blockIdx.x] = clock64() - t_begin;
    if (acc == 0x12345678u) sink[0] = acc;
    cluster.sync();
}

template <int CS>
void run(int nclusters, int steps, int variant, int compute, const char* label) {
    long long* out; unsigned int* fail; uint32_t* sink;
    const int ncta = nclusters * CS;
    CK(cudaMalloc(&out, ncta * 8)); CK(cudaMalloc(&fail, 4)); CK(cudaMemset(fail, 0, 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaFuncSetAttribute(dsmem_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int maxc = 0;
    CK(cudaOccupancyMaxActiveClusters(&maxc, dsmem_kernel<CS>, &cfg));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, dsmem_kernel<CS>, steps, variant, compute, out, fail, sink));
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned int f; CK(cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost));
    long long c0; CK(cudaMemcpy(&c0, out, 8, cudaMemcpyDeviceToHost));
    printf("%-34s CS=%2d clusters=%d (max co-resident %d) compute=%4d : %7.3f us/step (event), %6.0f cycles/step (cta0)%s\n", label, CS,
           nclusters, maxc, compute, ms * 1e3 / steps, (double)c0 / (steps - 16), f ? "  ** TIMEOUT **" : "");
    cudaFree(out); cudaFree(fail); cudaFree(sink);
}

int main() {
    const int steps = 2000;
    const char* vname[] = {"v4 staged, flag-in-data", "v2 direct, flag-in-data", "scalar direct, flag-in-data", "v4 staged, mbarrier arrive"};
    for (int v = 0; v < 4; ++v) run<16>(1, steps, v, 0, vname[v]);
    for (int v = 0; v < 4; ++v) run<16>(4, steps, v, 0, vname[v]);
    for (int v = 0; v < 4; ++v) run<16>(8, steps, v, 0, vname[v]);
    for (int v = 0; v < 4; ++v) run<8>(4, steps, v, 0, vname[v]);
    printf("---- with 500-cycle compute per step\n");
    for (int v = 0; v < 4; ++v) run<16>(4, steps, v, 500, vname[v]);
    return 0;
}
