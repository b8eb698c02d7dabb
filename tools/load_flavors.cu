// Raw L2 -> SM throughput of the load flavours that could carry the per-step exchange of the
// persistent LSTM kernels.  One CTA per SM re-reads an L2-resident 64 KB buffer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/load_flavors tools/load_flavors.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__device__ __forceinline__ uint4 ld(const uint4* p) {
    uint4 v;
    if (MODE == 0) asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (MODE == 1) asm volatile("ld.volatile.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (MODE == 2) asm volatile("ld.global.cg.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (MODE == 3) asm volatile("ld.global.ca.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (MODE == 4) asm volatile("ld.global.cv.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    if (MODE == 5) asm volatile("ld.acquire.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// register loads: each thread reads NV vectors per pass (all in flight), `passes` dependent passes
template <int MODE, int NV>
__global__ void reg_kernel(const uint4* buf, int vec_per_cta, int passes, long long* out, uint32_t* sink) {
    const uint4* base = buf + (size_t)blockIdx.x * vec_per_cta;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int p = 0; p < passes; ++p) {
        uint4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = ld<MODE>(base + ((threadIdx.x + blockDim.x * i + (acc & 1)) % vec_per_cta));
#pragma unroll
        for (int i = 0; i < NV; ++i) acc += v[i].x ^ v[i].w;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345) sink[0] = acc;
}

// cp.async (LDGSTS) 16 B into shared memory
template <int NV>
__global__ void cpasync_kernel(const uint4* buf, int vec_per_cta, int passes, long long* out, uint32_t* sink) {
    extern __shared__ uint4 sm[];
    const uint4* base = buf + (size_t)blockIdx.x * vec_per_cta;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int p = 0; p < passes; ++p) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int idx = threadIdx.x + blockDim.x * i;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&sm[idx])), "l"(base + idx) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        acc += sm[threadIdx.x].x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345) sink[0] = acc;
}

// TMA 1-D bulk copies: one thread issues `pieces` copies of bytes/pieces each
__global__ void bulk_kernel(const uint4* buf, int vec_per_cta, int passes, int pieces, long long* out, uint32_t* sink) {
    extern __shared__ __align__(128) uint4 sm[];
    __shared__ __align__(8) uint64_t bar;
    const uint4* base = buf + (size_t)blockIdx.x * vec_per_cta;
    const uint32_t bytes = vec_per_cta * 16;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t acc = 0, parity = 0;
    long long t0 = clock64();
    for (int p = 0; p < passes; ++p) {
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
            __syncwarp();
            const uint32_t pb = bytes / pieces;
            if (threadIdx.x < pieces)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32((char*)sm + threadIdx.x * pb)), "l"((const char*)base + threadIdx.x * pb), "r"(pb), "r"(smem_u32(&bar)) : "memory");
        }
        uint32_t ok = 0;
        long long w0 = clock64();
        while (!ok && clock64() - w0 < 400000000LL) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
        parity ^= 1;
        acc += sm[threadIdx.x].x;
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345) sink[0] = acc;
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int passes = 200;
    const char* names[] = {"ld.relaxed.gpu.v4", "ld.volatile.v4", "ld.global.cg.v4", "ld.global.ca.v4", "ld.global.cv.v4", "ld.acquire.gpu.v4"};
    long long* out; uint32_t* sink; CK(cudaMalloc(&out, sms * 2 * 8)); CK(cudaMalloc(&sink, 4));
    for (int kb : {16, 64}) {
        const int vec = kb * 1024 / 16;
        for (int ctas_per_sm : {1, 2}) {
            const int grid = sms * ctas_per_sm;
            uint4* buf; CK(cudaMalloc(&buf, (size_t)grid * vec * 16)); CK(cudaMemset(buf, 1, (size_t)grid * vec * 16));
            auto report = [&](const char* name, int threads) {
                CK(cudaDeviceSynchronize());
                long long c; CK(cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost));
                printf("%2d KB/CTA %d CTA/SM %3d thr  %-22s : %7.1f cycles/pass  %6.1f B/clk/CTA\n", kb, ctas_per_sm, threads, name,
                       (double)c / passes, (double)kb * 1024 * passes / c);
                fflush(stdout);
            };
#define RUNREG(MODE) { constexpr int NVR = 8; int threads = vec / NVR; if (threads > 1024) threads = 1024; \
            reg_kernel<MODE, NVR><<<grid, threads>>>(buf, vec, passes, out, sink); report(names[MODE], threads); }
            RUNREG(0) RUNREG(1) RUNREG(2) RUNREG(3) RUNREG(4) RUNREG(5)
            { constexpr int NVR = 8; int threads = vec / NVR; if (threads > 1024) threads = 1024;
              CK(cudaFuncSetAttribute(cpasync_kernel<NVR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
              cpasync_kernel<NVR><<<grid, threads, kb * 1024>>>(buf, vec, passes, out, sink); report("cp.async.cg 16B", threads); }
            for (int pieces : {1, 8, 32}) {
                CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
                bulk_kernel<<<grid, 128, kb * 1024>>>(buf, vec, passes, pieces, out, sink);
                char nm[64]; snprintf(nm, sizeof nm, "cp.async.bulk x%d", pieces); report(nm, 128);
            }
            cudaFree(buf);
        }
    }
    return 0;
}
