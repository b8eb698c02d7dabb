// Second exchange micro-benchmark: where do the ~2000 clocks of the 32-CTA all-to-all come from when a 2-CTA
// ping takes 905?  Flag-in-data protocol (st.relaxed.gpu / ld.volatile.v4), G independent groups of N CTAs
// (128 threads), each CTA publishes WPC words per step.
//   mode 0: every CTA polls all N*WPC words of its group (the recurrence kernels' pattern)
//   mode 1: CTA i polls only the words of CTA (i+1)%N  (same dependency depth, fan-in 1)
//   mode 2: as 0 but every thread keeps TWO sweeps in flight (issue sweep B before testing sweep A)
//   mode 3: as 0, the words of a step are spread one per 128-byte line (no line sharing between producers)
//   mode 4: push into private inboxes: every producer stores its WPC words once per consumer, [consumer][producer][WPC];
//           a consumer polls only its own inbox (no line is read by more than one SM)
//   mode 5: mode 4 with two sweeps in flight
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_bench2 tools/exchange_bench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void st_word(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_vec(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 ld4(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool ready(const uint4& v, uint32_t tag) { return v.x == tag && v.y == tag && v.z == tag && v.w == tag; }

template <int NV>
__global__ void __launch_bounds__(128) kern(uint32_t* ring_all, int N, int wpc, int steps, int mode, int vec_store, long long* out,
                                            unsigned int* fail) {
    const int tid = threadIdx.x;
    const int grp = blockIdx.x / N, me = blockIdx.x % N;
    const int stride = (mode == 3) ? 32 : 1;                 // word stride between consecutive vectors' lines (mode 3: 1 vec / line)
    const bool inbox = (mode == 4 || mode == 5);
    const size_t slot_words = (size_t)N * wpc * (mode == 3 ? 8 : 1) * (inbox ? N : 1);
    uint32_t* ring = ring_all + (size_t)grp * 2 * slot_words;
    const int nvec_all = N * wpc / 4, nvec_one = wpc / 4;
    const int nvec = (mode == 1) ? nvec_one : nvec_all;
    const int vbase = (mode == 1) ? ((me + 1) % N) * nvec_one : (inbox ? me * nvec_all : 0);
    long long t_begin = 0;
    unsigned failed = 0;
    auto vaddr = [&](uint32_t* slot, int v) { return slot + (size_t)v * (mode == 3 ? 32 : 4); };
    (void)stride;
    for (int s = 0; s < steps && !failed; ++s) {
        if (s == 16 && tid == 0) t_begin = clock64();
        const uint32_t tag = (uint32_t)(s + 1);
        uint32_t* slot = ring + (size_t)(s & 1) * slot_words;
        if (inbox) {
            // vector v of consumer c: [c][me][v]
            for (int i = tid; i < N * nvec_one; i += 128) st_vec(vaddr(slot, (i / nvec_one) * nvec_all + me * nvec_one + (i % nvec_one)), tag);
        } else if (vec_store) {
            for (int v = tid; v < nvec_one; v += 128) st_vec(vaddr(slot, me * nvec_one + v), tag);
        } else {
            for (int w = tid; w < wpc; w += 128) st_word(vaddr(slot, me * nvec_one + (w >> 2)) + (w & 3), tag);
        }
        uint4 v[NV], v2[NV];
        bool pending = true;
        long long t0 = clock64();
        bool first = true;
        while (pending) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int idx = tid + 128 * i;
                if (idx < nvec && (first || !ready(v[i], tag))) v[i] = ld4(vaddr(slot, vbase + idx));
            }
            if (mode == 2 || mode == 5) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int idx = tid + 128 * i;
                    if (idx < nvec) v2[i] = ld4(vaddr(slot, vbase + idx));
                }
            }
            first = false;
            pending = false;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int idx = tid + 128 * i;
                if (idx < nvec && !ready(v[i], tag)) {
                    if ((mode == 2 || mode == 5) && ready(v2[i], tag)) v[i] = v2[i];
                    else pending = true;
                }
            }
            if (pending && clock64() - t0 > 2000000000LL) { atomicExch(fail, 1u + s); failed = 1; break; }
        }
        failed = __syncthreads_or(failed);
    }
    if (tid == 0) out[blockIdx.x] = clock64() - t_begin;
}

template <int NV>
void run(int G, int N, int wpc, int mode, int vec_store) {
    const int steps = 2000;
    uint32_t* ring; long long* out; unsigned int* fail;
    const size_t words = (size_t)G * 2 * N * wpc * ((mode == 4 || mode == 5) ? N : 8);
    CK(cudaMalloc(&ring, words * 4)); CK(cudaMemset(ring, 0, words * 4));
    CK(cudaMalloc(&out, G * N * 8)); CK(cudaMalloc(&fail, 4)); CK(cudaMemset(fail, 0, 4));
    void* args[] = {&ring, &N, &wpc, (void*)&steps, &mode, &vec_store, &out, &fail};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((const void*)kern<NV>, dim3(G * N), dim3(128), args, 0, 0));
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned int f; CK(cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost));
    long long c0; CK(cudaMemcpy(&c0, out, 8, cudaMemcpyDeviceToHost));
    printf("groups=%d N=%2d wpc=%3d mode=%d %s : %7.3f us/step, %6.0f clk/step%s\n", G, N, wpc, mode, vec_store ? "v4-store" : "b32-store",
           ms * 1e3 / steps, (double)c0 / (steps - 16), f ? "  ** TIMEOUT **" : "");
    cudaFree(ring); cudaFree(out); cudaFree(fail);
}

int main() {
    printf("---- private inboxes (push), one group, fan-in scaling\n");
    run<1>(1, 2, 64, 4, 1); run<1>(1, 8, 64, 4, 1); run<2>(1, 16, 64, 4, 1); run<4>(1, 32, 64, 4, 1);
    printf("---- four groups of 32: shared ring vs private inboxes, 64 / 128 words per CTA\n");
    run<4>(4, 32, 64, 0, 1); run<4>(4, 32, 64, 2, 1); run<4>(4, 32, 64, 4, 1); run<4>(4, 32, 64, 5, 1);
    run<8>(4, 32, 128, 0, 1); run<8>(4, 32, 128, 2, 1); run<8>(4, 32, 128, 4, 1); run<8>(4, 32, 128, 5, 1);
    printf("---- 16 CTAs per group x 128 words, 4 groups\n");
    run<4>(4, 16, 128, 0, 1); run<4>(4, 16, 128, 2, 1); run<4>(4, 16, 128, 4, 1); run<4>(4, 16, 128, 5, 1);
    printf("---- 8 CTAs per group x 256 words, 4 groups\n");
    run<4>(4, 8, 256, 0, 1); run<4>(4, 8, 256, 2, 1); run<4>(4, 8, 256, 4, 1); run<4>(4, 8, 256, 5, 1);
    return 0;
}
