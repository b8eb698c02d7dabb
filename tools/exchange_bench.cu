// Micro-benchmark of the per-step inter-CTA exchange used by the persistent LSTM kernels:
// nCTA co-resident CTAs (128 threads) each publish WPC flagged words per step and gather all
// nCTA*WPC words before the next step.  Variants of the store / poll instructions are timed to
// pick the protocol (results: profiles/r01_lstm_handoff.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_bench tools/exchange_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum StoreMode { ST_RELAXED = 0, ST_VOLATILE = 1, ST_RELAXED_FENCE = 2, ST_ATOM_EXCH = 3, ST_RELEASE = 4, ST_WT = 5 };
enum PollMode { POLL_ALL = 0, POLL_CANARY = 1, POLL_ALL_VOLATILE = 2, POLL_ALL_BACKOFF = 3 };

__device__ __forceinline__ void store_word(uint32_t* p, uint32_t v, int mode) {
    switch (mode) {
        case ST_RELAXED: asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
        case ST_VOLATILE: asm volatile("st.volatile.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
        case ST_RELAXED_FENCE:
            asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
            __threadfence();
            break;
        case ST_ATOM_EXCH: atomicExch(p, v); break;
        case ST_RELEASE: asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
        case ST_WT: asm volatile("st.global.wt.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); break;
    }
}
__device__ __forceinline__ uint4 load4(const uint32_t* p, bool vol) {
    uint4 v;
    if (vol) asm volatile("ld.volatile.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool ready(const uint4& v, uint32_t tag) { return v.x == tag && v.y == tag && v.z == tag && v.w == tag; }

// NV = vectors gathered per thread per step (nCTA * WPC / 4 / 128)
template <int NV>
__global__ void __launch_bounds__(128) exchange_kernel(uint32_t* ring, int wpc, int steps, int store_mode, int poll_mode,
                                                       int compute_cycles, long long* out, unsigned int* fail) {
    const int tid = threadIdx.x;
    const int words = gridDim.x * wpc;  // words per slot
    long long t_begin = 0;
    for (int s = 0; s < steps; ++s) {
        if (s == 16 && tid == 0) t_begin = clock64();
        const uint32_t tag = (uint32_t)(s + 1);
        uint32_t* slot = ring + (size_t)(s & 1) * words;
        for (int w = tid; w < wpc; w += 128) store_word(slot + blockIdx.x * wpc + w, tag, store_mode);
        // gather
        uint4 v[NV];
        const bool vol = poll_mode == POLL_ALL_VOLATILE;
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = load4(slot + (size_t)(tid + 128 * i) * 4, vol);
        bool pending = false;
#pragma unroll
        for (int i = 0; i < NV; ++i) if (!ready(v[i], tag)) pending = true;
        long long t0 = clock64();
        if (pending && poll_mode == POLL_CANARY) {
            while (!ready(v[0], tag)) {
                v[0] = load4(slot + (size_t)tid * 4, vol);
                if (clock64() - t0 > 2000000000LL) { atomicExch(fail, 1u + s); break; }
            }
        }
        while (pending) {
            pending = false;
            if (poll_mode == POLL_ALL_BACKOFF) __nanosleep(100);
#pragma unroll
            for (int i = 0; i < NV; ++i) if (!ready(v[i], tag)) v[i] = load4(slot + (size_t)(tid + 128 * i) * 4, vol);
#pragma unroll
            for (int i = 0; i < NV; ++i) if (!ready(v[i], tag)) pending = true;
            if (pending && clock64() - t0 > 2000000000LL) { atomicExch(fail, 1u + s); break; }
        }
        __syncthreads();
        if (compute_cycles > 0) { long long c0 = clock64(); while (clock64() - c0 < compute_cycles) {} }
    }
    if (tid == 0) out[blockIdx.x] = clock64() - t_begin;
}

template <int NV>
void run(int ncta, int wpc, int steps, int sm, int pm, int compute, const char* label) {
    uint32_t* ring; long long* out; unsigned int* fail;
    size_t words = (size_t)ncta * wpc;
    CK(cudaMalloc(&ring, 2 * words * 4)); CK(cudaMemset(ring, 0, 2 * words * 4));
    CK(cudaMalloc(&out, ncta * 8)); CK(cudaMalloc(&fail, 4)); CK(cudaMemset(fail, 0, 4));
    void* args[] = {&ring, &wpc, &steps, &sm, &pm, &compute, &out, &fail};
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((const void*)exchange_kernel<NV>, dim3(ncta), dim3(128), args, 0, 0));
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned int f; CK(cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost));
    long long c0; CK(cudaMemcpy(&c0, out, 8, cudaMemcpyDeviceToHost));
    printf("%-52s ncta=%3d wpc=%3d NV=%d compute=%4d : %7.3f us/step (event), %6.0f cycles/step (cta0)%s\n", label, ncta, wpc, NV,
           compute, ms * 1e3 / steps, (double)c0 / (steps - 16), f ? "  ** TIMEOUT **" : "");
    cudaFree(ring); cudaFree(out); cudaFree(fail);
}

int main() {
    const int steps = 2000;
    const char* sname[] = {"st.relaxed.gpu", "st.volatile", "st.relaxed+threadfence", "atomicExch", "st.release.gpu", "st.wt"};
    const char* pname[] = {"poll-all", "canary", "poll-all-volatile", "poll-all-backoff100ns"};
    char label[128];
    // fwd H=256-like: 32 CTAs per group (x4 groups -> emulate with 128 CTAs all-to-all is too much; use one group)
    for (int sm = 0; sm < 6; ++sm)
        for (int pm = 0; pm < 4; ++pm) {
            snprintf(label, sizeof label, "%s / %s", sname[sm], pname[pm]);
            run<4>(32, 64, steps, sm, pm, 0, label);   // 32 CTAs x 64 words = 2048 words = 512 vectors = 4 per thread
        }
    printf("---- with 500-cycle compute per step\n");
    for (int sm = 0; sm < 6; ++sm) { snprintf(label, sizeof label, "%s / poll-all", sname[sm]); run<4>(32, 64, steps, sm, 0, 500, label); }
    printf("---- fwd H=512-like group: 64 CTAs x 64 words (8 vectors/thread)\n");
    for (int sm = 0; sm < 6; ++sm)
        for (int pm = 0; pm < 2; ++pm) { snprintf(label, sizeof label, "%s / %s", sname[sm], pname[pm]); run<8>(64, 64, steps, sm, pm, 0, label); }
    printf("---- bwd H=256-like group: 32 CTAs x 256 words (16 vectors/thread)\n");
    for (int sm = 0; sm < 6; ++sm)
        for (int pm = 0; pm < 2; ++pm) { snprintf(label, sizeof label, "%s / %s", sname[sm], pname[pm]); run<16>(32, 256, steps, sm, pm, 0, label); }
    printf("---- 2-CTA ping (1 word each)\n");
    for (int sm = 0; sm < 6; ++sm) { snprintf(label, sizeof label, "%s / poll-all", sname[sm]); run<1>(2, 256, steps, sm, 0, 0, label); }
    return 0;
}
