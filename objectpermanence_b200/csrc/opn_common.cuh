// Common device/host helpers for libopnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/opnet_b200.h"

namespace opn {

// ---- host-side error plumbing ------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
// the status words of a persistent launch: the registered sticky page of the current device, else the workspace head
unsigned int* status_page_or(void* workspace_head);

#define OPN_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            opn::set_error(__VA_ARGS__);         \
            return OPN_ERR_BAD_ARG;              \
        }                                        \
    } while (0)

#define OPN_CUDA(call)                                          \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return opn::cuda_fail(e__, #call); \
    } while (0)

// OPNet forward / backward as two concurrent kernels (the recurrence of LSTM2 on 128 CTAs, LSTM1 + who-to-track on the idle
// SMs: opn_opnet_l1head.cu, opn_opnet_l1bwd.cu): whether this device / environment allows it for a batch, and the library's
// side stream with its fork / join events (one set per device, created on first use)
bool opnet_split_wanted(int64_t B);
int opnet_split_groups_per_wave();
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    std::mutex enqueue;      // held from the fork record to the join wait: host threads sharing a device take turns
};
SideStream* opnet_side_stream();

// counts kernels launched by this library (bench.py reports it as gpu_launches)
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += (unsigned long long)n; }

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- device-side PTX wrappers --------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier (shared::cta)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make barrier initialisation visible to the async proxy (TMA unit)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// order generic-proxy accesses against async-proxy (TMA / tensor-core) accesses.  The state-space qualified
// forms matter: plain fence.proxy.async compiles to MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC (checked with
// cuobjdump -sass), the .shared::cta form to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC.S.
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

#endif  // __CUDACC__

}  // namespace opn
