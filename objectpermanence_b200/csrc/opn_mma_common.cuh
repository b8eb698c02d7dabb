// Device helpers shared by the tensor-core (mma.sync) recurrence kernels: opn_lstm_mma.cu and opn_opnet_fused.cu.
#pragma once
#include <cuda_fp16.h>

#include "opn_lstm_common.cuh"

// Optional per-phase cycle accounting of the step loop (development builds: -DOPN_LSTM_PHASES, read back with
// tools/lstm_phases.py): thread 0 of CTA 0 and CTA 77 sums clock64() deltas between the PH(i) marks into
// status words 64.. / 128..
#ifdef OPN_LSTM_PHASES
#define PH_DECL long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long ph_last = clock64();
#define PH(i)                                   \
    do {                                        \
        const long long now__ = clock64();      \
        ph_acc[i] += now__ - ph_last;           \
        ph_last = now__;                        \
    } while (0)
#define PH_COUNT(i, n) ph_acc[i] += (n);
#define PH_STORE(status)                                                                             \
    do {                                                                                             \
        if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 77)) {                             \
            unsigned long long* o = reinterpret_cast<unsigned long long*>(status) + (blockIdx.x ? 64 : 32); \
            for (int i = 0; i < 8; ++i) o[i] = (unsigned long long)ph_acc[i];                        \
        }                                                                                            \
    } while (0)
#else
#define PH_DECL
#define PH(i)
#define PH_COUNT(i, n)
#define PH_STORE(status)
#endif

namespace opn {

namespace {

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x0, x1) -> packed fp16 pairs hi, lo with x ~= hi + lo; x0 in the low half (the lower k index of a fragment word)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the SFU (ex2.approx + rcp.approx): absolute error < 3e-7 over the whole range,
// saturates correctly at +-1 (exp -> inf gives 2/inf = 0).  sigmoid(x) = 0.5 + 0.5 tanh(x/2), so the three gate
// non-linearities of a lane pair are one branch-free formula: act = s * tanh(s * a) + o with (s, o) = (1, 0) or
// (0.5, 0.5).  (The accurate expf / tanhf sequence of the FP32-FMA kernels cost ~550 clocks per step here.)
__device__ __forceinline__ float tanh_sfu(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// power-of-two scale bringing `amax` into [2^(target-1), 2^target); 1 for amax == 0
__device__ __forceinline__ void pow2_scale(float amax, int target, float& scale, float& inv) {
    int eb = (int)((__float_as_uint(amax) >> 23) & 0xffu);  // biased exponent: amax in [2^(eb-127), 2^(eb-126))
    eb = max(eb, target + 2);                                 // keeps both exponent fields in [1, 254]
    eb = min(eb, 254);
    // scale = 2^(target - 1 - (eb - 127)), inv = 2^((eb - 127) - (target - 1))
    scale = __uint_as_float((uint32_t)(127 + target - 1 - (eb - 127)) << 23);
    inv = __uint_as_float((uint32_t)(127 + (eb - 127) - (target - 1)) << 23);
}

// word index of (video b, hidden index k) in an operand tile laid out in B-fragment order
__device__ __forceinline__ int frag_word(int b, int k) {
    const int ks = k >> 4, r = k & 15;
    return 4 * (ks * 32 + b * 4 + ((r & 7) >> 1)) + 2 * (r >> 3) + (r & 1);
}

// Block-wide max of |W| over the CTA's slice -> power-of-two weight scale (same value in every thread)
template <int NW>
__device__ __forceinline__ void weight_scale(float local_max, float* red_s, float& scale, float& inv) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if (lane == 0) red_s[warp] = local_max;
    __syncthreads();
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < NW; ++i) m = fmaxf(m, red_s[i]);
    pow2_scale(m, 12, scale, inv);
    __syncthreads();
}

}  // namespace
}  // namespace opn
