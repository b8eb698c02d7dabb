// Probe of the legacy tensor path (mma.sync.m16n8k16, HMMA) on sm_100a for the LSTM recurrent matvec
// [4U gate rows x K] . [K x 8 videos]:
//   1. issue rate: MMAs per clock per SM with 1..4 independent accumulator chains per warp, 8 and 16 warps per SM;
//   2. accuracy of split-precision products against fp64 at K = 512 for operands shaped like the recurrence
//      (w ~ U(-1/sqrt(K), 1/sqrt(K)) * scale, h in (-1, 1)):  fp16x3 (hi.hi + hi.lo + lo.hi, weights pre-scaled by 2^s),
//      bf16x3, single chain vs small terms in their own accumulator, against plain fp32 FMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe tools/hmma_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <bool BF>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    if (BF)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void rate_kernel(int iters, float* out, long long* cyc) {
    uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
    uint32_t b[2] = {0x3c003c00u, 0x38003800u};
    float d[CHAINS][4];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) mma16816<false>(d[c], a, b);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
void rate(int warps) {
    float* out; long long* cyc;
    const int iters = 2000;
    CK(cudaMalloc(&out, 148 * warps * 32 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
    rate_kernel<CHAINS><<<148, warps * 32>>>(iters, out, cyc);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    const double mmas = (double)iters * 8 * CHAINS * warps;
    printf("rate: %2d warps/SM, %d chains/warp: %.3f MMA(m16n8k16)/clk/SM = %.0f dense FLOP/clk/SM; %.1f clk per dependent MMA\n", warps, CHAINS,
           mmas / c, mmas / c * 4096, (double)c / (iters * 8));
    cudaFree(out); cudaFree(cyc);
}

// ---- accuracy ------------------------------------------------------------------------
// One warp computes D[16 rows x 8 videos] = W[16 x K] . H[K x 8] with different schemes.
// mode 0: fp16x3 single chain; 1: fp16x3, hi.hi chain + (hi.lo, lo.hi) chain; 2: bf16x3 single; 3: bf16x3 two chains;
// mode 4: fp32 FMA (k ascending); 5: fp16x3 two chains with K split into 4 sub-chains summed in fp32
template <typename T> __device__ __forceinline__ T cvt(float x);
template <> __device__ __forceinline__ __half cvt<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <typename T> __device__ __forceinline__ float back(T x);
template <> __device__ __forceinline__ float back<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float back<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T>
__device__ __forceinline__ void split_pair(float x0, float x1, float scale, uint32_t& hi, uint32_t& lo) {
    const float s0 = x0 * scale, s1 = x1 * scale;
    const T h0 = cvt<T>(s0), h1 = cvt<T>(s1);
    const T l0 = cvt<T>(s0 - back<T>(h0)), l1 = cvt<T>(s1 - back<T>(h1));
    const unsigned short* p;
    p = reinterpret_cast<const unsigned short*>(&h0); uint32_t a = *p;
    p = reinterpret_cast<const unsigned short*>(&h1); uint32_t b = *p;
    hi = a | (b << 16);
    p = reinterpret_cast<const unsigned short*>(&l0); a = *p;
    p = reinterpret_cast<const unsigned short*>(&l1); b = *p;
    lo = a | (b << 16);
}

template <typename T, bool BF>
__device__ void split_dot(const float* W, const float* Hm, int K, float wscale, int mode, float* D) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float d_main[4] = {0, 0, 0, 0}, d_small[4] = {0, 0, 0, 0}, total[4] = {0, 0, 0, 0};
    const bool two = (mode & 1) || mode == 5;
    const int sub = (mode == 5) ? K / 4 : K;
    for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t ah[4], al[4], bh[2], bl[2];
        // A: a0 (row g, k 2t..), a1 (row g+8, k 2t..), a2 (row g, k 2t+8..), a3 (row g+8, k 2t+8..)
        split_pair<T>(W[g * K + k0 + 2 * t], W[g * K + k0 + 2 * t + 1], wscale, ah[0], al[0]);
        split_pair<T>(W[(g + 8) * K + k0 + 2 * t], W[(g + 8) * K + k0 + 2 * t + 1], wscale, ah[1], al[1]);
        split_pair<T>(W[g * K + k0 + 2 * t + 8], W[g * K + k0 + 2 * t + 9], wscale, ah[2], al[2]);
        split_pair<T>(W[(g + 8) * K + k0 + 2 * t + 8], W[(g + 8) * K + k0 + 2 * t + 9], wscale, ah[3], al[3]);
        // B: b0 (k 2t.., col g), b1 (k 2t+8.., col g);  Hm is [8 videos][K]
        split_pair<T>(Hm[g * K + k0 + 2 * t], Hm[g * K + k0 + 2 * t + 1], 1.0f, bh[0], bl[0]);
        split_pair<T>(Hm[g * K + k0 + 2 * t + 8], Hm[g * K + k0 + 2 * t + 9], 1.0f, bh[1], bl[1]);
        if (two) {
            mma16816<BF>(d_small, ah, bl);
            mma16816<BF>(d_small, al, bh);
            mma16816<BF>(d_main, ah, bh);
        } else {
            mma16816<BF>(d_main, ah, bl);
            mma16816<BF>(d_main, al, bh);
            mma16816<BF>(d_main, ah, bh);
        }
        if ((k0 + 16) % sub == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { total[i] += d_main[i] + d_small[i]; d_main[i] = 0; d_small[i] = 0; }
        }
    }
    const float inv = 1.0f / wscale;
    D[g * 8 + 2 * t] = total[0] * inv;
    D[g * 8 + 2 * t + 1] = total[1] * inv;
    D[(g + 8) * 8 + 2 * t] = total[2] * inv;
    D[(g + 8) * 8 + 2 * t + 1] = total[3] * inv;
}

__global__ void acc_kernel(const float* W, const float* Hm, int K, float wscale, int mode, float* D) {
    W += (size_t)blockIdx.x * 16 * K;
    D += (size_t)blockIdx.x * 128;
    if (mode == 4) {
        for (int idx = threadIdx.x; idx < 128; idx += 32) {
            const int r = idx >> 3, b = idx & 7;
            float a = 0.f;
            for (int k = 0; k < K; ++k) a = fmaf(W[r * K + k], Hm[b * K + k], a);
            D[idx] = a;
        }
    } else if (mode == 0 || mode == 1 || mode == 5) {
        split_dot<__half, false>(W, Hm, K, wscale, mode, D);
    } else {
        split_dot<__nv_bfloat16, true>(W, Hm, K, 1.0f, mode, D);
    }
}

int main() {
    rate<1>(8); rate<2>(8); rate<4>(8); rate<1>(16); rate<2>(16); rate<4>(16); rate<4>(4);

    const int K = 512, TILES = 64;
    for (int pass = 0; pass < 2; ++pass) {
        const float wmag = (pass == 0) ? 1.0f / sqrtf((float)K) : 6.0f / sqrtf((float)K);
        std::vector<float> W((size_t)TILES * 16 * K), Hm(8 * K);
        srand(1234 + pass);
        for (auto& w : W) w = wmag * (2.0f * rand() / RAND_MAX - 1.0f);
        for (auto& h : Hm) h = tanhf(1.5f * (2.0f * rand() / RAND_MAX - 1.0f));
        std::vector<double> ref((size_t)TILES * 128);
        double refmax = 0;
        for (int tile = 0; tile < TILES; ++tile)
            for (int r = 0; r < 16; ++r)
                for (int b = 0; b < 8; ++b) {
                    double a = 0;
                    for (int k = 0; k < K; ++k) a += (double)W[((size_t)tile * 16 + r) * K + k] * (double)Hm[b * K + k];
                    ref[(size_t)tile * 128 + r * 8 + b] = a;
                    refmax = fmax(refmax, fabs(a));
                }
        float *dW, *dH, *dD;
        CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dH, Hm.size() * 4)); CK(cudaMalloc(&dD, ref.size() * 4));
        CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dH, Hm.data(), Hm.size() * 4, cudaMemcpyHostToDevice));
        // power-of-two scale bringing max|w| to about 2^9 (fp16 max is 65504; lo parts stay normal)
        const float wscale = exp2f(floorf(log2f(512.0f / wmag)));
        const char* names[] = {"fp16x3 one chain", "fp16x3 main+small chains", "bf16x3 one chain", "bf16x3 main+small chains", "fp32 FMA", "fp16x3 two chains, 4 K-sub-chains"};
        printf("accuracy: K=%d, |w| <= %.4f, wscale 2^%d, max |ref| %.3f\n", K, wmag, (int)log2f(wscale), refmax);
        for (int mode = 0; mode < 6; ++mode) {
            acc_kernel<<<TILES, 32>>>(dW, dH, K, wscale, mode, dD);
            CK(cudaDeviceSynchronize());
            std::vector<float> D(ref.size());
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double emax = 0, esum = 0, bias = 0;
            for (size_t i = 0; i < ref.size(); ++i) { const double e = D[i] - ref[i]; emax = fmax(emax, fabs(e)); esum += e * e; bias += e; }
            printf("  %-36s max abs err %.3e  rms %.3e  mean %+.3e\n", names[mode], emax, sqrt(esum / ref.size()), bias / ref.size());
        }
        cudaFree(dW); cudaFree(dH); cudaFree(dD);
    }
    return 0;
}
