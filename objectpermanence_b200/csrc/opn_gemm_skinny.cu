// Skinny fp32 contractions: one operand is a [9600 x 512 .. 2048] activation / gradient tensor streamed from HBM
// exactly once, the other at most 16 wide.  These are the products around the bbox head, the who-to-track head and
// the K = 6 input of OPNet's second LSTM (baselines/learned_models.py:40,46,47 and their autograd backward); as
// 128 x 16 register tiles they used 16-75 CTAs and reached 0.4-1.4 TB/s, 194 us of the 3.04 ms OPNet step.
//
//   rowdot  : C[m][n] = sum_k A[m][k] * Bop[k][n],  N <= 8, A rows contiguous in k.   One warp per row, lanes stride k with
//             float4 loads, the small operand staged in shared memory as [n][K]; warp-shuffle reduction.
//   colred  : C = S^T L with S [K x W] (W <= 16), L [K x J] (J large, contiguous):  one thread per column j of L, W
//             accumulators, K split over the grid (fp32 atomics into a zeroed C).  Covers both dW = dy^T x (W = M) and
//             dW = dgates^T x (W = N) through output strides.
//   tinyk   : C[m][n] = sum_{k < 8} A[m][k] * Bop[k][n]: every thread produces 4 consecutive n of one row (16-byte
//             stores), the small operand in shared memory.
// All return false when the call does not fit (alignment, sizes): the caller falls through to the tiled kernel.
#include <stdlib.h>

#include "opn_common.cuh"

namespace opn {
namespace {

constexpr int kSkinnyThreads = 256;

// ---- rowdot ----------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(kSkinnyThreads) rowdot_kernel(const float* __restrict__ A, long long lda,
                                                                const float* __restrict__ B, long long ldb, int tb,
                                                                float* __restrict__ C, long long ldc, int M, int K,
                                                                float alpha, int beta_one) {
    extern __shared__ __align__(16) float bs[];   // [N][K]
    if (tb) {   // B[n][k]: rows of K contiguous floats
        for (int i = threadIdx.x; i < N * K; i += kSkinnyThreads) bs[i] = B[(long long)(i / K) * ldb + i % K];
    } else {    // B[k][n]: read along the (short) rows, scatter into [n][K]
        for (int i = threadIdx.x; i < N * K; i += kSkinnyThreads) bs[(i % N) * K + i / N] = B[(long long)(i / N) * ldb + i % N];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps = kSkinnyThreads / 32;
    for (long long m = (long long)blockIdx.x * warps + warp; m < M; m += (long long)gridDim.x * warps) {
        const float4* a4 = reinterpret_cast<const float4*>(A + m * lda);
        float acc[N];
#pragma unroll
        for (int n = 0; n < N; ++n) acc[n] = 0.0f;
#pragma unroll 8
        for (int k4 = lane; k4 < K / 4; k4 += 32) {
            const float4 a = __ldg(a4 + k4);
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float4 b = *reinterpret_cast<const float4*>(bs + n * K + 4 * k4);
                acc[n] = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc[n]))));
            }
        }
#pragma unroll
        for (int n = 0; n < N; ++n) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int n = 0; n < N; ++n) {
                float* c = C + m * ldc + n;
                *c = alpha * acc[n] + (beta_one ? *c : 0.0f);
            }
        }
    }
}

// ---- colred ----------------------------------------------------------------------------------------------------
// C[w, j] += alpha * sum_{k in chunk} S[k][w] * L[k][j];  element (w, j) of C at C[w*cs_w + j*cs_j].
// Every thread owns 4 consecutive columns j (one 16-byte load per row of L) and 4*W accumulators.
template <int W, int U>
__global__ void __launch_bounds__(kSkinnyThreads) colred_kernel(const float* __restrict__ S, long long lds,
                                                                const float* __restrict__ L, long long ldl,
                                                                float* __restrict__ C, long long cs_w, long long cs_j,
                                                                int J, int K, int k_per_cta, float alpha) {
    constexpr int KC = 64;                         // rows of S staged per pass
    __shared__ float ss[KC][W];
    const int j = 4 * (blockIdx.x * kSkinnyThreads + threadIdx.x);
    const int k_begin = blockIdx.y * k_per_cta, k_end = min(K, k_begin + k_per_cta);
    float acc[W][4];
#pragma unroll
    for (int w = 0; w < W; ++w) acc[w][0] = acc[w][1] = acc[w][2] = acc[w][3] = 0.0f;
    for (int k0 = k_begin; k0 < k_end; k0 += KC) {
        const int kc = min(KC, k_end - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kc * W; i += kSkinnyThreads) ss[i / W][i % W] = __ldg(S + (long long)(k0 + i / W) * lds + i % W);
        __syncthreads();
        if (j < J) {
            const float* lp = L + (long long)k0 * ldl + j;
            int kk = 0;
            for (; kk + U <= kc; kk += U) {   // U rows (16 bytes each per thread) in flight: the kernel is a pure stream
                float4 l[U];                  // over L and ran at 1.5 TB/s with U = 4, one CTA per SM
#pragma unroll
                for (int u = 0; u < U; ++u) l[u] = __ldg(reinterpret_cast<const float4*>(lp + (long long)(kk + u) * ldl));
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const float sv = ss[kk + u][w];
                        acc[w][0] = fmaf(l[u].x, sv, acc[w][0]);
                        acc[w][1] = fmaf(l[u].y, sv, acc[w][1]);
                        acc[w][2] = fmaf(l[u].z, sv, acc[w][2]);
                        acc[w][3] = fmaf(l[u].w, sv, acc[w][3]);
                    }
            }
            for (; kk < kc; ++kk) {
                const float4 l0 = __ldg(reinterpret_cast<const float4*>(lp + (long long)kk * ldl));
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    const float sv = ss[kk][w];
                    acc[w][0] = fmaf(l0.x, sv, acc[w][0]);
                    acc[w][1] = fmaf(l0.y, sv, acc[w][1]);
                    acc[w][2] = fmaf(l0.z, sv, acc[w][2]);
                    acc[w][3] = fmaf(l0.w, sv, acc[w][3]);
                }
            }
        }
    }
    if (j < J) {
#pragma unroll
        for (int w = 0; w < W; ++w)
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(C + w * cs_w + (j + c) * cs_j, alpha * acc[w][c]);
    }
}

__global__ void zero_strided2_kernel(float* C, long long s0, long long s1, int n0, int n1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)n0 * n1) C[(i / n1) * s0 + (i % n1) * s1] = 0.0f;
}

// ---- tinyk -----------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kSkinnyThreads) tinyk_kernel(const float* __restrict__ A, long long lda,
                                                               const float* __restrict__ B, long long ldb, int tb,
                                                               float* __restrict__ C, long long ldc, int M, int N,
                                                               float alpha, int beta_one) {
    extern __shared__ __align__(16) float bs[];   // [K][N]
    for (int i = threadIdx.x; i < K * N; i += kSkinnyThreads) {
        const int k = i / N, n = i % N;
        bs[i] = tb ? B[(long long)n * ldb + k] : B[(long long)k * ldb + n];
    }
    __syncthreads();
    const int n4 = N / 4;
    const long long total = (long long)M * n4;
    for (long long idx = (long long)blockIdx.x * kSkinnyThreads + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * kSkinnyThreads) {
        const long long m = idx / n4;
        const int c4 = (int)(idx % n4);
        float a[K];
#pragma unroll
        for (int k = 0; k < K; ++k) a[k] = __ldg(A + m * lda + k);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4 b = *reinterpret_cast<const float4*>(bs + k * N + 4 * c4);
            acc.x = fmaf(a[k], b.x, acc.x);
            acc.y = fmaf(a[k], b.y, acc.y);
            acc.z = fmaf(a[k], b.z, acc.z);
            acc.w = fmaf(a[k], b.w, acc.w);
        }
        float4* c = reinterpret_cast<float4*>(C + m * ldc + 4 * c4);
        float4 out = make_float4(alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w);
        if (beta_one) {
            const float4 old = *c;
            out.x += old.x, out.y += old.y, out.z += old.z, out.w += old.w;
        }
        *c = out;
    }
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <int N>
int launch_rowdot(const float* A, long long lda, const float* B, long long ldb, int tb, float* C, long long ldc, int M, int K,
                  float alpha, int beta_one, cudaStream_t s) {
    const size_t smem = (size_t)N * K * sizeof(float);
    OPN_CUDA(cudaFuncSetAttribute(rowdot_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (M + 7) / 8;
    if (grid > 148 * 2) grid = 148 * 2;   // the staging of B (up to 48 KB per CTA) is amortised over ~32 rows
    rowdot_kernel<N><<<(unsigned)grid, kSkinnyThreads, smem, s>>>(A, lda, B, ldb, tb, C, ldc, M, K, alpha, beta_one);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

template <int W>
int launch_colred(const float* S, long long lds, const float* L, long long ldl, float* C, long long cs_w, long long cs_j, int J,
                  int K, float alpha, int beta_one, cudaStream_t s) {
    if (!beta_one) {
        const long long n = (long long)W * J;
        zero_strided2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(C, cs_w, cs_j, W, J);
        OPN_CUDA(cudaGetLastError());
        count_launch();
    }
    static int legacy = -1;   // OPN_COLRED_LEGACY=1: four rows in flight, about one CTA per SM (the first version; kept for A/B)
    if (legacy < 0) {
        const char* e = getenv("OPN_COLRED_LEGACY");
        legacy = (e && e[0] == '1') ? 1 : 0;
    }
    const int gx = (J / 4 + kSkinnyThreads - 1) / kSkinnyThreads;
    int splits = (148 * (legacy ? 2 : 4) + gx - 1) / gx;   // up to four CTAs per SM in total (k_per_cta >= 64 caps it)
    int k_per_cta = (K + splits - 1) / splits;
    k_per_cta = (k_per_cta + 63) / 64 * 64;
    if (k_per_cta < 64) k_per_cta = 64;
    splits = (K + k_per_cta - 1) / k_per_cta;
    const dim3 grid((unsigned)gx, (unsigned)splits);
    if (legacy)
        colred_kernel<W, 4><<<grid, kSkinnyThreads, 0, s>>>(S, lds, L, ldl, C, cs_w, cs_j, J, K, k_per_cta, alpha);
    else
        colred_kernel<W, 8><<<grid, kSkinnyThreads, 0, s>>>(S, lds, L, ldl, C, cs_w, cs_j, J, K, k_per_cta, alpha);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

template <int K>
int launch_tinyk(const float* A, long long lda, const float* B, long long ldb, int tb, float* C, long long ldc, int M, int N,
                 float alpha, int beta_one, cudaStream_t s) {
    const size_t smem = (size_t)K * N * sizeof(float);
    OPN_CUDA(cudaFuncSetAttribute(tinyk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = ((long long)M * (N / 4) + kSkinnyThreads - 1) / kSkinnyThreads;
    if (grid > 148 * 8) grid = 148 * 8;
    tinyk_kernel<K><<<(unsigned)grid, kSkinnyThreads, smem, s>>>(A, lda, B, ldb, tb, C, ldc, M, N, alpha, beta_one);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

}  // namespace

// Tries the skinny kernels.  *handled = true when one was launched (return code = its status).
int gemm_skinny(bool ta, bool tb, long long M, long long N, long long K, float alpha, const float* A, long long lda,
                const float* B, long long ldb, float beta, float* C, long long ldc, cudaStream_t s, bool* handled) {
    *handled = false;
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("OPN_GEMM_SKINNY");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    if (!enabled || M <= 0 || N <= 0 || K <= 0) return OPN_OK;
    const int beta_one = beta == 1.0f;
#define OPN_SKINNY_CASES(MACRO) MACRO(1) MACRO(2) MACRO(3) MACRO(4) MACRO(5) MACRO(6) MACRO(7) MACRO(8)
    // rowdot: large M, N <= 8, long contiguous rows of A
    if (!ta && N <= 8 && M >= 512 && K >= 128 && K % 4 == 0 && lda % 4 == 0 && aligned16(A) && N * K * 4 <= 96 * 1024) {
        *handled = true;
        switch ((int)N) {
#define OPN_CASE(n) case n: return launch_rowdot<n>(A, lda, B, ldb, tb ? 1 : 0, C, ldc, (int)M, (int)K, alpha, beta_one, s);
            OPN_SKINNY_CASES(OPN_CASE)
#undef OPN_CASE
        }
    }
    // tinyk: K <= 8, wide N
    if (!ta && K <= 8 && N >= 64 && N % 4 == 0 && ldc % 4 == 0 && aligned16(C) && M >= 512 && K * N * 4 <= 96 * 1024) {
        *handled = true;
        switch ((int)K) {
#define OPN_CASE(k) case k: return launch_tinyk<k>(A, lda, B, ldb, tb ? 1 : 0, C, ldc, (int)M, (int)N, alpha, beta_one, s);
            OPN_SKINNY_CASES(OPN_CASE)
#undef OPN_CASE
        }
    }
    // colred: C = A^T B over a long K with one side at most 16 wide
    if (ta && !tb && K >= 1024) {
        if (N <= 16 && M >= 512 && M % 4 == 0 && lda % 4 == 0 && aligned16(A)) {   // skinny = B [K x N], large = A [K x M]; w = n, j = m
            *handled = true;
            switch ((int)N) {
#define OPN_CASE(w) case w: return launch_colred<w>(B, ldb, A, lda, C, 1, ldc, (int)M, (int)K, alpha, beta_one, s);
                OPN_SKINNY_CASES(OPN_CASE)
                OPN_CASE(9) OPN_CASE(10) OPN_CASE(11) OPN_CASE(12) OPN_CASE(13) OPN_CASE(14) OPN_CASE(15) OPN_CASE(16)
#undef OPN_CASE
            }
        }
        if (M <= 16 && N >= 512 && N % 4 == 0 && ldb % 4 == 0 && aligned16(B)) {   // skinny = A [K x M], large = B [K x N]; w = m, j = n
            *handled = true;
            switch ((int)M) {
#define OPN_CASE(w) case w: return launch_colred<w>(A, lda, B, ldb, C, ldc, 1, (int)N, (int)K, alpha, beta_one, s);
                OPN_SKINNY_CASES(OPN_CASE)
                OPN_CASE(9) OPN_CASE(10) OPN_CASE(11) OPN_CASE(12) OPN_CASE(13) OPN_CASE(14) OPN_CASE(15) OPN_CASE(16)
#undef OPN_CASE
            }
        }
    }
#undef OPN_SKINNY_CASES
    *handled = false;
    return OPN_OK;
}

}  // namespace opn
