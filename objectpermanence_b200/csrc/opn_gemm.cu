// fp32 dense contraction on the CUDA-core FFMA pipe (fp32 accumulate), sm_100a.
//
// C[M,N] = alpha * op(A) * op(B) + beta * C + bias, optional ReLU.  Used for every
// time-parallel product around the recurrences: LSTM input projections, weight gradients
// (K = B*T), input gradients, the bbox / who-to-track heads and the transformer variant's
// projections.  See include/opnet_b200.h for the reference call sites it replaces.
//
// 128 x BN x 16 tiles, 256 threads, 8 x (BN/16) register micro-tiles, register-staged double
// buffering, optional split-K (fp32 atomics) when the M x N grid alone cannot fill 148 SMs.
#include <stdlib.h>

#include "opn_common.cuh"

namespace opn {

// tensor-core path (opn_gemm_tc.cu)
long long gemm_tc_workspace_bytes(long long M, long long N, long long K);
int gemm_tc(int trans_a, int trans_b, long long M, long long N, long long K, float alpha, const float* A, long long lda,
            const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias, int relu,
            void* workspace, long long workspace_bytes, cudaStream_t s);

// short-K input projections (opn_gemm_proj.cu)
int gemm_proj(bool ta, bool tb, long long M, long long N, long long K, float alpha, const float* A, long long lda,
              const float* B, long long ldb, float beta, float* C, long long ldc, cudaStream_t s, bool* handled);
// skinny shapes (opn_gemm_skinny.cu)
int gemm_skinny(bool ta, bool tb, long long M, long long N, long long K, float alpha, const float* A, long long lda,
                const float* B, long long ldb, float beta, float* C, long long ldc, cudaStream_t s, bool* handled);

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int GEMM_THREADS = 256;

struct GemmParams {
    const float* A;
    const float* B;
    float* C;
    const float* bias;
    long long lda, ldb, ldc;
    int M, N, K;
    float alpha;
    int beta_one;
    int relu;
    int k_per_split;  // multiple of BK
    int atomic_out;
};

template <int BN, bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS) sgemm_kernel(const GemmParams p) {
    constexpr int TN = BN / 16;          // micro-tile width
    constexpr int A_PER = BM * BK / GEMM_THREADS;  // 8
    constexpr int B_PER = BN * BK / GEMM_THREADS;  // 8 or 4 or 1
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int m_blk = blockIdx.y * BM;
    const int n_blk = blockIdx.x * BN;
    const int k_begin = blockIdx.z * p.k_per_split;
    const int k_end = min(p.K, k_begin + p.k_per_split);

    const int ty = tid / 16, tx = tid % 16;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    float a_reg[A_PER], b_reg[B_PER];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int e = 0; e < A_PER; ++e) {
            const int idx = tid + GEMM_THREADS * e;
            int m, k;
            if (TA) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
            const int gm = m_blk + m, gk = k0 + k;
            float v = 0.0f;
            if (gm < p.M && gk < k_end) {
                if (TA) v = __ldg(p.A + (long long)gk * p.lda + gm);
                else v = __ldg(p.A + (long long)gm * p.lda + gk);
            }
            a_reg[e] = v;
        }
#pragma unroll
        for (int e = 0; e < B_PER; ++e) {
            const int idx = tid + GEMM_THREADS * e;
            int n, k;
            if (TB) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            const int gn = n_blk + n, gk = k0 + k;
            float v = 0.0f;
            if (gn < p.N && gk < k_end) {
                if (TB) v = __ldg(p.B + (long long)gn * p.ldb + gk);
                else v = __ldg(p.B + (long long)gk * p.ldb + gn);
            }
            b_reg[e] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int e = 0; e < A_PER; ++e) {
            const int idx = tid + GEMM_THREADS * e;
            int m, k;
            if (TA) { m = idx % BM; k = idx / BM; } else { k = idx % BK; m = idx / BK; }
            As[buf][k][m] = a_reg[e];
        }
#pragma unroll
        for (int e = 0; e < B_PER; ++e) {
            const int idx = tid + GEMM_THREADS * e;
            int n, k;
            if (TB) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            Bs[buf][k][n] = b_reg[e];
        }
    };

    if (k_begin < k_end) {
        load_tiles(k_begin);
        store_tiles(0);
        __syncthreads();
        int buf = 0;
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
            const bool has_next = (k0 + BK) < k_end;
            if (has_next) load_tiles(k0 + BK);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[8], b[TN];
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                if constexpr (TN == 8) {
                    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                    const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
                    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                    b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
                } else if constexpr (TN == 4) {
                    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                } else {
                    b[0] = Bs[buf][k][tx];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (has_next) {
                store_tiles(buf ^ 1);
                __syncthreads();
                buf ^= 1;
            }
        }
    }

    // epilogue.  Column of micro-tile element j: TN==8 -> (j<4 ? tx*4+j : 64+tx*4+j-4); TN==4 -> tx*4+j; TN==1 -> tx
    const bool lead_split = (blockIdx.z == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m_blk + ty * 8 + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int col;
            if constexpr (TN == 8) col = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
            else if constexpr (TN == 4) col = tx * 4 + j;
            else col = tx;
            const int gn = n_blk + col;
            if (gn >= p.N) continue;
            float v = p.alpha * acc[i][j];
            float* c = p.C + (long long)gm * p.ldc + gn;
            if (p.atomic_out) {
                if (lead_split && p.bias) v += p.bias[gn];
                atomicAdd(c, v);
            } else {
                if (p.bias) v += p.bias[gn];
                if (p.beta_one) v += *c;
                if (p.relu) v = fmaxf(v, 0.0f);
                *c = v;
            }
        }
    }
}

__global__ void zero_strided_kernel(float* C, long long ldc, int M, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)M * N) C[(i / N) * ldc + (i % N)] = 0.0f;
}

template <int BN, bool TA, bool TB>
int launch(const GemmParams& p, dim3 grid, cudaStream_t s) {
    sgemm_kernel<BN, TA, TB><<<grid, GEMM_THREADS, 0, s>>>(p);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

template <int BN>
int dispatch(const GemmParams& p, bool ta, bool tb, dim3 grid, cudaStream_t s) {
    if (!ta && !tb) return launch<BN, false, false>(p, grid, s);
    if (!ta && tb) return launch<BN, false, true>(p, grid, s);
    if (ta && !tb) return launch<BN, true, false>(p, grid, s);
    return launch<BN, true, true>(p, grid, s);
}

// Shapes worth the operand pre-pass of the tcgen05 path: both output dims at least one half tile, a K of at least 64
// to amortise it, and at least 2^27 multiply-adds.  OPN_GEMM_TC=0 forces the FFMA kernel everywhere.
bool tc_eligible(int64_t M, int64_t N, int64_t K) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("OPN_GEMM_TC");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    if (!enabled) return false;
    static long long min_k = -1;
    if (min_k < 0) {
        const char* e = getenv("OPN_GEMM_TC_MINK");   // tuning knob
        min_k = e ? atoll(e) : 64;
    }
    return M >= 64 && N >= 64 && K >= min_k && (double)M * (double)N * (double)K >= 134217728.0;
}

}  // namespace
}  // namespace opn

using namespace opn;

extern "C" int64_t opn_sgemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0 || !tc_eligible(M, N, K)) return 0;
    return (int64_t)gemm_tc_workspace_bytes(M, N, K);
}

extern "C" int opn_sgemm(int trans_a, int trans_b, int64_t M, int64_t N, int64_t K, float alpha, const float* A,
                         int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                         const float* bias, int relu, void* workspace, int64_t workspace_bytes, void* stream) {
    OPN_CHECK_ARG(M > 0 && N > 0 && K >= 0, "sgemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N,
                  (long long)K);
    OPN_CHECK_ARG(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "sgemm: dimension exceeds int32");
    OPN_CHECK_ARG(A && B && C, "sgemm: null pointer");
    OPN_CHECK_ARG(beta == 0.0f || beta == 1.0f, "sgemm: beta must be 0 or 1");
    const bool ta = trans_a != 0, tb = trans_b != 0;
    cudaStream_t s = as_stream(stream);
    if (K > 0 && bias == nullptr && !relu) {
        bool handled = false;
        const int rc = gemm_skinny(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, s, &handled);
        if (handled || rc != OPN_OK) return rc;
        const int rc2 = gemm_proj(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, s, &handled);
        if (handled || rc2 != OPN_OK) return rc2;
    }
    if (K > 0 && workspace != nullptr && tc_eligible(M, N, K) &&
        workspace_bytes >= (int64_t)gemm_tc_workspace_bytes(M, N, K))
        return gemm_tc(trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, relu, workspace,
                       workspace_bytes, s);

    int dev = 0, sms = 148;
    OPN_CUDA(cudaGetDevice(&dev));
    OPN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    const int bn = (N > 64) ? 128 : (N > 16 ? 64 : 16);
    const int gm = (int)((M + BM - 1) / BM), gn = (int)((N + bn - 1) / bn);
    // split K when the output grid is small and K is long (weight-gradient shapes)
    int splits = 1;
    const int k_tiles = (int)((K + BK - 1) / BK);
    if (!relu && gm * gn < sms && k_tiles >= 16) {
        splits = (2 * sms) / (gm * gn);
        const int max_splits = k_tiles / 8;
        if (splits > max_splits) splits = max_splits;
        if (splits > 64) splits = 64;
        if (splits < 1) splits = 1;
    }
    int k_per_split = ((k_tiles + splits - 1) / splits) * BK;
    splits = (int)((K + k_per_split - 1) / k_per_split);
    if (splits < 1) splits = 1;

    GemmParams p;
    p.A = A; p.B = B; p.C = C; p.bias = bias;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.alpha = alpha;
    p.beta_one = beta == 1.0f;
    p.relu = relu;
    p.k_per_split = k_per_split;
    p.atomic_out = splits > 1;

    if (splits > 1 && !p.beta_one) {
        const long long n = (long long)M * N;
        zero_strided_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(C, ldc, (int)M, (int)N);
        OPN_CUDA(cudaGetLastError());
        count_launch();
    }
    dim3 grid((unsigned)gn, (unsigned)gm, (unsigned)splits);
    if (bn == 128) return dispatch<128>(p, ta, tb, grid, s);
    if (bn == 64) return dispatch<64>(p, ta, tb, grid, s);
    return dispatch<16>(p, ta, tb, grid, s);
}
