// Input projections with a short contraction: C[M,N] = A[M,K] · B[N,K]^T, 16 <= K <= 128, M in the thousands.
//
// This is the x-projection of an LSTM layer (reference: the `weight_ih_l0` half of nn.LSTM,
// baselines/learned_models.py:29 (90 -> 4·256), :100 (75 -> 4·H)): 9,600 x 90 times 90 x 1,024 at the headline shape.  The
// general tensor-core path of opn_sgemm (opn_gemm_tc.cu) pays two operand pre-passes and a TMA / tcgen05 pipeline that is all
// prologue at K = 90 (44 us for a product that writes 39 MB); here one kernel reads the fp32 operands, splits them into bf16
// hi + lo planes in shared memory on the way in and runs the three products (hi·hi + lo·hi + hi·lo; one in
// OPN_PRECISION_16BIT) as warp-level bf16 MMAs with fp32 accumulators, so the kernel is bound by its 4·M·N bytes of output.
//
//   Persistent: a CTA keeps one 128-column tile of B (split once) and walks row tiles of A; 128 x 128 output tile per
//   iteration, 8 warps as 2 x 4 (64 x 32 per warp), whole K resident; two CTAs per SM overlap one tile's loads with the other's
//   products and stores.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "opn_common.cuh"

namespace opn {

int current_precision();   // opn_api.cu

namespace {

constexpr int PBM = 128, PBN = 128, PTHREADS = 256;
constexpr int STAGE_BATCH = 16;     // loads in flight per thread while a tile is staged (20 and more spill at 128 registers)

struct ProjParams {
    const float* A;
    const float* B;
    float* C;
    long long lda, ldb, ldc;
    int M, N, K, KP;        // KP: K rounded up to the MMA depth of 16
    int tiles_n, tiles_m, slots;   // `slots` CTAs share a column tile and take the row tiles slot, slot + slots, ...
    int single;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split2(float x, float y, __nv_bfloat162& hi, __nv_bfloat162& lo) {
    hi = __floats2bfloat162_rn(x, y);
    lo = __floats2bfloat162_rn(x - __low2float(hi), y - __high2float(hi));
}

// rows [row0, row0 + 128) x columns [0, K) of a row-major fp32 matrix -> bf16 hi / lo planes [128][lds]; rows beyond `rows`
// become zero (the padding columns up to KP are zeroed once per kernel).  Slot f = s * 256 + tid of a thread is the pair
// (f / PK, f % PK) of the tile, PK = ceil(K / 2): consecutive threads read consecutive words, and a batch of STAGE_BATCH loads
// is in flight per thread before the first conversion (dependent round trips made the first version latency bound).
template <bool VEC>
__device__ __forceinline__ void stage_planes(const float* __restrict__ src, long long ld, int row0, int rows, int K,
                                             __nv_bfloat16* hi, __nv_bfloat16* lo, int lds) {
    const int PK = (K + 1) >> 1;
    const float inv_pk = 1.0f / (float)PK;
    const int total = PBM * PK;
    for (int base = threadIdx.x; base < total; base += PTHREADS * STAGE_BATCH) {
        float x[STAGE_BATCH], y[STAGE_BATCH];
#pragma unroll
        for (int u = 0; u < STAGE_BATCH; ++u) {
            const int f = base + u * PTHREADS;
            const int r = (int)(((float)f + 0.5f) * inv_pk);      // exact: f < 2^14, PK <= 64
            const int c = 2 * (f - r * PK);
            const bool live = f < total && row0 + r < rows;
            const float* at = src + (long long)(row0 + r) * ld + c;
            x[u] = 0.0f;
            y[u] = 0.0f;
            if (VEC) {      // K even on this path: a pair never straddles the end of the row
                if (live) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(at));
                    x[u] = v.x;
                    y[u] = v.y;
                }
            } else {
                if (live) x[u] = __ldg(at);
                if (live && c + 1 < K) y[u] = __ldg(at + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < STAGE_BATCH; ++u) {
            const int f = base + u * PTHREADS;
            if (f < total) {
                const int r = (int)(((float)f + 0.5f) * inv_pk);
                const int c = 2 * (f - r * PK);
                __nv_bfloat162 h, l;
                split2(x[u], y[u], h, l);
                *reinterpret_cast<__nv_bfloat162*>(hi + r * lds + c) = h;
                *reinterpret_cast<__nv_bfloat162*>(lo + r * lds + c) = l;
            }
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(PTHREADS, 2) proj_kernel(const ProjParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lds = p.KP + 8;       // row stride in elements: an odd number of 16-byte chunks (ldmatrix without bank conflicts)
    __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* a_lo = a_hi + PBM * lds;
    __nv_bfloat16* b_hi = a_lo + PBM * lds;
    __nv_bfloat16* b_lo = b_hi + PBN * lds;

    const int tile_n = blockIdx.x % p.tiles_n, slot = blockIdx.x / p.tiles_n;
    const int col0 = tile_n * PBN;
    {   // padding columns 2·ceil(K / 2) .. KP of the four planes (back to back in memory: 512 rows of lds elements)
        const int c_first = ((p.K + 1) >> 1) << 1, width = p.KP - c_first;
        for (int f = threadIdx.x; f < 4 * PBM * width; f += PTHREADS) {
            const int r = f / width;
            a_hi[r * lds + c_first + (f - r * width)] = __float2bfloat16_rn(0.0f);
        }
    }
    stage_planes<VEC>(p.B, p.ldb, col0, p.N, p.K, b_hi, b_lo, lds);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    // ldmatrix lane addresses: A 16 x 16 block = rows (lane & 15), k half (lane >> 4); B two 8-column blocks per x4 =
    // column (lane & 7) + 8 (lane >> 4), k half ((lane >> 3) & 1)
    const int a_off = (wm + (lane & 15)) * lds + (lane >> 4) * 8;
    const int b_off = (wn + (lane & 7) + ((lane >> 4) << 3)) * lds + ((lane >> 3) & 1) * 8;
    const int ksteps = p.KP >> 4;

    for (int tile_m = slot; tile_m < p.tiles_m; tile_m += p.slots) {
        const int row0 = tile_m * PBM;
        stage_planes<VEC>(p.A, p.lda, row0, p.M, p.K, a_hi, a_lo, lds);
        __syncthreads();

        float acc[4][4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0f;

        for (int ks = 0; ks < ksteps; ++ks) {
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                uint32_t r[4];
                ldsm_x4(r, b_hi + b_off + j * 8 * lds + ks * 16);
                bh[j][0] = r[0]; bh[j][1] = r[1]; bh[j + 1][0] = r[2]; bh[j + 1][1] = r[3];
                if (!p.single) {
                    ldsm_x4(r, b_lo + b_off + j * 8 * lds + ks * 16);
                    bl[j][0] = r[0]; bl[j][1] = r[1]; bl[j + 1][0] = r[2]; bl[j + 1][1] = r[3];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t a[4];
                if (!p.single) {
                    // the two small products first, the large one on top of them
                    ldsm_x4(a, a_lo + a_off + i * 16 * lds + ks * 16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_bf16(acc[i][j], a, bh[j][0], bh[j][1]);
                }
                ldsm_x4(a, a_hi + a_off + i * 16 * lds + ks * 16);
                if (!p.single) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_bf16(acc[i][j], a, bl[j][0], bl[j][1]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_bf16(acc[i][j], a, bh[j][0], bh[j][1]);
            }
        }

        // accumulator (i, j): rows wm + 16 i + lane / 4 (+ 8), columns wn + 8 j + 2 (lane % 4) (+ 1): a quad writes one
        // 32-byte sector
        const int r_base = row0 + wm + (lane >> 2), c_base = col0 + wn + 2 * (lane & 3);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = r_base + i * 16 + half * 8;
                if (r >= p.M) continue;
                float* out = p.C + (long long)r * p.ldc;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = c_base + j * 8;
                    if (c < p.N)      // N even (checked by the host): c + 1 < N as well
                        *reinterpret_cast<float2*>(out + c) = make_float2(acc[i][j][2 * half], acc[i][j][2 * half + 1]);
                }
            }
        }
        __syncthreads();      // the A planes are rewritten by the next row tile
    }
}

}  // namespace

// Tries the short-K projection kernel.  *handled = true when it was launched (return code = its status).
int gemm_proj(bool ta, bool tb, long long M, long long N, long long K, float alpha, const float* A, long long lda,
              const float* B, long long ldb, float beta, float* C, long long ldc, cudaStream_t s, bool* handled) {
    *handled = false;
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("OPN_GEMM_PROJ");
        enabled = (e && e[0] == '0') ? 0 : 1;
    }
    if (!enabled || ta || !tb || alpha != 1.0f || beta != 0.0f) return OPN_OK;
    if (K < 16 || K > 128 || M < 1024 || N < 128 || (N & 1) || (ldc & 1) || M >= (1LL << 31) - PBM || N >= (1LL << 31) - PBN)
        return OPN_OK;
    if ((reinterpret_cast<uintptr_t>(C) & 7) != 0) return OPN_OK;
    int dev = 0, sms = 148;
    OPN_CUDA(cudaGetDevice(&dev));
    OPN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ProjParams p;
    p.A = A; p.B = B; p.C = C;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.KP = (int)((K + 15) / 16 * 16);
    p.tiles_n = (int)((N + PBN - 1) / PBN);
    p.tiles_m = (int)((M + PBM - 1) / PBM);
    if (p.tiles_n > 65535) return OPN_OK;
    // CTAs per column tile: two resident CTAs per SM in all (measured: 37 CTAs per column tile with 2 or 3 row tiles each beat
    // 25 with 3 each at [9600, 90] x [90, 1024]: 34.8 against 41.0 us)
    const int ctas_per_sm = p.KP <= 96 ? 2 : 1;
    int slots = ctas_per_sm * sms / p.tiles_n;
    if (slots < 1) slots = 1;
    if (slots > p.tiles_m) slots = p.tiles_m;
    if (const char* e = getenv("OPN_PROJ_SLOTS")) {     // tuning knob
        const int v = atoi(e);
        if (v >= 1 && v <= p.tiles_m) slots = v;
    }
    p.slots = slots;
    p.single = current_precision() == OPN_PRECISION_16BIT ? 1 : 0;
    const bool vec = (K % 2 == 0) && (lda % 2 == 0) && (ldb % 2 == 0) && (reinterpret_cast<uintptr_t>(A) & 7) == 0 &&
                     (reinterpret_cast<uintptr_t>(B) & 7) == 0;
    const size_t smem = (size_t)(PBM + PBN) * 2 * (p.KP + 8) * sizeof(__nv_bfloat16);
    auto kernel = vec ? proj_kernel<true> : proj_kernel<false>;
    OPN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *handled = true;
    kernel<<<(unsigned)(p.tiles_n * slots), PTHREADS, smem, s>>>(p);
    count_launch();
    OPN_CUDA(cudaGetLastError());
    return OPN_OK;
}

}  // namespace opn
