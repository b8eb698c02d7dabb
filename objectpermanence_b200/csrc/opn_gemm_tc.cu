// fp32-accurate dense contraction on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   C[M,N] = alpha * A[M,K] * B[N,K]^T (+ beta*C + bias, ReLU),  fp32 in / fp32 out.
//
// Every fp32 operand x is split into two bf16 numbers, x = hi + lo (hi = bf16(x), lo = bf16(x - hi),
// 16 mantissa bits together), and the product is accumulated as  Ahi*Bhi + Ahi*Blo + Alo*Bhi  in the fp32
// TMEM accumulator: three kind::f16 MMAs per K=16 slice, relative operand error 2^-17.  This keeps the
// 1e-4 fp32 parity bound of the hot path through K = 9600 reductions where a single bf16/tf32 pass does not.
//
// Structure (one CTA = one 128 x 128 output tile, 128 threads):
//   * a pre-pass kernel converts (and, if needed, transposes) each operand once into zero-padded K-major
//     bf16 hi / lo matrices;
//   * the GEMM kernel streams 128 x 64 tiles of the four matrices with cp.async (16-byte chunks written in
//     the SWIZZLE_128B K-major canonical layout the UMMA shared-memory descriptor expects) through a
//     3-stage ring; one thread issues tcgen05.mma (cta_group::1, M=128, N=128, K=16) and tcgen05.commit
//     signals an mbarrier when a stage may be overwritten;
//   * the accumulator lives in 128 TMEM columns; the four warps read it back with tcgen05.ld (32 lanes x
//     32 columns per instruction) and apply the epilogue; split-K partial tiles are added with fp32 atomics.
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (CUTLASS, vendored headers, read-only reference).
#include <cuda_bf16.h>

#include "opn_common.cuh"

namespace opn {
namespace {

constexpr int TBM = 128, TBN = 128, TBK = 64, TSTAGES = 3;
constexpr int TC_THREADS = 128;
constexpr int TILE_BYTES = TBM * TBK * 2;           // 16 KB, one operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // Ahi, Alo, Bhi, Blo
constexpr int TC_SMEM = TSTAGES * STAGE_BYTES + 1024;  // + alignment slack

// ---- pre-pass: fp32 -> (hi, lo) bf16, K-major, zero padded ---------------------------------------------------
// dst[r][k] for r < rows_pad, k < k_pad;  source element (r, k) is src[r*ld + k] or, transposed, src[k*ld + r].
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ src, long long ld, int rows, int kdim,
                                                         int transposed, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, int rows_pad, int k_pad) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    if (!transposed) {
        for (int j = ty; j < 32; j += 8) {
            const int r = r0 + j, k = k0 + tx;
            tile[j][tx] = (r < rows && k < kdim) ? src[(long long)r * ld + k] : 0.0f;
        }
    } else {
        for (int j = ty; j < 32; j += 8) {
            const int k = k0 + j, r = r0 + tx;
            tile[tx][j] = (r < rows && k < kdim) ? src[(long long)k * ld + r] : 0.0f;
        }
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, k = k0 + tx;
        if (r < rows_pad && k < k_pad) {
            const float x = tile[j][tx];
            const __nv_bfloat16 h = __float2bfloat16_rn(x);
            hi[(long long)r * k_pad + k] = h;
            lo[(long long)r * k_pad + k] = __float2bfloat16_rn(x - __bfloat162float(h));
        }
    }
}

// ---- tcgen05 helpers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    // K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO unused (=1), version 1
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (>>4), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset (>>4), bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: D=F32, A=B=BF16, both K-major, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kInstrDesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    // bounded spin: a broken pipeline must not hang the GPU
    for (unsigned int i = 0; i < (1u << 26); ++i)
        if (mbar_try_wait(bar, parity)) return;
}

struct TcParams {
    const __nv_bfloat16* a_hi;
    const __nv_bfloat16* a_lo;
    const __nv_bfloat16* b_hi;
    const __nv_bfloat16* b_lo;
    float* C;
    const float* bias;
    long long ldc;
    int M, N, k_pad;       // true M, N; padded K (multiple of 64)
    int kb_per_split;      // k-blocks (of 64) per split
    float alpha;
    int beta_one, relu, atomic_out;
};

__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const TcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t mma_done[TSTAGES];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 1024-byte aligned tile ring (SWIZZLE_128B atoms are 8 rows x 128 B)
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;

    const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
    const int nkb_total = p.k_pad / TBK;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int kb_end = min(nkb_total, kb_begin + p.kb_per_split);
    const int nkb = kb_end - kb_begin;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TSTAGES; ++s) mbar_init(&mma_done[s], 1);
        mbar_fence_init();
    }
    if (warp == 0) {  // TMEM: 128 fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_acc = tmem_base_s;

    // one stage = 4 tiles of [128 rows][64 bf16]; 16-byte chunk c of row r sits at r*128 + ((c ^ (r & 7)) << 4)
    auto load_stage = [&](int stage, int kb) {
        const uint32_t sbase = ring + stage * STAGE_BYTES;
        const long long kofs = (long long)kb * TBK;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int idx = tid + TC_THREADS * i;     // 0 .. 4095
            const int t = idx >> 10;                  // tile: 0 Ahi, 1 Alo, 2 Bhi, 3 Blo
            const int r = (idx >> 3) & 127, c = idx & 7;
            const __nv_bfloat16* base = (t == 0) ? p.a_hi : (t == 1) ? p.a_lo : (t == 2) ? p.b_hi : p.b_lo;
            const int row = ((t < 2) ? m0 : n0) + r;
            const __nv_bfloat16* src = base + (long long)row * p.k_pad + kofs + 8 * c;
            const uint32_t dst = sbase + t * TILE_BYTES + r * 128 + ((c ^ (r & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
    };

    if (nkb > 0) {
#pragma unroll
        for (int s = 0; s < TSTAGES - 1; ++s) {
            if (s < nkb) load_stage(s, kb_begin + s);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % TSTAGES;
            asm volatile("cp.async.wait_group %0;" ::"n"(TSTAGES - 2) : "memory");
            fence_proxy_async_smem();   // cp.async / generic writes -> tensor-core (async proxy) reads
            __syncthreads();
            // queue this k-block's MMAs first (the tensor pipe runs them in order behind k-block kb-1) ...
            if (tid == 0) {
                tcgen05_fence_after();
                const uint32_t sbase = ring + stage * STAGE_BYTES;
#pragma unroll
                for (int k16 = 0; k16 < TBK / 16; ++k16) {
                    const uint64_t ah = make_smem_desc_sw128(sbase + 0 * TILE_BYTES + k16 * 32);
                    const uint64_t al = make_smem_desc_sw128(sbase + 1 * TILE_BYTES + k16 * 32);
                    const uint64_t bh = make_smem_desc_sw128(sbase + 2 * TILE_BYTES + k16 * 32);
                    const uint64_t bl = make_smem_desc_sw128(sbase + 3 * TILE_BYTES + k16 * 32);
                    umma_bf16(tmem_acc, ah, bh, (kb > 0 || k16 > 0) ? 1u : 0u);
                    umma_bf16(tmem_acc, ah, bl, 1u);
                    umma_bf16(tmem_acc, al, bh, 1u);
                }
                umma_commit(&mma_done[stage]);   // implies tcgen05.fence::before_thread_sync
            }
            // ... then refill the stage k-block kb-1 used, once its MMAs have drained
            const int next = kb + TSTAGES - 1;
            if (next < nkb) {
                if (kb >= 1) mbar_wait_spin(&mma_done[(kb - 1) % TSTAGES], (uint32_t)(((kb - 1) / TSTAGES) & 1));
                load_stage(next % TSTAGES, kb_begin + next);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // all MMAs are complete when the last commit has arrived
        mbar_wait_spin(&mma_done[(nkb - 1) % TSTAGES], (uint32_t)(((nkb - 1) / TSTAGES) & 1));
        tcgen05_fence_after();
    }

    // ---- epilogue: TMEM -> registers -> padded shared tile -> coalesced rows of C ---------------------------
    // (all MMAs have completed, so the operand ring is free: the 128 x 128 fp32 tile is staged there with a
    //  row pitch of 132 floats, which makes both the per-row STS.128 of the TMEM read-out and the per-row
    //  LDS.128 of the write-out conflict free)
    constexpr int EP = TBN + 4;
    float* tile_s = reinterpret_cast<float*>(smem_dyn + (ring - smem_u32(smem_dyn)));
    {
        const int r = warp * 32 + lane;  // warp w may only touch TMEM lanes [32w, 32w+32)
#pragma unroll 1
        for (int col = 0; col < TBN; col += 32) {
            uint32_t v[32];
            if (nkb > 0) {
                const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)col;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<uint4*>(&tile_s[r * EP + col + j]) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
    __syncthreads();
    {
        const bool lead = (blockIdx.z == 0);
        const int n = n0 + 4 * lane;           // this lane's 4 columns
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (n + 3 < p.N);
        float b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias && (!p.atomic_out || lead)) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < p.N) b4[j] = p.bias[n + j];
        }
        for (int r = warp; r < TBM; r += 4) {   // one warp writes one 512-byte row per iteration
            const int m = m0 + r;
            if (m >= p.M) break;
            const float4 t = *reinterpret_cast<const float4*>(&tile_s[r * EP + 4 * lane]);
            float x[4] = {p.alpha * t.x + b4[0], p.alpha * t.y + b4[1], p.alpha * t.z + b4[2], p.alpha * t.w + b4[3]};
            float* c = p.C + (long long)m * p.ldc + n;
            if (p.atomic_out) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < p.N) atomicAdd(c + j, x[j]);
            } else if (vec_ok) {
                if (p.beta_one) {
                    const float4 o = *reinterpret_cast<const float4*>(c);
                    x[0] += o.x; x[1] += o.y; x[2] += o.z; x[3] += o.w;
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.0f);
                }
                *reinterpret_cast<float4*>(c) = make_float4(x[0], x[1], x[2], x[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < p.N) {
                        float y = x[j];
                        if (p.beta_one) y += c[j];
                        if (p.relu) y = fmaxf(y, 0.0f);
                        c[j] = y;
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(128) : "memory");
}

__global__ void zero_strided_tc_kernel(float* C, long long ldc, int M, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)M * N) C[(i / N) * ldc + (i % N)] = 0.0f;
}

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

}  // namespace

// Bytes of scratch the tensor-core path needs for an (M, N, K) contraction (split operands).
long long gemm_tc_workspace_bytes(long long M, long long N, long long K) {
    const long long mp = round_up(M, TBM), np = round_up(N, TBN), kp = round_up(K, TBK);
    return 2 * (mp + np) * kp * (long long)sizeof(__nv_bfloat16) + 1024;
}

// op(A) is [M,K]: trans_a == 0 -> A[m*lda + k], else A[k*lda + m];  op(B) is [K,N]: trans_b == 0 -> B[k*ldb + n],
// else B[n*ldb + k].  Same contract as opn_sgemm.
int gemm_tc(int trans_a, int trans_b, long long M, long long N, long long K, float alpha, const float* A, long long lda,
            const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias, int relu,
            void* workspace, long long workspace_bytes, cudaStream_t s) {
    const long long mp = round_up(M, TBM), np = round_up(N, TBN), kp = round_up(K, TBK);
    if (workspace_bytes < gemm_tc_workspace_bytes(M, N, K)) {
        set_error("gemm_tc: workspace too small");
        return OPN_ERR_BAD_ARG;
    }
    char* ws = reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* a_lo = a_hi + mp * kp;
    __nv_bfloat16* b_hi = a_lo + mp * kp;
    __nv_bfloat16* b_lo = b_hi + np * kp;

    // pre-pass: A as [M][K] (source is [K][M] when trans_a), B as [N][K] (source is [K][N] unless trans_b)
    split_bf16_kernel<<<dim3((unsigned)(kp / 32), (unsigned)(mp / 32)), 256, 0, s>>>(A, lda, (int)M, (int)K, trans_a ? 1 : 0,
                                                                                   a_hi, a_lo, (int)mp, (int)kp);
    OPN_CUDA(cudaGetLastError());
    split_bf16_kernel<<<dim3((unsigned)(kp / 32), (unsigned)(np / 32)), 256, 0, s>>>(B, ldb, (int)N, (int)K, trans_b ? 0 : 1,
                                                                                   b_hi, b_lo, (int)np, (int)kp);
    OPN_CUDA(cudaGetLastError());
    count_launch(2);

    int dev = 0, sms = 148;
    OPN_CUDA(cudaGetDevice(&dev));
    OPN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int gm = (int)(mp / TBM), gn = (int)(np / TBN), nkb = (int)(kp / TBK);
    int splits = 1;
    if (!relu && gm * gn < sms && nkb >= 8) {
        splits = sms / (gm * gn);
        if (splits > nkb / 4) splits = nkb / 4;
        if (splits > 32) splits = 32;
        if (splits < 1) splits = 1;
    }
    const int kb_per_split = (nkb + splits - 1) / splits;
    splits = (nkb + kb_per_split - 1) / kb_per_split;

    TcParams p;
    p.a_hi = a_hi; p.a_lo = a_lo; p.b_hi = b_hi; p.b_lo = b_lo;
    p.C = C; p.bias = bias; p.ldc = ldc;
    p.M = (int)M; p.N = (int)N; p.k_pad = (int)kp;
    p.kb_per_split = kb_per_split;
    p.alpha = alpha;
    p.beta_one = beta == 1.0f;
    p.relu = relu;
    p.atomic_out = splits > 1;
    if (splits > 1 && !p.beta_one) {
        const long long n = M * N;
        zero_strided_tc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(C, ldc, (int)M, (int)N);
        OPN_CUDA(cudaGetLastError());
        count_launch();
    }
    OPN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    gemm_tc_kernel<<<dim3((unsigned)gn, (unsigned)gm, (unsigned)splits), TC_THREADS, TC_SMEM, s>>>(p);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

}  // namespace opn
