// fp32-accurate dense contraction on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   C[M,N] = alpha * A[M,K] * B[N,K]^T (+ beta*C + bias, ReLU),  fp32 in / fp32 out.
//
// Every fp32 operand x is split into two bf16 numbers, x = hi + lo (hi = bf16(x), lo = bf16(x - hi),
// 16 mantissa bits together), and the product is accumulated as  Ahi*Bhi + Ahi*Blo + Alo*Bhi  in the fp32
// TMEM accumulator: three kind::f16 MMAs per K=16 slice, relative operand error 2^-17.  This keeps the
// 1e-4 fp32 parity bound of the hot path through K = 9600 reductions where a single bf16/tf32 pass does not.
//
// Structure (persistent, warp specialised; 192 threads, one CTA per SM):
//   * a pre-pass kernel converts (and, if needed, transposes) each operand once into zero-padded K-major
//     bf16 hi / lo matrices;
//   * warp 0 (one lane) is the TMA producer: per 64-wide k-block four cp.async.bulk.tensor.2d loads bring the
//     128 x 64 tiles of Ahi, Alo, Bhi, Blo into a 3-stage shared-memory ring in the SWIZZLE_128B K-major layout
//     the UMMA descriptors expect (full / empty mbarriers per stage);
//   * warp 1 (one lane) issues tcgen05.mma (cta_group::1, M=128, N=128, K=16, kind::f16), twelve per k-block;
//     tcgen05.commit releases the stage and, after the last k-block, hands the accumulator to the epilogue;
//   * the accumulator is double buffered in TMEM (2 x 128 columns), so warps 2-5 read tile i back with
//     tcgen05.ld, transpose it through a small per-warp shared tile and write 128-byte rows of C while the
//     mainloop of tile i+1 is already running; split-K partial tiles are added with fp32 atomics.
// The first version (cp.async by all threads + a CTA barrier per k-block, one tile per CTA) reached 11-13 %
// tensor-pipe activity (profiles/r01_gemm_tc_v1_ncu_summary.csv).
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (CUTLASS, vendored headers, read-only reference).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "opn_common.cuh"

namespace opn {
int current_precision();   // opn_api.cu
namespace {

constexpr int TBM = 128, TBN = 128, TBK = 64, TSTAGES = 3;
constexpr int TC_THREADS = 192;                     // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int TILE_BYTES = TBM * TBK * 2;           // 16 KB, one operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // Ahi, Alo, Bhi, Blo
constexpr int EPI_PITCH = 36;                       // per-warp 32 x 32 transpose tile: 16-byte aligned rows, conflict free
constexpr int TC_SMEM = TSTAGES * STAGE_BYTES + 4 * 32 * EPI_PITCH * 4 + 1024;  // + alignment slack

// ---- pre-pass: fp32 -> (hi, lo) bf16, K-major, zero padded ---------------------------------------------------
// dst[r][k] for r < rows_pad, k < k_pad;  source element (r, k) is src[r*ld + k] or, transposed, src[k*ld + r].
// 64 x 64 tiles, 256 threads: 16-byte loads along the contiguous source dimension, transposition through shared
// memory, 16-byte stores of 8 bf16 along k (the scalar 32 x 32 version reached 2.7 TB/s: 160 us of the OPNet step).
__device__ __forceinline__ void split_store8(const float (&x)[8], __nv_bfloat16* hi, __nv_bfloat16* lo) {
    __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i] = __float2bfloat16_rn(x[i]);
        l[i] = __float2bfloat16_rn(x[i] - __bfloat162float(h[i]));
    }
    *reinterpret_cast<uint4*>(hi) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(lo) = *reinterpret_cast<const uint4*>(l);
}

template <bool VEC>
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ src, long long ld, int rows, int kdim,
                                                         int transposed, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, int rows_pad, int k_pad) {
    __shared__ float tile[64][65];   // [r][k]
    const int r0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
    const int tid = threadIdx.x;
    // ---- load the 64 x 64 tile: 16 float4 per row of the contiguous dimension, 4 passes of 16 rows
    const int c4 = tid & 15, line = tid >> 4;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int j = line + 16 * pass;   // index along the strided dimension
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (!transposed) {                // rows r = r0 + j, contiguous k = k0 + 4*c4 ..
            const int r = r0 + j, k = k0 + 4 * c4;
            if (r < rows) {
                const float* p = src + (long long)r * ld + k;
                if (VEC && k + 3 < kdim) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                    v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (k + e < kdim) v[e] = __ldg(p + e);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) tile[j][4 * c4 + e] = v[e];
        } else {                          // rows k = k0 + j of the source, contiguous r = r0 + 4*c4 ..
            const int k = k0 + j, r = r0 + 4 * c4;
            if (k < kdim) {
                const float* p = src + (long long)k * ld + r;
                if (VEC && r + 3 < rows) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
                    v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (r + e < rows) v[e] = __ldg(p + e);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) tile[4 * c4 + e][j] = v[e];
        }
    }
    __syncthreads();
    // ---- store: 8 consecutive k per thread (16 bytes of bf16), 8 threads per row, 2 passes of 32 rows
    const int k8 = tid & 7, rr = tid >> 3;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int r = r0 + rr + 32 * pass, k = k0 + 8 * k8;
        if (r < rows_pad && k < k_pad) {   // k_pad and rows_pad are multiples of 64
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = tile[rr + 32 * pass][8 * k8 + e];
            split_store8(x, hi + (long long)r * k_pad + k, lo + (long long)r * k_pad + k);
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
    for (unsigned int i = 0; i < (1u << 22); ++i)
        if (mbar_try_wait(bar, parity)) return;
}

struct TcParams {
    float* C;
    const float* bias;
    long long ldc;
    int M, N;
    int tiles_m, tiles_n, splits;
    int nkb_total;         // k-blocks of 64 in the padded K
    int kb_per_split;
    float alpha;
    int beta_one, relu, atomic_out;
    int single;            // 1e-2 arithmetic mode: the hi.hi product alone (lo tiles neither loaded nor multiplied)
};

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
               const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo, const TcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[TSTAGES], empty_bar[TSTAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ring = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // 1024-byte aligned (SWIZZLE_128B atoms)
    float* epi_s = reinterpret_cast<float*>(smem_dyn + (ring - smem_u32(smem_dyn)) + TSTAGES * STAGE_BYTES);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TSTAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);   // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) {  // TMEM: two 128-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int n_work = p.tiles_m * p.tiles_n * p.splits;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int split = w % p.splits, tile = w / p.splits;
                const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
                const int kb0 = split * p.kb_per_split, kb1 = min(p.nkb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait_spin(&empty_bar[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[stage], p.single ? 2 * TILE_BYTES : STAGE_BYTES);
                    const uint32_t sbase = ring + stage * STAGE_BYTES;
                    tma_load_2d(sbase + 0 * TILE_BYTES, &map_ahi, kb * TBK, tm * TBM, &full_bar[stage]);
                    tma_load_2d(sbase + 2 * TILE_BYTES, &map_bhi, kb * TBK, tn * TBN, &full_bar[stage]);
                    if (!p.single) {
                        tma_load_2d(sbase + 1 * TILE_BYTES, &map_alo, kb * TBK, tm * TBM, &full_bar[stage]);
                        tma_load_2d(sbase + 3 * TILE_BYTES, &map_blo, kb * TBK, tn * TBN, &full_bar[stage]);
                    }
                    if (++stage == TSTAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int split = w % p.splits;
                const int kb0 = split * p.kb_per_split, kb1 = min(p.nkb_total, kb0 + p.kb_per_split);
                mbar_wait_spin(&tmem_empty[acc], acc_phase ^ 1u);   // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * TBN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait_spin(&full_bar[stage], phase);          // TMA bytes have landed
                    tcgen05_fence_after();
                    const uint32_t sbase = ring + stage * STAGE_BYTES;
#pragma unroll
                    for (int k16 = 0; k16 < TBK / 16; ++k16) {
                        const uint64_t ah = make_smem_desc_sw128(sbase + 0 * TILE_BYTES + k16 * 32);
                        const uint64_t al = make_smem_desc_sw128(sbase + 1 * TILE_BYTES + k16 * 32);
                        const uint64_t bh = make_smem_desc_sw128(sbase + 2 * TILE_BYTES + k16 * 32);
                        const uint64_t bl = make_smem_desc_sw128(sbase + 3 * TILE_BYTES + k16 * 32);
                        umma_bf16(tmem_acc, ah, bh, (kb > kb0 || k16 > 0) ? 1u : 0u);
                        if (!p.single) {
                            umma_bf16(tmem_acc, ah, bl, 1u);
                            umma_bf16(tmem_acc, al, bh, 1u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);                  // stage may be refilled when these retire
                    if (++stage == TSTAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&tmem_full[acc]);                        // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ================= epilogue warps (2..5): TMEM lane quadrant = warp % 4 =================
        const int q = warp & 3;
        float* my_s = epi_s + (warp - 2) * 32 * EPI_PITCH;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int split = w % p.splits, tile = w / p.splits;
            const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
            const int m0 = tm * TBM + q * 32, n0 = tn * TBN;
            const bool lead = (split == 0);
            mbar_wait_spin(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
#pragma unroll 1
            for (int col = 0; col < TBN; col += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TBN + col);
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
                if (col + 32 == TBN) {  // last read of this accumulator: hand it back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                // lane = row: transpose through shared memory so that a warp store covers whole 128-byte row segments
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(&my_s[lane * EPI_PITCH + j]) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                const int nq = n0 + col + 4 * (lane & 7);          // this lane's 4 columns
                const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (nq + 3 < p.N);
                float b4[4] = {0.f, 0.f, 0.f, 0.f};
                if (p.bias && (!p.atomic_out || lead)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (nq + j < p.N) b4[j] = p.bias[nq + j];
                }
#pragma unroll
                for (int rr = 0; rr < 32; rr += 4) {               // 4 rows x 128 bytes per warp instruction
                    const int r = rr + (lane >> 3);
                    const int m = m0 + r;
                    const float4 t = *reinterpret_cast<const float4*>(&my_s[r * EPI_PITCH + 4 * (lane & 7)]);
                    if (m < p.M && nq < p.N) {
                        float x[4] = {p.alpha * t.x + b4[0], p.alpha * t.y + b4[1], p.alpha * t.z + b4[2], p.alpha * t.w + b4[3]};
                        float* c = p.C + (long long)m * p.ldc + nq;
                        if (p.atomic_out) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (nq + j < p.N) atomicAdd(c + j, x[j]);
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
                                if (nq + j < p.N) {
                                    float y = x[j];
                                    if (p.beta_one) y += c[j];
                                    if (p.relu) y = fmaxf(y, 0.0f);
                                    c[j] = y;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

__global__ void zero_strided_tc_kernel(float* C, long long ldc, int M, int N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)M * N) C[(i / N) * ldc + (i % N)] = 0.0f;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
PFN_cuTensorMapEncodeTiled get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(ptr);
    }
    return fn;
}

// 2-D bf16 tensor [rows][k_pad] (K contiguous), box = 64 (K) x 128 (rows), 128-byte swizzle
int make_operand_map(CUtensorMap* map, const __nv_bfloat16* base, long long rows, long long k_pad) {
    PFN_cuTensorMapEncodeTiled encode = get_encode_fn();
    if (!encode) {
        set_error("gemm_tc: cuTensorMapEncodeTiled entry point not available");
        return OPN_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k_pad * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)TBM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return OPN_ERR_CUDA;
    }
    return OPN_OK;
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
    // 16-byte loads need 16-byte aligned rows (the scalar flavour handles everything else)
    const bool vec_a = (((uintptr_t)A & 15) == 0) && lda % 4 == 0, vec_b = (((uintptr_t)B & 15) == 0) && ldb % 4 == 0;
    const dim3 grid_a((unsigned)(kp / 64), (unsigned)(mp / 64)), grid_b((unsigned)(kp / 64), (unsigned)(np / 64));
    if (vec_a)
        split_bf16_kernel<true><<<grid_a, 256, 0, s>>>(A, lda, (int)M, (int)K, trans_a ? 1 : 0, a_hi, a_lo, (int)mp, (int)kp);
    else
        split_bf16_kernel<false><<<grid_a, 256, 0, s>>>(A, lda, (int)M, (int)K, trans_a ? 1 : 0, a_hi, a_lo, (int)mp, (int)kp);
    OPN_CUDA(cudaGetLastError());
    if (vec_b)
        split_bf16_kernel<true><<<grid_b, 256, 0, s>>>(B, ldb, (int)N, (int)K, trans_b ? 0 : 1, b_hi, b_lo, (int)np, (int)kp);
    else
        split_bf16_kernel<false><<<grid_b, 256, 0, s>>>(B, ldb, (int)N, (int)K, trans_b ? 0 : 1, b_hi, b_lo, (int)np, (int)kp);
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

    CUtensorMap map_ahi, map_alo, map_bhi, map_blo;
    int rc;
    if ((rc = make_operand_map(&map_ahi, a_hi, mp, kp)) != OPN_OK) return rc;
    if ((rc = make_operand_map(&map_alo, a_lo, mp, kp)) != OPN_OK) return rc;
    if ((rc = make_operand_map(&map_bhi, b_hi, np, kp)) != OPN_OK) return rc;
    if ((rc = make_operand_map(&map_blo, b_lo, np, kp)) != OPN_OK) return rc;

    TcParams p;
    p.C = C; p.bias = bias; p.ldc = ldc;
    p.M = (int)M; p.N = (int)N;
    p.tiles_m = gm; p.tiles_n = gn; p.splits = splits;
    p.nkb_total = nkb;
    p.kb_per_split = kb_per_split;
    p.alpha = alpha;
    p.beta_one = beta == 1.0f;
    p.relu = relu;
    p.atomic_out = splits > 1;
    p.single = current_precision() == OPN_PRECISION_16BIT ? 1 : 0;
    if (splits > 1 && !p.beta_one) {
        const long long n = M * N;
        zero_strided_tc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(C, ldc, (int)M, (int)N);
        OPN_CUDA(cudaGetLastError());
        count_launch();
    }
    const int n_work = gm * gn * splits;
    const int grid = n_work < sms ? n_work : sms;
    OPN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    gemm_tc_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(map_ahi, map_alo, map_bhi, map_blo, p);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

}  // namespace opn
