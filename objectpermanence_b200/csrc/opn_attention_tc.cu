// Fused (flash-style) self-attention on the 5th-generation tensor cores for transformer_lstm's encoder
// (nn.MultiheadAttention inside nn.TransformerEncoderLayer, baselines/learned_models.py:166-168,184), sm_100a.
//
//   ctx = softmax(Q K^T / sqrt(d)) V   per head over ONE sequence of S rows (S = B*T, head dimension d = 128)
//
// No [S,S] tensor exists anywhere: scores live in TMEM for one 128 x 64 tile at a time, the running softmax statistics and
// the running output in registers.  The backward pass recomputes the scores tile by tile from the saved row statistics
// (log-sum-exp) and is split in two kernels so that every accumulation stays inside one CTA (no atomics):
//   attn_bwd_q_kernel   CTA = 128-row query tile, loops over key tiles:  S, dP -> dS -> dQ += dS K
//   attn_bwd_kv_kernel  CTA = 128-row key tile,  loops over query tiles: S^T, dP^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q
// Arithmetic: every fp32 operand is a bf16 hi + lo pair (prepared once per call by attn_prep_kernel, row-major planes
// [row][128]); products are hi.hi + lo.hi + hi.lo into fp32 TMEM accumulators (one pass in the 1e-2 mode).  Operand tiles
// arrive by TMA (SWIZZLE_128B); the same row-major tile serves as a K-major operand (Q K^T: rows x d) and, through an
// MN-major descriptor, as the B operand of the products that contract over rows (P V, P^T dO, dS K, dS^T Q) -- no
// transposed copies.  P / dS never touch shared memory: the softmax threads (thread = tile row = TMEM lane) write them with
// tcgen05.st into tensor memory, from where the second product reads its A operand.
// Attention-weight dropout (train mode) uses the mask of opn_dropout: element (q, k) of head h is element q*S + k of
// the stream (seed, offset + h * ceil(S*S/4)).  The forward kernel generates it (extra warps, one tile ahead) and leaves it
// as one 64-bit word per (query, 64-key tile) in the workspace (23 MB at S = 9600, 2 heads); the backward kernels load
// the words (the key-tile kernel transposes them by shuffles) instead of running Philox again: regenerating the mask in
// both backward kernels had cost 0.40 ms of the 1.45 ms backward pass.
#include <stdlib.h>

#include <cuda_bf16.h>

#include "opn_tc_common.cuh"

namespace opn {

int current_precision();   // opn_api.cu

namespace {

constexpr int DH = 128;                 // head dimension
constexpr int TQ = 128;                 // rows of the CTA's resident tile (M of every MMA, TMEM lanes)
constexpr int TK = 64;                  // rows of a streamed tile (N of the score products, K of the second products)
constexpr int AT = 192;                 // warp 0: TMA, warp 1: MMA issue, warps 2-5: softmax threads (thread = resident row)
constexpr int BLK128 = TQ * 128;        // bytes of one [128 rows x 64 cols] bf16 block
constexpr int BLK64 = TK * 128;         // bytes of one [64 rows x 64 cols] bf16 block
constexpr int RES_TILE = 4 * BLK128;    // resident tile: 2 planes x 2 column blocks = 64 KB
constexpr int STR_TILE = 4 * BLK64;     // streamed tile: 32 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr long long kAttnTimeout = 4000000000LL;

struct AttnParams {
    const __nv_bfloat16* planes;   // operand planes (see plane_row)
    const float* ctx;        // [S, D]   forward output (backward: input)
    float* ctx_out;          // forward
    const float* dctx;       // [S, D]
    float* dqkv;             // [S, 3D]
    float* lse2;             // [heads][S_pad]  log2-domain log-sum-exp of the scaled scores
    float* delta;            // [heads][S_pad]  rowsum(dO * O)
    unsigned int* status;    // 4 words
    int S, S_pad, D, heads;
    float scale;             // 1 / sqrt(d)
    unsigned int drop_threshold;   // keep when word >= threshold; 0 = no dropout
    float drop_scale;
    unsigned long long seed, offset;
    // work split (see plan_split): CTAs [0, n_full) own a whole row of streamed tiles of one (resident tile, head); every
    // leftover (resident tile, head) is cut into `segs` segments of consecutive streamed tiles, one CTA each, whose partial
    // accumulators go to `part` and are merged by the combine kernels
    unsigned long long* keep;   // [head][key tile][S_pad]: dropout keep bits of 64 keys per query row, written by the forward
                                // kernel (train mode) and read by both backward kernels instead of re-running Philox
    int n_ktiles;
    float* part;             // [slot][2][128][128]
    float* part_ml;          // [slot][2][128]   forward: reference maximum and sum of exponentials of a segment
    int n_rtiles, n_full, segs, grid;
};

struct Work {
    int rt, h;       // resident tile and head
    int j0, nt;      // first streamed tile and their number
    int slot;        // partial-accumulator slot, -1: the CTA owns the whole row and writes the result itself
};
__device__ __forceinline__ Work my_work(const AttnParams& p, int n_tiles) {
    Work w;
    const int b = (int)blockIdx.x;
    int row;
    if (b < p.n_full) {
        row = b, w.j0 = 0, w.nt = n_tiles, w.slot = -1;
    } else {
        const int k = b - p.n_full, sg = k % p.segs;
        row = p.n_full + k / p.segs;
        w.j0 = (int)((long long)sg * n_tiles / p.segs);
        w.nt = (int)((long long)(sg + 1) * n_tiles / p.segs) - w.j0;
        w.slot = k;
    }
    w.rt = row % p.n_rtiles, w.h = row / p.n_rtiles;
    return w;
}

// ---- operand planes: [tensor t (0 Q', 1 K, 2 V, 3 dO)][plane (hi, lo)][head][S_pad][128] bf16 ---------------------------
__device__ __forceinline__ long long plane_row(const AttnParams& p, int t, int plane, int h, int row) {
    return (((long long)(t * 2 + plane) * p.heads + h) * p.S_pad) + row;
}

__global__ void __launch_bounds__(256) attn_prep_kernel(const float* __restrict__ src, long long ld, int col0, int t, float mul,
                                                        __nv_bfloat16* __restrict__ planes, int S, int S_pad, int heads) {
    // one thread = 8 consecutive columns of one (row, head); rows >= S are zero-filled
    const long long total = (long long)S_pad * heads * (DH / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % (DH / 8));
        const int h = (int)((i / (DH / 8)) % heads);
        const int row = (int)(i / ((DH / 8) * heads));
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row < S) {
            const float* s = src + (long long)row * ld + col0 + h * DH + 8 * c8;
            const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s + 4));
            x[0] = a.x * mul, x[1] = a.y * mul, x[2] = a.z * mul, x[3] = a.w * mul;
            x[4] = b.x * mul, x[5] = b.y * mul, x[6] = b.z * mul, x[7] = b.w * mul;
        }
        __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            hi[e] = __float2bfloat16_rn(x[e]);
            lo[e] = __float2bfloat16_rn(x[e] - __bfloat162float(hi[e]));
        }
        const long long r_hi = (((long long)(t * 2 + 0) * heads + h) * S_pad + row) * DH + 8 * c8;
        const long long r_lo = (((long long)(t * 2 + 1) * heads + h) * S_pad + row) * DH + 8 * c8;
        *reinterpret_cast<uint4*>(planes + r_hi) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(planes + r_lo) = *reinterpret_cast<const uint4*>(lo);
    }
}

// ---- descriptors ------------------------------------------------------------------------------------------------------
// MN-major SWIZZLE_128B operand: rows of the tile are the K index (128 bytes = 64 consecutive MN elements each, 8-row groups
// 1024 bytes apart = SBO); the next 64 MN elements are `lbo` bytes further (the second column block of the tile)
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t kIdescScore = tc::idesc_f16(TQ, TK, true);                       // [128 x 64], both operands K-major
constexpr uint32_t kIdescSecond = tc::idesc_f16(TQ, DH, true) | (1u << 16);         // [128 x 128], B operand MN-major

// bounded mbarrier wait: a broken pipeline reports instead of hanging the GPU
__device__ __forceinline__ bool await(uint64_t* bar, uint32_t parity, unsigned int* status, volatile int* abort_s) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0) {
            if (*abort_s) return false;
            if (clock64() - t0 > kAttnTimeout) {
                if (atomicCAS(status, 0u, 2u) == 0u) {
                    status[1] = blockIdx.x;
                    status[2] = 0u;
                    status[3] = threadIdx.x;
                }
                *abort_s = 1;
                return false;
            }
        }
    }
    return true;
}

__device__ __forceinline__ uint4 philox(unsigned long long ctr, unsigned long long seed) {
    unsigned int c0 = (unsigned int)ctr, c1 = (unsigned int)(ctr >> 32), c2 = 0u, c3 = 0u;
    unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// keep bits of the 64 consecutive elements i0 .. i0+63 of the dropout stream (bit j = element i0 + j is kept)
__device__ __forceinline__ unsigned long long keep_bits_row(unsigned long long i0, unsigned long long seed, unsigned long long offset,
                                                            unsigned int threshold) {
    unsigned long long bits = 0ull;
    if ((i0 & 3ull) == 0ull) {      // aligned (always when S % 4 == 0): exactly 16 blocks, one nibble each
        const unsigned long long b0 = offset + (i0 >> 2);
#pragma unroll 4
        for (int t = 0; t < 16; ++t) {
            const uint4 r = philox(b0 + (unsigned long long)t, seed);
            const unsigned long long nib = (unsigned long long)((r.x >= threshold ? 1u : 0u) | (r.y >= threshold ? 2u : 0u) |
                                                                (r.z >= threshold ? 4u : 0u) | (r.w >= threshold ? 8u : 0u));
            bits |= nib << (4 * t);
        }
        return bits;
    }
    const unsigned long long b0 = i0 >> 2, b1 = (i0 + 63) >> 2;
    for (unsigned long long b = b0; b <= b1; ++b) {
        const uint4 r = philox(offset + b, seed);
        const unsigned int w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long j = (long long)(b * 4 + e) - (long long)i0;
            if (j >= 0 && j < 64 && w[e] >= threshold) bits |= 1ull << j;
        }
    }
    return bits;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// (a, b) -> packed bf16 pair hi (a in the low half) and the pair of the residuals lo
__device__ __forceinline__ void split_bf16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
// 64 fp32 values of one tile row -> bf16 hi / lo planes in TENSOR MEMORY (lane = row, two K elements per 32-bit column:
// 32 columns per plane), the A operand of the second product.  No shared-memory tile, no swizzle, no proxy fence.
template <int PL>
__device__ __forceinline__ void store_row_tmem(uint32_t taddr, const float (&x)[TK]) {
    uint32_t hi[32], lo[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) split_bf16_pair(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
    uint32_t a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = hi[i], b[i] = hi[16 + i];
    tc::tmem_st16(taddr, a);
    tc::tmem_st16(taddr + 16, b);
    if (PL == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = lo[i], b[i] = lo[16 + i];
        tc::tmem_st16(taddr + 32, a);
        tc::tmem_st16(taddr + 48, b);
    }
    tc::tmem_st_wait();
}

// score-type product D[128 x 64] = A[128 x 128] . B[64 x 128]^T (both K-major, two 64-wide column blocks per plane)
template <int PASSES>
__device__ __forceinline__ void issue_score(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile) {
#pragma unroll
    for (int cb = 0; cb < 2; ++cb)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ah = tc::smem_desc_sw128(a_tile + cb * BLK128 + kk * 32), bh = tc::smem_desc_sw128(b_tile + cb * BLK64 + kk * 32);
            tc::umma_f16(d_tmem, ah, bh, kIdescScore, (cb > 0 || kk > 0) ? 1u : 0u);
            if (PASSES == 3) {
                const uint64_t al = tc::smem_desc_sw128(a_tile + 2 * BLK128 + cb * BLK128 + kk * 32);
                const uint64_t bl = tc::smem_desc_sw128(b_tile + 2 * BLK64 + cb * BLK64 + kk * 32);
                tc::umma_f16(d_tmem, al, bh, kIdescScore, 1u);
                tc::umma_f16(d_tmem, ah, bl, kIdescScore, 1u);
            }
        }
}
// the same product with the resident operand held in TENSOR MEMORY (64 columns per plane, lo plane 64 columns further):
// the 4 KB A slice no longer comes out of shared memory for every MMA (the score products were shared-memory bound:
// 64 clocks per MMA against 32 of math, profiles/r02_attention_phases.log)
template <int PASSES>
__device__ __forceinline__ void issue_score_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_tile) {
#pragma unroll
    for (int cb = 0; cb < 2; ++cb)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t a = a_tmem + cb * 32 + kk * 8;
            const uint64_t bh = tc::smem_desc_sw128(b_tile + cb * BLK64 + kk * 32);
            tc::umma_f16_ts(d_tmem, a, bh, kIdescScore, (cb > 0 || kk > 0) ? 1u : 0u);
            if (PASSES == 3) {
                const uint64_t bl = tc::smem_desc_sw128(b_tile + 2 * BLK64 + cb * BLK64 + kk * 32);
                tc::umma_f16_ts(d_tmem, a + 64, bh, kIdescScore, 1u);
                tc::umma_f16_ts(d_tmem, a, bl, kIdescScore, 1u);
            }
        }
}
// one row of a [rows x 128] operand (bf16 hi / lo planes in global memory) -> tensor memory, 64 columns per plane
template <int PL>
__device__ __forceinline__ void load_row_to_tmem(uint32_t taddr, const __nv_bfloat16* planes, long long row_hi, long long row_lo) {
#pragma unroll
    for (int pl = 0; pl < PL; ++pl) {
        const uint4* src = reinterpret_cast<const uint4*>(planes + (pl ? row_lo : row_hi) * DH);
#pragma unroll
        for (int c = 0; c < 4; ++c) {      // 16 columns = 32 bf16 = 4 x 16 bytes
            const uint4 v0 = __ldg(src + 4 * c), v1 = __ldg(src + 4 * c + 1), v2 = __ldg(src + 4 * c + 2), v3 = __ldg(src + 4 * c + 3);
            const uint32_t w[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
            tc::tmem_st16(taddr + pl * 64 + c * 16, w);
        }
    }
    tc::tmem_st_wait();
}

// second-type product D[128 x 128] (+)= A[128 x 64] . B[64 x 128]: A = the P / dS planes in tensor memory (8 columns per
// 16-wide k-step, lo plane 32 columns further), B = a streamed row-major tile read through an MN-major descriptor
template <int PASSES>
__device__ __forceinline__ void issue_second(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_tile, bool accumulate) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint64_t bh = smem_desc_mn_sw128(b_tile + kk * 2048, BLK64);
        tc::umma_f16_ts(d_tmem, a_tmem + kk * 8, bh, kIdescSecond, (accumulate || kk > 0) ? 1u : 0u);
        if (PASSES == 3) {
            const uint64_t bl = smem_desc_mn_sw128(b_tile + 2 * BLK64 + kk * 2048, BLK64);
            tc::umma_f16_ts(d_tmem, a_tmem + 32 + kk * 8, bh, kIdescSecond, 1u);
            tc::umma_f16_ts(d_tmem, a_tmem + kk * 8, bl, kIdescSecond, 1u);
        }
    }
}

// TMA: a [rows x 128] tile = 2 planes x 2 column blocks of [rows x 64]
template <int PL>
__device__ __forceinline__ void load_tile(uint32_t dst, const CUtensorMap* map, const AttnParams& p, int t, int h, int row0, int blk_bytes,
                                          uint64_t* bar) {
#pragma unroll
    for (int pl = 0; pl < PL; ++pl)
#pragma unroll
        for (int cb = 0; cb < 2; ++cb)
            tc::tma_load_2d(dst + (pl * 2 + cb) * blk_bytes, map, cb * 64, (int)plane_row(p, t, pl, h, row0), bar);
}

// per-role cycle counters of CTA (0,0) (development build: python -m objectpermanence_b200.build --phases; tools/attn_phases.py)
#ifdef OPN_LSTM_PHASES
#define APH_DECL long long aph[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long aph_last = clock64();
#define APH(i) do { const long long n__ = clock64(); aph[i] += n__ - aph_last; aph_last = n__; } while (0)
#define APH_STORE(status, role) do { if (blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long* o__ = reinterpret_cast<unsigned long long*>(status) + 32 + 8 * (role); \
        for (int i__ = 0; i__ < 8; ++i__) o__[i__] = (unsigned long long)aph[i__]; } } while (0)
#else
#define APH_DECL
#define APH(i)
#define APH_STORE(status, role)
#endif

struct Shared {
    uint64_t res_full, str_full[2], str_empty[2], kv_full[3], kv_empty[3], v_full, v_empty, s_full[2], s_empty[2], p_full[2], pv_full[2], done;
    uint64_t rng_full[2], rng_empty[2];
    uint32_t tmem_base;
    int abort_flag;
};
// Train mode: the Philox blocks of a tile's dropout mask (17 per row: ~2,000 integer instructions) are generated by extra
// warps one tile ahead and handed over through shared memory, so that they fill the issue slots the latency-bound softmax
// warps leave idle instead of doubling their work (forward 0.56 -> 1.05 ms with the mask generated in line).
constexpr int rng_warps(bool drop, bool fwd) { return drop ? (fwd ? 4 : 2) : 0; }
// the query-tile backward has the registers for four generator warps (168 per thread), the key-tile one (234) has not
constexpr int rng_warps_q(bool drop) { return drop ? 4 : 0; }

// ======================================================= forward =======================================================
// TMEM: S tiles 0-63 / 64-127 (double buffered), O 128-255 (accumulated over all key tiles), P planes 256-319 / 320-383,
// the Q' tile (A operand of every score product) 384-511.  Shared memory: three (K, V) stages.
// O stays in tensor memory for the whole row of key tiles: the softmax threads keep a reference maximum per row and only
// rescale O (tcgen05.ld / st) when the row maximum has grown by more than 2^8 since (P <= 256 is harmless in bf16 hi/lo).
template <int PASSES, bool DROP>
__global__ void __launch_bounds__(AT + 32 * rng_warps(DROP, true), 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap map128,
                                                         const __grid_constant__ CUtensorMap map64, const AttnParams p) {
    constexpr int PL = PASSES == 3 ? 2 : 1;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) Shared sh;
    __shared__ unsigned long long keep_bits[2][TQ];     // dropout keep bits of the current / next tile, one word per query row
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t kv_s = base;      // K/V stages: NST x (K tile, V tile)
    constexpr int NST = 3;
    const Work wk = my_work(p, (p.S + TK - 1) / TK);
    const int qt = wk.rt, h = wk.h, j0 = wk.j0;
    const int n_tiles = wk.nt;      // streamed tiles of THIS CTA: key tiles j0 .. j0 + n_tiles - 1

    if (tid == 0) {
        mbar_init(&sh.res_full, 4);
        for (int i = 0; i < 3; ++i) {
            mbar_init(&sh.kv_full[i], 1);
            mbar_init(&sh.kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sh.s_full[i], 1);
            mbar_init(&sh.s_empty[i], 4);
            mbar_init(&sh.p_full[i], 4);
            mbar_init(&sh.pv_full[i], 1);
            mbar_init(&sh.rng_full[i], rng_warps(true, true));
            mbar_init(&sh.rng_empty[i], 4);
        }
        mbar_fence_init();
        sh.abort_flag = 0;
    }
    if (warp == 1) tc::tmem_alloc(&sh.tmem_base, 512);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = sh.tmem_base;
    volatile int* abort_s = &sh.abort_flag;

    if (warp >= 6) {
        // ---- mask generators (train mode only): row r of tile j, one tile ahead of the softmax threads ------------------
        const int r = (warp - 6) * 32 + lane;
        const unsigned long long drop_off = p.offset + (unsigned long long)h * (((unsigned long long)p.S * p.S + 3ull) >> 2);
        for (int j = 0; j < n_tiles; ++j) {
            const int b = j & 1;
            if (j >= 2 && !await(&sh.rng_empty[b], ((uint32_t)(j >> 1) - 1u) & 1u, p.status, abort_s)) break;
            const unsigned long long bits = keep_bits_row((unsigned long long)(qt * TQ + r) * p.S + (unsigned long long)(j0 + j) * TK, p.seed,
                                                          drop_off, p.drop_threshold);
            keep_bits[b][r] = bits;
            p.keep[((size_t)h * p.n_ktiles + (j0 + j)) * p.S_pad + qt * TQ + r] = bits;      // for the backward kernels
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh.rng_full[b]);
        }
    } else if (warp == 0) {
        if (lane == 0) {
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % NST;
                if (j >= NST && !await(&sh.kv_empty[st], ((uint32_t)(j / NST) - 1u) & 1u, p.status, abort_s)) break;
                mbar_arrive_expect_tx(&sh.kv_full[st], 2 * PL * 2 * BLK64);
                load_tile<PL>(kv_s + st * 2 * STR_TILE, &map64, p, 1, h, (j0 + j) * TK, BLK64, &sh.kv_full[st]);
                load_tile<PL>(kv_s + st * 2 * STR_TILE + STR_TILE, &map64, p, 2, h, (j0 + j) * TK, BLK64, &sh.kv_full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            bool ok = await(&sh.res_full, 0, p.status, abort_s) && await(&sh.kv_full[0], 0, p.status, abort_s);
            if (ok) {
                tc::fence_after();
                issue_score_ts<PASSES>(tmem, tmem + 384, kv_s);
                tc::umma_commit(&sh.s_full[0]);
            }
            APH_DECL
            for (int j = 0; j < n_tiles && ok; ++j) {
                const int st = j % NST;
                if (j + 1 < n_tiles) {      // the next score tile is queued before this tile's P is ready
                    const int sn = (j + 1) % NST, sb = (j + 1) & 1;
                    ok = await(&sh.kv_full[sn], (uint32_t)((j + 1) / NST) & 1u, p.status, abort_s);
                    APH(0);
                    if (ok && j + 1 >= 2) ok = await(&sh.s_empty[sb], ((uint32_t)((j + 1) >> 1) - 1u) & 1u, p.status, abort_s);
                    if (!ok) break;
                    APH(1);
                    tc::fence_after();
                    issue_score_ts<PASSES>(tmem + sb * TK, tmem + 384, kv_s + sn * 2 * STR_TILE);
                    tc::umma_commit(&sh.s_full[sb]);
                    APH(2);
                }
                if (!await(&sh.p_full[j & 1], (uint32_t)(j >> 1) & 1u, p.status, abort_s)) break;
                APH(3);
                tc::fence_after();
                issue_second<PASSES>(tmem + 128, tmem + 256 + (j & 1) * 64, kv_s + st * 2 * STR_TILE + STR_TILE, j > 0);
                tc::umma_commit(&sh.pv_full[j & 1]);
                tc::umma_commit(&sh.kv_empty[st]);
                APH(4);
            }
            APH_STORE(p.status, 1);
        }
    } else {
        // ---- softmax threads: thread = query row -------------------------------------------------------------------------
        const int qd = warp & 3, r = qd * 32 + lane;
        const int q = qt * TQ + r;
        const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16);
        float m = -INFINITY, l = 0.0f;       // m: the reference maximum the stored exponentials are relative to
        // this thread's row of Q' goes to tensor memory once (plain loads, tcgen05.st): the A operand of every score product
        load_row_to_tmem<PL>(lane_addr + 384, p.planes, plane_row(p, 0, 0, h, q), plane_row(p, 0, 1, h, q));
        tc::fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&sh.res_full);
        bool ok = true;
        APH_DECL
        for (int j = 0; j < n_tiles; ++j) {
            const int sb = j & 1;
            APH(0);
            if (!await(&sh.s_full[sb], (uint32_t)(j >> 1) & 1u, p.status, abort_s)) { ok = false; break; }
            APH(1);
            tc::fence_after();
            float s[TK];
            {
                uint32_t x0[32], x1[32];
                tc::tmem_ld32(lane_addr + sb * TK, x0);
                tc::tmem_ld32(lane_addr + sb * TK + 32, x1);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(x0[i]), s[32 + i] = __uint_as_float(x1[i]);
            }
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh.s_empty[sb]);
            APH(2);
            const int kvalid = p.S - (j0 + j) * TK;      // keys of this tile that exist
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < TK; ++i) {
                if (i >= kvalid) s[i] = -INFINITY;
                mx = fmaxf(mx, s[i]);
            }
            // move the reference maximum only when the row maximum has outgrown it by 2^8 (always on the first tile)
            const bool grow = mx > m + 8.0f;
            if (__any_sync(0xffffffffu, grow)) {
                const float m_new = grow ? mx : m;
                const float corr = ex2(m - m_new);      // 0 on the first tile (m = -inf)
                l *= corr;
                m = m_new;
                if (j > 0) {
                    // every P V product issued so far must have landed in O before it is rescaled
                    if (!await(&sh.pv_full[(j - 1) & 1], (uint32_t)((j - 1) >> 1) & 1u, p.status, abort_s)) { ok = false; break; }
                    tc::fence_after();
#pragma unroll
                    for (int c = 0; c < DH; c += 32) {
                        uint32_t x[32];
                        tc::tmem_ld32(lane_addr + 128 + c, x);
                        tc::tmem_ld_wait();
                        uint32_t y0[16], y1[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            y0[i] = __float_as_uint(__uint_as_float(x[i]) * corr);
                            y1[i] = __float_as_uint(__uint_as_float(x[16 + i]) * corr);
                        }
                        tc::tmem_st16(lane_addr + 128 + c, y0);
                        tc::tmem_st16(lane_addr + 128 + c + 16, y1);
                    }
                    tc::tmem_st_wait();
                    tc::fence_before();
                }
            }
            float sum = 0.0f;
#pragma unroll
            for (int i = 0; i < TK; ++i) {
                s[i] = ex2(s[i] - m);
                sum += s[i];
            }
            l += sum;
            if (DROP) {
                if (!await(&sh.rng_full[sb], (uint32_t)(j >> 1) & 1u, p.status, abort_s)) { ok = false; break; }
                const unsigned long long bits = keep_bits[sb][r];
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh.rng_empty[sb]);
#pragma unroll
                for (int i = 0; i < TK; ++i) s[i] = ((bits >> i) & 1ull) ? s[i] * p.drop_scale : 0.0f;
            }
            APH(3);
            // the P planes of this parity were last read by the P V product of tile j-2
            if (j >= 2 && !await(&sh.pv_full[j & 1], ((uint32_t)(j >> 1) - 1u) & 1u, p.status, abort_s)) { ok = false; break; }
            APH(4);
            store_row_tmem<PL>(lane_addr + 256 + (j & 1) * 64, s);
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh.p_full[j & 1]);
            APH(5);
        }
        if (tid == 64) APH_STORE(p.status, 0);
        if (ok && await(&sh.pv_full[(n_tiles - 1) & 1], (uint32_t)((n_tiles - 1) >> 1) & 1u, p.status, abort_s)) {
            tc::fence_after();
            const bool whole = wk.slot < 0;      // a segment leaves its unnormalised output, reference maximum and sum
            const float inv = whole ? 1.0f / l : 1.0f;
            float* dst = whole ? p.ctx_out + (size_t)q * p.D + h * DH : p.part + ((size_t)wk.slot * 2 * TQ + r) * DH;
#pragma unroll
            for (int c = 0; c < DH; c += 32) {
                uint32_t x[32];
                tc::tmem_ld32(lane_addr + 128 + c, x);
                tc::tmem_ld_wait();
                if (q < p.S || !whole) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(dst + c + i) =
                            make_float4(__uint_as_float(x[i]) * inv, __uint_as_float(x[i + 1]) * inv, __uint_as_float(x[i + 2]) * inv,
                                        __uint_as_float(x[i + 3]) * inv);
                }
            }
            tc::fence_before();
            if (whole) {
                p.lse2[(size_t)h * p.S_pad + q] = (q < p.S) ? m + log2f(l) : INFINITY;
            } else {
                p.part_ml[((size_t)wk.slot * 2 + 0) * TQ + r] = m;
                p.part_ml[((size_t)wk.slot * 2 + 1) * TQ + r] = l;
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// ================================================= backward: dQ (query-tile CTAs) =======================================
// TMEM: (S, dP) pairs 0-127 / 128-255 (double buffered), dQ 256-383, dS planes 384-447 / 448-511.
// Shared memory: Q', dO resident; K double buffered (read by the score product and, a phase later, by dQ += dS K), V single.
template <int PASSES, bool DROP>
__global__ void __launch_bounds__(AT + 32 * rng_warps_q(DROP), 1) attn_bwd_q_kernel(const __grid_constant__ CUtensorMap map128,
                                                           const __grid_constant__ CUtensorMap map64, const AttnParams p) {
    constexpr int PL = PASSES == 3 ? 2 : 1;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) Shared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ unsigned long long keep_bits[TQ];     // dropout keep bits of one tile (single buffer: 224 KB of tiles leave 3 KB)
    const uint32_t q_s = base, do_s = base + RES_TILE, k_s = base + 2 * RES_TILE, v_s = k_s + 2 * STR_TILE;
    const Work wk = my_work(p, (p.S + TK - 1) / TK);
    const int qt = wk.rt, h = wk.h, j0 = wk.j0;
    const int n_tiles = wk.nt;      // key tiles j0 .. j0 + n_tiles - 1

    if (tid == 0) {
        mbar_init(&sh.res_full, 1);
        mbar_init(&sh.v_full, 1);
        mbar_init(&sh.v_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sh.str_full[i], 1);
            mbar_init(&sh.str_empty[i], 1);
            mbar_init(&sh.s_full[i], 1);
            mbar_init(&sh.s_empty[i], 4);
            mbar_init(&sh.p_full[i], 4);
            mbar_init(&sh.pv_full[i], 1);
            mbar_init(&sh.rng_full[i], rng_warps_q(true));
            mbar_init(&sh.rng_empty[i], 4);
        }
        mbar_fence_init();
        sh.abort_flag = 0;
    }
    if (warp == 1) tc::tmem_alloc(&sh.tmem_base, 512);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = sh.tmem_base;
    volatile int* abort_s = &sh.abort_flag;

    if (warp >= 6) {
        // ---- mask loaders (train mode only): the keep bits the forward kernel left, one row per thread, one tile ahead -------
        for (int j = 0; j < n_tiles; ++j) {
            const unsigned long long bits = __ldcs(p.keep + ((size_t)h * p.n_ktiles + (j0 + j)) * p.S_pad + qt * TQ + (warp - 6) * 32 + lane);
            if (j >= 1 && !await(&sh.rng_empty[0], (uint32_t)(j - 1) & 1u, p.status, abort_s)) break;     // tile j-1's bits have been read
            keep_bits[(warp - 6) * 32 + lane] = bits;
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh.rng_full[0]);
        }
    } else if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&sh.res_full, 2 * PL * 2 * BLK128);
            load_tile<PL>(q_s, &map128, p, 0, h, qt * TQ, BLK128, &sh.res_full);
            load_tile<PL>(do_s, &map128, p, 3, h, qt * TQ, BLK128, &sh.res_full);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                // K stage st was last read by dQ += dS K of tile j-2; V (single) by the score products of tile j-1
                if (j >= 2 && !await(&sh.str_empty[st], ((uint32_t)(j >> 1) - 1u) & 1u, p.status, abort_s)) break;
                mbar_arrive_expect_tx(&sh.str_full[st], PL * 2 * BLK64);
                load_tile<PL>(k_s + st * STR_TILE, &map64, p, 1, h, (j0 + j) * TK, BLK64, &sh.str_full[st]);
                if (j >= 1 && !await(&sh.v_empty, (uint32_t)(j - 1) & 1u, p.status, abort_s)) break;
                mbar_arrive_expect_tx(&sh.v_full, PL * 2 * BLK64);
                load_tile<PL>(v_s, &map64, p, 2, h, (j0 + j) * TK, BLK64, &sh.v_full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            bool ok = await(&sh.res_full, 0, p.status, abort_s);
            auto scores = [&](int j) {      // S = Q' K^T and dP = dO V^T of tile j into TMEM pair j & 1
                const int st = j & 1;
                if (!await(&sh.str_full[st], (uint32_t)(j >> 1) & 1u, p.status, abort_s)) return false;
                if (!await(&sh.v_full, (uint32_t)j & 1u, p.status, abort_s)) return false;
                if (j >= 2 && !await(&sh.s_empty[st], ((uint32_t)(j >> 1) - 1u) & 1u, p.status, abort_s)) return false;
                tc::fence_after();
                issue_score<PASSES>(tmem + st * 128, q_s, k_s + st * STR_TILE);
                issue_score<PASSES>(tmem + st * 128 + TK, do_s, v_s);
                tc::umma_commit(&sh.s_full[st]);
                tc::umma_commit(&sh.v_empty);
                return true;
            };
            if (ok) ok = scores(0);
            for (int j = 0; j < n_tiles && ok; ++j) {
                if (j + 1 < n_tiles && !(ok = scores(j + 1))) break;
                if (!await(&sh.p_full[j & 1], (uint32_t)(j >> 1) & 1u, p.status, abort_s)) break;
                tc::fence_after();
                issue_second<PASSES>(tmem + 256, tmem + 384 + (j & 1) * 64, k_s + (j & 1) * STR_TILE, j > 0);   // dQ += dS K
                tc::umma_commit(&sh.pv_full[j & 1]);
                tc::umma_commit(&sh.str_empty[j & 1]);
            }
        }
    } else {
        const int qd = warp & 3, r = qd * 32 + lane;
        const int q = qt * TQ + r;
        const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16);
        // delta = rowsum(dO * O) of this query row (also left in global memory for the key-tile kernel)
        float delta = 0.0f, lse = INFINITY;
        if (q < p.S) {
            const float* a = p.dctx + (size_t)q * p.D + h * DH;
            const float* b = p.ctx + (size_t)q * p.D + h * DH;
#pragma unroll 8
            for (int i = 0; i < DH; i += 4) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(a + i)), y = __ldg(reinterpret_cast<const float4*>(b + i));
                delta += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
            }
            lse = p.lse2[(size_t)h * p.S_pad + q];
        }
        p.delta[(size_t)h * p.S_pad + q] = delta;
        bool ok = true;
        for (int j = 0; j < n_tiles; ++j) {
            const int sb = j & 1;
            if (!await(&sh.s_full[sb], (uint32_t)(j >> 1) & 1u, p.status, abort_s)) { ok = false; break; }
            tc::fence_after();
            float ds[TK];
            {
                uint32_t s0[32], s1[32], d0[32], d1[32];
                tc::tmem_ld32(lane_addr + sb * 128, s0);
                tc::tmem_ld32(lane_addr + sb * 128 + 32, s1);
                tc::tmem_ld32(lane_addr + sb * 128 + TK, d0);
                tc::tmem_ld32(lane_addr + sb * 128 + TK + 32, d1);
                tc::tmem_ld_wait();
                tc::fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh.s_empty[sb]);
                const int kvalid = p.S - (j0 + j) * TK;
                unsigned long long bits = ~0ull;
                if (DROP) {
                    if (!await(&sh.rng_full[0], (uint32_t)j & 1u, p.status, abort_s)) { ok = false; break; }
                    bits = keep_bits[r];
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&sh.rng_empty[0]);
                }
#pragma unroll
                for (int i = 0; i < TK; ++i) {
                    const float sv = __uint_as_float(i < 32 ? s0[i & 31] : s1[i & 31]);
                    float dp = __uint_as_float(i < 32 ? d0[i & 31] : d1[i & 31]);
                    if (DROP) dp = ((bits >> i) & 1ull) ? dp * p.drop_scale : 0.0f;
                    const float pv = (i < kvalid) ? ex2(sv - lse) : 0.0f;
                    ds[i] = pv * (dp - delta);
                }
            }
            // the dS planes of this parity were last read by dQ += dS K of tile j-2
            if (j >= 2 && !await(&sh.pv_full[sb], ((uint32_t)(j >> 1) - 1u) & 1u, p.status, abort_s)) { ok = false; break; }
            store_row_tmem<PL>(lane_addr + 384 + sb * 64, ds);
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh.p_full[sb]);
        }
        if (ok && await(&sh.pv_full[(n_tiles - 1) & 1], (uint32_t)((n_tiles - 1) >> 1) & 1u, p.status, abort_s)) {
            tc::fence_after();
            const bool whole = wk.slot < 0;
            const float mul = whole ? p.scale : 1.0f;
            float* dst = whole ? p.dqkv + (size_t)q * (3 * p.D) + h * DH : p.part + ((size_t)wk.slot * 2 * TQ + r) * DH;
#pragma unroll
            for (int c = 0; c < DH; c += 32) {
                uint32_t x[32];
                tc::tmem_ld32(lane_addr + 256 + c, x);
                tc::tmem_ld_wait();
                if (q < p.S || !whole) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(dst + c + i) =
                            make_float4(__uint_as_float(x[i]) * mul, __uint_as_float(x[i + 1]) * mul,
                                        __uint_as_float(x[i + 2]) * mul, __uint_as_float(x[i + 3]) * mul);
                }
            }
            tc::fence_before();
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// ================================================= backward: dK, dV (key-tile CTAs) =====================================
// TMEM: S^T 0-63, dP^T 64-127, dV 128-255, dK 256-383, P^T planes 384-447, dS^T planes 448-511.
// Shared memory: K, V resident; Q' and dO tiles double buffered (read by the score products and by dK / dV).
template <int PASSES, bool DROP>
__global__ void __launch_bounds__(AT + 32 * rng_warps(DROP, false), 1) attn_bwd_kv_kernel(const __grid_constant__ CUtensorMap map128,
                                                            const __grid_constant__ CUtensorMap map64, const AttnParams p) {
    constexpr int PL = PASSES == 3 ? 2 : 1;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) Shared sh;
    __shared__ float lse_s[TK], delta_s[TK];          // row statistics of the current query tile (two named barriers per tile)
    __shared__ unsigned long long keep_bits[TQ];     // single buffer (see attn_bwd_q_kernel)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t k_s = base, v_s = base + RES_TILE, qd_s = base + 2 * RES_TILE;     // stages: (Q' tile, dO tile) x 1.5: see below
    // 224 KB: K 64 + V 64 + Q' 2 x 32 + dO 1 x 32 -- Q' double buffered (the TMA of the next tile overlaps), dO single
    const uint32_t do_s = qd_s + 2 * STR_TILE;
    const Work wk = my_work(p, (p.S + TK - 1) / TK);
    const int kt = wk.rt, h = wk.h, i0 = wk.j0;
    const int n_tiles = wk.nt;     // query tiles of 64 rows: i0 .. i0 + n_tiles - 1

    if (tid == 0) {
        mbar_init(&sh.res_full, 1);
        mbar_init(&sh.v_full, 1);
        mbar_init(&sh.v_empty, 1);
        mbar_init(&sh.done, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sh.str_full[i], 1);
            mbar_init(&sh.str_empty[i], 1);
        }
        mbar_init(&sh.s_full[0], 1);
        mbar_init(&sh.s_empty[0], 4);
        mbar_init(&sh.p_full[0], 4);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sh.rng_full[i], rng_warps(true, false));
            mbar_init(&sh.rng_empty[i], 4);
        }
        mbar_fence_init();
        sh.abort_flag = 0;
    }
    if (warp == 1) tc::tmem_alloc(&sh.tmem_base, 512);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = sh.tmem_base;
    volatile int* abort_s = &sh.abort_flag;

    if (warp >= 6) {
        // ---- mask loaders (train mode only): the forward kernel left one word per (query, 64-key tile); this CTA needs, per key
        // row, the bits of the 64 queries of the tile: a 64 x 64 bit transpose per half of the resident tile, by shuffles
        // (a lane holds the words of queries lane and lane + 32).  Two key rows per thread: lt and lt + 64.
        const int lt = (warp - 6) * 32 + lane;
        for (int i = 0; i < n_tiles; ++i) {
            unsigned long long out[2];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                unsigned long long wa = 0ull, wb = 0ull;
                if (kt * 2 + half < p.n_ktiles) {
                    const unsigned long long* src = p.keep + ((size_t)h * p.n_ktiles + (kt * 2 + half)) * p.S_pad + (size_t)(i0 + i) * TK;
                    wa = __ldcs(src + lane), wb = __ldcs(src + lane + 32);
                }
                unsigned long long bits = 0ull;
#pragma unroll 8
                for (int c = 0; c < 32; ++c) {
                    const unsigned long long x = __shfl_sync(0xffffffffu, wa, c), y = __shfl_sync(0xffffffffu, wb, c);
                    bits |= ((x >> lt) & 1ull) << c;
                    bits |= ((y >> lt) & 1ull) << (c + 32);
                }
                out[half] = bits;
            }
            if (i >= 1 && !await(&sh.rng_empty[0], (uint32_t)(i - 1) & 1u, p.status, abort_s)) break;     // tile i-1's bits have been read
            keep_bits[(warp - 6) * 32 + lane] = out[0];
            keep_bits[(warp - 6) * 32 + lane + 64] = out[1];
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh.rng_full[0]);
        }
    } else if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(&sh.res_full, 2 * PL * 2 * BLK128);
            load_tile<PL>(k_s, &map128, p, 1, h, kt * TQ, BLK128, &sh.res_full);
            load_tile<PL>(v_s, &map128, p, 2, h, kt * TQ, BLK128, &sh.res_full);
            for (int i = 0; i < n_tiles; ++i) {
                const int st = i & 1;
                if (i >= 2 && !await(&sh.str_empty[st], ((uint32_t)(i >> 1) - 1u) & 1u, p.status, abort_s)) break;
                mbar_arrive_expect_tx(&sh.str_full[st], PL * 2 * BLK64);
                load_tile<PL>(qd_s + st * STR_TILE, &map64, p, 0, h, (i0 + i) * TK, BLK64, &sh.str_full[st]);
                if (i >= 1 && !await(&sh.v_empty, (uint32_t)(i - 1) & 1u, p.status, abort_s)) break;     // dV += P^T dO of tile i-1 done
                mbar_arrive_expect_tx(&sh.v_full, PL * 2 * BLK64);
                load_tile<PL>(do_s, &map64, p, 3, h, (i0 + i) * TK, BLK64, &sh.v_full);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // Order per query tile i:  dP^T(i) | S^T(i+1) | dV(i), dK(i).  S^T of the NEXT tile is queued as soon as the
            // softmax threads hold S^T(i) and dP^T(i) in registers, so the tensor core works through it while they form
            // P^T / dS^T of tile i (the serial order S^T, dP^T -> softmax -> dV, dK left it idle 2,500 of 7,200 clocks per
            // tile).  dP^T(i+1) has to wait for the single dO buffer, i.e. for dV(i).
            bool ok = await(&sh.res_full, 0, p.status, abort_s) && await(&sh.str_full[0], 0, p.status, abort_s);
            if (ok) {
                tc::fence_after();
                issue_score<PASSES>(tmem, k_s, qd_s);     // S^T(0) = K Q'^T
            }
            APH_DECL
            for (int i = 0; i < n_tiles && ok; ++i) {
                const int st = i & 1;
                if (!await(&sh.v_full, (uint32_t)i & 1u, p.status, abort_s)) break;
                APH(0);
                tc::fence_after();
                issue_score<PASSES>(tmem + TK, v_s, do_s);                // dP^T(i) = V dO^T
                tc::umma_commit(&sh.s_full[0]);                          // ... and S^T(i), issued earlier
                APH(1);
                if (i + 1 < n_tiles) {
                    ok = await(&sh.str_full[st ^ 1], (uint32_t)((i + 1) >> 1) & 1u, p.status, abort_s) &&
                         await(&sh.s_empty[0], (uint32_t)i & 1u, p.status, abort_s);      // S^T(i), dP^T(i) are in registers
                    if (!ok) break;
                    tc::fence_after();
                    issue_score<PASSES>(tmem, k_s, qd_s + (st ^ 1) * STR_TILE);     // S^T(i+1)
                }
                APH(2);
                if (!await(&sh.p_full[0], (uint32_t)i & 1u, p.status, abort_s)) break;           // P^T and dS^T planes written
                APH(3);
                tc::fence_after();
                issue_second<PASSES>(tmem + 128, tmem + 384, do_s, i > 0);                       // dV += P^T dO
                tc::umma_commit(&sh.v_empty);
                issue_second<PASSES>(tmem + 256, tmem + 448, qd_s + st * STR_TILE, i > 0);       // dK += dS^T Q'
                tc::umma_commit(&sh.str_empty[st]);
                tc::umma_commit(&sh.done);
                APH(4);
            }
            APH_STORE(p.status, 3);
        }
    } else {
        const int qd = warp & 3, r = qd * 32 + lane;
        const int key = kt * TQ + r;
        const int et = tid - 64;      // 0..127 among the softmax threads
        const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16);
        bool dead = false;     // never leave the others alone at the named barrier below: keep stepping, the waits return at once
        APH_DECL
        for (int i = 0; i < n_tiles; ++i) {
            // row statistics of the 64 queries of this tile (the named barriers order fill and use)
            {
                const int qq = (i0 + i) * TK + (et & 63);
                const float v = (et < 64) ? p.lse2[(size_t)h * p.S_pad + qq] : p.delta[(size_t)h * p.S_pad + qq];
                if (i > 0) asm volatile("bar.sync 2, 128;" ::: "memory");     // everybody is done with the previous tile's values
                if (et < 64) lse_s[et] = v; else delta_s[et - 64] = v;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            APH(0);
            if (!dead && !await(&sh.s_full[0], (uint32_t)i & 1u, p.status, abort_s)) dead = true;
            APH(1);
            tc::fence_after();
            float pt[TK], ds[TK];
            {
                uint32_t s0[32], s1[32], d0[32], d1[32];
                tc::tmem_ld32(lane_addr, s0);
                tc::tmem_ld32(lane_addr + 32, s1);
                tc::tmem_ld32(lane_addr + TK, d0);
                tc::tmem_ld32(lane_addr + TK + 32, d1);
                tc::tmem_ld_wait();
                tc::fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sh.s_empty[0]);
                const bool kvalid = key < p.S;
                unsigned long long bits = ~0ull;
                if (DROP) {
                    if (!dead && !await(&sh.rng_full[0], (uint32_t)i & 1u, p.status, abort_s)) dead = true;
                    bits = keep_bits[r];
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&sh.rng_empty[0]);
                }
#pragma unroll
                for (int c = 0; c < TK; ++c) {
                    const float sv = __uint_as_float(c < 32 ? s0[c & 31] : s1[c & 31]);
                    float dp = __uint_as_float(c < 32 ? d0[c & 31] : d1[c & 31]);
                    float pv = kvalid ? ex2(sv - lse_s[c]) : 0.0f;      // lse = +inf for queries past the end
                    float pd = pv;
                    if (DROP) {
                        const bool keep = ((bits >> c) & 1ull) != 0ull;
                        dp = keep ? dp * p.drop_scale : 0.0f;
                        pd = keep ? pv * p.drop_scale : 0.0f;
                    }
                    pt[c] = pd;
                    ds[c] = pv * (dp - delta_s[c]);
                }
            }
            APH(2);
            // the planes were last read by the dV / dK products of tile i-1
            if (!dead && i >= 1 && !await(&sh.done, (uint32_t)(i - 1) & 1u, p.status, abort_s)) dead = true;
            APH(3);
            store_row_tmem<PL>(lane_addr + 384, pt);
            store_row_tmem<PL>(lane_addr + 448, ds);
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh.p_full[0]);
            APH(4);
        }
        if (tid == 64) APH_STORE(p.status, 2);
        if (!dead && await(&sh.done, (uint32_t)(n_tiles - 1) & 1u, p.status, abort_s)) {
            tc::fence_after();
            const bool whole = wk.slot < 0;
            float* dv = whole ? p.dqkv + (size_t)key * (3 * p.D) + 2 * p.D + h * DH : p.part + ((size_t)wk.slot * 2 * TQ + r) * DH;
            float* dk = whole ? p.dqkv + (size_t)key * (3 * p.D) + p.D + h * DH : p.part + (((size_t)wk.slot * 2 + 1) * TQ + r) * DH;
            const float kmul = whole ? 1.0f / kLog2e : 1.0f;     // Q' carries scale * log2(e)
#pragma unroll
            for (int c = 0; c < DH; c += 32) {
                uint32_t x[32], y[32];
                tc::tmem_ld32(lane_addr + 128 + c, x);
                tc::tmem_ld32(lane_addr + 256 + c, y);
                tc::tmem_ld_wait();
                if (key < p.S || !whole) {
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        *reinterpret_cast<float4*>(dv + c + e) = make_float4(__uint_as_float(x[e]), __uint_as_float(x[e + 1]),
                                                                             __uint_as_float(x[e + 2]), __uint_as_float(x[e + 3]));
                        *reinterpret_cast<float4*>(dk + c + e) =
                            make_float4(__uint_as_float(y[e]) * kmul, __uint_as_float(y[e + 1]) * kmul, __uint_as_float(y[e + 2]) * kmul,
                                        __uint_as_float(y[e + 3]) * kmul);
                    }
                }
            }
            tc::fence_before();
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// ---- merging the segments of the split rows -------------------------------------------------------------------------------
// grid (leftover rows, 128 rows of the resident tile), 128 threads = columns of the head.  Fixed summation order: deterministic.
__global__ void __launch_bounds__(DH) attn_combine_fwd_kernel(const AttnParams p) {
    const int row = p.n_full + (int)blockIdx.x, r = (int)blockIdx.y, c = (int)threadIdx.x;
    const int h = row / p.n_rtiles, q = (row % p.n_rtiles) * TQ + r;
    const size_t slot0 = (size_t)blockIdx.x * p.segs;
    // batches of 8 segments with all their loads in flight (a plain loop over the 74 segments is a chain of L2 latencies)
    const float* ml = p.part_ml + slot0 * 2 * TQ + r;
    const float* po = p.part + (slot0 * 2 * TQ + r) * DH + c;
    float M = -INFINITY;
    for (int s0 = 0; s0 < p.segs; s0 += 8) {
        float t[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) t[a] = (s0 + a < p.segs) ? __ldg(ml + (size_t)(s0 + a) * 2 * TQ) : -INFINITY;
#pragma unroll
        for (int a = 0; a < 8; ++a) M = fmaxf(M, t[a]);
    }
    float L = 0.0f, o = 0.0f;
    for (int s0 = 0; s0 < p.segs; s0 += 8) {
        float tm[8], tl[8], to[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const bool in = s0 + a < p.segs;
            tm[a] = in ? __ldg(ml + (size_t)(s0 + a) * 2 * TQ) : -INFINITY;
            tl[a] = in ? __ldg(ml + (size_t)(s0 + a) * 2 * TQ + TQ) : 0.0f;
            to[a] = in ? __ldg(po + (size_t)(s0 + a) * 2 * TQ * DH) : 0.0f;
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const float w = ex2(tm[a] - M);      // 0 for the padding of the last batch
            L += w * tl[a];
            o += w * to[a];
        }
    }
    if (q < p.S) p.ctx_out[(size_t)q * p.D + h * DH + c] = o / L;
    if (c == 0) p.lse2[(size_t)h * p.S_pad + q] = (q < p.S) ? M + log2f(L) : INFINITY;
}
// dst[q, col0 + h*128 + c] = mul * sum over segments of accumulator `arr`
__global__ void __launch_bounds__(DH) attn_combine_sum_kernel(const AttnParams p, int arr, int col0, float mul) {
    const int row = p.n_full + (int)blockIdx.x, r = (int)blockIdx.y, c = (int)threadIdx.x;
    const int h = row / p.n_rtiles, q = (row % p.n_rtiles) * TQ + r;
    if (q >= p.S) return;
    const size_t slot0 = (size_t)blockIdx.x * p.segs;
    float o = 0.0f;
    const float* po = p.part + ((slot0 * 2 + arr) * TQ + r) * DH + c;
    for (int s0 = 0; s0 < p.segs; s0 += 8) {      // 8 loads in flight
        float t[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) t[a] = (s0 + a < p.segs) ? __ldg(po + (size_t)(s0 + a) * 2 * TQ * DH) : 0.0f;
#pragma unroll
        for (int a = 0; a < 8; ++a) o += t[a];
    }
    p.dqkv[(size_t)q * (3 * p.D) + col0 + h * DH + c] = o * mul;
}

// ---- host side ------------------------------------------------------------------------------------------------------
constexpr int kMaxSlots = 148;      // one segment CTA per SM at most
struct AttnLayout {
    size_t planes_off, lse_off, delta_off, part_off, ml_off, keep_off, total;
    int S_pad;
};
AttnLayout attn_layout(int64_t S, int64_t heads) {
    AttnLayout l;
    l.S_pad = (int)((S + TQ - 1) / TQ * TQ);
    l.planes_off = 4096;
    const size_t planes = (size_t)4 * 2 * heads * l.S_pad * DH * sizeof(__nv_bfloat16);
    l.lse_off = l.planes_off + planes;
    l.delta_off = l.lse_off + (size_t)heads * l.S_pad * sizeof(float);
    l.part_off = (l.delta_off + (size_t)heads * l.S_pad * sizeof(float) + 255) & ~(size_t)255;
    l.ml_off = l.part_off + (size_t)kMaxSlots * 2 * TQ * DH * sizeof(float);
    l.keep_off = l.ml_off + (size_t)kMaxSlots * 2 * TQ * sizeof(float);
    l.total = l.keep_off + (size_t)heads * ((S + TK - 1) / TK) * l.S_pad * sizeof(unsigned long long);
    return l;
}
// Every CTA needs a whole SM (tiles + all of tensor memory), and a (resident tile, head) row costs the same everywhere, so
// R rows on n SMs take ceil(R / n) rounds: config 3 has R = 150 rows for 148 SMs -- two rounds, the second one for 2 rows.
// Rows beyond the last full round are therefore cut into floor(n / leftover) segments of consecutive streamed tiles each
// (150 rows: 148 whole rows + 2 x 74 segments of 2-3 tiles, 1.03 rounds), whose partial results the combine kernels merge.
void plan_split(AttnParams& p, int n_tiles) {
    int nsm = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148;
    if (const char* e = getenv("OPN_ATTN_SMS")) {      // tests: exercise whole rows + segments at small S
        const int v = atoi(e);
        if (v > 0) nsm = v;
    }
    if (nsm > kMaxSlots) nsm = kMaxSlots;
    const int R = p.n_rtiles * p.heads;
    p.n_full = R / nsm * nsm;
    const int left = R - p.n_full;
    p.segs = left > 0 ? nsm / left : 0;
    if (p.segs > n_tiles) p.segs = n_tiles;
    if (p.segs < 2) p.n_full = R, p.segs = 0;
    p.grid = p.n_full + (R - p.n_full) * p.segs;
}
constexpr size_t kFwdSmem = 1024 + 3 * 2 * STR_TILE;                        // 3 x (K, V) stages (Q' lives in tensor memory)
constexpr size_t kBwdSmem = 1024 + 2 * RES_TILE + 3 * STR_TILE;             // two resident tiles + 2 + 1 streamed tiles

int fill_params(AttnParams& p, const AttnLayout& l, int64_t S, int64_t D, int64_t heads, char* ws, float p_drop, uint64_t seed,
                uint64_t offset) {
    p.S = (int)S, p.S_pad = l.S_pad, p.D = (int)D, p.heads = (int)heads;
    p.scale = 1.0f / sqrtf((float)DH);
    p.lse2 = reinterpret_cast<float*>(ws + l.lse_off);
    p.delta = reinterpret_cast<float*>(ws + l.delta_off);
    p.status = reinterpret_cast<unsigned int*>(status_page_or(ws));
    double t = (double)p_drop * 4294967296.0;
    p.drop_threshold = p_drop > 0.0f ? (t >= 4294967295.0 ? 4294967295u : (unsigned int)t) : 0u;
    p.drop_scale = 1.0f / (1.0f - p_drop);
    p.seed = seed, p.offset = offset;
    p.planes = reinterpret_cast<const __nv_bfloat16*>(ws + l.planes_off);
    p.ctx = nullptr, p.ctx_out = nullptr, p.dctx = nullptr, p.dqkv = nullptr;
    p.part = reinterpret_cast<float*>(ws + l.part_off);
    p.part_ml = reinterpret_cast<float*>(ws + l.ml_off);
    p.keep = reinterpret_cast<unsigned long long*>(ws + l.keep_off);
    p.n_ktiles = (int)((S + TK - 1) / TK);
    p.n_rtiles = l.S_pad / TQ;
    plan_split(p, (int)((S + TK - 1) / TK));
    return OPN_OK;
}

template <typename Kernel>
int launch_attn(Kernel kernel, int threads, size_t smem, const CUtensorMap& m128, const CUtensorMap& m64, const AttnParams& p, cudaStream_t s) {
    OPN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<dim3((unsigned)p.grid), threads, smem, s>>>(m128, m64, p);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}
int combine_fwd(const AttnParams& p, cudaStream_t s) {
    const int left = p.n_rtiles * p.heads - p.n_full;
    if (left <= 0) return OPN_OK;
    attn_combine_fwd_kernel<<<dim3((unsigned)left, TQ), DH, 0, s>>>(p);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}
int combine_sum(const AttnParams& p, int arr, int col0, float mul, cudaStream_t s) {
    const int left = p.n_rtiles * p.heads - p.n_full;
    if (left <= 0) return OPN_OK;
    attn_combine_sum_kernel<<<dim3((unsigned)left, TQ), DH, 0, s>>>(p, arr, col0, mul);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

int prep(const float* src, long long ld, int col0, int t, float mul, __nv_bfloat16* planes, const AttnLayout& l, int64_t S, int64_t heads,
         cudaStream_t s) {
    const long long total = (long long)l.S_pad * heads * (DH / 8);
    long long grid = (total + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    attn_prep_kernel<<<(unsigned)grid, 256, 0, s>>>(src, ld, col0, t, mul, planes, (int)S, l.S_pad, (int)heads);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

}  // namespace
}  // namespace opn

using namespace opn;

extern "C" int64_t opn_attention_workspace_bytes(int64_t S, int64_t D, int64_t nhead) {
    if (S <= 0 || nhead <= 0 || D != nhead * DH) return 0;
    return (int64_t)attn_layout(S, nhead).total;
}

extern "C" int opn_attention_fwd(int64_t S, int64_t D, int64_t nhead, const float* qkv, float* ctx_out, void* workspace,
                                 int64_t workspace_bytes, float p_drop, uint64_t seed, uint64_t offset, void* stream) {
    OPN_CHECK_ARG(S > 0 && nhead > 0 && qkv && ctx_out && workspace, "attention_fwd: bad argument");
    if (D != nhead * DH) {
        set_error("attention_fwd: the fused kernel exists for head dimension %d (got D = %lld, %lld heads)", DH, (long long)D, (long long)nhead);
        return OPN_ERR_UNSUPPORTED;
    }
    OPN_CHECK_ARG(p_drop >= 0.0f && p_drop < 1.0f, "attention_fwd: p = %g outside [0, 1)", (double)p_drop);
    const AttnLayout l = attn_layout(S, nhead);
    OPN_CHECK_ARG(workspace_bytes >= (int64_t)l.total, "attention_fwd: workspace too small");
    cudaStream_t s = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    OPN_CUDA(cudaMemsetAsync(ws, 0, 4096, s));
    __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(ws + l.planes_off);
    int rc;
    const float qmul = (1.0f / sqrtf((float)DH)) * kLog2e;
    if ((rc = prep(qkv, 3 * D, 0, 0, qmul, planes, l, S, nhead, s)) != OPN_OK) return rc;
    if ((rc = prep(qkv, 3 * D, (int)D, 1, 1.0f, planes, l, S, nhead, s)) != OPN_OK) return rc;
    if ((rc = prep(qkv, 3 * D, (int)(2 * D), 2, 1.0f, planes, l, S, nhead, s)) != OPN_OK) return rc;
    CUtensorMap m128, m64;
    const long long rows = (long long)4 * 2 * nhead * l.S_pad;
    if ((rc = make_map_16bit(&m128, planes, rows, DH, TQ, true)) != OPN_OK) return rc;
    if ((rc = make_map_16bit(&m64, planes, rows, DH, TK, true)) != OPN_OK) return rc;
    AttnParams p;
    fill_params(p, l, S, D, nhead, ws, p_drop, seed, offset);
    p.ctx_out = ctx_out;
    const bool single = current_precision() == OPN_PRECISION_16BIT, drop = p_drop > 0.0f;
    const int th = AT + 32 * rng_warps(drop, true);
    if (single) rc = drop ? launch_attn(attn_fwd_kernel<1, true>, th, kFwdSmem, m128, m64, p, s) : launch_attn(attn_fwd_kernel<1, false>, th, kFwdSmem, m128, m64, p, s);
    else rc = drop ? launch_attn(attn_fwd_kernel<3, true>, th, kFwdSmem, m128, m64, p, s) : launch_attn(attn_fwd_kernel<3, false>, th, kFwdSmem, m128, m64, p, s);
    if (rc != OPN_OK) return rc;
    return combine_fwd(p, s);
}

extern "C" int opn_attention_bwd(int64_t S, int64_t D, int64_t nhead, const float* ctx, const float* dctx, float* dqkv, void* workspace,
                                 int64_t workspace_bytes, float p_drop, uint64_t seed, uint64_t offset, void* stream) {
    OPN_CHECK_ARG(S > 0 && nhead > 0 && ctx && dctx && dqkv && workspace, "attention_bwd: bad argument");
    if (D != nhead * DH) {
        set_error("attention_bwd: the fused kernel exists for head dimension %d", DH);
        return OPN_ERR_UNSUPPORTED;
    }
    const AttnLayout l = attn_layout(S, nhead);
    OPN_CHECK_ARG(workspace_bytes >= (int64_t)l.total, "attention_bwd: workspace too small");
    cudaStream_t s = as_stream(stream);
    char* ws = static_cast<char*>(workspace);       // still holds the Q', K, V planes and the row statistics of the forward call
    __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(ws + l.planes_off);
    int rc;
    if ((rc = prep(dctx, D, 0, 3, 1.0f, planes, l, S, nhead, s)) != OPN_OK) return rc;
    CUtensorMap m128, m64;
    const long long rows = (long long)4 * 2 * nhead * l.S_pad;
    if ((rc = make_map_16bit(&m128, planes, rows, DH, TQ, true)) != OPN_OK) return rc;
    if ((rc = make_map_16bit(&m64, planes, rows, DH, TK, true)) != OPN_OK) return rc;
    AttnParams p;
    fill_params(p, l, S, D, nhead, ws, p_drop, seed, offset);
    p.ctx = ctx, p.dctx = dctx, p.dqkv = dqkv;
    const bool single = current_precision() == OPN_PRECISION_16BIT, drop = p_drop > 0.0f;
    int th = AT + 32 * rng_warps_q(drop);
#define OPN_ATTN_BWD(K)                                                                                           \
    (single ? (drop ? launch_attn(K<1, true>, th, kBwdSmem, m128, m64, p, s) : launch_attn(K<1, false>, th, kBwdSmem, m128, m64, p, s)) \
            : (drop ? launch_attn(K<3, true>, th, kBwdSmem, m128, m64, p, s) : launch_attn(K<3, false>, th, kBwdSmem, m128, m64, p, s)))
    if ((rc = OPN_ATTN_BWD(attn_bwd_q_kernel)) != OPN_OK) return rc;      // writes delta, read by the key-tile kernel
    th = AT + 32 * rng_warps(drop, false);
    if ((rc = combine_sum(p, 0, 0, p.scale, s)) != OPN_OK) return rc;
    if ((rc = OPN_ATTN_BWD(attn_bwd_kv_kernel)) != OPN_OK) return rc;
    if ((rc = combine_sum(p, 0, (int)(2 * D), 1.0f, s)) != OPN_OK) return rc;      // dV
    return combine_sum(p, 1, (int)D, 1.0f / kLog2e, s);                            // dK (Q' carries scale * log2(e))
#undef OPN_ATTN_BWD
}
