// Time-parallel weight-gradient contractions of the LSTM layers on the 5th-generation tensor cores, sm_100a:
//
//   dW_ih[M, I] = sum_r        d_gates[r, :]^T x[r, :]          M = 4H gate rows, r = b*T + t over all B*T rows
//   dW_hh[M, H] = sum_{t(r)>0} d_gates[r, :]^T hs[r - 1, :]
//
// i.e. the weight gradients autograd forms for nn.LSTM (baselines/learned_models.py:39,46,76,113,146,192), several of them
// per launch ("jobs").  Both operands are contracted over their ROW index, so neither is K-major as it lies in memory; the
// general path (opn_sgemm -> opn_gemm_tc.cu) therefore ran a transposing bf16 hi/lo split pre-pass per operand (8 launches,
// 0.09 ms of the 2.5 ms headline step), zero fills, a correction product for the rows that straddle two videos and fp32
// atomics for split-K.  Here:
//   * the LARGE operand d_gates (78.6 + 39.3 MB of the headline step) is read once per use, as fp32, as it lies:
//     warps 0-11 (thread = gate column = TMEM lane, three groups alternating k-blocks) load 64 rows down their column
//     (coalesced across the warp), split to bf16 hi / lo pairs and write them with tcgen05.st: the A operand of the MMA
//     lives in TENSOR MEMORY (no shared-memory traffic, no transposition)
//   * the small operand (x / hs, re-read by every 128-column tile of d_gates) is split ONCE by wgrad_split_b_kernel into
//     row-major bf16 hi / lo planes -- the one-row shift of dW_hh and the exclusion of the first frame of every video
//     applied on the way -- and streamed by TMA (SWIZZLE_128B boxes of 64 rows x 64 columns) into a 3-stage ring, read by
//     the MMA through an MN-major descriptor.  (A first version converted this operand in the kernel too, from registers:
//     4,300 clocks per k-block, bound by the ~45 KB of loads an SM keeps in flight; profiles/r02_wgrad_*.log.)
//   * warp 13: hi.hi + lo.hi + hi.lo (one pass in the 1e-2 mode) per 16-row slice into a [128 x N<=256] fp32 accumulator
//   * CTA = (128 gate columns, <= 256 columns of the other operand, a range of 64-row k-blocks) of one job; the partial
//     sums of the k-ranges go to scratch and are added in a fixed order by wgrad_reduce_kernel (deterministic, no atomics,
//     no zero fill), which also transposes to [M][N].
#include <stdlib.h>

#include <cuda_bf16.h>

#include "opn_tc_common.cuh"

namespace opn {

int current_precision();   // opn_api.cu

namespace {

constexpr int WM = 128;                 // gate columns per CTA (M of the MMA, TMEM lanes)
constexpr int WK = 64;                  // rows per k-block
constexpr int WN = 256;                 // columns of the other operand per CTA at most (N of the MMA)
constexpr int NS = 3;                   // pipeline stages
constexpr int A_GROUPS = 3;             // groups of four A-converter warps (one per TMEM lane quadrant) alternate k-blocks
constexpr int A_WARPS = 4 * A_GROUPS;
constexpr int TMA_WARP = A_WARPS, MMA_WARP = A_WARPS + 1, PF_WARP = A_WARPS + 2;
constexpr int WT = 32 * (A_WARPS + 3);
#ifndef OPN_WGRAD_PF
#define OPN_WGRAD_PF 2      // 0: no L2 prefetch, 1: bulk prefetch per row (TMA unit), 2: prefetch.global.L2 per line (LSU)
#endif
constexpr int PF_DIST = 4;              // k-blocks the L2 prefetch of d_gates runs ahead of the pipeline
constexpr int BLK = WK * 128;           // one [64 k][64 n] bf16 block: 8 KB
constexpr int B_PLANE = (WN / 64) * BLK;
constexpr int B_STAGE = 2 * B_PLANE;    // hi and lo planes: 64 KB
constexpr int kMaxJobs = 8;
constexpr int kChainBlocks = 80;         // k-blocks per accumulator at most: bounds the accumulation chains (see n_acc in the kernel)
constexpr long long kWgradTimeout = 4000000000LL;
constexpr size_t kWgradSmem = 1024 + (size_t)NS * B_STAGE;

struct Job {
    const float* a;        // [rows][>= M] d(gates)
    const float* b;        // [rows][N]
    __nv_bfloat16* planes; // [2 (hi, lo)][rows_pad][n_pad]: b shifted / masked / zero padded, split
    float* part;           // [ksplit][n_pad][M] partial sums (column n, then gate row m)
    float* out;            // [M][N] (ldc)
    long long lda, ldb, ldc;
    int rows, rows_pad, T, M, N, n_pad;
    int shift;             // 1: row r of `a` meets row r - 1 of `b`; rows with r % T == 0 take no part
    int m_tiles, n_chunks, ksplit, cta0;
    int b_vec;             // rows of `b` are 16-byte aligned: float4 loads
    int trans_out;         // out is [N][M] (ldc): the partial sums already lie that way
};
struct WgradParams {
    Job jobs[kMaxJobs];
    int n_jobs;
    unsigned int* status;
};
struct WgradMaps {
    CUtensorMap m[kMaxJobs];
};

__device__ __forceinline__ bool await(uint64_t* bar, uint32_t parity, unsigned int* status, volatile int* abort_s) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0) {
            if (*abort_s) return false;
            if (clock64() - t0 > kWgradTimeout) {
                if (atomicCAS(status, 0u, 3u) == 0u) {
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

// (a, b) -> packed bf16 pair hi (a in the low half) and the pair of the residuals lo
__device__ __forceinline__ void split_pair_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
// MN-major SWIZZLE_128B operand in shared memory: rows of the tile are the K index (128 bytes = 64 consecutive MN elements
// each, 8-row groups 1024 bytes apart = SBO); the next 64 MN elements are `lbo` bytes further
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ---- the small operand: fp32 [rows][N] -> bf16 hi / lo planes [rows_pad][n_pad], shifted, masked, zero padded --------------
// grid (x: items, y: job); one thread = 8 consecutive columns of one row
__global__ void __launch_bounds__(256) wgrad_split_b_kernel(const __grid_constant__ WgradParams p, int planes_n) {
    const Job& jb = p.jobs[blockIdx.y];
    const int ipr = jb.n_pad >> 3;
    const long long total = (long long)jb.rows_pad * ipr;
    const size_t plane = (size_t)jb.rows_pad * jb.n_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / ipr), n = (int)(i - (long long)r * ipr) * 8;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < jb.rows && n < jb.N && !(jb.shift && (r % jb.T) == 0)) {
            const float* src = jb.b + (long long)(r - jb.shift) * jb.ldb + n;
            if (jb.b_vec && n + 8 <= jb.N) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(src)), w = __ldg(reinterpret_cast<const float4*>(src + 4));
                v[0] = u.x, v[1] = u.y, v[2] = u.z, v[3] = u.w, v[4] = w.x, v[5] = w.y, v[6] = w.z, v[7] = w.w;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (n + e < jb.N) v[e] = __ldg(src + e);
            }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_pair_bf16(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
        const size_t o = (size_t)r * jb.n_pad + n;
        *reinterpret_cast<uint4*>(jb.planes + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (planes_n == 2) *reinterpret_cast<uint4*>(jb.planes + plane + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// per-role cycle counters of CTA 0 (development build: python -m objectpermanence_b200.build --phases; tools/wgrad_phases.py)
#ifdef OPN_LSTM_PHASES
#define WPH_DECL long long wph[4] = {0, 0, 0, 0}; long long wph_last = clock64(); const long long wph_c0 = wph_last; unsigned long long wph_t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(wph_t0));
#define WPH_SPAN(status) do { unsigned long long t1__; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1__)); if (blockIdx.x < 224) { \
        unsigned long long* o__ = reinterpret_cast<unsigned long long*>(status) + 64 + 2 * blockIdx.x; o__[0] = wph_t0; o__[1] = t1__; \
        if (blockIdx.x == 0) { reinterpret_cast<unsigned long long*>(status)[60] = (unsigned long long)(clock64() - wph_c0); } } } while (0)
#define WPH(i) do { const long long n__ = clock64(); wph[i] += n__ - wph_last; wph_last = n__; } while (0)
#define WPH_STORE(status, role) do { if (blockIdx.x == 0) { unsigned long long* o__ = reinterpret_cast<unsigned long long*>(status) + 32 + 4 * (role); \
        for (int i__ = 0; i__ < 4; ++i__) o__[i__] = (unsigned long long)wph[i__]; } } while (0)
#else
#define WPH_DECL
#define WPH(i)
#define WPH_STORE(status, role)
#define WPH_SPAN(status)
#endif

template <int PASSES>
__global__ void __launch_bounds__(WT, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p, const __grid_constant__ WgradMaps maps) {
    constexpr int PL = PASSES == 3 ? 2 : 1;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[NS], empty[NS], done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int abort_flag;
    __shared__ int mma_progress;      // k-blocks the MMA thread has seen full (paces the prefetch warp; a phase parity could be lapped)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b_s = (smem_u32(smem_raw) + 1023u) & ~1023u;

    int ji = 0;
    while (ji + 1 < p.n_jobs && (int)blockIdx.x >= p.jobs[ji + 1].cta0) ++ji;
    const Job& jb = p.jobs[ji];
    const int local = (int)blockIdx.x - jb.cta0;
    const int mt = local % jb.m_tiles, nc = (local / jb.m_tiles) % jb.n_chunks, ks = local / (jb.m_tiles * jb.n_chunks);
    const int n0 = nc * WN;
    const int nw = min(WN, jb.n_pad - n0);      // multiple of 64
    const int nkb_all = jb.rows_pad / WK;
    const int kb0 = (int)((long long)ks * nkb_all / jb.ksplit), kb1 = (int)((long long)(ks + 1) * nkb_all / jb.ksplit);
    const int nkb = kb1 - kb0;
    // The tensor core truncates on every accumulation, so the error of a chain grows with its length (measured: 2.5e-4 of
    // the largest entry after 5,000 accumulations).  Narrow tiles leave accumulator columns free: consecutive k-blocks
    // rotate over n_acc accumulators, added up in the epilogue (the host also bounds the k-blocks per CTA, kChainBlocks).
    const int n_acc = nw <= 64 ? 4 : (nw <= 128 ? 2 : 1);

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 4 + 1);      // the four warps of an A group + the TMA thread (with the bytes of the B tile)
            mbar_init(&empty[i], 1);
        }
        mbar_init(&done, 1);
        mbar_fence_init();
        abort_flag = 0;
        mma_progress = 0;
    }
    if (warp == MMA_WARP) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    WPH_DECL
    const uint32_t tmem = tmem_base_s;      // accumulator: columns 0 .. nw-1; A planes of stage s: 256 + 64 s (hi 32, lo 32)
    volatile int* abort_s = &abort_flag;

    if (warp < A_WARPS) {
        // ---- A converters: thread = gate column m, down the 64 rows of the k-block; group g takes k-blocks it % A_GROUPS == g
        const int grp = warp >> 2, qd = warp & 3;
        const int m = mt * WM + qd * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(qd * 32) << 16);
        bool ok = true;
        for (int it = grp; it < nkb; it += A_GROUPS) {
            const int s = it % NS;
            const int r0 = (kb0 + it) * WK;
            float x[WK];
            const float* src = jb.a + (long long)r0 * jb.lda + m;
            if (r0 + WK <= jb.rows) {
#pragma unroll
                for (int k = 0; k < WK; ++k) x[k] = __ldg(src + (long long)k * jb.lda);
            } else {
#pragma unroll
                for (int k = 0; k < WK; ++k) x[k] = (r0 + k < jb.rows) ? __ldg(src + (long long)k * jb.lda) : 0.0f;
            }
            WPH(0);
            if (it >= NS && !await(&empty[s], ((uint32_t)(it / NS) - 1u) & 1u, p.status, abort_s)) { ok = false; break; }
            WPH(1);
            tc::fence_after();
            const uint32_t taddr = lane_addr + 256 + s * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {      // 32 rows -> 16 columns per plane at a time
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) split_pair_bf16(x[32 * half + 2 * i], x[32 * half + 2 * i + 1], hi[i], lo[i]);
                tc::tmem_st16(taddr + 16 * half, hi);
                if (PL == 2) tc::tmem_st16(taddr + 32 + 16 * half, lo);
            }
            tc::tmem_st_wait();
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full[s]);
            WPH(2);
        }
        // ---- epilogue: accumulator -> partial sums [n][m] (a warp stores 128 contiguous bytes per column); the groups
        // take alternate 32-column slices
        float* dst = jb.part + ((size_t)ks * jb.n_pad + n0) * jb.M + m;
        if (nkb == 0) {
            for (int c = grp; c < nw; c += A_GROUPS) dst[(size_t)c * jb.M] = 0.0f;
        } else if (ok && await(&done, 0, p.status, abort_s)) {
            tc::fence_after();
            const int used = min(n_acc, nkb);
            for (int c = 32 * grp; c < nw; c += 32 * A_GROUPS) {
                uint32_t v[32];
                tc::tmem_ld32(lane_addr + c, v);
                tc::tmem_ld_wait();
                for (int a = 1; a < used; ++a) {
                    uint32_t w[32];
                    tc::tmem_ld32(lane_addr + a * nw + c, w);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) dst[(size_t)(c + i) * jb.M] = __uint_as_float(v[i]);
            }
            tc::fence_before();
        }
        WPH(3);
        if (tid == 0) WPH_STORE(p.status, 0);
    } else if (warp == TMA_WARP) {
        // ---- the [64 x nw] tile of the split small operand: nw / 64 boxes per plane ------------------------------------------
        if (lane == 0) {
            const CUtensorMap* map = &maps.m[ji];
            const int nb = nw >> 6;
            for (int it = 0; it < nkb; ++it) {
                const int s = it % NS;
                if (it >= NS && !await(&empty[s], ((uint32_t)(it / NS) - 1u) & 1u, p.status, abort_s)) break;
                const int r0 = (kb0 + it) * WK;
                const uint32_t stage = b_s + s * B_STAGE;
                mbar_arrive_expect_tx(&full[s], (uint32_t)(PL * nb * BLK));
#pragma unroll
                for (int pl = 0; pl < PL; ++pl)
                    for (int cb = 0; cb < nb; ++cb)
                        tc::tma_load_2d(stage + pl * B_PLANE + cb * BLK, map, n0 + 64 * cb, pl * jb.rows_pad + r0, &full[s]);
            }
        }
    } else if (warp == PF_WARP) {
        // ---- L2 prefetch of the d_gates tile PF_DIST k-blocks ahead (one bulk prefetch per row), paced by the pipeline: the
        // tile is read by the CTAs of n_chunks column ranges only, so without it every k-block pays HBM latency
        volatile int* prog = &mma_progress;
        for (int it = -PF_DIST; OPN_WGRAD_PF != 0 && it < nkb - PF_DIST; ++it) {
            if (it >= 0) {
                const long long t0 = clock64();
                while (*prog <= it && !*abort_s && clock64() - t0 < kWgradTimeout) __nanosleep(64);
                if (*prog <= it) break;
            }
            const int r0 = (kb0 + it + PF_DIST) * WK;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = r0 + lane + 32 * h;
                if (r < jb.rows) {
                    const float* row = jb.a + (long long)r * jb.lda + mt * WM;
#if OPN_WGRAD_PF == 1
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"(WM * 4) : "memory");
#else
#pragma unroll
                    for (int l = 0; l < WM * 4 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 32 * l) : "memory");
#endif
                }
            }
        }
    } else if (lane == 0) {
        // ---- MMA issue ------------------------------------------------------------------------------------------------
        const uint32_t idesc = tc::idesc_f16(WM, nw, true) | (1u << 16);      // B operand MN-major
        bool ok = true;
        for (int it = 0; it < nkb && ok; ++it) {
            const int s = it % NS;
            if (!await(&full[s], (uint32_t)(it / NS) & 1u, p.status, abort_s)) { ok = false; break; }
            WPH(0);
            *reinterpret_cast<volatile int*>(&mma_progress) = it + 1;
            tc::fence_after();
            const uint32_t stage = b_s + s * B_STAGE, a_t = tmem + 256 + s * 64;
            const uint32_t acc = tmem + (uint32_t)((it % n_acc) * nw);
#pragma unroll
            for (int kk = 0; kk < WK / 16; ++kk) {
                const uint64_t bh = desc_mn_sw128(stage + kk * 2048, BLK);
                tc::umma_f16_ts(acc, a_t + kk * 8, bh, idesc, (it >= n_acc || kk > 0) ? 1u : 0u);
                if (PASSES == 3) {
                    const uint64_t bl = desc_mn_sw128(stage + B_PLANE + kk * 2048, BLK);
                    tc::umma_f16_ts(acc, a_t + 32 + kk * 8, bh, idesc, 1u);
                    tc::umma_f16_ts(acc, a_t + kk * 8, bl, idesc, 1u);
                }
            }
            tc::umma_commit(&empty[s]);
            WPH(1);
        }
        if (ok && nkb > 0) tc::umma_commit(&done);
        WPH_STORE(p.status, 2);
    }
    tc::fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc(tmem, 512);
    if (tid == 0) WPH_SPAN(p.status);
}

// out[m][n] = sum over the k-ranges of part[ks][n][m]: 32 x 32 tiles transposed through shared memory, fixed order
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgradParams p) {
    const Job& jb = p.jobs[blockIdx.z];
    const int m0 = (int)blockIdx.x * 32, n0 = (int)blockIdx.y * 32;
    if (m0 >= jb.M || n0 >= jb.N) return;
    __shared__ float t_s[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // 16 independent loads in flight per thread (4 columns x 4 k-ranges): as a plain accumulation loop the kernel was a chain
    // of L2 latencies (48 us for 19 MB of partial sums)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* base = jb.part + (size_t)(n0 + ty) * jb.M + m0 + tx;      // column n0 + ty + 8 j < n_pad (a multiple of 64)
    const size_t ks_stride = (size_t)jb.n_pad * jb.M, j_stride = (size_t)8 * jb.M;
    for (int ks = 0; ks < jb.ksplit; ks += 4) {
        float t[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) t[a][j] = (ks + a < jb.ksplit) ? __ldg(base + (size_t)(ks + a) * ks_stride + j * j_stride) : 0.0f;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += t[a][j];
    }
    if (jb.trans_out) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + ty + 8 * j;
            if (n < jb.N) jb.out[(long long)n * jb.ldc + m0 + tx] = acc[j];
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) t_s[ty + 8 * j][tx] = acc[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int m = m0 + ty + 8 * j, n = n0 + tx;
        if (n < jb.N) jb.out[(long long)m * jb.ldc + n] = t_s[tx][ty + 8 * j];
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
struct Plan {
    WgradParams prm;
    size_t total_bytes;
    int grid;
};

// fewest k-ranges that keep every accumulation chain within kChainBlocks k-blocks (narrow tiles rotate over 2 or 4 accumulators)
int chain_floor(const Job& jb) {
    const int nw = jb.n_pad < WN ? jb.n_pad : WN, n_acc = nw <= 64 ? 4 : (nw <= 128 ? 2 : 1);
    const int nkb = jb.rows_pad / WK, cap = kChainBlocks * n_acc;
    return (nkb + cap - 1) / cap;
}

int make_plan(int n_jobs, const opn_wgrad_job* jobs, Plan& pl, char* ws) {
    OPN_CHECK_ARG(n_jobs >= 1 && n_jobs <= kMaxJobs && jobs, "wgrad: 1 .. %d jobs", kMaxJobs);
    int nsm = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148;
    const bool single = current_precision() == OPN_PRECISION_16BIT;
    double unit_cost[kMaxJobs], total = 0.0;
    int order[kMaxJobs];
    Job tmp[kMaxJobs];
    for (int j = 0; j < n_jobs; ++j) {
        const opn_wgrad_job& u = jobs[j];
        OPN_CHECK_ARG(u.a && u.b && u.out && u.rows > 0 && u.T > 0 && u.M > 0 && u.N > 0, "wgrad: job %d: bad argument", j);
        OPN_CHECK_ARG(u.M % WM == 0, "wgrad: job %d: M = %lld is not a multiple of %d", j, (long long)u.M, WM);
        OPN_CHECK_ARG(u.lda >= u.M && u.ldb >= u.N && u.ldc >= (u.trans_out ? u.M : u.N), "wgrad: job %d: row stride shorter than the row", j);
        OPN_CHECK_ARG(u.rows < (1LL << 31) && u.M < (1LL << 24) && u.N < (1LL << 24), "wgrad: job %d: size out of range", j);
        Job& jb = tmp[j];
        jb.a = u.a, jb.b = u.b, jb.out = u.out, jb.part = nullptr;
        jb.lda = u.lda, jb.ldb = u.ldb, jb.ldc = u.ldc;
        jb.rows = (int)u.rows, jb.T = (int)u.T, jb.M = (int)u.M, jb.N = (int)u.N;
        jb.n_pad = (jb.N + 63) / 64 * 64;
        jb.rows_pad = (jb.rows + WK - 1) / WK * WK;
        jb.planes = nullptr;
        jb.shift = u.shift ? 1 : 0;
        jb.trans_out = u.trans_out ? 1 : 0;
        jb.m_tiles = jb.M / WM;
        jb.n_chunks = (jb.n_pad + WN - 1) / WN;
        jb.b_vec = (u.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(u.b) & 15) == 0) ? 1 : 0;
        // cost of one (m-tile, n-chunk) unit per k-block: MMA issue (128 N / 256 clocks per instruction) or the conversion
        const int nw = jb.n_pad < WN ? jb.n_pad : WN;
        const double mma = 4.0 * (single ? 1 : 3) * nw * 0.5, conv = 900.0;
        const int nkb = (jb.rows + WK - 1) / WK;
        unit_cost[j] = (mma > conv ? mma : conv) * nkb;
        total += unit_cost[j] * jb.m_tiles * jb.n_chunks;
        order[j] = j;
    }
    // k-ranges per job: the smallest per-CTA budget for which all CTAs fit into ONE round of the SMs (CTAs of equal length
    // in a second, partial round would double the time); huge problems, whose chains the cap bounds, take several rounds
    const double overhead = 6000.0;      // prologue + epilogue of a CTA, clocks
    double budget = total / nsm;
    for (int iter = 0; iter < 200; ++iter, budget *= 1.03) {
        int ctas = 0;
        for (int j = 0; j < n_jobs; ++j) {
            Job& jb = tmp[j];
            const int nkb = jb.rows_pad / WK;
            int ksplit = (int)(unit_cost[j] / budget) + 1;
            if (ksplit < chain_floor(jb)) ksplit = chain_floor(jb);
            if (ksplit > nkb) ksplit = nkb;
            jb.ksplit = ksplit;
            ctas += jb.m_tiles * jb.n_chunks * ksplit;
        }
        if (ctas <= nsm) break;
        bool floored = true;      // nothing left to merge: every job sits at its chain-length floor
        for (int j = 0; j < n_jobs; ++j) floored = floored && tmp[j].ksplit <= chain_floor(tmp[j]);
        if (floored) break;
    }
    for (int j = 0; j < n_jobs; ++j) unit_cost[j] = unit_cost[j] / tmp[j].ksplit + overhead;
    // the longest CTAs first (the hardware hands out CTAs in index order)
    for (int i = 0; i < n_jobs; ++i)
        for (int j = i + 1; j < n_jobs; ++j)
            if (unit_cost[order[j]] > unit_cost[order[i]]) {
                const int t = order[i];
                order[i] = order[j], order[j] = t;
            }
    size_t off = 4096;
    int cta = 0;
    pl.prm.n_jobs = n_jobs;
    for (int i = 0; i < n_jobs; ++i) {
        Job& jb = pl.prm.jobs[i];
        jb = tmp[order[i]];
        jb.cta0 = cta;
        cta += jb.m_tiles * jb.n_chunks * jb.ksplit;
        jb.part = ws ? reinterpret_cast<float*>(ws + off) : nullptr;
        off += ((size_t)jb.ksplit * jb.n_pad * jb.M * sizeof(float) + 255) & ~(size_t)255;
        jb.planes = ws ? reinterpret_cast<__nv_bfloat16*>(ws + off) : nullptr;
        off += ((size_t)2 * jb.rows_pad * jb.n_pad * sizeof(__nv_bfloat16) + 255) & ~(size_t)255;
    }
    pl.grid = cta;
    pl.total_bytes = off;
    pl.prm.status = nullptr;
    return OPN_OK;
}

}  // namespace
}  // namespace opn

using namespace opn;

extern "C" int64_t opn_wgrad_workspace_bytes(int32_t n_jobs, const opn_wgrad_job* jobs) {
    Plan pl;
    if (make_plan(n_jobs, jobs, pl, nullptr) != OPN_OK) return 0;
    return (int64_t)pl.total_bytes;
}

extern "C" int opn_wgrad(int32_t n_jobs, const opn_wgrad_job* jobs, void* workspace, int64_t workspace_bytes, void* stream) {
    OPN_CHECK_ARG(workspace, "wgrad: no workspace");
    Plan pl;
    char* ws = static_cast<char*>(workspace);
    int rc = make_plan(n_jobs, jobs, pl, ws);
    if (rc != OPN_OK) return rc;
    OPN_CHECK_ARG(workspace_bytes >= (int64_t)pl.total_bytes, "wgrad: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                  (long long)pl.total_bytes);
    cudaStream_t s = as_stream(stream);
    OPN_CUDA(cudaMemsetAsync(ws, 0, 4096, s));
    pl.prm.status = reinterpret_cast<unsigned int*>(status_page_or(ws));
    const bool single = current_precision() == OPN_PRECISION_16BIT;
    WgradMaps maps;
    long long items = 0;
    for (int j = 0; j < pl.prm.n_jobs; ++j) {
        const Job& jb = pl.prm.jobs[j];
        if ((rc = make_map_16bit(&maps.m[j], jb.planes, (long long)2 * jb.rows_pad, jb.n_pad, WK, true)) != OPN_OK) return rc;
        const long long it = (long long)jb.rows_pad * (jb.n_pad / 8);
        if (it > items) items = it;
    }
    for (int j = pl.prm.n_jobs; j < kMaxJobs; ++j) maps.m[j] = maps.m[0];
    long long sgrid = (items + 255) / 256;
    if (sgrid > 148 * 4) sgrid = 148 * 4;
    wgrad_split_b_kernel<<<dim3((unsigned)sgrid, (unsigned)pl.prm.n_jobs), 256, 0, s>>>(pl.prm, single ? 1 : 2);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    if (single) {
        OPN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmem));
        wgrad_tc_kernel<1><<<pl.grid, WT, kWgradSmem, s>>>(pl.prm, maps);
    } else {
        OPN_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmem));
        wgrad_tc_kernel<3><<<pl.grid, WT, kWgradSmem, s>>>(pl.prm, maps);
    }
    OPN_CUDA(cudaGetLastError());
    count_launch();
    int mx = 0, nx = 0;
    for (int j = 0; j < pl.prm.n_jobs; ++j) {
        if (pl.prm.jobs[j].M > mx) mx = pl.prm.jobs[j].M;
        if (pl.prm.jobs[j].N > nx) nx = pl.prm.jobs[j].N;
    }
    wgrad_reduce_kernel<<<dim3((unsigned)(mx / 32), (unsigned)((nx + 31) / 32), (unsigned)pl.prm.n_jobs), 256, 0, s>>>(pl.prm);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}
