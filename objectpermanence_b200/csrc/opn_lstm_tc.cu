// Batch-wide persistent LSTM recurrence on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// The mma.sync kernels (opn_lstm_mma.cu, the fused OPNet kernels) give a CTA 8 videos: per-GPU batches above 32 run as
// sequential waves and throughput is flat in B.  Here a batch group is 128 videos = the M dimension of one tcgen05.mma
// and the 128 lanes of TMEM; the weights are the (stationary) B operand:
//
//   forward,  CTA (slice s of H/16, group g):  D[128 videos x 64 gate rows] = h_{t-1}[128 x H] . W_hh[rows of s]^T
//     * the CTA's 64 x H slice of W_hh (rows gate*16 + unit) lives in shared memory for the whole sequence as fp16 hi / lo
//       planes in the K-major SWIZZLE_128B layout of the UMMA descriptors (128 KB at H = 512), pre-scaled by a power of two;
//     * h_{t-1} of the group is exchanged through an L2-resident ring as fp16 hi / lo planes [video][H]: every CTA
//       publishes its 16 units (32 bytes per video and plane), a release-increment of a per-(group, k-block) counter hands
//       them over, and warp 0 streams the 128 x 64 k-block tiles back in with TMA as soon as their four producers are in;
//     * warp 1 issues the MMAs (hi.hi + lo.hi + hi.lo: fp32-grade; or one pass in the 1e-2 mode) into a TMEM accumulator;
//     * warps 2-5 (thread = video = TMEM lane) read the 64 pre-activations back with tcgen05.ld, add the x-projection,
//       apply the gates, keep c in registers and write h / gates / cells.
//   backward, same decomposition: the epilogue threads form d gates[128 x 64] (cell backward, per-video power-of-two scale),
//     write it as the fp16 hi / lo A tile (one 128-byte swizzled row per video), the MMA produces the CTA's partial
//     dh_{t-1}[128 x H] = d gates . W_hh[rows of s] into TMEM (H columns), and the partials are reduce-scattered through an
//     fp32 ring: consumer c sums the H/16 producers' [128 x 16] blocks for its own units.
// Inter-CTA ordering is release / acquire on counters at GPU scope (plus the generic -> async proxy fence where TMA reads):
// no flag-in-data tricks.  Every wait has a clock64 time-out that reports through the status word.
//
// Replaces nn.LSTM at baselines/learned_models.py:29,32,66,100,131,170 (and its autograd backward) for large per-GPU batches
// (lstm_tc_wanted below); selected inside opn_lstm_fwd / opn_lstm_bwd.  The gates / cells stash it writes has its own
// (video-minor, coalesced) layout and is only ever read by the backward kernel of this file.
#include <stdlib.h>

#include "opn_lstm_common.cuh"
#include "opn_mma_common.cuh"
#include "opn_tc_common.cuh"

namespace opn {

int current_precision();   // opn_api.cu: 0 = fp32-grade (split operands, three passes), 1 = one 16-bit pass (1e-2 mode)

namespace {

constexpr int MV = 128;          // videos per batch group: M of the MMA, lanes of TMEM
constexpr int UC = 16;           // hidden units per CTA
constexpr int NR = 4 * UC;       // gate rows per CTA
constexpr int KBE = 64;          // K elements per k-block: one 128-byte swizzle row of fp16
constexpr int TCT = 320;         // warp 0: TMA / counters, warp 1: MMA issue, warps 2-9: epilogue, thread = (video, 8 of the 16 units)
constexpr int NEW = 8;           // epilogue warps: two per TMEM lane quadrant (warp % 4), one per half of the CTA's units
constexpr int UH = UC / 2;       // units per epilogue thread
constexpr int NSTAGE = 2;
constexpr int A_TILE = MV * 128; // 16 KB: 128 videos x 64 fp16

struct TcFwdParams {
    const float* xproj;      // [B,T,4H]
    const float* w_hh;       // [4H,H]
    float *hs, *gates, *cells;
    __half* ring;            // [2 slots][planes][groups_total][128][H]
    unsigned int* counters;  // [groups_total][H/64]
    unsigned int* status;
    int B, T, groups_total, group_offset;
};

struct TcBwdParams {
    const float *w_hh, *gates, *cells, *dh_out;
    float* dgates;
    float* ring;             // [2 slots][groups_total][H/16 consumers][H/16 producers][128][16]
    unsigned int* counters;  // [groups_total]
    unsigned int* status;
    int B, T, groups_total, group_offset;
};

// per-role cycle counters of CTA 0 (development build: python -m objectpermanence_b200.build --phases; tools/lstm_tc_phases.py)
#ifdef OPN_LSTM_PHASES
#define TPH_DECL long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tph_last = clock64();
#define TPH(i) do { const long long n__ = clock64(); tph[i] += n__ - tph_last; tph_last = n__; } while (0)
#define TPH_STORE(status, role) do { if (blockIdx.x == 0) { unsigned long long* o__ = reinterpret_cast<unsigned long long*>(status) + 32 + 8 * (role); \
        for (int i__ = 0; i__ < 8; ++i__) o__[i__] = (unsigned long long)tph[i__]; } } while (0)
// wall-clock stamps (globaltimer, ns) of step 100 of every CTA: [cta][0 acc_full seen, 1 published, 2 counters satisfied, 3 first tile landed]
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TSTAMP(ctrs, step, which) do { if ((step) == 100) reinterpret_cast<unsigned long long*>(ctrs)[64 + 4 * blockIdx.x + (which)] = gtime(); } while (0)
#else
#define TPH_DECL
#define TPH(i)
#define TPH_STORE(status, role)
#define TSTAMP(ctrs, step, which)
#endif

struct Waiter {
    volatile int* abort_s;
    unsigned int* status;
    int step;
    __device__ __forceinline__ bool expired(long long t0) {
        if (*abort_s) return true;
        if (ld_relaxed(status) != 0) {
            *abort_s = 1;
            return true;
        }
        if (clock64() - t0 > kTimeoutCycles) {
            if (atomicCAS(status, 0u, kStatusPollTimeout) == 0u) {
                status[1] = (unsigned int)step;
                status[2] = blockIdx.x;
                status[3] = threadIdx.x;
            }
            *abort_s = 1;
            return true;
        }
        return false;
    }
    __device__ __forceinline__ bool barrier(uint64_t* bar, uint32_t parity) {
        if (mbar_try_wait(bar, parity)) return true;
        const long long t0 = clock64();
        unsigned spins = 0;
        while (!mbar_try_wait(bar, parity))
            if ((++spins & 63u) == 0 && expired(t0)) return false;
        return true;
    }
    // N consecutive counters (N = 4 or 8), all >= target; relaxed polling, acquire fence at the end
    template <int N>
    __device__ __forceinline__ bool counters(const unsigned int* ctr, unsigned int target) {
        const long long t0 = clock64();
        unsigned spins = 0;
        for (;;) {
            unsigned int lo = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < N; i += 4) {
                uint4 v;
                asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ctr + i) : "memory");
                lo = min(lo, min(min(v.x, v.y), min(v.z, v.w)));
            }
            if (lo >= target) break;
            if ((++spins & 63u) == 0 && expired(t0)) return false;
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        return true;
    }
    __device__ __forceinline__ bool counter(const unsigned int* ctr, unsigned int target) {
        if (tc::ld_acquire(ctr) >= target) return true;
        const long long t0 = clock64();
        unsigned spins = 0;
        while (tc::ld_acquire(ctr) < target)
            if ((++spins & 63u) == 0 && expired(t0)) return false;
        return true;
    }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_cg4(const float* p) {   // L2 only: the line was written by another SM a moment ago
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float sigmoid_sfu(float x) { return fmaf(0.5f, tanh_sfu(0.5f * x), 0.5f); }

// local gate row n of the CTA (the N / K index of its weight operand): n = half*32 + gate*8 + uu, so that the 32 values an
// epilogue thread (video, half) needs are 32 consecutive TMEM columns / one half of the d gates row
__device__ __forceinline__ int w_row(int n, int u0, int H) { return ((n >> 3) & 3) * H + u0 + (n >> 5) * UH + (n & 7); }

// The CTA's 64 gate rows of W_hh (local order n, see w_row), scaled by a power of two, as fp16 hi (/ lo)
// planes.  FWD: B operand [n][k], one 64 x 128-byte SWIZZLE_128B tile per 64-wide k-block.  BWD (TRANSPOSED): B operand
// [k][n]: H rows of 128 bytes (the 64 own rows along K).
template <int H, int PL, bool TRANSPOSED>
__device__ __forceinline__ float stage_weights(const float* __restrict__ w_hh, int u0, unsigned char* w_s, float* red_s) {
    const int tid = threadIdx.x;
    float m = 0.0f;
    for (int idx = tid; idx < NR * (H / 4); idx += TCT) {
        const int n = idx / (H / 4), k4 = idx % (H / 4);
        const float4 v = ldg4(w_hh + (size_t)w_row(n, u0, H) * H + 4 * k4);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    float wscale, winv;
    weight_scale<TCT / 32>(m, red_s, wscale, winv);
    constexpr int PLANE = TRANSPOSED ? H * 128 : (H / KBE) * NR * 128;
    if (!TRANSPOSED) {
        for (int idx = tid; idx < NR * (H / 8); idx += TCT) {
            const int n = idx / (H / 8), c8 = idx % (H / 8);
            const float* src = w_hh + (size_t)w_row(n, u0, H) * H + 8 * c8;
            const float4 a = ldg4(src), b = ldg4(src + 4);
            uint4 hi, lo;
            tc::split_pair(a.x * wscale, a.y * wscale, hi.x, lo.x);
            tc::split_pair(a.z * wscale, a.w * wscale, hi.y, lo.y);
            tc::split_pair(b.x * wscale, b.y * wscale, hi.z, lo.z);
            tc::split_pair(b.z * wscale, b.w * wscale, hi.w, lo.w);
            const uint32_t off = (uint32_t)(c8 / 8) * (NR * 128) + tc::sw128_offset(n, c8 % 8);
            *reinterpret_cast<uint4*>(w_s + off) = hi;
            if (PL == 2) *reinterpret_cast<uint4*>(w_s + PLANE + off) = lo;
        }
    } else {
        // element (k, n): 8 consecutive n of one k per thread (strided global reads, once per launch)
        for (int idx = tid; idx < H * (NR / 8); idx += TCT) {
            const int k = idx % H, c = idx / H;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int n = 8 * c + e;
                v[e] = __ldg(w_hh + (size_t)w_row(n, u0, H) * H + k) * wscale;
            }
            uint4 hi, lo;
            tc::split_pair(v[0], v[1], hi.x, lo.x);
            tc::split_pair(v[2], v[3], hi.y, lo.y);
            tc::split_pair(v[4], v[5], hi.z, lo.z);
            tc::split_pair(v[6], v[7], hi.w, lo.w);
            const uint32_t off = tc::sw128_offset(k, c);
            *reinterpret_cast<uint4*>(w_s + off) = hi;
            if (PL == 2) *reinterpret_cast<uint4*>(w_s + PLANE + off) = lo;
        }
    }
    return winv;
}

// ======================================================= forward =======================================================
template <int H, int PASSES>
__global__ void __launch_bounds__(TCT, 1) lstm_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_ring, const TcFwdParams p) {
    constexpr int NKB = H / KBE, PL = PASSES == 3 ? 2 : 1, NSL = H / UC;
    constexpr int W_TILE = NR * 128, W_PLANE = NKB * W_TILE;
    constexpr uint32_t IDESC = tc::idesc_f16(MV, NR);
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], acc_full, acc_empty, go_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int abort_flag;
    __shared__ float red_s[TCT / 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    unsigned char* w_s = sm;                               // [PL][NKB][64 x 128 B]
    const uint32_t a_s = base + PL * W_PLANE;              // [NSTAGE][PL][128 x 128 B]
    const int slice = blockIdx.x % NSL, group = p.group_offset + blockIdx.x / NSL;
    const int u0 = slice * UC, T = p.T;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_full, 1);
        mbar_init(&acc_empty, NEW);
        mbar_init(&go_bar, 1);
        mbar_fence_init();
        abort_flag = 0;
    }
    // accumulators: [0] hi.hi over the first half of K, [1] hi.hi over the second half, [2] the two cross products.  The
    // tensor core truncates on every accumulation: one chain of 96 MMAs was 5x less accurate than the mma.sync kernels
    // on the x8-weights stress case (4.8e-4 against 1e-4), chains of 16 are not.
    constexpr int NACC = PASSES == 3 ? 3 : 2, TMEM_COLS = 256;
    if (warp == 1) tc::tmem_alloc(&tmem_base_s, TMEM_COLS);
    const float winv = stage_weights<H, PL, false>(p.w_hh, u0, w_s, red_s);
    fence_proxy_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    Waiter wait = {&abort_flag, p.status, 0};
    const size_t ring_rows_per_plane = (size_t)p.groups_total * MV;

    if (warp == 0) {
        // ---- exchange watcher + TMA producer: tile kb of h_{t-1} as soon as its four producers have published it ----
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            TPH_DECL
            for (int t = 1; t < T && ok; ++t) {
                wait.step = t;
                const int slot = (t - 1) & 1;
                TPH(0);
                // every producer of the group has published h_{t-1}: all NKB counters at 32 t (4 slices x 8 epilogue warps
                // per k-block and step).  The counters are polled together with relaxed loads (an acquire load per k-block
                // serialised 8 L2 round trips per step); one acquire fence + proxy fence then covers all of them.
                if (!wait.counters<NKB>(p.counters + (size_t)group * NKB, (unsigned)(4 * NEW) * (unsigned)t)) break;
                TSTAMP(p.counters, t, 2);
                TPH(1);
                tc::acquire_for_tma();
                TPH(2);
                for (int kb = 0; kb < NKB; ++kb) {
                    if (!wait.barrier(&empty_bar[stage], phase ^ 1u)) { ok = false; break; }
                    mbar_arrive_expect_tx(&full_bar[stage], PL * A_TILE);
#pragma unroll
                    for (int pl = 0; pl < PL; ++pl)
                        tc::tma_load_2d(a_s + (stage * PL + pl) * A_TILE, &map_ring, kb * KBE,
                                        (int)(((size_t)(slot * PL + pl)) * ring_rows_per_plane + (size_t)group * MV), &full_bar[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
                }
                TPH(3);
            }
            TPH_STORE(p.status, 0);
        }
    } else if (warp == 1) {
        // ---- MMA issuer ------------------------------------------------------------------------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            TPH_DECL
            for (int t = 1; t < T && ok; ++t) {
                wait.step = t;
                TPH(0);
                if (!wait.barrier(&acc_empty, ((uint32_t)(t - 1) & 1u) ^ 1u)) break;   // epilogue of step t-1 has drained TMEM
                tc::fence_after();
                TPH(1);
                for (int kb = 0; kb < NKB; ++kb) {
                    if (!wait.barrier(&full_bar[stage], phase)) { ok = false; break; }
                    tc::fence_after();
                    if (kb == 0) {
                        // the exchange of this step is over: the epilogue warps may now issue their bulk global traffic
                        // (stash stores of step t-1, x-projection loads of step t) without delaying anybody's hand-over
                        tc::mbar_arrive(&go_bar);
                        TPH(2);
                        TSTAMP(p.counters, t, 3);
                    } else {
                        TPH(3);
                    }
                    const uint32_t a_hi = a_s + (stage * PL) * A_TILE, a_lo = a_hi + A_TILE;
                    const uint32_t w_hi = base + kb * W_TILE, w_lo = w_hi + W_PLANE;
#pragma unroll
                    for (int k16 = 0; k16 < KBE / 16; ++k16) {
                        const uint64_t ah = tc::smem_desc_sw128(a_hi + k16 * 32), wh = tc::smem_desc_sw128(w_hi + k16 * 32);
                        const int half = kb >= NKB / 2 ? 1 : 0;
                        tc::umma_f16(tmem_base + half * NR, ah, wh, IDESC, (kb % (NKB / 2) > 0 || k16 > 0) ? 1u : 0u);
                        if (PASSES == 3) {
                            const uint64_t al = tc::smem_desc_sw128(a_lo + k16 * 32), wl = tc::smem_desc_sw128(w_lo + k16 * 32);
                            tc::umma_f16(tmem_base + 2 * NR, al, wh, IDESC, (kb > 0 || k16 > 0) ? 1u : 0u);
                            tc::umma_f16(tmem_base + 2 * NR, ah, wl, IDESC, 1u);
                        }
                    }
                    tc::umma_commit(&empty_bar[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
                    TPH(4);
                }
                if (ok) tc::umma_commit(&acc_full);
            }
            TPH_STORE(p.status, 1);
        }
    } else {
        // ---- epilogue: thread = (video = TMEM lane, half of the CTA's units) ----------------------------------------------
        const int q = warp & 3, hf = (warp - 2) >> 2, v = q * 32 + lane;
        const int b = group * MV + v;
        const bool valid = b < p.B;
        const size_t row0 = (size_t)(valid ? b : 0) * T;
        const int uh0 = u0 + hf * UH;
        float c[UH];
#pragma unroll
        for (int u = 0; u < UH; ++u) c[u] = 0.0f;
        unsigned int* my_counter = p.counters + (size_t)group * NKB + slice / (KBE / UC);
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        TPH_DECL
        // x-projection of a step: 4 gates x 8 units, loaded one step ahead (behind the go barrier, see below)
        float4 xp[4][2];
        auto load_xproj = [&](int t) {
            const float* xrow = p.xproj + (row0 + t) * (size_t)(4 * H) + uh0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                xp[g][0] = valid ? ldg4(xrow + g * H) : zero4;
                xp[g][1] = valid ? ldg4(xrow + g * H + 4) : zero4;
            }
        };
        load_xproj(0);
        // internal stash layout of the tcgen05 path (written here, read by lstm_bwd_tc_kernel only): video-minor float4
        // groups, gates [t][slice][half][gate][j][B][4], cells [t][slice][half][j][B][4] -- a warp stores 512 contiguous bytes
        const size_t Bz = (size_t)p.B;
        for (int t = 0; t < T; ++t) {
            wait.step = t;
            TPH(0);
            float d[32];      // pre-activations, n' = gate*8 + uu
            if (t > 0) {
                if (!wait.barrier(&acc_full, (uint32_t)(t - 1) & 1u)) break;
                tc::fence_after();
                TPH(1);
                if (tid == 64) TSTAMP(p.counters, t + 1, 0);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 32);
                uint32_t x0[32], x1[32];
                tc::tmem_ld32(taddr, x0);
                tc::tmem_ld32(taddr + NR, x1);
                if (NACC == 3) {
                    uint32_t x2[32];
                    tc::tmem_ld32(taddr + 2 * NR, x2);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) d[i] = ((__uint_as_float(x0[i]) + __uint_as_float(x1[i])) + __uint_as_float(x2[i])) * winv;
                } else {
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) d[i] = (__uint_as_float(x0[i]) + __uint_as_float(x1[i])) * winv;
                }
                tc::fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc_empty);
                TPH(2);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) d[i] = 0.0f;
            }
            const float xv[4][UH] = {{xp[0][0].x, xp[0][0].y, xp[0][0].z, xp[0][0].w, xp[0][1].x, xp[0][1].y, xp[0][1].z, xp[0][1].w},
                                     {xp[1][0].x, xp[1][0].y, xp[1][0].z, xp[1][0].w, xp[1][1].x, xp[1][1].y, xp[1][1].z, xp[1][1].w},
                                     {xp[2][0].x, xp[2][0].y, xp[2][0].z, xp[2][0].w, xp[2][1].x, xp[2][1].y, xp[2][1].z, xp[2][1].w},
                                     {xp[3][0].x, xp[3][0].y, xp[3][0].z, xp[3][0].w, xp[3][1].x, xp[3][1].y, xp[3][1].z, xp[3][1].w}};
            float gi[UH], gf[UH], gg[UH], go[UH], hv[UH];
#pragma unroll
            for (int u = 0; u < UH; ++u) {
                gi[u] = sigmoid_sfu(d[u] + xv[0][u]);
                gf[u] = sigmoid_sfu(d[UH + u] + xv[1][u]);
                gg[u] = tanh_sfu(d[2 * UH + u] + xv[2][u]);
                go[u] = sigmoid_sfu(d[3 * UH + u] + xv[3][u]);
                c[u] = fmaf(gf[u], c[u], gi[u] * gg[u]);
                hv[u] = go[u] * tanh_sfu(c[u]);
            }
            TPH(3);
            if (t + 1 < T) {
                // publish h_t FIRST (the release below waits for every earlier store of the warp: the stash stores follow
                // it): 8 units x fp16 = 16 bytes per plane; videos past the end of the batch publish zeros
                uint4 hi, lo;
                tc::split_pair(hv[0], hv[1], hi.x, lo.x);
                tc::split_pair(hv[2], hv[3], hi.y, lo.y);
                tc::split_pair(hv[4], hv[5], hi.z, lo.z);
                tc::split_pair(hv[6], hv[7], hi.w, lo.w);
                if (!valid) hi = lo = make_uint4(0u, 0u, 0u, 0u);
                __half* dst = p.ring + (((size_t)((t & 1) * PL) * ring_rows_per_plane + (size_t)group * MV + v) * H + uh0);
                *reinterpret_cast<uint4*>(dst) = hi;
                if (PL == 2) *reinterpret_cast<uint4*>(dst + ring_rows_per_plane * H) = lo;
                __syncwarp();
                if (lane == 0) tc::publish(my_counter);
                if (tid == 64) TSTAMP(p.counters, t + 1, 1);
            }
            TPH(4);
            // bulk global traffic only once the hand-over of the step is through (the MMA warp has its first tile of step
            // t+1): before, the (uncoalesced: lane = video) stores below queued in front of the counter polls and the
            // hand-over took 2-5 us instead of 1 (tools/lstm_tc_skew.py, profiles/r02_lstm_tc_*.log)
            if (t + 1 < T && !wait.barrier(&go_bar, (uint32_t)t & 1u)) break;
            if (valid) {
                const size_t row = row0 + t;
                float* hp = p.hs + row * H + uh0;
                *reinterpret_cast<float4*>(hp) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(hp + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                if (p.gates) {
                    float4* gp = reinterpret_cast<float4*>(p.gates) + ((((size_t)t * NSL + slice) * 2 + hf) * 8) * Bz + (size_t)b;
                    gp[0 * Bz] = make_float4(gi[0], gi[1], gi[2], gi[3]);
                    gp[1 * Bz] = make_float4(gi[4], gi[5], gi[6], gi[7]);
                    gp[2 * Bz] = make_float4(gf[0], gf[1], gf[2], gf[3]);
                    gp[3 * Bz] = make_float4(gf[4], gf[5], gf[6], gf[7]);
                    gp[4 * Bz] = make_float4(gg[0], gg[1], gg[2], gg[3]);
                    gp[5 * Bz] = make_float4(gg[4], gg[5], gg[6], gg[7]);
                    gp[6 * Bz] = make_float4(go[0], go[1], go[2], go[3]);
                    gp[7 * Bz] = make_float4(go[4], go[5], go[6], go[7]);
                    float4* cp = reinterpret_cast<float4*>(p.cells) + ((((size_t)t * NSL + slice) * 2 + hf) * 2) * Bz + (size_t)b;
                    cp[0] = make_float4(c[0], c[1], c[2], c[3]);
                    cp[Bz] = make_float4(c[4], c[5], c[6], c[7]);
                }
            }
            if (t + 1 < T) load_xproj(t + 1);
        }
        if (tid == 64) TPH_STORE(p.status, 2);
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ======================================================= backward ======================================================
// RED: the reduce-scatter of the partial products is done by the L2 (vector atomic adds into one [128 x 16] accumulation
// block per consumer and slot, zeroed by its consumer after reading) instead of by the consumer summing H/16 stored blocks
template <int H, int PASSES, bool RED>
__global__ void __launch_bounds__(TCT, 1) lstm_bwd_tc_kernel(const TcBwdParams p) {
    constexpr int PL = PASSES == 3 ? 2 : 1, NSL = H / UC, NHALF = H / 256;
    constexpr int W_PLANE = H * 128;
    constexpr uint32_t IDESC = tc::idesc_f16(MV, 256);
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t a_full, acc_full, acc_empty;
    __shared__ uint32_t tmem_base_s;
    __shared__ int abort_flag;
    __shared__ float red_s[TCT / 32];
    __shared__ float amax_s[2][MV];      // per-video max |d gates| of each half of the units (the row scale is common)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    unsigned char* w_s = sm;                                    // [PL][H rows x 128 B]   W_hh[own rows]^T
    unsigned char* a_tile = sm + PL * W_PLANE;                  // [PL][128 x 128 B]      d gates of the step
    const uint32_t a_addr = base + PL * W_PLANE;
    const int slice = blockIdx.x % NSL, group = p.group_offset + blockIdx.x / NSL;
    const int u0 = slice * UC, T = p.T;

    if (tid == 0) {
        mbar_init(&a_full, NEW);
        mbar_init(&acc_full, 1);
        mbar_init(&acc_empty, NEW);
        mbar_fence_init();
        abort_flag = 0;
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base_s, H);
    const float winv = stage_weights<H, PL, true>(p.w_hh, u0, w_s, red_s);
    fence_proxy_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    Waiter wait = {&abort_flag, p.status, 0};
    constexpr size_t kBlock = (size_t)MV * UC;                         // floats of one (consumer, producer) block
    constexpr size_t kSlot = (size_t)NSL * (RED ? 1 : NSL) * kBlock;   // floats per slot and group
    float* ring_g = p.ring + (size_t)group * 2 * kSlot;
    unsigned int* counter = p.counters + group;

    if (warp == 1) {
        // ---- MMA issuer: partial dh_{t-1}[128 x H] = d gates_t[128 x 64] . W_hh[own 64 rows][H] ------------------------
        if (lane == 0) {
            for (int s = 0; s + 1 < T; ++s) {       // step s handles frame t = T-1-s; the product is needed for t >= 1
                wait.step = s;
                if (!wait.barrier(&a_full, (uint32_t)s & 1u)) break;           // d gates tile written (4 epilogue warps)
                if (!wait.barrier(&acc_empty, ((uint32_t)s & 1u) ^ 1u)) break; // the previous partial has left TMEM
                tc::fence_after();
#pragma unroll
                for (int k16 = 0; k16 < NR / 16; ++k16) {
                    const uint64_t ah = tc::smem_desc_sw128(a_addr + k16 * 32), al = tc::smem_desc_sw128(a_addr + A_TILE + k16 * 32);
#pragma unroll
                    for (int hf = 0; hf < NHALF; ++hf) {
                        const uint32_t wb = base + hf * 256 * 128 + k16 * 32;
                        const uint64_t wh = tc::smem_desc_sw128(wb), wl = tc::smem_desc_sw128(wb + W_PLANE);
                        const uint32_t dcol = tmem_base + hf * 256;
                        tc::umma_f16(dcol, ah, wh, IDESC, k16 > 0 ? 1u : 0u);
                        if (PASSES == 3) {
                            tc::umma_f16(dcol, al, wh, IDESC, 1u);
                            tc::umma_f16(dcol, ah, wl, IDESC, 1u);
                        }
                    }
                }
                tc::umma_commit(&acc_full);
            }
        }
    } else if (warp >= 2) {
        // ---- epilogue: thread = (video = TMEM lane, half of the CTA's units / of the H output columns) -----------------------
        const int q = warp & 3, hf = (warp - 2) >> 2, v = q * 32 + lane;
        const int b = group * MV + v;
        const bool valid = b < p.B;
        const size_t row0 = (size_t)(valid ? b : 0) * T;
        const int uh0 = u0 + hf * UH;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float dc[UH];
#pragma unroll
        for (int u = 0; u < UH; ++u) dc[u] = 0.0f;
        // A failed wait (time-out / abort) does NOT leave the loop: the two warps of a quadrant meet at a named barrier every
        // step, so both keep stepping -- every later wait returns at once because the abort flag is set -- and the launch
        // ends with garbage and a non-zero status word instead of a hang.
        bool dead = false;
        TPH_DECL
        for (int s = 0; s < T; ++s) {
            wait.step = s;
            const int t = T - 1 - s;
            const size_t row = row0 + t;
            TPH(0);
            // ---- stash of the frame (issued before the wait on the partial products) ----------------------------------
            float4 st[7][2];     // i, f, g, o, c, c_prev, dh
            {
                const size_t Bz = (size_t)p.B, bb = valid ? (size_t)b : 0;
                const float4* gp = reinterpret_cast<const float4*>(p.gates) + ((((size_t)t * NSL + slice) * 2 + hf) * 8) * Bz + bb;
                const float4* cp = reinterpret_cast<const float4*>(p.cells) + ((((size_t)t * NSL + slice) * 2 + hf) * 2) * Bz + bb;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) st[g][j] = valid ? __ldg(gp + (2 * g + j) * Bz) : zero4;
                    st[4][j] = valid ? __ldg(cp + j * Bz) : zero4;
                    st[5][j] = (valid && t > 0) ? __ldg(cp + j * Bz - (size_t)NSL * 4 * Bz) : zero4;    // cells of frame t-1
                    st[6][j] = valid ? ldg4(p.dh_out + row * H + uh0 + 4 * j) : zero4;
                }
            }
            // ---- recurrent part of dh_t: the producers' partial products of step s-1 for the own units ----------------
            float4 acc0 = zero4, acc1 = zero4;
            if (s > 0) {
                bool ok = true;
                if (lane == 0 && !dead) ok = wait.counter(counter, (unsigned)(NEW * NSL) * (unsigned)s);
                if (dead) ok = false;
                ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
                if (!ok) dead = true;
                TPH(1);
                // block (consumer, producer): [4 float4 groups][128 videos][4 floats] -- lane = video, 512 contiguous bytes per warp load
                if constexpr (RED) {
                    float* src = ring_g + (size_t)((s - 1) & 1) * kSlot + (size_t)slice * kBlock + (size_t)(2 * hf) * (MV * 4) + (size_t)v * 4;
                    acc0 = ld_cg4(src), acc1 = ld_cg4(src + MV * 4);
                    // hand the block back zeroed: the producers add into this slot again two steps from now, after they
                    // have acquired the counter this thread releases at the end of the step
                    *reinterpret_cast<float4*>(src) = zero4;
                    *reinterpret_cast<float4*>(src + MV * 4) = zero4;
                } else {
                    const float* src = ring_g + (size_t)((s - 1) & 1) * kSlot + (size_t)slice * NSL * kBlock + (size_t)(2 * hf) * (MV * 4) + (size_t)v * 4;
#pragma unroll 8      // 16 loads in flight per thread; 32 (unroll 16) ran 1.8x slower: 12880 against 7266 clocks for this phase
                    for (int pr = 0; pr < NSL; ++pr) {
                        const float4 x0 = ld_cg4(src + (size_t)pr * kBlock), x1 = ld_cg4(src + (size_t)pr * kBlock + MV * 4);
                        acc0.x += x0.x, acc0.y += x0.y, acc0.z += x0.z, acc0.w += x0.w;
                        acc1.x += x1.x, acc1.y += x1.y, acc1.z += x1.z, acc1.w += x1.w;
                    }
                }
            }
            TPH(2);
            // ---- cell backward -> d(pre-activation gates), n' = gate*8 + uu ---------------------------------------------
            const float sv[7][UH] = {
                {st[0][0].x, st[0][0].y, st[0][0].z, st[0][0].w, st[0][1].x, st[0][1].y, st[0][1].z, st[0][1].w},
                {st[1][0].x, st[1][0].y, st[1][0].z, st[1][0].w, st[1][1].x, st[1][1].y, st[1][1].z, st[1][1].w},
                {st[2][0].x, st[2][0].y, st[2][0].z, st[2][0].w, st[2][1].x, st[2][1].y, st[2][1].z, st[2][1].w},
                {st[3][0].x, st[3][0].y, st[3][0].z, st[3][0].w, st[3][1].x, st[3][1].y, st[3][1].z, st[3][1].w},
                {st[4][0].x, st[4][0].y, st[4][0].z, st[4][0].w, st[4][1].x, st[4][1].y, st[4][1].z, st[4][1].w},
                {st[5][0].x, st[5][0].y, st[5][0].z, st[5][0].w, st[5][1].x, st[5][1].y, st[5][1].z, st[5][1].w},
                {st[6][0].x + acc0.x, st[6][0].y + acc0.y, st[6][0].z + acc0.z, st[6][0].w + acc0.w,
                 st[6][1].x + acc1.x, st[6][1].y + acc1.y, st[6][1].z + acc1.z, st[6][1].w + acc1.w}};
            float dg[4 * UH];
            float amax = 0.0f;
#pragma unroll
            for (int u = 0; u < UH; ++u) {
                const float gi = sv[0][u], gf = sv[1][u], gg = sv[2][u], go = sv[3][u], dh = sv[6][u];
                const float tcv = tanh_sfu(sv[4][u]);
                const float d_o = dh * tcv;
                const float dcv = fmaf(dh * go, 1.0f - tcv * tcv, dc[u]);
                dc[u] = dcv * gf;
                dg[u] = dcv * gg * gi * (1.0f - gi);
                dg[UH + u] = dcv * sv[5][u] * gf * (1.0f - gf);
                dg[2 * UH + u] = dcv * gi * (1.0f - gg * gg);
                dg[3 * UH + u] = d_o * go * (1.0f - go);
            }
            if (t > 0) {
#pragma unroll
                for (int i = 0; i < 4 * UH; ++i) amax = fmaxf(amax, fabsf(dg[i]));
                // per-video power-of-two scale, common to both halves of the row (the two warps of a TMEM quadrant meet at
                // a 64-thread named barrier): the fp16 hi / lo pair keeps 22 bits whatever the magnitude of the gradient.
                // (The MMAs of step s-1 have finished reading the d gates tile: this thread waited for their commit,
                // acc_full, before it read their result out of TMEM in the previous iteration.)
                amax_s[hf][v] = amax;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
                amax = fmaxf(amax_s[0][v], amax_s[1][v]);
                float sc, inv;
                pow2_scale(amax, 11, sc, inv);
                inv *= winv;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 hi, lo;
                    tc::split_pair(dg[g * UH] * sc, dg[g * UH + 1] * sc, hi.x, lo.x);
                    tc::split_pair(dg[g * UH + 2] * sc, dg[g * UH + 3] * sc, hi.y, lo.y);
                    tc::split_pair(dg[g * UH + 4] * sc, dg[g * UH + 5] * sc, hi.z, lo.z);
                    tc::split_pair(dg[g * UH + 6] * sc, dg[g * UH + 7] * sc, hi.w, lo.w);
                    const uint32_t off = tc::sw128_offset(v, hf * 4 + g);
                    *reinterpret_cast<uint4*>(a_tile + off) = hi;
                    if (PL == 2) *reinterpret_cast<uint4*>(a_tile + A_TILE + off) = lo;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&a_full);
                amax = inv;     // carried to the read-out below
            }
            TPH(3);
            if (valid) {     // behind the hand-over to the MMA warp: these stores overlap the product
                float* dp = p.dgates + row * (size_t)(4 * H) + uh0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    *reinterpret_cast<float4*>(dp + g * H) = make_float4(dg[g * UH], dg[g * UH + 1], dg[g * UH + 2], dg[g * UH + 3]);
                    *reinterpret_cast<float4*>(dp + g * H + 4) = make_float4(dg[g * UH + 4], dg[g * UH + 5], dg[g * UH + 6], dg[g * UH + 7]);
                }
            }
            if (t == 0) break;      // no frame before the first: nothing to send back
            const float inv = amax;
            // ---- the CTA's partial dh_{t-1}[video][this half of the H columns]: TMEM -> ring, block (consumer, this producer)
            if (!dead && !wait.barrier(&acc_full, (uint32_t)s & 1u)) dead = true;
            tc::fence_after();
            TPH(4);
            float* dst = ring_g + (size_t)(s & 1) * kSlot + (RED ? (size_t)0 : (size_t)slice * kBlock) + (size_t)v * 4;
            const int col0 = hf * (H / 2);
#pragma unroll 2
            for (int col = col0; col < col0 + H / 2; col += 32) {
                uint32_t x[32];
                tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, x);
                tc::tmem_ld_wait();
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {     // two consumers of 16 units per 32 columns
                    float* o = dst + (size_t)(col / UC + cb) * (RED ? 1 : NSL) * kBlock;
#pragma unroll
                    for (int j = 0; j < UC / 4; ++j) {
                        const float y0 = __uint_as_float(x[16 * cb + 4 * j]) * inv, y1 = __uint_as_float(x[16 * cb + 4 * j + 1]) * inv;
                        const float y2 = __uint_as_float(x[16 * cb + 4 * j + 2]) * inv, y3 = __uint_as_float(x[16 * cb + 4 * j + 3]) * inv;
                        if constexpr (RED)
                            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + (size_t)j * (MV * 4)), "f"(y0), "f"(y1), "f"(y2), "f"(y3) : "memory");
                        else
                            *reinterpret_cast<float4*>(o + (size_t)j * (MV * 4)) = make_float4(y0, y1, y2, y3);
                    }
                }
            }
            tc::fence_before();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&acc_empty);
                asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
            }
            TPH(5);
        }
        if (tid == 64) TPH_STORE(p.status, 3);
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, H);
}

// ---- host side ------------------------------------------------------------------------------------------------------
struct TcLayout {
    size_t counters_off, ring_off, total;
};
TcLayout tc_layout(int64_t B, int64_t H) {
    const size_t groups = (size_t)((B + MV - 1) / MV);
    TcLayout l;
    l.counters_off = 4096;
    l.ring_off = 8192;
    const size_t fwd = 2 * 2 * groups * MV * (size_t)H * sizeof(__half);
    const size_t nsl = (size_t)(H / UC);
    const size_t bwd = 2 * groups * nsl * nsl * MV * UC * sizeof(float);
    l.total = l.ring_off + (fwd > bwd ? fwd : bwd);
    return l;
}

template <typename Kernel>
int tc_capacity(Kernel kernel, size_t smem, int* cap) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> cache;
    int dev = 0;
    OPN_CUDA(cudaGetDevice(&dev));
    const std::pair<const void*, int> key((const void*)kernel, dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *cap = it->second;
            return OPN_OK;
        }
    }
    const int rc = max_coresident(kernel, TCT, smem, cap);
    if (rc != OPN_OK) return rc;
    std::lock_guard<std::mutex> lock(mu);
    cache[key] = *cap;
    return OPN_OK;
}

template <int H, int PASSES>
int run_fwd_tc(const FwdParams& p0, int64_t B, char* ws, cudaStream_t s) {
    constexpr int PL = PASSES == 3 ? 2 : 1, NSL = H / UC;
    constexpr size_t smem = 1024 + (size_t)PL * (H / KBE) * NR * 128 + (size_t)NSTAGE * PL * A_TILE;
    const TcLayout l = tc_layout(B, H);
    const int groups = (int)((B + MV - 1) / MV);
    int cap = 0;
    int rc = tc_capacity(lstm_fwd_tc_kernel<H, PASSES>, smem, &cap);
    if (rc != OPN_OK) return rc;
    const int per_launch = cap / NSL;
    if (per_launch < 1) {
        set_error("lstm_fwd (tcgen05): device cannot co-schedule %d CTAs (capacity %d)", NSL, cap);
        return OPN_ERR_UNSUPPORTED;
    }
    OPN_CUDA(cudaMemsetAsync(ws, 0, l.total, s));
    TcFwdParams p;
    p.xproj = p0.xproj, p.w_hh = p0.w_hh, p.hs = p0.hs, p.gates = p0.gates, p.cells = p0.cells;
    p.ring = reinterpret_cast<__half*>(ws + l.ring_off);
    p.counters = reinterpret_cast<unsigned int*>(ws + l.counters_off);
    p.status = p0.status;
    p.B = p0.B, p.T = p0.T, p.groups_total = groups;
    CUtensorMap map;
    rc = make_map_16bit(&map, p.ring, (long long)2 * PL * groups * MV, H, MV, false);
    if (rc != OPN_OK) return rc;
    for (int g0 = 0; g0 < groups; g0 += per_launch) {
        const int ng = groups - g0 < per_launch ? groups - g0 : per_launch;
        p.group_offset = g0;
        void* args[] = {(void*)&map, (void*)&p};
        OPN_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_fwd_tc_kernel<H, PASSES>, dim3(NSL * ng), dim3(TCT), args, smem, s));
        count_launch();
    }
    return OPN_OK;
}

template <int H, int PASSES, bool RED>
int run_bwd_tc(const BwdParams& p0, int64_t B, char* ws, cudaStream_t s) {
    constexpr int PL = PASSES == 3 ? 2 : 1, NSL = H / UC;
    constexpr size_t smem = 1024 + (size_t)PL * H * 128 + (size_t)PL * A_TILE;
    const TcLayout l = tc_layout(B, H);
    const int groups = (int)((B + MV - 1) / MV);
    int cap = 0;
    int rc = tc_capacity(lstm_bwd_tc_kernel<H, PASSES, RED>, smem, &cap);
    if (rc != OPN_OK) return rc;
    const int per_launch = cap / NSL;
    if (per_launch < 1) {
        set_error("lstm_bwd (tcgen05): device cannot co-schedule %d CTAs (capacity %d)", NSL, cap);
        return OPN_ERR_UNSUPPORTED;
    }
    // status + counters; the stored-block ring is written before it is read, the accumulation blocks of the RED flavour
    // (2 slots x H/16 consumers x [128 x 16] floats per group) start from zero
    OPN_CUDA(cudaMemsetAsync(ws, 0, l.ring_off + (RED ? (size_t)groups * 2 * NSL * MV * UC * sizeof(float) : 0), s));
    TcBwdParams p;
    p.w_hh = p0.w_hh, p.gates = p0.gates, p.cells = p0.cells, p.dh_out = p0.dh_out, p.dgates = p0.dgates;
    p.ring = reinterpret_cast<float*>(ws + l.ring_off);
    p.counters = reinterpret_cast<unsigned int*>(ws + l.counters_off);
    p.status = p0.status;
    p.B = p0.B, p.T = p0.T, p.groups_total = groups;
    for (int g0 = 0; g0 < groups; g0 += per_launch) {
        const int ng = groups - g0 < per_launch ? groups - g0 : per_launch;
        p.group_offset = g0;
        void* args[] = {(void*)&p};
        OPN_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_bwd_tc_kernel<H, PASSES, RED>, dim3(NSL * ng), dim3(TCT), args, smem, s));
        count_launch();
    }
    return OPN_OK;
}

}  // namespace

// The batch-wide kernels exist for H = 256 / 512 and pay off once the 128-video groups fill the SMs.  Forward and backward
// share an internal stash layout, so one rule decides for both; measured on the B200 (tools/lstm_tc_time.py,
// profiles/r02_lstm_tc_time.log; forward + backward of one layer, T = 300):
//   H = 512: B = 128  mma.sync 5.8 ms / tcgen05 6.8 ms;  B = 256  11.6 / 6.9;  B = 512  23.2 / 7.3   -> from B = 192
//   H = 256: B = 256  mma.sync 4.3 ms / tcgen05 4.5 ms;  B = 512   8.0 / 4.5                         -> from B = 384
// OPN_LSTM_TC=0 keeps the mma.sync kernels for every batch, OPN_LSTM_TC=1 forces the tcgen05 kernels (tests).
bool lstm_tc_wanted(int64_t B, int64_t H) {
    if (H != 256 && H != 512) return false;
    const char* e = getenv("OPN_LSTM_TC");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return H == 512 ? B >= 192 : B >= 384;
}

size_t lstm_tc_workspace_bytes(int64_t B, int64_t H) { return (H == 256 || H == 512) ? tc_layout(B, H).total : 0; }

int lstm_fwd_tc(const FwdParams& p, int64_t B, int64_t H, void* workspace, cudaStream_t s) {
    char* ws = static_cast<char*>(workspace);
    const bool single = current_precision() == 1;
    if (H == 512) return single ? run_fwd_tc<512, 1>(p, B, ws, s) : run_fwd_tc<512, 3>(p, B, ws, s);
    return single ? run_fwd_tc<256, 1>(p, B, ws, s) : run_fwd_tc<256, 3>(p, B, ws, s);
}

int lstm_bwd_tc(const BwdParams& p, int64_t B, int64_t H, void* workspace, cudaStream_t s) {
    char* ws = static_cast<char*>(workspace);
    const bool single = current_precision() == 1;
    static int red = -1;     // OPN_LSTM_TC_RED=0: the consumers sum stored blocks (the first version)
    if (red < 0) {
        const char* e = getenv("OPN_LSTM_TC_RED");
        red = (e && e[0] == '0') ? 0 : 1;
    }
    if (red) {
        if (H == 512) return single ? run_bwd_tc<512, 1, true>(p, B, ws, s) : run_bwd_tc<512, 3, true>(p, B, ws, s);
        return single ? run_bwd_tc<256, 1, true>(p, B, ws, s) : run_bwd_tc<256, 3, true>(p, B, ws, s);
    }
    if (H == 512) return single ? run_bwd_tc<512, 1, false>(p, B, ws, s) : run_bwd_tc<512, 3, false>(p, B, ws, s);
    return single ? run_bwd_tc<256, 1, false>(p, B, ws, s) : run_bwd_tc<256, 3, false>(p, B, ws, s);
}

}  // namespace opn
