// Persistent LSTM recurrence with the per-step matvec on the tensor cores (mma.sync.m16n8k16, HMMA), sm_100a.
//
// Same decomposition, exchange protocol and outputs as the FP32-FMA kernels of opn_lstm.cu (one CTA = U hidden
// units x one batch group of 8 videos, W_hh slice stationary in registers, flag-in-data exchange); the
// [4U gate rows x H] . [H x 8 videos] product of a step (forward) and the [H columns x 4U rows] . [4U x 8 videos]
// product (backward) run as warp-level MMAs with N = 8 = the batch group:
//
//   * split precision: every fp32 operand x is carried as two fp16 numbers hi = fp16(x), lo = fp16(x - hi)
//     (22 significand bits) and a product is hi*hi + hi*lo + lo*hi, accumulated in fp32 by the tensor core.
//     W_hh is pre-scaled per CTA by a power of two (max |w| -> [2^11, 2^12)) so its lo parts stay normal;
//     h lies in (-1, 1) and needs no scale; the backward operand d(gates) is scaled per (video, step) by a
//     power of two taken from its largest magnitude in the CTA (one redux.sync), undone in fp32 afterwards.
//     Measured on B200 at K = 512 (tools/hmma_probe.cu, profiles/r01_hmma_probe.log): max error 1.1e-6 / rms 2.3e-7
//     against 1.4e-6 / 1.5e-7 for the fp32 FMA chain -- the split product is as accurate as the FFMA kernel
//     (bf16 splits are 6x worse and are not used);  hi*hi and the two cross terms go to separate accumulators.
//   * the weight fragments (A operand, hi and lo) fill the same registers the fp32 weights did: H/4 per thread.
//   * the recurrent operand is exchanged as flagged fp32 words laid out in B-fragment order
//     ([k-step][video][t][k 2t, 2t+1, 2t+8, 2t+9]): the 16-byte vector a thread polls is exactly one lane's
//     fragment of one k-step; it is split to fp16 and stored with one conflict-free STS.128.
//   * HMMA issue rate on B200 is 0.5 m16n8k16 per clock per SM (2048 dense FLOP/clk/SM): the 384 MMAs of an
//     H = 512 step take ~770 clocks where the FFMA tile + its 62-shuffle reduction took ~2500.
//
// Replaces the nn.LSTM calls of baselines/learned_models.py:39,46,76,113,146,192 (forward) and their autograd
// backward, as opn_lstm.cu does.
#include <stdlib.h>

#include "opn_mma_common.cuh"

namespace opn {

namespace {

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
// CTA: NT = 128*RG threads, U = 8*RG units, R = 32*RG gate rows (local row lr = unit*4 + gate), MT = 2*RG m-tiles
// of 16 rows; the KS = H/16 k-steps are split in two halves, warp = (m-tile, K half).  Per step:
//   gather (poll flagged fp32 fragments, split to fp16, STS.128)  | barrier |  KS/2 x 3 MMAs per warp, partial
//   pre-activations to shared memory  | barrier |  pointwise (two lanes per cell), publish h_t.
// SINGLE: the 1e-2 arithmetic mode (opn_set_precision): the hi.hi product alone, a third of the MMAs
template <int H, int RG, bool SINGLE>
__global__ void __launch_bounds__(kThreads* RG, 1) lstm_fwd_mma_kernel(const FwdParams p) {
    constexpr int NT = kThreads * RG;
    constexpr int NW = 4 * RG;
    constexpr int U = kUnits * RG;
    constexpr int R = 4 * U;
    constexpr int MT = R / 16;
    constexpr int KS = H / 16;
    constexpr int KPW = KS / 2;               // k-steps per warp
    constexpr int NV = KS * 32 / NT;          // fragment vectors polled per thread
    static_assert(NW == 2 * MT && KS % 2 == 0 && (KS * 32) % NT == 0, "warp tiling");

    __shared__ __align__(16) uint4 bfrag_s[KS * 32];   // h_{t-1} as fp16 hi/lo B fragments {b0hi, b1hi, b0lo, b1lo}
    __shared__ __align__(16) float d_s[2][R][8];       // partial pre-activations of the two K halves
    __shared__ float red_s[NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int slice = blockIdx.x % p.n_slices;
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int u0 = slice * U;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    uint32_t* ring = p.ring + (size_t)group * (2 * kGroup * H);

    for (int i = tid; i < KS * 32; i += NT) bfrag_s[i] = make_uint4(0u, 0u, 0u, 0u);  // absent videos stay zero

    // ---- weights: A fragments of this warp's m-tile for its K half, fp16 hi/lo of W * 2^s --------------------
    const int mt = warp % MT, kp = warp / MT;
    const int lr_a = mt * 16 + g, lr_b = lr_a + 8;
    const float* wrow_a = p.w_hh + (size_t)((lr_a & 3) * H + u0 + (lr_a >> 2)) * H;
    const float* wrow_b = p.w_hh + (size_t)((lr_b & 3) * H + u0 + (lr_b >> 2)) * H;
    float wmax = 0.0f;
#pragma unroll 4
    for (int j = 0; j < KPW; ++j) {
        const int k0 = 16 * (kp * KPW + j) + 2 * tq;
        const float2 a0 = __ldg(reinterpret_cast<const float2*>(wrow_a + k0));
        const float2 a1 = __ldg(reinterpret_cast<const float2*>(wrow_b + k0));
        const float2 a2 = __ldg(reinterpret_cast<const float2*>(wrow_a + k0 + 8));
        const float2 a3 = __ldg(reinterpret_cast<const float2*>(wrow_b + k0 + 8));
        wmax = fmaxf(wmax, fmaxf(fmaxf(fmaxf(fabsf(a0.x), fabsf(a0.y)), fmaxf(fabsf(a1.x), fabsf(a1.y))),
                                 fmaxf(fmaxf(fabsf(a2.x), fabsf(a2.y)), fmaxf(fabsf(a3.x), fabsf(a3.y)))));
    }
    float wscale, winv;
    weight_scale<NW>(wmax, red_s, wscale, winv);
    uint32_t ahi[KPW][4], alo[KPW][4];
#pragma unroll
    for (int j = 0; j < KPW; ++j) {
        const int k0 = 16 * (kp * KPW + j) + 2 * tq;
        const float2 a0 = __ldg(reinterpret_cast<const float2*>(wrow_a + k0));
        const float2 a1 = __ldg(reinterpret_cast<const float2*>(wrow_b + k0));
        const float2 a2 = __ldg(reinterpret_cast<const float2*>(wrow_a + k0 + 8));
        const float2 a3 = __ldg(reinterpret_cast<const float2*>(wrow_b + k0 + 8));
        split2(a0.x * wscale, a0.y * wscale, ahi[j][0], alo[j][0]);
        split2(a1.x * wscale, a1.y * wscale, ahi[j][1], alo[j][1]);
        split2(a2.x * wscale, a2.y * wscale, ahi[j][2], alo[j][2]);
        split2(a3.x * wscale, a3.y * wscale, ahi[j][3], alo[j][3]);
    }

    // ---- pointwise ownership: two lanes per (unit, video) cell, lane gh computes gates 2gh, 2gh+1 -----------
    // The cells are dealt to the lanes so that the values of one fragment vector of the exchange tile meet in one
    // warp and are published with a single vector store (4x / 2x fewer L2 store transactions than scalar words):
    //   U = 16: a vector = units {2q, 2q+1, 2q+8, 2q+9} of one video; warp w: q = w>>1, videos 4*(w&1)..+3;
    //           lane = (video bq : 2 bits | j : 2 bits | gh), unit = 2q + (j&1) + 8*(j>>1)
    //   U =  8: the CTA owns half a k-step: a pair = units {2q, 2q+1}; warp w: q = w; lane = (video : 3 | j : 1 | gh)
    const int gh = lane & 1;
    int ul, bl;
    if (U == 16) {
        const int j = (lane >> 1) & 3;
        ul = 2 * (warp >> 1) + (j & 1) + 8 * (j >> 1);
        bl = 4 * (warp & 1) + (lane >> 3);
    } else {
        ul = 2 * warp + ((lane >> 1) & 1);
        bl = lane >> 2;
    }
    const bool leader = (U == 16) ? ((lane & 7) == 0) : ((lane & 3) == 0);  // stores the vector of its lane group
    const int u = u0 + ul;
    const int bb = b0 + bl;
    const bool valid = bb < p.B;
    const size_t row0 = (size_t)(valid ? bb : 0) * T;
    const float* xp_ptr = p.xproj + row0 * (4 * H) + (size_t)(2 * gh) * H + u;
    const int pub_word = frag_word(bl, u);  // leader: first word of the vector (units 2q.. of the k-step half)
    const float s0 = gh ? 1.0f : 0.5f;

    float c_state = 0.0f;
    int my_abort = 0;
    float xp0 = 0.f, xp1 = 0.f;
    if (valid) {
        xp0 = __ldg(xp_ptr);
        xp1 = __ldg(xp_ptr + H);
    }
    __syncthreads();

    PH_DECL
    for (int t = 0; t < T; ++t) {
        float a0 = xp0, a1 = xp1;
#ifdef OPN_EXP_SYNC_AFTER_PUBLISH
        __syncthreads();
#endif
#ifdef OPN_EXP_SLEEP_BEFORE_POLL
        if (t > 0) __nanosleep(OPN_EXP_SLEEP_BEFORE_POLL);
#endif
        PH(0);  // publish + stash stores + prefetch of the previous step
        if (t > 0) {
            // ---- h_{t-1}: poll this thread's fragment vectors, split, store as fp16 fragments -----------------
            const uint32_t par = step_parity(t - 1);
            const uint32_t* src = ring + (size_t)((t - 1) & 1) * (kGroup * H);
            auto vec_valid = [&](int i) { return (((tid + NT * i) & 31) >> 2) < nvalid; };
            uint4 v[NV];
            if (!gather_flagged(v, [&](int i) { return src + (size_t)(tid + NT * i) * 4; }, vec_valid, par, p.status, t))
                my_abort = 1;
            PH(1);  // poll
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (vec_valid(i)) {
                    uint4 f;
                    split2(__uint_as_float(v[i].x), __uint_as_float(v[i].y), f.x, f.z);
                    split2(__uint_as_float(v[i].z), __uint_as_float(v[i].w), f.y, f.w);
                    bfrag_s[tid + NT * i] = f;
                }
            }
            if (__syncthreads_or(my_abort)) break;
            PH(2);  // split + STS + barrier

            // ---- this warp's 16 rows x its K half: hi*hi on two alternating chains, cross terms on two more --
            float dm0[4] = {0.f, 0.f, 0.f, 0.f}, dm1[4] = {0.f, 0.f, 0.f, 0.f};
            float ds0[4] = {0.f, 0.f, 0.f, 0.f}, ds1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < KPW; ++j) {
                const uint4 b = bfrag_s[(kp * KPW + j) * 32 + lane];
                if (j & 1)
                    mma_f16(dm1, ahi[j], b.x, b.y);
                else
                    mma_f16(dm0, ahi[j], b.x, b.y);
                if constexpr (!SINGLE) {
                    mma_f16(ds0, ahi[j], b.z, b.w);
                    mma_f16(ds1, alo[j], b.x, b.y);
                }
            }
            // D fragment: (row g, videos 2tq, 2tq+1), (row g+8, same videos)
            const float2 lo2 = make_float2(((dm0[0] + dm1[0]) + (ds0[0] + ds1[0])) * winv,
                                           ((dm0[1] + dm1[1]) + (ds0[1] + ds1[1])) * winv);
            const float2 hi2 = make_float2(((dm0[2] + dm1[2]) + (ds0[2] + ds1[2])) * winv,
                                           ((dm0[3] + dm1[3]) + (ds0[3] + ds1[3])) * winv);
            *reinterpret_cast<float2*>(&d_s[kp][lr_a][2 * tq]) = lo2;
            *reinterpret_cast<float2*>(&d_s[kp][lr_b][2 * tq]) = hi2;
            PH(3);  // MMAs
            __syncthreads();
            PH(4);  // barrier
            const int lr0 = ul * 4 + 2 * gh;
            a0 += d_s[0][lr0][bl] + d_s[1][lr0][bl];
            a1 += d_s[0][lr0 + 1][bl] + d_s[1][lr0 + 1][bl];
        }

        // fused pointwise: gate activations, cell update
        const float act0 = fmaf(s0, tanh_sfu(s0 * a0), 1.0f - s0);      // gh=0: i = sigmoid   gh=1: g = tanh
        const float act1 = fmaf(0.5f, tanh_sfu(0.5f * a1), 0.5f);       // gh=0: f             gh=1: o
        const float oth0 = __shfl_xor_sync(0xffffffffu, act0, 1);
        const float oth1 = __shfl_xor_sync(0xffffffffu, act1, 1);
        const float gi = gh ? oth0 : act0;
        const float gf = gh ? oth1 : act1;
        const float gg = gh ? act0 : oth0;
        const float go = gh ? act1 : oth1;
        c_state = fmaf(gf, c_state, gi * gg);
        const float hval = go * tanh_sfu(c_state);
        PH(5);  // pointwise
        // critical path first: publish h_t to the other CTAs of this batch group (one vector per lane group)
        {
            const uint32_t par = step_parity(t);
            const uint32_t w0 = flagged(hval, par);
            const uint32_t w1 = __shfl_xor_sync(0xffffffffu, w0, 2);
            uint32_t* dst = ring + (size_t)(t & 1) * (kGroup * H) + pub_word;
            if (U == 16) {
                const uint32_t w2 = __shfl_xor_sync(0xffffffffu, w0, 4);
                const uint32_t w3 = __shfl_xor_sync(0xffffffffu, w0, 6);
                if (leader && valid && t + 1 < T)
                    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(w0), "r"(w1), "r"(w2),
                                 "r"(w3)
                                 : "memory");
            } else {
                if (leader && valid && t + 1 < T)
                    asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1,%2};" ::"l"(dst), "r"(w0), "r"(w1) : "memory");
            }
        }
        if (valid) {
            const size_t row = row0 + t;
            if (gh == 0) {
                p.hs[row * H + u] = hval;
                if (p.gates) {
                    p.gates[row * (4 * H) + u] = gi;
                    p.gates[row * (4 * H) + H + u] = gf;
                }
            } else {
                if (p.gates) {
                    p.gates[row * (4 * H) + 2 * H + u] = gg;
                    p.gates[row * (4 * H) + 3 * H + u] = go;
                }
                if (p.cells) p.cells[row * H + u] = c_state;
            }
            if (t + 1 < T) {  // prefetch next step's input projection
                xp0 = __ldg(xp_ptr + (size_t)(t + 1) * (4 * H));
                xp1 = __ldg(xp_ptr + (size_t)(t + 1) * (4 * H) + H);
            }
        }
    }
    PH_STORE(p.status);
}

// ------------------------------------------------------------------------------------
// backward (reduce-scatter formulation, see opn_lstm.cu)
// ------------------------------------------------------------------------------------
// partial[k][b] = sum_{own rows lr} W_hh[row(lr)][k] * da[b][lr]:  M = H columns (H/16 m-tiles, H/(64*RG) per
// warp), N = 8 videos, K = R = 4U own gate rows (2*RG k-steps).  The B operand is da of the CTA's own cells, scaled
// per video to [2^10, 2^11) and split to fp16 by the lanes that compute it.
template <int H, int RG, bool SINGLE>
__global__ void __launch_bounds__(kThreads* RG, 1) lstm_bwd_mma_kernel(const BwdParams p) {
    constexpr int NT = kThreads * RG;
    constexpr int NW = 4 * RG;
    constexpr int U = kUnits * RG;
    constexpr int R = 4 * U;
    constexpr int KSB = R / 16;                // k-steps (own gate rows)
    constexpr int MPW = H / 16 / NW;           // m-tiles (16 columns of W_hh each) per warp
    constexpr int NS = H / U;                  // producers (slices) per batch group, <= 32
    constexpr int VPC = U / 4;                 // uint4 per (producer, video) holding this CTA's units
    static_assert(NS <= 32 && 8 * VPC == 4 * (NT / 32) && MPW >= 1, "tiling");

    __shared__ __align__(16) uint4 dafrag_s[2][KSB * 32];  // scaled da as fp16 hi/lo B fragments, double buffered
    __shared__ float dainv_s[2][8];                        // 1 / (per-video scale)
    __shared__ float red_s[NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int slice = blockIdx.x % p.n_slices;
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int u0 = slice * U;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    constexpr size_t kSlotWords = (size_t)NS * kGroup * H;
    uint32_t* ring = p.ring + (size_t)group * (2 * kSlotWords);

    for (int i = tid; i < 2 * KSB * 32; i += NT) (&dafrag_s[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 16) (&dainv_s[0][0])[tid] = 1.0f;

    // ---- weights: A[m = column k][kk = own row lr] = W_hh[row(lr)][k], local row lr = unit*4 + gate -------------
    auto wrow = [&](int lr) { return p.w_hh + (size_t)((lr & 3) * H + u0 + (lr >> 2)) * H; };
    float wmax = 0.0f;
    for (int mi = 0; mi < MPW; ++mi)
        for (int ks = 0; ks < KSB; ++ks) {
            const int kc = (warp * MPW + mi) * 16 + g, lr = 16 * ks + 2 * tq;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int l = lr + (q & 1) + 8 * (q >> 1);
                wmax = fmaxf(wmax, fmaxf(fabsf(__ldg(wrow(l) + kc)), fabsf(__ldg(wrow(l) + kc + 8))));
            }
        }
    float wscale, winv;
    weight_scale<NW>(wmax, red_s, wscale, winv);
    uint32_t ahi[MPW][KSB][4], alo[MPW][KSB][4];
#pragma unroll
    for (int mi = 0; mi < MPW; ++mi)
#pragma unroll
        for (int ks = 0; ks < KSB; ++ks) {
            const int kc = (warp * MPW + mi) * 16 + g, lr = 16 * ks + 2 * tq;
            // a0: (m g, kk 2t,2t+1)  a1: (m g+8, kk 2t,2t+1)  a2: (m g, kk 2t+8,+9)  a3: (m g+8, kk 2t+8,+9)
            split2(__ldg(wrow(lr) + kc) * wscale, __ldg(wrow(lr + 1) + kc) * wscale, ahi[mi][ks][0], alo[mi][ks][0]);
            split2(__ldg(wrow(lr) + kc + 8) * wscale, __ldg(wrow(lr + 1) + kc + 8) * wscale, ahi[mi][ks][1],
                   alo[mi][ks][1]);
            split2(__ldg(wrow(lr + 8) + kc) * wscale, __ldg(wrow(lr + 9) + kc) * wscale, ahi[mi][ks][2], alo[mi][ks][2]);
            split2(__ldg(wrow(lr + 8) + kc + 8) * wscale, __ldg(wrow(lr + 9) + kc + 8) * wscale, ahi[mi][ks][3],
                   alo[mi][ks][3]);
        }

    // Cell ownership after the reduction of step 3 (as in opn_lstm.cu):
    //   RG = 1: warp w gathers videos 2w, 2w+1; lane l ends with video 2w + (l>>4), unit 4*(l&1) + ((l>>2)&3)
    //   RG = 2: warp w gathers video w;         lane l ends with unit 4*(l&3) + ((l>>3)&3)
    // two lanes hold every cell (they differ in lane bit 1 resp. 2): `half` splits the gate work between them.
    int bl, ul, half;
    if (RG == 1) {
        bl = 2 * warp + (lane >> 4);
        ul = 4 * (lane & 1) + ((lane >> 2) & 3);
        half = (lane >> 1) & 1;
    } else {
        bl = warp;
        ul = 4 * (lane & 3) + ((lane >> 3) & 3);
        half = (lane >> 2) & 1;
    }
    const unsigned video_mask = (RG == 1) ? ((lane & 16) ? 0xffff0000u : 0x0000ffffu) : 0xffffffffu;
    const int u = u0 + ul;
    const int bb = b0 + bl;
    const bool valid = bb < p.B;
    const size_t row0 = (size_t)(valid ? bb : 0) * T;
    // fragment word of this lane's row pair lr = ul*4 + 2*half, +1:  k-step lr>>4, r = lr&15
    const int lr_pair = ul * 4 + 2 * half;
    const int da_word = 4 * ((lr_pair >> 4) * 32 + bl * 4 + (((lr_pair & 15) & 7) >> 1)) + ((lr_pair & 15) >> 3);
    // word offset of this thread's first published partial sum (m-tile warp*MPW, column g, video 2*tq) in a ring slot
    const int pub_off = (((warp * MPW) * (U == 16 ? 1 : 2)) * kGroup * NS + slice) * U + g + (2 * tq) * (NS * U);

    float dc_carry = 0.0f, dh_rec = 0.0f;
    // stash of the step being processed (s*) and of the next one (n*): prefetched two steps ahead, the loads of a
    // step are issued ~2 step times before their first use (one step ahead was measured to stall the cell backward
    // for ~1000 clocks: the stash streams from HBM behind the polling traffic)
    float si = 0.f, sf = 0.f, sg = 0.f, so = 0.f, sc = 0.f, scp = 0.f, sdh = 0.f;
    float ni = 0.f, nf = 0.f, ng = 0.f, no = 0.f, ndh = 0.f;
    auto load_next = [&](int t) {  // row t into n*; the cell value of row t is already held in scp
        const size_t row = row0 + t;
        const float* gp = p.gates + row * (size_t)(4 * H) + u;
        ni = __ldg(gp);
        nf = __ldg(gp + H);
        ng = __ldg(gp + 2 * H);
        no = __ldg(gp + 3 * H);
        ndh = __ldg(p.dh_out + row * H + u);
    };
    float ncp = 0.f;  // cells[t-2] for the step after next
    if (valid) {
        const size_t row = row0 + (T - 1);
        const float* gp = p.gates + row * (size_t)(4 * H) + u;
        si = __ldg(gp);
        sf = __ldg(gp + H);
        sg = __ldg(gp + 2 * H);
        so = __ldg(gp + 3 * H);
        sc = __ldg(p.cells + row * H + u);
        scp = (T > 1) ? __ldg(p.cells + (row - 1) * H + u) : 0.0f;
        sdh = __ldg(p.dh_out + row * H + u);
        if (T > 1) {
            load_next(T - 2);
            ncp = (T > 2) ? __ldg(p.cells + (row - 2) * H + u) : 0.0f;
        }
    }
    __syncthreads();

    int my_abort = 0;

    PH_DECL
    for (int t = T - 1; t >= 0; --t) {
        const int s = T - 1 - t;  // step number of the reverse recurrence: ring slot s&1, parity of s
        const int buf = s & 1;
        PH(0);  // reduction of the gathered partials
        // ---- 1. cell backward for the CTA's own units ---------------------------------------------------
        // Critical path first: d(pre-activation gates) -> scaled fp16 fragments in shared memory.  The exact copies
        // to `dgates` and the stash prefetch are issued behind the barrier, in the shadow of the MMAs.
        float v0 = 0.0f, v1 = 0.0f;
        if (valid) {
            const float dh = sdh + dh_rec;
            const float tc = tanh_sfu(sc);
            const float d_o = dh * tc;
            const float dc = fmaf(dh * so, 1.0f - tc * tc, dc_carry);
            const float d_i = dc * sg;
            const float d_g = dc * si;
            const float d_f = dc * scp;
            dc_carry = dc * sf;
            if (half == 0) {
                v0 = d_i * si * (1.0f - si);
                v1 = d_f * sf * (1.0f - sf);
            } else {
                v0 = d_g * (1.0f - sg * sg);
                v1 = d_o * so * (1.0f - so);
            }
            if (t > 0) {
                // per-video power-of-two scale from the largest |da| of the CTA's cells of this video
                const unsigned mbits =
                    __reduce_max_sync(video_mask, __float_as_uint(fmaxf(fabsf(v0), fabsf(v1))) & 0x7fffffffu);
                float sc2, inv2;
                pow2_scale(__uint_as_float(mbits), 11, sc2, inv2);
                uint32_t hi, lo;
                split2(v0 * sc2, v1 * sc2, hi, lo);
                uint32_t* w = reinterpret_cast<uint32_t*>(&dafrag_s[buf][0]) + da_word;
                w[0] = hi;
                w[2] = lo;
                if (ul == 0 && half == 0) dainv_s[buf][bl] = inv2 * winv;
            }
        }
        auto store_dgates = [&]() {
            float* dg = p.dgates + (row0 + t) * (size_t)(4 * H) + u + (half ? 2 * H : 0);
            dg[0] = v0;
            dg[H] = v1;
        };
        if (t == 0) {
            if (valid) store_dgates();
            break;
        }
        PH(1);  // cell backward
        if (__syncthreads_or(my_abort)) break;  // dafrag_s[buf] complete (double buffered: one barrier per step)
        PH(2);  // barrier
        if (valid) {
            store_dgates();
            // rotate: step t-1 becomes current, issue the loads of step t-2
            si = ni, sf = nf, sg = ng, so = no, sdh = ndh;
            sc = scp;
            scp = ncp;
            if (t > 1) {
                load_next(t - 2);
                ncp = (t > 2) ? __ldg(p.cells + (row0 + t - 3) * H + u) : 0.0f;
            }
        }

        // ---- 2. partial[k][b] over the own rows; publish ------------------------------------------------
        const uint32_t par = step_parity(s);
        uint32_t* slot = ring + (size_t)buf * kSlotWords;
        {
            uint4 bf[KSB];
#pragma unroll
            for (int ks = 0; ks < KSB; ++ks) bf[ks] = dafrag_s[buf][ks * 32 + lane];
            const float inv0 = dainv_s[buf][2 * tq], inv1 = dainv_s[buf][2 * tq + 1];
            uint32_t* pub = slot + pub_off;
#pragma unroll
            for (int mi = 0; mi < MPW; ++mi) {
                float dm[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < KSB; ++ks) {
                    mma_f16(dm, ahi[mi][ks], bf[ks].x, bf[ks].y);
                    if constexpr (!SINGLE) {
                        mma_f16(ds, ahi[mi][ks], bf[ks].z, bf[ks].w);
                        mma_f16(ds, alo[mi][ks], bf[ks].x, bf[ks].y);
                    }
                }
                // D: (column kc, videos 2tq, 2tq+1), (column kc+8, same videos), kc = (warp*MPW + mi)*16 + g;
                // column k belongs to consumer k / U, unit k % U: word [consumer][b][producer = slice][unit].
                // U = 16: consumer warp*MPW + mi, unit g + 8*(q>>1);  U = 8: consumer 2*(warp*MPW + mi) + (q>>1), unit g.
                // One base pointer per step, compile-time offsets, and no bounds check per word: videos past the end
                // of the batch are stored too (zeros; nobody reads them) -- 12 -> 5 instructions per published word.
                constexpr int kTileStride = (U == 16 ? 1 : 2) * kGroup * NS * U;   // words between consecutive m-tiles
                constexpr int kHalfStride = (U == 16) ? 8 : kGroup * NS * U;       // words between columns kc and kc + 8
                uint32_t* base = pub + mi * kTileStride;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    st_flagged(base + (q >> 1) * kHalfStride + (q & 1) * (NS * U), (dm[q] + ds[q]) * ((q & 1) ? inv1 : inv0), par);
            }
        }

        PH(3);  // MMAs + publish
        // ---- 3. reduce-scatter: sum the producers' partials for the own units ---------------------------
        {
            uint4 v[4];
            // load i of lane l:  RG=1: video 2w + (i>>1), vector (i&1)*32 + l  ->  producer vec/2, unit quad l&1
            //                    RG=2: video w,            vector i*32 + l      ->  producer vec/4, unit quad l&3
            auto vec_b = [&](int i) { return RG == 1 ? 2 * warp + (i >> 1) : warp; };
            auto vec_id = [&](int i) { return RG == 1 ? (i & 1) * 32 + lane : i * 32 + lane; };
            auto vec_valid = [&](int i) { return vec_id(i) < NS * VPC && vec_b(i) < nvalid; };
            const uint32_t* src = slot + (size_t)slice * kGroup * NS * U;
            if (!gather_flagged(
                    v, [&](int i) { return src + ((size_t)vec_b(i) * NS * VPC + vec_id(i)) * 4; }, vec_valid, par,
                    p.status, t))
                my_abort = 1;
            PH(4);  // poll
            float f[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = vec_valid(i);
                f[i][0] = ok ? __uint_as_float(v[i].x) : 0.0f;
                f[i][1] = ok ? __uint_as_float(v[i].y) : 0.0f;
                f[i][2] = ok ? __uint_as_float(v[i].z) : 0.0f;
                f[i][3] = ok ? __uint_as_float(v[i].w) : 0.0f;
            }
            if (RG == 1) {
                float r8[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    r8[j] = f[0][j] + f[1][j];
                    r8[4 + j] = f[2][j] + f[3][j];
                }
                butterfly_stage<8, 2, 8>(r8, (lane & 16) != 0, 16);
                butterfly_stage<4, 1, 8>(r8, (lane & 8) != 0, 8);
                butterfly_stage<2, 0, 8>(r8, (lane & 4) != 0, 4);
                dh_rec = r8[0] + __shfl_xor_sync(0xffffffffu, r8[0], 2);
            } else {
                float r4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r4[j] = (f[0][j] + f[1][j]) + (f[2][j] + f[3][j]);
                butterfly_stage<4, 1, 4>(r4, (lane & 16) != 0, 16);
                butterfly_stage<2, 0, 4>(r4, (lane & 8) != 0, 8);
                dh_rec = r4[0] + __shfl_xor_sync(0xffffffffu, r4[0], 4);
            }
        }
    }
    PH_STORE(p.status);
}

}  // namespace

// ---- dispatch (called from opn_lstm.cu) ----------------------------------------------------------------
bool lstm_mma_supported(int64_t H) { return H == 256 || H == 512; }

int current_precision();   // opn_api.cu

int lstm_fwd_mma(const FwdParams& p, int64_t B, int64_t H, cudaStream_t s) {
    const bool single = current_precision() == OPN_PRECISION_16BIT;
    if (H == 256) {
        const char* e = getenv("OPN_LSTM_FWD256_RG");
        if (e && e[0] == '2') return launch_ring(lstm_fwd_mma_kernel<256, 2, false>, p, 2 * kThreads, 16, 0, B, s, "lstm_fwd");
        if (single) return launch_ring(lstm_fwd_mma_kernel<256, 1, true>, p, kThreads, 32, 0, B, s, "lstm_fwd");
        return launch_ring(lstm_fwd_mma_kernel<256, 1, false>, p, kThreads, 32, 0, B, s, "lstm_fwd");
    }
    if (single) return launch_ring(lstm_fwd_mma_kernel<512, 2, true>, p, 2 * kThreads, 32, 0, B, s, "lstm_fwd");
    return launch_ring(lstm_fwd_mma_kernel<512, 2, false>, p, 2 * kThreads, 32, 0, B, s, "lstm_fwd");
}

int lstm_bwd_mma(const BwdParams& p, int64_t B, int64_t H, cudaStream_t s) {
    const bool single = current_precision() == OPN_PRECISION_16BIT;
    if (H == 256) {
        // 16-unit CTAs: 16 producers per batch group instead of 32 halve the reduce-scatter traffic (every producer sends
        // a partial [8, H] whatever its size): 1.92 against 2.31 us/step.  OPN_LSTM_BWD256_RG=1 selects 8-unit CTAs.
        const char* e = getenv("OPN_LSTM_BWD256_RG");
        if (e && e[0] == '1') return launch_ring(lstm_bwd_mma_kernel<256, 1, false>, p, kThreads, 32, 0, B, s, "lstm_bwd");
        if (single) return launch_ring(lstm_bwd_mma_kernel<256, 2, true>, p, 2 * kThreads, 16, 0, B, s, "lstm_bwd");
        return launch_ring(lstm_bwd_mma_kernel<256, 2, false>, p, 2 * kThreads, 16, 0, B, s, "lstm_bwd");
    }
    if (single) return launch_ring(lstm_bwd_mma_kernel<512, 2, true>, p, 2 * kThreads, 32, 0, B, s, "lstm_bwd");
    return launch_ring(lstm_bwd_mma_kernel<512, 2, false>, p, 2 * kThreads, 32, 0, B, s, "lstm_bwd");
}

}  // namespace opn
