// Persistent, weight-stationary LSTM recurrence for sm_100a (forward and backward).
//
// Replaces the nn.LSTM calls of baselines/learned_models.py:39,46,76,113,146,192 and their
// autograd backward.  Design:
//
//   * One cooperative launch runs all T steps.  A CTA (128 threads) owns 8 hidden units
//     (forward: their 32 gate rows of W_hh; backward: their 8 columns of W_hh, i.e. 8 rows
//     of W_hh^T) for one *batch group* of 8 videos.  The CTA's weight slice lives in
//     registers for the whole sequence (8 rows x KPT values per thread, KPT = H/32).
//   * Per step every CTA needs the full recurrent vector of its batch group (forward:
//     h_{t-1} [8,H]; backward: dgates_t [8,4H]).  It is exchanged through a 2-slot ring in
//     global memory (L2 resident) with a *ready bit carried in the data*: the producer
//     replaces the mantissa LSB of every exchanged fp32 word by the step parity and stores it
//     with st.relaxed.gpu; consumers ld.relaxed.gpu the words they need and retry until every
//     LSB shows the expected parity.  32-bit accesses are single-copy atomic, so a word is
//     either stale (old parity) or complete: no fence, no atomic, no inter-CTA barrier is on
//     the per-step critical path (the first version of this kernel used arrival counters +
//     red.release / ld.acquire + TMA bulk copies and measured 3.5 us per step at H=256 where
//     the FFMA work is 0.13 us; see profiles/r01_lstm_handoff.md).  The perturbation of the
//     recurrent operand is <= 1 ulp (6e-8 relative); the tensors handed to the rest of the
//     graph (hs, gates, cells, dgates) are stored exactly, off the critical path.
//   * Batch groups are independent chains; two CTAs are resident per SM so one chain's
//     exchange latency overlaps another chain's FFMA work.
//   * The [8 rows x 8 videos] partial sums of the thread tile are reduced across the K-split
//     with a 62-shuffle transposing butterfly.
//   * sigmoid/tanh, the cell update and the stash writes are fused into the step.
//
// Every spin wait has a clock64() time-out: on expiry the kernel records a status word and
// all CTAs leave the time loop (no hang, no trap).
#include <stdlib.h>

#include "opn_lstm_common.cuh"

namespace opn {

namespace {

template <int KPT, int NS>
__device__ __forceinline__ void load_weight_row(float (&w)[KPT], const float* __restrict__ row, int ks) {
    if (KPT >= 4) {
#pragma unroll
        for (int j = 0; j < KPT / 4; ++j) {
            float4 v = __ldg(reinterpret_cast<const float4*>(row + 4 * ks + (4 * NS) * j));
            w[4 * j + 0] = v.x;
            w[4 * j + 1] = v.y;
            w[4 * j + 2] = v.z;
            w[4 * j + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < KPT; ++e) w[e] = __ldg(row + ks + NS * e);
    }
}

// acc[rr*8 + b] += sum_e w[rr][e] * v(b, k(e)) for the operand tile v_s = [8 videos][H] (row-major) in shared
// memory.  Lane ks owns k = 4*ks + 128*j + i (KPT >= 4: one LDS.128 per (j, b), the warp reads 32 consecutive
// 16-byte words) or k = ks + 32*e (KPT < 4).
template <int KPT>
__device__ __forceinline__ void matvec_tile(const float (&w)[kUnits][KPT], const float* v_s, int lane, float (&acc)[64]) {
    constexpr int H = 32 * KPT;
    if (KPT >= 4) {
        const float4* v4 = reinterpret_cast<const float4*>(v_s);
#pragma unroll
        for (int j = 0; j < KPT / 4; ++j) {
#pragma unroll
            for (int b = 0; b < kGroup; ++b) {
                const float4 hv = v4[b * (H / 4) + j * 32 + lane];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    float a = acc[rr * 8 + b];
                    a = fmaf(w[rr][4 * j + 0], hv.x, a);
                    a = fmaf(w[rr][4 * j + 1], hv.y, a);
                    a = fmaf(w[rr][4 * j + 2], hv.z, a);
                    a = fmaf(w[rr][4 * j + 3], hv.w, a);
                    acc[rr * 8 + b] = a;
                }
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < KPT; ++e) {
#pragma unroll
            for (int b = 0; b < kGroup; ++b) {
                const float hv = v_s[b * H + e * 32 + lane];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) acc[rr * 8 + b] = fmaf(w[rr][e], hv, acc[rr * 8 + b]);
            }
        }
    }
}

// Sum acc[64] over the 32 lanes of the warp.  On return lane l holds in v[0], v[1] the totals
// of index  a = ((l>>4)&1)<<5 | (l&1)<<4 | j<<3 | ((l>>1)&7)   for j = 0, 1.
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[64], int lane) {
    butterfly_stage<64, 5, 64>(v, (lane & 16) != 0, 16);
    butterfly_stage<32, 2, 64>(v, (lane & 8) != 0, 8);
    butterfly_stage<16, 1, 64>(v, (lane & 4) != 0, 4);
    butterfly_stage<8, 0, 64>(v, (lane & 2) != 0, 2);
    butterfly_stage<4, 1, 64>(v, (lane & 1) != 0, 1);
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
// Shared-memory operand tile: h_{t-1} of the batch group as [8 videos][H], double buffered.
// RG = row groups per CTA (128*RG threads, 8*RG hidden units).  RG = 2 is used at H = 512 so that a
// batch of 32 videos is exactly one CTA per SM: two 128-thread CTAs per SM were measured to serialise
// (5.2 us/step against 3.0 us/step with one), because they share the SM's L2 request path while polling.
//
// CL = false: the exchange goes through the global-memory ring (any H, cooperative launch).
// CL = true : the H/(8*RG) <= 16 CTAs of a batch group form one thread-block cluster; every CTA pushes its
//             flagged h_t values directly into the operand tile of all CTAs of the cluster (DSMEM) and polls
//             only its own shared memory -- no L2 round trip on the step path, no cooperative launch (clusters
//             are co-scheduled by the hardware, batch groups are independent clusters).
template <int KPT, int RG, bool CL>
__global__ void __launch_bounds__(kThreads* RG, (RG == 1 && !CL) ? 2 : 1) lstm_fwd_kernel(const FwdParams p) {
    constexpr int H = 32 * KPT;
    constexpr int NT = kThreads * RG;
    constexpr int U = kUnits * RG;
    constexpr int CS = H / U;                          // CTAs per batch group (= cluster size when CL)
    constexpr int NV = (8 * H / 4 + NT - 1) / NT;      // 16-byte vectors of the operand tile per thread

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* h_s = reinterpret_cast<float*>(smem_raw);  // [2][8][H]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slice = blockIdx.x % p.n_slices;
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int u0 = slice * U;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    uint32_t* ring = CL ? nullptr : p.ring + (size_t)group * (2 * kGroup * H);

    for (int i = tid; i < 2 * kGroup * H; i += NT) h_s[i] = 0.0f;

    // this warp's 8 gate rows: rr = uu*4 + gate, unit = u0 + 2*warp + uu
    float w[8][KPT];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const int row = (rr & 3) * H + u0 + 2 * warp + (rr >> 2);
        load_weight_row<KPT, 32>(w[rr], p.w_hh + (size_t)row * H, lane);
    }
    __syncthreads();
    if (CL) cg::this_cluster().sync();  // every tile of the cluster is zeroed before the first remote store

    // after the butterfly this lane owns gates (2*gh, 2*gh+1) of cell (unit u, video bb)
    const int uu = lane >> 4, bl = (lane >> 1) & 7, gh = lane & 1;
    const int u = u0 + 2 * warp + uu;
    const int bb = b0 + bl;
    const bool valid = bb < p.B;
    const size_t row0 = (size_t)(valid ? bb : 0) * T;
    const float* xp_ptr = p.xproj + row0 * (4 * H) + (size_t)(2 * gh) * H + u;

    float c_state = 0.0f;
    int my_abort = 0;
    float xp0 = 0.f, xp1 = 0.f;
    if (valid) {
        xp0 = __ldg(xp_ptr);
        xp1 = __ldg(xp_ptr + H);
    }

    for (int t = 0; t < T; ++t) {
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.0f;

        if (t > 0) {
            // ---- h_{t-1} of this batch group: wait until every word of the tile carries the step parity ----
            const uint32_t par = step_parity(t - 1);
            float* dst = h_s + ((t - 1) & 1) * (kGroup * H);
            // vector idx = tid + NT*i covers video idx / (H/4), columns 4*(idx % (H/4)) ..
            auto vec_valid = [&](int i) { return (tid + NT * i) < 8 * H / 4 && (tid + NT * i) / (H / 4) < nvalid; };
            uint4 v[NV];
            if (CL) {
                const uint32_t tile = smem_u32(dst);
                if (!gather_flagged(v, [&](int i) { return tile + (uint32_t)(tid + NT * i) * 16u; }, vec_valid, par,
                                    p.status, t))
                    my_abort = 1;
            } else {
                const uint32_t* src = ring + (size_t)((t - 1) & 1) * (kGroup * H);
                if (!gather_flagged(v, [&](int i) { return src + (size_t)(tid + NT * i) * 4; }, vec_valid, par,
                                    p.status, t))
                    my_abort = 1;
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (vec_valid(i)) reinterpret_cast<uint4*>(dst)[tid + NT * i] = v[i];
            }
            if (__syncthreads_or(my_abort)) break;
            matvec_tile<KPT>(w, dst, lane, acc);
            warp_transpose_reduce(acc, lane);
        }

        // fused pointwise: gate activations, cell update
        const float a0 = acc[0] + xp0;
        const float a1 = acc[1] + xp1;
        const float act0 = gh ? tanhf(a0) : sigmoid_acc(a0);  // gh=0: i      gh=1: g
        const float act1 = sigmoid_acc(a1);                    // gh=0: f      gh=1: o
        const float oth0 = __shfl_xor_sync(0xffffffffu, act0, 1);
        const float oth1 = __shfl_xor_sync(0xffffffffu, act1, 1);
        const float gi = gh ? oth0 : act0;
        const float gf = gh ? oth1 : act1;
        const float gg = gh ? act0 : oth0;
        const float go = gh ? act1 : oth1;
        c_state = fmaf(gf, c_state, gi * gg);
        const float hval = go * tanhf(c_state);
        if (CL) {
            // critical path first: push h_t of the warp's unit pair into the tile of every CTA of the cluster.
            // All four lanes (uu, gh) of a video hold both values after one shuffle; each serves CS/4 peers with
            // 8-byte stores.
            const float other = __shfl_xor_sync(0xffffffffu, hval, 16);
            if (valid && t + 1 < T) {
                const uint32_t par = step_parity(t);
                const uint32_t x = flagged(uu ? other : hval, par), y = flagged(uu ? hval : other, par);
                const uint32_t local = smem_u32(h_s + (t & 1) * (kGroup * H) + bl * H + u0 + 2 * warp);
                constexpr int PER = (CS >= 4) ? CS / 4 : 1;
                const int q = uu * 2 + gh;
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int d = q * PER + i;
                    if (d < CS) st_peer_v2(map_to_cta(local, (uint32_t)d), x, y);
                }
            }
        }
        if (valid) {
            // critical path first: publish h_t to the other CTAs of this batch group
            if (!CL && gh == 0 && t + 1 < T)
                st_flagged(ring + (size_t)(t & 1) * (kGroup * H) + bl * H + u, hval, step_parity(t));
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
    if (CL) cg::this_cluster().sync();  // no CTA leaves while peers may still push into its shared memory
}

// ------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------
// Reduce-scatter formulation.  The CTA keeps the same slice of W_hh as the forward CTA -- the 4*U gate rows
// of its U = 8*RG hidden units, all H columns, in registers (thread tid owns columns KC*tid .. KC*tid+KC-1).
// Per step t (descending):
//   1. the lane pairs that own the U x 8 (unit, video) cells turn dh_t (upstream + recurrent) and the stash
//      into d(pre-activation gates) for the CTA's own units, store them (exactly) to `dgates` and into
//      shared memory [4U][8];
//   2. every thread accumulates partial[b][k] = sum_{own rows r} da[b][r] * W_hh[r][k] for its columns and
//      publishes the 8 x KC partial sums as flagged words into its producer region of the ring;
//   3. the CTA gathers, for its own units only, the partials of all H/U producers of the batch group and
//      sums them (in-thread adds + a short shuffle butterfly): dh_rec_{t-1}[unit, video] lands in the lane
//      pair that owns that cell.  The ring is laid out per CONSUMER, [consumer][video][producer][unit], so
//      this gather is one contiguous, fully coalesced block (4 uint4 per thread); with a per-producer
//      layout every warp load touched 32 different lines and H=256 ran at 4.05 us/step.
// Per CTA and step 8*H*4 bytes are written and read (16 KB at H=512) against 8*4H*4 bytes read by the
// gather formulation of the first version (64 KB, 6.05 us/step); no transposed copy of W_hh is needed.
//
// CL = true: the NS <= 16 CTAs of a batch group form one thread-block cluster; producers push their partial sums
// directly into the consumer's shared memory ([2][video][producer][unit], 8*H words per slot) and the consumer
// polls its own shared memory.
template <int KPT, int RG, bool CL>
__global__ void __launch_bounds__(kThreads* RG, (RG == 1 && !CL) ? 2 : 1) lstm_bwd_kernel(const BwdParams p) {
    constexpr int H = 32 * KPT;
    constexpr int NT = kThreads * RG;
    constexpr int U = kUnits * RG;            // hidden units of this CTA
    constexpr int R = 4 * U;                  // its gate rows
    constexpr int KC = (H >= NT) ? H / NT : 1;  // columns per active thread
    constexpr int NACT = H / KC;              // threads that take part in the matvec
    constexpr int NS = H / U;                 // producers (slices) per batch group, <= 32
    constexpr int VPC = U / 4;                // uint4 per (producer, video) holding this CTA's units
    static_assert(NS <= 32 && 8 * VPC == 4 * (NT / 32), "4 gather vectors per thread");

    __shared__ __align__(16) float da_s[2][R][8];
    __shared__ __align__(16) uint32_t rx_s[CL ? 2 * kGroup * H : 4];  // cluster flavour: inbox of partial sums

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slice = blockIdx.x % p.n_slices;
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int u0 = slice * U;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    constexpr size_t kSlotWords = (size_t)NS * kGroup * H;
    uint32_t* ring = CL ? nullptr : p.ring + (size_t)group * (2 * kSlotWords);

    for (int i = tid; i < 2 * R * 8; i += NT) (&da_s[0][0][0])[i] = 0.0f;  // rows of absent videos stay zero
    if (CL)
        for (int i = tid; i < 2 * kGroup * H; i += NT) rx_s[i] = 0u;

    // weights: local row r = gate*U + unit  ->  W_hh row gate*H + u0 + unit, columns KC*tid ..
    float w[R][KC];
    if (tid < NACT) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float* row = p.w_hh + (size_t)((r / U) * H + u0 + (r % U)) * H + KC * tid;
            if (KC == 2) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(row));
                w[r][0] = v.x;
                w[r][KC - 1] = v.y;
            } else {
                w[r][0] = __ldg(row);
            }
        }
    }

    // Cell ownership after the reduction of step 3 (see there):
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
    const int u = u0 + ul;
    const int bb = b0 + bl;
    const bool valid = bb < p.B;
    const size_t row0 = (size_t)(valid ? bb : 0) * T;

    float dc_carry = 0.0f, dh_rec = 0.0f;
    float si = 0.f, sf = 0.f, sg = 0.f, so = 0.f, sc = 0.f, scp = 0.f, sdh = 0.f;
    auto load_stash = [&](int t) {
        const size_t row = row0 + t;
        const float* g = p.gates + row * (size_t)(4 * H) + u;
        si = __ldg(g);
        sf = __ldg(g + H);
        sg = __ldg(g + 2 * H);
        so = __ldg(g + 3 * H);
        sc = __ldg(p.cells + row * H + u);
        scp = (t > 0) ? __ldg(p.cells + (row - 1) * H + u) : 0.0f;
        sdh = __ldg(p.dh_out + row * H + u);
    };
    if (valid) load_stash(T - 1);
    __syncthreads();
    if (CL) cg::this_cluster().sync();  // every inbox of the cluster is zeroed before the first remote store

    int my_abort = 0;

    for (int t = T - 1; t >= 0; --t) {
        const int s = T - 1 - t;  // step number of the reverse recurrence: ring slot s&1, parity of s
        const int buf = s & 1;
        // ---- 1. cell backward for the CTA's own units ---------------------------------------------------
        if (valid) {
            const float dh = sdh + dh_rec;
            const float tc = tanhf(sc);
            const float d_o = dh * tc;
            const float dc = fmaf(dh * so, 1.0f - tc * tc, dc_carry);
            const float d_i = dc * sg;
            const float d_g = dc * si;
            const float d_f = dc * scp;
            dc_carry = dc * sf;
            float* dg = p.dgates + (row0 + t) * (size_t)(4 * H) + u;
            if (half == 0) {
                const float ai = d_i * si * (1.0f - si);
                const float af = d_f * sf * (1.0f - sf);
                da_s[buf][ul][bl] = ai;
                da_s[buf][U + ul][bl] = af;
                dg[0] = ai;
                dg[H] = af;
            } else {
                const float ag = d_g * (1.0f - sg * sg);
                const float ao = d_o * so * (1.0f - so);
                da_s[buf][2 * U + ul][bl] = ag;
                da_s[buf][3 * U + ul][bl] = ao;
                dg[2 * H] = ag;
                dg[3 * H] = ao;
            }
            if (t > 0) load_stash(t - 1);  // prefetch: lands while the matvec and the exchange run
        }
        if (t == 0) break;
        if (__syncthreads_or(my_abort)) break;  // da_s[buf] complete (double buffered: one barrier per step)

        // ---- 2. partial[b][k] over the own rows; publish ------------------------------------------------
        const uint32_t par = step_parity(s);
        uint32_t* slot = CL ? nullptr : ring + (size_t)buf * kSlotWords;
        if (tid < NACT) {
            float acc[8][KC];
#pragma unroll
            for (int b = 0; b < 8; ++b)
#pragma unroll
                for (int c = 0; c < KC; ++c) acc[b][c] = 0.0f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 d0 = *reinterpret_cast<const float4*>(&da_s[buf][r][0]);
                const float4 d1 = *reinterpret_cast<const float4*>(&da_s[buf][r][4]);
                const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int b = 0; b < 8; ++b)
#pragma unroll
                    for (int c = 0; c < KC; ++c) acc[b][c] = fmaf(d[b], w[r][c], acc[b][c]);
            }
            // column k = KC*tid belongs to consumer k / U, unit k % U: word [consumer][b][producer = slice][unit]
            const int k0 = KC * tid;
            if (CL) {
                const uint32_t inbox = map_to_cta(smem_u32(&rx_s[((size_t)buf * kGroup * NS + slice) * U + (k0 % U)]),
                                                  (uint32_t)(k0 / U));
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (b < nvalid) {
                        const uint32_t dst = inbox + (uint32_t)(b * NS * U) * 4u;
                        if (KC == 2)
                            st_peer_v2(dst, flagged(acc[b][0], par), flagged(acc[b][KC - 1], par));
                        else
                            st_peer_b32(dst, flagged(acc[b][0], par));
                    }
                }
            } else {
                uint32_t* out = slot + ((size_t)(k0 / U) * kGroup * NS + slice) * U + (k0 % U);
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (b < nvalid) {
                        uint32_t* dst = out + (size_t)b * NS * U;
                        if (KC == 2) {
                            const uint32_t x = (__float_as_uint(acc[b][0]) & ~1u) | par;
                            const uint32_t y = (__float_as_uint(acc[b][KC - 1]) & ~1u) | par;
                            asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(dst), "r"(x), "r"(y) : "memory");
                        } else {
                            st_flagged(dst, acc[b][0], par);
                        }
                    }
                }
            }
        }

        // ---- 3. reduce-scatter: sum the producers' partials for the own units ---------------------------
        {
            // this CTA's block of the slot: [b][producer][U] words = per video NS*VPC uint4, contiguous
            uint4 v[4];
            // load i of lane l:  RG=1: video 2w + (i>>1), vector (i&1)*32 + l  ->  producer vec/2, unit quad l&1
            //                    RG=2: video w,            vector i*32 + l      ->  producer vec/4, unit quad l&3
            auto vec_b = [&](int i) { return RG == 1 ? 2 * warp + (i >> 1) : warp; };
            auto vec_id = [&](int i) { return RG == 1 ? (i & 1) * 32 + lane : i * 32 + lane; };
            auto vec_valid = [&](int i) { return vec_id(i) < NS * VPC && vec_b(i) < nvalid; };
            if (CL) {
                const uint32_t inbox = smem_u32(&rx_s[(size_t)buf * kGroup * H]);
                if (!gather_flagged(
                        v, [&](int i) { return inbox + (uint32_t)(vec_b(i) * NS * VPC + vec_id(i)) * 16u; }, vec_valid,
                        par, p.status, t))
                    my_abort = 1;
            } else {
                const uint32_t* src = slot + (size_t)slice * kGroup * NS * U;
                if (!gather_flagged(
                        v, [&](int i) { return src + ((size_t)vec_b(i) * NS * VPC + vec_id(i)) * 4; }, vec_valid, par,
                        p.status, t))
                    my_abort = 1;
            }
            float f[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = vec_id(i) < NS * VPC && vec_b(i) < nvalid;
                f[i][0] = ok ? __uint_as_float(v[i].x) : 0.0f;
                f[i][1] = ok ? __uint_as_float(v[i].y) : 0.0f;
                f[i][2] = ok ? __uint_as_float(v[i].z) : 0.0f;
                f[i][3] = ok ? __uint_as_float(v[i].w) : 0.0f;
            }
            if (RG == 1) {
                // 8 values (video bit, j): in-thread sum over the two producer halves, then over lane bits 4,3,2
                // with a butterfly (value index bit 2 = video, bits 1..0 = j) and over lane bit 1 with a plain add
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
                // 4 values (j): in-thread sum over the four producer octets, butterfly over lane bits 4,3,
                // plain add over lane bit 2
                float r4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r4[j] = (f[0][j] + f[1][j]) + (f[2][j] + f[3][j]);
                butterfly_stage<4, 1, 4>(r4, (lane & 16) != 0, 16);
                butterfly_stage<2, 0, 4>(r4, (lane & 8) != 0, 8);
                dh_rec = r4[0] + __shfl_xor_sync(0xffffffffu, r4[0], 4);
            }
        }
    }
    if (CL) cg::this_cluster().sync();  // no CTA leaves while peers may still push into its inbox
}

// ---- host side ----------------------------------------------------------------------
struct WorkspaceLayout {
    size_t status_off, fwd_ring_off, bwd_ring_off, total;
};

// units per CTA (8 or 16): 16 at H = 512 so that B = 32 is one CTA per SM
int units_per_cta(int64_t H) { return H == 512 ? 2 * kUnits : kUnits; }

WorkspaceLayout layout(int64_t B, int64_t T, int64_t H) {
    (void)T;
    const size_t groups = (size_t)((B + kGroup - 1) / kGroup);
    WorkspaceLayout l;
    l.status_off = 0;
    l.fwd_ring_off = 4096;  // status block (4 words used)
    l.bwd_ring_off = l.fwd_ring_off + groups * 2 * kGroup * (size_t)H * sizeof(float);
    const size_t producers = (size_t)(H / units_per_cta(H));
    l.total = l.bwd_ring_off + groups * 2 * producers * kGroup * (size_t)H * sizeof(float);
    return l;
}

bool supported_hidden(int64_t H) { return H == 32 || H == 64 || H == 128 || H == 256 || H == 512; }

// exchange medium of the FP32-FMA kernels: "cluster" (DSMEM, H <= 256) or "l2" (global-memory ring, any H; default:
// with 8-byte peer stores the cluster flavour measured 2.86 us/step against 2.08 for the ring at H = 256 forward)
bool want_cluster() {
    const char* e = getenv("OPN_LSTM_EXCHANGE");
    return e && e[0] == 'c';
}

template <int KPT, int RG>
int run_fwd(const FwdParams& p, int64_t B, cudaStream_t s, bool cluster_ok) {
    constexpr int H = 32 * KPT;
    if constexpr (H / (kUnits * RG) <= 16) {
        if (cluster_ok && want_cluster()) {
            bool launched = false;
            const int rc = launch_cluster(lstm_fwd_kernel<KPT, RG, true>, p, kThreads * RG, H / (kUnits * RG),
                                          (size_t)2 * kGroup * H * sizeof(float), B, s, &launched);
            if (rc != OPN_OK || launched) return rc;
        }
    }
    return launch_ring(lstm_fwd_kernel<KPT, RG, false>, p, kThreads * RG, H / (kUnits * RG),
                       (size_t)2 * kGroup * H * sizeof(float), B, s, "lstm_fwd");
}

template <int KPT, int RG>
int run_bwd(const BwdParams& p, int64_t B, cudaStream_t s, bool cluster_ok) {
    constexpr int H = 32 * KPT;
    if constexpr (H / (kUnits * RG) <= 16) {
        if (cluster_ok && want_cluster()) {
            bool launched = false;
            const int rc = launch_cluster(lstm_bwd_kernel<KPT, RG, true>, p, kThreads * RG, H / (kUnits * RG), 0, B, s,
                                          &launched);
            if (rc != OPN_OK || launched) return rc;
        }
    }
    return launch_ring(lstm_bwd_kernel<KPT, RG, false>, p, kThreads * RG, H / (kUnits * RG), 0, B, s, "lstm_bwd");
}

}  // namespace
}  // namespace opn

namespace opn {
// batch-wide tcgen05 flavour (opn_lstm_tc.cu): groups of 128 videos, weights in shared memory, accumulators in TMEM
bool lstm_tc_wanted(int64_t B, int64_t H);
size_t lstm_tc_workspace_bytes(int64_t B, int64_t H);
int lstm_fwd_tc(const FwdParams& p, int64_t B, int64_t H, void* workspace, cudaStream_t s);
int lstm_bwd_tc(const BwdParams& p, int64_t B, int64_t H, void* workspace, cudaStream_t s);
// tensor-core flavour of the recurrence (opn_lstm_mma.cu)
bool lstm_mma_supported(int64_t H);
int lstm_fwd_mma(const FwdParams& p, int64_t B, int64_t H, cudaStream_t s);
int lstm_bwd_mma(const BwdParams& p, int64_t B, int64_t H, cudaStream_t s);
}  // namespace opn

using namespace opn;

// per-step matvec engine: split-fp16 tensor-core MMAs where they exist (H = 256, 512), FP32 FMA otherwise or when
// OPN_LSTM_MATH=ffma asks for it
static bool want_mma(int64_t H) {
    const char* e = getenv("OPN_LSTM_MATH");
    if (e && e[0] == 'f') return false;
    return lstm_mma_supported(H);
}

extern "C" int64_t opn_lstm_workspace_bytes(int64_t B, int64_t T, int64_t H) {
    if (B <= 0 || T <= 0 || H <= 0) return 0;
    const size_t ring = layout(B, T, H).total, tcb = lstm_tc_workspace_bytes(B, H);
    return (int64_t)(ring > tcb ? ring : tcb);
}

extern "C" int opn_lstm_fwd(int64_t B, int64_t T, int64_t H, const float* xproj, const float* w_hh, float* hs,
                            float* gates, float* cells, void* workspace, int64_t workspace_bytes, void* stream) {
    OPN_CHECK_ARG(B > 0 && T > 0, "lstm_fwd: B and T must be positive (got %lld, %lld)", (long long)B, (long long)T);
    if (!supported_hidden(H)) {
        set_error("lstm_fwd: hidden size %lld unsupported (supported: 32, 64, 128, 256, 512)", (long long)H);
        return OPN_ERR_UNSUPPORTED;
    }
    OPN_CHECK_ARG(xproj && w_hh && hs && workspace, "lstm_fwd: null pointer");
    OPN_CHECK_ARG((gates == nullptr) == (cells == nullptr), "lstm_fwd: gates and cells must both be given or both NULL");
    const WorkspaceLayout l = layout(B, T, H);
    OPN_CHECK_ARG(workspace_bytes >= opn_lstm_workspace_bytes(B, T, H), "lstm_fwd: workspace too small (%lld < %lld)",
                  (long long)workspace_bytes, (long long)opn_lstm_workspace_bytes(B, T, H));
    cudaStream_t s = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    if (lstm_tc_wanted(B, H)) {
        FwdParams q;
        q.xproj = xproj, q.w_hh = w_hh, q.hs = hs, q.gates = gates, q.cells = cells;
        q.ring = nullptr;
        q.status = status_page_or(ws);
        q.B = (int)B, q.T = (int)T, q.group_offset = 0, q.n_slices = 0;
        return lstm_fwd_tc(q, B, H, ws, s);
    }
    OPN_CUDA(cudaMemsetAsync(ws + l.status_off, 0, l.bwd_ring_off - l.status_off, s));  // status + forward ring
    FwdParams p;
    p.xproj = xproj;
    p.w_hh = w_hh;
    p.hs = hs;
    p.gates = gates;
    p.cells = cells;
    p.ring = reinterpret_cast<uint32_t*>(ws + l.fwd_ring_off);
    p.status = status_page_or(ws + l.status_off);
    p.B = (int)B;
    p.T = (int)T;
    p.group_offset = 0;
    p.n_slices = 0;
    if (want_mma(H)) return lstm_fwd_mma(p, B, H, s);
    // cluster size = H / units per CTA must be <= 16: 8-unit CTAs up to H = 128, 16-unit CTAs at H = 256
    switch (H) {
        case 32: return run_fwd<1, 1>(p, B, s, true);
        case 64: return run_fwd<2, 1>(p, B, s, true);
        case 128: return run_fwd<4, 1>(p, B, s, true);
        case 256: return want_cluster() ? run_fwd<8, 2>(p, B, s, true) : run_fwd<8, 1>(p, B, s, false);
        default: return run_fwd<16, 2>(p, B, s, false);
    }
}

extern "C" int opn_lstm_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates,
                            const float* cells, const float* dh_out, float* dgates, void* workspace,
                            int64_t workspace_bytes, void* stream) {
    OPN_CHECK_ARG(B > 0 && T > 0, "lstm_bwd: B and T must be positive");
    if (!supported_hidden(H)) {
        set_error("lstm_bwd: hidden size %lld unsupported (supported: 32, 64, 128, 256, 512)", (long long)H);
        return OPN_ERR_UNSUPPORTED;
    }
    OPN_CHECK_ARG(w_hh && gates && cells && dh_out && dgates && workspace, "lstm_bwd: null pointer");
    const WorkspaceLayout l = layout(B, T, H);
    OPN_CHECK_ARG(workspace_bytes >= opn_lstm_workspace_bytes(B, T, H), "lstm_bwd: workspace too small");
    cudaStream_t s = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    if (lstm_tc_wanted(B, H)) {
        BwdParams q;
        q.w_hh = w_hh, q.gates = gates, q.cells = cells, q.dh_out = dh_out, q.dgates = dgates;
        q.ring = nullptr;
        q.status = status_page_or(ws);
        q.B = (int)B, q.T = (int)T, q.group_offset = 0, q.n_slices = 0;
        return lstm_bwd_tc(q, B, H, ws, s);
    }
    OPN_CUDA(cudaMemsetAsync(ws + l.status_off, 0, 4096, s));
    OPN_CUDA(cudaMemsetAsync(ws + l.bwd_ring_off, 0, l.total - l.bwd_ring_off, s));
    BwdParams p;
    p.w_hh = w_hh;
    p.gates = gates;
    p.cells = cells;
    p.dh_out = dh_out;
    p.dgates = dgates;
    p.ring = reinterpret_cast<uint32_t*>(ws + l.bwd_ring_off);
    p.status = status_page_or(ws + l.status_off);
    p.B = (int)B;
    p.T = (int)T;
    p.group_offset = 0;
    p.n_slices = 0;
    if (want_mma(H)) return lstm_bwd_mma(p, B, H, s);
    switch (H) {
        case 32: return run_bwd<1, 1>(p, B, s, true);
        case 64: return run_bwd<2, 1>(p, B, s, true);
        case 128: return run_bwd<4, 1>(p, B, s, true);
        case 256: return want_cluster() ? run_bwd<8, 2>(p, B, s, true) : run_bwd<8, 1>(p, B, s, false);
        default: return run_bwd<16, 2>(p, B, s, false);
    }
}

extern "C" int opn_lstm_batchwide(int64_t B, int64_t H) { return lstm_tc_wanted(B, H) ? 1 : 0; }

extern "C" int opn_lstm_status(const void* workspace, uint32_t* info) {
    OPN_CHECK_ARG(workspace != nullptr, "lstm_status: null workspace");
    uint32_t words[4] = {0, 0, 0, 0};
    OPN_CUDA(cudaDeviceSynchronize());
    OPN_CUDA(cudaMemcpy(words, workspace, sizeof(words), cudaMemcpyDeviceToHost));
    if (info) {
        info[0] = words[1];
        info[1] = words[2];
        info[2] = words[3];
    }
    if (words[0] != 0) {
        set_error("persistent LSTM kernel timed out (code %u, step %u, cta %u, thread %u)", words[0], words[1],
                  words[2], words[3]);
        return OPN_ERR_TIMEOUT;
    }
    return OPN_OK;
}
