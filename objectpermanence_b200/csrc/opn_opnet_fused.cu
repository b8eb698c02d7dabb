// OPNet forward as ONE persistent kernel: LSTM1 (90 -> 256), the who-to-track head (Linear 256 -> 15, softmax,
// probability-weighted box sum) and LSTM2 (6 -> 512) advance together, one frame per loop iteration
// (baselines/learned_models.py:36-47; the K = 90 input projection of LSTM1 stays a time-parallel contraction
// in front of the kernel, the Linear 512 -> 4 head one behind it).
//
// Why: each recurrence alone is bound by the latency of its inter-CTA exchange (~2000 clocks of a 3200-4100
// clock step during which the SM only polls, profiles/r01_lstm_mma_phases.log).  LSTM2 at frame t needs only
// frames_boxes[t], i.e. h1[t]; so while the CTAs wait for h2[t-1] of the other CTAs they run LSTM1 frame t+1 and
// the who-to-track head of frame t, and vice versa.  The loop time is that of LSTM2 alone; LSTM1, the head, the
// K = 6 input projection of LSTM2 and the [B,T,4*512] x-projection tensor (79 MB written and re-read) disappear.
//
// CTA (slice s of 32, batch group g of 8 videos), 256 threads:
//   * LSTM2: units [16s, 16s+16) -- exactly lstm_fwd_mma_kernel<512, 2>: W_hh2 slice as fp16 hi/lo A fragments in
//     registers (128 per thread), h2 exchanged through the L2 ring in B-fragment order.
//   * LSTM1: units [8s, 8s+8): the 32 x 256 slice of W_hh1 as fp16 hi/lo A fragments in SHARED memory (32 KB,
//     the registers are taken), 96 MMAs per frame spread over the 8 warps (m-tile x K quarter).
//   * head: every CTA of a group holds h1[t] of its 8 videos after the LSTM1 gather, so each computes the 15 logits
//     (W_pred as 16 x 256 A fragments in shared memory, 48 MMAs over the 8 warps), the softmax and the weighted
//     box sum redundantly (8 videos x 15 objects) -- no extra exchange; slice 0 writes logits / probs /
//     frames_boxes for the backward pass.
//   * pointwise: warps 0-3 own the LSTM1 cells while warps 4-7 run the head; all 8 warps own the LSTM2 cells and
//     form the K = 6 input projection of their gates from frames_boxes in shared memory.
// Outputs are the tensors the unfused forward produces (hs1, gates1, cells1, logits, probs, frames_boxes, hs2,
// gates2, cells2), so the backward pass is unchanged.
#include <stdlib.h>

#include "opn_mma_common.cuh"

namespace opn {

struct FusedFwdParams {
    const float* boxes;    // [B,T,15,6]
    const float* xproj1;   // [B,T,4*256]   boxes . W_ih1^T
    const float* w_hh1;    // [1024,256]
    const float* w_pred;   // [15,256]
    const float* w_ih2;    // [2048,6]
    const float* w_hh2;    // [2048,512]
    float *hs1, *gates1, *cells1;   // gates1 / cells1 NULL in inference
    float *logits, *probs, *fb;     // [B,15,T], [B,T,15], [B,T,6]
    float *hs2, *gates2, *cells2;   // gates2 / cells2 NULL in inference
    uint32_t *ring1, *ring2;        // [groups][2][8*256], [groups][2][8*512] flagged words, fragment order
    const unsigned int* fbx;        // EXT mode: [groups][T][8][8] frames_boxes left by the LSTM1 / head producer kernel (ready bit in the LSB)
    const unsigned int* flags;      // unused (reserved)
    unsigned int* status;
    int B, T;
    int group_offset, n_slices;
};

namespace {

constexpr int H1 = 256, H2 = 512, NOBJ = 15, NFEAT = 6, BOXROW = NOBJ * NFEAT;
constexpr int NT = 256, NW = 8;
constexpr int U1 = 8, U2 = 16;
constexpr int KS1 = H1 / 16, KS2 = H2 / 16;
// Every exchange tile exists in kReplicas copies in L2 (a producer stores its vector once per copy, a consumer
// reads copy slice % kReplicas).  Measured: 4 copies change nothing (3.98 vs 3.86 us/frame) -- the sweeps are bound
// by the bytes every SM pulls out of L2, not by the fan-out per line -- so the default is 1.
#ifndef OPN_RING_REPLICAS
#define OPN_RING_REPLICAS 1
#endif
constexpr int kReplicas = OPN_RING_REPLICAS;


// shared memory carve-up (bytes)
constexpr int OFF_BFRAG2 = 0;                                   // uint4 [KS2*32]
constexpr int OFF_D2 = OFF_BFRAG2 + KS2 * 32 * 16;              // float [2][64][8]
constexpr int OFF_A1 = OFF_D2 + 2 * 64 * 8 * 4;                 // uint4 [2 mt][KS1][2 hi/lo][32]
constexpr int OFF_AP = OFF_A1 + 2 * KS1 * 2 * 32 * 16;          // uint4 [KS1][2][32]
constexpr int OFF_BFRAG1 = OFF_AP + KS1 * 2 * 32 * 16;          // uint4 [KS1*32]
constexpr int OFF_D1 = OFF_BFRAG1 + KS1 * 32 * 16;              // float [4][32][8]
constexpr int OFF_DL = OFF_D1 + 4 * 32 * 8 * 4;                 // float [8][16][8]
constexpr int OFF_BOX = OFF_DL + 8 * 16 * 8 * 4;                // float [2][8][96]
constexpr int OFF_WIH2 = OFF_BOX + 2 * 8 * 96 * 4;              // float [64][8]
constexpr int OFF_FB = OFF_WIH2 + 64 * 8 * 4;                   // float [8][8]
constexpr int OFF_PROBS = OFF_FB + 8 * 8 * 4;                   // float [8][16]
constexpr int OFF_RED = OFF_PROBS + 8 * 16 * 4;                 // float [8]
constexpr int OFF_LAND1 = OFF_RED + 64;                         // uint4 [2][KS1*32]  landing tiles of the h1 sweeps
constexpr int OFF_LAND2 = OFF_LAND1 + 2 * KS1 * 32 * 16;        // uint4 [2][KS2*32]  landing tiles of the h2 sweeps
constexpr int OFF_BARS = OFF_LAND2 + 2 * KS2 * 32 * 16;         // uint64 [2 tiles][2 buffers]
constexpr int SMEM_BYTES = OFF_BARS + 32;

// The 32 CTAs of a batch group all need the same tile.  CTAs are launched as thread-block clusters of kCluster
// consecutive slices; a sweep is kCluster TMA bulk copies, each CTA fetching 1/kCluster of the tile from L2 ONCE and
// multicasting it into the landing tile of every CTA of its cluster: the bytes pulled out of L2 per sweep drop
// kCluster-fold (2 MB per frame for the h2 tiles without it, the burst that made a sweep take ~2000 clocks).
constexpr int kCluster = 4;   // 8 does not fit: only 15 clusters of 8 are co-resident on the 148 SMs (GPC granularity), 32 of 4 are

__device__ __forceinline__ void mma_f16(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// A fragment (hi, lo) of a row-major matrix W (row stride ld): rows (ra, rb) = the lane's two rows (g, g+8 of the
// m-tile; null = zero row), columns k0 = 16*ks + 2*tq .. ; values scaled by `scale`.
__device__ __forceinline__ void make_a_frag(const float* ra, const float* rb, int k0, float scale, uint4& hi, uint4& lo) {
    const float2 z = make_float2(0.f, 0.f);
    const float2 a0 = ra ? __ldg(reinterpret_cast<const float2*>(ra + k0)) : z;
    const float2 a1 = rb ? __ldg(reinterpret_cast<const float2*>(rb + k0)) : z;
    const float2 a2 = ra ? __ldg(reinterpret_cast<const float2*>(ra + k0 + 8)) : z;
    const float2 a3 = rb ? __ldg(reinterpret_cast<const float2*>(rb + k0 + 8)) : z;
    split2(a0.x * scale, a0.y * scale, hi.x, lo.x);
    split2(a1.x * scale, a1.y * scale, hi.y, lo.y);
    split2(a2.x * scale, a2.y * scale, hi.z, lo.z);
    split2(a3.x * scale, a3.y * scale, hi.w, lo.w);
}
__device__ __forceinline__ float a_frag_max(const float* ra, const float* rb, int k0) {
    float m = 0.f;
    if (ra) {
        const float2 a0 = __ldg(reinterpret_cast<const float2*>(ra + k0));
        const float2 a2 = __ldg(reinterpret_cast<const float2*>(ra + k0 + 8));
        m = fmaxf(fmaxf(fabsf(a0.x), fabsf(a0.y)), fmaxf(fabsf(a2.x), fabsf(a2.y)));
    }
    if (rb) {
        const float2 a1 = __ldg(reinterpret_cast<const float2*>(rb + k0));
        const float2 a3 = __ldg(reinterpret_cast<const float2*>(rb + k0 + 8));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(a1.x), fabsf(a1.y)), fmaxf(fabsf(a3.x), fabsf(a3.y))));
    }
    return m;
}

__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                   uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}

// Landing tile of one exchange: two buffers + two mbarriers, used by alternate sweeps (a CTA may already receive
// the pieces of sweep k+1 from faster peers while it still checks sweep k).
struct SweepState {
    uint32_t count;   // sweeps issued so far
    bool in_flight;   // the sweep `count - 1` has been issued and not yet consumed
};

// one sweep: every CTA of the cluster arms its barrier for the whole tile and multicasts its 1/kCluster piece
template <int TILE_BYTES>
__device__ __forceinline__ void issue_sweep(SweepState& st, uint64_t* bars, unsigned char* land, const uint32_t* src,
                                            uint32_t rank) {
    constexpr int PIECE = TILE_BYTES / kCluster;
    const uint32_t buf = st.count & 1u;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[buf], TILE_BYTES);
        bulk_g2s_multicast(land + buf * TILE_BYTES + rank * PIECE, reinterpret_cast<const unsigned char*>(src) + rank * PIECE,
                           PIECE, &bars[buf], (uint16_t)((1u << kCluster) - 1u));
    }
    st.count++;
    st.in_flight = true;
}

// Wait for the sweep in flight (or issue one), check the ready bits of this thread's NV fragment vectors in the
// landing tile, split them to fp16 fragments; repeat with fresh sweeps while any word of the tile is stale.  All
// CTAs of a cluster see the same bytes per sweep, hence take the same decision and stay in step.  The sweeps are
// asynchronous: the caller issues the first one early and computes something else while the tile streams in.
// Ends with a CTA barrier (fragments visible to all warps).  Returns the number of sweeps, 0 = time-out / abort.
template <int NV>
__device__ __forceinline__ int tma_gather(SweepState& st, uint64_t* bars, unsigned char* land, uint4* bfrag,
                                          const uint32_t* src, uint32_t rank, uint32_t par, int nvalid,
                                          unsigned int* status, int t) {
    constexpr int TILE_BYTES = NV * NT * 16;
    const int tid = threadIdx.x;
    const long long t0 = clock64();
    for (int sweeps = 1;; ++sweeps) {
        if (!st.in_flight) issue_sweep<TILE_BYTES>(st, bars, land, src, rank);
        st.in_flight = false;
        const uint32_t k = st.count - 1u, buf = k & 1u, phase = (k >> 1) & 1u;
        int bad = 0;
        unsigned spins = 0;
        while (!mbar_try_wait(&bars[buf], phase)) {
            if ((++spins & 1023u) == 0 && poll_expired(t0, status, t)) {
                bad = 1;
                break;
            }
        }
        const uint4* tile = reinterpret_cast<const uint4*>(land + buf * TILE_BYTES);
        int stale = 0;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int idx = tid + NT * q;
            if (((idx & 31) >> 2) < nvalid) {
                const uint4 v = tile[idx];
                if (!ready4(v, par)) {
                    stale = 1;
                } else {
                    uint4 f;
                    split2(__uint_as_float(v.x), __uint_as_float(v.y), f.x, f.z);
                    split2(__uint_as_float(v.z), __uint_as_float(v.w), f.y, f.w);
                    bfrag[idx] = f;
                }
            }
        }
        if (!__syncthreads_or(stale | bad)) return sweeps;
        // slow path bookkeeping (another CTA failed / this wait expired) costs an L2 round trip: every 32nd sweep
        if (bad || (sweeps & 31) == 0) {
            const int give_up = bad || poll_expired(t0, status, t);
            if (__syncthreads_or(give_up)) return 0;
        }
    }
}

// SINGLE: the 1e-2 arithmetic mode (opn_set_precision): the hi.hi product alone, a third of the MMAs
// EXT: LSTM1 and the who-to-track head run in the producer kernel of opn_opnet_l1head.cu on the SMs this launch leaves idle;
//      this kernel is then the LSTM2 loop alone and takes frames_boxes[t] from the producer (release flag per frame and slice)
template <bool SINGLE, bool EXT>
__global__ void __launch_bounds__(NT, 1) opnet_fwd_fused_kernel(const FusedFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint4* bfrag2_s = reinterpret_cast<uint4*>(smem + OFF_BFRAG2);
    float* d2_s = reinterpret_cast<float*>(smem + OFF_D2);        // [kp][lr][8]
    uint4* a1_s = reinterpret_cast<uint4*>(smem + OFF_A1);
    uint4* ap_s = reinterpret_cast<uint4*>(smem + OFF_AP);
    uint4* bfrag1_s = reinterpret_cast<uint4*>(smem + OFF_BFRAG1);
    float* d1_s = reinterpret_cast<float*>(smem + OFF_D1);        // [kp][lr][8]
    float* dl_s = reinterpret_cast<float*>(smem + OFF_DL);        // [warp][row][8]
    float* box_s = reinterpret_cast<float*>(smem + OFF_BOX);      // [2][8][96]
    float* wih2_s = reinterpret_cast<float*>(smem + OFF_WIH2);    // [lr][8]
    float* fb_s = reinterpret_cast<float*>(smem + OFF_FB);        // [b][8]
    float* probs_s = reinterpret_cast<float*>(smem + OFF_PROBS);  // [b][16]
    float* red_s = reinterpret_cast<float*>(smem + OFF_RED);
    unsigned char* land1_s = smem + OFF_LAND1;
    unsigned char* land2_s = smem + OFF_LAND2;
    uint64_t* bars1 = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
    uint64_t* bars2 = bars1 + 2;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int slice = blockIdx.x % p.n_slices;
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    // rings: [group][replica][slot][tile]; ringX = the copy this CTA reads, ringX_w = copy 0 (stores go to every copy)
    constexpr size_t kCopy1 = 2 * kGroup * H1, kCopy2 = 2 * kGroup * H2;
    uint32_t* ring1_w = p.ring1 + (size_t)group * (kReplicas * kCopy1);
    uint32_t* ring2_w = p.ring2 + (size_t)group * (kReplicas * kCopy2);
    uint32_t* ring1 = ring1_w + (size_t)(slice % kReplicas) * kCopy1;
    uint32_t* ring2 = ring2_w + (size_t)(slice % kReplicas) * kCopy2;
    const int u0_1 = slice * U1, u0_2 = slice * U2;

    if (tid == 0) {
        mbar_init(&bars1[0], 1);
        mbar_init(&bars1[1], 1);
        mbar_init(&bars2[0], 1);
        mbar_init(&bars2[1], 1);
        mbar_fence_init();
    }
    const uint32_t crank = cg::this_cluster().block_rank();
    for (int i = tid; i < KS2 * 32; i += NT) bfrag2_s[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < KS1 * 32; i += NT) bfrag1_s[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * 8 * 96; i += NT) box_s[i] = 0.0f;
    for (int i = tid; i < 64; i += NT) fb_s[i] = 0.0f;

    // ---- LSTM2 weights: A fragments in registers (warp = m-tile x K half), as lstm_fwd_mma_kernel<512, 2> ------
    const int mt2 = warp % 4, kp2 = warp / 4;
    const int lr2_a = mt2 * 16 + g, lr2_b = lr2_a + 8;   // local row = unit*4 + gate
    const float* w2a = p.w_hh2 + (size_t)((lr2_a & 3) * H2 + u0_2 + (lr2_a >> 2)) * H2;
    const float* w2b = p.w_hh2 + (size_t)((lr2_b & 3) * H2 + u0_2 + (lr2_b >> 2)) * H2;
    constexpr int KPW2 = KS2 / 2;
    float wmax = 0.0f;
#pragma unroll 4
    for (int j = 0; j < KPW2; ++j) wmax = fmaxf(wmax, a_frag_max(w2a, w2b, 16 * (kp2 * KPW2 + j) + 2 * tq));
    float wscale2, winv2;
    weight_scale<NW>(wmax, red_s, wscale2, winv2);
    uint32_t ahi[KPW2][4], alo[KPW2][4];
#pragma unroll
    for (int j = 0; j < KPW2; ++j) {
        uint4 hi, lo;
        make_a_frag(w2a, w2b, 16 * (kp2 * KPW2 + j) + 2 * tq, wscale2, hi, lo);
        ahi[j][0] = hi.x, ahi[j][1] = hi.y, ahi[j][2] = hi.z, ahi[j][3] = hi.w;
        alo[j][0] = lo.x, alo[j][1] = lo.y, alo[j][2] = lo.z, alo[j][3] = lo.w;
    }

    // ---- LSTM1 weights: 2 m-tiles x 16 k-steps of A fragments in shared memory; 1024 fragments, 4 per thread ---
    float winv1 = 0.0f, winvp = 0.0f;
    if constexpr (!EXT) {
        auto rows_of = [&](int mt, int gg, const float*& ra, const float*& rb) {
            const int la = mt * 16 + gg, lb = la + 8;
            ra = p.w_hh1 + (size_t)((la & 3) * H1 + u0_1 + (la >> 2)) * H1;
            rb = p.w_hh1 + (size_t)((lb & 3) * H1 + u0_1 + (lb >> 2)) * H1;
        };
        float m = 0.0f;
        for (int e = tid; e < 2 * KS1 * 32; e += NT) {
            const int l = e & 31, ks = (e >> 5) % KS1, mt = e / (32 * KS1);
            const float *ra, *rb;
            rows_of(mt, l >> 2, ra, rb);
            m = fmaxf(m, a_frag_max(ra, rb, 16 * ks + 2 * (l & 3)));
        }
        float wscale1;
        weight_scale<NW>(m, red_s, wscale1, winv1);
        for (int e = tid; e < 2 * KS1 * 32; e += NT) {
            const int l = e & 31, ks = (e >> 5) % KS1, mt = e / (32 * KS1);
            const float *ra, *rb;
            rows_of(mt, l >> 2, ra, rb);
            uint4 hi, lo;
            make_a_frag(ra, rb, 16 * ks + 2 * (l & 3), wscale1, hi, lo);
            a1_s[((mt * KS1 + ks) * 2 + 0) * 32 + l] = hi;
            a1_s[((mt * KS1 + ks) * 2 + 1) * 32 + l] = lo;
        }
        // ---- W_pred: one m-tile (row 15 = zero) x 16 k-steps, 512 fragments, 2 per thread -----------------------
        m = 0.0f;
        for (int e = tid; e < KS1 * 32; e += NT) {
            const int l = e & 31, ks = e >> 5, gg = l >> 2;
            const float* ra = p.w_pred + (size_t)gg * H1;
            const float* rb = (gg + 8 < NOBJ) ? p.w_pred + (size_t)(gg + 8) * H1 : nullptr;
            m = fmaxf(m, a_frag_max(ra, rb, 16 * ks + 2 * (l & 3)));
        }
        float wscalep;
        weight_scale<NW>(m, red_s, wscalep, winvp);
        for (int e = tid; e < KS1 * 32; e += NT) {
            const int l = e & 31, ks = e >> 5, gg = l >> 2;
            const float* ra = p.w_pred + (size_t)gg * H1;
            const float* rb = (gg + 8 < NOBJ) ? p.w_pred + (size_t)(gg + 8) * H1 : nullptr;
            uint4 hi, lo;
            make_a_frag(ra, rb, 16 * ks + 2 * (l & 3), wscalep, hi, lo);
            ap_s[(ks * 2 + 0) * 32 + l] = hi;
            ap_s[(ks * 2 + 1) * 32 + l] = lo;
        }
    }
    {
        // ---- W_ih2 rows of the CTA's LSTM2 cells, local row order ----------------------------------------------
        for (int e = tid; e < 64 * 8; e += NT) {
            const int lr = e >> 3, f = e & 7;
            wih2_s[e] = (f < NFEAT) ? __ldg(p.w_ih2 + (size_t)((lr & 3) * H2 + u0_2 + (lr >> 2)) * NFEAT + f) : 0.0f;
        }
    }

    // ---- cell ownership -----------------------------------------------------------------------------------------
    // LSTM2 (all warps): lane = (video : 2 | j : 2 | gh), unit = 2*(warp>>1) + (j&1) + 8*(j>>1)  (vector publish)
    const int gh = lane & 1;
    const int j2 = (lane >> 1) & 3;
    const int ul2 = 2 * (warp >> 1) + (j2 & 1) + 8 * (j2 >> 1);
    const int bl2 = 4 * (warp & 1) + (lane >> 3);
    const bool leader2 = (lane & 7) == 0;
    const int uu2 = u0_2 + ul2;
    const bool valid2 = b0 + bl2 < p.B;
    const size_t row2 = (size_t)(valid2 ? b0 + bl2 : 0) * T;
    const int pub2 = frag_word(bl2, uu2);
    // LSTM1 (warps 0-3): lane = (video : 3 | unit parity | gh), unit = 2*warp + parity
    const int ul1 = 2 * (warp & 3) + ((lane >> 1) & 1);
    const int bl1 = lane >> 2;
    const bool leader1 = (lane & 3) == 0;
    const int uu1 = u0_1 + ul1;
    const bool valid1 = b0 + bl1 < p.B;
    const size_t row1 = (size_t)(valid1 ? b0 + bl1 : 0) * T;
    const int pub1 = frag_word(bl1, uu1);
    const float* xp_ptr = p.xproj1 + row1 * (4 * H1) + (size_t)(2 * gh) * H1 + uu1;
    // head (warps 4-7): video hb, object ho
    const int hb = (tid - 128) >> 4, ho = tid & 15;
    const bool validh = warp >= 4 && b0 + hb < p.B;
    const float s0 = gh ? 1.0f : 0.5f;

    float c1 = 0.0f, c2 = 0.0f;
    float xp0 = 0.f, xp1 = 0.f;
    if (!EXT && warp < 4 && valid1) {
        xp0 = __ldg(xp_ptr);
        xp1 = __ldg(xp_ptr + H1);
    }
    // boxes rows of the 8 videos, 720 floats per frame: 3 per thread, held in registers one frame ahead
    float bx[3];
    auto load_boxes = [&](int t) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int idx = tid + NT * q;
            const int b = idx / BOXROW, e = idx % BOXROW;
            bx[q] = (idx < 8 * BOXROW && b < nvalid && t < T) ? __ldg(p.boxes + ((size_t)(b0 + b) * T + t) * BOXROW + e) : 0.0f;
        }
    };
    auto store_boxes = [&](int t) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int idx = tid + NT * q;
            if (idx < 8 * BOXROW) box_s[(t & 1) * (8 * 96) + (idx / BOXROW) * 96 + idx % BOXROW] = bx[q];
        }
    };
    load_boxes(0);
    SweepState sw1 = {0u, false}, sw2 = {0u, false};
    unsigned nsweeps1 = 0, nsweeps2 = 0;   // statistics: sweeps it took (>= one per frame), left in status[8], status[9]
    __syncthreads();
    cg::this_cluster().sync();   // every barrier of the cluster is initialised before the first multicast

    // iteration i: LSTM1 frame i (i < T), head of frame i-1 and LSTM2 frame i-1 (i >= 1)
    PH_DECL
    for (int i = 0; i <= T; ++i) {
        PH(0);  // publish h2 + stash stores of the previous iteration
        store_boxes(i);       // frame i (zeros past the end), read by the head in iteration i+1
        load_boxes(i + 1);
        // ================= LSTM1 frame i + head of frame i-1: both need h1[i-1] ===================================
        if (i >= 1 && !EXT) {
            // the sweep for h1[i-1] was issued behind the LSTM2 MMAs of the previous iteration (or is issued now)
            const int n = tma_gather<KS1 * 32 / NT>(sw1, bars1, land1_s, bfrag1_s,
                                                    ring1 + (size_t)((i - 1) & 1) * (kGroup * H1), crank,
                                                    step_parity(i - 1), nvalid, p.status, i);
            if (n == 0) break;
            nsweeps1 += n;
        } else {
            __syncthreads();
        }
        PH(2);  // h1 tile
        if (i >= 1 && !EXT) {
            // LSTM1: warp = (m-tile warp&1, K quarter warp>>1), 4 k-steps;  head: k-steps 2*warp, 2*warp+1
            if (i < T) {
                const int mt1 = warp & 1, kq = warp >> 1;
                float dm[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ks = 4 * kq + j;
                    const uint4 b = bfrag1_s[ks * 32 + lane];
                    const uint4 ah = a1_s[((mt1 * KS1 + ks) * 2 + 0) * 32 + lane];
                    const uint4 al = a1_s[((mt1 * KS1 + ks) * 2 + 1) * 32 + lane];
                    mma_f16(dm, ah, b.x, b.y);
                    if constexpr (!SINGLE) {
                        mma_f16(ds, ah, b.z, b.w);
                        mma_f16(ds, al, b.x, b.y);
                    }
                }
                float* d = d1_s + (kq * 32 + mt1 * 16 + g) * 8 + 2 * tq;
                *reinterpret_cast<float2*>(d) = make_float2((dm[0] + ds[0]) * winv1, (dm[1] + ds[1]) * winv1);
                *reinterpret_cast<float2*>(d + 64) = make_float2((dm[2] + ds[2]) * winv1, (dm[3] + ds[3]) * winv1);
            }
            {
                float dm[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int ks = 2 * warp + j;
                    const uint4 b = bfrag1_s[ks * 32 + lane];
                    const uint4 ah = ap_s[(ks * 2 + 0) * 32 + lane];
                    const uint4 al = ap_s[(ks * 2 + 1) * 32 + lane];
                    mma_f16(dm, ah, b.x, b.y);
                    if constexpr (!SINGLE) {
                        mma_f16(ds, ah, b.z, b.w);
                        mma_f16(ds, al, b.x, b.y);
                    }
                }
                float* d = dl_s + (warp * 16 + g) * 8 + 2 * tq;
                *reinterpret_cast<float2*>(d) = make_float2((dm[0] + ds[0]) * winvp, (dm[1] + ds[1]) * winvp);
                *reinterpret_cast<float2*>(d + 64) = make_float2((dm[2] + ds[2]) * winvp, (dm[3] + ds[3]) * winvp);
            }
        }
        __syncthreads();
        // h2[i-2] was published a whole LSTM1 phase ago: stream it in while the cells / the head are computed.
        // (Issued one phase earlier, before the LSTM1 MMAs, 10 % of the sweeps came back stale: 1.04 ms against 1.00.
        // Gating the early issue on a canary -- warp 0 samples the last word of each producer with a polling load a
        // phase ahead -- removed the stale sweeps but ran at 1.35 ms: the extra loads on the lines being written cost
        // more than the wait they save.)
        if (i >= 2) issue_sweep<KS2 * 32 * 16>(sw2, bars2, land2_s, ring2 + (size_t)((i - 2) & 1) * (kGroup * H2), crank);
        PH(3);  // LSTM1 + head MMAs + barrier
        if (EXT && i >= 1 && tid < 64) {
            // frames_boxes[i-1] of this group from the producer kernel (it normally runs frames ahead): words with the ready
            // bit in the mantissa LSB, polled like the exchange tiles; consumed behind the barrier of the h2 gather
            const int vb = tid >> 3, f = tid & 7;
            float x = 0.0f;
            if (f < NFEAT && vb < nvalid) {
                const unsigned int* src = p.fbx + (((size_t)group * T + (i - 1)) * 8 + vb) * 8 + f;
                unsigned int w = ld_relaxed(src);
                if (!(w & 1u)) {
                    const long long t0 = clock64();
                    unsigned spins = 0;
                    while (!((w = ld_relaxed(src)) & 1u)) {
                        if ((++spins & 63u) == 0 && poll_expired(t0, p.status, i)) break;
                    }
                }
                x = __uint_as_float(w);
            }
            fb_s[tid] = x;
        }
        if (EXT) {
        } else if (warp < 4) {
            if (i < T) {
                // ---- LSTM1 pointwise, frame i ---------------------------------------------------------------------
                float a0 = xp0, a1 = xp1;
                if (i >= 1) {
                    const int lr0 = ul1 * 4 + 2 * gh;
#pragma unroll
                    for (int kq = 0; kq < 4; ++kq) {
                        a0 += d1_s[(kq * 32 + lr0) * 8 + bl1];
                        a1 += d1_s[(kq * 32 + lr0 + 1) * 8 + bl1];
                    }
                }
                const float act0 = fmaf(s0, tanh_sfu(s0 * a0), 1.0f - s0);
                const float act1 = fmaf(0.5f, tanh_sfu(0.5f * a1), 0.5f);
                const float oth0 = __shfl_xor_sync(0xffffffffu, act0, 1);
                const float oth1 = __shfl_xor_sync(0xffffffffu, act1, 1);
                const float gi = gh ? oth0 : act0, gf = gh ? oth1 : act1, gg = gh ? act0 : oth0, go = gh ? act1 : oth1;
                c1 = fmaf(gf, c1, gi * gg);
                const float hval = go * tanh_sfu(c1);
                const uint32_t w0 = flagged(hval, step_parity(i));
                const uint32_t w1 = __shfl_xor_sync(0xffffffffu, w0, 2);
                if (leader1 && valid1) {
#pragma unroll
                    for (int r = 0; r < kReplicas; ++r)
                        asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1,%2};" ::"l"(ring1_w + r * kCopy1 +
                                                                                       (size_t)(i & 1) * (kGroup * H1) + pub1),
                                     "r"(w0), "r"(w1)
                                     : "memory");
                }
                if (valid1) {
                    const size_t row = row1 + i;
                    if (gh == 0) {
                        p.hs1[row * H1 + uu1] = hval;
                        if (p.gates1) {
                            p.gates1[row * (4 * H1) + uu1] = gi;
                            p.gates1[row * (4 * H1) + H1 + uu1] = gf;
                        }
                    } else {
                        if (p.gates1) {
                            p.gates1[row * (4 * H1) + 2 * H1 + uu1] = gg;
                            p.gates1[row * (4 * H1) + 3 * H1 + uu1] = go;
                        }
                        if (p.cells1) p.cells1[row * H1 + uu1] = c1;
                    }
                    if (i + 1 < T) {
                        xp0 = __ldg(xp_ptr + (size_t)(i + 1) * (4 * H1));
                        xp1 = __ldg(xp_ptr + (size_t)(i + 1) * (4 * H1) + H1);
                    }
                }
            }
        } else if (i >= 1 && !EXT) {
            // ---- who-to-track head, frame t = i-1: thread = (video hb, object ho), 16 lanes per video ----------------
            const int t = i - 1;
            float logit = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) logit += dl_s[(w * 16 + ho) * 8 + hb];
            float mx = (ho < NOBJ) ? logit : -INFINITY;
#pragma unroll
            for (int m = 8; m > 0; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
            const float pe = (ho < NOBJ) ? expf(logit - mx) : 0.0f;
            float den = pe;
#pragma unroll
            for (int m = 8; m > 0; m >>= 1) den += __shfl_xor_sync(0xffffffffu, den, m);
            const float pr = pe * (1.0f / den);
            probs_s[hb * 16 + ho] = pr;
            __syncwarp();
            float fbv = 0.0f;
            if (ho < NFEAT) {
                const float* bxr = box_s + (t & 1) * (8 * 96) + hb * 96;
#pragma unroll
                for (int o = 0; o < NOBJ; ++o) fbv = fmaf(probs_s[hb * 16 + o], bxr[o * NFEAT + ho], fbv);
                fb_s[hb * 8 + ho] = fbv;
            }
            if (slice == 0 && validh) {
                const size_t bb = (size_t)(b0 + hb);
                if (ho < NOBJ) {
                    p.logits[(bb * NOBJ + ho) * T + t] = logit;
                    p.probs[(bb * T + t) * NOBJ + ho] = pr;
                }
                if (ho < NFEAT) p.fb[(bb * T + t) * NFEAT + ho] = fbv;
            }
        }
        PH(4);  // LSTM1 pointwise + publish
        if (i == 0) continue;

        // ================= LSTM2 frame t2 = i-1: needs h2[t2-1] and frames_boxes[t2] ==============================
        const int t2 = i - 1;
        if (t2 >= 1) {
            const int n = tma_gather<KS2 * 32 / NT>(sw2, bars2, land2_s, bfrag2_s,
                                                    ring2 + (size_t)((t2 - 1) & 1) * (kGroup * H2), crank,
                                                    step_parity(t2 - 1), nvalid, p.status, i);
            if (n == 0) break;
            nsweeps2 += n;
        } else {
            __syncthreads();   // publishes fb_s of this frame to all warps
        }
        PH(5);  // h2 tile
        if (t2 >= 1) {
            float dm0[4] = {0.f, 0.f, 0.f, 0.f}, dm1[4] = {0.f, 0.f, 0.f, 0.f};
            float ds0[4] = {0.f, 0.f, 0.f, 0.f}, ds1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < KPW2; ++j) {
                const uint4 b = bfrag2_s[(kp2 * KPW2 + j) * 32 + lane];
                if (j & 1)
                    opn::mma_f16(dm1, ahi[j], b.x, b.y);
                else
                    opn::mma_f16(dm0, ahi[j], b.x, b.y);
                if constexpr (!SINGLE) {
                    opn::mma_f16(ds0, ahi[j], b.z, b.w);
                    opn::mma_f16(ds1, alo[j], b.x, b.y);
                }
            }
            float* d = d2_s + (kp2 * 64 + lr2_a) * 8 + 2 * tq;
            *reinterpret_cast<float2*>(d) = make_float2(((dm0[0] + dm1[0]) + (ds0[0] + ds1[0])) * winv2,
                                                        ((dm0[1] + dm1[1]) + (ds0[1] + ds1[1])) * winv2);
            *reinterpret_cast<float2*>(d + 64) = make_float2(((dm0[2] + dm1[2]) + (ds0[2] + ds1[2])) * winv2,
                                                             ((dm0[3] + dm1[3]) + (ds0[3] + ds1[3])) * winv2);
        }
        __syncthreads();
        // h1[i] was published before the LSTM2 phase: stream it in while the LSTM2 cells are computed
        if (i < T && !EXT) issue_sweep<KS1 * 32 * 16>(sw1, bars1, land1_s, ring1 + (size_t)(i & 1) * (kGroup * H1), crank);
        PH(6);  // LSTM2 MMAs + barrier
        {
            // ---- LSTM2 pointwise, frame t2: K = 6 input projection from frames_boxes in shared memory -----------
            const int lr0 = ul2 * 4 + 2 * gh;
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int f = 0; f < NFEAT; ++f) {
                const float x = fb_s[bl2 * 8 + f];
                a0 = fmaf(wih2_s[lr0 * 8 + f], x, a0);
                a1 = fmaf(wih2_s[(lr0 + 1) * 8 + f], x, a1);
            }
            if (t2 >= 1) {
                a0 += d2_s[lr0 * 8 + bl2] + d2_s[(64 + lr0) * 8 + bl2];
                a1 += d2_s[(lr0 + 1) * 8 + bl2] + d2_s[(64 + lr0 + 1) * 8 + bl2];
            }
            const float act0 = fmaf(s0, tanh_sfu(s0 * a0), 1.0f - s0);
            const float act1 = fmaf(0.5f, tanh_sfu(0.5f * a1), 0.5f);
            const float oth0 = __shfl_xor_sync(0xffffffffu, act0, 1);
            const float oth1 = __shfl_xor_sync(0xffffffffu, act1, 1);
            const float gi = gh ? oth0 : act0, gf = gh ? oth1 : act1, gg = gh ? act0 : oth0, go = gh ? act1 : oth1;
            c2 = fmaf(gf, c2, gi * gg);
            const float hval = go * tanh_sfu(c2);
            const uint32_t w0 = flagged(hval, step_parity(t2));
            const uint32_t w1 = __shfl_xor_sync(0xffffffffu, w0, 2);
            const uint32_t w2 = __shfl_xor_sync(0xffffffffu, w0, 4);
            const uint32_t w3 = __shfl_xor_sync(0xffffffffu, w0, 6);
            if (leader2 && valid2 && t2 + 1 < T) {
#pragma unroll
                for (int r = 0; r < kReplicas; ++r)
                    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(ring2_w + r * kCopy2 +
                                                                                         (size_t)(t2 & 1) * (kGroup * H2) + pub2),
                                 "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                                 : "memory");
            }
            if (valid2) {
                const size_t row = row2 + t2;
                if (gh == 0) {
                    p.hs2[row * H2 + uu2] = hval;
                    if (p.gates2) {
                        p.gates2[row * (4 * H2) + uu2] = gi;
                        p.gates2[row * (4 * H2) + H2 + uu2] = gf;
                    }
                } else {
                    if (p.gates2) {
                        p.gates2[row * (4 * H2) + 2 * H2 + uu2] = gg;
                        p.gates2[row * (4 * H2) + 3 * H2 + uu2] = go;
                    }
                    if (p.cells2) p.cells2[row * H2 + uu2] = c2;
                }
            }
        }
        PH(7);  // LSTM2 pointwise
    }
    PH_STORE(p.status);
    if (blockIdx.x == 0 && tid == 0) {
        p.status[8] = nsweeps1;
        p.status[9] = nsweeps2;
    }
    cg::this_cluster().sync();   // no CTA leaves while a peer's multicast may still target its shared memory
}

struct FusedLayout {
    size_t status_off, ring1_off, ring2_off, flags_off, fbx_off, zero_bytes, total;
};
FusedLayout fused_layout(int64_t B, int64_t T) {
    const size_t groups = (size_t)((B + kGroup - 1) / kGroup);
    FusedLayout l;
    l.status_off = 0;
    l.ring1_off = 4096;
    l.ring2_off = l.ring1_off + groups * kReplicas * 2 * kGroup * H1 * sizeof(float);
    l.flags_off = l.ring2_off + groups * kReplicas * 2 * kGroup * H2 * sizeof(float);
    l.fbx_off = (l.flags_off + groups * (size_t)T * 4 * sizeof(unsigned int) + 255) & ~(size_t)255;
    l.total = l.fbx_off + groups * (size_t)T * 64 * sizeof(float);
    l.zero_bytes = l.total;        // status, rings and the ready bits of fbx are zeroed per call
    return l;
}

}  // namespace
}  // namespace opn

namespace opn {
int current_precision();
struct L1HeadParams {      // opn_opnet_l1head.cu
    const float* boxes;
    const float* xproj1;
    const float* w_hh1;
    const float* w_pred;
    float *hs1, *gates1, *cells1;
    float *logits, *probs, *fb;
    uint32_t* fbx;
    unsigned int* flags;
    uint32_t* ring1;
    unsigned int* status;
    int B, T;
    int group_offset, n_slices;
};
int launch_opnet_l1head(const L1HeadParams& p, int64_t B, bool single, cudaStream_t s, int group_begin, int group_end);
int preload_opnet_l1head();
}  // namespace opn
using namespace opn;

extern "C" int64_t opn_opnet_fwd_workspace_bytes(int64_t B, int64_t T) {
    if (B <= 0 || T <= 0) return 0;
    return (int64_t)fused_layout(B, T).total;
}

extern "C" int opn_opnet_fwd(int64_t B, int64_t T, int64_t H1_, int64_t H2_, const float* boxes, const float* xproj1,
                             const float* w_hh1, const float* w_pred, const float* w_ih2, const float* w_hh2, float* hs1,
                             float* gates1, float* cells1, float* logits_bpt, float* probs, float* frames_boxes, float* hs2,
                             float* gates2, float* cells2, void* workspace, int64_t workspace_bytes, void* stream) {
    OPN_CHECK_ARG(B > 0 && T > 0, "opnet_fwd: B and T must be positive");
    if (H1_ != H1 || H2_ != H2) {
        set_error("opnet_fwd: the fused forward exists for the shipped OPNet config (H1 = 256, H2 = 512), got %lld / %lld",
                  (long long)H1_, (long long)H2_);
        return OPN_ERR_UNSUPPORTED;
    }
    OPN_CHECK_ARG(boxes && xproj1 && w_hh1 && w_pred && w_ih2 && w_hh2 && hs1 && logits_bpt && probs && frames_boxes && hs2 &&
                      workspace,
                  "opnet_fwd: null pointer");
    OPN_CHECK_ARG((gates1 == nullptr) == (cells1 == nullptr) && (gates2 == nullptr) == (cells2 == nullptr) &&
                      (gates1 == nullptr) == (gates2 == nullptr),
                  "opnet_fwd: the stash tensors must all be given or all NULL");
    const FusedLayout l = fused_layout(B, T);
    OPN_CHECK_ARG(workspace_bytes >= (int64_t)l.total, "opnet_fwd: workspace too small (%lld < %lld)",
                  (long long)workspace_bytes, (long long)l.total);
    cudaStream_t s = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    OPN_CUDA(cudaMemsetAsync(ws, 0, l.zero_bytes, s));
    FusedFwdParams p;
    p.boxes = boxes;
    p.xproj1 = xproj1;
    p.w_hh1 = w_hh1;
    p.w_pred = w_pred;
    p.w_ih2 = w_ih2;
    p.w_hh2 = w_hh2;
    p.hs1 = hs1;
    p.gates1 = gates1;
    p.cells1 = cells1;
    p.logits = logits_bpt;
    p.probs = probs;
    p.fb = frames_boxes;
    p.hs2 = hs2;
    p.gates2 = gates2;
    p.cells2 = cells2;
    p.ring1 = reinterpret_cast<uint32_t*>(ws + l.ring1_off);
    p.ring2 = reinterpret_cast<uint32_t*>(ws + l.ring2_off);
    p.status = status_page_or(ws + l.status_off);
    p.B = (int)B;
    p.T = (int)T;
    p.group_offset = 0;
    p.n_slices = 32;
    p.fbx = reinterpret_cast<const unsigned int*>(ws + l.fbx_off);
    p.flags = reinterpret_cast<const unsigned int*>(ws + l.flags_off);
    const bool single = current_precision() == OPN_PRECISION_16BIT;
    SideStream* side = opnet_split_wanted(B) ? opnet_side_stream() : nullptr;
    if (side) {
        // consumer (this kernel, LSTM2 alone) on the caller's stream, producer (LSTM1 + head) on the library's side stream:
        // fork behind the memset and whatever produced the inputs, join so that the caller's stream sees hs1 / logits / ...
        // the producer's module must be resident before the consumer starts: a lazy load behind a running kernel waits for it
        // (first call: the consumer ran to its time-out)
        int rc = preload_opnet_l1head();
        if (rc != OPN_OK) return rc;
        std::lock_guard<std::mutex> turn(side->enqueue);
        OPN_CUDA(cudaEventRecord(side->fork, s));
        OPN_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
        const char* dbg = getenv("OPN_OPNET_SPLIT");
        const bool producer_only = dbg && dbg[0] == '2';      // timing of the producer alone (tools/split_fwd_debug.py)
        L1HeadParams q;
        q.boxes = boxes, q.xproj1 = xproj1, q.w_hh1 = w_hh1, q.w_pred = w_pred;
        q.hs1 = hs1, q.gates1 = gates1, q.cells1 = cells1, q.logits = logits_bpt, q.probs = probs, q.fb = frames_boxes;
        q.fbx = reinterpret_cast<uint32_t*>(ws + l.fbx_off);
        q.flags = reinterpret_cast<unsigned int*>(ws + l.flags_off);
        q.ring1 = p.ring1, q.status = p.status, q.B = (int)B, q.T = (int)T, q.group_offset = 0, q.n_slices = 5;
        // waves of as many batch groups as fit with both kernels co-resident (4 on 148 SMs); the consumer launches queue on the
        // caller's stream, the producer launches on the side stream: wave k+1 of either starts when its wave k ends
        const int groups = (int)((B + kGroup - 1) / kGroup), per_wave = opnet_split_groups_per_wave();
        for (int g0 = 0; g0 < groups; g0 += per_wave) {
            const int g1 = g0 + per_wave < groups ? g0 + per_wave : groups;
            if (!producer_only) {
                rc = single ? launch_ring(opnet_fwd_fused_kernel<true, true>, p, NT, 32, (size_t)SMEM_BYTES, B, s, "opnet_fwd", kCluster, g0, g1)
                            : launch_ring(opnet_fwd_fused_kernel<false, true>, p, NT, 32, (size_t)SMEM_BYTES, B, s, "opnet_fwd", kCluster, g0, g1);
                if (rc != OPN_OK) return rc;
            }
            rc = launch_opnet_l1head(q, B, single, side->stream, g0, g1);
            if (rc != OPN_OK) return rc;
        }
        OPN_CUDA(cudaEventRecord(side->join, side->stream));
        OPN_CUDA(cudaStreamWaitEvent(s, side->join, 0));
        return OPN_OK;
    }
    if (single) return launch_ring(opnet_fwd_fused_kernel<true, false>, p, NT, 32, (size_t)SMEM_BYTES, B, s, "opnet_fwd", kCluster);
    return launch_ring(opnet_fwd_fused_kernel<false, false>, p, NT, 32, (size_t)SMEM_BYTES, B, s, "opnet_fwd", kCluster);
}
