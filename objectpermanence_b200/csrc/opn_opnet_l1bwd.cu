// OPNet backward of the who-to-track head and of LSTM1 (reverse recurrence) as a kernel on the SMs the fused backward leaves
// idle -- the mirror image of opn_opnet_l1head.cu (autograd backward of baselines/learned_models.py:36-43).
//
// opn_opnet_fused_bwd.cu runs the LSTM2 reverse recurrence, the head backward and the LSTM1 reverse recurrence on the same 128
// CTAs; the head / LSTM1 work of a frame (3,500 of 7,000 clocks: gather and sum of the d frames_boxes shares, head, LSTM1
// gather, cells, 64 MMAs, publish) is longer than the LSTM2 exchange it hides.  In EXT mode that kernel is the LSTM2 loop
// alone and leaves its 32 shares of d frames_boxes[t] per frame in a flagged buffer; this kernel consumes them on the 20
// idle SMs, 5 CTAs per batch group of 8 videos:
//   * head CTA: sums the 32 shares, head backward (softmax Jacobian against the probability-weighted box sum) -> d logits
//     (output), then d h1[t] (head part) = W_pred^T d logits[t] for all 256 units, published with a ready bit.
//   * four unit CTAs, 64 hidden units each: d h1[t] = head part + recurrent part (sum of the four CTAs' partial products of
//     the previous step, exchanged through an L2 ring; warp = video, so a step needs one block barrier), cell backward ->
//     d gates1 (output), scaled fp16 hi / lo B fragments,
//     W_hh1^T product for the CTA's 256 gate rows: [256 columns x 256 rows] . [256 x 8 videos], A fragments hi plane in
//     REGISTERS (128 per thread), lo plane in SHARED memory (128 KB); partial d h1[t-1] published per consumer.
// Outputs: d_gates1 [B,T,1024], d_logits [B,T,15] -- as the fused kernel leaves them for the weight-gradient kernel.
#include <stdlib.h>

#include "opn_mma_common.cuh"

namespace opn {

struct L1BwdParams {
    const float *boxes, *probs;          // [B,T,15,6], [B,T,15]
    const float *w_hh1, *w_pred;         // [1024,256], [15,256]
    const float *gates1, *cells1;        // forward stash
    float *dgates1, *dl;                 // [B,T,1024], [B,T,15]
    const uint32_t* dfbx;                // [groups][T][32 producers][64]: shares of d frames_boxes from the LSTM2 kernel (ready bit)
    uint32_t* dhx;                       // [groups][T][8 videos][256]: head part of d h1 (ready bit), head CTA -> unit CTAs
    uint32_t* ring;                      // [groups][2 slots][4 consumers][4 producers][8 videos][64 units] partial products
    unsigned int* status;
    int B, T;
    int group_offset, n_slices;
};

namespace {
namespace l1b {

constexpr int H1 = 256, NOBJ = 15, NFEAT = 6, BOXROW = NOBJ * NFEAT;
constexpr int NT = 256, NW = 8, NSL = 4, U = 64, KS = 16, MT = 16;
constexpr size_t kSlot = (size_t)NSL * NSL * kGroup * U;      // words per ring slot and batch group

// shared memory carve-up (bytes); the head CTA reuses the first region for its own tiles
constexpr int OFF_ALO = 0;                                  // uint4 [MT][KS][32]   lo plane of the W_hh1^T fragments, 128 KB
constexpr int OFF_DA = OFF_ALO + MT * KS * 32 * 16;         // uint4 [2][KS*32]     scaled d gates1 fragments, double buffered by step
constexpr int OFF_INV = OFF_DA + 2 * KS * 32 * 16;          // float [2][8]
constexpr int OFF_RED = OFF_INV + 64;                       // float [8]
constexpr int SMEM_BYTES = OFF_RED + 64;
// head CTA
constexpr int HOFF_WP = 0;                                  // float [15][256]
constexpr int HOFF_DFBT = HOFF_WP + NOBJ * H1 * 4;          // float [32][64]
constexpr int HOFF_DFB = HOFF_DFBT + 32 * 64 * 4;           // float [8][8]
constexpr int HOFF_DL = HOFF_DFB + 8 * 8 * 4;               // float [8][16]
constexpr int HOFF_BOX = HOFF_DL + 8 * 16 * 4;              // float [3][8][96]
constexpr int HOFF_PRB = HOFF_BOX + 3 * 8 * 96 * 4;         // float [3][8][16]
static_assert(HOFF_PRB + 3 * 8 * 16 * 4 <= SMEM_BYTES, "head CTA tiles fit");

__device__ __forceinline__ void mma4(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// one flagged word, polled until its ready bit is set (bounded); returns false on time-out
__device__ __forceinline__ bool poll_word(const uint32_t* src, uint32_t& w, unsigned int* status, int t) {
    w = ld_relaxed(src);
    if (w & 1u) return true;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!((w = ld_relaxed(src)) & 1u)) {
        if ((++spins & 63u) == 0 && poll_expired(t0, status, t)) return false;
    }
    return true;
}

// SINGLE: the 1e-2 arithmetic mode (the hi.hi product alone)
template <bool SINGLE>
__global__ void __launch_bounds__(NT, 1) opnet_l1bwd_kernel(const L1BwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int slice = blockIdx.x % p.n_slices;      // 0 .. 3: 64 hidden units each; 4: the head of the group
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    uint32_t* dhx = p.dhx + (size_t)group * T * (kGroup * H1);
    int my_abort = 0;

    if (slice == NSL) {
        // ============================================ head CTA =============================================================
        float* wp_s = reinterpret_cast<float*>(smem + HOFF_WP);
        float* dfbt_s = reinterpret_cast<float*>(smem + HOFF_DFBT);
        float* dfb_s = reinterpret_cast<float*>(smem + HOFF_DFB);
        float* dl_s = reinterpret_cast<float*>(smem + HOFF_DL);
        float* box_s = reinterpret_cast<float*>(smem + HOFF_BOX);
        float* prb_s = reinterpret_cast<float*>(smem + HOFF_PRB);
        for (int e = tid; e < NOBJ * H1; e += NT) wp_s[e] = __ldg(p.w_pred + e);
        for (int e = tid; e < 8 * 16; e += NT) dl_s[e] = 0.0f;
        const uint32_t* dfbx = p.dfbx + (size_t)group * T * (32 * 64);
        // probs / boxes of frame t -> buffer t % 3, asynchronously, a frame ahead (warps 4-7; warps 0-3 may lag a frame behind)
        auto prefetch_head = [&](int t) {
            if (warp >= 4) {
                if (t >= 0) {
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        const int e = (tid - 128) + 128 * q;      // 8 videos x (90 boxes + 15 probs) = 840 words
                        if (e < 8 * BOXROW) {
                            const int v = e / BOXROW, o = e % BOXROW;
                            if (v < nvalid) cp_async4(box_s + ((t % 3) * 8 + v) * 96 + o, p.boxes + ((size_t)(b0 + v) * T + t) * BOXROW + o);
                        } else if (e < 8 * BOXROW + 8 * NOBJ) {
                            const int v = (e - 8 * BOXROW) / NOBJ, o = (e - 8 * BOXROW) % NOBJ;
                            if (v < nvalid) cp_async4(prb_s + ((t % 3) * 8 + v) * 16 + o, p.probs + ((size_t)(b0 + v) * T + t) * NOBJ + o);
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        };
        prefetch_head(T - 1);
        const int hb = tid >> 4, ho = tid & 15;
        __syncthreads();
        for (int t = T - 1; t >= 0; --t) {
            {   // the 32 shares of d frames_boxes[t]: tile [32 producers][64 words], 48 used: vectors with (idx & 15) < 12
                const uint32_t* src = dfbx + (size_t)t * (32 * 64);
                auto vec_valid = [&](int q) { return ((tid + NT * q) & 15) < 12; };
                uint4 v[2];
                if (!gather_flagged(v, [&](int q) { return src + (size_t)(tid + NT * q) * 4; }, vec_valid, 1u, p.status, t)) my_abort = 1;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (vec_valid(q)) reinterpret_cast<uint4*>(dfbt_s)[tid + NT * q] = v[q];
            }
            prefetch_head(t - 1);
            if (__syncthreads_or(my_abort)) break;
            if (tid < 192) {   // 4 threads per value, 8 producers each, two shuffles
                const int out = tid >> 2, part = tid & 3;
                float sum = 0.0f;
#pragma unroll
                for (int pr = 0; pr < 8; ++pr) sum += dfbt_s[(part * 8 + pr) * 64 + out];
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                if (part == 0) dfb_s[(out & 7) * 8 + (out >> 3)] = sum;   // [video][feature]
            }
            if (warp >= 4) asm volatile("cp.async.wait_group 1;" ::: "memory");      // probs / boxes of frame t have landed
            __syncthreads();
            if (warp < 4) {
                // ---- head backward: thread = (video hb, object ho) ------------------------------------------------------
                const float* bx = box_s + ((t % 3) * 8 + hb) * 96 + ho * NFEAT;
                const float po = (ho < NOBJ && hb < nvalid) ? prb_s[((t % 3) * 8 + hb) * 16 + ho] : 0.0f;
                float dp = 0.0f;
                if (ho < NOBJ && hb < nvalid) {
#pragma unroll
                    for (int c = 0; c < NFEAT; ++c) dp = fmaf(dfb_s[hb * 8 + c], bx[c], dp);
                }
                float dot = po * dp;
#pragma unroll
                for (int m = 8; m > 0; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
                const float dlv = po * (dp - dot);
                dl_s[hb * 16 + ho] = (ho < NOBJ) ? dlv : 0.0f;
                if (ho < NOBJ && hb < nvalid) p.dl[((size_t)(b0 + hb) * T + t) * NOBJ + ho] = dlv;
            }
            __syncthreads();
            {   // head part of d h1[t]: thread = unit, the 8 videos; published with the ready bit in the mantissa LSB
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int o = 0; o < NOBJ; ++o) {
                    const float w = wp_s[o * H1 + tid];
#pragma unroll
                    for (int v = 0; v < 8; ++v) acc[v] = fmaf(dl_s[v * 16 + o], w, acc[v]);
                }
                uint32_t* dst = dhx + (size_t)t * (kGroup * H1) + tid;
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    if (v < nvalid) asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(dst + v * H1), "r"(flagged(acc[v], 1u)) : "memory");
            }
            // dl_s / dfb_s are rewritten behind the next iteration's first barrier; dfbt_s by its gather: order them behind this use
            __syncthreads();
        }
        return;
    }

    // ================================================ unit CTAs ================================================================
    uint4* alo_s = reinterpret_cast<uint4*>(smem + OFF_ALO);
    uint4* da_s = reinterpret_cast<uint4*>(smem + OFF_DA);
    float* inv_s = reinterpret_cast<float*>(smem + OFF_INV);
    float* red_s = reinterpret_cast<float*>(smem + OFF_RED);
    const int u0 = slice * U;
    uint32_t* ring = p.ring + (size_t)group * (2 * kSlot);

    for (int i = tid; i < 2 * KS * 32; i += NT) da_s[i] = make_uint4(0u, 0u, 0u, 0u);      // absent videos stay zero
    // ---- W_hh1^T slice: A[m = column k][kk = own local row lr = unit*4 + gate]; warp w owns columns 32w .. 32w+31 ------------
    auto wrow = [&](int lr) { return p.w_hh1 + (size_t)((lr & 3) * H1 + u0 + (lr >> 2)) * H1; };
    auto frag = [&](int mt, int ks, int l, float scale, uint4& hi, uint4& lo, float& m) {
        const int kc = mt * 16 + (l >> 2), lr = 16 * ks + 2 * (l & 3);
        const float v[8] = {__ldg(wrow(lr) + kc),         __ldg(wrow(lr + 1) + kc),     __ldg(wrow(lr) + kc + 8),     __ldg(wrow(lr + 1) + kc + 8),
                            __ldg(wrow(lr + 8) + kc),     __ldg(wrow(lr + 9) + kc),     __ldg(wrow(lr + 8) + kc + 8), __ldg(wrow(lr + 9) + kc + 8)};
#pragma unroll
        for (int q = 0; q < 8; ++q) m = fmaxf(m, fabsf(v[q]));
        split2(v[0] * scale, v[1] * scale, hi.x, lo.x);
        split2(v[2] * scale, v[3] * scale, hi.y, lo.y);
        split2(v[4] * scale, v[5] * scale, hi.z, lo.z);
        split2(v[6] * scale, v[7] * scale, hi.w, lo.w);
    };
    float wmax = 0.0f;
    {
        uint4 hi, lo;
        for (int m = 0; m < 2; ++m)
            for (int ks = 0; ks < KS; ++ks) frag(2 * warp + m, ks, lane, 1.0f, hi, lo, wmax);
    }
    float wscale, winv;
    weight_scale<NW>(wmax, red_s, wscale, winv);
    uint4 ahi[2][KS];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint4 lo;
            float dummy = 0.0f;
            frag(2 * warp + m, ks, lane, wscale, ahi[m][ks], lo, dummy);
            alo_s[((2 * warp + m) * KS + ks) * 32 + lane] = lo;
        }

    // ---- cells: warp = video, lane = units 2*lane, 2*lane + 1: the per-video maximum is a warp reduction, the four producers'
    // partial sums of a cell meet in its own thread, and a step needs ONE block barrier (fragments complete -> products)
    const int bl = warp, ul = 2 * lane;
    const bool valid = b0 + bl < p.B;
    const size_t row0 = (size_t)(valid ? b0 + bl : 0) * T;
    const int uu = u0 + ul;
    float dc[2] = {0.0f, 0.0f};
    __syncthreads();

    PH_DECL
    // step s: frame t = T - 1 - s
    for (int s = 0; s < T; ++s) {
        const int t = T - 1 - s;
        PH(0);
        // stash of frame t (the latency hides behind the polls below)
        float2 qi = make_float2(0.f, 0.f), qf = qi, qg = qi, qo = qi, qc = qi, qcp = qi;
        if (valid) {
            const float* gp = p.gates1 + (row0 + t) * (size_t)(4 * H1) + uu;
            qi = __ldg(reinterpret_cast<const float2*>(gp));
            qf = __ldg(reinterpret_cast<const float2*>(gp + H1));
            qg = __ldg(reinterpret_cast<const float2*>(gp + 2 * H1));
            qo = __ldg(reinterpret_cast<const float2*>(gp + 3 * H1));
            qc = __ldg(reinterpret_cast<const float2*>(p.cells1 + (row0 + t) * H1 + uu));
            if (t > 0) qcp = __ldg(reinterpret_cast<const float2*>(p.cells1 + (row0 + t - 1) * H1 + uu));
        }
        // d h1[t] of the two units: head part (head CTA) + recurrent part (the four producers' partial products of step s - 1)
        float dh[2] = {0.0f, 0.0f};
        if (valid) {
            const uint32_t* hsrc = dhx + (size_t)t * (kGroup * H1) + bl * H1 + uu;
            uint2 w[5];
            bool pending = true;
            const uint32_t par = s >= 1 ? step_parity(s - 1) : 0u;
            const uint32_t* rsrc = ring + (size_t)((s - 1) & 1) * kSlot + (size_t)slice * (NSL * kGroup * U) + (size_t)bl * U + ul;
            const long long t0 = clock64();
            unsigned spins = 0;
            while (pending) {
                asm volatile("ld.volatile.global.v2.b32 {%0,%1}, [%2];" : "=r"(w[4].x), "=r"(w[4].y) : "l"(hsrc) : "memory");
                pending = !((w[4].x & w[4].y) & 1u);
                if (s >= 1) {
#pragma unroll
                    for (int q = 0; q < NSL; ++q) {
                        asm volatile("ld.volatile.global.v2.b32 {%0,%1}, [%2];" : "=r"(w[q].x), "=r"(w[q].y) : "l"(rsrc + (size_t)q * (kGroup * U)) : "memory");
                        pending |= (((w[q].x ^ par) | (w[q].y ^ par)) & 1u) != 0u;
                    }
                }
                if (pending && (++spins & 63u) == 0 && poll_expired(t0, p.status, s)) {
                    my_abort = 1;
                    break;
                }
            }
            dh[0] = __uint_as_float(w[4].x), dh[1] = __uint_as_float(w[4].y);
            if (s >= 1) {
                dh[0] += (__uint_as_float(w[0].x) + __uint_as_float(w[1].x)) + (__uint_as_float(w[2].x) + __uint_as_float(w[3].x));
                dh[1] += (__uint_as_float(w[0].y) + __uint_as_float(w[1].y)) + (__uint_as_float(w[2].y) + __uint_as_float(w[3].y));
            }
        }
        PH(1);  // polls
        float dgv[2][4];
        {
            const float qi_[2] = {qi.x, qi.y}, qf_[2] = {qf.x, qf.y}, qg_[2] = {qg.x, qg.y}, qo_[2] = {qo.x, qo.y}, qc_[2] = {qc.x, qc.y},
                        qcp_[2] = {qcp.x, qcp.y};
            float mx = 0.0f;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float tc = tanh_sfu(qc_[j]);
                const float d_o = dh[j] * tc;
                const float d = fmaf(dh[j] * qo_[j], 1.0f - tc * tc, dc[j]);
                dc[j] = d * qf_[j];
                dgv[j][0] = d * qg_[j] * qi_[j] * (1.0f - qi_[j]);
                dgv[j][1] = d * qcp_[j] * qf_[j] * (1.0f - qf_[j]);
                dgv[j][2] = d * qi_[j] * (1.0f - qg_[j] * qg_[j]);
                dgv[j][3] = d_o * qo_[j] * (1.0f - qo_[j]);
                if (!valid) dgv[j][0] = dgv[j][1] = dgv[j][2] = dgv[j][3] = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; ++q) mx = fmaxf(mx, fabsf(dgv[j][q]));
            }
            if (t > 0) {
                // per-video power-of-two scale of the B operand: the warp holds all 256 rows of its video
                const unsigned mbits = __reduce_max_sync(0xffffffffu, __float_as_uint(mx) & 0x7fffffffu);
                float sc, inv;
                pow2_scale(__uint_as_float(mbits), 11, sc, inv);
                // rows lr = unit*4 + gate: the thread's 8 rows are words (lr, lr+1) of the B fragments
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int lr = (ul + j) * 4 + 2 * h2;
                        uint32_t hi, lo;
                        split2(dgv[j][2 * h2] * sc, dgv[j][2 * h2 + 1] * sc, hi, lo);
                        uint32_t* w = reinterpret_cast<uint32_t*>(da_s + (s & 1) * (KS * 32)) + 4 * ((lr >> 4) * 32 + bl * 4 + (((lr & 15) & 7) >> 1)) + ((lr & 15) >> 3);
                        w[0] = hi;
                        w[2] = lo;
                    }
                if (lane == 0) inv_s[(s & 1) * 8 + bl] = inv * winv;
            }
        }
        PH(2);  // cells + fragments
        if (__syncthreads_or(my_abort)) break;      // B fragments of all videos complete
        PH(3);  // barrier
        if (t > 0) {
            // partial[k][b] = sum over own rows of W_hh1[row][k] * d gates1[b][row]: 2 m-tiles of 16 columns per warp
            float dm[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, ds[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint4 b = da_s[(s & 1) * (KS * 32) + ks * 32 + lane];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    mma4(dm[m], ahi[m][ks], b.x, b.y);
                    if constexpr (!SINGLE) {
                        mma4(ds[m], ahi[m][ks], b.z, b.w);
                        mma4(ds[m], alo_s[((2 * warp + m) * KS + ks) * 32 + lane], b.x, b.y);
                    }
                }
            }
            // D: (column k = mt*16 + g (+8), videos 2tq, 2tq+1); column k belongs to consumer k >> 6, unit k & 63;
            // ring word [consumer][producer = slice][video][unit]
            const uint32_t par = step_parity(s);
            uint32_t* pub = ring + (size_t)(s & 1) * kSlot;
            const float inv0 = inv_s[(s & 1) * 8 + 2 * tq], inv1 = inv_s[(s & 1) * 8 + 2 * tq + 1];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = (2 * warp + m) * 16 + g + 8 * (q >> 1), b = 2 * tq + (q & 1);
                    st_flagged(pub + ((size_t)((k >> 6) * NSL + slice) * kGroup + b) * U + (k & 63), (dm[m][q] + ds[m][q]) * ((q & 1) ? inv1 : inv0), par);
                }
            // the fragments and scales of the NEXT step go to the other buffer; the one read here is rewritten two steps on,
            // behind the next step's barrier: no second barrier per step
        }
        // d gates1 of frame t (exact), behind the publish: off the critical path of the other CTAs
        if (valid) {
            float* dg = p.dgates1 + (row0 + t) * (size_t)(4 * H1) + uu;
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<float2*>(dg + q * H1) = make_float2(dgv[0][q], dgv[1][q]);
        }
        PH(4);  // MMAs + publish + d gates stores
    }
#ifdef OPN_LSTM_PHASES
    if (threadIdx.x == 0 && blockIdx.x == 0) {      // behind the fused kernel's own counters (words 32.., 64..)
        unsigned long long* o = reinterpret_cast<unsigned long long*>(p.status) + 96;
        for (int i = 0; i < 8; ++i) o[i] = (unsigned long long)ph_acc[i];
    }
#endif
}

}  // namespace l1b
}  // namespace

int launch_opnet_l1bwd(const L1BwdParams& p, int64_t B, bool single, cudaStream_t s, int group_begin, int group_end) {
    if (single)
        return launch_ring(l1b::opnet_l1bwd_kernel<true>, p, l1b::NT, l1b::NSL + 1, (size_t)l1b::SMEM_BYTES, B, s, "opnet_l1bwd", 1, group_begin, group_end);
    return launch_ring(l1b::opnet_l1bwd_kernel<false>, p, l1b::NT, l1b::NSL + 1, (size_t)l1b::SMEM_BYTES, B, s, "opnet_l1bwd", 1, group_begin, group_end);
}

// loads both variants of the kernel before the LSTM2 kernel starts (a lazy module load behind a running kernel waits for it)
int preload_opnet_l1bwd() {
    static bool done = false;
    if (done) return OPN_OK;
    cudaFuncAttributes a;
    OPN_CUDA(cudaFuncGetAttributes(&a, l1b::opnet_l1bwd_kernel<false>));
    OPN_CUDA(cudaFuncGetAttributes(&a, l1b::opnet_l1bwd_kernel<true>));
    OPN_CUDA(cudaFuncSetAttribute(l1b::opnet_l1bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, l1b::SMEM_BYTES));
    OPN_CUDA(cudaFuncSetAttribute(l1b::opnet_l1bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, l1b::SMEM_BYTES));
    done = true;
    return OPN_OK;
}

size_t opnet_l1bwd_ring_words_per_group() { return 2 * l1b::kSlot; }

}  // namespace opn
