// OPNet LSTM1 (90 -> 256) and the who-to-track head as a PRODUCER kernel on the SMs the fused forward leaves idle
// (baselines/learned_models.py:36-43).
//
// Why: opn_opnet_fused.cu runs LSTM1, the head and LSTM2 on the same 128 CTAs so that each layer's exchange wait hides the
// other's work -- but the LSTM1 / head work of a frame (2,690 clocks: gather-split, 144 MMAs, cells, softmax) is longer
// than the h2 exchange it hides, so a frame costs the SUM of both layers' busy phases (6,020 clocks).  Without that work
// on its SMs the same loop runs at 2.41 us per frame instead of 3.30 (profiles/r02_fused_l2only_probe.log).  The 128-CTA launch
// leaves 20 of the 148 SMs idle; this kernel puts LSTM1 + head there: 5 CTAs per batch group of 8 videos -- four with 64
// hidden units each and one for the who-to-track head -- and the fused kernel (EXT mode) only consumes frames_boxes[t]
// as it becomes available.
//
// CTA (slice s of 4, batch group g), 256 threads, one per SM:
//   * W_hh1 slice: 256 gate rows x 256 as fp16 hi / lo A fragments (x 2^s): the hi plane in REGISTERS (128 per thread),
//     the lo plane in SHARED memory (128 KB) -- together they are the whole register file of an SM.
//   * per frame: poll the flagged h1[t-1] tile of the group (L2 ring, B-fragment order, as the fused kernel), split to fp16,
//     96 MMAs per warp (2 m-tiles x 16 k-steps x {hi.hi, hi.lo, lo.hi}), cells (2 per thread), publish h1[t];
//     head of frame t-1: 6 MMAs per warp against W_pred fragments, softmax and the probability-weighted box sum for the
//     slice's two videos (finished one iteration later, in the shadow of the next poll), exact outputs for the backward
//     pass, and frames_boxes[t-1] with a ready bit in the mantissa LSB for the consumer.
//   * outputs: hs1, gates1, cells1, logits, probs, frames_boxes exactly as the fused kernel leaves them.
#include <stdlib.h>

#include "opn_mma_common.cuh"

namespace opn {

struct L1HeadParams {
    const float* boxes;    // [B,T,15,6]
    const float* xproj1;   // [B,T,4*256]
    const float* w_hh1;    // [1024,256]
    const float* w_pred;   // [15,256]
    float *hs1, *gates1, *cells1;   // gates1 / cells1 NULL in inference
    float *logits, *probs, *fb;     // [B,15,T], [B,T,15], [B,T,6]
    uint32_t* fbx;         // [groups][T][8 videos][8]: frames_boxes for the consumer kernel, ready bit in the LSB (zeroed per call)
    unsigned int* flags;   // unused (reserved)
    uint32_t* ring1;       // [groups][2][8*256] flagged words, fragment order
    unsigned int* status;
    int B, T;
    int group_offset, n_slices;
};

namespace {

namespace l1 {

constexpr int H1 = 256, NOBJ = 15, NFEAT = 6, BOXROW = NOBJ * NFEAT;
constexpr int NT = 256, NW = 8, NSL = 4, U = 64, KS = H1 / 16, MT = 16;
constexpr int DPAD = 68;      // d_s row: 64 units + pad (conflict-free fragment stores)

constexpr int OFF_ALO = 0;                                  // uint4 [MT][KS][32]        128 KB
constexpr int OFF_AP = OFF_ALO + MT * KS * 32 * 16;         // uint4 [KS][2][32]          16 KB
constexpr int OFF_BFRAG = OFF_AP + KS * 2 * 32 * 16;        // uint4 [KS*32]               8 KB
constexpr int OFF_D = OFF_BFRAG + KS * 32 * 16;             // float [4 gates][8][DPAD]
constexpr int OFF_DL = OFF_D + 4 * 8 * DPAD * 4;            // float [8 warps][16][8]
constexpr int OFF_PROBS = OFF_DL + 8 * 16 * 8 * 4;          // float [8][16]
constexpr int OFF_RED = OFF_PROBS + 8 * 16 * 4;             // float [8]
constexpr int OFF_BOX = OFF_RED + 64;                       // float [3 frames][8 videos][96]: boxes one frame ahead (head CTA; the
                                                            // warps that prefetch run up to a frame ahead of those that consume)
constexpr int SMEM_BYTES = OFF_BOX + 3 * 8 * 96 * 4;

__device__ __forceinline__ void mma4(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float frag_max(const float* ra, const float* rb, int k0) {
    float m = 0.f;
    if (ra) {
        const float2 a0 = __ldg(reinterpret_cast<const float2*>(ra + k0)), a2 = __ldg(reinterpret_cast<const float2*>(ra + k0 + 8));
        m = fmaxf(fmaxf(fabsf(a0.x), fabsf(a0.y)), fmaxf(fabsf(a2.x), fabsf(a2.y)));
    }
    if (rb) {
        const float2 a1 = __ldg(reinterpret_cast<const float2*>(rb + k0)), a3 = __ldg(reinterpret_cast<const float2*>(rb + k0 + 8));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(a1.x), fabsf(a1.y)), fmaxf(fabsf(a3.x), fabsf(a3.y))));
    }
    return m;
}
__device__ __forceinline__ void make_frag(const float* ra, const float* rb, int k0, float scale, uint4& hi, uint4& lo) {
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

// SINGLE: the 1e-2 arithmetic mode (the hi.hi product alone)
template <bool SINGLE>
__global__ void __launch_bounds__(NT, 1) opnet_l1head_kernel(const L1HeadParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint4* alo_s = reinterpret_cast<uint4*>(smem + OFF_ALO);
    uint4* ap_s = reinterpret_cast<uint4*>(smem + OFF_AP);
    uint4* bfrag_s = reinterpret_cast<uint4*>(smem + OFF_BFRAG);
    float* d_s = reinterpret_cast<float*>(smem + OFF_D);          // [(gate*8 + video)*DPAD + unit]
    float* dl_s = reinterpret_cast<float*>(smem + OFF_DL);        // [warp][object row][8 videos]
    float* probs_s = reinterpret_cast<float*>(smem + OFF_PROBS);  // [8 videos][16]
    float* red_s = reinterpret_cast<float*>(smem + OFF_RED);
    float* box_s = reinterpret_cast<float*>(smem + OFF_BOX);      // [3 frames][8 videos][96]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int slice = blockIdx.x % p.n_slices;      // 0 .. 3: 64 hidden units each; 4: the who-to-track head of the group
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    uint32_t* ring = p.ring1 + (size_t)group * (2 * kGroup * H1);
    uint32_t* fbx = p.fbx + (size_t)group * T * 64;

    for (int i = tid; i < KS * 32; i += NT) bfrag_s[i] = make_uint4(0u, 0u, 0u, 0u);   // absent videos stay zero
    int my_abort = 0;
    // h1[t] of the group: poll this thread's two fragment vectors of the ring slot, split to fp16 B fragments
    auto gather_h1 = [&](int t) {
        const uint32_t par = step_parity(t);
        const uint32_t* src = ring + (size_t)(t & 1) * (kGroup * H1);
        auto vec_valid = [&](int q) { return (((tid + NT * q) & 31) >> 2) < nvalid; };
        uint4 v[2];
        if (!gather_flagged(v, [&](int q) { return src + (size_t)(tid + NT * q) * 4; }, vec_valid, par, p.status, t)) my_abort = 1;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (vec_valid(q)) {
                uint4 f;
                split2(__uint_as_float(v[q].x), __uint_as_float(v[q].y), f.x, f.z);
                split2(__uint_as_float(v[q].z), __uint_as_float(v[q].w), f.y, f.w);
                bfrag_s[tid + NT * q] = f;
            }
        }
    };

    if (slice == NSL) {
        // ======================================= head CTA: the 8 videos of the group =======================================
        float winvp;
        {
            float m = 0.0f;
            for (int e = tid; e < KS * 32; e += NT) {
                const int l = e & 31, ks = e >> 5, gg = l >> 2;
                const float* ra = p.w_pred + (size_t)gg * H1;
                const float* rb = (gg + 8 < NOBJ) ? p.w_pred + (size_t)(gg + 8) * H1 : nullptr;
                m = fmaxf(m, frag_max(ra, rb, 16 * ks + 2 * (l & 3)));
            }
            float wscalep;
            weight_scale<NW>(m, red_s, wscalep, winvp);
            for (int e = tid; e < KS * 32; e += NT) {
                const int l = e & 31, ks = e >> 5, gg = l >> 2;
                const float* ra = p.w_pred + (size_t)gg * H1;
                const float* rb = (gg + 8 < NOBJ) ? p.w_pred + (size_t)(gg + 8) * H1 : nullptr;
                uint4 hi, lo;
                make_frag(ra, rb, 16 * ks + 2 * (l & 3), wscalep, hi, lo);
                ap_s[(ks * 2 + 0) * 32 + l] = hi;
                ap_s[(ks * 2 + 1) * 32 + l] = lo;
            }
        }
        // boxes of frame t of the 8 videos -> shared memory, asynchronously, a frame ahead (they come from HBM): warps 4-7
        auto prefetch_boxes = [&](int t) {
            if (warp >= 4) {
                if (t < T) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const int e = (tid - 128) + 128 * q, v = e / BOXROW, o = e % BOXROW;
                        if (e < 8 * BOXROW && v < nvalid) {
                            const float* src = p.boxes + ((size_t)(b0 + v) * T + t) * BOXROW + o;
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(box_s + ((t % 3) * 8 + v) * 96 + o)), "l"(src) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        };
        prefetch_boxes(0);
        const int hb = tid >> 4, ho = tid & 15;      // finalisation (warps 0-3): video, object
        __syncthreads();
        for (int t = 0; t < T; ++t) {
            gather_h1(t);
            prefetch_boxes(t + 1);
            if (__syncthreads_or(my_abort)) break;
            {   // logits of frame t: k-steps 2*warp, 2*warp + 1
                float dm[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int ks = 2 * warp + j;
                    const uint4 b = bfrag_s[ks * 32 + lane];
                    const uint4 ah = ap_s[(ks * 2 + 0) * 32 + lane];
                    mma4(dm, ah, b.x, b.y);
                    if constexpr (!SINGLE) {
                        mma4(ds, ah, b.z, b.w);
                        mma4(ds, ap_s[(ks * 2 + 1) * 32 + lane], b.x, b.y);
                    }
                }
                float* d = dl_s + (warp * 16 + g) * 8 + 2 * tq;
                *reinterpret_cast<float2*>(d) = make_float2((dm[0] + ds[0]) * winvp, (dm[1] + ds[1]) * winvp);
                *reinterpret_cast<float2*>(d + 64) = make_float2((dm[2] + ds[2]) * winvp, (dm[3] + ds[3]) * winvp);
            }
            if (warp >= 4) asm volatile("cp.async.wait_group 1;" ::: "memory");      // boxes of frame t have landed
            __syncthreads();
            if (warp < 4) {
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
                    const float* bxr = box_s + ((t % 3) * 8 + hb) * 96;
#pragma unroll
                    for (int o = 0; o < NOBJ; ++o) fbv = fmaf(probs_s[hb * 16 + o], bxr[o * NFEAT + ho], fbv);
                }
                if (hb < nvalid) {
                    const size_t bb = (size_t)(b0 + hb);
                    if (ho < NFEAT) {
                        // for the consumer kernel: the value with the ready bit in its mantissa LSB (<= 1 ulp; the exact value
                        // goes to frames_boxes for the backward pass).  A 32-bit store is single-copy atomic: no fence, no flag.
                        asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(fbx + ((size_t)t * 8 + hb) * 8 + ho), "r"(flagged(fbv, 1u)) : "memory");
                        p.fb[(bb * T + t) * NFEAT + ho] = fbv;
                    }
                    if (ho < NOBJ) {
                        p.logits[(bb * NOBJ + ho) * T + t] = logit;
                        p.probs[(bb * T + t) * NOBJ + ho] = pr;
                    }
                }
                __syncwarp();
            }
        }
        return;
    }

    // ============================================ unit CTAs: 64 hidden units each ============================================
    const int u0 = slice * U;
    // ---- W_hh1 slice: local row lr = gate*64 + unit; warp w owns m-tiles 2w, 2w+1 (gate w/2, units 32*(w&1) ..) ----------
    auto rows_of = [&](int mt, int gg, const float*& ra, const float*& rb) {
        const int la = mt * 16 + gg, lb = la + 8;
        ra = p.w_hh1 + (size_t)((la >> 6) * H1 + u0 + (la & 63)) * H1;
        rb = p.w_hh1 + (size_t)((lb >> 6) * H1 + u0 + (lb & 63)) * H1;
    };
    float wmax = 0.0f;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const float *ra, *rb;
        rows_of(2 * warp + m, g, ra, rb);
        for (int ks = 0; ks < KS; ++ks) wmax = fmaxf(wmax, frag_max(ra, rb, 16 * ks + 2 * tq));
    }
    float wscale, winv;
    weight_scale<NW>(wmax, red_s, wscale, winv);
    uint4 ahi[2][KS];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const float *ra, *rb;
        rows_of(2 * warp + m, g, ra, rb);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint4 lo;
            make_frag(ra, rb, 16 * ks + 2 * tq, wscale, ahi[m][ks], lo);
            alo_s[((2 * warp + m) * KS + ks) * 32 + lane] = lo;
        }
    }
    // ---- cells: warp = video, lane = units 2*lane, 2*lane + 1: every load / store of the two cells is one 8-byte access
    // (x-projection, stash, the published pair of ring words), half the memory instructions of two unrelated cells
    const int bl = warp, ul = 2 * lane;
    const bool valid = b0 + bl < p.B;
    const size_t row0 = (size_t)(valid ? b0 + bl : 0) * T;
    const int uu = u0 + ul;
    float cst[2] = {0.0f, 0.0f};
    float2 xp[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) xp[q] = valid ? __ldg(reinterpret_cast<const float2*>(p.xproj1 + row0 * (4 * H1) + q * H1 + uu)) : make_float2(0.f, 0.f);
    const int pub_word = frag_word(bl, uu);      // units uu (even), uu + 1 are adjacent words of a fragment vector
    __syncthreads();

    PH_DECL
    // iteration i: LSTM1 frame i, needs h1[i-1]
    for (int i = 0; i < T; ++i) {
        PH(0);  // stash stores of the previous iteration
        if (i >= 2 && tid == 0) {
            // back-pressure: h1[i] will overwrite the ring slot of h1[i-2]; the head CTA (not part of the h1 dependency chain)
            // must have read it -- it has once frames_boxes[i-2] is out (normally long ago: one relaxed load)
            const uint32_t* w = fbx + (size_t)(i - 2) * 64;
            const long long t0 = clock64();
            unsigned spins = 0;
            while (!(ld_relaxed(w) & 1u)) {
                if ((++spins & 63u) == 0 && poll_expired(t0, p.status, i)) {
                    my_abort = 1;
                    break;
                }
            }
        }
        if (i >= 1) gather_h1(i - 1);
        PH(1);  // poll + split
        if (__syncthreads_or(my_abort)) break;
        PH(2);  // barrier
        if (i >= 1) {
            float dm[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, ds[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const uint4 b = bfrag_s[ks * 32 + lane];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    mma4(dm[m], ahi[m][ks], b.x, b.y);
                    if constexpr (!SINGLE) {
                        mma4(ds[m], ahi[m][ks], b.z, b.w);
                        mma4(ds[m], alo_s[((2 * warp + m) * KS + ks) * 32 + lane], b.x, b.y);
                    }
                }
            }
            // D fragment: (row g, videos 2tq, 2tq+1), (row g+8, same videos); local row = gate*64 + unit
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int la = (2 * warp + m) * 16 + g, gate = la >> 6, un = la & 63;
                float* d = d_s + (gate * 8 + 2 * tq) * DPAD + un;
                d[0] = (dm[m][0] + ds[m][0]) * winv;
                d[DPAD] = (dm[m][1] + ds[m][1]) * winv;
                d[8] = (dm[m][2] + ds[m][2]) * winv;
                d[DPAD + 8] = (dm[m][3] + ds[m][3]) * winv;
            }
        }
        PH(3);  // MMAs
        __syncthreads();
        PH(4);  // barrier
        {
            // ---- cells of frame i: gates, cell update, publish first, stash after ------------------------------------------
            float hv[2], act[4][2];
            float2 dpre[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                dpre[q] = i >= 1 ? *reinterpret_cast<const float2*>(d_s + (q * 8 + bl) * DPAD + ul) : make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float a0 = (j ? xp[0].y + dpre[0].y : xp[0].x + dpre[0].x), a1 = (j ? xp[1].y + dpre[1].y : xp[1].x + dpre[1].x);
                const float a2 = (j ? xp[2].y + dpre[2].y : xp[2].x + dpre[2].x), a3 = (j ? xp[3].y + dpre[3].y : xp[3].x + dpre[3].x);
                act[0][j] = fmaf(0.5f, tanh_sfu(0.5f * a0), 0.5f);
                act[1][j] = fmaf(0.5f, tanh_sfu(0.5f * a1), 0.5f);
                act[2][j] = tanh_sfu(a2);
                act[3][j] = fmaf(0.5f, tanh_sfu(0.5f * a3), 0.5f);
                cst[j] = fmaf(act[1][j], cst[j], act[0][j] * act[2][j]);
                hv[j] = act[3][j] * tanh_sfu(cst[j]);
            }
            if (valid) {      // also the last frame: its head still needs h1[T-1]
                const uint32_t par = step_parity(i);
                asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1,%2};" ::"l"(ring + (size_t)(i & 1) * (kGroup * H1) + pub_word),
                             "r"(flagged(hv[0], par)), "r"(flagged(hv[1], par))
                             : "memory");
                const size_t r = row0 + i;
                if (i + 1 < T) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) xp[q] = __ldg(reinterpret_cast<const float2*>(p.xproj1 + (r + 1) * (4 * H1) + q * H1 + uu));
                }
                *reinterpret_cast<float2*>(p.hs1 + r * H1 + uu) = make_float2(hv[0], hv[1]);
                if (p.gates1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) *reinterpret_cast<float2*>(p.gates1 + r * (4 * H1) + q * H1 + uu) = make_float2(act[q][0], act[q][1]);
                }
                if (p.cells1) *reinterpret_cast<float2*>(p.cells1 + r * H1 + uu) = make_float2(cst[0], cst[1]);
            }
        }
        PH(5);  // cells + publish + stash
    }
    PH_STORE(p.status);
}

}  // namespace l1
}  // namespace

// host side: called by opn_opnet_fwd (opn_opnet_fused.cu)
int launch_opnet_l1head(const L1HeadParams& p, int64_t B, bool single, cudaStream_t s, int group_begin, int group_end) {
    if (single)
        return launch_ring(l1::opnet_l1head_kernel<true>, p, l1::NT, l1::NSL + 1, (size_t)l1::SMEM_BYTES, B, s, "opnet_l1head", 1, group_begin, group_end);
    return launch_ring(l1::opnet_l1head_kernel<false>, p, l1::NT, l1::NSL + 1, (size_t)l1::SMEM_BYTES, B, s, "opnet_l1head", 1, group_begin, group_end);
}


// loads both variants of the kernel (CUDA loads modules lazily) -- see the call site
int preload_opnet_l1head() {
    static bool done = false;
    if (done) return OPN_OK;
    cudaFuncAttributes a;
    OPN_CUDA(cudaFuncGetAttributes(&a, l1::opnet_l1head_kernel<false>));
    OPN_CUDA(cudaFuncGetAttributes(&a, l1::opnet_l1head_kernel<true>));
    OPN_CUDA(cudaFuncSetAttribute(l1::opnet_l1head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, l1::SMEM_BYTES));
    OPN_CUDA(cudaFuncSetAttribute(l1::opnet_l1head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, l1::SMEM_BYTES));
    done = true;
    return OPN_OK;
}

}  // namespace opn
