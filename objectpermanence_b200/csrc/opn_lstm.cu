// Persistent, weight-stationary LSTM recurrence for sm_100a (forward and backward).
//
// Replaces the nn.LSTM calls of baselines/learned_models.py:39,46,76,113,146,192 and their
// autograd backward.  Design:
//
//   * One cooperative launch runs all T steps.  A CTA (128 threads) owns 8 hidden units
//     (forward: their 32 gate rows of W_hh; backward: their 8 columns of W_hh, i.e. 8 rows
//     of W_hh^T) for one *batch group* of 8 videos.  The CTA's weight slice lives in
//     registers for the whole sequence (8 rows x KPT values per thread, KPT = H/32).
//   * Per step every CTA needs the full recurrent vector of its batch group
//     (forward: h_{t-1} [8,H]; backward: dgates_t [8,4H]).  It is fetched from L2 into
//     shared memory with TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) in up to four
//     chunks, each completing its own mbarrier so the FFMA loop starts on chunk 0 while
//     the rest is still in flight.
//   * Cross-CTA ordering is a per-(batch group, step) arrival counter in global memory:
//     producers red.release.gpu after storing their slice, the consumer's elected thread
//     ld.acquire.gpu-polls it.  Batch groups are independent chains; two CTAs are resident
//     per SM so one chain's wait overlaps another chain's FFMA work.
//   * The [8 rows x 8 videos] partial sums of the thread tile are reduced across the
//     K-split with a 62-shuffle transposing butterfly (no shared memory in the forward).
//   * sigmoid/tanh, the cell update and the stash writes are fused into the step.
//
// Every spin wait has a clock64() time-out: on expiry the kernel records a status word and
// all CTAs leave the time loop (no hang, no trap).
#include "opn_common.cuh"

namespace opn {

namespace {

constexpr int kThreads = 128;
constexpr int kGroup = 8;   // videos per batch group
constexpr int kUnits = 8;   // hidden units per CTA
constexpr long long kTimeoutCycles = 3000000000LL;  // ~1.5 s at 2 GHz

constexpr uint32_t kStatusPollTimeout = 1;
constexpr uint32_t kStatusMbarTimeout = 2;

struct FwdParams {
    const float* xproj;   // [B,T,4H]
    const float* w_hh;    // [4H,H]
    float* hs;            // [B,T,H]
    float* gates;         // [B,T,4H] or null
    float* cells;         // [B,T,H] or null
    unsigned int* counters;  // [n_groups_total][T]
    unsigned int* status;    // 4 words
    int B, T;
    int group_offset;  // first batch group handled by this launch
    int n_slices;      // H / 8
};

struct BwdParams {
    const float* w_t;     // [H,4H]  (W_hh transposed)
    const float* gates;   // [B,T,4H]
    const float* cells;   // [B,T,H]
    const float* dh_out;  // [B,T,H]
    float* dgates;        // [B,T,4H]
    unsigned int* counters;
    unsigned int* status;
    int B, T;
    int group_offset;
    int n_slices;
};

// k index of the e-th weight held by K-split `ks` (NS splits in total)
template <int KPT, int NS>
__device__ __forceinline__ int k_index(int ks, int e) {
    if (KPT >= 4) return 4 * ks + (4 * NS) * (e >> 2) + (e & 3);
    return ks + NS * e;
}

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

// Wait for an mbarrier phase with a time-out.  Returns false on time-out / observed abort.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* status) {
    if (mbar_try_wait(bar, parity)) return true;
    long long t0 = clock64();
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 255u) == 0) {
            if (ld_relaxed(status) != 0) return false;
            if (clock64() - t0 > kTimeoutCycles) {
                atomicCAS(status, 0u, kStatusMbarTimeout);
                return false;
            }
        }
    }
    return true;
}

// Elected-thread wait until *counter >= target.
__device__ __forceinline__ bool poll_counter(const unsigned int* counter, unsigned int target, unsigned int* status,
                                             int t) {
    if (ld_acquire(counter) >= target) return true;
    long long t0 = clock64();
    unsigned int spins = 0;
    while (ld_acquire(counter) < target) {
        if ((++spins & 63u) == 0) {
            if (ld_relaxed(status) != 0) return false;
            if (clock64() - t0 > kTimeoutCycles) {
                if (atomicCAS(status, 0u, kStatusPollTimeout) == 0u) {
                    status[1] = (unsigned int)t;
                    status[2] = blockIdx.x;
                    status[3] = *counter;
                }
                return false;
            }
        }
    }
    return true;
}

// acc[rr*8 + b] += sum_e w[rr][e] * v_s[b][k_index(ks, e)]  for the thread's K-split.
// v_s is the shared-memory copy of the recurrent vector, [8][K]; chunk c (K/NCH values per
// row) is valid once bars[c] has completed the phase `parity`.
template <int KPT, int NS>
__device__ __forceinline__ bool matvec_tile(const float (&w)[kUnits][KPT], const float* v_s, int ks, uint64_t* bars,
                                            uint32_t parity, unsigned int* status, float (&acc)[64]) {
    constexpr int K = NS * KPT;
    bool ok = true;
    if (KPT >= 4) {
#pragma unroll
        for (int j = 0; j < KPT / 4; ++j) {
            ok = mbar_wait(&bars[j], parity, status) && ok;
#pragma unroll
            for (int b = 0; b < kGroup; ++b) {
                const float4 hv = *reinterpret_cast<const float4*>(v_s + b * K + 4 * ks + (4 * NS) * j);
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
        ok = mbar_wait(&bars[0], parity, status);
#pragma unroll
        for (int e = 0; e < KPT; ++e) {
#pragma unroll
            for (int b = 0; b < kGroup; ++b) {
                const float hv = v_s[b * K + ks + NS * e];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) acc[rr * 8 + b] = fmaf(w[rr][e], hv, acc[rr * 8 + b]);
            }
        }
    }
    return ok;
}

// One stage of the transposing butterfly: the N live values (compact index) are halved; bit
// ABIT of the compact index is resolved by lane bit `mask`.
template <int N, int ABIT>
__device__ __forceinline__ void butterfly_stage(float (&v)[64], bool hi, int mask) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const int lo = ((i >> ABIT) << (ABIT + 1)) | (i & ((1 << ABIT) - 1));
        const int up = lo | (1 << ABIT);
        const float send = hi ? v[lo] : v[up];
        const float keep = hi ? v[up] : v[lo];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
}

// Sum acc[64] over the 32 lanes of the warp.  On return lane l holds in v[0], v[1] the totals
// of index  a = ((l>>4)&1)<<5 | (l&1)<<4 | j<<3 | ((l>>1)&7)   for j = 0, 1.
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[64], int lane) {
    butterfly_stage<64, 5>(v, (lane & 16) != 0, 16);
    butterfly_stage<32, 2>(v, (lane & 8) != 0, 8);
    butterfly_stage<16, 1>(v, (lane & 4) != 0, 4);
    butterfly_stage<8, 0>(v, (lane & 2) != 0, 2);
    butterfly_stage<4, 1>(v, (lane & 1) != 0, 1);
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
template <int KPT>
__global__ void __launch_bounds__(kThreads, 2) lstm_fwd_kernel(const FwdParams p) {
    constexpr int H = 32 * KPT;
    constexpr int NCH = (KPT >= 4) ? KPT / 4 : 1;
    constexpr int CH = H / NCH;  // floats per chunk per video row

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* h_s = reinterpret_cast<float*>(smem_raw);  // [8][H]
    __shared__ __align__(8) uint64_t bars[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slice = blockIdx.x % p.n_slices;
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int u0 = slice * kUnits;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    unsigned int* cnt = p.counters + (size_t)group * T;

    for (int i = tid; i < kGroup * H; i += kThreads) h_s[i] = 0.0f;
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) mbar_init(&bars[c], 1);
        mbar_fence_init();
    }

    // this warp's 8 gate rows: rr = uu*4 + gate, unit = u0 + 2*warp + uu
    float w[8][KPT];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const int row = (rr & 3) * H + u0 + 2 * warp + (rr >> 2);
        load_weight_row<KPT, 32>(w[rr], p.w_hh + (size_t)row * H, lane);
    }
    fence_proxy_async_smem();  // zero-fill of h_s (generic proxy) before any TMA write to it
    __syncthreads();

    // after the butterfly this lane owns gates (2*gh, 2*gh+1) of cell (unit u, video bb)
    const int uu = lane >> 4, bl = (lane >> 1) & 7, gh = lane & 1;
    const int u = u0 + 2 * warp + uu;
    const int bb = b0 + bl;
    const bool valid = bb < p.B;
    const size_t row0 = (size_t)(valid ? bb : 0) * T;
    const float* xp_ptr = p.xproj + row0 * (4 * H) + (size_t)(2 * gh) * H + u;

    float c_state = 0.0f;
    uint32_t parity = 0;
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
            if (warp == 0) {
                int ok = 1;
                if (lane == 0) ok = poll_counter(&cnt[t - 1], (unsigned int)p.n_slices, p.status, t) ? 1 : 0;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) {
                    fence_proxy_async_global();
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < NCH; ++c) mbar_arrive_expect_tx(&bars[c], (uint32_t)(nvalid * CH * 4));
                    }
                    __syncwarp();
                    for (int piece = lane; piece < kGroup * NCH; piece += 32) {
                        const int b = piece / NCH, c = piece % NCH;
                        if (b < nvalid)
                            bulk_g2s(h_s + b * H + c * CH, p.hs + ((size_t)(b0 + b) * T + (t - 1)) * H + c * CH,
                                     (uint32_t)(CH * 4), &bars[c]);
                    }
                } else {
                    my_abort = 1;
                }
            }
            if (!matvec_tile<KPT, 32>(w, h_s, lane, bars, parity, p.status, acc)) my_abort = 1;
            parity ^= 1u;
            warp_transpose_reduce(acc, lane);
        }

        // fused pointwise: gate activations, cell update, stash
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
        fence_proxy_async_global();  // h_t stores (generic proxy) -> later TMA reads by other CTAs
        const int abort = __syncthreads_or(my_abort);
        if (abort) break;
        if (tid == 0) red_release_add(&cnt[t], 1u);
    }
}

// ------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------
template <int KPT>
__global__ void __launch_bounds__(kThreads, 2) lstm_bwd_kernel(const BwdParams p) {
    constexpr int H = 32 * KPT;
    constexpr int K = 4 * H;
    constexpr int NCH = (KPT >= 4) ? KPT / 4 : 1;
    constexpr int CH = K / NCH;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* da_s = reinterpret_cast<float*>(smem_raw);  // [8][4H]
    __shared__ float red_s[4][64];
    __shared__ __align__(8) uint64_t bars[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slice = blockIdx.x % p.n_slices;
    const int group = p.group_offset + blockIdx.x / p.n_slices;
    const int u0 = slice * kUnits;
    const int b0 = group * kGroup;
    const int T = p.T;
    const int nvalid = min(kGroup, p.B - b0);
    unsigned int* cnt = p.counters + (size_t)group * T;

    for (int i = tid; i < kGroup * K; i += kThreads) da_s[i] = 0.0f;
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) mbar_init(&bars[c], 1);
        mbar_fence_init();
    }

    // rows rr = the CTA's 8 units; this thread's K-split is tid (128 splits of 4H)
    float w[8][KPT];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) load_weight_row<KPT, kThreads>(w[rr], p.w_t + (size_t)(u0 + rr) * K, tid);
    fence_proxy_async_smem();
    __syncthreads();

    // threads 0..63 own one cell (unit u0 + tid/8, video b0 + tid%8) of the recurrence state
    const bool cell_thread = tid < 64;
    const int ul = tid >> 3, bl = tid & 7;
    const int u = u0 + ul;
    const int bb = b0 + bl;
    const bool valid = cell_thread && (bb < p.B);
    const size_t row0 = (size_t)(valid ? bb : 0) * T;

    float dc_carry = 0.0f, dh_rec = 0.0f;
    float si = 0.f, sf = 0.f, sg = 0.f, so = 0.f, sc = 0.f, scp = 0.f, sdh = 0.f;
    auto load_stash = [&](int t) {
        const size_t row = row0 + t;
        const float* g = p.gates + row * (size_t)K + u;
        si = __ldg(g);
        sf = __ldg(g + H);
        sg = __ldg(g + 2 * H);
        so = __ldg(g + 3 * H);
        sc = __ldg(p.cells + row * H + u);
        scp = (t > 0) ? __ldg(p.cells + (row - 1) * H + u) : 0.0f;
        sdh = __ldg(p.dh_out + row * H + u);
    };
    if (valid) load_stash(T - 1);

    uint32_t parity = 0;
    int my_abort = 0;

    for (int t = T - 1; t >= 0; --t) {
        if (valid) {
            const float dh = sdh + dh_rec;
            const float tc = tanhf(sc);
            const float d_o = dh * tc;
            const float dc = fmaf(dh * so, 1.0f - tc * tc, dc_carry);
            const float d_i = dc * sg;
            const float d_g = dc * si;
            const float d_f = dc * scp;
            dc_carry = dc * sf;
            float* dg = p.dgates + (row0 + t) * (size_t)K + u;
            dg[0] = d_i * si * (1.0f - si);
            dg[H] = d_f * sf * (1.0f - sf);
            dg[2 * H] = d_g * (1.0f - sg * sg);
            dg[3 * H] = d_o * so * (1.0f - so);
            if (t > 0) load_stash(t - 1);  // prefetch: lands while the matvec runs
        }
        fence_proxy_async_global();
        const int abort = __syncthreads_or(my_abort);
        if (abort) break;
        if (tid == 0) red_release_add(&cnt[t], 1u);
        if (t == 0) break;

        if (warp == 0) {
            int ok = 1;
            if (lane == 0) ok = poll_counter(&cnt[t], (unsigned int)p.n_slices, p.status, t) ? 1 : 0;
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (ok) {
                fence_proxy_async_global();
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) mbar_arrive_expect_tx(&bars[c], (uint32_t)(nvalid * CH * 4));
                }
                __syncwarp();
                for (int piece = lane; piece < kGroup * NCH; piece += 32) {
                    const int b = piece / NCH, c = piece % NCH;
                    if (b < nvalid)
                        bulk_g2s(da_s + b * K + c * CH, p.dgates + ((size_t)(b0 + b) * T + t) * K + c * CH,
                                 (uint32_t)(CH * 4), &bars[c]);
                }
            } else {
                my_abort = 1;
            }
        }

        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.0f;
        if (!matvec_tile<KPT, kThreads>(w, da_s, tid, bars, parity, p.status, acc)) my_abort = 1;
        parity ^= 1u;
        warp_transpose_reduce(acc, lane);
        {
            const int a_lo = (((lane >> 4) & 1) << 5) | ((lane & 1) << 4) | ((lane >> 1) & 7);
            red_s[warp][a_lo] = acc[0];
            red_s[warp][a_lo | 8] = acc[1];
        }
        __syncthreads();
        if (cell_thread) dh_rec = red_s[0][tid] + red_s[1][tid] + red_s[2][tid] + red_s[3][tid];
    }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

// ---- host side ----------------------------------------------------------------------
struct WorkspaceLayout {
    size_t status_off, fwd_cnt_off, bwd_cnt_off, wt_off, total;
};

WorkspaceLayout layout(int64_t B, int64_t T, int64_t H) {
    const size_t groups = (size_t)((B + kGroup - 1) / kGroup);
    WorkspaceLayout l;
    l.status_off = 0;
    l.fwd_cnt_off = 256;
    const size_t cnt_bytes = ((groups * (size_t)T * sizeof(unsigned int)) + 255) / 256 * 256;
    l.bwd_cnt_off = l.fwd_cnt_off + cnt_bytes;
    l.wt_off = l.bwd_cnt_off + cnt_bytes;
    l.total = l.wt_off + (size_t)H * 4 * H * sizeof(float);
    return l;
}

bool supported_hidden(int64_t H) { return H == 32 || H == 64 || H == 128 || H == 256 || H == 512; }

template <typename Kernel>
int max_coresident(Kernel kernel, size_t smem, int* out) {
    int dev = 0, sms = 0, per_sm = 0;
    OPN_CUDA(cudaGetDevice(&dev));
    OPN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    OPN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OPN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
    *out = sms * per_sm;
    return OPN_OK;
}

template <int KPT>
int launch_fwd(FwdParams p, int64_t B, cudaStream_t stream) {
    constexpr int H = 32 * KPT;
    const size_t smem = (size_t)kGroup * H * sizeof(float);
    int cap = 0;
    int rc = max_coresident(lstm_fwd_kernel<KPT>, smem, &cap);
    if (rc != OPN_OK) return rc;
    const int n_slices = H / kUnits;
    const int groups = (int)((B + kGroup - 1) / kGroup);
    const int per_launch = cap / n_slices;
    if (per_launch < 1) {
        set_error("lstm_fwd: device cannot co-schedule %d CTAs (capacity %d)", n_slices, cap);
        return OPN_ERR_UNSUPPORTED;
    }
    p.n_slices = n_slices;
    for (int g0 = 0; g0 < groups; g0 += per_launch) {
        const int ng = groups - g0 < per_launch ? groups - g0 : per_launch;
        p.group_offset = g0;
        void* args[] = {(void*)&p};
        OPN_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_fwd_kernel<KPT>, dim3(n_slices * ng), dim3(kThreads),
                                             args, smem, stream));
        count_launch();
    }
    return OPN_OK;
}

template <int KPT>
int launch_bwd(BwdParams p, int64_t B, cudaStream_t stream) {
    constexpr int H = 32 * KPT;
    const size_t smem = (size_t)kGroup * 4 * H * sizeof(float);
    int cap = 0;
    int rc = max_coresident(lstm_bwd_kernel<KPT>, smem, &cap);
    if (rc != OPN_OK) return rc;
    const int n_slices = H / kUnits;
    const int groups = (int)((B + kGroup - 1) / kGroup);
    const int per_launch = cap / n_slices;
    if (per_launch < 1) {
        set_error("lstm_bwd: device cannot co-schedule %d CTAs (capacity %d)", n_slices, cap);
        return OPN_ERR_UNSUPPORTED;
    }
    p.n_slices = n_slices;
    for (int g0 = 0; g0 < groups; g0 += per_launch) {
        const int ng = groups - g0 < per_launch ? groups - g0 : per_launch;
        p.group_offset = g0;
        void* args[] = {(void*)&p};
        OPN_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_bwd_kernel<KPT>, dim3(n_slices * ng), dim3(kThreads),
                                             args, smem, stream));
        count_launch();
    }
    return OPN_OK;
}

}  // namespace
}  // namespace opn

using namespace opn;

extern "C" int64_t opn_lstm_workspace_bytes(int64_t B, int64_t T, int64_t H) {
    if (B <= 0 || T <= 0 || H <= 0) return 0;
    return (int64_t)layout(B, T, H).total;
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
    OPN_CHECK_ARG(workspace_bytes >= (int64_t)l.total, "lstm_fwd: workspace too small (%lld < %lld)",
                  (long long)workspace_bytes, (long long)l.total);
    cudaStream_t s = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    OPN_CUDA(cudaMemsetAsync(ws + l.status_off, 0, l.bwd_cnt_off - l.status_off, s));
    FwdParams p;
    p.xproj = xproj;
    p.w_hh = w_hh;
    p.hs = hs;
    p.gates = gates;
    p.cells = cells;
    p.counters = reinterpret_cast<unsigned int*>(ws + l.fwd_cnt_off);
    p.status = reinterpret_cast<unsigned int*>(ws + l.status_off);
    p.B = (int)B;
    p.T = (int)T;
    p.group_offset = 0;
    p.n_slices = 0;
    switch (H) {
        case 32: return launch_fwd<1>(p, B, s);
        case 64: return launch_fwd<2>(p, B, s);
        case 128: return launch_fwd<4>(p, B, s);
        case 256: return launch_fwd<8>(p, B, s);
        default: return launch_fwd<16>(p, B, s);
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
    OPN_CHECK_ARG(workspace_bytes >= (int64_t)l.total, "lstm_bwd: workspace too small");
    cudaStream_t s = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    OPN_CUDA(cudaMemsetAsync(ws + l.status_off, 0, 256, s));
    OPN_CUDA(cudaMemsetAsync(ws + l.bwd_cnt_off, 0, l.wt_off - l.bwd_cnt_off, s));
    float* w_t = reinterpret_cast<float*>(ws + l.wt_off);
    {
        dim3 grid((unsigned)((H + 31) / 32), (unsigned)((4 * H + 31) / 32));
        transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(w_hh, w_t, (int)(4 * H), (int)H);
        OPN_CUDA(cudaGetLastError());
        count_launch();
    }
    BwdParams p;
    p.w_t = w_t;
    p.gates = gates;
    p.cells = cells;
    p.dh_out = dh_out;
    p.dgates = dgates;
    p.counters = reinterpret_cast<unsigned int*>(ws + l.bwd_cnt_off);
    p.status = reinterpret_cast<unsigned int*>(ws + l.status_off);
    p.B = (int)B;
    p.T = (int)T;
    p.group_offset = 0;
    p.n_slices = 0;
    switch (H) {
        case 32: return launch_bwd<1>(p, B, s);
        case 64: return launch_bwd<2>(p, B, s);
        case 128: return launch_bwd<4>(p, B, s);
        case 256: return launch_bwd<8>(p, B, s);
        default: return launch_bwd<16>(p, B, s);
    }
}

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
        set_error("persistent LSTM kernel timed out (code %u, step %u, cta %u, counter %u)", words[0], words[1],
                  words[2], words[3]);
        return OPN_ERR_TIMEOUT;
    }
    return OPN_OK;
}
