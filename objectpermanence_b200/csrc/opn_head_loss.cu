// The bbox head, the training loss and their backward as ONE pass over the hidden states ("the loss as an
// epilogue of the head"):
//
//   y = h W^T                       prediction(s)_layer, Linear H -> 4 without bias (baselines/learned_models.py:33,47,117,150,196)
//   loss = mean |y - labels| [* mask] [+ 0.5 mean_t ||y[t+1] - y[t]||]      (baselines/training_main.py:192-210)
//   dy = d loss / dy,   d_h = dy W,   d_W = dy^T h
//
// As separate launches (row-dot head, loss, dy W, zero fill, column reduction) this read h twice and took 52 us of the
// 2.36 ms headline step, none of them near the HBM rate.  Here a warp owns a run of consecutive rows: it keeps a two-row
// window of h in registers, forms y of the next row (halo rows at both ends of the run, for the consistency term), then the
// loss terms and dy of the current row, writes d_h and accumulates its share of d_W in registers.  h is read once (+ halo),
// d_h written once.  The d_W shares are summed per CTA in shared memory (fixed order) and added to the zeroed result with
// fp32 reductions, like the loss terms.
#include "opn_common.cuh"

namespace opn {
namespace {

constexpr int HL_THREADS = 256, HL_WARPS = HL_THREADS / 32;

__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

struct HeadLossParams {
    const float* h;        // [rows, H]
    const float* w;        // [4, H]
    const float* labels;   // [rows, 4]
    const uint8_t* mask;   // [rows, 4] or NULL
    float* y;              // [rows, 4]
    float* loss;           // 3 floats, zeroed by the caller
    float* dh;             // [rows, H]
    float* dw;             // [4, H], zeroed by the caller
    int B, T, H, consistency, rows_per_warp;
};

// VPL: float4 vectors per lane (H = 128 * VPL)
template <int VPL>
__global__ void __launch_bounds__(HL_THREADS) head_loss_kernel(const HeadLossParams p) {
    extern __shared__ __align__(16) float smem[];
    float* w_s = smem;                 // [4][H]
    float* dw_s = smem + 4 * p.H;      // [4][H] CTA sum of the warps' shares
    __shared__ float red_s[2][HL_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, T = p.T;
    const long long rows = (long long)p.B * T;
    for (int i = tid; i < 4 * H; i += HL_THREADS) {
        w_s[i] = __ldg(p.w + i);
        dw_s[i] = 0.0f;
    }
    __syncthreads();

    const float w_pred = 1.0f / (float)(rows * 4);
    const long long pairs = (long long)p.B * (T - 1);
    const float w_cons = (p.consistency && pairs > 0) ? 0.5f / (float)pairs : 0.0f;
    float pred = 0.0f, cons = 0.0f;
    float4 acc[4][VPL];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[c][v] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_row = [&](long long r, float4 (&x)[VPL]) {
        const float4* src = reinterpret_cast<const float4*>(p.h + r * H) + lane;
#pragma unroll
        for (int v = 0; v < VPL; ++v) x[v] = __ldg(src + 32 * v);
    };
    auto head = [&](const float4 (&x)[VPL], float (&yy)[4]) {      // the four dot products, the result in every lane
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float s = 0.0f;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const float4 wv = reinterpret_cast<const float4*>(w_s + c * H)[lane + 32 * v];
                s = fmaf(x[v].x, wv.x, s), s = fmaf(x[v].y, wv.y, s), s = fmaf(x[v].z, wv.z, s), s = fmaf(x[v].w, wv.w, s);
            }
            yy[c] = warp_sum_all(s);
        }
    };

    const long long gw = (long long)blockIdx.x * HL_WARPS + warp;
    const long long r0 = gw * p.rows_per_warp;
    const long long r1 = (r0 + p.rows_per_warp < rows) ? r0 + p.rows_per_warp : rows;
    if (r0 < rows) {
        float4 hc[VPL], hn[VPL];
        float yp[4] = {0.f, 0.f, 0.f, 0.f}, yc[4], yn[4] = {0.f, 0.f, 0.f, 0.f};
        if (r0 % T != 0) {      // halo: the frame before the run, same video
            load_row(r0 - 1, hc);
            head(hc, yp);
        }
        load_row(r0, hc);
        head(hc, yc);
        for (long long r = r0; r < r1; ++r) {
            const int t = (int)(r % T);
            const bool has_next = t + 1 < T;      // row r + 1 is the next frame of the same video (may be the halo behind the run)
            if (has_next || r + 1 < r1) {         // ... or the first frame of the next video of this run
                load_row(r + 1, hn);
                head(hn, yn);
            }
            const float4 lv = __ldg(reinterpret_cast<const float4*>(p.labels + r * 4));
            const float ll[4] = {lv.x, lv.y, lv.z, lv.w};
            float g[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float d = yc[c] - ll[c];
                const float m = p.mask ? (float)p.mask[r * 4 + c] : 1.0f;
                pred += fabsf(d) * m;
                g[c] = ((d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f)) * m * w_pred;
            }
            if (has_next) {
                const float d0 = yn[0] - yc[0], d1 = yn[1] - yc[1], d2 = yn[2] - yc[2], d3 = yn[3] - yc[3];
                const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
                cons += nrm;
                if (nrm > 0.0f) {
                    const float s = w_cons / nrm;
                    g[0] -= d0 * s, g[1] -= d1 * s, g[2] -= d2 * s, g[3] -= d3 * s;
                }
            }
            if (t > 0) {
                const float d0 = yc[0] - yp[0], d1 = yc[1] - yp[1], d2 = yc[2] - yp[2], d3 = yc[3] - yp[3];
                const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
                if (nrm > 0.0f) {
                    const float s = w_cons / nrm;
                    g[0] += d0 * s, g[1] += d1 * s, g[2] += d2 * s, g[3] += d3 * s;
                }
            }
            if (lane == 0) *reinterpret_cast<float4*>(p.y + r * 4) = make_float4(yc[0], yc[1], yc[2], yc[3]);
            // d_h[r] = g W,  d_W += g^T h[r]
            float4* dst = reinterpret_cast<float4*>(p.dh + r * H) + lane;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 wv = reinterpret_cast<const float4*>(w_s + c * H)[lane + 32 * v];
                    o.x = fmaf(g[c], wv.x, o.x), o.y = fmaf(g[c], wv.y, o.y), o.z = fmaf(g[c], wv.z, o.z), o.w = fmaf(g[c], wv.w, o.w);
                    acc[c][v].x = fmaf(g[c], hc[v].x, acc[c][v].x), acc[c][v].y = fmaf(g[c], hc[v].y, acc[c][v].y);
                    acc[c][v].z = fmaf(g[c], hc[v].z, acc[c][v].z), acc[c][v].w = fmaf(g[c], hc[v].w, acc[c][v].w);
                }
                dst[32 * v] = o;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) yp[c] = yc[c], yc[c] = yn[c];
#pragma unroll
            for (int v = 0; v < VPL; ++v) hc[v] = hn[v];
        }
    }
    // ---- CTA sum of the d_W shares (shared-memory adds, warp after warp: fixed order) and of the loss terms -------------
    for (int w = 0; w < HL_WARPS; ++w) {
        if (warp == w) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    float4* d = reinterpret_cast<float4*>(dw_s + c * H) + lane + 32 * v;
                    float4 o = *d;
                    o.x += acc[c][v].x, o.y += acc[c][v].y, o.z += acc[c][v].z, o.w += acc[c][v].w;
                    *d = o;
                }
        }
        __syncthreads();
    }
    if (lane == 0) red_s[0][warp] = pred, red_s[1][warp] = cons;      // every lane holds the same sums
    // the CTA's share of d_W onto the zeroed result with fp32 reductions (a fixed-order sum of the 148 per-CTA shares by the
    // last CTA to finish was a 60 us serial tail: 1,184 dependent-latency loads per thread)
    for (int i = tid; i < 4 * H; i += HL_THREADS) atomicAdd(p.dw + i, dw_s[i]);
    __syncthreads();
    if (tid == 0) {
        float ps = 0.0f, cs = 0.0f;
        for (int w = 0; w < HL_WARPS; ++w) ps += red_s[0][w], cs += red_s[1][w];
        const float pred_mean = ps * w_pred;
        const float cons_mean = pairs > 0 ? cs / (float)pairs : 0.0f;
        atomicAdd(p.loss + 1, pred_mean);
        atomicAdd(p.loss + 2, cons_mean);
        atomicAdd(p.loss + 0, p.consistency ? pred_mean + 0.5f * cons_mean : pred_mean);
    }
}

int plan(int64_t B, int64_t T, int64_t H, int& grid, int& rows_per_warp) {
    OPN_CHECK_ARG(B > 0 && T > 0 && B * T < (1LL << 31), "head_loss: bad shape");
    if (H % 128 != 0 || H < 128 || H > 512) {
        set_error("head_loss: hidden size %lld unsupported (a multiple of 128 up to 512)", (long long)H);
        return OPN_ERR_UNSUPPORTED;
    }
    int nsm = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || nsm <= 0) nsm = 148;
    const long long rows = B * T, warps = (long long)nsm * HL_WARPS;
    rows_per_warp = (int)((rows + warps - 1) / warps);
    if (rows_per_warp < 4) rows_per_warp = 4;      // the halo rows cost two extra rows per run
    grid = (int)((rows + (long long)rows_per_warp * HL_WARPS - 1) / ((long long)rows_per_warp * HL_WARPS));
    return OPN_OK;
}

}  // namespace
}  // namespace opn

using namespace opn;

extern "C" int64_t opn_head_loss_workspace_bytes(int64_t B, int64_t T, int64_t H) {
    int grid = 0, rpw = 0;
    if (plan(B, T, H, grid, rpw) != OPN_OK) return 0;
    return 256;      // reserved
}

extern "C" int opn_head_loss(int64_t B, int64_t T, int64_t H, const float* h, const float* w, const float* labels, const uint8_t* mask,
                             int consistency, float* y, float* loss_out, float* d_h, float* d_w, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    OPN_CHECK_ARG(h && w && labels && y && loss_out && d_h && d_w && workspace, "head_loss: null pointer");
    HeadLossParams p;
    int grid = 0;
    int rc = plan(B, T, H, grid, p.rows_per_warp);
    if (rc != OPN_OK) return rc;
    (void)workspace, (void)workspace_bytes;      // reserved (no scratch needed at present)
    cudaStream_t s = as_stream(stream);
    OPN_CUDA(cudaMemsetAsync(d_w, 0, (size_t)4 * H * sizeof(float), s));
    OPN_CUDA(cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), s));
    p.h = h, p.w = w, p.labels = labels, p.mask = mask, p.y = y, p.loss = loss_out, p.dh = d_h, p.dw = d_w;
    p.B = (int)B, p.T = (int)T, p.H = (int)H, p.consistency = consistency;
    const size_t smem = (size_t)8 * H * sizeof(float);
    switch (H / 128) {
        case 1: head_loss_kernel<1><<<grid, HL_THREADS, smem, s>>>(p); break;
        case 2: head_loss_kernel<2><<<grid, HL_THREADS, smem, s>>>(p); break;
        case 3: head_loss_kernel<3><<<grid, HL_THREADS, smem, s>>>(p); break;
        default: head_loss_kernel<4><<<grid, HL_THREADS, smem, s>>>(p); break;
    }
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}
