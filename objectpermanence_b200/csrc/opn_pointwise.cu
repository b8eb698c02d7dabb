// Row-wise / element-wise kernels of the OPNet hot path (sm_100a):
//   * the "who to track" stage of OPNet (head mat-vec + softmax over 15 objects + weighted box
//     sum, baselines/learned_models.py:40-43) and its backward -- one warp per (video, frame),
//     warp-shuffle reductions for the small hidden-dim mat-vecs, coalesced box-row reads
//   * LayerNorm / row softmax / ReLU / bias-gradient helpers of the transformer_lstm encoder
//   * the fused training loss of baselines/training_main.py:192-210
#include "opn_common.cuh"

namespace opn {
namespace {

constexpr int NOBJ = OPN_MAX_OBJECTS;   // 15
constexpr int NFEAT = OPN_BOX_FEATURES; // 6
constexpr int BOXROW = NOBJ * NFEAT;    // 90

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// ---- who-to-track forward -------------------------------------------------------------
__global__ void __launch_bounds__(256) wtt_fwd_kernel(const float* __restrict__ boxes, const float* __restrict__ hs1,
                                                      const float* __restrict__ w_pred, float* __restrict__ logits_bpt,
                                                      float* __restrict__ probs, float* __restrict__ fb, int B, int T,
                                                      int H1) {
    extern __shared__ __align__(16) float sm[];
    float* wp_s = sm;                                  // [15][H1]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* box_s = sm + NOBJ * H1 + warp * 96;         // [90] per warp
    for (int i = threadIdx.x; i < NOBJ * H1; i += blockDim.x) wp_s[i] = w_pred[i];
    __syncthreads();

    const long long rows = (long long)B * T;
    for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
        float acc[NOBJ];
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) acc[o] = 0.0f;
        const float* h = hs1 + row * H1;
        for (int k = lane; k < H1; k += 32) {
            const float hv = __ldg(h + k);
#pragma unroll
            for (int o = 0; o < NOBJ; ++o) acc[o] = fmaf(wp_s[o * H1 + k], hv, acc[o]);
        }
        for (int e = lane; e < BOXROW; e += 32) box_s[e] = __ldg(boxes + row * BOXROW + e);
        float mx = -INFINITY;
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
            acc[o] = warp_sum(acc[o]);
            mx = fmaxf(mx, acc[o]);
        }
        float pr[NOBJ];
        float den = 0.0f;
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
            pr[o] = expf(acc[o] - mx);
            den += pr[o];
        }
        const float inv = 1.0f / den;
        __syncwarp();
        const int b = (int)(row / T), t = (int)(row % T);
        float my_logit = 0.0f, my_prob = 0.0f, my_fb = 0.0f;
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
            const float po = pr[o] * inv;
            if (lane == o) {
                my_logit = acc[o];
                my_prob = po;
            }
            if (lane < NFEAT) my_fb = fmaf(po, box_s[o * NFEAT + lane], my_fb);
        }
        if (lane < NOBJ) {
            logits_bpt[((long long)b * NOBJ + lane) * T + t] = my_logit;
            probs[row * NOBJ + lane] = my_prob;
        }
        if (lane < NFEAT) fb[row * NFEAT + lane] = my_fb;
        __syncwarp();
    }
}

// ---- who-to-track backward ------------------------------------------------------------
__global__ void __launch_bounds__(256) wtt_bwd_kernel(const float* __restrict__ boxes, const float* __restrict__ probs,
                                                      const float* __restrict__ w_pred, const float* __restrict__ dfb,
                                                      const float* __restrict__ dlogits_bpt,
                                                      float* __restrict__ dlogits, float* __restrict__ dhs1, int B,
                                                      int T, int H1) {
    extern __shared__ __align__(16) float sm[];
    float* wp_s = sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* box_s = sm + NOBJ * H1 + warp * 96;
    for (int i = threadIdx.x; i < NOBJ * H1; i += blockDim.x) wp_s[i] = w_pred[i];
    __syncthreads();

    const long long rows = (long long)B * T;
    for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
        for (int e = lane; e < BOXROW; e += 32) box_s[e] = __ldg(boxes + row * BOXROW + e);
        float g[NFEAT];
#pragma unroll
        for (int c = 0; c < NFEAT; ++c) g[c] = __ldg(dfb + row * NFEAT + c);
        __syncwarp();
        const int b = (int)(row / T), t = (int)(row % T);
        float dl[NOBJ];
        float dot = 0.0f;
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
            float dp = 0.0f;
#pragma unroll
            for (int c = 0; c < NFEAT; ++c) dp = fmaf(g[c], box_s[o * NFEAT + c], dp);
            const float po = __ldg(probs + row * NOBJ + o);
            dot = fmaf(po, dp, dot);
            dl[o] = dp;
        }
        float mine = 0.0f;
#pragma unroll
        for (int o = 0; o < NOBJ; ++o) {
            const float po = __ldg(probs + row * NOBJ + o);
            float v = po * (dl[o] - dot);
            if (dlogits_bpt) v += __ldg(dlogits_bpt + ((long long)b * NOBJ + o) * T + t);
            dl[o] = v;
            if (lane == o) mine = v;
        }
        if (lane < NOBJ) dlogits[row * NOBJ + lane] = mine;
        float* dh = dhs1 + row * H1;
        for (int k = lane; k < H1; k += 32) {
            float v = 0.0f;
#pragma unroll
            for (int o = 0; o < NOBJ; ++o) v = fmaf(dl[o], wp_s[o * H1 + k], v);
            dh[k] = v;
        }
        __syncwarp();
    }
}

// ---- element-wise ----------------------------------------------------------------------
__global__ void relu_bwd_kernel(long long n, const float* __restrict__ y, float* __restrict__ dy) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        if (!(y[i] > 0.0f)) dy[i] = 0.0f;
}
__global__ void add_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] + b[i];
}

// ---- dropout (train mode of nn.TransformerEncoderLayer, baselines/learned_models.py:166) ----
// Counter-based mask: element i belongs to Philox4x32-10 block (offset + i/4), word i%4, key = seed; it is kept when
// its word >= p * 2^32.  Nothing is stashed: the backward pass calls the same entry with the same (seed, offset).
__device__ __forceinline__ uint4 philox4x32_10(unsigned long long ctr, unsigned long long seed) {
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
__global__ void __launch_bounds__(256) dropout_kernel(long long n, const float* __restrict__ x, float* __restrict__ out,
                                                      unsigned int threshold, float scale, unsigned long long seed,
                                                      unsigned long long offset, int vec) {
    const long long blocks = (n + 3) >> 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < blocks;
         q += (long long)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(offset + (unsigned long long)q, seed);
        const long long i = q << 2;
        if (vec && i + 3 < n) {
            float4 v = *reinterpret_cast<const float4*>(x + i);
            v.x = r.x >= threshold ? v.x * scale : 0.0f;
            v.y = r.y >= threshold ? v.y * scale : 0.0f;
            v.z = r.z >= threshold ? v.z * scale : 0.0f;
            v.w = r.w >= threshold ? v.w * scale : 0.0f;
            *reinterpret_cast<float4*>(out + i) = v;
        } else {
            const unsigned int w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i + j < n) out[i + j] = w[j] >= threshold ? x[i + j] * scale : 0.0f;
        }
    }
}

// column sums: grid.x covers columns (32 per block), grid.y splits rows; fp32 atomics
__global__ void __launch_bounds__(256) colsum_kernel(long long rows, int cols, const float* __restrict__ x,
                                                     long long ld, float* __restrict__ out) {
    __shared__ float part[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    const long long chunk = (rows + gridDim.y - 1) / gridDim.y;
    const long long r0 = blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
    float s = 0.0f;
    if (col < cols)
        for (long long r = r0 + ry; r < r1; r += 8) s += x[r * ld + col];
    part[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && col < cols) {
        float v = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v += part[i][cx];
        atomicAdd(out + col, v);
    }
}
__global__ void zero_kernel(long long n, float* x) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = 0.0f;
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    float r = is_max ? -INFINITY : 0.0f;
    for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, scratch[i]) : r + scratch[i];
    return r;
}

// softmax over one row per CTA, row cached in shared memory
__global__ void __launch_bounds__(256) softmax_rows_kernel(int cols, float* __restrict__ x, long long ld, float scale) {
    extern __shared__ __align__(16) float row_s[];
    __shared__ float scratch[32];
    float* row = x + (long long)blockIdx.x * ld;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        const float v = row[c] * scale;
        row_s[c] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_reduce(mx, true, scratch);
    float sum = 0.0f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        const float e = expf(row_s[c] - mx);
        row_s[c] = e;
        sum += e;
    }
    sum = block_reduce(sum, false, scratch);
    const float inv = 1.0f / sum;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) row[c] = row_s[c] * inv;
}

__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(int cols, const float* __restrict__ p,
                                                               float* __restrict__ dp, long long ld, float scale) {
    __shared__ float scratch[32];
    const float* pr = p + (long long)blockIdx.x * ld;
    float* dr = dp + (long long)blockIdx.x * ld;
    float dot = 0.0f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) dot = fmaf(pr[c], dr[c], dot);
    dot = block_reduce(dot, false, scratch);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) dr[c] = scale * pr[c] * (dr[c] - dot);
}

// LayerNorm over D (one warp per row): y = (z - mean) * rstd * w + b with z = x + res
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(long long rows, int D, const float* __restrict__ x,
                                                            const float* __restrict__ res,
                                                            const float* __restrict__ w, const float* __restrict__ b,
                                                            float eps, float* __restrict__ y,
                                                            float* __restrict__ xhat, float* __restrict__ rstd_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float* xr = x + row * D;
    const float* rr = res ? res + row * D : nullptr;
    float s = 0.0f;
    for (int k = lane; k < D; k += 32) s += xr[k] + (rr ? rr[k] : 0.0f);
    const float mean = warp_sum(s) / D;
    float v = 0.0f;
    for (int k = lane; k < D; k += 32) {
        const float d = xr[k] + (rr ? rr[k] : 0.0f) - mean;
        v = fmaf(d, d, v);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(v) / D + eps);
    for (int k = lane; k < D; k += 32) {
        const float xh = (xr[k] + (rr ? rr[k] : 0.0f) - mean) * rstd;
        xhat[row * D + k] = xh;
        y[row * D + k] = fmaf(xh, w[k], b[k]);
    }
    if (lane == 0) rstd_out[row] = rstd;
}

// dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), dxhat = dy * w
// dw += sum_rows dy * xhat ; db += sum_rows dy      (per-CTA partials -> atomics)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(long long rows, int D, const float* __restrict__ xhat,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ w, const float* __restrict__ dy,
                                                            float* __restrict__ dx, float* __restrict__ dw,
                                                            float* __restrict__ db, int rows_per_cta) {
    extern __shared__ __align__(16) float sm[];  // [2][D]
    float* dw_s = sm;
    float* db_s = sm + D;
    for (int k = threadIdx.x; k < 2 * D; k += blockDim.x) sm[k] = 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = min(rows, r_begin + rows_per_cta);
    for (long long row = r_begin + warp; row < r_end; row += 8) {
        const float* xh = xhat + row * D;
        const float* g = dy + row * D;
        float s1 = 0.0f, s2 = 0.0f;
        for (int k = lane; k < D; k += 32) {
            const float d = g[k] * w[k];
            s1 += d;
            s2 = fmaf(d, xh[k], s2);
        }
        s1 = warp_sum(s1) / D;
        s2 = warp_sum(s2) / D;
        const float rs = rstd[row];
        for (int k = lane; k < D; k += 32) {
            const float d = g[k] * w[k];
            dx[row * D + k] = rs * (d - s1 - xh[k] * s2);
            atomicAdd(&dw_s[k], g[k] * xh[k]);
            atomicAdd(&db_s[k], g[k]);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        atomicAdd(dw + k, dw_s[k]);
        atomicAdd(db + k, db_s[k]);
    }
}

// ---- fused training loss ---------------------------------------------------------------
// out[0] = total, out[1] = prediction term, out[2] = consistency term (always reported, added to the
// total only when `consistency` != 0).  Every term is linear in the per-row sums, so the CTAs add their
// share to the zeroed 3-vector with fp32 atomics (dy, the gradient, does not depend on the reduction).
__global__ void __launch_bounds__(256) loss_kernel(int B, int T, const float* __restrict__ y,
                                                    const float* __restrict__ labels,
                                                    const uint8_t* __restrict__ mask, int consistency,
                                                    float* __restrict__ out, float* __restrict__ dy) {
    __shared__ float scratch[32];
    const long long n = (long long)B * T * 4;
    const float w_pred = 1.0f / (float)n;
    const long long pairs = (long long)B * (T - 1);
    const float w_cons = (consistency && pairs > 0) ? 0.5f / (float)pairs : 0.0f;
    float pred = 0.0f, cons = 0.0f;
    // one thread per (video, frame): 4 coordinates
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < (long long)B * T;
         r += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(r % T);
        const float4 yv = *reinterpret_cast<const float4*>(y + r * 4);
        const float4 lv = *reinterpret_cast<const float4*>(labels + r * 4);
        float yy[4] = {yv.x, yv.y, yv.z, yv.w};
        float ll[4] = {lv.x, lv.y, lv.z, lv.w};
        float g[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float d = yy[c] - ll[c];
            const float m = mask ? (float)mask[r * 4 + c] : 1.0f;
            pred += fabsf(d) * m;
            g[c] = ((d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f)) * m * w_pred;
        }
        // consistency: pairs (t-1,t) and (t,t+1) both touch frame t
        if (t + 1 < T) {
            const float4 nv = *reinterpret_cast<const float4*>(y + (r + 1) * 4);
            const float d0 = nv.x - yy[0], d1 = nv.y - yy[1], d2 = nv.z - yy[2], d3 = nv.w - yy[3];
            const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
            cons += nrm;
            if (nrm > 0.0f) {
                const float s = w_cons / nrm;
                g[0] -= d0 * s; g[1] -= d1 * s; g[2] -= d2 * s; g[3] -= d3 * s;
            }
        }
        if (t > 0) {
            const float4 pv = *reinterpret_cast<const float4*>(y + (r - 1) * 4);
            const float d0 = yy[0] - pv.x, d1 = yy[1] - pv.y, d2 = yy[2] - pv.z, d3 = yy[3] - pv.w;
            const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
            if (nrm > 0.0f) {
                const float s = w_cons / nrm;
                g[0] += d0 * s; g[1] += d1 * s; g[2] += d2 * s; g[3] += d3 * s;
            }
        }
        *reinterpret_cast<float4*>(dy + r * 4) = make_float4(g[0], g[1], g[2], g[3]);
    }
    pred = block_reduce(pred, false, scratch);
    cons = block_reduce(cons, false, scratch);
    if (threadIdx.x == 0) {
        const float pred_mean = pred * w_pred;
        const float cons_mean = pairs > 0 ? cons / (float)pairs : 0.0f;
        atomicAdd(out + 1, pred_mean);
        atomicAdd(out + 2, cons_mean);
        atomicAdd(out + 0, consistency ? pred_mean + 0.5f * cons_mean : pred_mean);
    }
}

inline unsigned grid_for(long long n, int per_block, int cap = 148 * 16) {
    long long g = (n + per_block - 1) / per_block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace
}  // namespace opn

using namespace opn;

extern "C" int opn_wtt_fwd(int64_t B, int64_t T, int64_t H1, const float* boxes, const float* hs1,
                           const float* w_pred, float* logits_bpt, float* probs, float* frames_boxes, void* stream) {
    OPN_CHECK_ARG(B > 0 && T > 0 && H1 > 0 && H1 <= 2048, "wtt_fwd: bad shape");
    OPN_CHECK_ARG(boxes && hs1 && w_pred && logits_bpt && probs && frames_boxes, "wtt_fwd: null pointer");
    const size_t smem = (size_t)(NOBJ * H1 + 8 * 96) * sizeof(float);
    OPN_CUDA(cudaFuncSetAttribute(wtt_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wtt_fwd_kernel<<<grid_for(B * T, 8, 148 * 4), 256, smem, as_stream(stream)>>>(boxes, hs1, w_pred, logits_bpt, probs,
                                                                                 frames_boxes, (int)B, (int)T, (int)H1);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_wtt_bwd(int64_t B, int64_t T, int64_t H1, const float* boxes, const float* probs,
                           const float* w_pred, const float* d_frames_boxes, const float* d_logits_bpt,
                           float* d_logits, float* d_hs1, void* stream) {
    OPN_CHECK_ARG(B > 0 && T > 0 && H1 > 0 && H1 <= 2048, "wtt_bwd: bad shape");
    OPN_CHECK_ARG(boxes && probs && w_pred && d_frames_boxes && d_logits && d_hs1, "wtt_bwd: null pointer");
    const size_t smem = (size_t)(NOBJ * H1 + 8 * 96) * sizeof(float);
    OPN_CUDA(cudaFuncSetAttribute(wtt_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wtt_bwd_kernel<<<grid_for(B * T, 8, 148 * 4), 256, smem, as_stream(stream)>>>(
        boxes, probs, w_pred, d_frames_boxes, d_logits_bpt, d_logits, d_hs1, (int)B, (int)T, (int)H1);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_relu_bwd(int64_t n, const float* y, float* dy, void* stream) {
    OPN_CHECK_ARG(n >= 0 && y && dy, "relu_bwd: bad argument");
    if (n == 0) return OPN_OK;
    relu_bwd_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(n, y, dy);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_add(int64_t n, const float* a, const float* b, float* out, void* stream) {
    OPN_CHECK_ARG(n >= 0 && a && b && out, "add: bad argument");
    if (n == 0) return OPN_OK;
    add_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(n, a, b, out);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_dropout(int64_t n, const float* x, float* out, float p, uint64_t seed, uint64_t offset, void* stream) {
    OPN_CHECK_ARG(n >= 0 && x && out, "dropout: bad argument");
    OPN_CHECK_ARG(p >= 0.0f && p < 1.0f, "dropout: p = %g outside [0, 1)", (double)p);
    if (n == 0) return OPN_OK;
    double t = (double)p * 4294967296.0;
    const unsigned int threshold = t >= 4294967295.0 ? 4294967295u : (unsigned int)t;
    const float scale = 1.0f / (1.0f - p);
    const int vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    dropout_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(stream)>>>(n, x, out, threshold, scale, seed, offset,
                                                                             vec);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_colsum(int64_t rows, int64_t cols, const float* x, int64_t ld, float* out, int accumulate,
                          void* stream) {
    OPN_CHECK_ARG(rows > 0 && cols > 0 && x && out, "colsum: bad argument");
    cudaStream_t s = as_stream(stream);
    if (!accumulate) {
        zero_kernel<<<grid_for(cols, 256), 256, 0, s>>>(cols, out);
        OPN_CUDA(cudaGetLastError());
        count_launch();
    }
    unsigned gy = (unsigned)((rows + 255) / 256);
    if (gy > 64) gy = 64;
    dim3 grid((unsigned)((cols + 31) / 32), gy);
    colsum_kernel<<<grid, 256, 0, s>>>(rows, (int)cols, x, ld, out);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_softmax_rows(int64_t rows, int64_t cols, float* x, int64_t ld, float scale, void* stream) {
    OPN_CHECK_ARG(rows > 0 && cols > 0 && x, "softmax_rows: bad argument");
    OPN_CHECK_ARG(cols * 4 <= 200 * 1024, "softmax_rows: row of %lld columns does not fit shared memory",
                  (long long)cols);
    const size_t smem = (size_t)cols * sizeof(float);
    OPN_CUDA(cudaFuncSetAttribute(softmax_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    softmax_rows_kernel<<<(unsigned)rows, 256, smem, as_stream(stream)>>>((int)cols, x, ld, scale);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_softmax_rows_bwd(int64_t rows, int64_t cols, const float* p, float* dp, int64_t ld, float scale,
                                    void* stream) {
    OPN_CHECK_ARG(rows > 0 && cols > 0 && p && dp, "softmax_rows_bwd: bad argument");
    softmax_rows_bwd_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>((int)cols, p, dp, ld, scale);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_layernorm_fwd(int64_t rows, int64_t D, const float* x, const float* res, const float* w,
                                 const float* b, float eps, float* y, float* xhat, float* rstd, void* stream) {
    OPN_CHECK_ARG(rows > 0 && D > 0 && x && w && b && y && xhat && rstd, "layernorm_fwd: bad argument");
    layernorm_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(rows, (int)D, x, res, w, b, eps, y,
                                                                                   xhat, rstd);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_layernorm_bwd(int64_t rows, int64_t D, const float* xhat, const float* rstd, const float* w,
                                 const float* dy, float* dx, float* dw, float* db, void* stream) {
    OPN_CHECK_ARG(rows > 0 && D > 0 && D <= 4096 && xhat && rstd && w && dy && dx && dw && db,
                  "layernorm_bwd: bad argument");
    const int rows_per_cta = 64;
    const size_t smem = (size_t)2 * D * sizeof(float);
    layernorm_bwd_kernel<<<(unsigned)((rows + rows_per_cta - 1) / rows_per_cta), 256, smem, as_stream(stream)>>>(
        rows, (int)D, xhat, rstd, w, dy, dx, dw, db, rows_per_cta);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_loss_fwd_bwd(int64_t B, int64_t T, const float* y, const float* labels, const uint8_t* mask,
                                int consistency, float* loss_out, float* dy, void* stream) {
    OPN_CHECK_ARG(B > 0 && T > 0 && y && labels && loss_out && dy, "loss_fwd_bwd: bad argument");
    OPN_CUDA(cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), as_stream(stream)));
    long long blocks = (B * T + 255) / 256;
    if (blocks > 148) blocks = 148;
    loss_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((int)B, (int)T, y, labels, mask, consistency, loss_out, dy);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}
