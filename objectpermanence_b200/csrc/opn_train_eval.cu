// The two per-step / per-epoch pieces of the reference training loop that sit right behind the hot path
// (SURVEY 8f rows 1 and 3):
//   * opn_adam_step   -- torch.optim.Adam(model.parameters(), lr) (baselines/training_main.py:150,217) as ONE launch
//                        over a flat parameter / gradient buffer (HBM bound: 28 bytes per parameter).
//   * opn_iou_eval    -- the IoU evaluation of inference_and_iou_comp (baselines/training_main.py:97-112) with the
//                        reference's integer semantics: (x * [320,240,320,240]) in double -> truncation to int32 ->
//                        per-frame IoU with the +1 pixel convention (baselines/tracking_utils.py:138-159) -> per-video
//                        mean and mean over the masked ("containment") frames.  Replaces output.cpu().numpy() of
//                        the whole prediction tensor by 20 bytes per video.
#include "opn_common.cuh"

namespace opn {
namespace {

__global__ void __launch_bounds__(256) adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float lr, float beta1,
                                                   float beta2, float eps, float bias1, float inv_sqrt_bias2,
                                                   float weight_decay) {
    // torch.optim.Adam (amsgrad = False, maximize = False):
    //   g += wd * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / bias1) * m / (sqrt(v) / sqrt(bias2) + eps)
    const float step_size = lr / bias1;
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4*>(p)[i];
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        float4 mv = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pp = &pv.x;
        const float* gp = &gv.x;
        float* mp = &mv.x;
        float* vp = &vv.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float gg = fmaf(weight_decay, pp[c], gp[c]);
            mp[c] = fmaf(beta1, mp[c], (1.0f - beta1) * gg);
            vp[c] = fmaf(beta2, vp[c], (1.0f - beta2) * gg * gg);
            const float denom = sqrtf(vp[c]) * inv_sqrt_bias2 + eps;
            pp[c] -= step_size * (mp[c] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pv;
        reinterpret_cast<float4*>(m)[i] = mv;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float gg = fmaf(weight_decay, p[i], g[i]);
        const float mm = fmaf(beta1, m[i], (1.0f - beta1) * gg);
        const float vv = fmaf(beta2, v[i], (1.0f - beta2) * gg * gg);
        m[i] = mm;
        v[i] = vv;
        p[i] -= step_size * (mm / (sqrtf(vv) * inv_sqrt_bias2 + eps));
    }
}

// One CTA per video; thread t handles frames t, t+128, ...; fixed-order tree reduction in double (deterministic).
__global__ void __launch_bounds__(128) iou_eval_kernel(int T, const float* __restrict__ y, const float* __restrict__ labels,
                                                       const uint8_t* __restrict__ mask, double* __restrict__ video_mean,
                                                       double* __restrict__ masked_mean, int* __restrict__ masked_frames,
                                                       double* __restrict__ frame_iou) {
    __shared__ double s_all[128], s_msk[128];
    __shared__ int s_cnt[128];
    const int n = blockIdx.x, tid = threadIdx.x;
    const double shape[4] = {320.0, 240.0, 320.0, 240.0};
    double sum_all = 0.0, sum_msk = 0.0;
    int cnt = 0;
    for (int t = tid; t < T; t += 128) {
        const size_t r = (size_t)n * T + t;
        const float4 pv = *reinterpret_cast<const float4*>(y + r * 4);
        const float4 lv = *reinterpret_cast<const float4*>(labels + r * 4);
        const float pf[4] = {pv.x, pv.y, pv.z, pv.w}, lf[4] = {lv.x, lv.y, lv.z, lv.w};
        long long a[4], b[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            // numpy: float32 array * int64 array -> float64 product, .astype(np.int32) truncates toward zero
            a[c] = (long long)(int)((double)pf[c] * shape[c]);
            b[c] = (long long)(int)((double)lf[c] * shape[c]);
        }
        const long long xa = max(a[0], b[0]), ya = max(a[1], b[1]), xb = min(a[2], b[2]), yb = min(a[3], b[3]);
        const long long inter = max(xb - xa + 1, 0LL) * max(yb - ya + 1, 0LL);
        const long long area_a = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
        const long long area_b = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
        const double iou = (double)inter / (double)(area_a + area_b - inter);   // 0/0 -> nan, x/0 -> inf as in numpy
        if (frame_iou) frame_iou[r] = iou;
        sum_all += iou;
        if (mask) {
            // containment frame = any of the 4 mask entries set (torch.sum(mask, -1).type(bool), training_main.py:88)
            const uchar4 mk = *reinterpret_cast<const uchar4*>(mask + r * 4);
            if (mk.x | mk.y | mk.z | mk.w) {
                sum_msk += iou;
                ++cnt;
            }
        }
    }
    s_all[tid] = sum_all;
    s_msk[tid] = sum_msk;
    s_cnt[tid] = cnt;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o) {
            s_all[tid] += s_all[tid + o];
            s_msk[tid] += s_msk[tid + o];
            s_cnt[tid] += s_cnt[tid + o];
        }
        __syncthreads();
    }
    if (tid == 0) {
        video_mean[n] = s_all[0] / (double)T;
        if (masked_mean) masked_mean[n] = s_cnt[0] > 0 ? s_msk[0] / (double)s_cnt[0] : __longlong_as_double(0x7ff8000000000000LL);
        if (masked_frames) masked_frames[n] = s_cnt[0];
    }
}

}  // namespace
}  // namespace opn

using namespace opn;

extern "C" int opn_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream) {
    OPN_CHECK_ARG(n > 0 && params && grads && exp_avg && exp_avg_sq && step >= 1, "adam_step: bad argument");
    OPN_CHECK_ARG((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                  "adam_step: buffers must be 16-byte aligned");
    const double bias1 = 1.0 - pow((double)beta1, (double)step);
    const double bias2 = 1.0 - pow((double)beta2, (double)step);
    long long blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((long long)n, params, grads, exp_avg, exp_avg_sq, lr, beta1,
                                                                 beta2, eps, (float)bias1, (float)(1.0 / sqrt(bias2)),
                                                                 weight_decay);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

extern "C" int opn_iou_eval(int64_t N, int64_t T, const float* y, const float* labels, const uint8_t* mask,
                            double* video_mean, double* masked_mean, int32_t* masked_frames, double* frame_iou,
                            void* stream) {
    OPN_CHECK_ARG(N > 0 && T > 0 && y && labels && video_mean, "iou_eval: bad argument");
    OPN_CHECK_ARG(mask || (!masked_mean && !masked_frames), "iou_eval: masked outputs need a mask");
    iou_eval_kernel<<<(unsigned)N, 128, 0, as_stream(stream)>>>((int)T, y, labels, mask, video_mean, masked_mean, masked_frames,
                                                                frame_iou);
    OPN_CUDA(cudaGetLastError());
    count_launch();
    return OPN_OK;
}

// ---- pixel boxes for the inference writer (baselines/inference_main.py:214-215) -------------------------------
namespace opn {
namespace {
__global__ void __launch_bounds__(256) to_pixels_kernel(long long rows, const float* __restrict__ x, int* __restrict__ out) {
    const double shape[4] = {320.0, 240.0, 320.0, 240.0};
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const float4 v = *reinterpret_cast<const float4*>(x + r * 4);
        int4 o;
        // numpy: float32 array * int64 array -> float64, .astype(np.int32) truncates toward zero
        o.x = (int)((double)v.x * shape[0]);
        o.y = (int)((double)v.y * shape[1]);
        o.z = (int)((double)v.z * shape[2]);
        o.w = (int)((double)v.w * shape[3]);
        *reinterpret_cast<int4*>(out + r * 4) = o;
    }
}
}  // namespace
}  // namespace opn

extern "C" int opn_to_pixels(int64_t rows, const float* boxes, int32_t* pixels, void* stream) {
    OPN_CHECK_ARG(rows > 0 && boxes && pixels, "to_pixels: bad argument");
    long long blocks = (rows + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    opn::to_pixels_kernel<<<(unsigned)blocks, 256, 0, opn::as_stream(stream)>>>((long long)rows, boxes, pixels);
    OPN_CUDA(cudaGetLastError());
    opn::count_launch();
    return OPN_OK;
}
