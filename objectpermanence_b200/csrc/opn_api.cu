// Library-level entry points and host-side error plumbing of libopnet_b200.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>

#include "opn_common.cuh"

namespace opn {

static thread_local char g_error[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// Sticky status page (opn_set_status_page): when the caller registers one for a device, the persistent kernels report
// time-outs there instead of into the head of their (per-launch, memset) workspace, so the word survives until the host
// reads it where it synchronises anyway (training step, inference) and later launches bail out at once instead of each
// waiting out its own time-out.
static void* g_status_page[64] = {nullptr};

unsigned int* status_page_or(void* workspace_head) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && g_status_page[dev])
        return static_cast<unsigned int*>(g_status_page[dev]);
    return static_cast<unsigned int*>(workspace_head);
}

// arithmetic mode of the recurrence / contraction kernels (opn_set_precision)
static thread_local int g_precision = OPN_PRECISION_FP32;
int current_precision() { return g_precision; }

// The producer / consumer split (LSTM1 + head on the idle SMs, opn_opnet_l1head.cu) needs both launches co-resident:
// 32 + 4 CTAs per batch group, one per SM.  OPN_OPNET_SPLIT=0 keeps the single fused kernel.
bool opnet_split_wanted(int64_t B) {
    const char* e = getenv("OPN_OPNET_SPLIT");
    if (e && e[0] == '0') return false;
    // the consumer waits for a producer launched behind it: tools that serialise kernels (Nsight Compute replay,
    // compute-sanitizer, CUDA_LAUNCH_BLOCKING) would run it to its time-out
    const char* lb = getenv("CUDA_LAUNCH_BLOCKING");
    if ((lb && lb[0] == '1') || getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")) return false;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
    (void)B;
    return sms >= 37;      // 32 consumer + 4 unit + 1 head CTA per batch group, one per SM: at least one group per wave
}

// batch groups per wave of the split form: consumer and producer launches of a wave are co-resident
int opnet_split_groups_per_wave() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 1;
    return sms / 37 > 0 ? sms / 37 : 1;
}

SideStream* opnet_side_stream() {
    static std::mutex mu;
    static std::map<int, SideStream> per_dev;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    SideStream& s = per_dev[dev];
    if (!s.stream) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
            (void)cudaGetLastError();
            s.stream = nullptr;
            return nullptr;
        }
    }
    return &s;
}


int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return OPN_ERR_CUDA;
}

}  // namespace opn

extern "C" int opn_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* opn_last_error(void) { return opn::g_error; }

extern "C" unsigned long long opn_launch_count(void) { return opn::g_launch_count; }

extern "C" int opn_set_precision(int mode) {
    OPN_CHECK_ARG(mode == OPN_PRECISION_FP32 || mode == OPN_PRECISION_16BIT, "set_precision: unknown mode %d", mode);
    opn::g_precision = mode;
    return OPN_OK;
}

extern "C" int opn_get_precision(void) { return opn::g_precision; }

extern "C" int opn_set_status_page(void* device_page) {
    int dev = 0;
    OPN_CUDA(cudaGetDevice(&dev));
    OPN_CHECK_ARG(dev >= 0 && dev < 64, "set_status_page: device ordinal %d out of range", dev);
    opn::g_status_page[dev] = device_page;   // NULL: back to the workspace head
    return OPN_OK;
}

extern "C" int opn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    OPN_CUDA(cudaGetDevice(&dev));
    if (sm_count) OPN_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) OPN_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) OPN_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return OPN_OK;
}
