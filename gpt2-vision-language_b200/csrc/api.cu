// Library-level entry points: version, error text, device query, launch counter.
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace vlk {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int device_sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    int major = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    if (major != 10) return -1;  // built for sm_100a only; no other architecture is supported
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = sms;
    return sms;
}

}  // namespace vlk

extern "C" int vlk_version(void) { return VLK_VERSION; }
extern "C" const char* vlk_last_error_string(void) { return vlk::last_error_buf(); }
extern "C" int vlk_num_sms(void) {
    int n = vlk::device_sm_count();
    if (n <= 0) return vlk::set_error(VLK_ERR_ARCH, "vlk_num_sms: current device is not sm_100 (B200)");
    return n;
}
extern "C" long long vlk_launch_count(void) { return vlk::g_launches.load(std::memory_order_relaxed); }
