// Host-side bookkeeping of the C ABI: thread-local error text, launch counter.
#include "common.cuh"
#include <stdarg.h>

namespace xc {

static thread_local char g_err[512] = "";
static thread_local long g_launches = 0;

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += n; }

int sm_count()
{
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;              // B200; only reached without a device (size queries)
    }
    cached = n;
    return n;
}

}  // namespace xc

extern "C" const char* xc_last_error(void) { return xc::g_err; }
extern "C" int xc_abi_version(void) { return XC_ABI_VERSION; }
extern "C" long xc_launch_count(void) { return xc::g_launches; }
extern "C" void xc_reset_launch_count(void) { xc::g_launches = 0; }
