// Error handling and device queries for the C ABI.
#include <stdarg.h>
#include <string.h>

#include "dig_common.cuh"

namespace dig {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace dig

extern "C" {

int dig_version(void) { return 100; }

const char *dig_last_error(void) { return dig::g_err; }

int dig_device_sm_count(void) { return dig::sm_count(); }

}
