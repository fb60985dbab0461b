// Library-level plumbing of the C ABI: error string, version, device query.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"
#include <stdarg.h>
#include <stdio.h>

static thread_local char g_err[512] = "";
unsigned long long g_uwr_launches = 0;

void uwr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int uwr_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;  // B200
    }
    return sms;
}

extern "C" const char* uwr_last_error(void) { return g_err; }
extern "C" int uwr_abi_version(void) { return 1; }
extern "C" int uwr_device_sm_count(void) { return uwr_sm_count(); }
extern "C" unsigned long long uwr_launch_count(void) { return g_uwr_launches; }
