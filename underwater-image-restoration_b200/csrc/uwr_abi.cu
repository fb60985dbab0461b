// Library-level plumbing of the C ABI: error string, version, device query.
#include "uwr_common.cuh"
#include "../../include/uwr_b200.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";
unsigned long long g_uwr_launches = 0;
int g_uwr_gemm_passes = 1;  // 1: TF32 (operands rounded where produced), 3: 3xTF32 (fp32 operands)

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

// programmatic dependent launch of the library's kernels (uwr_common.cuh): off unless UWR_PDL=1; uwr_set_pdl() overrides.
// Measured on B200: AST step 495.1 img/s off, 493.6 on; SpectralTransformer (2 256 launches of ~15 us) 226.5
// off, 231.2 on -- so it is an opt-in for launch-bound models, not the default.
static int g_uwr_pdl = -1;
int uwr_pdl_enabled() {
    if (g_uwr_pdl < 0) {
        const char* e = getenv("UWR_PDL");
        g_uwr_pdl = (e && e[0] == '1') ? 1 : 0;
    }
    return g_uwr_pdl;
}
extern "C" int uwr_set_pdl(int on) {
    g_uwr_pdl = on ? 1 : 0;
    return 0;
}

extern "C" const char* uwr_last_error(void) { return g_err; }
extern "C" int uwr_abi_version(void) { return 1; }
extern "C" int uwr_device_sm_count(void) { return uwr_sm_count(); }
extern "C" unsigned long long uwr_launch_count(void) { return g_uwr_launches; }
extern "C" int uwr_set_gemm_precision(int passes) {
    if (passes != 1 && passes != 3) {
        uwr_set_error("uwr_set_gemm_precision: passes must be 1 (tf32) or 3 (tf32x3)");
        return -1;
    }
    g_uwr_gemm_passes = passes;
    return 0;
}
extern "C" int uwr_get_gemm_precision(void) { return g_uwr_gemm_passes; }
