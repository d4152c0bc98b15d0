// Thread-local error string + ABI version for the retinanet_b200 C ABI.
#include <stdarg.h>

#include "rn_common.cuh"

static thread_local char g_err[512] = "";

void rn_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *rn_last_error(void) { return g_err; }
extern "C" int rn_abi_version(void) { return RN_ABI_VERSION; }
