// Thread-local error string + ABI version for the retinanet_b200 C ABI.
#include <stdarg.h>

#include <atomic>

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

// Every kernel launch of the library passes through RN_CHECK_LAUNCH, which counts it here — during CUDA-graph capture
// too, where the count is the number of kernel nodes recorded.  bench.py reports its `gpu_launches` from this counter.
static std::atomic<unsigned long long> g_launches{0};
void rn_note_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t rn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
