// Receive buffers of the image-sharded loss exchange (include/retinanet_b200.h: rn_exchange_t, rn_comm_*).
// The exchange itself lives in loss.cu (finalize_image): peers store into each other's buffers over NVLink.
// These helpers only allocate the buffer with cudaMalloc (a torch caching-allocator block cannot be exported
// reliably) and move CUDA IPC handles; the Python side ships the 64-byte handles with torch.distributed.
#include <cstring>

#include "rn_common.cuh"

namespace {
constexpr size_t COMM_BYTES = 2048;   // 2 parities x RN_MAX_PEERS senders x 4 words x 8 B = 1 KB, + seq / error words
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI documents 64-byte handles");
}  // namespace

#define RN_CUDA_TRY(expr, what)                                                   \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            rn_set_error("%s: %s", what, cudaGetErrorString(e__));                \
            (void)cudaGetLastError();                                             \
            return (int)e__;                                                      \
        }                                                                         \
    } while (0)

extern "C" size_t rn_comm_bytes(void) { return COMM_BYTES; }

extern "C" int rn_comm_alloc(void **out_ptr) {
    RN_CHECK_ARG(out_ptr, RN_E_BADARG, "rn_comm_alloc: null pointer");
    void *p = nullptr;
    RN_CUDA_TRY(cudaMalloc(&p, COMM_BYTES), "rn_comm_alloc: cudaMalloc");
    RN_CUDA_TRY(cudaMemset(p, 0, COMM_BYTES), "rn_comm_alloc: cudaMemset");
    RN_CUDA_TRY(cudaDeviceSynchronize(), "rn_comm_alloc: synchronize");
    *out_ptr = p;
    return 0;
}

extern "C" int rn_comm_free(void *ptr) {
    if (ptr) RN_CUDA_TRY(cudaFree(ptr), "rn_comm_free");
    return 0;
}

extern "C" int rn_comm_export(void *ptr, void *handle_out_host) {
    RN_CHECK_ARG(ptr && handle_out_host, RN_E_BADARG, "rn_comm_export: null pointer");
    RN_CUDA_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out_host, ptr), "rn_comm_export: cudaIpcGetMemHandle");
    return 0;
}

extern "C" int rn_comm_import(const void *handle_host, void **out_ptr) {
    RN_CHECK_ARG(handle_host && out_ptr, RN_E_BADARG, "rn_comm_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    void *p = nullptr;
    RN_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "rn_comm_import: cudaIpcOpenMemHandle");
    *out_ptr = p;
    return 0;
}

extern "C" int rn_comm_unmap(void *imported_ptr) {
    if (imported_ptr) RN_CUDA_TRY(cudaIpcCloseMemHandle(imported_ptr), "rn_comm_unmap");
    return 0;
}

extern "C" int rn_comm_error(const void *local_buf, int32_t *out_host) {
    RN_CHECK_ARG(local_buf && out_host, RN_E_BADARG, "rn_comm_error: null pointer");
    RN_CUDA_TRY(cudaMemcpy(out_host, (const char *)local_buf + 1024 + 4, sizeof(int32_t), cudaMemcpyDeviceToHost),
                "rn_comm_error: cudaMemcpy");
    return 0;
}
