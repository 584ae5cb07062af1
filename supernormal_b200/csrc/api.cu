// api.cu -- error reporting and version of the C ABI (include/snb200.h).
#include <stdarg.h>
#include "common.cuh"

namespace snb {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace snb

extern "C" const char *snb_last_error(void) { return snb::g_err; }
extern "C" int32_t snb_version(void) { return 100; }

// Asynchronous copy on a caller-chosen stream (kind: 1 host->device, 2 device->host, 3 device->device; host memory must be pinned for the
// copy to be asynchronous).  What a C++ loader calls directly; the Python feeder uses it instead of Tensor.copy_ under a stream context
// manager, which costs ~30 us of interpreter time per step.
extern "C" int32_t snb_copy_async(void *dst, const void *src, int64_t bytes, int32_t kind, snb_stream_t stream) {
    SNB_REQUIRE(bytes >= 0 && kind >= 1 && kind <= 3, SNB_ERR_ARG, "copy_async: bad size / kind");
    if (bytes == 0) return SNB_OK;
    SNB_REQUIRE(dst && src, SNB_ERR_NULL, "copy_async: null pointer");
    const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    const cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, k, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) { snb::set_error("copy_async: %s", cudaGetErrorString(e)); return SNB_ERR_LAUNCH; }
    return SNB_OK;
}
