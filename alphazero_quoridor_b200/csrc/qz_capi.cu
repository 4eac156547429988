// qz_capi.cu -- error plumbing + device queries of the C ABI (include/qzb200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "qz_common.cuh"

static thread_local char g_err[512] = "";

char *qz_err_buf() { return g_err; }

int qz_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int qz_check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return qz_fail((int)e, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    // QZ_SYNC_CHECK: wait for the device after every launch so that an asynchronous fault is reported by the call
    // that caused it (debugging aid; launches are on the caller's stream, a device-wide wait covers all of them)
    static const bool sync_check = getenv("QZ_SYNC_CHECK") != nullptr;
    if (sync_check) {
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return qz_fail((int)e, "%s (after the launch): %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    }
    return 0;
}

extern "C" int qz_stream_check(void *stream) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return qz_fail((int)e, "qz_stream_check: %s (%s)", cudaGetErrorString(e), cudaGetErrorName(e));
    return 0;
}

extern "C" int qz_version(void) { return QZ_ABI_VERSION; }

extern "C" const char *qz_last_error_string(void) { return g_err; }

extern "C" int qz_device_sm_count(int *out_sm_count) {
    QZ_REQUIRE_PTR(out_sm_count);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(out_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return qz_fail((int)e, "qz_device_sm_count: %s", cudaGetErrorString(e));
    return 0;
}
