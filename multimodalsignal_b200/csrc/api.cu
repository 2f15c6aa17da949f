// Process-level entry points of the C ABI: error state, device check.
#include "mms_common.cuh"
#include <string.h>

namespace mms {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    return MMS_E_CUDA;
}

}  // namespace mms

using namespace mms;

extern "C" int mms_version(void) { return 100; }

extern "C" const char* mms_last_error(void) { return g_error; }

extern "C" int mms_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device is visible (%s); libmms_b200 has no CPU fallback", cudaGetErrorString(e));
        return MMS_E_ARCH;
    }
    MMS_REQUIRE(device >= 0 && device < count, "device %d outside [0,%d)", device, count);
    cudaDeviceProp prop;
    MMS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libmms_b200 is built for sm_100a (B200) only and has no fallback", device, prop.major,
                  prop.minor);
        return MMS_E_ARCH;
    }
    MMS_CUDA(cudaSetDevice(device));
    return MMS_OK;
}
