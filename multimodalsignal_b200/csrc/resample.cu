// FFT resampling (reference preprocess.py:70-75 -> scipy.signal.resample).  Placeholder for the
// first bring-up run; replaced by the Bluestein/stockham implementation.
#include "mms_common.cuh"
using namespace mms;
extern "C" int64_t mms_resample_workspace_bytes(int64_t n_in, int64_t n_out, int32_t n_sig) { return -1; }
extern "C" int mms_resample_f64(const double* x, int64_t n_in, int64_t n_out, int32_t n_sig, double* y, void* workspace,
                                int64_t workspace_bytes, mms_stream_t stream) {
    MMS_REQUIRE(false, "resample_f64: not built yet");
}
