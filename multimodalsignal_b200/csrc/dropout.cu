// Inter-layer GRU dropout (reference models.py:62, nn.GRU(dropout=0.5) between layers in train mode)
// as one vectorised streaming pass: out[i] = in[i] * m(i), m(i) in {0, 1/(1-p)} from the counter-based
// stream of mms_common.cuh.  The backward applies the same multipliers to the incoming gradient.
#include "mms_common.cuh"

namespace mms {

__global__ void __launch_bounds__(256) dropout_apply_kernel(const float* in, float* out, int64_t n,   // in == out allowed
                                                            int64_t base_id, float p, uint64_t seed, uint64_t offset,
                                                            const int64_t* offset_dev) {
    MMS_PDL_PROLOGUE();
    DropRng rng;
    rng.init(seed, resolve_offset(offset, offset_dev), p);
    const int64_t n4 = n >> 2;
    const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = reinterpret_cast<const float4*>(in)[i];
            const uint64_t e = (uint64_t)(base_id + 4 * i);
            v.x *= rng.mult(e); v.y *= rng.mult(e + 1); v.z *= rng.mult(e + 2); v.w *= rng.mult(e + 3);
            reinterpret_cast<float4*>(out)[i] = v;
        }
        for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            out[i] = in[i] * rng.mult((uint64_t)(base_id + i));
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            out[i] = in[i] * rng.mult((uint64_t)(base_id + i));
    }
}

// rows of `row_len` contiguous floats whose element ids advance by id_row_stride per row (a strided slice
// of a larger tensor, e.g. the t = L-1 rows of [B, L, 2H]); in == out allowed.
__global__ void __launch_bounds__(256) dropout_rows_kernel(const float* in, float* out, int rows, int row_len, int64_t base_id,
                                                           int64_t id_row_stride, float p, uint64_t seed, uint64_t offset,
                                                           const int64_t* offset_dev) {
    DropRng rng;
    rng.init(seed, resolve_offset(offset, offset_dev), p);
    const int64_t n = (int64_t)rows * row_len;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / row_len, c = i - r * row_len;
        out[i] = in[i] * rng.mult((uint64_t)(base_id + r * id_row_stride + c));
    }
}

int launch_dropout_rows(const float* in, float* out, int rows, int row_len, int64_t base_id, int64_t id_row_stride, float p,
                        uint64_t seed, uint64_t offset, const int64_t* offset_dev, cudaStream_t st) {
    if (rows <= 0 || row_len <= 0) return MMS_OK;
    const int blocks = cdiv((int64_t)rows * row_len, 256);
    MMS_PROF_BEGIN(st);
    dropout_rows_kernel<<<blocks > 592 ? 592 : blocks, 256, 0, st>>>(in, out, rows, row_len, base_id, id_row_stride, p, seed, offset, offset_dev);
    MMS_LAUNCH_CHECK("dropout_rows_kernel");
    return MMS_OK;
}

int launch_dropout_apply(const float* in, float* out, int64_t n, int64_t base_id, float p, uint64_t seed, uint64_t offset,
                         const int64_t* offset_dev, cudaStream_t st) {
    if (n <= 0) return MMS_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    MMS_PROF_BEGIN(st);
    MMS_LAUNCH(dropout_apply_kernel, dim3((int)blocks), dim3(256), 0, st, in, out, n, base_id, p, seed, offset, offset_dev);
    MMS_LAUNCH_CHECK("dropout_apply_kernel");
    return MMS_OK;
}

}  // namespace mms

using namespace mms;

extern "C" int mms_dropout_apply(const float* in, float* out, int64_t n, int64_t base_id, float dropout_p, uint64_t rng_seed,
                                 uint64_t rng_offset, const int64_t* rng_offset_dev, mms_stream_t stream) {
    MMS_REQUIRE(in && out && n >= 0 && dropout_p >= 0.f && dropout_p < 1.f, "dropout_apply: bad arguments");
    return launch_dropout_apply(in, out, n, base_id, dropout_p, rng_seed, rng_offset, rng_offset_dev, (cudaStream_t)stream);
}
