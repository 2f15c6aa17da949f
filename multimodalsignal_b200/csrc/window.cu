// Window stacking (reference preprocess.py:189-200) and the per-subject normalisation statistics of
// dataset.py:37-48, as coalesced HBM-streaming kernels.  Window start indices are computed on the
// host in float64 exactly as the reference does (bit-exact) and passed in.
#include "mms_common.cuh"

namespace mms {

constexpr int WIN_MAX_CH = 32;

struct StreamPtrs { const double* p[WIN_MAX_CH]; };

// float64 [n_win, win, n_ch] (the .npy layout of preprocess.py:218).  One CTA per (chunk, window):
// reads are contiguous per channel, the interleave goes through shared memory so that the
// stores are contiguous too.       grid = (ceil(win/128), n_win), block = 256
__global__ void __launch_bounds__(256) window_gather_f64_kernel(StreamPtrs s, int n_ch, int64_t stream_len,
                                                                const int64_t* __restrict__ starts, int win,
                                                                double* __restrict__ out) {
    extern __shared__ double tile[];          // [128][n_ch]
    const int w = blockIdx.y, i0 = blockIdx.x * 128;
    const int64_t st = starts[w];
    const int rows = min(128, win - i0);
    for (int idx = threadIdx.x; idx < n_ch * 128; idx += 256) {
        const int c = idx >> 7, i = idx & 127;
        if (i < rows) {
            const int64_t g = st + i0 + i;
            tile[i * n_ch + c] = (g >= 0 && g < stream_len) ? s.p[c][g] : 0.0;
        }
    }
    __syncthreads();
    double* o = out + ((int64_t)w * win + i0) * n_ch;
    for (int idx = threadIdx.x; idx < rows * n_ch; idx += 256) o[idx] = tile[idx];
}

// float32 [n_win, n_ch, win] = what WesadDataset.__getitem__ yields (dataset.py:62-65: normalised,
// cast to float32, permuted to [C, W]) -- written directly, so the 6x-expanded float64 window
// array never exists.               grid = (ceil(win/1024), n_ch, n_win), block = 256
__global__ void __launch_bounds__(256) window_gather_f32_kernel(StreamPtrs s, int n_ch, int64_t stream_len,
                                                                const int64_t* __restrict__ starts, int win,
                                                                const double* __restrict__ shift,
                                                                const double* __restrict__ scale,
                                                                const int32_t* __restrict__ log_flag,
                                                                float* __restrict__ out) {
    const int w = blockIdx.z, c = blockIdx.y;
    const int64_t st = starts[w];
    const double sh = shift ? shift[c] : 0.0, sc = scale ? scale[c] : 1.0;
    const bool lg = log_flag && log_flag[c];
    const double* src = s.p[c];
    float* o = out + ((int64_t)w * n_ch + c) * win;
    for (int i = blockIdx.x * 1024 + threadIdx.x; i < min(win, (int)(blockIdx.x + 1) * 1024); i += 256) {
        const int64_t g = st + i;
        double v = (g >= 0 && g < stream_len) ? src[g] : 0.0;
        if (lg) v = log1p(v);
        o[i] = (float)((v - sh) * sc);
    }
}

// Overlap-weighted sums: every stream sample counts once per window that covers it.
// grid = (ceil(win/1024), n_ch, n_win), block = 256; float64 atomics.
__global__ void __launch_bounds__(256) window_stats_kernel(StreamPtrs s, int n_ch, int64_t stream_len,
                                                           const int64_t* __restrict__ starts, int win,
                                                           const int32_t* __restrict__ log_flag,
                                                           double* __restrict__ sums) {
    __shared__ double part[8][2];
    const int w = blockIdx.z, c = blockIdx.y;
    const int64_t st = starts[w];
    const bool lg = log_flag && log_flag[c];
    const double* src = s.p[c];
    double a = 0.0, q = 0.0;
    for (int i = blockIdx.x * 1024 + threadIdx.x; i < min(win, (int)(blockIdx.x + 1) * 1024); i += 256) {
        const int64_t g = st + i;
        double v = (g >= 0 && g < stream_len) ? src[g] : 0.0;
        if (lg) v = log1p(v);
        a += v;
        q += v * v;
    }
    a = warp_sum(a);
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { part[threadIdx.x >> 5][0] = a; part[threadIdx.x >> 5][1] = q; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
        atomicAdd(sums + c * 2 + threadIdx.x, t);
    }
}

}  // namespace mms

using namespace mms;

static int fill_ptrs(const double* const* streams_dev_or_host, int n_ch, StreamPtrs* sp) {
    // `streams` is a HOST array of device pointers (n_ch is tiny); it is passed by value in the
    // kernel parameter block so no extra device allocation or copy is needed.
    MMS_REQUIRE(n_ch >= 1 && n_ch <= WIN_MAX_CH, "window: channel count %d outside [1,%d]", n_ch, WIN_MAX_CH);
    for (int c = 0; c < n_ch; ++c) {
        MMS_REQUIRE(streams_dev_or_host[c], "window: null stream pointer");
        sp->p[c] = streams_dev_or_host[c];
    }
    for (int c = n_ch; c < WIN_MAX_CH; ++c) sp->p[c] = nullptr;
    return MMS_OK;
}

extern "C" int mms_window_gather(const double* const* streams, int32_t n_ch, int64_t stream_len, const int64_t* starts,
                                 int32_t n_win, int32_t win, int32_t out_f32, const double* shift, const double* scale,
                                 const int32_t* log_flag, void* out, mms_stream_t stream) {
    MMS_REQUIRE(streams && starts && out && win > 0 && n_win >= 0, "window_gather: bad arguments");
    if (n_win == 0) return MMS_OK;
    StreamPtrs sp;
    int rc = fill_ptrs(streams, n_ch, &sp);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_f32) {
        dim3 grid(cdiv(win, 1024), n_ch, n_win);
        MMS_PROF_BEGIN(st);
        window_gather_f32_kernel<<<grid, 256, 0, st>>>(sp, n_ch, stream_len, starts, win, shift, scale, log_flag, (float*)out);
    } else {
        dim3 grid(cdiv(win, 128), n_win);
        MMS_PROF_BEGIN(st);
        window_gather_f64_kernel<<<grid, 256, (size_t)128 * n_ch * sizeof(double), st>>>(sp, n_ch, stream_len, starts, win, (double*)out);
    }
    MMS_LAUNCH_CHECK("window_gather");
    return MMS_OK;
}

extern "C" int mms_window_stats(const double* const* streams, int32_t n_ch, int64_t stream_len, const int64_t* starts,
                                int32_t n_win, int32_t win, const int32_t* log_flag, double* sums, mms_stream_t stream) {
    MMS_REQUIRE(streams && starts && sums && win > 0 && n_win >= 0, "window_stats: bad arguments");
    if (n_win == 0) return MMS_OK;
    StreamPtrs sp;
    int rc = fill_ptrs(streams, n_ch, &sp);
    if (rc) return rc;
    dim3 grid(cdiv(win, 1024), n_ch, n_win);
    MMS_PROF_BEGIN((cudaStream_t)stream);
    window_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sp, n_ch, stream_len, starts, win, log_flag, sums);
    MMS_LAUNCH_CHECK("window_stats_kernel");
    return MMS_OK;
}
