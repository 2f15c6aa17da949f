// ChannelAttention (reference models.py:7-31): squeeze (mean over T) -> Linear(C, C/4) -> ReLU
// -> Linear(C/4, C) -> Sigmoid -> scale.  HBM-streaming kernels: 128-bit loads, one pass for the
// squeeze; in the model the scale pass is folded into conv1's weight staging, so x is read once.
#include "mms_common.cuh"

namespace mms {

constexpr int CA_MAX_C = 32;
constexpr int CA_MAX_A = 8;

// grid = B, block = 256.  One CTA reduces the C rows of one batch element and evaluates the
// two tiny linears in shared memory.
__global__ void __launch_bounds__(256) chan_gate_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                        const float* __restrict__ w2, int C, int A, int T,
                                                        float* __restrict__ mean_out, float* __restrict__ gate_out) {
    MMS_PDL_PROLOGUE();
    __shared__ float s_mean[CA_MAX_C];
    __shared__ float s_hid[CA_MAX_A];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const float* xb = x + (size_t)b * C * T;
    const bool vec = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int c = warp; c < C; c += nwarp) {
        const float* row = xb + (size_t)c * T;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        if (vec) {
            const float4* r4 = reinterpret_cast<const float4*>(row);
            const int n4 = T >> 2;
            for (int i = lane; i < n4; i += 32) {
                float4 v = __ldg(r4 + i);
                acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w;
            }
        } else {
            for (int i = lane; i < T; i += 32) acc0 += __ldg(row + i);
        }
        float s = warp_sum((acc0 + acc1) + (acc2 + acc3));
        if (lane == 0) s_mean[c] = s / (float)T;
    }
    __syncthreads();
    if (threadIdx.x < A) {
        float h = 0.f;
        for (int c = 0; c < C; ++c) h += w1[threadIdx.x * C + c] * s_mean[c];
        s_hid[threadIdx.x] = fmaxf(h, 0.f);
    }
    __syncthreads();
    if (threadIdx.x < C) {
        float g = 0.f;
        for (int a = 0; a < A; ++a) g += w2[threadIdx.x * A + a] * s_hid[a];
        gate_out[b * C + threadIdx.x] = sigmoid_f(g);       // A == 0 -> sigmoid(0) = 0.5
        mean_out[b * C + threadIdx.x] = s_mean[threadIdx.x];
    }
}

// y[b,c,t] = x[b,c,t] * gate[b,c]      grid = (ceil(T/1024), B*C), block = 256
__global__ void __launch_bounds__(256) chan_scale_kernel(const float* __restrict__ x, const float* __restrict__ gate,
                                                         int T, float* __restrict__ y) {
    const int row = blockIdx.y;
    const float g = gate[row];
    const size_t base = (size_t)row * T;
    const int t0 = blockIdx.x * 1024 + threadIdx.x * 4;
    if ((T % 4 == 0) && t0 + 3 < T && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x + base + t0));
        v.x *= g; v.y *= g; v.z *= g; v.w *= g;
        *reinterpret_cast<float4*>(y + base + t0) = v;
    } else {
        for (int t = t0; t < min(t0 + 4, T); ++t) y[base + t] = x[base + t] * g;
    }
}

// dg[row] = sum_t dy[row,t] * x[row,t]     grid = B*C, block = 256
__global__ void __launch_bounds__(256) chan_dot_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       int T, float* __restrict__ dg) {
    __shared__ float s_part[8];
    const size_t base = (size_t)blockIdx.x * T;
    float acc = 0.f;
    for (int t = threadIdx.x; t < T; t += blockDim.x) acc += __ldg(x + base + t) * __ldg(dy + base + t);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) s += s_part[w];
        dg[blockIdx.x] = s;
    }
}

// Parameter gradients of the excite MLP from dg[b,c] = dL/dgate.  Single CTA (B*C*A is tiny).
// scratch: dpre2 [B,C] | hid [B,A] | dhid [B,A];  ds [B,C] (grad w.r.t. the channel means) is
// written when ds != nullptr.
__global__ void __launch_bounds__(256) chan_param_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ mean,
                                                             const float* __restrict__ gate, const float* __restrict__ w1,
                                                             const float* __restrict__ w2, int B, int C, int A,
                                                             float* __restrict__ scratch, float* __restrict__ ds,
                                                             float* __restrict__ dw1, float* __restrict__ dw2) {
    MMS_PDL_PROLOGUE();
    float* dpre2 = scratch;
    float* hid = scratch + (size_t)B * C;
    float* dhid = hid + (size_t)B * A;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float h[CA_MAX_A], dh[CA_MAX_A];
        for (int a = 0; a < A; ++a) {
            float s = 0.f;
            for (int c = 0; c < C; ++c) s += w1[a * C + c] * mean[b * C + c];
            h[a] = fmaxf(s, 0.f);
            dh[a] = 0.f;
        }
        for (int c = 0; c < C; ++c) {
            const float g = gate[b * C + c];
            const float dp = dg[b * C + c] * g * (1.f - g);
            dpre2[b * C + c] = dp;
            for (int a = 0; a < A; ++a) dh[a] += dp * w2[c * A + a];
        }
        for (int a = 0; a < A; ++a) {
            const float d = h[a] > 0.f ? dh[a] : 0.f;
            hid[b * A + a] = h[a];
            dhid[b * A + a] = d;
            dh[a] = d;
        }
        if (ds) {
            for (int c = 0; c < C; ++c) {
                float s = 0.f;
                for (int a = 0; a < A; ++a) s += dh[a] * w1[a * C + c];
                ds[b * C + c] = s;
            }
        }
    }
    __syncthreads();
    // thread <-> (c, a) element of dw2 and (a, c) element of dw1
    for (int e = threadIdx.x; e < C * A; e += blockDim.x) {
        const int c2 = e / A, a2 = e % A;        // dw2[c2, a2]
        const int a1 = e / C, c1 = e % C;        // dw1[a1, c1]
        float s2 = 0.f, s1 = 0.f;
        for (int b = 0; b < B; ++b) {
            s2 += dpre2[b * C + c2] * hid[b * A + a2];
            s1 += dhid[b * A + a1] * mean[b * C + c1];
        }
        dw2[e] += s2;
        dw1[e] += s1;
    }
}

// dx[b,c,t] = dy[b,c,t] * gate[b,c] + ds[b,c] / T
__global__ void __launch_bounds__(256) chan_dx_kernel(const float* __restrict__ dy, const float* __restrict__ gate,
                                                      const float* __restrict__ ds, int T, float* __restrict__ dx) {
    const int row = blockIdx.y;
    const float g = gate[row];
    const float add = ds ? ds[row] / (float)T : 0.f;
    const size_t base = (size_t)row * T;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x)
        dx[base + t] = dy[base + t] * g + add;
}

// ---- host launchers (used by the C ABI and by model.cu) -----------------------------------
int launch_chan_gate(const float* x, const float* w1, const float* w2, int B, int C, int T, float* mean_out,
                     float* gate_out, cudaStream_t st) {
    MMS_REQUIRE(C >= 1 && C <= CA_MAX_C, "chan_attn: in_channels %d outside [1,%d]", C, CA_MAX_C);
    const int A = C / 4;
    MMS_PROF_BEGIN(st);
    // plain launch on purpose (never MMS_LAUNCH): the first kernel of a step must depend on the previous step's Adam in full,
    // so that no kernel of this step can start -- and read parameters ahead of its MMS_PDL_WAIT -- while Adam still writes them
    chan_gate_kernel<<<B, 256, 0, st>>>(x, w1, w2, C, A, T, mean_out, gate_out);
    MMS_LAUNCH_CHECK("chan_gate_kernel");
    return MMS_OK;
}

int launch_chan_param_bwd(const float* dg, const float* mean, const float* gate, const float* w1, const float* w2,
                          int B, int C, float* scratch, float* ds, float* dw1, float* dw2, cudaStream_t st) {
    const int A = C / 4;
    if (A == 0) return MMS_OK;          // no parameters
    MMS_PROF_BEGIN(st);
    MMS_LAUNCH(chan_param_bwd_kernel, dim3(1), dim3(256), 0, st, dg, mean, gate, w1, w2, B, C, A, scratch, ds, dw1, dw2);
    MMS_LAUNCH_CHECK("chan_param_bwd_kernel");
    return MMS_OK;
}

int launch_chan_dx(const float* dy, const float* gate, const float* ds, int B, int C, int T, float* dx, cudaStream_t st) {
    dim3 grid(cdiv(T, 1024), B * C);
    MMS_PROF_BEGIN(st);
    chan_dx_kernel<<<grid, 256, 0, st>>>(dy, gate, ds, T, dx);
    MMS_LAUNCH_CHECK("chan_dx_kernel");
    return MMS_OK;
}

}  // namespace mms

using namespace mms;

extern "C" int mms_chan_attn_fwd(const float* x, const float* w1, const float* w2, int32_t B, int32_t C, int32_t T,
                                 float* mean_out, float* gate_out, float* y, mms_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MMS_REQUIRE(x && mean_out && gate_out && B > 0 && T > 0, "chan_attn_fwd: bad arguments");
    int rc = launch_chan_gate(x, w1, w2, B, C, T, mean_out, gate_out, st);
    if (rc) return rc;
    if (y) {
        dim3 grid(cdiv(T, 1024), B * C);
        MMS_PROF_BEGIN(st);
        chan_scale_kernel<<<grid, 256, 0, st>>>(x, gate_out, T, y);
        MMS_LAUNCH_CHECK("chan_scale_kernel");
    }
    return MMS_OK;
}

extern "C" int mms_chan_attn_bwd(const float* x, const float* dy, const float* w1, const float* w2, const float* mean,
                                 const float* gate, int32_t B, int32_t C, int32_t T, float* dx, float* dw1, float* dw2,
                                 float* scratch, mms_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MMS_REQUIRE(x && dy && mean && gate && scratch && B > 0 && C >= 1 && C <= CA_MAX_C, "chan_attn_bwd: bad arguments");
    // scratch: dg [B,C] | ds [B,C] | dpre2 [B,C] | hid,dhid [2*B*A]   (<= 4*B*C floats)
    float* dg = scratch;
    float* ds = scratch + (size_t)B * C;
    float* rest = ds + (size_t)B * C;
    const int A = C / 4;
    MMS_PROF_BEGIN(st);
    chan_dot_kernel<<<B * C, 256, 0, st>>>(x, dy, T, dg);
    MMS_LAUNCH_CHECK("chan_dot_kernel");
    int rc = launch_chan_param_bwd(dg, mean, gate, w1, w2, B, C, rest, ds, dw1, dw2, st);
    if (rc) return rc;
    if (dx) return launch_chan_dx(dy, gate, A > 0 ? ds : nullptr, B, C, T, dx, st);
    return MMS_OK;
}
