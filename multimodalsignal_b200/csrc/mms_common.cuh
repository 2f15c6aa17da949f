// Shared device/host helpers for libmms_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/mms_b200.h"

namespace mms {

// ---- error plumbing ------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define MMS_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) return ::mms::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// Launch bookkeeping: every kernel launch is preceded by MMS_PROF_BEGIN(stream) and followed by
// MMS_LAUNCH_CHECK(name).  Both are a counter increment unless mms_profile_enable(1) was called, in
// which case the launch is bracketed by CUDA events on its stream (bench.py's live kernel times).
void prof_begin(cudaStream_t st);
void prof_end(const char* name);

#define MMS_PROF_BEGIN(st) ::mms::prof_begin(st)

#define MMS_LAUNCH_CHECK(name)                                                      \
    do {                                                                            \
        cudaError_t _e = cudaPeekAtLastError();                                     \
        if (_e != cudaSuccess) return ::mms::cuda_fail(_e, name, __FILE__, __LINE__);  \
        ::mms::prof_end(name);                                                      \
    } while (0)

#define MMS_REQUIRE(cond, ...)                                                      \
    do {                                                                            \
        if (!(cond)) { ::mms::set_error(__VA_ARGS__); return MMS_E_INVALID; }       \
    } while (0)

// Integer run-time switches (A/B measurements, opt-in kernels): the value set by mms_set_option(name, v) wins, else the
// environment variable MMS_<name> read once, else `dflt`.
int option_get(const char* name, int dflt);

// One weight-gradient (TN) product, the argument list of launch_tc_gemm_tn / mms_gemm_tn_acc as a record, so that the
// products of a GRU layer can be handed to the batched tensor-core kernel in one call.
struct TnCall {
    const float* A; int64_t lda; int a_split, a_skip;
    const float* Bm; int64_t ldb; int shift, seq;
    float* C; int64_t ldc; float* bias_grad;
    int M, N1, N2;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: the opt-in guards at the launch
// sites remember it per device (several devices in one process: mms_init(dev), per-device side streams), not per process.
struct PerDeviceOnce {
    bool done[64] = {};
    bool need() {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- shapes of the reference CNN stack (models.py:46-53) ------------------------------
static inline int conv_out_len(int l_in, int k, int s, int p) { return (l_in + 2 * p - k) / s + 1; }
static inline int pool_out_len(int l_in) { return (l_in + 2 - 3) / 2 + 1; }

constexpr int CONV1_CO = 16, CONV1_K = 7, CONV1_S = 2, CONV1_P = 3;
constexpr int CONV2_CI = 16, CONV2_K = 5, CONV2_S = 2, CONV2_P = 2;
constexpr int HEAD_HID = 64;
constexpr float BN_EPS = 1e-5f;
constexpr float BN_MOMENTUM = 0.1f;

// BatchNorm backward folded into the consumers of a conv layer's output gradient (conv1d_dgrad / conv1d_wgrad): when
// y != nullptr the `dy` those kernels are given is the UN-normalised gradient dyn = d(relu/pool) and they form
//     dy = a * (dyn - mean(dyn) - xhat * mean(dyn * xhat)),   a = gamma / sqrt(var + eps),  xhat = (y - mean) / sqrt(var + eps)
// themselves while staging their tile (red = the two float64 reductions of pool_relu_bwd; eval mode: dy = a * dyn), which
// removes the separate BN-apply pass (one launch and one read + write of the whole gradient) from the critical path.
// dgamma / dbeta (optional) are added once, by the weight-gradient kernel.
struct BnBwd {
    const float* y;
    const double* stats;
    const float* gamma;
    const float* beta;
    const float* rm;
    const float* rv;
    const double* red;
    float* dgamma;
    float* dbeta;
    int Bstat;
    int training;
    float grad_scale;
};

// ---- programmatic dependent launch (build variant, -DMMS_PDL; default build: every macro below is a no-op) ---------
// 25 dependent launches make up the critical path of a step, and each edge costs the drain of the producer plus the launch,
// CTA scheduling and prologue of the consumer.  With MMS_PDL the kernels of the main chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: MMS_PDL_TRIGGER() (griddepcontrol.launch_dependents) at the top of a
// kernel lets the NEXT kernel of the stream start while this one runs, and that kernel executes MMS_PDL_WAIT()
// (griddepcontrol.wait: every prerequisite grid complete and its memory visible) before it touches anything an earlier
// kernel of the step wrote or still reads -- after its weight / invariant prologue where that prologue only reads
// parameters (constant within a step; the first kernel of a step is not launched this way).  Rule: a kernel may be passed to
// MMS_LAUNCH only if it executes MMS_PDL_WAIT() on every path before its first such access.  The first kernel of the
// forward and of the backward pass follows a memset node (full dependency) and chan_gate is a plain launch, so every kernel
// of a step starts after the previous step's Adam has completed: reading parameters ahead of the wait is safe.
// Build: MMS_NVCC_EXTRA="-DMMS_PDL" python -m multimodalsignal_b200.build --force   (parity-green on a B200 but slower than the default build in round 2: not the default)
#ifdef MMS_PDL
#define MMS_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define MMS_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#else
#define MMS_PDL_TRIGGER() ((void)0)
#define MMS_PDL_WAIT() ((void)0)
#endif
#define MMS_PDL_PROLOGUE() do { MMS_PDL_TRIGGER(); MMS_PDL_WAIT(); } while (0)

// ---- device helpers -----------------------------------------------------------------
#ifdef __CUDACC__

#ifdef MMS_PDL
template <typename... KArgs, typename... Args>
static inline void pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface in MMS_LAUNCH_CHECK (cudaPeekAtLastError)
}
#define MMS_LAUNCH(kern, grid, block, smem, st, ...) ::mms::pdl_launch(kern, grid, block, smem, st, __VA_ARGS__)
#else
#define MMS_LAUNCH(kern, grid, block, smem, st, ...) kern<<<grid, block, smem, st>>>(__VA_ARGS__)
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Counter-based dropout stream: murmur3 finaliser over (seed, offset, element index).
// Not bit-compatible with torch's Philox stream (nothing outside torch can be); parity
// tests run with p = 0 or in eval mode, exactly as SURVEY.md §7 hard part 4 prescribes.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
struct DropRng {
    uint32_t k0, k1;
    float p, scale;
    __device__ __forceinline__ void init(uint64_t seed, uint64_t offset, float p_) {
        k0 = fmix32((uint32_t)seed ^ 0x9e3779b9u) ^ fmix32((uint32_t)(offset) + 0x7f4a7c15u);
        k1 = fmix32((uint32_t)(seed >> 32) ^ 0x85ebca6bu) + fmix32((uint32_t)(offset >> 32) ^ 0xc2b2ae35u);
        p = p_;
        scale = p_ < 1.f ? 1.f / (1.f - p_) : 0.f;
    }
    // multiplier applied to element `idx`: 0 (dropped) or 1/(1-p)
    __device__ __forceinline__ float mult(uint64_t idx) const {
        uint32_t h = fmix32((uint32_t)idx ^ k0);
        h = fmix32(h + k1 + (uint32_t)(idx >> 32) * 0x27d4eb2fu);
        float u = (float)(h >> 8) * (1.0f / 16777216.0f);
        return u >= p ? scale : 0.f;
    }
};

__device__ __forceinline__ uint64_t resolve_offset(uint64_t host_off, const int64_t* dev_off) {
    return dev_off ? (uint64_t)(*dev_off) : host_off;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

#endif  // __CUDACC__

}  // namespace mms
