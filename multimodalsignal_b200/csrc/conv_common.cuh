// BatchNorm helpers shared by the CNN-encoder kernels (conv_bn_pool.cu, conv_fused.cu).  Forward and backward kernels MUST
// form z = relu(fmaf(a, y, b)) from the SAME (a, b): the max-pool backward recomputes the arg-max from z, and ties are
// only routed consistently when both passes see bit-identical values.
#pragma once
#include "mms_common.cuh"

namespace mms {

// cp.async copies with zero fill (src-size 0 when !ok): the source address must still be valid, callers clamp it
__device__ __forceinline__ void cp_async4_zfill(float* smem_dst, const float* gsrc, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(ok ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(float* smem_dst, const float* gsrc, bool ok) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Contiguous global -> shared copy of n floats (n % 4 == 0, both 16-byte aligned) by the whole CTA, every copy in flight at once
// (cp.async; the caller commits / waits).  Tiles that a loop of "load, then store" would fetch one L2 latency at a time
// arrive in one: ncu had 40-80 % of the stall samples of the first version of the fused kernels on exactly those stores.
__device__ __forceinline__ void cp_async_floats(float* smem_dst, const float* gsrc, int n, int tid, int nthreads) {
    for (int i = 4 * tid; i < n; i += 4 * nthreads) cp_async16_zfill(smem_dst + i, gsrc + i, true);
}
__device__ __forceinline__ void cp_async_commit_only() { asm volatile("cp.async.commit_group;" ::: "memory"); }

struct BnAffine { float a, b, mean, inv; };

__device__ __forceinline__ BnAffine bn_affine(int training, const double* stats, const float* gamma, const float* beta,
                                              const float* rm, const float* rv, int c, int C, double n) {
    double mean, var;
    if (training) {
        mean = stats[c] / n;
        var = stats[C + c] / n - mean * mean;
        if (var < 0.0) var = 0.0;
    } else {
        mean = (double)rm[c];
        var = (double)rv[c];
    }
    const double inv = 1.0 / sqrt(var + (double)BN_EPS);
    BnAffine r;
    r.inv = (float)inv;
    r.mean = (float)mean;
    r.a = gamma[c] * r.inv;
    r.b = beta[c] - r.mean * r.a;
    return r;
}

// Block-uniform version: thread 0 does the float64 arithmetic once, everybody reads shared memory.
__device__ __forceinline__ BnAffine bn_affine_block(int training, const double* stats, const float* gamma, const float* beta,
                                                    const float* rm, const float* rv, int c, int C, double n) {
    __shared__ BnAffine s_af;
    if (threadIdx.x == 0) s_af = bn_affine(training, stats, gamma, beta, rm, rv, c, C, n);
    __syncthreads();
    return s_af;
}

__device__ __forceinline__ void bn_running_update(const double* stats, float* rm, float* rv, int64_t* nbt, int C,
                                                  double n, int tid) {
    if (tid < C) {
        const double mean = stats[tid] / n;
        double var = stats[C + tid] / n - mean * mean;
        if (var < 0.0) var = 0.0;
        const double unbiased = n > 1.0 ? var * (n / (n - 1.0)) : var;
        rm[tid] = (1.f - BN_MOMENTUM) * rm[tid] + BN_MOMENTUM * (float)mean;
        rv[tid] = (1.f - BN_MOMENTUM) * rv[tid] + BN_MOMENTUM * (float)unbiased;
    }
    if (tid == 0 && nbt) *nbt += 1;
}


// per-channel constants of the folded BatchNorm backward: s_bn[o] = {a, mean, inv, m1, m2}
template <int CO>
__device__ __forceinline__ void bn_bwd_constants(const BnBwd& bn, int Lout, float (*s_bn)[5]) {
    if (threadIdx.x < CO) {
        const int o = threadIdx.x;
        const double n = (double)bn.Bstat * (double)Lout;
        const BnAffine af = bn_affine(bn.training, bn.stats, bn.gamma, bn.beta, bn.rm, bn.rv, o, CO, n);
        s_bn[o][0] = af.a;
        s_bn[o][1] = af.mean;
        s_bn[o][2] = af.inv;
        s_bn[o][3] = bn.training ? (float)(bn.red[o] / n) : 0.f;
        s_bn[o][4] = bn.training ? (float)(bn.red[CO + o] / n) : 0.f;
    }
}

// sums of N values per lane over the 32 lanes of a warp by a transposing butterfly (N - 1 shuffles for N = 32 instead of
// 5 N): after the call lane L holds in v[0] the warp total of value index L (N == 32).
template <int N>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[N], int lane) {
    static_assert(N == 32, "one value per lane at the end");
    int n = N;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        n >>= 1;
        const bool upper = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            if (i < n) {
                const float keep = upper ? v[i + n] : v[i];
                const float send = upper ? v[i] : v[i + n];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
            }
        }
    }
}

}  // namespace mms
