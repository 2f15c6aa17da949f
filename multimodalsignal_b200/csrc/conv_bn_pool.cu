// CNN encoder of the reference (models.py:45-54): Conv1d(bias=False) -> BatchNorm1d -> ReLU ->
// MaxPool1d(3,2,1), twice.  fp32 SIMT kernels staged through shared memory:
//   conv1d_fwd    : x tile + (gate-scaled) weights in smem, one output position per thread, all
//                   C_out accumulators in registers, BN batch statistics reduced in the epilogue
//                   (float64 atomics -> order-independent to ~1e-16);
//   bn_relu_pool  : one streaming pass, optional time-major store so the permute(0,2,1) of
//                   models.py:77 costs nothing;
//   backward      : pool/ReLU/BN-reduction pass, BN apply pass, dgrad (optionally reduced against x
//                   for the attention gate) and wgrad.
#include "conv_common.cuh"
#include <stdlib.h>

namespace mms {

constexpr int CONV_CI_PAD = 16;

// ------------------------------------------------------------------------------------------
template <int CO, int KW, int S, int P, int TL>
__global__ void __launch_bounds__(TL) conv1d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ gate, float* __restrict__ y,
                                                        double* __restrict__ stats, int CI, int Lin, int Lout) {
    MMS_PDL_PROLOGUE();
    constexpr int SPAN = (TL - 1) * S + KW;
    extern __shared__ __align__(16) float smem[];
    float* ws = smem;                       // [CI*KW][CO]
    float* xs = smem + CI * KW * CO;        // [CI][SPAN]
    __shared__ double red[TL / 32][2 * CO];

    const int b = blockIdx.y, l0 = blockIdx.x * TL, tid = threadIdx.x;
    const int in0 = l0 * S - P;
    const float* xb = x + (size_t)b * CI * Lin;
    // the input tile goes to shared memory with cp.async (all copies of a thread in flight at once: one memory latency
    // for the whole tile); the weight staging below overlaps with it
    for (int idx = tid; idx < CI * SPAN; idx += TL) {
        const int c = idx / SPAN, i = idx - c * SPAN, gi = in0 + i;
        const bool ok = gi >= 0 && gi < Lin;
        cp_async4_zfill(xs + idx, xb + (size_t)c * Lin + (ok ? gi : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 4
    for (int idx = tid; idx < CO * CI * KW; idx += TL) {
        const int o = idx / (CI * KW), ck = idx - o * (CI * KW), c = ck / KW;
        const float g = gate ? gate[b * CI + c] : 1.f;
        ws[ck * CO + o] = w[idx] * g;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
    const float* xr = xs + tid * S;
    for (int c = 0; c < CI; ++c) {
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            const float xv = xr[c * SPAN + k];
            const float4* wv = reinterpret_cast<const float4*>(ws + (c * KW + k) * CO);
#pragma unroll
            for (int o4 = 0; o4 < CO / 4; ++o4) {
                const float4 wq = wv[o4];
                acc[4 * o4 + 0] += wq.x * xv;
                acc[4 * o4 + 1] += wq.y * xv;
                acc[4 * o4 + 2] += wq.z * xv;
                acc[4 * o4 + 3] += wq.w * xv;
            }
        }
    }
    const int l = l0 + tid;
    const bool valid = l < Lout;
    if (valid) {
        float* yb = y + (size_t)b * CO * Lout + l;
#pragma unroll
        for (int o = 0; o < CO; ++o) yb[(size_t)o * Lout] = acc[o];
    }
    if (stats) {
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int o = 0; o < CO; ++o) {
            const float v = valid ? acc[o] : 0.f;
            const float s = warp_sum(v), q = warp_sum(v * v);
            if (lane == 0) { red[warp][o] = (double)s; red[warp][CO + o] = (double)q; }
        }
        __syncthreads();
        if (tid < 2 * CO) {
            double t = 0.0;
#pragma unroll
            for (int wq = 0; wq < TL / 32; ++wq) t += red[wq][tid];
            atomicAdd(stats + tid, t);
        }
    }
}

// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float pooled_value(const float* __restrict__ row, int j, int Lin, float a, float bsh) {
    float m = -INFINITY;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const int i = 2 * j - 1 + e;
        if (i >= 0 && i < Lin) m = fmaxf(m, fmaxf(fmaf(a, __ldg(row + i), bsh), 0.f));
    }
    return m;
}

// Stage one row of y in shared memory as z = relu(a*y + b) with 128-bit loads where possible.
__device__ __forceinline__ void stage_row_relu(const float* __restrict__ row, int Lin, float a, float bsh, float* __restrict__ zs) {
    if ((Lin & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        for (int i = threadIdx.x; i < (Lin >> 2); i += blockDim.x) {
            const float4 v = __ldg(r4 + i);
            float4 z;
            z.x = fmaxf(fmaf(a, v.x, bsh), 0.f); z.y = fmaxf(fmaf(a, v.y, bsh), 0.f);
            z.z = fmaxf(fmaf(a, v.z, bsh), 0.f); z.w = fmaxf(fmaf(a, v.w, bsh), 0.f);
            reinterpret_cast<float4*>(zs)[i] = z;
        }
    } else {
        for (int i = threadIdx.x; i < Lin; i += blockDim.x) zs[i] = fmaxf(fmaf(a, __ldg(row + i), bsh), 0.f);
    }
}

// out[b,c,j] (channel-major)       grid = (C, B), block = 256, dynamic smem = Lin floats: one CTA per row
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_ncl_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float* rm, float* rv, int64_t* nbt, int Bn, int C, int Lin,
                                                                   int Lout, int training, float* __restrict__ out) {
    MMS_PDL_PROLOGUE();
    extern __shared__ __align__(16) float zs[];
    const int c = blockIdx.x, b = blockIdx.y;
    const double n = (double)Bn * (double)Lin;
    const BnAffine af = bn_affine_block(training, stats, gamma, beta, rm, rv, c, C, n);
    stage_row_relu(y + ((size_t)b * C + c) * Lin, Lin, af.a, af.b, zs);
    __syncthreads();
    float* orow = out + ((size_t)b * C + c) * Lout;
    for (int j = threadIdx.x; j < Lout; j += blockDim.x) {
        const int i = 2 * j;
        float m = zs[i];
        if (i - 1 >= 0) m = fmaxf(m, zs[i - 1]);
        if (i + 1 < Lin) m = fmaxf(m, zs[i + 1]);
        orow[j] = m;
    }
    if (training && blockIdx.x == 0 && blockIdx.y == 0) bn_running_update(stats, rm, rv, nbt, C, n, threadIdx.x);
}

// out[b,j,c] (time-major)          grid = (ceil(Lout/32), B), block = 256, C <= 64
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_tm_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float* rm, float* rv, int64_t* nbt, int Bn, int C, int Lin,
                                                                  int Lout, int training, float* __restrict__ out) {
    MMS_PDL_PROLOGUE();
    __shared__ float tile[32][65];
    __shared__ float sa[64], sb[64];
    const int b = blockIdx.y, j0 = blockIdx.x * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double n = (double)Bn * (double)Lin;
    // even rows, 8-byte aligned: the raw values of every (channel, window) of this thread are requested BEFORE the float64
    // BatchNorm constants are formed (they do not depend on them): one 64-bit load per window, the left neighbour by shuffle
    const bool fast = (C & 7) == 0 && (Lin & 1) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0;
    float2 raw[8];
    float left[8];
    if (fast) {
        const int j = j0 + lane;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            raw[i] = make_float2(0.f, 0.f);
            left[i] = 0.f;
            if (8 * i < C && j < Lout) {
                const float* row = y + ((size_t)b * C + warp + 8 * i) * Lin;
                raw[i] = __ldg(reinterpret_cast<const float2*>(row + 2 * j));
                if (lane == 0 && j > 0) left[i] = __ldg(row + 2 * j - 1);
            }
        }
    }
    if (threadIdx.x < C) {
        const BnAffine af = bn_affine(training, stats, gamma, beta, rm, rv, threadIdx.x, C, n);
        sa[threadIdx.x] = af.a;
        sb[threadIdx.x] = af.b;
    }
    __syncthreads();
    if (fast) {
        const int j = j0 + lane;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (8 * i < C) {
                const int c = warp + 8 * i;
                const float a = sa[c], bs = sb[c];
                const float z0 = fmaxf(fmaf(a, raw[i].x, bs), 0.f), z1 = fmaxf(fmaf(a, raw[i].y, bs), 0.f);
                float zl = __shfl_up_sync(0xffffffffu, z1, 1);
                // first window of the tile: its left element comes from the previous tile (none for j = 0: the pool pads with
                // -inf, and 0 is as good since every candidate is >= 0 after the ReLU)
                if (lane == 0) zl = j > 0 ? fmaxf(fmaf(a, left[i], bs), 0.f) : 0.f;
                if (j < Lout) tile[lane][c] = fmaxf(zl, fmaxf(z0, z1));
            }
        }
    } else {
        for (int c = warp; c < C; c += 8) {
            const int j = j0 + lane;
            if (j < Lout) tile[lane][c] = pooled_value(y + ((size_t)b * C + c) * Lin, j, Lin, sa[c], sb[c]);
        }
    }
    __syncthreads();
    for (int jj = warp; jj < 32; jj += 8) {
        const int j = j0 + jj;
        if (j < Lout)
            for (int c = lane; c < C; c += 32) out[((size_t)b * Lout + j) * C + c] = tile[jj][c];
    }
    if (training && blockIdx.x == 0 && blockIdx.y == 0) {
        __syncthreads();
        bn_running_update(stats, rm, rv, nbt, C, n, threadIdx.x);
    }
}

// ---- backward pass A: d(pool) -> d(relu) -> dyn, plus the two BN reductions --------------------
// grid = (C, B), block = 256, dynamic smem = (Lin + Lout) floats: one CTA per (b, c) row -> 128-bit row loads and
// one pair of float64 atomics per row (was per 256 elements).  dyn is written to dy.
__global__ void __launch_bounds__(256) pool_relu_bwd_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ rm, const float* __restrict__ rv,
                                                            const float* __restrict__ dout, int Bn, int C, int Lin, int Lout,
                                                            int training, int time_major, float* __restrict__ dy,
                                                            double* __restrict__ red) {
    MMS_PDL_PROLOGUE();
    extern __shared__ __align__(16) float bsm[];
    float* zs = bsm;                 // [Lin]  relu(bn(y))
    float* ds = bsm + ((Lin + 3) & ~3);   // [Lout] upstream gradient of this row
    __shared__ double part[8][2];
    const int c = blockIdx.x, b = blockIdx.y;
    const double n = (double)Bn * (double)Lin;
    const BnAffine af = bn_affine_block(training, stats, gamma, beta, rm, rv, c, C, n);
    const float* row = y + ((size_t)b * C + c) * Lin;
    stage_row_relu(row, Lin, af.a, af.b, zs);
    if (time_major) {
        for (int j = threadIdx.x; j < Lout; j += blockDim.x) ds[j] = __ldg(dout + ((size_t)b * Lout + j) * C + c);
    } else {
        const float* drow = dout + ((size_t)b * C + c) * Lout;
        for (int j = threadIdx.x; j < Lout; j += blockDim.x) ds[j] = __ldg(drow + j);
    }
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    float* dyrow = dy + ((size_t)b * C + c) * Lin;
    for (int i = threadIdx.x; i < Lin; i += blockDim.x) {
        // windows that contain i: centre j = i/2 for even i; j = (i+1)/2 (i is element 0) and j = (i-1)/2 (i is
        // element 2) for odd i.  First maximal element wins (ATen: val > maxval), -inf padding never wins.
        const float zi = zs[i];
        float dsum = 0.f;
        if ((i & 1) == 0) {
            const int j = i >> 1;                       // window (i-1, i, i+1)
            const bool left_wins = i - 1 >= 0 && zs[i - 1] >= zi;       // an earlier element with the same value wins
            const bool right_wins = i + 1 < Lin && zs[i + 1] > zi;
            if (j < Lout && !left_wins && !right_wins) dsum += ds[j];
        } else {
            const int ja = (i + 1) >> 1;                // window (i, i+1, i+2): i is the first element
            if (ja < Lout) {
                const bool lose = (i + 1 < Lin && zs[i + 1] > zi) || (i + 2 < Lin && zs[i + 2] > zi);
                if (!lose) dsum += ds[ja];
            }
            const int jb = (i - 1) >> 1;                // window (i-2, i-1, i): i is the last element
            if (jb < Lout) {
                const bool lose = (i - 2 >= 0 && zs[i - 2] >= zi) || zs[i - 1] >= zi;
                if (!lose) dsum += ds[jb];
            }
        }
        const float dyn = zi > 0.f ? dsum : 0.f;
        const float xhat = (__ldg(row + i) - af.mean) * af.inv;
        dyrow[i] = dyn;
        s1 += dyn;
        s2 += dyn * xhat;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { part[warp][0] = (double)s1; part[warp][1] = (double)s2; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int wq = 0; wq < 8; ++wq) t += part[wq][threadIdx.x];
        atomicAdd(red + threadIdx.x * C + c, t);      // red[0][c] = sum dyn, red[1][c] = sum dyn*xhat
    }
}

// ---- backward pass B: BN input gradient in place, dgamma / dbeta ---------------------------
// grid = (ceil(Lin/1024), C, B), block = 256: four elements per thread, 128-bit loads / stores
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ rm, const float* __restrict__ rv,
                                                           const double* __restrict__ red, int Bn, int C, int Lin, int training,
                                                           float* __restrict__ dy, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, float grad_scale) {
    const int c = blockIdx.y, b = blockIdx.z;
    const double n = (double)Bn * (double)Lin;
    const BnAffine af = bn_affine_block(training, stats, gamma, beta, rm, rv, c, C, n);
    const float m1 = training ? (float)(red[c] / n) : 0.f;
    const float m2 = training ? (float)(red[C + c] / n) : 0.f;
    const size_t base = ((size_t)b * C + c) * Lin;
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const bool vec = (Lin & 3) == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0;
    if (vec) {
        if (i0 < Lin) {
            const float4 yv = __ldg(reinterpret_cast<const float4*>(y + base + i0));
            float4 d = *reinterpret_cast<float4*>(dy + base + i0);
            d.x = af.a * (d.x - m1 - (yv.x - af.mean) * af.inv * m2);
            d.y = af.a * (d.y - m1 - (yv.y - af.mean) * af.inv * m2);
            d.z = af.a * (d.z - m1 - (yv.z - af.mean) * af.inv * m2);
            d.w = af.a * (d.w - m1 - (yv.w - af.mean) * af.inv * m2);
            *reinterpret_cast<float4*>(dy + base + i0) = d;
        }
    } else {
        for (int i = i0; i < min(i0 + 4, Lin); ++i) {
            const float xhat = (__ldg(y + base + i) - af.mean) * af.inv;
            dy[base + i] = af.a * (dy[base + i] - m1 - xhat * m2);
        }
    }
    if (blockIdx.x == 0 && b == 0 && threadIdx.x == 0) {
        dgamma[c] += grad_scale * (float)red[C + c];     // grad_scale = 1/world under data parallelism: every
        dbeta[c] += grad_scale * (float)red[c];          // rank holds the GLOBAL sums, the all-reduce adds them up
    }
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int floor_div2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

// dx[b,c,i] = sum_{o,k : S*l + k - P == i} w[o,c,k] * dy[b,o,l];   optional dot with xdot -> dgate
// CPAD = input channels rounded up to 8 or 16: the accumulator count (C = 6 needs 8, not 16)
template <int CO, int KW, int S, int P, int TI, int CPAD>
__global__ void __launch_bounds__(TI) conv1d_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          float* __restrict__ dx, const float* __restrict__ xdot,
                                                          float* __restrict__ dgate, int CI, int Lin, int Lout, const BnBwd bn) {
    MMS_PDL_PROLOGUE();
    static_assert(S == 2, "stride-2 convolutions only");
    static_assert(CO <= TI, "one thread per output channel for the BN constants");
    constexpr int NL = (TI - 1 + KW - 1) / S + 2;
    extern __shared__ __align__(16) float smem[];
    float* ws = smem;                           // [KW][CO][16]
    float* dys = smem + KW * CO * CPAD;  // [CO][NL]
    float* ys = dys + CO * NL;                  // [CO][NL], only with the folded BatchNorm backward
    __shared__ float red[TI / 32][CPAD];
    __shared__ float s_bn[CO][5];
    if (bn.y) bn_bwd_constants<CO>(bn, Lout, s_bn);

    const int b = blockIdx.y, i0 = blockIdx.x * TI, tid = threadIdx.x;
    const int lbase = floor_div2(i0 + P - (KW - 1));
    const float* dyb = dy + (size_t)b * CO * Lout;
    for (int idx = tid; idx < CO * NL; idx += TI) {          // cp.async: the whole dy tile in flight at once
        const int o = idx / NL, ll = idx - o * NL, l = lbase + ll;
        const bool ok = l >= 0 && l < Lout;
        cp_async4_zfill(dys + idx, dyb + (size_t)o * Lout + (ok ? l : 0), ok);
        if (bn.y) cp_async4_zfill(ys + idx, bn.y + (size_t)b * CO * Lout + (size_t)o * Lout + (ok ? l : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 4
    for (int idx = tid; idx < KW * CO * CPAD; idx += TI) {
        const int c = idx % CPAD, ko = idx / CPAD, o = ko % CO, k = ko / CO;
        ws[idx] = c < CI ? __ldg(w + ((size_t)o * CI + c) * KW + k) : 0.f;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (bn.y) {          // dyn -> dy in place (positions outside the tensor stay zero)
        for (int idx = tid; idx < CO * NL; idx += TI) {
            const int o = idx / NL, l = lbase + (idx - o * NL);
            if (l >= 0 && l < Lout)
                dys[idx] = s_bn[o][0] * (dys[idx] - s_bn[o][3] - (ys[idx] - s_bn[o][1]) * s_bn[o][2] * s_bn[o][4]);
        }
        __syncthreads();
    }

    const int i = i0 + tid;
    float acc[CPAD];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) acc[c] = 0.f;
    for (int k = (i + P) & 1; k < KW; k += S) {
        const int ll = (i + P - k) / S - lbase;     // i + P - k is even; may be negative -> staged as zero
        if (ll < 0 || ll >= NL) continue;
        for (int o = 0; o < CO; ++o) {
            const float d = dys[o * NL + ll];
            const float4* wv = reinterpret_cast<const float4*>(ws + (k * CO + o) * CPAD);
#pragma unroll
            for (int c4 = 0; c4 < CPAD / 4; ++c4) {
                const float4 wq = wv[c4];
                acc[4 * c4 + 0] += wq.x * d;
                acc[4 * c4 + 1] += wq.y * d;
                acc[4 * c4 + 2] += wq.z * d;
                acc[4 * c4 + 3] += wq.w * d;
            }
        }
    }
    const bool valid = i < Lin;
    if (dx && valid) {
#pragma unroll
        for (int c = 0; c < CPAD; ++c)
            if (c < CI) dx[((size_t)b * CI + c) * Lin + i] = acc[c];
    }
    if (xdot) {
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
            float v = 0.f;
            if (valid && c < CI) v = acc[c] * __ldg(xdot + ((size_t)b * CI + c) * Lin + i);
            v = warp_sum(v);
            if (lane == 0) red[warp][c] = v;
        }
        __syncthreads();
        if (tid < CI) {
            float t = 0.f;
#pragma unroll
            for (int wq = 0; wq < TI / 32; ++wq) t += red[wq][tid];
            atomicAdd(dgate + b * CI + tid, t);
        }
    }
}

// dw[o,c,k] += gate[b,c] * sum_l dy[b,o,l] * x[b,c,S*l+k-P]
// One CTA = one batch row x TLW output positions, staged in shared memory.  A thread owns one input channel c, one group
// of 16 output channels and every NPL-th position of the tile: its 16 x KW partial sums live in registers, so a position
// costs 16 + KW shared-memory loads for 16 * KW FMAs (the first version re-read both operands for every FMA pair and kept
// 96 of 256 threads busy).  The NPL position lanes of a (c, group) pair sit in one warp and are combined with shuffles;
// one atomicAdd per weight and CTA follows (4x fewer CTAs than before).  grid = (ceil(Lout/TLW), B), block = NT.
template <int CO, int KW, int S, int P, int TLW, int NPL, int NT>
__global__ void __launch_bounds__(NT) conv1d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          const float* __restrict__ gate, float* __restrict__ dw, int CI,
                                                          int Lin, int Lout, const BnBwd bn) {
    constexpr int SPAN = (TLW - 1) * S + KW;
    constexpr int OG = 16, NOG = CO / OG;
    constexpr int DLD = TLW + 4;            // padded row of the dy tile
    static_assert(CO % OG == 0 && 32 % NPL == 0 && TLW % NPL == 0 && TLW % 4 == 0 && CO <= NT, "conv1d_wgrad: bad tiling");
    extern __shared__ __align__(16) float smem[];
    float* dys = smem;                 // [CO][DLD]
    float* xs = smem + CO * DLD;       // [CI][SPAN]
    float* ys = xs + ((CI * SPAN + 3) & ~3);      // [CO][DLD], only with the folded BatchNorm backward
    __shared__ float s_bn[CO][5];
    if (bn.y) {
        bn_bwd_constants<CO>(bn, Lout, s_bn);
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < CO) {      // dgamma / dbeta once per launch
            if (bn.dgamma) bn.dgamma[threadIdx.x] += bn.grad_scale * (float)bn.red[CO + threadIdx.x];
            if (bn.dbeta) bn.dbeta[threadIdx.x] += bn.grad_scale * (float)bn.red[threadIdx.x];
        }
    }
    const int b = blockIdx.y, l0 = blockIdx.x * TLW, tid = threadIdx.x;
    const int in0 = l0 * S - P;
    const float* xb = x + (size_t)b * CI * Lin;
    const float* dyb = dy + (size_t)b * CO * Lout;
    // Both tiles go to shared memory with cp.async (zero-filled outside the tensors): every copy of a thread is in flight at
    // once, so the load phase costs one memory latency instead of one per row.
    for (int idx = tid; idx < CI * SPAN; idx += NT) {
        const int c = idx / SPAN, i = idx - c * SPAN, gi = in0 + i;
        const bool ok = gi >= 0 && gi < Lin;
        cp_async4_zfill(xs + idx, xb + (size_t)c * Lin + (ok ? gi : 0), ok);
    }
    if ((Lout & 3) == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(bn.y)) & 15) == 0) {
        for (int idx = tid; idx < CO * (TLW / 4); idx += NT) {
            const int o = idx / (TLW / 4), i = (idx - o * (TLW / 4)) * 4;
            const bool ok = l0 + i < Lout;            // Lout % 4 == 0: a 16-byte group is inside or outside as a whole
            cp_async16_zfill(dys + o * DLD + i, dyb + (size_t)o * Lout + (ok ? l0 + i : 0), ok);
            if (bn.y) cp_async16_zfill(ys + o * DLD + i, bn.y + (size_t)b * CO * Lout + (size_t)o * Lout + (ok ? l0 + i : 0), ok);
        }
    } else {
        for (int idx = tid; idx < CO * TLW; idx += NT) {
            const int o = idx / TLW, i = idx - o * TLW;
            const bool ok = l0 + i < Lout;
            cp_async4_zfill(dys + o * DLD + i, dyb + (size_t)o * Lout + (ok ? l0 + i : 0), ok);
            if (bn.y) cp_async4_zfill(ys + o * DLD + i, bn.y + (size_t)b * CO * Lout + (size_t)o * Lout + (ok ? l0 + i : 0), ok);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (bn.y) {          // dyn -> dy in place (positions outside the tensor stay zero)
        for (int idx = tid; idx < CO * TLW; idx += NT) {
            const int o = idx / TLW, i = idx - o * TLW;
            if (l0 + i < Lout) {
                float* d = dys + o * DLD + i;
                *d = s_bn[o][0] * (*d - s_bn[o][3] - (ys[o * DLD + i] - s_bn[o][1]) * s_bn[o][2] * s_bn[o][4]);
            }
        }
        __syncthreads();
    }
    // thread -> (pair = (c, og), position lane j); the NPL lanes of a pair are adjacent lanes of one warp
    const int pair = tid / NPL, j = tid % NPL;
    const int npairs = CI * NOG;
    const bool active = pair < npairs;  // idle lanes still take part in the shuffles below
    const int c = active ? pair / NOG : 0, og = active ? pair % NOG : 0;
    float acc[OG][KW];
#pragma unroll
    for (int o = 0; o < OG; ++o)
#pragma unroll
        for (int k = 0; k < KW; ++k) acc[o][k] = 0.f;
    const float* xr = xs + c * SPAN;
    const float* dr = dys + (og * OG) * DLD;
    for (int l = active ? j : TLW; l < TLW; l += NPL) {
        float xv[KW];
#pragma unroll
        for (int k = 0; k < KW; ++k) xv[k] = xr[l * S + k];
#pragma unroll
        for (int o = 0; o < OG; ++o) {
            const float d = dr[o * DLD + l];
#pragma unroll
            for (int k = 0; k < KW; ++k) acc[o][k] = fmaf(d, xv[k], acc[o][k]);
        }
    }
    const float g = gate ? gate[b * CI + c] : 1.f;
    float* dwp = dw + ((size_t)(og * OG) * CI + c) * KW;
#pragma unroll
    for (int o = 0; o < OG; ++o)
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            float v = acc[o][k];
#pragma unroll
            for (int m = NPL / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
            if (active && j == (o * KW + k) % NPL) atomicAdd(dwp + (size_t)o * CI * KW + k, v * g);
        }
}

// ---- launchers -------------------------------------------------------------------------------
template <int CO, int KW, int S, int P, int TL>
static int conv_fwd_launch(const float* x, const float* w, const float* gate, int B, int CI, int Lin, float* y,
                           double* stats, cudaStream_t st) {
    const int Lout = conv_out_len(Lin, KW, S, P);
    constexpr int SPAN = (TL - 1) * S + KW;
    const size_t smem = (size_t)(CI * KW * CO + CI * SPAN) * sizeof(float);
    auto kern = conv1d_fwd_kernel<CO, KW, S, P, TL>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) { MMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); }
    MMS_REQUIRE(smem <= 96 * 1024, "conv1d_fwd: shared memory %zu too large", smem);
    dim3 grid(cdiv(Lout, TL), B);
    MMS_PROF_BEGIN(st);
    MMS_LAUNCH(kern, grid, dim3(TL), smem, st, x, w, gate, y, stats, CI, Lin, Lout);
    MMS_LAUNCH_CHECK("conv1d_fwd_kernel");
    return MMS_OK;
}

template <int CO, int KW, int S, int P, int TI, int CPAD>
static int conv_dgrad_launch_pad(const float* dy, const float* w, int B, int CI, int Lin, float* dx, const float* xdot,
                                 float* dgate, cudaStream_t st, const BnBwd& bn) {
    const int Lout = conv_out_len(Lin, KW, S, P);
    constexpr int NL = (TI - 1 + KW - 1) / S + 2;
    const size_t smem = (size_t)(KW * CO * CPAD + (bn.y ? 2 : 1) * CO * NL) * sizeof(float);
    auto kern = conv1d_dgrad_kernel<CO, KW, S, P, TI, CPAD>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) { MMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); }
    MMS_REQUIRE(smem <= 96 * 1024, "conv1d_dgrad: shared memory %zu too large", smem);
    dim3 grid(cdiv(Lin, TI), B);
    MMS_PROF_BEGIN(st);
    MMS_LAUNCH(kern, grid, dim3(TI), smem, st, dy, w, dx, xdot, dgate, CI, Lin, Lout, bn);
    MMS_LAUNCH_CHECK("conv1d_dgrad_kernel");
    return MMS_OK;
}

template <int CO, int KW, int S, int P, int TI>
static int conv_dgrad_launch(const float* dy, const float* w, int B, int CI, int Lin, float* dx, const float* xdot,
                             float* dgate, cudaStream_t st, const BnBwd& bn) {
    if (CI <= 8) return conv_dgrad_launch_pad<CO, KW, S, P, TI, 8>(dy, w, B, CI, Lin, dx, xdot, dgate, st, bn);
    return conv_dgrad_launch_pad<CO, KW, S, P, TI, 16>(dy, w, B, CI, Lin, dx, xdot, dgate, st, bn);
}

template <int CO, int KW, int S, int P, int TLW, int NPL, int NT>
static int conv_wgrad_launch(const float* x, const float* dy, const float* gate, int B, int CI, int Lin, float* dw,
                             cudaStream_t st, const BnBwd& bn) {
    const int Lout = conv_out_len(Lin, KW, S, P);
    constexpr int SPAN = (TLW - 1) * S + KW;
    MMS_REQUIRE(CI * (CO / 16) * NPL <= NT, "conv1d_wgrad: %d input channels do not fit the thread mapping", CI);
    const size_t smem = (size_t)((bn.y ? 2 : 1) * CO * (TLW + 4) + ((CI * SPAN + 3) & ~3)) * sizeof(float);
    auto kern = conv1d_wgrad_kernel<CO, KW, S, P, TLW, NPL, NT>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) { MMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); }
    MMS_REQUIRE(smem <= 220 * 1024, "conv1d_wgrad: shared memory %zu too large", smem);
    dim3 grid(cdiv(Lout, TLW), B);
    MMS_PROF_BEGIN(st);
    kern<<<grid, NT, smem, st>>>(x, dy, gate, dw, CI, Lin, Lout, bn);
    MMS_LAUNCH_CHECK("conv1d_wgrad_kernel");
    return MMS_OK;
}

static int check_conv(int which, int c_in, int c_out) {
    if (which == 1) {
        MMS_REQUIRE(c_out == CONV1_CO && c_in >= 1 && c_in <= CONV_CI_PAD, "conv1: need C_out=16 and 1<=C_in<=16 (got %d,%d)", c_out, c_in);
    } else if (which == 2) {
        MMS_REQUIRE(c_in == CONV2_CI && (c_out == 16 || c_out == 32 || c_out == 64), "conv2: need C_in=16 and C_out in {16,32,64} (got %d,%d)", c_in, c_out);
    } else {
        MMS_REQUIRE(false, "conv: which must be 1 or 2");
    }
    return MMS_OK;
}

bool conv_fwd_tc_supported(int which, const float* x, int c_in, int c_out, int l_in);
int launch_conv_fwd_tc(int which, const float* x, const float* w, const float* gate, int B, int c_in, int c_out, int l_in, float* y,
                       double* stats, cudaStream_t st);

// The tcgen05 implicit-GEMM convolution (conv_tc.cu) is correct to fp32 accuracy but, measured inside the training step
// (B = 64, T = 3840), it does not beat the SIMT kernel: 22.4 us against 21.0 us per launch and +19 us per step -- with
// K = 42 / 80 the tensor-core work is ~1 us and the per-CTA set-up (TMEM allocation, TMA round trip, operand expansion,
// commit / wait) is exposed.  It is therefore opt-in: MMS_CONV_TC=1 (profiles/r1_conv_tc_vs_simt.md).
static bool conv_use_tc() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MMS_CONV_TC"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

int launch_conv_fwd(int which, const float* x, const float* w, const float* gate, int B, int c_in, int c_out, int l_in,
                    float* y, double* stats, cudaStream_t st) {
    int rc = check_conv(which, c_in, c_out);
    if (rc) return rc;
    if (conv_use_tc() && conv_fwd_tc_supported(which, x, c_in, c_out, l_in))     // implicit GEMM on tcgen05 (conv_tc.cu)
        return launch_conv_fwd_tc(which, x, w, gate, B, c_in, c_out, l_in, y, stats, st);
    if (which == 1) return conv_fwd_launch<16, CONV1_K, CONV1_S, CONV1_P, 256>(x, w, gate, B, c_in, l_in, y, stats, st);
    if (c_out == 16) return conv_fwd_launch<16, CONV2_K, CONV2_S, CONV2_P, 128>(x, w, gate, B, c_in, l_in, y, stats, st);
    if (c_out == 32) return conv_fwd_launch<32, CONV2_K, CONV2_S, CONV2_P, 128>(x, w, gate, B, c_in, l_in, y, stats, st);
    return conv_fwd_launch<64, CONV2_K, CONV2_S, CONV2_P, 128>(x, w, gate, B, c_in, l_in, y, stats, st);
}

static const BnBwd NO_BN = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 1.f};

int launch_conv_dgrad(int which, const float* dy, const float* w, int B, int c_in, int c_out, int l_in, float* dx,
                      const float* xdot, float* dgate, cudaStream_t st, const BnBwd* bnp = nullptr) {
    int rc = check_conv(which, c_in, c_out);
    if (rc) return rc;
    const BnBwd& bn = bnp ? *bnp : NO_BN;
    if (which == 1) return conv_dgrad_launch<16, CONV1_K, CONV1_S, CONV1_P, 256>(dy, w, B, c_in, l_in, dx, xdot, dgate, st, bn);
    if (c_out == 16) return conv_dgrad_launch<16, CONV2_K, CONV2_S, CONV2_P, 128>(dy, w, B, c_in, l_in, dx, xdot, dgate, st, bn);
    if (c_out == 32) return conv_dgrad_launch<32, CONV2_K, CONV2_S, CONV2_P, 128>(dy, w, B, c_in, l_in, dx, xdot, dgate, st, bn);
    return conv_dgrad_launch<64, CONV2_K, CONV2_S, CONV2_P, 128>(dy, w, B, c_in, l_in, dx, xdot, dgate, st, bn);
}

int launch_conv_wgrad(int which, const float* x, const float* dy, const float* gate, int B, int c_in, int c_out,
                      int l_in, float* dw, cudaStream_t st, const BnBwd* bnp = nullptr) {
    int rc = check_conv(which, c_in, c_out);
    if (rc) return rc;
    const BnBwd& bn = bnp ? *bnp : NO_BN;
    // conv1: C_in <= 16 pairs x 32 position lanes (512 threads); conv2: 16 channels x C_out/16 groups x 8 lanes
    if (which == 1) {
        // MMS_WGRAD1_TILE=480 (experiment): the 480-position tile of the earlier version (85 KB instead of 170 KB of shared memory
        // per CTA, twice the atomics) -- the weight-gradient kernels run beside the main chain and block SMs by their footprint
        if (c_in <= 8 && option_get("WGRAD1_TILE", 960) == 480)
            return conv_wgrad_launch<16, CONV1_K, CONV1_S, CONV1_P, 480, 32, 256>(x, dy, gate, B, c_in, l_in, dw, st, bn);
        if (c_in <= 8) return conv_wgrad_launch<16, CONV1_K, CONV1_S, CONV1_P, 960, 32, 256>(x, dy, gate, B, c_in, l_in, dw, st, bn);
        return conv_wgrad_launch<16, CONV1_K, CONV1_S, CONV1_P, 480, 32, 512>(x, dy, gate, B, c_in, l_in, dw, st, bn);
    }
    if (c_out == 16) return conv_wgrad_launch<16, CONV2_K, CONV2_S, CONV2_P, 240, 8, 128>(x, dy, gate, B, c_in, l_in, dw, st, bn);
    // MMS_WGRAD2_TILE=120 (experiment): 47 KB instead of 93 KB of shared memory per CTA, twice the CTAs and atomics
    if (c_out == 32 && option_get("WGRAD2_TILE", 240) == 120)
        return conv_wgrad_launch<32, CONV2_K, CONV2_S, CONV2_P, 120, 8, 256>(x, dy, gate, B, c_in, l_in, dw, st, bn);
    if (c_out == 32) return conv_wgrad_launch<32, CONV2_K, CONV2_S, CONV2_P, 240, 8, 256>(x, dy, gate, B, c_in, l_in, dw, st, bn);
    return conv_wgrad_launch<64, CONV2_K, CONV2_S, CONV2_P, 240, 8, 512>(x, dy, gate, B, c_in, l_in, dw, st, bn);
}

int launch_bn_relu_pool_fwd(const float* y, const double* stats, const float* gamma, const float* beta, float* rm,
                            float* rv, int64_t* nbt, int B, int C, int l_in, int training, int time_major, float* out,
                            cudaStream_t st, int Bstat = 0) {
    if (Bstat <= 0) Bstat = B;        // batch the statistics were reduced over (global batch under data parallelism)
    const int Lout = pool_out_len(l_in);
    MMS_REQUIRE(C >= 1 && C <= 64, "bn_relu_pool: channels %d outside [1,64]", C);
    MMS_REQUIRE(!training || stats, "bn_relu_pool: training mode needs batch statistics");
    if (time_major) {
        dim3 grid(cdiv(Lout, 32), B);
        MMS_PROF_BEGIN(st);
        MMS_LAUNCH(bn_relu_pool_fwd_tm_kernel, grid, dim3(256), 0, st, y, stats, gamma, beta, rm, rv, nbt, Bstat, C, l_in, Lout, training, out);
    } else {
        dim3 grid(C, B);
        MMS_REQUIRE((size_t)l_in * sizeof(float) <= 40 * 1024, "bn_relu_pool: row length %d too long for one CTA's shared memory", l_in);
        MMS_PROF_BEGIN(st);
        MMS_LAUNCH(bn_relu_pool_fwd_ncl_kernel, grid, dim3(256), (size_t)l_in * sizeof(float), st, y, stats, gamma, beta, rm, rv, nbt, Bstat, C, l_in, Lout, training, out);
    }
    MMS_LAUNCH_CHECK("bn_relu_pool_fwd");
    return MMS_OK;
}

// which: bit 0 = pool/ReLU backward + the two BN reductions into `red`; bit 1 = BN apply (+ dgamma, dbeta).
// Under data parallelism the caller all-reduces `red` between the two.
int launch_bn_relu_pool_bwd(const float* y, const double* stats, const float* gamma, const float* beta, const float* rm,
                            const float* rv, const float* dout, int B, int C, int l_in, int training, int time_major,
                            float* dy, float* dgamma, float* dbeta, double* red, cudaStream_t st, int which = 3, int Bstat = 0,
                            float grad_scale = 1.f) {
    const int Lout = pool_out_len(l_in);
    if (Bstat <= 0) Bstat = B;
    MMS_REQUIRE(C >= 1 && C <= 64, "bn_relu_pool_bwd: channels %d outside [1,64]", C);
    if (which & 1) {
        dim3 grid(C, B);
        const size_t smem = (size_t)(((l_in + 3) & ~3) + Lout) * sizeof(float);
        MMS_REQUIRE(smem <= 44 * 1024, "bn_relu_pool_bwd: row length %d too long for one CTA's shared memory", l_in);
        MMS_PROF_BEGIN(st);
        MMS_LAUNCH(pool_relu_bwd_kernel, grid, dim3(256), smem, st, y, stats, gamma, beta, rm, rv, dout, Bstat, C, l_in, Lout, training, time_major, dy, red);
        MMS_LAUNCH_CHECK("pool_relu_bwd_kernel");
    }
    if (which & 2) {
        dim3 grid(cdiv(l_in, 1024), C, B);
        MMS_PROF_BEGIN(st);
        bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(y, stats, gamma, beta, rm, rv, red, Bstat, C, l_in, training, dy, dgamma, dbeta, grad_scale);
        MMS_LAUNCH_CHECK("bn_bwd_apply_kernel");
    }
    return MMS_OK;
}

}  // namespace mms

using namespace mms;

extern "C" int mms_conv1d_fwd_tc(int32_t which, const float* x, const float* w, const float* gate, int32_t B, int32_t c_in, int32_t c_out,
                                 int32_t l_in, float* y, double* stats, mms_stream_t stream) {
    MMS_REQUIRE(x && w && y && B > 0 && l_in > 0, "conv1d_fwd_tc: bad arguments");
    int rc = check_conv(which, c_in, c_out);
    if (rc) return rc;
    return launch_conv_fwd_tc(which, x, w, gate, B, c_in, c_out, l_in, y, stats, (cudaStream_t)stream);
}

extern "C" int mms_conv1d_fwd(int32_t which, const float* x, const float* w, const float* gate, int32_t B, int32_t c_in,
                              int32_t c_out, int32_t l_in, float* y, double* stats, mms_stream_t stream) {
    MMS_REQUIRE(x && w && y && B > 0 && l_in > 0, "conv1d_fwd: bad arguments");
    return launch_conv_fwd(which, x, w, gate, B, c_in, c_out, l_in, y, stats, (cudaStream_t)stream);
}
extern "C" int mms_conv1d_dgrad(int32_t which, const float* dy, const float* w, int32_t B, int32_t c_in, int32_t c_out,
                                int32_t l_in, float* dx, const float* xdot, float* dgate, mms_stream_t stream) {
    MMS_REQUIRE(dy && w && B > 0 && l_in > 0 && (dx || (xdot && dgate)), "conv1d_dgrad: bad arguments");
    return launch_conv_dgrad(which, dy, w, B, c_in, c_out, l_in, dx, xdot, dgate, (cudaStream_t)stream);
}
extern "C" int mms_conv1d_wgrad(int32_t which, const float* x, const float* dy, const float* gate, int32_t B, int32_t c_in,
                                int32_t c_out, int32_t l_in, float* dw, mms_stream_t stream) {
    MMS_REQUIRE(x && dy && dw && B > 0 && l_in > 0, "conv1d_wgrad: bad arguments");
    return launch_conv_wgrad(which, x, dy, gate, B, c_in, c_out, l_in, dw, (cudaStream_t)stream);
}
extern "C" int mms_bn_relu_pool_fwd(const float* y, const double* stats, const float* gamma, const float* beta,
                                    float* running_mean, float* running_var, int64_t* nbt, int32_t B, int32_t C, int32_t l_in,
                                    int32_t training, int32_t time_major, float* out, mms_stream_t stream) {
    MMS_REQUIRE(y && gamma && beta && running_mean && running_var && out, "bn_relu_pool_fwd: bad arguments");
    return launch_bn_relu_pool_fwd(y, stats, gamma, beta, running_mean, running_var, nbt, B, C, l_in, training, time_major, out,
                                   (cudaStream_t)stream);
}
extern "C" int mms_bn_relu_pool_bwd(const float* y, const double* stats, const float* gamma, const float* beta,
                                    const float* running_mean, const float* running_var, const float* dout, int32_t B, int32_t C,
                                    int32_t l_in, int32_t training, int32_t time_major, float* dy, float* dgamma, float* dbeta,
                                    double* red, mms_stream_t stream) {
    MMS_REQUIRE(y && gamma && beta && dout && dy && dgamma && dbeta && red, "bn_relu_pool_bwd: bad arguments");
    return launch_bn_relu_pool_bwd(y, stats, gamma, beta, running_mean, running_var, dout, B, C, l_in, training, time_major, dy,
                                   dgamma, dbeta, red, (cudaStream_t)stream);
}
