// Backward kernels of the CNN encoder (reverse of reference models.py:45-53 and 24-31), the default path when T % 8 == 0.
// Main chain of a training step:  pool_relu_bwd_tile (stage 2) -> conv2_dgrad -> pool_relu_bwd_tile (stage 1) ->
// conv1_wgrad_dgate; the weight gradient of conv2 stays on a side stream.
//
//   pool_relu_bwd_tile_kernel  MaxPool1d(3,2,1) + ReLU backward of one (batch row, position tile) for ALL channels, plus the
//                              two BatchNorm reductions (sum dyn, sum dyn * xhat).  The upstream gradient of stage 2 is
//                              time-major [B, L, C] (the GRU's layout): it is read in full 128-byte rows and transposed in
//                              shared memory instead of being gathered with stride C.
//   conv2_dgrad_kernel         input gradient of Conv1d(16, C_out, k5, s2, p2), BatchNorm-backward apply folded into the
//                              tile staging (BnBwd).  A thread owns 4 (even, odd) position pairs x 8 input channels: both
//                              parities use the same three upstream values, the 5 taps of a pair are 20 packed FFMA2 per
//                              output channel against 10 broadcast 128-bit weight loads.
//   conv1_wgrad_dgate_kernel   G[b,o,c,k] = sum_l dy1[b,o,l] * x[b,c,2l+k-3] per batch row, from which BOTH remaining gradients of
//                              the first stage follow:  dW1[o,c,k] += gate[b,c] * G  and  dgate[b,c] = sum_{o,k} W1[o,c,k] * G --
//                              the input gradient of conv1 is never formed (it was only ever reduced against x), and the
//                              ChannelAttention parameter gradients are finished in the same launch.  One thread-block CLUSTER
//                              per batch row: a warp owns an input channel and all 16 x 7 accumulators, lanes stride over the
//                              positions; partial G meets through distributed shared memory.
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mms {

__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float cl_ld_f32(const void* smem_ptr, uint32_t rank) {
    uint32_t a;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(smem_ptr)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
// grid = (ceil(Lin / TI), B), block = 256.  dynamic smem: zs [C][TI + 8] | ds [C][TI/2 + 1]
template <int C, int TI, bool TIME_MAJOR>
__global__ void __launch_bounds__(256) pool_relu_bwd_tile_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ rm, const float* __restrict__ rv,
                                                                 const float* __restrict__ dout, int Bn, int Lin, int Lout,
                                                                 int training, float* __restrict__ dy, double* __restrict__ red) {
    constexpr int ZW = TI + 8, NJ = TI / 2 + 1;
    static_assert(TI % 64 == 0 && (NJ & 1) == 1, "tile");
    extern __shared__ __align__(16) float pb_smem[];
    float* zs = pb_smem;
    float* ds = pb_smem + C * ZW;
    __shared__ float s_a[C], s_b[C], s_mean[C], s_inv[C];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, i0 = blockIdx.x * TI, j0 = i0 >> 1;
    if (tid < C) {
        const BnAffine af = bn_affine(training, stats, gamma, beta, rm, rv, tid, C, (double)Bn * (double)Lin);
        s_a[tid] = af.a; s_b[tid] = af.b; s_mean[tid] = af.mean; s_inv[tid] = af.inv;
    }
    // upstream gradient of the windows j0 .. j0 + TI/2 (the last one belongs to the next tile's first element too)
    const int nj = min(NJ, Lout - j0);
    if (TIME_MAJOR) {
        const float* src = dout + ((size_t)b * Lout + j0) * C;
        for (int idx = tid; idx < nj * C; idx += 256) {
            const int jj = idx / C, c = idx - jj * C;
            ds[c * NJ + jj] = __ldg(src + idx);
        }
    } else {
        for (int c = warp; c < C; c += 8) {
            const float* src = dout + ((size_t)b * C + c) * Lout + j0;
            for (int jj = lane; jj < nj; jj += 32) ds[c * NJ + jj] = __ldg(src + jj);
        }
    }
    __syncthreads();
    // z = relu(bn(y)) for i in [i0 - 4, i0 + TI + 4); -inf outside the row (never wins a window)
    const bool vec = (Lin & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    for (int c = warp; c < C; c += 8) {
        const float* row = y + ((size_t)b * C + c) * Lin;
        const float a = s_a[c], bsh = s_b[c];
        for (int q = lane; q < ZW / 4; q += 32) {
            const int i = i0 - 4 + 4 * q;
            float4 z;
            if (vec && i >= 0 && i + 3 < Lin) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row + i));
                z.x = fmaxf(fmaf(a, v.x, bsh), 0.f); z.y = fmaxf(fmaf(a, v.y, bsh), 0.f);
                z.z = fmaxf(fmaf(a, v.z, bsh), 0.f); z.w = fmaxf(fmaf(a, v.w, bsh), 0.f);
            } else {
                float t[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) t[e] = (i + e >= 0 && i + e < Lin) ? fmaxf(fmaf(a, __ldg(row + i + e), bsh), 0.f) : -INFINITY;
                z = make_float4(t[0], t[1], t[2], t[3]);
            }
            *reinterpret_cast<float4*>(zs + c * ZW + 4 * q) = z;
        }
    }
    __syncthreads();
    for (int c = warp; c < C; c += 8) {
        const float* row = y + ((size_t)b * C + c) * Lin;
        float* dyrow = dy + ((size_t)b * C + c) * Lin;
        const float* zr = zs + c * ZW + 4;          // zr[ii] = z[i0 + ii]
        const float* dr = ds + c * NJ - j0;         // dr[j]
        const float mean = s_mean[c], inv = s_inv[c];
        float s1 = 0.f, s2 = 0.f;
        for (int ii = lane; ii < TI; ii += 32) {
            const int i = i0 + ii;
            if (i >= Lin) break;
            // windows that contain i: centre j = i/2 for even i; j = (i+1)/2 (i is element 0) and j = (i-1)/2 (i is element 2)
            // for odd i.  First maximal element wins (ATen: val > maxval).
            const float zi = zr[ii];
            float dsum = 0.f;
            if ((i & 1) == 0) {
                const int j = i >> 1;
                if (j < Lout && !(zr[ii - 1] >= zi) && !(zr[ii + 1] > zi)) dsum = dr[j];
            } else {
                const int ja = (i + 1) >> 1, jb = (i - 1) >> 1;
                if (ja < Lout && !(zr[ii + 1] > zi) && !(zr[ii + 2] > zi)) dsum += dr[ja];
                if (jb < Lout && !(zr[ii - 2] >= zi) && !(zr[ii - 1] >= zi)) dsum += dr[jb];
            }
            const float dyn = zi > 0.f ? dsum : 0.f;
            const float xhat = (__ldg(row + i) - mean) * inv;
            dyrow[i] = dyn;
            s1 += dyn;
            s2 = fmaf(dyn, xhat, s2);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
            atomicAdd(red + c, (double)s1);          // red[0][c] = sum dyn, red[1][c] = sum dyn * xhat
            atomicAdd(red + C + c, (double)s2);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// grid = (ceil(P1 / 256), B), block = 64: thread (h, t) = (tid >> 5, tid & 31) owns the input channels [8h, 8h + 8) and the
// position pairs u = u0 + 4t + e (e < 4), i.e. dp1[:, 2u] and dp1[:, 2u + 1]; u0 = 128 * blockIdx.x.
//   even j = 2u   : taps k = 0, 2, 4 with m = u + 1, u, u - 1;    odd j = 2u + 1 : taps k = 1, 3 with m = u + 1, u
// dynamic smem: wd [CO2 * 5][16] | dys [CO2][132] | ys [CO2][132] (BatchNorm fold only); dys[o][mm] holds m = u0 - 1 + mm.
template <int CO2>
__global__ void __launch_bounds__(64) conv2_dgrad_kernel(const float* __restrict__ dyn, const float* __restrict__ w,
                                                         float* __restrict__ dx, int Lin, int Lout, const BnBwd bn) {
    constexpr int NM = 130, NMP = 132;
    extern __shared__ __align__(16) float dg_smem[];
    float* wd = dg_smem;
    float* dys = wd + CO2 * 80;
    float* ys = dys + CO2 * NMP;
    __shared__ float s_bn[CO2][5];
    const int tid = threadIdx.x, h = tid >> 5, t = tid & 31;
    const int b = blockIdx.y, u0 = blockIdx.x * 128;
    if (bn.y) bn_bwd_constants<CO2>(bn, Lout, s_bn);
    // wd[(o*5 + k)*16 + ci] = w[o][ci][k]
    for (int idx = tid; idx < CO2 * 80; idx += 64) {
        const int o = idx / 80, r = idx - o * 80, ci = r / 5, k = r - ci * 5;
        wd[(o * 5 + k) * 16 + ci] = __ldg(w + idx);
    }
    const float* dyb = dyn + (size_t)b * CO2 * Lout;
    const float* yb = bn.y ? bn.y + (size_t)b * CO2 * Lout : nullptr;
    for (int idx = tid; idx < CO2 * NM; idx += 64) {
        const int o = idx / NM, mm = idx - o * NM, m = u0 - 1 + mm;
        const bool ok = m >= 0 && m < Lout;
        dys[o * NMP + mm] = ok ? __ldg(dyb + (size_t)o * Lout + m) : 0.f;
        if (yb) ys[o * NMP + mm] = ok ? __ldg(yb + (size_t)o * Lout + m) : 0.f;
    }
    __syncthreads();
    if (bn.y) {          // dyn -> dy in place (positions outside the tensor stay zero)
        for (int idx = tid; idx < CO2 * NM; idx += 64) {
            const int o = idx / NM, mm = idx - o * NM, m = u0 - 1 + mm;
            if (m >= 0 && m < Lout) {
                float* d = dys + o * NMP + mm;
                *d = s_bn[o][0] * (*d - s_bn[o][3] - (ys[o * NMP + mm] - s_bn[o][1]) * s_bn[o][2] * s_bn[o][4]);
            }
        }
        __syncthreads();
    }
    float2 ae[4][4], ao[4][4];      // [pair e][input-channel pair]: even / odd position of the pair
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int i = 0; i < 4; ++i) { ae[e][i] = make_float2(0.f, 0.f); ao[e][i] = make_float2(0.f, 0.f); }
    const float* drow = dys + 4 * t;
    const float* wrow = wd + 8 * h;
#pragma unroll 2
    for (int o = 0; o < CO2; ++o) {
        float d[6];
        {
            const float4 v = *reinterpret_cast<const float4*>(drow + o * NMP);
            const float2 v2 = *reinterpret_cast<const float2*>(drow + o * NMP + 4);
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; d[4] = v2.x; d[5] = v2.y;
        }
        float2 wk[5][4];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float4 v0 = *reinterpret_cast<const float4*>(wrow + (o * 5 + k) * 16);
            const float4 v1 = *reinterpret_cast<const float4*>(wrow + (o * 5 + k) * 16 + 4);
            wk[k][0] = make_float2(v0.x, v0.y); wk[k][1] = make_float2(v0.z, v0.w);
            wk[k][2] = make_float2(v1.x, v1.y); wk[k][3] = make_float2(v1.z, v1.w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // pair u = u0 + 4t + e: d[e] = dy[u - 1], d[e + 1] = dy[u], d[e + 2] = dy[u + 1]
            const float2 dm = make_float2(d[e], d[e]), dc = make_float2(d[e + 1], d[e + 1]), dp = make_float2(d[e + 2], d[e + 2]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ae[e][i] = __ffma2_rn(wk[0][i], dp, ae[e][i]);
                ae[e][i] = __ffma2_rn(wk[2][i], dc, ae[e][i]);
                ae[e][i] = __ffma2_rn(wk[4][i], dm, ae[e][i]);
                ao[e][i] = __ffma2_rn(wk[1][i], dp, ao[e][i]);
                ao[e][i] = __ffma2_rn(wk[3][i], dc, ao[e][i]);
            }
        }
    }
    // dp1[b][ci][2 u0 + 8 t + 2 e + {0, 1}]: 8 consecutive positions per thread and channel
    const int j = 2 * u0 + 8 * t;
    const bool vec = (Lin & 3) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && j + 7 < Lin;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int ci = 8 * h + 2 * i + hh;
            float* dst = dx + ((size_t)b * 16 + ci) * Lin + j;
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                v[2 * e] = hh ? ae[e][i].y : ae[e][i].x;
                v[2 * e + 1] = hh ? ao[e][i].y : ao[e][i].x;
            }
            if (vec) {
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (j + e < Lin) dst[e] = v[e];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// conv1_wgrad_dgate_kernel.  grid = (CS, B), cluster = (CS, 1, 1), block = 32 * NW (NW = min(C, 8) warps).
// CTA `rank` owns the positions [rank * chunk, rank * chunk + chunk) of its batch row (chunk a multiple of 32, <= W1_MAXCH).
// dynamic smem: xs [C][2 chunk + 8] | dyT [chunk][20] (dy1 transposed, 16 + 4 padding) | s_G [C * 112]
constexpr int W1_MAXCH = 512, W1_DLD = 20;

__global__ void __launch_bounds__(256) conv1_wgrad_dgate_kernel(const float* __restrict__ x, const float* __restrict__ dyn,
                                                                const float* __restrict__ w, const float* __restrict__ gate,
                                                                const float* __restrict__ mean, const float* __restrict__ ca_w1,
                                                                const float* __restrict__ ca_w2, int C, int A, int T, int Lout,
                                                                int chunk, float* __restrict__ dw, float* __restrict__ dca_w1,
                                                                float* __restrict__ dca_w2, const BnBwd bn) {
    extern __shared__ __align__(16) float w1_smem[];
    const int XS = 2 * chunk + 8;
    float* xs = w1_smem;
    float* dyT = xs + C * XS;
    float* s_G = dyT + chunk * W1_DLD;
    __shared__ float s_bn[16][5];
    __shared__ float s_dgp[16], s_dg[16], s_hid[4], s_dh[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NT = blockDim.x, NW = NT >> 5;
    const int b = blockIdx.y;
    const uint32_t rank = cl_rank(), CS = gridDim.x;
    const int l_lo = (int)rank * chunk;
    const int nl = max(0, min(chunk, Lout - l_lo));
    if (bn.y) {
        bn_bwd_constants<16>(bn, Lout, s_bn);
        if (rank == 0 && b == 0 && tid < 16) {       // dgamma / dbeta once per launch
            if (bn.dgamma) bn.dgamma[tid] += bn.grad_scale * (float)bn.red[16 + tid];
            if (bn.dbeta) bn.dbeta[tid] += bn.grad_scale * (float)bn.red[tid];
        }
    }
    if (tid < 16) s_dgp[tid] = 0.f;
    // x tile: columns [2 l_lo - 4, 2 l_lo + 2 chunk + 4) of every channel (zeros outside the row); T % 4 == 0
    const float* xb = x + (size_t)b * C * T;
    for (int idx = tid; idx < C * (XS / 4); idx += NT) {
        const int c = idx / (XS / 4), q = idx - c * (XS / 4);
        const int tcol = 2 * l_lo - 4 + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tcol >= 0 && tcol + 3 < T) v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)c * T + tcol));
        *reinterpret_cast<float4*>(xs + c * XS + 4 * q) = v;
    }
    __syncthreads();        // s_bn
    // dy1 tile, transposed: dyT[ll][o], BatchNorm backward applied on the way (BnBwd)
    const float* dyb = dyn + (size_t)b * 16 * Lout;
    const float* yb = bn.y ? bn.y + (size_t)b * 16 * Lout : nullptr;
    for (int idx = tid; idx < 16 * chunk; idx += NT) {
        const int o = idx / chunk, ll = idx - o * chunk;
        float d = 0.f;
        if (ll < nl) {
            d = __ldg(dyb + (size_t)o * Lout + l_lo + ll);
            if (yb) d = s_bn[o][0] * (d - s_bn[o][3] - (__ldg(yb + (size_t)o * Lout + l_lo + ll) - s_bn[o][1]) * s_bn[o][2] * s_bn[o][4]);
        }
        dyT[ll * W1_DLD + o] = d;
    }
    __syncthreads();

    for (int c = warp; c < C; c += NW) {
        float2 acc[8][7];
#pragma unroll
        for (int o2 = 0; o2 < 8; ++o2)
#pragma unroll
            for (int k = 0; k < 7; ++k) acc[o2][k] = make_float2(0.f, 0.f);
        const float* xr = xs + c * XS;
        for (int ll = lane; ll < nl; ll += 32) {
            float2 d2[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(dyT + ll * W1_DLD + 4 * q);
                d2[2 * q] = make_float2(v.x, v.y);
                d2[2 * q + 1] = make_float2(v.z, v.w);
            }
            float xw[8];        // x[c][2 l + k - 3] = xs[2 ll + k + 1]
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 v = *reinterpret_cast<const float2*>(xr + 2 * ll + 2 * q);
                xw[2 * q] = v.x; xw[2 * q + 1] = v.y;
            }
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float2 xv = make_float2(xw[k + 1], xw[k + 1]);
#pragma unroll
                for (int o2 = 0; o2 < 8; ++o2) acc[o2][k] = __ffma2_rn(d2[o2], xv, acc[o2][k]);
            }
        }
        // sum over the 32 lanes: 112 values, flattened as a[o * 7 + k]; after four halvings (112 -> 7) lane L holds the
        // values of output channel o = L >> 1 (both lanes of a pair after the last exchange)
        float a[112];
#pragma unroll
        for (int o2 = 0; o2 < 8; ++o2)
#pragma unroll
            for (int k = 0; k < 7; ++k) { a[(2 * o2) * 7 + k] = acc[o2][k].x; a[(2 * o2 + 1) * 7 + k] = acc[o2][k].y; }
        int n = 112;
#pragma unroll
        for (int msk = 16; msk >= 2; msk >>= 1) {
            n >>= 1;
            const bool upper = (lane & msk) != 0;
#pragma unroll
            for (int i = 0; i < 56; ++i) {
                if (i < n) {
                    const float keep = upper ? a[i + n] : a[i];
                    const float send = upper ? a[i] : a[i + n];
                    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, msk);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], 1);
        if ((lane & 1) == 0) {
            const int o = lane >> 1;
#pragma unroll
            for (int k = 0; k < 7; ++k) s_G[(c * 16 + o) * 7 + k] = a[k];
        }
    }
    __syncthreads();
    cl_sync();                                            // #1: every CTA's partial G is published
    // rank r finishes the slice [r * per, r * per + per) of the C * 112 values: sums the partials of all ranks
    {
        const int total = C * 112, per = (total + (int)CS - 1) / (int)CS;
        const int e0 = (int)rank * per, e1 = min(total, e0 + per);
        for (int e = e0 + tid; e < e1; e += NT) {
            float G = 0.f;
            for (uint32_t r = 0; r < CS; ++r) G += cl_ld_f32(s_G + e, r);
            const int c = e / 112, rem = e - c * 112, o = rem / 7, k = rem - o * 7;
            const int widx = (o * C + c) * 7 + k;
            const float g = gate ? __ldg(gate + b * C + c) : 1.f;
            atomicAdd(dw + widx, g * G);
            if (gate) atomicAdd(&s_dgp[c], __ldg(w + widx) * G);
        }
    }
    __syncthreads();
    cl_sync();                                            // #2: dgate partials published; nobody reads s_G any more
    if (gate && A > 0 && rank == 0) {
        // ChannelAttention parameter gradients of this batch row (reverse of models.py:28-31)
        if (tid < C) {
            float dg = 0.f;
            for (uint32_t r = 0; r < CS; ++r) dg += cl_ld_f32(&s_dgp[tid], r);
            const float g = __ldg(gate + b * C + tid);
            s_dg[tid] = dg * g * (1.f - g);               // d(pre-sigmoid)
        }
        __syncthreads();
        if (tid < A) {
            float hsum = 0.f, dh = 0.f;
            for (int c = 0; c < C; ++c) {
                hsum += __ldg(ca_w1 + tid * C + c) * __ldg(mean + b * C + c);
                dh += s_dg[c] * __ldg(ca_w2 + c * A + tid);
            }
            const float hr = fmaxf(hsum, 0.f);
            s_hid[tid] = hr;
            s_dh[tid] = hr > 0.f ? dh : 0.f;
        }
        __syncthreads();
        for (int e = tid; e < C * A; e += NT) {
            const int c2 = e / A, a2 = e - c2 * A;        // dw2[c2, a2]
            const int a1 = e / C, c1 = e - a1 * C;        // dw1[a1, c1]
            atomicAdd(dca_w2 + e, s_dg[c2] * s_hid[a2]);
            atomicAdd(dca_w1 + e, s_dh[a1] * __ldg(mean + b * C + c1));
        }
    }
    cl_sync();                                            // #3: no CTA leaves while a peer may still read its shared memory
}

// ---- host side -------------------------------------------------------------------------------------------------------
template <int C, int TI, bool TM>
static int pool_bwd_tile_launch(const float* y, const double* stats, const float* gamma, const float* beta, const float* rm,
                                const float* rv, const float* dout, int B, int Bstat, int Lin, int training, float* dy, double* red,
                                cudaStream_t st) {
    const int Lout = pool_out_len(Lin);
    const size_t smem = (size_t)(C * (TI + 8) + C * (TI / 2 + 1)) * sizeof(float);
    auto kern = pool_relu_bwd_tile_kernel<C, TI, TM>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) MMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    dim3 grid(cdiv(Lin, TI), B);
    MMS_PROF_BEGIN(st);
    kern<<<grid, 256, smem, st>>>(y, stats, gamma, beta, rm, rv, dout, Bstat, Lin, Lout, training, dy, red);
    MMS_LAUNCH_CHECK("pool_relu_bwd_tile_kernel");
    return MMS_OK;
}

bool pool_bwd_tile_supported(int C, int time_major) { return (C == 16 && !time_major) || ((C == 16 || C == 32 || C == 64) && time_major); }

// pool / ReLU backward + the two BN reductions (the `which & 1` pass of launch_bn_relu_pool_bwd), tile version
int launch_pool_relu_bwd_tile(const float* y, const double* stats, const float* gamma, const float* beta, const float* rm,
                              const float* rv, const float* dout, int B, int C, int Lin, int training, int time_major, float* dy,
                              double* red, cudaStream_t st, int Bstat) {
    MMS_REQUIRE(pool_bwd_tile_supported(C, time_major), "pool_relu_bwd_tile: unsupported channel count %d", C);
    if (Bstat <= 0) Bstat = B;
    if (!time_major) return pool_bwd_tile_launch<16, 256, false>(y, stats, gamma, beta, rm, rv, dout, B, Bstat, Lin, training, dy, red, st);
    if (C == 16) return pool_bwd_tile_launch<16, 128, true>(y, stats, gamma, beta, rm, rv, dout, B, Bstat, Lin, training, dy, red, st);
    if (C == 32) return pool_bwd_tile_launch<32, 128, true>(y, stats, gamma, beta, rm, rv, dout, B, Bstat, Lin, training, dy, red, st);
    return pool_bwd_tile_launch<64, 64, true>(y, stats, gamma, beta, rm, rv, dout, B, Bstat, Lin, training, dy, red, st);
}

template <int CO2>
static int conv2_dgrad_launch(const float* dyn, const float* w, int B, int Lin, float* dx, cudaStream_t st, const BnBwd& bn) {
    const int Lout = conv_out_len(Lin, CONV2_K, CONV2_S, CONV2_P);
    const size_t smem = (size_t)(CO2 * 80 + 2 * CO2 * 132) * sizeof(float);
    auto kern = conv2_dgrad_kernel<CO2>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) MMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    dim3 grid(cdiv(Lin, 256), B);
    MMS_PROF_BEGIN(st);
    kern<<<grid, 64, smem, st>>>(dyn, w, dx, Lin, Lout, bn);
    MMS_LAUNCH_CHECK("conv2_dgrad_kernel");
    return MMS_OK;
}

int launch_conv2_dgrad_v3(const float* dyn, const float* w, int B, int O, int Lin, float* dx, cudaStream_t st, const BnBwd* bnp) {
    static const BnBwd none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 1.f};
    const BnBwd& bn = bnp ? *bnp : none;
    if (O == 16) return conv2_dgrad_launch<16>(dyn, w, B, Lin, dx, st, bn);
    if (O == 32) return conv2_dgrad_launch<32>(dyn, w, B, Lin, dx, st, bn);
    MMS_REQUIRE(O == 64, "conv2_dgrad: C_out %d not in {16,32,64}", O);
    return conv2_dgrad_launch<64>(dyn, w, B, Lin, dx, st, bn);
}

bool conv1_wgrad_dgate_supported(const float* x, int C, int T) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) || T % 8 != 0 || T < 16 || C < 1 || C > 16) return false;
    return T / 2 <= 8 * W1_MAXCH;
}

// Weight gradient of conv1 (dw +=), and -- with gate != nullptr -- the ChannelAttention parameter gradients (dca_w1, dca_w2 +=).
int launch_conv1_wgrad_dgate(const float* x, const float* dyn, const float* w, const float* gate, const float* mean,
                             const float* ca_w1, const float* ca_w2, int B, int C, int T, float* dw, float* dca_w1, float* dca_w2,
                             cudaStream_t st, const BnBwd* bnp) {
    static const BnBwd none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 1.f};
    const BnBwd& bn = bnp ? *bnp : none;
    MMS_REQUIRE(conv1_wgrad_dgate_supported(x, C, T), "conv1_wgrad_dgate: unsupported shape / alignment");
    const int L1 = T / 2;
    int CS = cdiv(L1, 256);
    if (CS > 8) CS = 8;
    if (CS < 1) CS = 1;
    const int chunk = ((cdiv(L1, CS) + 31) / 32) * 32;
    MMS_REQUIRE(chunk <= W1_MAXCH, "conv1_wgrad_dgate: sequence too long (chunk %d)", chunk);
    const int NW = C < 8 ? C : 8;
    const size_t smem = (size_t)(C * (2 * chunk + 8) + chunk * W1_DLD + C * 112) * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.need()) MMS_CUDA(cudaFuncSetAttribute(conv1_wgrad_dgate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    MMS_REQUIRE(smem <= 160 * 1024, "conv1_wgrad_dgate: shared memory %zu too large", smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, B);
    cfg.blockDim = dim3(32 * NW);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MMS_PROF_BEGIN(st);
    MMS_CUDA(cudaLaunchKernelEx(&cfg, conv1_wgrad_dgate_kernel, x, dyn, w, gate, mean, ca_w1, ca_w2, C, C / 4, T, L1, chunk, dw, dca_w1,
                                dca_w2, bn));
    MMS_LAUNCH_CHECK("conv1_wgrad_dgate_kernel");
    return MMS_OK;
}

}  // namespace mms
