// Fused forward kernels of the CNN encoder (reference models.py:24-31,45-53), the default path when T % 8 == 0:
//
//   attn_conv1_fwd_kernel   ChannelAttention (squeeze, excite MLP, sigmoid gate; models.py:24-31) + Conv1d(C,16,k7,s2,p3)
//                           (models.py:46) + the BatchNorm batch sums of its output, ONE launch.  A thread-block CLUSTER owns
//                           a batch row: every CTA sums the channels of its own x tile, the partial sums meet through
//                           distributed shared memory (mapa / ld.shared::cluster), every CTA evaluates the C x C/4 MLP
//                           itself and applies the gate while it forms the FFMA2 operands.  The BN sums are reduced over the
//                           cluster first: one float64 atomic per channel and ROW instead of per CTA.
//   bn_pool_conv2_fwd_kernel BatchNorm1d(16)+ReLU+MaxPool1d(3,2,1) (models.py:47-49) + Conv1d(16,C_out,k5,s2,p2)
//                           (models.py:50) + the batch sums of ITS output, one launch: the pooled tile never leaves shared
//                           memory on its way into the convolution (it is also written out once: conv2's weight gradient
//                           reads it).
// Both stage their input tile with TMA (cp.async.bulk.tensor, zero fill outside the tensor = the convolution's padding,
// no boundary code, no per-thread address arithmetic) into 256-column panels and compute on the fp32 pipes with packed
// fma.rn.f32x2: a thread owns P consecutive output positions x 16 output channels (8 float2 accumulators per position), so
// one 128-bit broadcast load of weights feeds P x 2 FFMA2 and the sliding input window of its positions is loaded once per
// input channel.  The previous kernels spent 3 instructions per FMA (one shared-memory load per 3-4 FMAs; ncu: 7.8 M warp
// instructions for 2.6 M warp FMAs, 55 % issue-bound); these need ~1.2 per packed pair.
// Why not tcgen05 here (conv_tc.cu stays opt-in): with K = 7C = 42 / 80 the tensor work is < 0.2 us; building the im2col
// operand in the canonical MN-major 3xTF32 layout costs ~6 instructions per operand element and position -- as many issue
// slots as computing the convolution directly on the FMA pipes (profiles/r1_conv_tc_vs_simt.md).
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mms {

constexpr int CF_PANEL = 256;       // TMA box width (the hardware limit of a box dimension)

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void* smem_ptr, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(smem_ptr)), "r"(rank));
    return a;
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double ld_cluster_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }

// ------------------------------------------------------------------------------------------------------------------
// attn_conv1_fwd_kernel.  grid = (CS, B), cluster = (CS, 1, 1), block = 64.  CTA `rank` of a cluster handles the position
// tiles rank * tpc .. rank * tpc + tpc - 1 of batch row blockIdx.y (256 positions each, 4 per thread).
// x tile of a position tile starting at l0: columns [2 l0 - 4, 2 l0 + 516) of the row (the first needed column is 2 l0 - 3;
// one more to the left keeps the TMA start coordinate a multiple of 4 floats), as panels [C][256] | [C][256] | [C][8].
// dynamic smem: tiles [tpc][C * 520] | w1s [C * 7][16]
constexpr int A1_TL = 256, A1_P = 4, A1_NT = A1_TL / A1_P, A1_COLS = 2 * A1_TL + 8, A1_TAIL = A1_COLS - 2 * CF_PANEL;
constexpr int A1_MAX_TPC = 4;

__device__ __forceinline__ int a1_off(int i, int c, int C) {      // float offset of local column i of channel c inside a tile
    return i < 2 * CF_PANEL ? (((i >> 8) * C + c) << 8) + (i & 255) : 2 * CF_PANEL * C + c * A1_TAIL + (i - 2 * CF_PANEL);
}

__global__ void __launch_bounds__(A1_NT) attn_conv1_fwd_kernel(const __grid_constant__ CUtensorMap map_panel,
                                                               const __grid_constant__ CUtensorMap map_tail,
                                                               const float* __restrict__ w, const float* __restrict__ ca_w1,
                                                               const float* __restrict__ ca_w2, int use_gate, int C, int A, int T,
                                                               int Lout, int tpc, float* __restrict__ mean_out,
                                                               float* __restrict__ gate_out, float* __restrict__ y,
                                                               double* __restrict__ stats, const float* __restrict__ w2, int O2,
                                                               float* __restrict__ w2_fwd, float* __restrict__ w2_bwd) {
    extern __shared__ __align__(128) float a1_smem[];
    __shared__ __align__(8) uint64_t load_bar;
    __shared__ float s_wsum[2][16], s_part[16], s_all[8 * 16], s_mean[16], s_gate[16];
    __shared__ float s_red[2][32];
    __shared__ __align__(8) double s_stat[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const uint32_t rank = cluster_ctarank(), CS = gridDim.x;
    const int tile_bytes = C * A1_COLS * 4;
    const int tile_floats = (C * A1_COLS + 31) & ~31;     // tile stride: every TMA destination stays 128-byte aligned
    float* w1s = a1_smem + tpc * tile_floats;
    const int ntiles = (Lout + A1_TL - 1) / A1_TL;
    const int g0 = (int)rank * tpc;                       // first position tile of this CTA
    int mine = ntiles - g0;
    mine = mine < 0 ? 0 : (mine > tpc ? tpc : mine);      // tiles this CTA really has

    if (tid == 0) {
        mbar_init(&load_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && mine > 0) {
        mbar_expect_tx(&load_bar, (uint32_t)(mine * tile_bytes));
        for (int t = 0; t < mine; ++t) {
            float* dst = a1_smem + t * tile_floats;
            const int col0 = 2 * (g0 + t) * A1_TL - 4;
            tma_load_2d(&map_panel, &load_bar, dst, col0, b * C);
            tma_load_2d(&map_panel, &load_bar, dst + C * CF_PANEL, col0 + CF_PANEL, b * C);
            tma_load_2d(&map_tail, &load_bar, dst + 2 * C * CF_PANEL, col0 + 2 * CF_PANEL, b * C);
        }
    }
    // weights while the tiles are in flight: w1s[(c*7 + k)*16 + o] = w[o][c][k]
    for (int idx = tid; idx < 16 * C * 7; idx += A1_NT) {
        const int o = idx / (C * 7), ck = idx - o * (C * 7);
        w1s[ck * 16 + o] = __ldg(w + idx);
    }
    if (tid < 16) { s_part[tid] = 0.f; s_gate[tid] = 1.f; }
    // conv2's weights w2[o][ci][k] in the orders the NEXT kernels stage them with one bulk copy per CTA (they follow this
    // launch on the stream): [ci*5 + k][o] for bn_pool_conv2_fwd_kernel, [o*5 + k][ci] for conv2_bwd_kernel.  Rank 0 of every
    // cluster takes a 64-element slice while its tiles are in flight.
    if (w2_fwd && rank == 0) {
        for (int idx = b * A1_NT + tid; idx < O2 * 80; idx += (int)gridDim.y * A1_NT) {
            const float v = __ldg(w2 + idx);
            const int o = idx / 80, r = idx - o * 80, ci = r / 5, k = r - ci * 5;
            w2_fwd[r * O2 + o] = v;
            if (w2_bwd) w2_bwd[(o * 5 + k) * 16 + ci] = v;
        }
    }
    if (mine > 0) mbar_wait(&load_bar, 0);
    __syncthreads();

    if (use_gate) {
        // squeeze: channel sums over the columns this CTA owns (local 4 .. 515 of each tile; columns beyond T arrive as zeros)
        for (int c = 0; c < C; ++c) {
            float s = 0.f;
            for (int t = 0; t < mine; ++t) {
                const float* tl = a1_smem + t * tile_floats;
                const int i = 4 + 8 * tid;
                const float4 v0 = *reinterpret_cast<const float4*>(tl + a1_off(i, c, C));
                const float4 v1 = *reinterpret_cast<const float4*>(tl + a1_off(i + 4, c, C));
                s += ((v0.x + v0.y) + (v0.z + v0.w)) + ((v1.x + v1.y) + (v1.z + v1.w));
            }
            s = warp_sum(s);
            if (lane == 0) s_wsum[warp][c] = s;
        }
        __syncthreads();
        if (tid < C) s_part[tid] = s_wsum[0][tid] + s_wsum[1][tid];
        __syncthreads();
    }
    cluster_sync_all();                                   // #1: every CTA's partial sums are published
    if (use_gate) {
        for (int idx = tid; idx < (int)CS * C; idx += A1_NT) {
            const int r = idx / C, c = idx - r * C;
            s_all[idx] = ld_cluster_f32(map_to_rank(&s_part[c], (uint32_t)r));
        }
        __syncthreads();
        if (tid < C) {
            float tot = 0.f;
            for (int r = 0; r < (int)CS; ++r) tot += s_all[r * C + tid];      // rank order: every CTA gets the same bits
            s_mean[tid] = tot / (float)T;
        }
        __syncthreads();
        if (tid < C) {      // excite MLP, evaluated by thread c for its own channel (A = C/4 <= 4 hidden units)
            float z = 0.f;
            for (int a = 0; a < A; ++a) {
                float h = 0.f;
                for (int c2 = 0; c2 < C; ++c2) h += __ldg(ca_w1 + a * C + c2) * s_mean[c2];
                z += __ldg(ca_w2 + tid * A + a) * fmaxf(h, 0.f);
            }
            const float g = sigmoid_f(z);                 // A == 0 -> sigmoid(0) = 0.5
            s_gate[tid] = g;
            if (rank == 0) {
                gate_out[b * C + tid] = g;
                mean_out[b * C + tid] = s_mean[tid];
            }
        }
        __syncthreads();
    }

    double stat_acc = 0.0;                                // threads 0..31: value `tid` of (16 sums | 16 sums of squares)
    for (int t = 0; t < mine; ++t) {
        const float* tl = a1_smem + t * tile_floats;
        const int l0 = (g0 + t) * A1_TL;
        // local columns of this thread's window: positions l0 + 4 tid + pp, tap k -> local column 8 tid + 2 pp + k + 1
        int off[4], cs[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = 8 * tid + 4 * q;
            off[q] = a1_off(i, 0, C);
            cs[q] = i < 2 * CF_PANEL ? CF_PANEL : A1_TAIL;
        }
        float2 acc[A1_P][8];
#pragma unroll
        for (int pp = 0; pp < A1_P; ++pp)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[pp][i] = make_float2(0.f, 0.f);
        for (int c = 0; c < C; ++c) {
            const float g = s_gate[c];
            float xw[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(tl + off[q] + c * cs[q]);
                xw[4 * q + 0] = v.x * g; xw[4 * q + 1] = v.y * g; xw[4 * q + 2] = v.z * g; xw[4 * q + 3] = v.w * g;
            }
            const float4* wc = reinterpret_cast<const float4*>(w1s + c * 7 * 16);
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                float2 wv[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = wc[k * 4 + q];
                    wv[2 * q] = make_float2(v.x, v.y);
                    wv[2 * q + 1] = make_float2(v.z, v.w);
                }
#pragma unroll
                for (int pp = 0; pp < A1_P; ++pp) {
                    const float2 xv = dup2(xw[2 * pp + k + 1]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[pp][i] = __ffma2_rn(wv[i], xv, acc[pp][i]);
                }
            }
        }
        // epilogue: 4 consecutive positions per channel = one 128-bit store (Lout % 4 == 0: a thread's positions are all
        // inside or all outside the row)
        const int l = l0 + A1_P * tid;
        const bool valid = l < Lout;
        if (valid) {
            float* yb = y + (size_t)b * 16 * Lout + l;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                *reinterpret_cast<float4*>(yb + (size_t)(2 * i) * Lout) = make_float4(acc[0][i].x, acc[1][i].x, acc[2][i].x, acc[3][i].x);
                *reinterpret_cast<float4*>(yb + (size_t)(2 * i + 1) * Lout) = make_float4(acc[0][i].y, acc[1][i].y, acc[2][i].y, acc[3][i].y);
            }
        }
        if (stats) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float sx = 0.f, sy = 0.f, qx = 0.f, qy = 0.f;
#pragma unroll
                for (int pp = 0; pp < A1_P; ++pp) {
                    sx += acc[pp][i].x; sy += acc[pp][i].y;
                    qx = fmaf(acc[pp][i].x, acc[pp][i].x, qx); qy = fmaf(acc[pp][i].y, acc[pp][i].y, qy);
                }
                v[2 * i] = valid ? sx : 0.f; v[2 * i + 1] = valid ? sy : 0.f;
                v[16 + 2 * i] = valid ? qx : 0.f; v[16 + 2 * i + 1] = valid ? qy : 0.f;
            }
            warp_transpose_reduce<32>(v, lane);
            s_red[warp][lane] = v[0];
            __syncthreads();
            if (tid < 32) stat_acc += (double)s_red[0][tid] + (double)s_red[1][tid];
            __syncthreads();
        }
    }
    if (tid < 32) s_stat[tid] = stat_acc;
    __syncthreads();
    cluster_sync_all();                                   // #2: every CTA's batch sums are published (and nobody reads s_part any more)
    if (stats && rank == 0 && tid < 32) {
        double tot = 0.0;
        for (int r = 0; r < (int)CS; ++r) tot += ld_cluster_f64(map_to_rank(&s_stat[tid], (uint32_t)r));
        atomicAdd(stats + tid, tot);                      // stats[0..15] = sums, stats[16..31] = sums of squares
    }
    cluster_sync_all();                                   // #3: no CTA leaves while rank 0 still reads its shared memory
}

// ------------------------------------------------------------------------------------------------------------------
// bn_pool_conv2_fwd_kernel.  grid = (ceil(L2 / TM), B), block = (TM / P) * (CO2 / 16).
// CTA tile: conv2 outputs m in [m0, m0 + TM); pooled inputs j in [2 m0 - 2, 2 m0 + 2 TM + 1) (local jj = j - 2 m0 + 2);
// conv1 outputs i in [4 m0 - 5, 4 m0 + 4 TM + 2), staged from column 4 m0 - 8 (local ii = i - 4 m0 + 8; window of pooled
// element jj = local columns 2 jj + 3 .. 2 jj + 5).
// dynamic smem: y1 panels [nfull][16][256] | tail [16][TAILW] | p1s [16][NJP] | w2s [16 * 5][CO2]
template <int CO2, int TM, int P>
__global__ void __launch_bounds__((TM / P) * (CO2 / 16)) bn_pool_conv2_fwd_kernel(
    const __grid_constant__ CUtensorMap map_panel, const __grid_constant__ CUtensorMap map_tail, const double* __restrict__ stats1,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* rm, float* rv, int64_t* nbt, int Bstat, int training,
    const float* __restrict__ w, int w_arranged, int L1, int P1, int L2, float* __restrict__ p1_out, float* __restrict__ y2,
    double* __restrict__ stats2) {
    constexpr int NPG = TM / P, NCG = CO2 / 16, NT = NPG * NCG;
    constexpr int NI = 4 * TM + 12, NFULL = NI / CF_PANEL, TAILW = NI % CF_PANEL;
    constexpr int NJ = 2 * TM + 3, NJP = 2 * TM + 4;
    constexpr int NW = (2 * P + 3 + 3) / 4;               // float4 loads that cover a thread's window of 2P + 3 pooled values
    static_assert(NPG % 32 == 0, "a warp must not mix channel groups");
    static_assert(TAILW % 4 == 0 && TAILW > 0, "tail box");
    extern __shared__ __align__(128) float c2_smem[];
    __shared__ __align__(8) uint64_t load_bar;
    __shared__ float s_a[16], s_b[16];
    __shared__ float s_red[NT / 32][32];
    float* ys = c2_smem;
    float* p1s = ys + 16 * NI;
    float* w2s = p1s + 16 * NJP;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, m0 = blockIdx.x * TM;
    if (tid == 0) {
        mbar_init(&load_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&load_bar, (uint32_t)(16 * NI * 4) + (w_arranged ? (uint32_t)(CO2 * 80 * 4) : 0u));
        const int col0 = 4 * m0 - 8;
#pragma unroll
        for (int p = 0; p < NFULL; ++p) tma_load_2d(&map_panel, &load_bar, ys + p * 16 * CF_PANEL, col0 + p * CF_PANEL, b * 16);
        tma_load_2d(&map_tail, &load_bar, ys + NFULL * 16 * CF_PANEL, col0 + NFULL * CF_PANEL, b * 16);
        if (w_arranged) bulk_g2s(w2s, w, (uint32_t)(CO2 * 80 * 4), &load_bar);     // already [ci*5 + k][o] (conv2_w_relayout_fwd_kernel)
    }
    const double n1 = (double)Bstat * (double)L1;
    if (tid < 16) {
        const BnAffine af = bn_affine(training, stats1, gamma, beta, rm, rv, tid, 16, n1);
        s_a[tid] = af.a;
        s_b[tid] = af.b;
    }
    // w2s[(ci*5 + k)*CO2 + o] = w[o][ci][k]
    if (!w_arranged) {
        for (int idx = tid; idx < CO2 * 80; idx += NT) {
            const int o = idx / 80, ck = idx - o * 80;
            w2s[ck * CO2 + o] = __ldg(w + idx);
        }
    }
    mbar_wait(&load_bar, 0);
    __syncthreads();

    // BatchNorm + ReLU + MaxPool(3,2,1) of the tile; positions outside [0, P1) are the convolution's zero padding.
    // NT / 16 threads per channel, four pooled values per visit: three 128-bit loads of conv1 outputs (a quad of pooled
    // values jj .. jj+3 needs the local columns 2 jj + 3 .. 2 jj + 11), one 128-bit store of the tile, 64-bit stores of p1.
    {
        constexpr int TPC = NT / 16, NQ = NJP / 4;
        const int c = tid / TPC;
        const float a = s_a[c], bsh = s_b[c];
        float* p1row = p1_out + ((size_t)b * 16 + c) * P1;
        auto ycol = [&](int ii) -> const float* {
            return ii < NFULL * CF_PANEL ? ys + ((((ii >> 8) * 16 + c) << 8) + (ii & 255))
                                         : ys + (NFULL * 16 * CF_PANEL + c * TAILW + (ii - NFULL * CF_PANEL));
        };
        for (int q = tid - c * TPC; q < NQ; q += TPC) {
            const int jj = 4 * q, ii0 = 2 * jj;
            float yv[12];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ii0 + 4 * t < NI) v = *reinterpret_cast<const float4*>(ycol(ii0 + 4 * t));
                yv[4 * t] = v.x; yv[4 * t + 1] = v.y; yv[4 * t + 2] = v.z; yv[4 * t + 3] = v.w;
            }
            float z[9];           // relu(bn(y)) at the local columns ii0 + 3 .. ii0 + 11; -inf outside the row
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int i = 4 * m0 - 8 + ii0 + 3 + t;
                z[t] = (i >= 0 && i < L1) ? fmaxf(fmaf(a, yv[t + 3], bsh), 0.f) : -INFINITY;
            }
            float pv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 2 * m0 - 2 + jj + e;
                pv[e] = (j >= 0 && j < P1) ? fmaxf(fmaxf(z[2 * e], z[2 * e + 1]), z[2 * e + 2]) : 0.f;
            }
            *reinterpret_cast<float4*>(p1s + c * NJP + jj) = make_float4(pv[0], pv[1], pv[2], pv[3]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int je = jj + 2 * h, j = 2 * m0 - 2 + je;       // the pair (je, je + 1): j is even
                if (je >= 2 && je < 2 + 2 * TM && j >= 0 && j + 1 < P1)
                    *reinterpret_cast<float2*>(p1row + j) = make_float2(pv[2 * h], pv[2 * h + 1]);
                else if (je >= 2 && je < 2 + 2 * TM && j >= 0 && j < P1)
                    p1row[j] = pv[2 * h];
            }
        }
    }
    __syncthreads();

    const int cg = tid / NPG, pg = tid - cg * NPG;
    float2 acc[P][8];
#pragma unroll
    for (int pp = 0; pp < P; ++pp)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[pp][i] = make_float2(0.f, 0.f);
    const float* xrow = p1s + 2 * P * pg;                 // window of position pp, tap k: xrow[2 pp + k]
    const float* wrow = w2s + cg * 16;
#pragma unroll 2
    for (int ci = 0; ci < 16; ++ci) {
        float xw[4 * NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(xrow + ci * NJP + 4 * q);
            xw[4 * q + 0] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            float2 wv[8];
            const float4* w4 = reinterpret_cast<const float4*>(wrow + (ci * 5 + k) * CO2);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = w4[q];
                wv[2 * q] = make_float2(v.x, v.y);
                wv[2 * q + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int pp = 0; pp < P; ++pp) {
                const float2 xv = dup2(xw[2 * pp + k]);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[pp][i] = __ffma2_rn(wv[i], xv, acc[pp][i]);
            }
        }
    }
    // epilogue
    const int m = m0 + P * pg;
    float* yb = y2 + ((size_t)b * CO2 + cg * 16) * L2 + m;
    if (P == 2 && (L2 & 1) == 0 && m + 1 < L2 && (reinterpret_cast<uintptr_t>(y2) & 7) == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            *reinterpret_cast<float2*>(yb + (size_t)(2 * i) * L2) = make_float2(acc[0][i].x, acc[P - 1][i].x);
            *reinterpret_cast<float2*>(yb + (size_t)(2 * i + 1) * L2) = make_float2(acc[0][i].y, acc[P - 1][i].y);
        }
    } else {
#pragma unroll
        for (int pp = 0; pp < P; ++pp) {
            if (m + pp < L2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    yb[(size_t)(2 * i) * L2 + pp] = acc[pp][i].x;
                    yb[(size_t)(2 * i + 1) * L2 + pp] = acc[pp][i].y;
                }
            }
        }
    }
    if (stats2) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float sx = 0.f, sy = 0.f, qx = 0.f, qy = 0.f;
#pragma unroll
            for (int pp = 0; pp < P; ++pp) {
                const bool ok = m + pp < L2;
                const float ax = ok ? acc[pp][i].x : 0.f, ay = ok ? acc[pp][i].y : 0.f;
                sx += ax; sy += ay;
                qx = fmaf(ax, ax, qx); qy = fmaf(ay, ay, qy);
            }
            v[2 * i] = sx; v[2 * i + 1] = sy; v[16 + 2 * i] = qx; v[16 + 2 * i + 1] = qy;
        }
        warp_transpose_reduce<32>(v, lane);
        s_red[warp][lane] = v[0];
        __syncthreads();
        if (tid < 2 * CO2) {
            const int which = tid / CO2, o = tid - which * CO2, g = o >> 4, oc = o & 15;
            double t = 0.0;
#pragma unroll
            for (int wq = 0; wq < NPG / 32; ++wq) t += (double)s_red[g * (NPG / 32) + wq][which * 16 + oc];
            atomicAdd(stats2 + which * CO2 + o, t);
        }
    }
    if (training && blockIdx.x == 0 && blockIdx.y == 0) bn_running_update(stats1, rm, rv, nbt, 16, n1, tid);
}

// wt[(ci*5 + k) * O + o] = w[o][ci][k]: conv2's weights in the order bn_pool_conv2_fwd_kernel keeps them in shared memory
__global__ void __launch_bounds__(256) conv2_w_relayout_fwd_kernel(const float* __restrict__ w, int O, float* __restrict__ wt) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= O * 80) return;
    const int o = idx / 80, ck = idx - o * 80;
    wt[ck * O + o] = __ldg(w + idx);
}

// ---- host side -------------------------------------------------------------------------------------------------------
int launch_conv2_w_relayout_fwd(const float* w, int O, float* wt, cudaStream_t st) {
    MMS_PROF_BEGIN(st);
    conv2_w_relayout_fwd_kernel<<<cdiv(O * 80, 256), 256, 0, st>>>(w, O, wt);
    MMS_LAUNCH_CHECK("conv2_w_relayout_fwd_kernel");
    return MMS_OK;
}

bool conv_fused_supported(const float* x, int C, int T, int O) {
    if (!encode_tiled_fn() || (reinterpret_cast<uintptr_t>(x) & 15) || T % 8 != 0 || T < 16 || C < 1 || C > 16) return false;
    if (!(O == 16 || O == 32 || O == 64)) return false;
    const int L1 = T / 2, ntiles = (L1 + A1_TL - 1) / A1_TL, tpc = (ntiles + 7) / 8;
    return tpc <= A1_MAX_TPC;
}

// ChannelAttention + conv1 (+ BN batch sums).  use_gate == 0: plain convolution (cnn_gru baseline).
// w2 / w2_fwd / w2_bwd (optional): conv2's weights [O2][16][5] and where to leave their re-arranged copies (w2_bwd: O2 == 32 only)
int launch_attn_conv1_fwd(const float* x, const float* w, const float* ca_w1, const float* ca_w2, int use_gate, int B, int C, int T,
                          float* mean_out, float* gate_out, float* y1, double* stats, const float* w2, int O2, float* w2_fwd,
                          float* w2_bwd, cudaStream_t st) {
    MMS_REQUIRE(!w2_fwd || (w2 && O2 > 0), "attn_conv1_fwd: re-layout requested without weights");
    MMS_REQUIRE(!w2_bwd || (w2_fwd && O2 == 32), "attn_conv1_fwd: the backward order of conv2's weights needs C_out = 32");
    MMS_REQUIRE(conv_fused_supported(x, C, T, 16), "attn_conv1_fwd: unsupported shape / alignment");
    MMS_REQUIRE((reinterpret_cast<uintptr_t>(y1) & 15) == 0, "attn_conv1_fwd: y1 must be 16-byte aligned");
    const int L1 = T / 2, ntiles = (L1 + A1_TL - 1) / A1_TL, tpc = (ntiles + 7) / 8, CS = (ntiles + tpc - 1) / tpc;
    CUtensorMap mp, mt;
    int rc = make_map(&mp, x, (int64_t)B * C, T, T, C, CU_TENSOR_MAP_SWIZZLE_NONE, CF_PANEL);
    if (rc) return rc;
    rc = make_map(&mt, x, (int64_t)B * C, T, T, C, CU_TENSOR_MAP_SWIZZLE_NONE, A1_TAIL);
    if (rc) return rc;
    const size_t smem = (size_t)(tpc * ((C * A1_COLS + 31) & ~31) + C * 7 * 16) * sizeof(float);
    static PerDeviceOnce attr_once;
    if (attr_once.need()) MMS_CUDA(cudaFuncSetAttribute(attn_conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    MMS_REQUIRE(smem <= 160 * 1024, "attn_conv1_fwd: shared memory %zu too large", smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, B);
    cfg.blockDim = dim3(A1_NT);
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
    MMS_CUDA(cudaLaunchKernelEx(&cfg, attn_conv1_fwd_kernel, mp, mt, w, ca_w1, ca_w2, use_gate, C, C / 4, T, L1, tpc, mean_out, gate_out,
                                y1, stats, w2, O2, w2_fwd, w2_bwd));
    MMS_LAUNCH_CHECK("attn_conv1_fwd_kernel");
    return MMS_OK;
}

template <int CO2, int TM, int P>
static int bn_pool_conv2_launch(const float* y1, const double* stats1, const float* gamma, const float* beta, float* rm, float* rv,
                                int64_t* nbt, int Bstat, int training, const float* w, int w_arranged, int B, int L1, float* p1, float* y2,
                                double* stats2, cudaStream_t st) {
    constexpr int NI = 4 * TM + 12, TAILW = NI % CF_PANEL, NJP = 2 * TM + 4, NT = (TM / P) * (CO2 / 16);
    const int P1 = pool_out_len(L1), L2 = conv_out_len(P1, CONV2_K, CONV2_S, CONV2_P);
    CUtensorMap mp, mt;
    int rc = make_map(&mp, y1, (int64_t)B * 16, L1, L1, 16, CU_TENSOR_MAP_SWIZZLE_NONE, CF_PANEL);
    if (rc) return rc;
    rc = make_map(&mt, y1, (int64_t)B * 16, L1, L1, 16, CU_TENSOR_MAP_SWIZZLE_NONE, TAILW);
    if (rc) return rc;
    const size_t smem = (size_t)(16 * NI + 16 * NJP + 80 * CO2) * sizeof(float);
    auto kern = bn_pool_conv2_fwd_kernel<CO2, TM, P>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) MMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    MMS_REQUIRE(smem <= 100 * 1024, "bn_pool_conv2_fwd: shared memory %zu too large", smem);
    dim3 grid(cdiv(L2, TM), B);
    MMS_PROF_BEGIN(st);
    MMS_REQUIRE(!w_arranged || (reinterpret_cast<uintptr_t>(w) & 15) == 0, "bn_pool_conv2_fwd: re-arranged weights must be 16-byte aligned");
    kern<<<grid, NT, smem, st>>>(mp, mt, stats1, gamma, beta, rm, rv, nbt, Bstat, training, w, w_arranged, L1, P1, L2, p1, y2, stats2);
    MMS_LAUNCH_CHECK("bn_pool_conv2_fwd_kernel");
    return MMS_OK;
}

// BN1 + ReLU + MaxPool + conv2 (+ BN2 batch sums, BN1 running statistics).  L1 = T / 2 (multiple of 4).
// w_arranged != 0: `w` is the [ci*5 + k][o] copy made by launch_conv2_w_relayout_fwd (one bulk copy per CTA instead of a gather)
int launch_bn_pool_conv2_fwd(const float* y1, const double* stats1, const float* gamma, const float* beta, float* rm, float* rv,
                             int64_t* nbt, int Bstat, int training, const float* w, int w_arranged, int B, int O, int L1, float* p1,
                             float* y2, double* stats2, cudaStream_t st) {
    MMS_REQUIRE(L1 % 4 == 0 && (reinterpret_cast<uintptr_t>(y1) & 15) == 0, "bn_pool_conv2_fwd: unsupported shape / alignment");
    MMS_REQUIRE(!training || stats1, "bn_pool_conv2_fwd: training mode needs batch statistics");
    if (Bstat <= 0) Bstat = B;
    if (O == 16) return bn_pool_conv2_launch<16, 128, 4>(y1, stats1, gamma, beta, rm, rv, nbt, Bstat, training, w, w_arranged, B, L1, p1, y2, stats2, st);
    if (O == 32) return bn_pool_conv2_launch<32, 128, 2>(y1, stats1, gamma, beta, rm, rv, nbt, Bstat, training, w, w_arranged, B, L1, p1, y2, stats2, st);
    return bn_pool_conv2_launch<64, 128, 4>(y1, stats1, gamma, beta, rm, rv, nbt, Bstat, training, w, w_arranged, B, L1, p1, y2, stats2, st);
}

}  // namespace mms
