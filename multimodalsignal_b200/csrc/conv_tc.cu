// Conv1d (reference models.py:46,50: k7 s2 p3 and k5 s2 p2, bias=False) as an implicit GEMM on Blackwell's tensor cores:
//     y[b, o, l] = sum_{(c,k)} (gate[b,c] * w[o,c,k]) * x[b, c, S*l + k - P]
// One CTA = one batch row x 128 output positions:
//   * the input tile x[b, :, S*l0 - P .. ] arrives by TMA (two boxes: 256 + 16 columns, started on a 16-byte boundary
//     as TMA requires; the zero padding of the convolution is TMA's out-of-bounds fill, so there is no boundary code);
//   * the 128 threads expand it in shared memory into the im2col operand A[position, (c,k)], already split into the
//     3xTF32 hi / lo parts, in the canonical MN-major layout of 32-bit operands (positions contiguous, SWIZZLE_128B with
//     32-byte atoms -- see tc_common.cuh); the gate-scaled weights W^T[(c,k), o] are laid out the same way as operand B;
//   * one thread issues K/8 x 3 tcgen05.mma.kind::tf32 (M = 128 positions, N = C_out, accumulator in TMEM);
//   * epilogue: tcgen05.ld -> one position per thread, coalesced channel-major stores, BatchNorm batch statistics
//     (sum, sum of squares per channel) reduced per warp and added with float64 atomics, as the SIMT kernel does.
// fp32-class accuracy (3xTF32), so the parity tolerances of the SIMT kernel hold unchanged.
#include "tc_common.cuh"

namespace mms {

constexpr int CT_TILE = 128;        // output positions per CTA = UMMA M
constexpr int CT_BOX0 = 256;        // columns of the first TMA box
constexpr int CT_BOX1 = 16;         // columns of the second one (SPAN <= 272)

// byte offset of element (krow r, column j < 32) inside one [rows x 32 floats] MN-major block (Swizzle<2,5,2>)
__device__ __forceinline__ uint32_t mn_off(int r, int j) { return (uint32_t)(r * 128 + ((((j >> 3) ^ (r & 3)) << 5) | ((j & 7) << 2))); }

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    lo = v - hi;
}

// dynamic smem (1024-byte aligned): xs0 [CI][256] | xs1 [CI][16] | A_hi [4][KP][32] | A_lo | B_hi [KP][32] | B_lo
template <int CO, int KW, int S, int P>
__global__ void __launch_bounds__(CT_TILE) conv1d_fwd_tc_kernel(const __grid_constant__ CUtensorMap map0,
                                                                const __grid_constant__ CUtensorMap map1,
                                                                const float* __restrict__ w, const float* __restrict__ gate,
                                                                float* __restrict__ y, double* __restrict__ stats, int CI, int KP,
                                                                int Lout) {
    static_assert(CO == 16 || CO == 32, "conv1d_fwd_tc: C_out must be 16 or 32");
    extern __shared__ __align__(1024) uint8_t ct_smem[];
    __shared__ __align__(8) uint64_t load_bar, mma_bar;
    __shared__ uint32_t tmem_base_sh;
    __shared__ double red[CT_TILE / 32][2 * CO];
    __shared__ float s_gate[16];
    constexpr int KP_MAX = (16 * KW + 7) / 8 * 8;        // C_in <= 16

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, l0 = blockIdx.x * CT_TILE;
    const int K = CI * KW;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ct_smem) + 1023) & ~(uintptr_t)1023);
    float* xs0 = reinterpret_cast<float*>(base);
    float* xs1 = xs0 + CI * CT_BOX0;
    const uint32_t xs_bytes = (uint32_t)(((CI * (CT_BOX0 + CT_BOX1) * 4) + 1023) & ~1023);
    const uint32_t a_blk = (uint32_t)KP * 128, a_bytes = 4 * a_blk, b_bytes = (uint32_t)KP * 128;
    uint8_t* A_hi = base + xs_bytes;
    uint8_t* A_lo = A_hi + a_bytes;
    uint8_t* B_hi = A_lo + a_bytes;
    uint8_t* B_lo = B_hi + b_bytes;

    // weights of this thread's share of operand B, fetched before the set-up latencies (TMEM allocation, barrier init,
    // TMA) so that they overlap: element idx = tid + i * 128 -> (k-row r = idx / 32, channel o = idx % 32)
    constexpr int WPT = (KP_MAX * 32) / CT_TILE;
    float wreg[WPT];
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
        const int idx = tid + i * CT_TILE, r = idx >> 5, o = idx & 31;
        wreg[i] = (idx < KP * 32 && r < K && o < CO) ? __ldg(w + (size_t)o * K + r) : 0.f;
    }
    if (tid < CI) s_gate[tid] = gate ? __ldg(gate + b * CI + tid) : 1.f;

    if (tid == 0) {
        mbar_init(&load_bar, 1);
        mbar_init(&mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_sh;

    // TMA needs the innermost start coordinate 16-byte aligned: start XOFF columns to the left of S*l0 - P
    constexpr int XOFF = (4 - P % 4) % 4;
    static_assert((CT_TILE * S) % 4 == 0 && (CT_TILE - 1) * S + KW + XOFF <= CT_BOX0 + CT_BOX1, "tile span exceeds the two TMA boxes");
    if (tid == 0) {      // input tile: rows b*CI .. b*CI+CI-1 (negative / >= Lin columns read as zero)
        mbar_expect_tx(&load_bar, (uint32_t)CI * (CT_BOX0 + CT_BOX1) * 4);
        tma_load_2d(&map0, &load_bar, xs0, l0 * S - P - XOFF, b * CI);
        tma_load_2d(&map1, &load_bar, xs1, l0 * S - P - XOFF + CT_BOX0, b * CI);
    }
    // operand B while the tile is in flight: W^T[(c,k), o], zero rows for the K padding, zero columns o >= CO
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
        const int idx = tid + i * CT_TILE;
        if (idx < KP * 32) {
            float hi, lo;
            split_tf32(wreg[i], hi, lo);
            const uint32_t off = mn_off(idx >> 5, idx & 31);
            *reinterpret_cast<float*>(B_hi + off) = hi;
            *reinterpret_cast<float*>(B_lo + off) = lo;
        }
    }
    mbar_wait(&load_bar, 0);
    // operand A: thread = position p; A[p, (c,k)] = gate[b,c] * x[c][S*p + k]
    {
        const int p = tid, blk = p >> 5, j = p & 31;
        uint8_t* ah = A_hi + (size_t)blk * a_blk;
        uint8_t* al = A_lo + (size_t)blk * a_blk;
        int r = 0;
        for (int c = 0; c < CI; ++c) {
            const float g = s_gate[c];
#pragma unroll
            for (int k = 0; k < KW; ++k, ++r) {
                const int i = p * S + k + XOFF;
                const float v = g * (i < CT_BOX0 ? xs0[c * CT_BOX0 + i] : xs1[c * CT_BOX1 + (i - CT_BOX0)]);
                float hi, lo;
                split_tf32(v, hi, lo);
                const uint32_t off = mn_off(r, j);
                *reinterpret_cast<float*>(ah + off) = hi;
                *reinterpret_cast<float*>(al + off) = lo;
            }
        }
        for (; r < KP; ++r) {
            const uint32_t off = mn_off(r, j);
            *reinterpret_cast<float*>(ah + off) = 0.f;
            *reinterpret_cast<float*>(al + off) = 0.f;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = umma_idesc_tf32_mn(CO);
        const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo), b_hi = smem_u32(B_hi), b_lo = smem_u32(B_lo);
        for (int ks = 0; ks < KP / TC_UMMA_K; ++ks) {
            const uint32_t off = (uint32_t)ks * 1024;               // 8 k-rows of 128 bytes
            const uint64_t dAh = umma_desc_mnmajor_sw128(a_hi + off, a_blk), dAl = umma_desc_mnmajor_sw128(a_lo + off, a_blk);
            const uint64_t dBh = umma_desc_mnmajor_sw128(b_hi + off, b_bytes), dBl = umma_desc_mnmajor_sw128(b_lo + off, b_bytes);
            umma_tf32(tmem_d, dAl, dBh, idesc, ks != 0);
            umma_tf32(tmem_d, dAh, dBl, idesc, 1);
            umma_tf32(tmem_d, dAh, dBh, idesc, 1);
        }
        umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: TMEM lane = position
    float acc[CO];
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) acc[c0 + jj] = __uint_as_float(r[jj]);
    }
    const int l = l0 + tid;
    const bool valid = l < Lout;
    if (valid) {
        float* yb = y + (size_t)b * CO * Lout + l;
#pragma unroll
        for (int o = 0; o < CO; ++o) yb[(size_t)o * Lout] = acc[o];
    }
    if (stats) {
        // per-warp sums of every channel over the 32 positions: transposing butterfly (each exchange halves the number
        // of values a lane carries): 2 * (CO - 1 + extra) shuffles instead of 2 * CO * 5
        float sv[CO], qv[CO];
#pragma unroll
        for (int o = 0; o < CO; ++o) { sv[o] = valid ? acc[o] : 0.f; qv[o] = sv[o] * sv[o]; }
        int n = CO;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            if (n > 1) {
                const bool upper = (lane & m) != 0;
                n >>= 1;
#pragma unroll
                for (int i = 0; i < CO / 2; ++i) {
                    if (i < n) {
                        const float ks = upper ? sv[i + n] : sv[i], ss = upper ? sv[i] : sv[i + n];
                        const float kq = upper ? qv[i + n] : qv[i], sq = upper ? qv[i] : qv[i + n];
                        sv[i] = ks + __shfl_xor_sync(0xffffffffu, ss, m);
                        qv[i] = kq + __shfl_xor_sync(0xffffffffu, sq, m);
                    }
                }
            } else {
                sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], m);
                qv[0] += __shfl_xor_sync(0xffffffffu, qv[0], m);
            }
        }
        // lane holds channel ch(lane): the bits of `lane` consumed while n > 1, most significant exchange first
        int ch = 0, nn = CO;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            if (nn > 1) { nn >>= 1; if (lane & m) ch += nn; }
        }
        constexpr int DUP = 32 / CO;                 // lanes that end up with the same channel (1 or 2)
        if ((lane & (DUP - 1)) == 0) { red[warp][ch] = (double)sv[0]; red[warp][CO + ch] = (double)qv[0]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (stats && tid < 2 * CO) {
        double t = 0.0;
#pragma unroll
        for (int wq = 0; wq < CT_TILE / 32; ++wq) t += red[wq][tid];
        atomicAdd(stats + tid, t);
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(32u) : "memory");
    }
}

template <int CO, int KW, int S, int P>
static int conv_fwd_tc_launch(const float* x, const float* w, const float* gate, int B, int CI, int Lin, float* y, double* stats,
                              cudaStream_t st) {
    const int Lout = conv_out_len(Lin, KW, S, P);
    const int K = CI * KW, KP = (K + 7) / 8 * 8;
    CUtensorMap map0, map1;
    int rc = make_map(&map0, x, (int64_t)B * CI, Lin, Lin, CI, CU_TENSOR_MAP_SWIZZLE_NONE, CT_BOX0);
    if (rc) return rc;
    rc = make_map(&map1, x, (int64_t)B * CI, Lin, Lin, CI, CU_TENSOR_MAP_SWIZZLE_NONE, CT_BOX1);
    if (rc) return rc;
    const size_t xs_bytes = (size_t)((CI * (CT_BOX0 + CT_BOX1) * 4 + 1023) & ~1023);
    const size_t smem = xs_bytes + (size_t)2 * 4 * KP * 128 + (size_t)2 * KP * 128 + 1024;
    auto kern = conv1d_fwd_tc_kernel<CO, KW, S, P>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) { MMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); }
    MMS_REQUIRE(smem <= 200 * 1024, "conv1d_fwd_tc: shared memory %zu too large", smem);
    dim3 grid(cdiv(Lout, CT_TILE), B);
    MMS_PROF_BEGIN(st);
    kern<<<grid, CT_TILE, smem, st>>>(map0, map1, w, gate, y, stats, CI, KP, Lout);
    MMS_LAUNCH_CHECK("conv1d_fwd_tc_kernel");
    return MMS_OK;
}

bool conv_fwd_tc_supported(int which, const float* x, int c_in, int c_out, int l_in) {
    if (!encode_tiled_fn() || (reinterpret_cast<uintptr_t>(x) & 15) || l_in % 4 || l_in < 64) return false;
    if (which == 1) return c_out == 16 && c_in >= 1 && c_in <= 16;
    return c_in == 16 && (c_out == 16 || c_out == 32);
}

int launch_conv_fwd_tc(int which, const float* x, const float* w, const float* gate, int B, int c_in, int c_out, int l_in, float* y,
                       double* stats, cudaStream_t st) {
    MMS_REQUIRE(conv_fwd_tc_supported(which, x, c_in, c_out, l_in), "conv1d_fwd_tc: unsupported shape / alignment");
    if (which == 1) return conv_fwd_tc_launch<16, CONV1_K, CONV1_S, CONV1_P>(x, w, gate, B, c_in, l_in, y, stats, st);
    if (c_out == 16) return conv_fwd_tc_launch<16, CONV2_K, CONV2_S, CONV2_P>(x, w, gate, B, c_in, l_in, y, stats, st);
    return conv_fwd_tc_launch<32, CONV2_K, CONV2_S, CONV2_P>(x, w, gate, B, c_in, l_in, y, stats, st);
}

}  // namespace mms
