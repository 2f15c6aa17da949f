// Tensor-core GEMMs for the GRU input projections and their gradients (the x @ W_ih^T + b_ih part of
// nn.GRU, reference models.py:56-63,78) on Blackwell's 5th-generation tensor cores:
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) straight from the fp32 tensors;
//   * tcgen05.mma.kind::tf32 issued by one elected thread, accumulators in tensor memory (TMEM),
//     read back with tcgen05.ld for the epilogue (bias add, optional accumulate, store);
//   * fp32-level accuracy through the 3xTF32 split: every operand tile x is rewritten in shared
//     memory as hi = x with the low 13 mantissa bits cleared (exactly representable in tf32) and
//     lo = x - hi, and each k-step issues  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.  The dropped term
//     lo*lo is ~2^-22 relative -- the result is indistinguishable from an fp32 FMA chain at the
//     1e-4 logit tolerance, which a plain tf32 GEMM (2^-11) is not.
// Warp roles in a 192-thread CTA: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2-5 = operand splitters, then epilogue (one TMEM lane quarter each).
//
//   NT :  C[m, n] (+)= sum_k A[m, k] * B[n, k] + bias[n]        (A, B both K-major = row-major)
#include "tc_common.cuh"
#include <string.h>

namespace mms {

struct TcGemmParams {
    float* C;
    const float* bias;
    int64_t ldc;
    int M, N, K, BN, accumulate;
    // optional dropout on the OUTPUT (the gradient through nn.GRU's inter-layer dropout, models.py:62): C[m][n] is multiplied by
    // the multiplier of element drop_base + m * ldc + n of the counter-based stream (mms_common.cuh) -- what a separate
    // dropout_apply pass over C would do, without the extra launch and the read + write of C.  drop_p == 0: off.
    // drop_on_a: the multipliers go onto the A operand instead (element drop_base + m * drop_lda + k), while it is split for the
    // tensor core: the product then reads the UN-dropped tensor and a separate dropout pass leaves the critical path.
    float drop_p;
    uint64_t drop_seed, drop_offset;
    const int64_t* drop_offset_dev;
    int64_t drop_base, drop_lda;
    int drop_on_a;
};

// dynamic smem: [stage][A_hi | A_lo | B_hi | B_lo] (1024-byte aligned tiles); NS = pipeline stages (2, or 4 for long reductions)
// 384 threads = 12 warps: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-11 = operand splitters (clock stamps inside a CTA
// showed the splitters setting the pace of a k-block with FOUR warps -- one per scheduler, every latency exposed; ten warps
// overlap each other's).  ALL warps run the epilogue: a warp reads the TMEM lane quarter (warp % 4), so every quarter has three
// warps that take every third block of 32 columns each (tcgen05.ld, shared-memory transpose, bias / dropout, 128-byte stores:
// 0.7-5.5 us of a CTA's 12 us with four warps).  68 registers: two such CTAs still share an SM for the single-k-block product.
constexpr int NT_THREADS = 384, NT_WARPS = NT_THREADS / 32, NT_SPLIT_THREADS = NT_THREADS - 64, NT_GROUPS = NT_WARPS / 4;
template <int NS>
__global__ void __launch_bounds__(NT_THREADS, 1) tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB, const TcGemmParams p) {
    MMS_PDL_TRIGGER();
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t full_bar[NS], split_bar[NS], empty_bar[NS], acc_bar;
    __shared__ uint32_t tmem_base_sh;
    __shared__ __align__(16) float s_bias[256];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int BN = p.BN;
    const uint32_t a_bytes = TC_BM * TC_BK * 4, b_bytes = (uint32_t)BN * TC_BK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint8_t* base = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);      // pointer arithmetic keeps the address space: LDS / STS, not generic LD / ST
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
    const int nkb = (p.K + TC_BK - 1) / TC_BK;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < BN) tmem_cols <<= 1;
    for (int i = threadIdx.x; i < 256; i += NT_THREADS) s_bias[i] = (p.bias && i < BN && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&split_bar[s], NT_WARPS - 2);      // one arrival per splitter warp
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    MMS_PDL_WAIT();       // barrier / TMEM set-up and the bias (a parameter) are done; A and C belong to the kernels before this one
    const uint32_t tmem_d = tmem_base_sh;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NS, round = kb / NS;
                if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
                uint8_t* st = base + (size_t)s * stage_bytes;
                mbar_expect_tx(&full_bar[s], a_bytes + b_bytes);
                tma_load_2d(&mapA, &full_bar[s], st, kb * TC_BK, m0);
                tma_load_2d(&mapB, &full_bar[s], st + 2 * a_bytes, kb * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NS, round = kb / NS;
                mbar_wait(&split_bar[s], round & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(base + (size_t)s * stage_bytes);
                const uint32_t a_hi = st, a_lo = st + a_bytes, b_hi = st + 2 * a_bytes, b_lo = st + 2 * a_bytes + b_bytes;
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    const uint32_t off = k * TC_UMMA_K * 4;
                    const uint64_t dAh = umma_desc_kmajor_sw128(a_hi + off), dAl = umma_desc_kmajor_sw128(a_lo + off);
                    const uint64_t dBh = umma_desc_kmajor_sw128(b_hi + off), dBl = umma_desc_kmajor_sw128(b_lo + off);
                    umma_tf32(tmem_d, dAl, dBh, idesc, (kb | k) != 0);
                    umma_tf32(tmem_d, dAh, dBl, idesc, 1);
                    umma_tf32(tmem_d, dAh, dBh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);            // frees the stage once these MMAs have read it
            }
            umma_commit(&acc_bar);                     // accumulator complete
        }
    } else {
        // ===== operand splitters (warps 2 .. NT_WARPS - 1) =====
        const int t = threadIdx.x - 64;                // 0 .. NT_SPLIT_THREADS - 1
        DropRng rng_a;
        const bool drop_a = p.drop_p > 0.f && p.drop_on_a;
        if (drop_a) rng_a.init(p.drop_seed, resolve_offset(p.drop_offset, p.drop_offset_dev), p.drop_p);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % NS, round = kb / NS;
            mbar_wait(&full_bar[s], round & 1);
            uint8_t* st = base + (size_t)s * stage_bytes;
            // A: hi in place, lo to the second tile (same offsets, so the swizzle is irrelevant)
            {
                float4* hi = reinterpret_cast<float4*>(st);
                float4* lo = reinterpret_cast<float4*>(st + a_bytes);
                for (int i = t; i < (int)(a_bytes / 16); i += NT_SPLIT_THREADS) {
                    float4 v = hi[i];
                    if (drop_a) {
                        // 128-byte swizzle: row r of the tile holds its 16-byte chunk c at position c ^ (r & 7)
                        const int r = i >> 3, c = (i & 7) ^ (r & 7);
                        const uint64_t e = (uint64_t)(p.drop_base + (int64_t)(m0 + r) * p.drop_lda + kb * TC_BK + 4 * c);
                        v.x *= rng_a.mult(e); v.y *= rng_a.mult(e + 1); v.z *= rng_a.mult(e + 2); v.w *= rng_a.mult(e + 3);
                    }
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    hi[i] = h;
                    lo[i] = l;
                }
            }
            {
                float4* hi = reinterpret_cast<float4*>(st + 2 * a_bytes);
                float4* lo = reinterpret_cast<float4*>(st + 2 * a_bytes + b_bytes);
                for (int i = t; i < (int)(b_bytes / 16); i += NT_SPLIT_THREADS) {
                    const float4 v = hi[i];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                    hi[i] = h;
                    lo[i] = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&split_bar[s]);
        }
    }
    __syncwarp();             // lanes 1..31 of the producer / issuer warps wait here (no spinning beside lane 0's loop)
    {
        // ===== epilogue (all warps): TMEM lane quarter (warp % 4) -> registers -> global; warp group (warp / 4) takes the column
        // blocks group, group + NT_GROUPS, ... =====
        mbar_wait(&acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;
        const int group = warp >> 2;
        const bool aligned = !p.accumulate && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (n0 & 3) == 0 &&
                             (p.N & 3) == 0;
        if (aligned) {
            // Transposed store: a thread owns one accumulator ROW in TMEM, but a warp-wide store of per-row pieces touches 32
            // different 128-byte lines (ncu: the store queue was the second-largest stall).  Each warp therefore passes its
            // 32 x 32 block through a private shared-memory tile (the operand stages are free once acc_bar has fired) and
            // writes 4 rows x 128 contiguous bytes per instruction.  The bias comes from shared memory (loaded at kernel
            // start; fetching it here from global memory was the largest stall).
            float* tb = reinterpret_cast<float*>(base) + warp * (32 * 36);
            const int rbase = m0 + quarter * 32;
            DropRng rng;
            const bool drop = p.drop_p > 0.f && !p.drop_on_a;
            if (drop) rng.init(p.drop_seed, resolve_offset(p.drop_offset, p.drop_offset_dev), p.drop_p);
            for (int c0 = group * 32; c0 < BN; c0 += NT_GROUPS * 32) {
                const int wcols = min(32, BN - c0);           // 32 or 16 (BN is a multiple of 16)
                uint32_t r[32];
                const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                if (wcols == 32) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr + 16));
                } else {
#pragma unroll
                    for (int j = 16; j < 32; ++j) r[j] = 0u;
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float4* trow = reinterpret_cast<float4*>(tb + lane * 36);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    trow[j4] = make_float4(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                                           __uint_as_float(r[4 * j4 + 3]));
                __syncwarp();
                const int col4 = lane & 7, rsub = lane >> 3;
                const int n = n0 + c0 + 4 * col4;
                if (4 * col4 < wcols && n < p.N) {
                    const float4 bv = *reinterpret_cast<const float4*>(&s_bias[c0 + 4 * col4]);
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + rsub;
                        if (rbase + rr < p.M) {
                            float4 v = *reinterpret_cast<const float4*>(tb + rr * 36 + 4 * col4);
                            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                            if (drop) {
                                const uint64_t e = (uint64_t)(p.drop_base + (int64_t)(rbase + rr) * p.ldc + n);
                                v.x *= rng.mult(e); v.y *= rng.mult(e + 1); v.z *= rng.mult(e + 2); v.w *= rng.mult(e + 3);
                            }
                            *reinterpret_cast<float4*>(p.C + (int64_t)(rbase + rr) * p.ldc + n) = v;
                        }
                    }
                }
                __syncwarp();
            }
        } else {
            const int row = m0 + quarter * 32 + lane;
            float* crow = p.C + (int64_t)row * p.ldc + n0;
            for (int c0 = group * 16; c0 < BN; c0 += NT_GROUPS * 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < p.M) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = n0 + c0 + j;
                        if (n < p.N) {
                            float v = __uint_as_float(r[j]) + s_bias[c0 + j];
                            if (p.accumulate) v += crow[c0 + j];
                            crow[c0 + j] = v;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

static int pick_bn(int N) {
    // UMMA N for M = 128: multiple of 16 in [16, 256]; split wide outputs into equal tiles
    int tiles = (N + 255) / 256;
    int bn = (N + tiles - 1) / tiles;
    bn = (bn + 15) / 16 * 16;
    return bn;
}

bool tc_gemm_supported(const float* A, int64_t lda, const float* W, int64_t ldw, int M, int N, int K) {
    return encode_tiled_fn() && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
           lda % 4 == 0 && ldw % 4 == 0 && M >= 1 && N >= 1 && K >= 1;
}

// true when launch_tc_gemm_nt_drop can apply the output dropout itself (the coalesced epilogue)
bool tc_gemm_nt_drop_supported(const float* C, int64_t ldc, int N) {
    const int BN = pick_bn(N);
    return (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (N & 3) == 0 && (BN & 3) == 0;
}

int launch_tc_gemm_nt_drop(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc, int M,
                           int N, int K, int accumulate, cudaStream_t st, float drop_p, uint64_t seed, uint64_t offset,
                           const int64_t* offset_dev, int64_t drop_base, int drop_on_a);

int launch_tc_gemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc, int M,
                      int N, int K, int accumulate, cudaStream_t st) {
    return launch_tc_gemm_nt_drop(A, lda, W, ldw, bias, C, ldc, M, N, K, accumulate, st, 0.f, 0, 0, nullptr, 0, 0);
}

int launch_tc_gemm_nt_drop(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc, int M,
                           int N, int K, int accumulate, cudaStream_t st, float drop_p, uint64_t seed, uint64_t offset,
                           const int64_t* offset_dev, int64_t drop_base, int drop_on_a) {
    if (M <= 0 || N <= 0) return MMS_OK;
    MMS_REQUIRE(drop_p == 0.f || drop_on_a || (!accumulate && tc_gemm_nt_drop_supported(C, ldc, N)),
                "tc_gemm_nt: output dropout needs the aligned epilogue");
    const int BN = pick_bn(N);
    CUtensorMap mapA, mapB;
    int rc = make_map(&mapA, A, M, K, lda, TC_BM);
    if (rc) return rc;
    rc = make_map(&mapB, W, N, K, ldw, BN);
    if (rc) return rc;
    const size_t stage = (size_t)2 * TC_BM * TC_BK * 4 + (size_t)2 * BN * TC_BK * 4;
    // MMS_NT_TRIM_STAGES=1 (experiment): a product with a single k-block (K <= 32: the layer-0 input projection) only
    // ever touches stage 0, so it can ask for one stage of shared memory and let two CTAs share an SM (81 KB, 2 x 256
    // TMEM columns), overlapping one CTA's epilogue with the other's TMA / split / MMA
    const int nkb = (K + TC_BK - 1) / TC_BK;
    // pipeline depth: 4 stages when the reduction has at least 4 k-blocks and they fit (MMS_NT_STAGES=2 keeps two): with two, a
    // CTA's TMA latency is exposed once per k-block pair -- its 16 k-blocks made the K = 512 input-gradient product the
    // slowest of the four; 1 stage for a single k-block (two CTAs per SM)
    int stages = TC_STAGES;
    if (nkb >= 4 && option_get("NT_STAGES", 4) >= 4 && stage * 4 + 1024 <= 200 * 1024) stages = 4;
    const bool four = stages == 4;
    if (option_get("NT_TRIM_STAGES", 1) == 1 && nkb < TC_STAGES) stages = nkb;
    // the epilogue's per-warp transpose tiles (NT_WARPS x 4.5 KB) reuse the operand stages
    const size_t smem = (stage * stages > (size_t)NT_WARPS * 32 * 36 * 4 ? stage * stages : (size_t)NT_WARPS * 32 * 36 * 4) + 1024;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        MMS_CUDA(cudaFuncSetAttribute(tc_gemm_nt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MMS_CUDA(cudaFuncSetAttribute(tc_gemm_nt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    MMS_REQUIRE(smem <= 200 * 1024, "tc_gemm: shared memory %zu too large", smem);
    TcGemmParams p;
    p.C = C; p.bias = bias; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.BN = BN; p.accumulate = accumulate;
    p.drop_p = drop_p; p.drop_seed = seed; p.drop_offset = offset; p.drop_offset_dev = offset_dev; p.drop_base = drop_base;
    p.drop_lda = lda; p.drop_on_a = drop_on_a;
    dim3 grid(cdiv(M, TC_BM), cdiv(N, BN));
    MMS_PROF_BEGIN(st);
    if (four) MMS_LAUNCH(tc_gemm_nt_kernel<4>, grid, dim3(NT_THREADS), smem, st, mapA, mapB, p);
    else MMS_LAUNCH(tc_gemm_nt_kernel<2>, grid, dim3(NT_THREADS), smem, st, mapA, mapB, p);
    MMS_LAUNCH_CHECK("tc_gemm_nt_kernel");
    return MMS_OK;
}

// ---- TN form: weight gradients of the GRU projections ----------------------------------------------
//   C[i, j] += sum_m A[m, acol(i)] * Bm[m + shift, j]        bias_grad[i] += sum_m A[m, acol(i)]
// with acol(i) = i < a_split ? i : i + a_skip (skips the n-gate or dq column block of the backward's D rows) and
// rows whose shifted partner leaves its sequence (t + shift outside [0, seq)) reading as zero (h_{t-1} / h_{t+1}).
// The reduction index m is the ROW index of both row-major operands, i.e. both are MN-major for the tensor core:
// a TMA box of [KB rows x 32 floats] with the 128-byte / 32-byte-atom swizzle is exactly the canonical MN-major layout
// of 32-bit operands (see umma_desc_mnmajor_sw128), so the fp32 tensors are consumed in place, no transposition pass.
//   * M side = up to 256 columns of A as two UMMA M = 128 tiles sharing the B tile;
//   * the bias gradient comes out of the same MMAs: B gets one extra 32-column block whose first column is 1.0;
//   * split-K over the grid (the outputs are tiny, the reduction is B*L rows long); the partial tiles are added
//     to C with vector red.global.add.f32 (C is the zero-initialised flat gradient buffer);
//   * 3xTF32 operand split and warp roles as in the NT kernel; the splitter warps also zero the B rows that the
//     shift moves across a sequence boundary.
// 16 rows per k-block and 3 stages (round 2, after clock stamps inside a CTA): with 32-row k-blocks a stage is 88 KB, only two fit,
// and a k-block then cost one exposed TMA latency (~2000 cycles) + its split + its MMAs = 3300 cycles; half-size k-blocks four deep
// keep the loads ahead of the splitters (four deep measured 1.7 % slower in the step than three: the fourth stage's
// 44 KB of shared memory keep the chain's CTAs off the SM).
constexpr int TN_KB = 16;           // reduction rows per k-block
constexpr int TN_STAGES = 3;        // pipeline depth of the default instantiation (NS below): 3 x 44 KB leave room for a co-resident CTA of the chain
constexpr int TN_BLK = TN_KB * 128; // bytes of one [KB x 32 floats] column block
// 384 threads as in the NT kernel: warp 0 = TMA, warp 1 = MMA, warps 2-11 split the operands, all twelve run the epilogue
constexpr int TN_THREADS = 384, TN_WARPS = TN_THREADS / 32, TN_SPLIT_THREADS = TN_THREADS - 64, TN_GROUPS = TN_WARPS / 4;

struct TcGemmTnParams {
    float* C;
    float* bias_grad;
    int64_t ldc;
    int M, N1, N2, a_split, a_skip, shift, seq, chunk, nblkA, nblkB;   // nblkB counts the data blocks (without the ones block)
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// dynamic smem per stage: [A_hi (nblkA blocks) | A_lo | B_hi (nblkB + 1 blocks) | B_lo], 1024-byte aligned.
// NS = pipeline stages: 2 by default; 1 (MMS_TN_STAGES=1, experiment) halves the shared-memory footprint (<= 88 KB), so
// that a weight-gradient CTA on a side stream no longer blocks its SM for the conv / pool kernels of the main chain.
template <int NS>
__global__ void __launch_bounds__(TN_THREADS, 1) tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB, const TcGemmTnParams p) {
#define TN_MAPA (&mapA)
#define TN_MAPB (&mapB)
#define TN_CHUNK_IDX blockIdx.x
#include "tc_gemm_tn_body.inc"
#undef TN_MAPA
#undef TN_MAPB
#undef TN_CHUNK_IDX
}

// Up to TN_MAX_BATCH independent problems in ONE launch (MMS_TN_BATCH=1, experiment): the weight-gradient products of a
// GRU layer are four small split-K GEMMs over the same B*L rows; launched one after the other they cost four launches,
// four prologues / epilogues and 4 x 96 partial-tile reductions.  Batched, ~148 CTAs cover all of them at once, each with
// a 2-4 times longer reduction chunk (better pipelining, proportionally fewer red.global.add tiles).
constexpr int TN_MAX_BATCH = 4;
struct TnBatchMaps { CUtensorMap a[TN_MAX_BATCH], b[TN_MAX_BATCH]; };
struct TnBatchParams { TcGemmTnParams p[TN_MAX_BATCH]; int nchunks[TN_MAX_BATCH]; };

template <int NS>
__global__ void __launch_bounds__(TN_THREADS, 1) tc_gemm_tn_batch_kernel(const __grid_constant__ TnBatchMaps maps, const TnBatchParams bp) {
    const int j = blockIdx.y;
    if ((int)blockIdx.x >= bp.nchunks[j]) return;      // whole CTA leaves before any barrier / TMEM allocation
    const TcGemmTnParams p = bp.p[j];
#define TN_MAPA (&maps.a[j])
#define TN_MAPB (&maps.b[j])
#define TN_CHUNK_IDX blockIdx.x
#include "tc_gemm_tn_body.inc"
#undef TN_MAPA
#undef TN_MAPB
#undef TN_CHUNK_IDX
}

bool tc_gemm_tn_supported(const float* A, int64_t lda, int a_split, int a_skip, const float* Bm, int64_t ldb, float* C, int64_t ldc,
                          int M, int N1, int N2) {
    if (!encode_tiled_fn() || M < 4 * TN_KB) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || lda % 4 || N1 < 1 || N1 > 256 || N1 % 32 || a_split % 32 || a_skip % 32) return false;
    if (N2 > 0 && ((reinterpret_cast<uintptr_t>(Bm) & 15) || ldb % 4 || N2 % 32 || N2 > 224 || !C ||
                   (reinterpret_cast<uintptr_t>(C) & 15) || ldc % 4)) return false;
    return true;
}

int launch_tc_gemm_tn(const float* A, int64_t lda, int a_split, int a_skip, const float* Bm, int64_t ldb, int shift, int seq,
                      float* C, int64_t ldc, float* bias_grad, int M, int N1, int N2, cudaStream_t st) {
    if (M <= 0 || N1 <= 0) return MMS_OK;
    if (N2 <= 0 || !Bm) { N2 = 0; Bm = nullptr; C = nullptr; }
    MMS_REQUIRE(tc_gemm_tn_supported(A, lda, a_split, a_skip, Bm, ldb, C, ldc, M, N1, N2), "tc_gemm_tn: unsupported shape / alignment");
    MMS_REQUIRE(seq >= 1 && (shift == 0 || M % seq == 0), "tc_gemm_tn: M must be a multiple of seq when rows are shifted");
    CUtensorMap mapA, mapB;
    int rc = make_map(&mapA, A, M, (a_split < N1 ? a_skip : 0) + N1, lda, TN_KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    if (N2 > 0) rc = make_map(&mapB, Bm, M, N2, ldb, TN_KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    else mapB = mapA;
    if (rc) return rc;
    TcGemmTnParams p;
    p.C = C; p.bias_grad = bias_grad; p.ldc = ldc; p.M = M; p.N1 = N1; p.N2 = N2; p.a_split = a_split; p.a_skip = a_skip;
    p.shift = shift; p.seq = seq; p.nblkA = N1 / 32; p.nblkB = N2 / 32;
    // split-K: ~one CTA per SM pair for long reductions, never fewer than 2 k-blocks per CTA
    // MMS_TN_SPLIT (experiment): target CTA count of the split (fewer CTAs = fewer red.global.add partial tiles)
    int nsplit = option_get("TN_SPLIT", 96);
    if (nsplit < 1) nsplit = 1;
    int chunk = (int)align_up(cdiv(M, nsplit), TN_KB);
    if (chunk < 2 * TN_KB) chunk = 2 * TN_KB;
    p.chunk = chunk;
    const size_t stage = (size_t)2 * p.nblkA * TN_BLK + (size_t)2 * (p.nblkB + 1) * TN_BLK;
    const int ns = option_get("TN_STAGES", TN_STAGES) == 1 ? 1 : TN_STAGES;
    const size_t smem = stage * ns + 1024;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        MMS_CUDA(cudaFuncSetAttribute(tc_gemm_tn_kernel<TN_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MMS_CUDA(cudaFuncSetAttribute(tc_gemm_tn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
       
    }
    MMS_REQUIRE(smem <= 220 * 1024, "tc_gemm_tn: shared memory %zu too large", smem);
    MMS_PROF_BEGIN(st);
    if (ns == 1) tc_gemm_tn_kernel<1><<<cdiv(M, chunk), TN_THREADS, smem, st>>>(mapA, mapB, p);
    else tc_gemm_tn_kernel<TN_STAGES><<<cdiv(M, chunk), TN_THREADS, smem, st>>>(mapA, mapB, p);
    MMS_LAUNCH_CHECK("tc_gemm_tn_kernel");
    return MMS_OK;
}

// All of `calls` (2 .. TN_MAX_BATCH problems, each acceptable to tc_gemm_tn_supported) in one launch of the batched kernel.
// The CTA budget (MMS_TN_BATCH_CTAS, default 148 = one per SM) is shared equally: every problem of a GRU layer reduces over
// the same B*L rows.
int launch_tc_gemm_tn_batch(const TnCall* calls, int n, cudaStream_t st) {
    MMS_REQUIRE(calls && n >= 1 && n <= TN_MAX_BATCH, "tc_gemm_tn_batch: 1..%d problems", TN_MAX_BATCH);
    TnBatchMaps maps;
    TnBatchParams bp;
    memset(&bp, 0, sizeof(bp));
    // 80 CTAs, not one per SM: the products run beside the critical chain (the dseq product, pool / conv backward), and 68 free SMs
    // serve it better than a 1.9x shorter weight-gradient launch (A/B r2c60: 148 -> 0.451, 120 -> 0.448, 100 -> 0.443, 80 -> 0.440, 64 -> 0.442 ms)
    int budget = option_get("TN_BATCH_CTAS", 80);
    if (budget < n) budget = n;
    int ns = option_get("TN_STAGES", TN_STAGES);           // 1 .. 4 pipeline stages; reduced until the largest problem's stages fit
    ns = ns < 1 ? 1 : (ns > 4 ? 4 : ns);
    size_t stage_max = 0;
    int max_chunks = 0;
    for (int j = 0; j < n; ++j) {
        TnCall c = calls[j];
        MMS_REQUIRE(c.M > 0 && c.N1 > 0, "tc_gemm_tn_batch: problem %d is empty", j);
        if (c.N2 <= 0 || !c.Bm) { c.N2 = 0; c.Bm = nullptr; c.C = nullptr; }
        MMS_REQUIRE(tc_gemm_tn_supported(c.A, c.lda, c.a_split, c.a_skip, c.Bm, c.ldb, c.C, c.ldc, c.M, c.N1, c.N2),
                    "tc_gemm_tn_batch: problem %d has an unsupported shape / alignment", j);
        MMS_REQUIRE(c.seq >= 1 && (c.shift == 0 || c.M % c.seq == 0), "tc_gemm_tn_batch: M must be a multiple of seq when rows are shifted");
        int rc = make_map(&maps.a[j], c.A, c.M, (c.a_split < c.N1 ? c.a_skip : 0) + c.N1, c.lda, TN_KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
        if (c.N2 > 0) rc = make_map(&maps.b[j], c.Bm, c.M, c.N2, c.ldb, TN_KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        else maps.b[j] = maps.a[j];
        if (rc) return rc;
        TcGemmTnParams& p = bp.p[j];
        p.C = c.C; p.bias_grad = c.bias_grad; p.ldc = c.ldc; p.M = c.M; p.N1 = c.N1; p.N2 = c.N2; p.a_split = c.a_split; p.a_skip = c.a_skip;
        p.shift = c.shift; p.seq = c.seq; p.nblkA = c.N1 / 32; p.nblkB = c.N2 / 32;
        int chunk = (int)align_up(cdiv(c.M, budget / n), TN_KB);
        if (chunk < 2 * TN_KB) chunk = 2 * TN_KB;
        p.chunk = chunk;
        bp.nchunks[j] = cdiv(c.M, chunk);
        if (bp.nchunks[j] > max_chunks) max_chunks = bp.nchunks[j];
        const size_t stage = (size_t)2 * p.nblkA * TN_BLK + (size_t)2 * (p.nblkB + 1) * TN_BLK;
        if (stage > stage_max) stage_max = stage;
    }
    while (ns > 1 && stage_max * ns + 1024 > 220 * 1024) --ns;
    const size_t smem = stage_max * ns + 1024;
    for (int j = n; j < TN_MAX_BATCH; ++j) { maps.a[j] = maps.a[0]; maps.b[j] = maps.b[0]; }     // unused slots: valid bytes, never read
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        MMS_CUDA(cudaFuncSetAttribute(tc_gemm_tn_batch_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MMS_CUDA(cudaFuncSetAttribute(tc_gemm_tn_batch_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MMS_CUDA(cudaFuncSetAttribute(tc_gemm_tn_batch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        MMS_CUDA(cudaFuncSetAttribute(tc_gemm_tn_batch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    }
    MMS_REQUIRE(smem <= 220 * 1024, "tc_gemm_tn_batch: shared memory %zu too large", smem);
    dim3 grid(max_chunks, n);
    MMS_PROF_BEGIN(st);
    if (ns == 1) tc_gemm_tn_batch_kernel<1><<<grid, TN_THREADS, smem, st>>>(maps, bp);
    else if (ns == 2) tc_gemm_tn_batch_kernel<2><<<grid, TN_THREADS, smem, st>>>(maps, bp);
    else if (ns == 3) tc_gemm_tn_batch_kernel<3><<<grid, TN_THREADS, smem, st>>>(maps, bp);
    else tc_gemm_tn_batch_kernel<4><<<grid, TN_THREADS, smem, st>>>(maps, bp);
    MMS_LAUNCH_CHECK("tc_gemm_tn_batch_kernel");
    return MMS_OK;
}

// out[c * ldo + col_off + r] = W[r * cols + c]   (W is [rows, cols] row-major)
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ out,
                                                            int64_t ldo, int col_off) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8)
        if (r0 + i < rows && c0 + tx < cols) tile[i][tx] = __ldg(W + (int64_t)(r0 + i) * cols + c0 + tx);
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) out[(int64_t)(c0 + i) * ldo + col_off + r0 + tx] = tile[tx][i];
}

int launch_transpose_pad(const float* W, int rows, int cols, float* out, int64_t ldo, int col_off, cudaStream_t st) {
    dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
    MMS_PROF_BEGIN(st);
    transpose_pad_kernel<<<grid, 256, 0, st>>>(W, rows, cols, out, ldo, col_off);
    MMS_LAUNCH_CHECK("transpose_pad_kernel");
    return MMS_OK;
}

// Up to 4 transposes in one launch (the W_ih^T copies of every GRU layer for the backward's dx products are three 4 us launches
// of a few dozen CTAs otherwise); `pad` further columns behind each transposed block are zero-filled (the dq columns of the
// bottom layers' D rows meet zero weights), which replaces the memset of the destination.
struct TransposeJobs {
    const float* W[4];
    float* out[4];
    int64_t ldo[4];
    int rows[4], cols[4], col_off[4], pad[4];
};
__global__ void __launch_bounds__(256) transpose_pad_multi_kernel(const TransposeJobs jobs) {
    __shared__ float tile[32][33];
    const int j = blockIdx.z;
    const int rows = jobs.rows[j], cols = jobs.cols[j], pad = jobs.pad[j];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    if (r0 >= rows + pad || c0 >= cols) return;
    const float* __restrict__ W = jobs.W[j];
    float* __restrict__ out = jobs.out[j] + jobs.col_off[j];
    const int64_t ldo = jobs.ldo[j];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8)
        tile[i][tx] = (r0 + i < rows && c0 + tx < cols) ? __ldg(W + (int64_t)(r0 + i) * cols + c0 + tx) : 0.f;     // rows beyond the matrix: the pad
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows + pad) out[(int64_t)(c0 + i) * ldo + r0 + tx] = tile[tx][i];
}

int launch_transpose_pad_multi(const TransposeJobs& jobs, int n, cudaStream_t st) {
    MMS_REQUIRE(n >= 1 && n <= 4, "transpose_pad_multi: 1..4 jobs");
    int gx = 1, gy = 1;
    for (int j = 0; j < n; ++j) {
        gx = max(gx, cdiv(jobs.cols[j], 32));
        gy = max(gy, cdiv(jobs.rows[j] + jobs.pad[j], 32));
    }
    MMS_PROF_BEGIN(st);
    transpose_pad_multi_kernel<<<dim3(gx, gy, n), 256, 0, st>>>(jobs);
    MMS_LAUNCH_CHECK("transpose_pad_multi_kernel");
    return MMS_OK;
}

}  // namespace mms

using namespace mms;

extern "C" int mms_tc_gemm_tn_batch(const mms_tn_call* calls_host, int32_t n, mms_stream_t stream) {
    MMS_REQUIRE(calls_host && n >= 1 && n <= TN_MAX_BATCH, "tc_gemm_tn_batch: 1..%d problems", TN_MAX_BATCH);
    TnCall c[TN_MAX_BATCH];
    for (int j = 0; j < n; ++j) {
        const mms_tn_call& h = calls_host[j];
        c[j] = {h.A, h.lda, h.a_split, h.a_skip, h.Bm, h.ldb, h.shift, h.seq, h.C, h.ldc, h.bias_grad, h.M, h.N1, h.N2};
    }
    return launch_tc_gemm_tn_batch(c, n, (cudaStream_t)stream);
}

extern "C" int mms_tc_gemm_tn(const float* A, int64_t lda, int32_t a_split, int32_t a_skip, const float* Bm, int64_t ldb, int32_t shift,
                              int32_t seq, float* C, int64_t ldc, float* bias_grad, int32_t M, int32_t N1, int32_t N2, mms_stream_t stream) {
    MMS_REQUIRE(A && (N2 == 0 || (Bm && C)), "tc_gemm_tn: null pointer");
    return launch_tc_gemm_tn(A, lda, a_split, a_skip, Bm, ldb, shift, seq, C, ldc, bias_grad, M, N1, N2, (cudaStream_t)stream);
}

extern "C" int mms_tc_gemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
                              int32_t M, int32_t N, int32_t K, int32_t accumulate, mms_stream_t stream) {
    MMS_REQUIRE(A && W && C && K > 0, "tc_gemm_nt: bad arguments");
    return launch_tc_gemm_nt(A, lda, W, ldw, bias, C, ldc, M, N, K, accumulate, (cudaStream_t)stream);
}
