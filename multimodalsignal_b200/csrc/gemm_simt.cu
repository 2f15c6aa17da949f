// fp32 SIMT GEMMs for the GRU input projections and their gradients (the x @ W_ih^T + b_ih part
// of nn.GRU, reference models.py:56-63,78).  64x64 output tile, BK = 16, 256 threads, 4x4 register
// micro-tile, k-major shared-memory tiles.  These are the exact-fp32 path; see tc_gemm.cu for
// the tcgen05 path.
#include "mms_common.cuh"

namespace mms {

constexpr int BM = 64, BN = 64, BK = 16, LDS_PAD = 4;

// Tile whose global rows are contiguous along the reduction index k:  S[k][r] = G[(r0+r)*ld + k0+k]
__device__ __forceinline__ void load_kcontig(float (*S)[BM + LDS_PAD], const float* __restrict__ G, int64_t ld, int r0,
                                             int rows, int k0, int K, bool vec_ok) {
    const int t = threadIdx.x, r = t >> 2, kq = (t & 3) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r0 + r < rows) {
        const float* p = G + (int64_t)(r0 + r) * ld + k0 + kq;
        if (vec_ok && k0 + kq + 3 < K) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k0 + kq + j < K) v[j] = __ldg(p + j);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) S[kq + j][r] = v[j];
}

// Tile whose global rows are indexed by the reduction index:  S[k][c] = G[(k0+k)*ld + c0+c]
__device__ __forceinline__ void load_ncontig(float (*S)[BN + LDS_PAD], const float* __restrict__ G, int64_t ld, int k0,
                                             int K, int c0, int cols, bool vec_ok) {
    const int t = threadIdx.x, k = t >> 4, c4 = (t & 15) * 4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + k < K) {
        const float* p = G + (int64_t)(k0 + k) * ld + c0 + c4;
        if (vec_ok && c0 + c4 + 3 < cols) {
            q = __ldg(reinterpret_cast<const float4*>(p));
        } else {
            if (c0 + c4 + 0 < cols) q.x = __ldg(p + 0);
            if (c0 + c4 + 1 < cols) q.y = __ldg(p + 1);
            if (c0 + c4 + 2 < cols) q.z = __ldg(p + 2);
            if (c0 + c4 + 3 < cols) q.w = __ldg(p + 3);
        }
    }
    *reinterpret_cast<float4*>(&S[k][c4]) = q;
}

__device__ __forceinline__ void tile_fma(const float (*As)[BM + LDS_PAD], const float (*Bs)[BN + LDS_PAD], float acc[4][4]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// C[m,n] = sum_k A[m,k] W[n,k] + bias[n]
__global__ void __launch_bounds__(256) gemm_nt_bias_kernel(const float* __restrict__ A, int64_t lda,
                                                           const float* __restrict__ W, int64_t ldw,
                                                           const float* __restrict__ bias, float* __restrict__ C, int64_t ldc,
                                                           int M, int N, int K, int vecA, int vecW) {
    __shared__ __align__(16) float As[BK][BM + LDS_PAD];
    __shared__ __align__(16) float Bs[BK][BN + LDS_PAD];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        load_kcontig(As, A, lda, m0, M, k0, K, vecA);
        load_kcontig(Bs, W, ldw, n0, N, k0, K, vecW);
        __syncthreads();
        tile_fma(As, Bs, acc);
        __syncthreads();
    }
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) C[(int64_t)m * ldc + n] = acc[i][j] + (bias ? bias[n] : 0.f);
        }
    }
}

// C[m,n] (+)= sum_k A[m,k] W[k,n]
__global__ void __launch_bounds__(256) gemm_nn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W,
                                                      int64_t ldw, float* __restrict__ C, int64_t ldc, int M, int N, int K,
                                                      int accumulate, int vecA, int vecW) {
    __shared__ __align__(16) float As[BK][BM + LDS_PAD];
    __shared__ __align__(16) float Bs[BK][BN + LDS_PAD];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        load_kcontig(As, A, lda, m0, M, k0, K, vecA);
        load_ncontig(Bs, W, ldw, k0, K, n0, N, vecW);
        __syncthreads();
        tile_fma(As, Bs, acc);
        __syncthreads();
    }
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) {
                float* c = C + (int64_t)m * ldc + n;
                *c = accumulate ? *c + acc[i][j] : acc[i][j];
            }
        }
    }
}

// C[i,j] += sum_m A[m, acol(i)] * Bm[row(m), j];  bias_grad[i] += sum_m A[m, acol(i)]
// grid = (ceil(N2/64), ceil(N1/64), m_splits); every CTA reduces `chunk` rows and adds with atomics.
__global__ void __launch_bounds__(256) gemm_tn_acc_kernel(const float* __restrict__ A, int64_t lda, int a_split, int a_skip,
                                                          const float* __restrict__ Bm, int64_t ldb, int shift, int seq,
                                                          float* __restrict__ C, int64_t ldc, float* __restrict__ bias_grad,
                                                          int M, int N1, int N2, int chunk) {
    __shared__ __align__(16) float As[BK][BM + LDS_PAD];
    __shared__ __align__(16) float Bs[BK][BN + LDS_PAD];
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int mbeg = blockIdx.z * chunk, mend = min(M, mbeg + chunk);
    float acc[4][4] = {};
    float bsum = 0.f;
    const int t = threadIdx.x, kk = t >> 4, c4 = (t & 15) * 4;
    for (int m0 = mbeg; m0 < mend; m0 += BK) {
        const int m = m0 + kk;
        // A tile: As[kk][i] = A[m, acol(i0+i)]
        {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < mend) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + c4 + j;
                    if (i < N1) v[j] = __ldg(A + (int64_t)m * lda + (i < a_split ? i : i + a_skip));
                }
            }
            *reinterpret_cast<float4*>(&As[kk][c4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        // B tile: Bs[kk][j] = Bm[row(m), j0+j] with the h_{t-1} shift inside each sequence
        {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < mend) {
                const int tt = (m % seq) + shift;
                if (tt >= 0 && tt < seq) {
                    const float* p = Bm + (int64_t)(m + shift) * ldb + j0 + c4;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j0 + c4 + j < N2) v[j] = __ldg(p + j);
                }
            }
            *reinterpret_cast<float4*>(&Bs[kk][c4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
        tile_fma(As, Bs, acc);
        if (bias_grad && blockIdx.x == 0 && t < BM) {
#pragma unroll
            for (int k = 0; k < BK; ++k) bsum += As[k][t];
        }
        __syncthreads();
    }
    const int tx = t & 15, ty = t >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ii = i0 + ty * 4 + i;
        if (ii >= N1) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = j0 + tx * 4 + j;
            if (jj < N2) atomicAdd(C + (int64_t)ii * ldc + jj, acc[i][j]);
        }
    }
    if (bias_grad && blockIdx.x == 0 && t < BM && i0 + t < N1) atomicAdd(bias_grad + i0 + t, bsum);
}

// Products with few rows (M <= a few dozen: the B rows of the top layer's single reverse step, forward and backward).  The
// 64 x 64-tile kernels above run them on N / 64 = 2-3 CTAs with a serial load -> barrier -> FMA loop per 16 reduction
// elements (14-18 us for 1.5 MFLOP, and the recurrence that follows waits for them); here a CTA owns ALL reduction elements
// of 32 rows x 8 columns, stages both operands with every load in flight at once, and N / 8 CTAs run side by side.
//   C[m,n] = sum_k a(m,k) W(n,k) (+ bias[n]),   W(n,k) = W[n*ldw + k] (w_kmajor) or W[k*ldw + n]
//   drop_mode 1: a(m,k) = A[m,k] * mult(drop_base + m*drop_row_stride + k)   (inter-layer dropout on the operand)
//   drop_mode 2: C[m,n] *= mult(drop_base + m*drop_row_stride + n)            (its gradient, on the result)
constexpr int SK_ROWS = 32, SK_COLS = 8, SK_MAXK = 256;

__global__ void __launch_bounds__(256) gemm_skinny_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W,
                                                          int64_t ldw, int w_kmajor, const float* __restrict__ bias,
                                                          float* __restrict__ C, int64_t ldc, int M, int N, int K, int drop_mode,
                                                          int64_t drop_base, int64_t drop_row_stride, float p, uint64_t seed,
                                                          uint64_t offset, const int64_t* offset_dev) {
    extern __shared__ __align__(16) float sk_smem[];
    const int KP = K + 4;                                  // row pitch: rows 4 banks apart -> conflict-free 128-bit reads
    float* As = sk_smem;                                   // [SK_ROWS][KP]
    float* Ws = sk_smem + SK_ROWS * KP;                    // [SK_COLS][KP]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * SK_ROWS, n0 = blockIdx.x * SK_COLS;
    const int K4 = K >> 2;
    DropRng rng;
    if (drop_mode) rng.init(seed, resolve_offset(offset, offset_dev), p);

    for (int i = tid; i < SK_ROWS * K4; i += 256) {
        const int r = i / K4, k = (i - r * K4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + r < M) {
            v = __ldg(reinterpret_cast<const float4*>(A + (int64_t)(m0 + r) * lda + k));
            if (drop_mode == 1) {
                const uint64_t e = (uint64_t)(drop_base + (int64_t)(m0 + r) * drop_row_stride + k);
                v.x *= rng.mult(e); v.y *= rng.mult(e + 1); v.z *= rng.mult(e + 2); v.w *= rng.mult(e + 3);
            }
        }
        *reinterpret_cast<float4*>(As + r * KP + k) = v;
    }
    if (w_kmajor) {
        for (int i = tid; i < SK_COLS * K4; i += 256) {
            const int c = i / K4, k = (i - c * K4) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + c < N) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(n0 + c) * ldw + k));
            *reinterpret_cast<float4*>(Ws + c * KP + k) = v;
        }
    } else {
        for (int i = tid; i < K * 2; i += 256) {           // two 16-byte pieces of the 8 columns per reduction index
            const int k = i >> 1, c = (i & 1) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + c + 3 < N) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)k * ldw + n0 + c));
            else {
                if (n0 + c + 0 < N) v.x = __ldg(W + (int64_t)k * ldw + n0 + c + 0);
                if (n0 + c + 1 < N) v.y = __ldg(W + (int64_t)k * ldw + n0 + c + 1);
                if (n0 + c + 2 < N) v.z = __ldg(W + (int64_t)k * ldw + n0 + c + 2);
            }
            Ws[(c + 0) * KP + k] = v.x; Ws[(c + 1) * KP + k] = v.y; Ws[(c + 2) * KP + k] = v.z; Ws[(c + 3) * KP + k] = v.w;
        }
    }
    __syncthreads();

    // thread (c, r) = (tid & 7, tid >> 3) owns one output: a warp reads 4 rows of A and 8 rows of W, 128 bits per lane
    const int c = tid & 7, r = tid >> 3;                   // r in 0..31
    const float4* ap = reinterpret_cast<const float4*>(As + r * KP);
    const float4* wp = reinterpret_cast<const float4*>(Ws + c * KP);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 4
    for (int k4 = 0; k4 < K4; ++k4) {
        const float4 a = ap[k4], w = wp[k4];
        acc0 = fmaf(a.x, w.x, acc0);
        acc1 = fmaf(a.y, w.y, acc1);
        acc2 = fmaf(a.z, w.z, acc2);
        acc3 = fmaf(a.w, w.w, acc3);
    }
    const int m = m0 + r, n = n0 + c;
    if (m < M && n < N) {
        float v = (acc0 + acc1) + (acc2 + acc3);
        if (bias) v += __ldg(bias + n);
        if (drop_mode == 2) v *= rng.mult((uint64_t)(drop_base + (int64_t)m * drop_row_stride + n));
        C[(int64_t)m * ldc + n] = v;
    }
}

bool gemm_skinny_supported(const float* A, int64_t lda, const float* W, int64_t ldw, int w_kmajor, int M, int N, int K) {
    if (option_get("GEMM_SKINNY", 1) != 1) return false;
    if (M < 1 || M > 256 || N < 1 || K < 4 || K > SK_MAXK || K % 4 != 0) return false;
    if (!aligned16(A) || lda % 4 != 0 || !aligned16(W) || ldw % 4 != 0) return false;
    return true;
}

int launch_gemm_skinny(const float* A, int64_t lda, const float* W, int64_t ldw, int w_kmajor, const float* bias, float* C, int64_t ldc,
                       int M, int N, int K, int drop_mode, int64_t drop_base, int64_t drop_row_stride, float p, uint64_t seed,
                       uint64_t offset, const int64_t* offset_dev, cudaStream_t st) {
    MMS_REQUIRE(gemm_skinny_supported(A, lda, W, ldw, w_kmajor, M, N, K), "gemm_skinny: unsupported shape / alignment");
    if (p <= 0.f) drop_mode = 0;
    const size_t smem = (size_t)(SK_ROWS + SK_COLS) * (K + 4) * sizeof(float);      // <= 41.6 KB
    dim3 grid(cdiv(N, SK_COLS), cdiv(M, SK_ROWS));
    MMS_PROF_BEGIN(st);
    gemm_skinny_kernel<<<grid, 256, smem, st>>>(A, lda, W, ldw, w_kmajor, bias, C, ldc, M, N, K, drop_mode, drop_base, drop_row_stride,
                                                p, seed, offset, offset_dev);
    MMS_LAUNCH_CHECK("gemm_skinny_kernel");
    return MMS_OK;
}

int launch_gemm_nt_bias(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
                        int M, int N, int K, cudaStream_t st) {
    if (M <= 0 || N <= 0) return MMS_OK;
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    MMS_PROF_BEGIN(st);
    gemm_nt_bias_kernel<<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, aligned16(A) && lda % 4 == 0,
                                              aligned16(W) && ldw % 4 == 0);
    MMS_LAUNCH_CHECK("gemm_nt_bias_kernel");
    return MMS_OK;
}

int launch_gemm_nn(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int M, int N, int K,
                   int accumulate, cudaStream_t st) {
    if (M <= 0 || N <= 0) return MMS_OK;
    dim3 grid(cdiv(N, BN), cdiv(M, BM));
    MMS_PROF_BEGIN(st);
    gemm_nn_kernel<<<grid, 256, 0, st>>>(A, lda, W, ldw, C, ldc, M, N, K, accumulate, aligned16(A) && lda % 4 == 0,
                                         aligned16(W) && ldw % 4 == 0);
    MMS_LAUNCH_CHECK("gemm_nn_kernel");
    return MMS_OK;
}

int launch_gemm_tn_acc(const float* A, int64_t lda, int a_split, int a_skip, const float* Bm, int64_t ldb, int shift, int seq,
                       float* C, int64_t ldc, float* bias_grad, int M, int N1, int N2, cudaStream_t st) {
    if (M <= 0 || N1 <= 0) return MMS_OK;
    MMS_REQUIRE(seq >= 1 && (shift == 0 || M % seq == 0), "gemm_tn: M must be a multiple of seq when rows are shifted");
    int chunk = 256;
    while (chunk < M && cdiv(M, chunk) > 96) chunk *= 2;
    dim3 grid(cdiv(N2 > 0 ? N2 : 1, BN), cdiv(N1, BM), cdiv(M, chunk));
    MMS_PROF_BEGIN(st);
    gemm_tn_acc_kernel<<<grid, 256, 0, st>>>(A, lda, a_split, a_skip, Bm, ldb, shift, seq, C, ldc, bias_grad, M, N1, N2, chunk);
    MMS_LAUNCH_CHECK("gemm_tn_acc_kernel");
    return MMS_OK;
}

}  // namespace mms

using namespace mms;

extern "C" int mms_gemm_nt_bias(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C,
                                int64_t ldc, int32_t M, int32_t N, int32_t K, mms_stream_t stream) {
    MMS_REQUIRE(A && W && C && K > 0, "gemm_nt_bias: bad arguments");
    return launch_gemm_nt_bias(A, lda, W, ldw, bias, C, ldc, M, N, K, (cudaStream_t)stream);
}
extern "C" int mms_gemm_skinny(const float* A, int64_t lda, const float* W, int64_t ldw, int32_t w_kmajor, const float* bias, float* C,
                               int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t drop_mode, int64_t drop_base,
                               int64_t drop_row_stride, float dropout_p, uint64_t rng_seed, uint64_t rng_offset,
                               const int64_t* rng_offset_dev, mms_stream_t stream) {
    MMS_REQUIRE(A && W && C && drop_mode >= 0 && drop_mode <= 2 && dropout_p >= 0.f && dropout_p < 1.f, "gemm_skinny: bad arguments");
    return launch_gemm_skinny(A, lda, W, ldw, w_kmajor, bias, C, ldc, M, N, K, drop_mode, drop_base, drop_row_stride, dropout_p, rng_seed,
                              rng_offset, rng_offset_dev, (cudaStream_t)stream);
}
extern "C" int mms_gemm_nn(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int32_t M,
                           int32_t N, int32_t K, int32_t accumulate, mms_stream_t stream) {
    MMS_REQUIRE(A && W && C && K > 0, "gemm_nn: bad arguments");
    return launch_gemm_nn(A, lda, W, ldw, C, ldc, M, N, K, accumulate, (cudaStream_t)stream);
}
extern "C" int mms_gemm_tn_acc(const float* A, int64_t lda, int32_t a_split, int32_t a_skip, const float* Bm, int64_t ldb,
                               int32_t shift, int32_t seq, float* C, int64_t ldc, float* bias_grad, int32_t M, int32_t N1,
                               int32_t N2, mms_stream_t stream) {
    MMS_REQUIRE(A && (N2 == 0 || (Bm && C)), "gemm_tn_acc: bad arguments");
    return launch_gemm_tn_acc(A, lda, a_split, a_skip, Bm, ldb, shift, seq, C, ldc, bias_grad, M, N1, N2, (cudaStream_t)stream);
}
