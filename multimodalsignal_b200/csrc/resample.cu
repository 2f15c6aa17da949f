// FFT resampling with scipy.signal.resample semantics (reference preprocess.py:70-75; algorithm of
// scipy/signal/_signaltools.py for real input, no window):
//     X = rfft(x);  keep the m2 = min(N, num)//2 + 1 lowest bins;  fix the unpaired bin when
//     min(N, num) is even and num != N;  y = irfft(X * num/N, n = num).
// Recording lengths are arbitrary (N ~ 4.2e6 with large prime factors), so both transforms are
// evaluated as chirp-z transforms (Bluestein): a length-P DFT restricted to the bins that are
// actually needed becomes a circular convolution of power-of-two length M, done with radix-2^k
// passes through shared memory in float64.
//   stage A:  X[k] = sum_n x[n] e^{-2 pi i nk/N},           k < m2   (only the kept bins)
//   stage C:  y[j] = Re sum_k G[k] e^{+2 pi i jk/num} / num,  j < num  (one-sided Hermitian sum)
// The forward passes leave the spectrum in digit-reversed order and the inverse passes undo exactly
// that permutation, so no reordering pass exists: the point-wise product with the (identically
// permuted) chirp-filter spectrum happens in the permuted domain.
// HBM-bound: every pass streams the M complex values once (32*M bytes of traffic).
// Two REAL signals share one COMPLEX stage-A transform (z = x1 + i x2, the textbook pairing): the chirp-z transform is asked
// for the bins -(m2-1) .. m2-1 of z instead of 0 .. m2-1 (same convolution length for the decimating case N >> num), and
// X1[k] = (Z[k] + conj Z[-k]) / 2, X2[k] = (Z[k] - conj Z[-k]) / 2i are separated while the spectrum is fixed up.  Stage A is
// ~90 % of the traffic of a 700 Hz -> 64 Hz resampling, so the bytes moved per signal nearly halve.
#include "mms_common.cuh"
#include "fft_fast.cuh"
#include <string.h>

namespace mms {

constexpr int FFT_TC = 16;          // tile columns
constexpr int FFT_THREADS = 256;

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// e^{sign * pi * i * (n^2 mod 2P) / P} and the chirp times the linear phase e^{sign * 2 pi i n k / P} (0 <= k < P): fft_fast.cuh
__device__ __forceinline__ double2 chirp(int64_t n, int64_t P, int sign) { return ff::chirp_at(n, P, sign); }
__device__ __forceinline__ double2 chirp_shift(int64_t n, int64_t k, int64_t P, int sign) { return ff::chirp_shift_at(n, k, P, sign); }

__device__ __forceinline__ int bitrev(int v, int bits) { return bits ? (int)(__brev((unsigned)v) >> (32 - bits)) : 0; }

// One radix-R pass over blocks of size Mc = R*S.  Forward (DIF): R-point butterflies over the
// strided index, then the twiddle w_Mc^{r*k1}.  Inverse (DIT): conjugate twiddle first, then the
// butterflies in reverse stage order -- the exact inverse of the forward pass (up to the factor R).
template <int R>
__global__ void __launch_bounds__(FFT_THREADS) fft_pass_kernel(double2* __restrict__ a, int64_t sig_stride, int64_t Mc, int tc,
                                                               int inverse) {
    constexpr int LOGR = R == 256 ? 8 : R == 128 ? 7 : R == 64 ? 6 : R == 32 ? 5 : R == 16 ? 4 : R == 8 ? 3 : R == 4 ? 2 : 1;
    constexpr int LD = FFT_TC + 1;
    extern __shared__ __align__(16) double2 fsm[];
    double2* tile = fsm;                 // [R][LD]
    double2* tw = fsm + R * LD;          // [R/2]  w_R^j (conjugated for the inverse)
    const int tid = threadIdx.x;
    const int64_t S = Mc / R;
    double2* base = a + (int64_t)blockIdx.y * sig_stride;
    const int64_t g0 = (int64_t)blockIdx.x * tc;          // first (block, r) column of this tile
    const double tsign = inverse ? 1.0 : -1.0;

    for (int j = tid; j < R / 2; j += FFT_THREADS) {
        double s, c;
        sincospi(2.0 * (double)j / (double)R, &s, &c);
        tw[j] = make_double2(c, tsign * s);
    }
    const bool rowwise = S >= tc;
    const int64_t blk0 = g0 / S, r0 = g0 % S;
    double2* region = base + blk0 * Mc;                    // contiguous tc*R elements when !rowwise
    const int total = R * tc;

    auto twiddle = [&](int qrow, int64_t r) {
        const int k1 = bitrev(qrow, LOGR);
        const uint64_t e = ((uint64_t)r * (uint64_t)k1) & (uint64_t)(Mc - 1);
        double s, c;
        sincospi(2.0 * (double)e / (double)Mc, &s, &c);
        return make_double2(c, tsign * s);
    };

    for (int idx = tid; idx < total; idx += FFT_THREADS) {
        int q, c;
        int64_t r;
        double2 v;
        if (rowwise) {
            q = idx / tc; c = idx - q * tc; r = r0 + c;
            v = base[blk0 * Mc + (int64_t)q * S + r];
        } else {
            const int64_t blk = idx / Mc, within = idx - blk * Mc;
            q = (int)(within / S); r = within - (int64_t)q * S; c = (int)(blk * S + r);
            v = region[idx];
        }
        if (inverse && S > 1) v = cmul(v, twiddle(q, r));
        tile[q * LD + c] = v;
    }
    __syncthreads();

    const int nbf = (R / 2) * tc;
    if (!inverse) {
#pragma unroll 1
        for (int h = R / 2; h >= 1; h >>= 1) {
            const int tstep = R / (2 * h);
            for (int bf = tid; bf < nbf; bf += FFT_THREADS) {
                const int c = bf % tc, pair = bf / tc, grp = pair / h, j = pair - grp * h;
                const int q0 = grp * 2 * h + j, q1 = q0 + h;
                const double2 x0 = tile[q0 * LD + c], x1 = tile[q1 * LD + c];
                tile[q0 * LD + c] = make_double2(x0.x + x1.x, x0.y + x1.y);
                tile[q1 * LD + c] = cmul(make_double2(x0.x - x1.x, x0.y - x1.y), tw[j * tstep]);
            }
            __syncthreads();
        }
    } else {
#pragma unroll 1
        for (int h = 1; h <= R / 2; h <<= 1) {
            const int tstep = R / (2 * h);
            for (int bf = tid; bf < nbf; bf += FFT_THREADS) {
                const int c = bf % tc, pair = bf / tc, grp = pair / h, j = pair - grp * h;
                const int q0 = grp * 2 * h + j, q1 = q0 + h;
                const double2 u0 = tile[q0 * LD + c];
                const double2 t = cmul(tile[q1 * LD + c], tw[j * tstep]);
                tile[q0 * LD + c] = make_double2(u0.x + t.x, u0.y + t.y);
                tile[q1 * LD + c] = make_double2(u0.x - t.x, u0.y - t.y);
            }
            __syncthreads();
        }
    }

    for (int idx = tid; idx < total; idx += FFT_THREADS) {
        int q, c;
        int64_t r;
        double2* dst;
        if (rowwise) {
            q = idx / tc; c = idx - q * tc; r = r0 + c;
            dst = base + blk0 * Mc + (int64_t)q * S + r;
        } else {
            const int64_t blk = idx / Mc, within = idx - blk * Mc;
            q = (int)(within / S); r = within - (int64_t)q * S; c = (int)(blk * S + r);
            dst = region + idx;
        }
        double2 v = tile[q * LD + c];
        if (!inverse && S > 1) v = cmul(v, twiddle(q, r));
        *dst = v;
    }
}

// a[sig][n] = x[sig][n] * chirp(n, P, sign) for n < n_in, zero up to M
__global__ void __launch_bounds__(256) czt_pre_real_kernel(const double* __restrict__ x, int64_t n_in, int64_t P, int sign,
                                                           double2* __restrict__ a, int64_t M) {
    const double* xs = x + (int64_t)blockIdx.y * n_in;
    double2* as = a + (int64_t)blockIdx.y * M;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < M; n += (int64_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (n < n_in) {
            const double2 c = chirp(n, P, sign);
            const double xv = xs[n];
            v = make_double2(xv * c.x, xv * c.y);
        }
        as[n] = v;
    }
}

// Paired version: a[p][n] = (x[2p][n] + i x[2p+1][n]) * chirp(n, P, sign) * e^{sign 2 pi i n k0 / P}; the second signal of the last
// pair may be missing (odd n_sig): zero imaginary part.
__global__ void __launch_bounds__(256) czt_pre_pair_kernel(const double* __restrict__ x, int64_t n_in, int n_sig, int64_t P, int64_t k0,
                                                           int sign, double2* __restrict__ a, int64_t M) {
    const int p = blockIdx.y;
    const double* x1 = x + (int64_t)(2 * p) * n_in;
    const double* x2 = 2 * p + 1 < n_sig ? x + (int64_t)(2 * p + 1) * n_in : nullptr;
    double2* as = a + (int64_t)p * M;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < M; n += (int64_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (n < n_in) {
            const double2 c = chirp_shift(n, k0, P, sign);
            const double re = x1[n], im = x2 ? x2[n] : 0.0;
            v = make_double2(re * c.x - im * c.y, re * c.y + im * c.x);
        }
        as[n] = v;
    }
}

// Chirp filter b[j mod M] = conj(chirp(j, P, sign)) for j in (-n_in, n_out), zero elsewhere.
__global__ void __launch_bounds__(256) czt_filter_kernel(double2* __restrict__ b, int64_t M, int64_t n_in, int64_t n_out,
                                                         int64_t P, int sign) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (i < n_out) v = chirp(i, P, -sign);
        else if (M - i < n_in) v = chirp(M - i, P, -sign);
        b[i] = v;
    }
}

__global__ void __launch_bounds__(256) cmul_kernel(double2* __restrict__ a, const double2* __restrict__ b, int64_t M) {
    double2* as = a + (int64_t)blockIdx.y * M;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
        as[i] = cmul(as[i], b[i]);
}

// Stage A epilogue fused with stage C prologue.  From the convolution result conv[k] (k < m2):
//   X[k] = conv[k]/M * chirp(k, N, -1) * (num/N)          the kept rfft bins, already rescaled
//   unpaired-bin fix, one-sided weights (G = 2X, except the real DC / Nyquist bins)
//   a2[k] = G[k] * chirp(k, num, +1),   zero up to M2
// Reads and writes the same buffer (in place, element-wise); `src_stride`/`dst_stride` are the
// per-signal strides of the two layouts (both views of the same allocation, processed signal by
// signal from the host so that they never overlap destructively).
__global__ void __launch_bounds__(256) spectrum_fix_kernel(const double2* __restrict__ conv, double2* __restrict__ a2, int64_t M,
                                                           int64_t M2, int64_t N, int64_t num, int64_t m2) {
    const int64_t m = N < num ? N : num;
    const double scale = ((double)num / (double)N) / (double)M;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M2; k += (int64_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (k < m2) {
            double2 X = cmul(conv[k], chirp(k, N, -1));
            X.x *= scale; X.y *= scale;
            if ((m & 1) == 0 && num != N && k == m / 2) {
                const double f = num < N ? 2.0 : 0.5;
                X.x *= f; X.y *= f;
            }
            double2 G = make_double2(2.0 * X.x, 2.0 * X.y);
            if (k == 0) G = make_double2(X.x, 0.0);
            if ((num & 1) == 0 && k == num / 2) G = make_double2(X.x, 0.0);
            v = cmul(G, chirp(k, num, +1));
        }
        a2[k] = v;
    }
}

// Paired version of spectrum_fix_kernel: conv holds the bins k = j - (m2 - 1), j < 2 m2 - 1, of z = x1 + i x2 (pair blockIdx.y at
// conv + blockIdx.y * M); the two real signals' spectra are separated, fixed up as above and written to a2 + (2p) * M2 and
// a2 + (2p + 1) * M2 (a different region of the workspace: nothing is read after it has been overwritten).
__global__ void __launch_bounds__(256, 2) spectrum_fix_pair_kernel(const double2* __restrict__ conv, double2* __restrict__ a2, int n_sig,
                                                                int64_t M, int64_t M2, int64_t N, int64_t num, int64_t m2) {
    const int p = blockIdx.y;
    const double2* cv = conv + (int64_t)p * M;
    double2* o1 = a2 + (int64_t)(2 * p) * M2;
    double2* o2 = 2 * p + 1 < n_sig ? a2 + (int64_t)(2 * p + 1) * M2 : nullptr;
    const int64_t m = N < num ? N : num;
    const double scale = ((double)num / (double)N) / (double)M;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M2; k += (int64_t)gridDim.x * blockDim.x) {
        double2 v1 = make_double2(0.0, 0.0), v2 = v1;
        if (k < m2) {
            const int64_t jp = m2 - 1 + k, jn = m2 - 1 - k;
            const double2 Zp = cmul(cv[jp], chirp(jp, N, -1));
            const double2 Zn = cmul(cv[jn], chirp(jn, N, -1));
            double2 X1 = make_double2(0.5 * (Zp.x + Zn.x) * scale, 0.5 * (Zp.y - Zn.y) * scale);
            double2 X2 = make_double2(0.5 * (Zp.y + Zn.y) * scale, -0.5 * (Zp.x - Zn.x) * scale);
            if ((m & 1) == 0 && num != N && k == m / 2) {
                const double f = num < N ? 2.0 : 0.5;
                X1.x *= f; X1.y *= f; X2.x *= f; X2.y *= f;
            }
            double2 G1 = make_double2(2.0 * X1.x, 2.0 * X1.y), G2 = make_double2(2.0 * X2.x, 2.0 * X2.y);
            if (k == 0 || ((num & 1) == 0 && k == num / 2)) { G1 = make_double2(X1.x, 0.0); G2 = make_double2(X2.x, 0.0); }
            const double2 c = chirp(k, num, +1);
            v1 = cmul(G1, c);
            v2 = cmul(G2, c);
        }
        o1[k] = v1;
        if (o2) o2[k] = v2;
    }
}

// y[j] = Re(conv[j] * chirp(j, num, +1)) / (M2 * num)
__global__ void __launch_bounds__(256) czt_post_real_kernel(const double2* __restrict__ a, int64_t M2, int64_t num,
                                                            double* __restrict__ y) {
    const double2* as = a + (int64_t)blockIdx.y * M2;
    double* ys = y + (int64_t)blockIdx.y * num;
    const double scale = 1.0 / ((double)M2 * (double)num);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < num; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 c = chirp(j, num, +1);
        const double2 v = as[j];
        ys[j] = (v.x * c.x - v.y * c.y) * scale;
    }
}


// Paired stage A -> PAIRED stage C (MMS_RESAMPLE_PAIRED_C): the two real OUTPUT signals of a pair come out of ONE complex inverse
// chirp-z transform.  A real y_s[j] = Re sum_{k=0..K} G_s[k] e^{i theta jk} is the two-sided sum over k = -K .. K of c_s[k] = G_s[k] / 2
// (k > 0), conj(G_s[-k]) / 2 (k < 0), G_s[0] (k = 0), so y_1 + i y_2 = sum_k (c_1[k] + i c_2[k]) e^{i theta jk}: a transform with
// 2K + 1 = 2 m2 - 1 inputs (index n = k + K; the shift by K is the linear phase e^{-i theta jK} of the post kernel) instead of two
// with m2 inputs each -- for the 700 -> 64 Hz case 4 transforms of 3 * 2^18 points instead of 8 of 9 * 2^16.
__global__ void __launch_bounds__(256, 2) spectrum_fix_pair2_kernel(const double2* __restrict__ conv, double2* __restrict__ a2, int n_sig,
                                                                    int64_t M, int64_t M2, int64_t N, int64_t num, int64_t m2) {
    const int p = blockIdx.y;
    const double2* cv = conv + (int64_t)p * M;
    double2* out = a2 + (int64_t)p * M2;
    const bool has2 = 2 * p + 1 < n_sig;
    const int64_t m = N < num ? N : num, K = m2 - 1;
    const double scale = ((double)num / (double)N) / (double)M;
    // one thread per |k|: X1, X2 of bin |k| give both d[+k] (index K + k) and d[-k] (index K - k); the tail beyond 2K is zeroed
    for (int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; kk < M2 - K; kk += (int64_t)gridDim.x * blockDim.x) {
        if (kk > K) {
            out[K + kk] = make_double2(0.0, 0.0);
            continue;
        }
        const int64_t jp = K + kk, jn = K - kk;
        const double2 Zp = cmul(cv[jp], chirp(jp, N, -1));
        const double2 Zn = cmul(cv[jn], chirp(jn, N, -1));
        double2 X1 = make_double2(0.5 * (Zp.x + Zn.x) * scale, 0.5 * (Zp.y - Zn.y) * scale);
        double2 X2 = make_double2(0.5 * (Zp.y + Zn.y) * scale, -0.5 * (Zp.x - Zn.x) * scale);
        if ((m & 1) == 0 && num != N && kk == m / 2) {
            const double f = num < N ? 2.0 : 0.5;
            X1.x *= f; X1.y *= f; X2.x *= f; X2.y *= f;
        }
        double2 G1 = make_double2(2.0 * X1.x, 2.0 * X1.y), G2 = make_double2(2.0 * X2.x, 2.0 * X2.y);
        if (kk == 0 || ((num & 1) == 0 && kk == num / 2)) { G1 = make_double2(X1.x, 0.0); G2 = make_double2(X2.x, 0.0); }
        if (!has2) G2 = make_double2(0.0, 0.0);
        if (kk == 0) {
            out[K] = cmul(make_double2(G1.x, G2.x), chirp(K, num, +1));
        } else {
            out[K + kk] = cmul(make_double2(0.5 * (G1.x - G2.y), 0.5 * (G1.y + G2.x)), chirp(K + kk, num, +1));
            out[K - kk] = cmul(make_double2(0.5 * (G1.x + G2.y), 0.5 * (G2.x - G1.y)), chirp(K - kk, num, +1));
        }
    }
}

// (y_1 + i y_2)[j] = conv[j] * chirp(j) * e^{-2 pi i jK / num} / (M2 * num)
__global__ void __launch_bounds__(256) czt_post_pair_kernel(const double2* __restrict__ a, int64_t M2, int64_t num, int64_t K, int n_sig,
                                                            double* __restrict__ y) {
    const int p = blockIdx.y;
    const double2* as = a + (int64_t)p * M2;
    double* y1 = y + (int64_t)(2 * p) * num;
    double* y2 = 2 * p + 1 < n_sig ? y + (int64_t)(2 * p + 1) * num : nullptr;
    const double scale = 1.0 / ((double)M2 * (double)num);
    const int64_t ks = (num - K % num) % num;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < num; j += (int64_t)gridDim.x * blockDim.x) {
        const double2 z = cmul(as[j], chirp_shift(j, ks, num, +1));
        y1[j] = z.x * scale;
        if (y2) y2[j] = z.y * scale;
    }
}

// ---- register-resident passes (fft_fast.cuh) ---------------------------------------------------------------------------
// One CTA = one tile of TC columns, 256 threads, one block barrier; two CTAs per SM (a thread holds 16 complex doubles).
// grid = (signals, tiles): the CTAs of one tile position run next to each other, so the filter spectrum tile they all multiply
// by (last forward pass) is read from DRAM once.
template <int N1, int N2, int TC>
__global__ void __launch_bounds__(256, 2) fft_fast_fwd_kernel(const ff::PassArgs p) {
    extern __shared__ __align__(16) double2 ffs[];
    ff::fwd_stage1<N1, N2, TC, 256>(p, threadIdx.x, blockIdx.y, blockIdx.x, ffs);
    __syncthreads();
    ff::fwd_stage2<N1, N2, TC, 256>(p, threadIdx.x, blockIdx.y, blockIdx.x, ffs);
}
template <int N1, int N2, int TC>
__global__ void __launch_bounds__(256, 2) fft_fast_inv_kernel(const ff::PassArgs p) {
    extern __shared__ __align__(16) double2 ffs[];
    ff::inv_stage0<N1, N2, TC, 256>(p, threadIdx.x, blockIdx.y, ffs);
    __syncthreads();
    ff::inv_stage2<N1, N2, TC, 256>(p, threadIdx.x, blockIdx.y, blockIdx.x, ffs);
    __syncthreads();
    ff::inv_stage1<N1, N2, TC, 256>(p, threadIdx.x, blockIdx.y, blockIdx.x, ffs);
}

template <int N1, int N2, int TC>
static int launch_fast_pass(const ff::PassArgs& p, int n_sig, int inverse, cudaStream_t st) {
    constexpr int R = N1 * N2;
    const size_t smem = (size_t)ff::tile_elems<R, N2, TC>() * sizeof(double2);
    auto kf = fft_fast_fwd_kernel<N1, N2, TC>;
    auto ki = fft_fast_inv_kernel<N1, N2, TC>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) {
        MMS_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MMS_CUDA(cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int64_t tiles = (p.ncols + TC - 1) / TC;
    MMS_REQUIRE(tiles <= 65535, "fft_fast: %lld tiles exceed the grid limit", (long long)tiles);
    dim3 grid(n_sig, (unsigned)tiles);
    MMS_PROF_BEGIN(st);
    if (inverse) ki<<<grid, 256, smem, st>>>(p);
    else kf<<<grid, 256, smem, st>>>(p);
    MMS_LAUNCH_CHECK(inverse ? "fft_fast_inv_kernel" : "fft_fast_fwd_kernel");
    return MMS_OK;
}

// Hooks of a fast transform: loads of the first forward pass, multiplier of the last forward pass, outputs kept by the last
// inverse pass (0 = all).  Only the fields named here are read.
struct FastHooks {
    int load_op = ff::LD_PLAIN;
    const double* x = nullptr;
    int paired = 0, n_x = 0, sign = 0;
    int64_t n_in = 0, n_out = 0, P = 0, k0 = 0;
    const double2* mul = nullptr;
    cudaEvent_t mul_ready = nullptr;       // recorded on another stream once `mul` is complete: waited for in front of the last forward pass
    int64_t n_keep = 0;
};

// Side stream of the filter transforms.  The two chirp-filter spectra of a call depend on the lengths only, are needed by the
// LAST forward pass of their stage, and are transforms of ONE signal whose grids leave most of the GPU idle: for long signals
// they run beside the first passes of the signal transforms instead of in front of them.
struct FilterStream {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, ready_a = nullptr, ready_c = nullptr;
    bool ok = false;
};
static FilterStream* filter_stream() {
    static FilterStream per_dev[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    FilterStream& f = per_dev[dev];
    if (!f.ok) {
        if (cudaStreamCreateWithFlags(&f.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;    // default priority: highest measured slower (1.435 vs 1.403 ms)
        if (cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&f.ready_a, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&f.ready_c, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        f.ok = true;
    }
    return &f;
}

static int fft_fast(double2* a, int64_t sig_stride, int64_t M, int n_sig, int inverse, const FastHooks& h, cudaStream_t st) {
    const ff::FastPlan pl = ff::make_fast_plan(M);
    MMS_REQUIRE(pl.ok, "fft_fast: no plan for length %lld", (long long)M);
    for (int ii = 0; ii < pl.n; ++ii) {
        const int i = inverse ? pl.n - 1 - ii : ii;
        const ff::FastPass& fp = pl.p[i];
        ff::PassArgs p;
        memset(&p, 0, sizeof(p));
        p.a = a; p.sig_stride = sig_stride; p.M = M; p.Mc = fp.mc; p.contig = i == pl.n - 1 ? 1 : 0;
        p.ncols = M / (fp.n1 * fp.n2);
        p.n_keep = M;
        if (!inverse && i == 0) {
            p.load_op = h.load_op; p.x = h.x; p.paired = h.paired; p.n_in = h.n_in; p.n_x = h.n_x; p.P = h.P; p.k0 = h.k0; p.sign = h.sign;
            p.n_out = h.n_out;
        }
        if (!inverse && i == pl.n - 1) {
            p.mul = h.mul;
            if (h.mul && h.mul_ready) MMS_CUDA(cudaStreamWaitEvent(st, h.mul_ready, 0));
        }
        if (inverse && i == 0 && h.n_keep > 0) p.n_keep = h.n_keep;
        int rc;
        if (fp.n1 == 16 && fp.n2 == 16) rc = launch_fast_pass<16, 16, 16>(p, n_sig, inverse, st);
        else if (fp.n1 == 16 && fp.n2 == 8) rc = launch_fast_pass<16, 8, 32>(p, n_sig, inverse, st);
        else if (fp.n1 == 8 && fp.n2 == 8) rc = launch_fast_pass<8, 8, 64>(p, n_sig, inverse, st);
        else if (fp.n1 == 9 && fp.n2 == 16) rc = launch_fast_pass<9, 16, 16>(p, n_sig, inverse, st);
        else if (fp.n1 == 12 && fp.n2 == 16) rc = launch_fast_pass<12, 16, 16>(p, n_sig, inverse, st);
        else { set_error("fft_fast: no kernel for radix %d x %d", fp.n1, fp.n2); return MMS_E_INVALID; }
        if (rc) return rc;
    }
    return MMS_OK;
}

// ---- host side -----------------------------------------------------------------------------------
static int64_t pow2_at_least(int64_t v) { int64_t m = 1; while (m < v) m <<= 1; return m; }

struct FftPlan { int n; int radix[8]; };

static FftPlan make_plan(int64_t M) {
    int bits = 0;
    while (((int64_t)1 << bits) < M) ++bits;
    FftPlan p;
    p.n = bits == 0 ? 0 : (bits + 7) / 8;
    int left = bits;
    for (int i = 0; i < p.n; ++i) {
        const int b = (left + (p.n - i) - 1) / (p.n - i);   // spread the bits evenly, larger radices first
        p.radix[i] = 1 << b;
        left -= b;
    }
    return p;
}

template <int R>
static int launch_pass(double2* a, int64_t sig_stride, int64_t M, int64_t Mc, int n_sig, int inverse, cudaStream_t st) {
    const int64_t cols = M / R;
    const int tc = cols < FFT_TC ? (int)cols : FFT_TC;
    const size_t smem = (size_t)(R * (FFT_TC + 1) + R / 2 + 1) * sizeof(double2);
    auto kern = fft_pass_kernel<R>;
    static PerDeviceOnce attr_once;
    if (attr_once.need()) { MMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); }
    dim3 grid((unsigned)(cols / tc), n_sig);
    MMS_PROF_BEGIN(st);
    kern<<<grid, FFT_THREADS, smem, st>>>(a, sig_stride, Mc, tc, inverse);
    MMS_LAUNCH_CHECK("fft_pass_kernel");
    return MMS_OK;
}

static int run_pass(int R, double2* a, int64_t sig_stride, int64_t M, int64_t Mc, int n_sig, int inverse, cudaStream_t st) {
    switch (R) {
        case 256: return launch_pass<256>(a, sig_stride, M, Mc, n_sig, inverse, st);
        case 128: return launch_pass<128>(a, sig_stride, M, Mc, n_sig, inverse, st);
        case 64: return launch_pass<64>(a, sig_stride, M, Mc, n_sig, inverse, st);
        case 32: return launch_pass<32>(a, sig_stride, M, Mc, n_sig, inverse, st);
        case 16: return launch_pass<16>(a, sig_stride, M, Mc, n_sig, inverse, st);
        case 8: return launch_pass<8>(a, sig_stride, M, Mc, n_sig, inverse, st);
        case 4: return launch_pass<4>(a, sig_stride, M, Mc, n_sig, inverse, st);
        case 2: return launch_pass<2>(a, sig_stride, M, Mc, n_sig, inverse, st);
    }
    set_error("fft: unsupported radix %d", R);
    return MMS_E_INVALID;
}

// digit-reversed forward transform (inverse = 0) or its exact inverse (inverse = 1, unnormalised)
static int fft_inplace(double2* a, int64_t sig_stride, int64_t M, int n_sig, int inverse, cudaStream_t st) {
    const FftPlan p = make_plan(M);
    int64_t mc[8];
    int64_t cur = M;
    for (int i = 0; i < p.n; ++i) { mc[i] = cur; cur /= p.radix[i]; }
    for (int ii = 0; ii < p.n; ++ii) {
        const int i = inverse ? p.n - 1 - ii : ii;
        int rc = run_pass(p.radix[i], a, sig_stride, M, mc[i], n_sig, inverse, st);
        if (rc) return rc;
    }
    return MMS_OK;
}

static inline unsigned grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

struct ResampleDims { int64_t N, num, m2, M1, M2, Mmax; };

// The fast path (MMS_RESAMPLE_FAST, default 1): smooth convolution lengths 2^a / 3 * 2^a / 9 * 2^a just above what the
// chirp-z transforms need (the power of two above N + J - 1 = 4.58 M is 8.39 M; 9 * 2^19 = 4.72 M), register-resident
// passes with the chirp multiplications, the filter product and the output pruning fused in.  Workspace layout:
// filter A [M1] | filter C [M2] | stage-A transforms [nA][M1] | stage-C transforms [nC][M2].
struct FastDims { bool ok; bool paired, paired_c; int nA, nC; int64_t J, JC, M1, M2, Mmax; };

static FastDims fast_dims(const ResampleDims& d, int n_sig) {
    FastDims f;
    memset(&f, 0, sizeof(f));
    if (option_get("RESAMPLE_FAST", 1) != 1) return f;
    f.paired = n_sig >= 2 && option_get("RESAMPLE_PAIRED", 1) == 1;
    f.J = f.paired ? 2 * d.m2 - 1 : d.m2;
    f.nA = f.paired ? (n_sig + 1) / 2 : n_sig;
    f.paired_c = f.paired && option_get("RESAMPLE_PAIRED_C", 1) == 1;
    f.JC = f.paired_c ? 2 * d.m2 - 1 : d.m2;
    f.nC = f.paired_c ? f.nA : n_sig;
    f.M1 = ff::fast_length_at_least(d.N + f.J - 1);
    f.M2 = ff::fast_length_at_least(f.JC + d.num - 1);
    f.Mmax = f.M1 > f.M2 ? f.M1 : f.M2;
    f.ok = f.M1 > 0 && f.M2 > 0;
    return f;
}

static int resample_dims(int64_t n_in, int64_t n_out, ResampleDims* d) {
    MMS_REQUIRE(n_in >= 1 && n_out >= 1 && n_in < ((int64_t)1 << 26) && n_out < ((int64_t)1 << 26),
                "resample: lengths must be in [1, 2^26) (got %lld -> %lld)", (long long)n_in, (long long)n_out);
    d->N = n_in; d->num = n_out;
    const int64_t m = n_in < n_out ? n_in : n_out;
    d->m2 = m / 2 + 1;
    d->M1 = pow2_at_least(n_in + d->m2 - 1);
    d->M2 = pow2_at_least(d->m2 + n_out - 1);
    d->Mmax = d->M1 > d->M2 ? d->M1 : d->M2;
    return MMS_OK;
}

}  // namespace mms

using namespace mms;

extern "C" int64_t mms_resample_workspace_bytes(int64_t n_in, int64_t n_out, int32_t n_sig) {
    ResampleDims d;
    if (resample_dims(n_in, n_out, &d) || n_sig < 1) return -1;
    const FastDims f = fast_dims(d, n_sig);
    if (f.ok) return (int64_t)sizeof(double2) * (f.M1 + f.M2 + (int64_t)f.nA * f.M1 + (int64_t)f.nC * f.M2);
    return (int64_t)sizeof(double2) * d.Mmax * ((int64_t)n_sig + 1);
}

extern "C" int mms_resample_f64(const double* x, int64_t n_in, int64_t n_out, int32_t n_sig, double* y, void* workspace,
                                int64_t workspace_bytes, mms_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ResampleDims d;
    int rc = resample_dims(n_in, n_out, &d);
    if (rc) return rc;
    MMS_REQUIRE(x && y && workspace && n_sig >= 1, "resample_f64: bad arguments");
    const FastDims f = fast_dims(d, n_sig);
    if (f.ok) {
        const int64_t need_f = (int64_t)sizeof(double2) * (f.M1 + f.M2 + (int64_t)f.nA * f.M1 + (int64_t)f.nC * f.M2);
        if (workspace_bytes < need_f) {
            set_error("resample_f64: workspace %lld bytes < %lld", (long long)workspace_bytes, (long long)need_f);
            return MMS_E_WORKSPACE;
        }
        double2* filtA = (double2*)workspace;
        double2* filtC = filtA + f.M1;
        double2* workA = filtC + f.M2;
        double2* workC = workA + (int64_t)f.nA * f.M1;
        // The two chirp filter spectra (generated by the loads of their first pass): beside the signal transforms on the filter
        // stream for long signals (MMS_RESAMPLE_FILTER_STREAM, default 1), in front of them otherwise
        FastHooks hf, hg;
        hf.load_op = ff::LD_FILTER; hf.n_in = d.N; hf.n_out = f.J; hf.P = d.N; hf.sign = -1;
        hg.load_op = ff::LD_FILTER; hg.n_in = f.JC; hg.n_out = d.num; hg.P = d.num; hg.sign = +1;
        FilterStream* fs = f.M1 >= ((int64_t)1 << 20) && option_get("RESAMPLE_FILTER_STREAM", 1) == 1 ? filter_stream() : nullptr;
        cudaStream_t sf = fs ? fs->s : st;
        if (fs) {
            MMS_CUDA(cudaEventRecord(fs->fork, st));            // everything the caller enqueued before (and the previous call's reads of this workspace)
            MMS_CUDA(cudaStreamWaitEvent(sf, fs->fork, 0));
        }
        rc = fft_fast(filtA, f.M1, f.M1, 1, 0, hf, sf);
        if (rc) return rc;
        if (fs) MMS_CUDA(cudaEventRecord(fs->ready_a, sf));
        rc = fft_fast(filtC, f.M2, f.M2, 1, 0, hg, sf);
        if (rc) return rc;
        if (fs) MMS_CUDA(cudaEventRecord(fs->ready_c, sf));
        // stage A: the transforms of the (paired) signals with the chirp pre-multiplication on the loads and the filter product
        // on the last pass's stores; the inverse keeps the J outputs that are used
        FastHooks ha;
        ha.load_op = ff::LD_PAIR; ha.x = x; ha.paired = f.paired ? 1 : 0; ha.n_x = n_sig; ha.n_in = d.N; ha.P = d.N; ha.sign = -1;
        ha.k0 = f.paired ? d.N - (d.m2 - 1) : 0;           // first bin: -(m2 - 1) mod N
        ha.mul = filtA;
        ha.mul_ready = fs ? fs->ready_a : nullptr;
        rc = fft_fast(workA, f.M1, f.M1, f.nA, 0, ha, st);
        if (rc) return rc;
        FastHooks hi;
        hi.n_keep = f.J;
        rc = fft_fast(workA, f.M1, f.M1, f.nA, 1, hi, st);
        if (rc) return rc;
        if (f.paired_c) {
            dim3 grid(grid_for(f.M2), f.nA);
            MMS_PROF_BEGIN(st);
            spectrum_fix_pair2_kernel<<<grid, 256, 0, st>>>(workA, workC, n_sig, f.M1, f.M2, d.N, d.num, d.m2);
            MMS_LAUNCH_CHECK("spectrum_fix_pair2_kernel");
        } else if (f.paired) {
            dim3 grid(grid_for(f.M2), f.nA);
            MMS_PROF_BEGIN(st);
            spectrum_fix_pair_kernel<<<grid, 256, 0, st>>>(workA, workC, n_sig, f.M1, f.M2, d.N, d.num, d.m2);
            MMS_LAUNCH_CHECK("spectrum_fix_pair_kernel");
        } else {
            for (int s = 0; s < n_sig; ++s) {
                MMS_PROF_BEGIN(st);
                spectrum_fix_kernel<<<grid_for(f.M2), 256, 0, st>>>(workA + (int64_t)s * f.M1, workC + (int64_t)s * f.M2, f.M1, f.M2, d.N, d.num,
                                                                  d.m2);
                MMS_LAUNCH_CHECK("spectrum_fix_kernel");
            }
        }
        // stage C
        FastHooks hc;
        hc.mul = filtC;
        hc.mul_ready = fs ? fs->ready_c : nullptr;
        rc = fft_fast(workC, f.M2, f.M2, f.nC, 0, hc, st);
        if (rc) return rc;
        FastHooks hj;
        hj.n_keep = d.num;
        rc = fft_fast(workC, f.M2, f.M2, f.nC, 1, hj, st);
        if (rc) return rc;
        MMS_PROF_BEGIN(st);
        if (f.paired_c) {
            dim3 grid(grid_for(d.num), f.nC);
            czt_post_pair_kernel<<<grid, 256, 0, st>>>(workC, f.M2, d.num, d.m2 - 1, n_sig, y);
            MMS_LAUNCH_CHECK("czt_post_pair_kernel");
        } else {
            dim3 grid(grid_for(d.num), n_sig);
            czt_post_real_kernel<<<grid, 256, 0, st>>>(workC, f.M2, d.num, y);
            MMS_LAUNCH_CHECK("czt_post_real_kernel");
        }
        return MMS_OK;
    }
    const int64_t need = (int64_t)sizeof(double2) * d.Mmax * ((int64_t)n_sig + 1);
    if (workspace_bytes < need) {
        set_error("resample_f64: workspace %lld bytes < %lld", (long long)workspace_bytes, (long long)need);
        return MMS_E_WORKSPACE;
    }
    double2* filt = (double2*)workspace;                 // [Mmax]
    double2* work = filt + d.Mmax;                       // [n_sig][Mmax]

    // ---- stage A: the m2 lowest bins of the length-N DFT ------------------------------------------
    // Paired (two real signals per complex transform) when that needs no longer convolution and the stage-C operands of all
    // signals fit in the part of the workspace the halved stage A leaves free; otherwise one transform per signal.
    const int n_pairs = (n_sig + 1) / 2;
    const int64_t J = 2 * d.m2 - 1;
    const bool paired = n_sig >= 2 && option_get("RESAMPLE_PAIRED", 1) == 1 && pow2_at_least(n_in + J - 1) == d.M1 &&
                        (int64_t)n_sig * d.M2 <= (int64_t)(n_sig - n_pairs) * d.Mmax;
    double2* workC = work;                               // where stage C finds its n_sig operands (stride M2)
    if (paired) {
        workC = work + (int64_t)n_pairs * d.Mmax;
        MMS_PROF_BEGIN(st);
        czt_filter_kernel<<<grid_for(d.M1), 256, 0, st>>>(filt, d.M1, d.N, J, d.N, -1);
        MMS_LAUNCH_CHECK("czt_filter_kernel");
        rc = fft_inplace(filt, d.M1, d.M1, 1, 0, st);
        if (rc) return rc;
        {
            dim3 grid(grid_for(d.M1), n_pairs);
            MMS_PROF_BEGIN(st);
            czt_pre_pair_kernel<<<grid, 256, 0, st>>>(x, d.N, n_sig, d.N, d.N - (d.m2 - 1), -1, work, d.M1);     // first bin: -(m2 - 1) mod N
            MMS_LAUNCH_CHECK("czt_pre_pair_kernel");
        }
        rc = fft_inplace(work, d.M1, d.M1, n_pairs, 0, st);
        if (rc) return rc;
        {
            dim3 grid(grid_for(d.M1), n_pairs);
            MMS_PROF_BEGIN(st);
            cmul_kernel<<<grid, 256, 0, st>>>(work, filt, d.M1);
            MMS_LAUNCH_CHECK("cmul_kernel");
        }
        rc = fft_inplace(work, d.M1, d.M1, n_pairs, 1, st);
        if (rc) return rc;
        {
            dim3 grid(grid_for(d.M2), n_pairs);
            MMS_PROF_BEGIN(st);
            spectrum_fix_pair_kernel<<<grid, 256, 0, st>>>(work, workC, n_sig, d.M1, d.M2, d.N, d.num, d.m2);
            MMS_LAUNCH_CHECK("spectrum_fix_pair_kernel");
        }
    } else {
    MMS_PROF_BEGIN(st);
    czt_filter_kernel<<<grid_for(d.M1), 256, 0, st>>>(filt, d.M1, d.N, d.m2, d.N, -1);
    MMS_LAUNCH_CHECK("czt_filter_kernel");
    rc = fft_inplace(filt, d.M1, d.M1, 1, 0, st);
    if (rc) return rc;
    {
        dim3 grid(grid_for(d.M1), n_sig);
        MMS_PROF_BEGIN(st);
        czt_pre_real_kernel<<<grid, 256, 0, st>>>(x, d.N, d.N, -1, work, d.M1);
        MMS_LAUNCH_CHECK("czt_pre_real_kernel");
    }
    rc = fft_inplace(work, d.M1, d.M1, n_sig, 0, st);
    if (rc) return rc;
    {
        dim3 grid(grid_for(d.M1), n_sig);
        MMS_PROF_BEGIN(st);
        cmul_kernel<<<grid, 256, 0, st>>>(work, filt, d.M1);
        MMS_LAUNCH_CHECK("cmul_kernel");
    }
    rc = fft_inplace(work, d.M1, d.M1, n_sig, 1, st);
    if (rc) return rc;

    // ---- spectrum fix-up; re-pack the signals from stride M1 to stride M2 ---------------------------
    // Signal s moves from work + s*M1 to work + s*M2.  With M2 <= M1 going upwards in s never
    // overwrites an unread source; with M2 > M1 going downwards does the same.
    for (int i = 0; i < n_sig; ++i) {
        const int s = d.M2 <= d.M1 ? i : n_sig - 1 - i;
        MMS_PROF_BEGIN(st);
        spectrum_fix_kernel<<<grid_for(d.M2), 256, 0, st>>>(work + (int64_t)s * d.M1, work + (int64_t)s * d.M2, d.M1, d.M2, d.N,
                                                          d.num, d.m2);
        MMS_LAUNCH_CHECK("spectrum_fix_kernel");
    }
    }

    // ---- stage C: one-sided inverse transform of length num ----------------------------------------
    MMS_PROF_BEGIN(st);
    czt_filter_kernel<<<grid_for(d.M2), 256, 0, st>>>(filt, d.M2, d.m2, d.num, d.num, +1);
    MMS_LAUNCH_CHECK("czt_filter_kernel");
    rc = fft_inplace(filt, d.M2, d.M2, 1, 0, st);
    if (rc) return rc;
    rc = fft_inplace(workC, d.M2, d.M2, n_sig, 0, st);
    if (rc) return rc;
    {
        dim3 grid(grid_for(d.M2), n_sig);
        MMS_PROF_BEGIN(st);
        cmul_kernel<<<grid, 256, 0, st>>>(workC, filt, d.M2);
        MMS_LAUNCH_CHECK("cmul_kernel");
    }
    rc = fft_inplace(workC, d.M2, d.M2, n_sig, 1, st);
    if (rc) return rc;
    {
        dim3 grid(grid_for(d.num), n_sig);
        MMS_PROF_BEGIN(st);
        czt_post_real_kernel<<<grid, 256, 0, st>>>(workC, d.M2, d.num, y);
        MMS_LAUNCH_CHECK("czt_post_real_kernel");
    }
    return MMS_OK;
}
