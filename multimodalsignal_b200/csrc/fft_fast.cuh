// Register-resident FFT passes for the chirp-z resampler (resample.cu; reference preprocess.py:70-75 through
// scipy.signal.resample), float64.
//
// A transform of length M = R_0 * R_1 * ... is a sequence of in-place passes.  Pass i works on blocks of Mc = R_i * S
// consecutive elements (S = the product of the later radices): the R_i elements q * S + r of a block form one column,
// which is transformed by an R_i-point DFT and (forward) multiplied by the inter-pass twiddle w_Mc^{r k}.  The previous
// kernel ran the column DFT as log2(R) radix-2 stages through shared memory (8 round trips of 16-byte elements per pass,
// a sincospi per element); here R = N1 * N2 and a thread holds a whole N1- or N2-point sub-transform in REGISTERS:
//     stage 1   thread (column c, q0):  v[q1] = x[N2 q1 + q0],  DFT_N1 over q1,  v[k1] *= w_R^{q0 k1},  -> smem row N2 k1 + q0
//     stage 2   thread (column c, k1):  v[q0] = smem row N2 k1 + q0,  DFT_N2 over q0,  v[k0] *= w_Mc^{r (k1 + N1 k0)},
//                                       -> global row N2 k1 + k0  (which therefore holds frequency k1 + N1 k0)
// so a pass moves every element global -> registers -> shared -> registers -> global: ONE shared-memory round trip, all
// global loads of a thread in flight at once, twiddles as running products (one sincospi per thread for the intra-pass
// factors w_R^{q0 k1}, two for the inter-pass ones) instead of a sincospi per element.  The inverse pass is the exact reverse (conjugate twiddles, DFTs with the
// opposite sign, stages in the opposite order); the frequency order a forward transform leaves (digit-reversed in the
// mixed radix N1, N2 of every pass) is what the inverse expects, and point-wise products are formed in that order.
// N1 in {16, 12, 9, 8}, N2 in {16, 8}: radices 256, 192, 144, 128, 64 -- lengths 2^a, 3 * 2^a and 9 * 2^a.
// Fused into the passes: the chirp pre-multiplication of (pairs of) real signals and the chirp filter (loads of the first
// forward pass; the zero-padded part is not read), the point-wise product with the filter spectrum (stores of the last
// forward pass), pruning of the outputs that are never used (stores of the last inverse pass).
//
// Everything here is __host__ __device__ so that tests/test_fft_fast_host.py can run the same pass bodies on the CPU
// (threads emulated by loops) against a direct DFT.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FF_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define FF_HD inline
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
static inline void sincospi(double x, double* s, double* c) {
    const long double a = 3.14159265358979323846264338327950288L * (long double)x;
    *s = (double)sinl(a); *c = (double)cosl(a);
}
#endif

namespace ff {

FF_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
FF_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
FF_HD double2 cmulf(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
FF_HD double2 cmulc(double2 a, double2 b) { return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y)); }   // a * conj(b)

// w_N^j = (cos, sin)(2 pi j / N), j < N, for the composite register DFTs
// (local constant arrays: indexed by unrolled loop counters they fold into immediates, on the host and in device code)
#define FF_SQ 0.70710678118654752440
#define FF_C16 0.92387953251128675613
#define FF_S16 0.38268343236508977173
#define FF_H3 0.86602540378443864676
#define FF_C91 0.76604444311897803520
#define FF_S91 0.64278760968653932632
#define FF_C92 0.17364817766693034885
#define FF_S92 0.98480775301220805937
#define FF_C94 0.93969262078590838405
#define FF_S94 0.34202014332566873304
template <int N>
FF_HD double tw_cos(int j) {
    if (N == 8) { const double t[8] = {1.0, FF_SQ, 0.0, -FF_SQ, -1.0, -FF_SQ, 0.0, FF_SQ}; return t[j & 7]; }
    if (N == 9) { const double t[9] = {1.0, FF_C91, FF_C92, -0.5, -FF_C94, -FF_C94, -0.5, FF_C92, FF_C91}; return t[j % 9]; }
    if (N == 12) { const double t[12] = {1.0, FF_H3, 0.5, 0.0, -0.5, -FF_H3, -1.0, -FF_H3, -0.5, 0.0, 0.5, FF_H3}; return t[j % 12]; }
    const double t[16] = {1.0, FF_C16, FF_SQ, FF_S16, 0.0, -FF_S16, -FF_SQ, -FF_C16, -1.0, -FF_C16, -FF_SQ, -FF_S16, 0.0, FF_S16, FF_SQ, FF_C16};
    return t[j & 15];
}
template <int N>
FF_HD double tw_sin(int j) {
    if (N == 8) { const double t[8] = {0.0, FF_SQ, 1.0, FF_SQ, 0.0, -FF_SQ, -1.0, -FF_SQ}; return t[j & 7]; }
    if (N == 9) { const double t[9] = {0.0, FF_S91, FF_S92, FF_H3, FF_S94, -FF_S94, -FF_H3, -FF_S92, -FF_S91}; return t[j % 9]; }
    if (N == 12) { const double t[12] = {0.0, 0.5, FF_H3, 1.0, FF_H3, 0.5, 0.0, -0.5, -FF_H3, -1.0, -FF_H3, -0.5}; return t[j % 12]; }
    const double t[16] = {0.0, FF_S16, FF_SQ, FF_C16, 1.0, FF_C16, FF_SQ, FF_S16, 0.0, -FF_S16, -FF_SQ, -FF_C16, -1.0, -FF_C16, -FF_SQ, -FF_S16};
    return t[j & 15];
}

// w_R^{q} = (cos, sin)(2 pi q / R), q < 16, for the five radices (rows: R = 64, 128, 256, 144, 192): the base of the
// intra-pass twiddles of a stage-1 thread, from constant memory instead of a sincospi
#ifdef __CUDACC__
#define FF_CONSTANT static __constant__
#else
#define FF_CONSTANT static const
#endif
FF_CONSTANT double TW_BASE_C[5][16] = {
    {1.0, 0.9951847266721969, 0.9807852804032304, 0.9569403357322088, 0.9238795325112867, 0.881921264348355, 0.8314696123025452, 0.773010453362737, 0.7071067811865476, 0.6343932841636455, 0.5555702330196023, 0.4713967368259978, 0.38268343236508984, 0.29028467725446233, 0.19509032201612833, 0.09801714032956077},
    {1.0, 0.9987954562051724, 0.9951847266721969, 0.989176509964781, 0.9807852804032304, 0.970031253194544, 0.9569403357322088, 0.9415440651830208, 0.9238795325112867, 0.9039892931234433, 0.881921264348355, 0.8577286100002721, 0.8314696123025452, 0.8032075314806449, 0.773010453362737, 0.7409511253549591},
    {1.0, 0.9996988186962042, 0.9987954562051724, 0.9972904566786902, 0.9951847266721969, 0.99247953459871, 0.989176509964781, 0.9852776423889412, 0.9807852804032304, 0.9757021300385286, 0.970031253194544, 0.9637760657954398, 0.9569403357322088, 0.9495281805930367, 0.9415440651830208, 0.932992798834739},
    {1.0, 0.9990482215818578, 0.9961946980917455, 0.9914448613738104, 0.984807753012208, 0.9762960071199334, 0.9659258262890683, 0.9537169507482269, 0.9396926207859084, 0.9238795325112867, 0.9063077870366499, 0.8870108331782217, 0.8660254037844387, 0.8433914458128857, 0.8191520442889918, 0.7933533402912353},
    {1.0, 0.9994645874763657, 0.9978589232386035, 0.9951847266721969, 0.9914448613738104, 0.986643332084879, 0.9807852804032304, 0.9738769792773336, 0.9659258262890683, 0.9569403357322088, 0.9469301294951057, 0.9359059267573258, 0.9238795325112867, 0.9108638249211758, 0.8968727415326884, 0.881921264348355}};
FF_CONSTANT double TW_BASE_S[5][16] = {
    {0.0, 0.0980171403295606, 0.19509032201612825, 0.29028467725446233, 0.3826834323650898, 0.47139673682599764, 0.5555702330196022, 0.6343932841636455, 0.7071067811865475, 0.773010453362737, 0.8314696123025452, 0.8819212643483549, 0.9238795325112867, 0.9569403357322089, 0.9807852804032304, 0.9951847266721968},
    {0.0, 0.049067674327418015, 0.0980171403295606, 0.14673047445536175, 0.19509032201612825, 0.24298017990326387, 0.29028467725446233, 0.33688985339222005, 0.3826834323650898, 0.4275550934302821, 0.47139673682599764, 0.5141027441932217, 0.5555702330196022, 0.5956993044924334, 0.6343932841636455, 0.6715589548470183},
    {0.0, 0.024541228522912288, 0.049067674327418015, 0.07356456359966743, 0.0980171403295606, 0.1224106751992162, 0.14673047445536175, 0.17096188876030122, 0.19509032201612825, 0.2191012401568698, 0.24298017990326387, 0.26671275747489837, 0.29028467725446233, 0.3136817403988915, 0.33688985339222005, 0.3598950365349881},
    {0.0, 0.043619387365336, 0.08715574274765817, 0.13052619222005157, 0.17364817766693033, 0.21643961393810288, 0.25881904510252074, 0.3007057995042731, 0.3420201433256687, 0.3826834323650898, 0.42261826174069944, 0.46174861323503386, 0.49999999999999994, 0.5372996083468239, 0.573576436351046, 0.6087614290087205},
    {0.0, 0.03271908282177614, 0.06540312923014306, 0.0980171403295606, 0.13052619222005157, 0.16289547339458874, 0.19509032201612825, 0.2270762630343732, 0.25881904510252074, 0.29028467725446233, 0.3214394653031616, 0.3522500479212335, 0.3826834323650898, 0.4127070298043947, 0.44228869021900125, 0.4713967368259976}};
template <int R>
FF_HD double2 tw_base(int q0) {      // e^{-2 pi i q0 / R}
#if defined(__CUDA_ARCH__) || !defined(__CUDACC__)
    constexpr int row = R == 64 ? 0 : R == 128 ? 1 : R == 256 ? 2 : R == 144 ? 3 : 4;
    static_assert(R == 64 || R == 128 || R == 256 || R == 144 || R == 192, "radix without a table");
    return make_double2(TW_BASE_C[row][q0], -TW_BASE_S[row][q0]);
#else
    double sn, cs;
    sincospi(2.0 * (double)q0 / (double)R, &sn, &cs);
    return make_double2(cs, -sn);
#endif
}
// b^k for 0 <= k < 16 by squaring (three squarings, at most three products)
FF_HD double2 cpow16(double2 b, int k, double2* b8_out) {
    const double2 b2 = make_double2(fma(b.x, b.x, -b.y * b.y), 2.0 * b.x * b.y);
    const double2 b4 = make_double2(fma(b2.x, b2.x, -b2.y * b2.y), 2.0 * b2.x * b2.y);
    const double2 b8 = make_double2(fma(b4.x, b4.x, -b4.y * b4.y), 2.0 * b4.x * b4.y);
    double2 w = (k & 1) ? b : make_double2(1.0, 0.0);
    if (k & 2) w = make_double2(fma(w.x, b2.x, -w.y * b2.y), fma(w.x, b2.y, w.y * b2.x));
    if (k & 4) w = make_double2(fma(w.x, b4.x, -w.y * b4.y), fma(w.x, b4.y, w.y * b4.x));
    if (k & 8) w = make_double2(fma(w.x, b8.x, -w.y * b8.y), fma(w.x, b8.y, w.y * b8.x));
    *b8_out = b8;
    return w;
}

// v *= (c + i SIGN s): SIGN = -1 forward (e^{-2 pi i jk/N}), +1 inverse; the trivial factors cost nothing once unrolled
template <int SIGN>
FF_HD double2 mul_const(double2 v, double c, double s) {
    if (c == 1.0 && s == 0.0) return v;
    if (c == -1.0 && s == 0.0) return make_double2(-v.x, -v.y);
    const double ss = SIGN > 0 ? s : -s;
    if (c == 0.0 && ss == 1.0) return make_double2(-v.y, v.x);
    if (c == 0.0 && ss == -1.0) return make_double2(v.y, -v.x);
    return make_double2(fma(v.x, c, -v.y * ss), fma(v.x, ss, v.y * c));
}

// natural order in, natural order out
template <int N, int SIGN> struct Dft;
template <int SIGN> struct Dft<2, SIGN> {
    static FF_HD void run(double2* v) {
        const double2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <int SIGN> struct Dft<3, SIGN> {
    static FF_HD void run(double2* v) {
        const double2 s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
        const double2 m = make_double2(fma(-0.5, s.x, v[0].x), fma(-0.5, s.y, v[0].y));
        const double h = SIGN > 0 ? FF_H3 : -FF_H3;           // X1 = m + i h d (inverse), m - i |h| d (forward)
        v[0] = cadd(v[0], s);
        v[1] = make_double2(fma(-h, d.y, m.x), fma(h, d.x, m.y));
        v[2] = make_double2(fma(h, d.y, m.x), fma(-h, d.x, m.y));
    }
};
template <int SIGN> struct Dft<4, SIGN> {
    static FF_HD void run(double2* v) {
        const double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        const double2 t3 = SIGN > 0 ? make_double2(-d.y, d.x) : make_double2(d.y, -d.x);       // (+-i) d
        v[0] = cadd(t0, t2);
        v[2] = csub(t0, t2);
        v[1] = cadd(t1, t3);
        v[3] = csub(t1, t3);
    }
};
// N = A * B:  n = B n1 + n0,  k = k1 + A k0:  X[k1 + A k0] = sum_n0 w_B^{n0 k0} ( w_N^{n0 k1} sum_n1 x[B n1 + n0] w_A^{n1 k1} )
template <int A, int B, int SIGN>
FF_HD void dft_comp(double2* v) {
    constexpr int N = A * B;
    double2 t[N];
#pragma unroll
    for (int n0 = 0; n0 < B; ++n0) {
        double2 u[A];
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1) u[n1] = v[B * n1 + n0];
        Dft<A, SIGN>::run(u);
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) t[n0 * A + k1] = mul_const<SIGN>(u[k1], tw_cos<N>(n0 * k1), tw_sin<N>(n0 * k1));
    }
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) {
        double2 u[B];
#pragma unroll
        for (int n0 = 0; n0 < B; ++n0) u[n0] = t[n0 * A + k1];
        Dft<B, SIGN>::run(u);
#pragma unroll
        for (int k0 = 0; k0 < B; ++k0) v[k1 + A * k0] = u[k0];
    }
}
template <int SIGN> struct Dft<8, SIGN> { static FF_HD void run(double2* v) { dft_comp<4, 2, SIGN>(v); } };
template <int SIGN> struct Dft<9, SIGN> { static FF_HD void run(double2* v) { dft_comp<3, 3, SIGN>(v); } };
template <int SIGN> struct Dft<12, SIGN> { static FF_HD void run(double2* v) { dft_comp<4, 3, SIGN>(v); } };
template <int SIGN> struct Dft<16, SIGN> { static FF_HD void run(double2* v) { dft_comp<4, 4, SIGN>(v); } };

// e^{sign * 2 pi i e / Mc} for an integer exponent already reduced mod Mc
FF_HD double2 cis_frac(uint64_t e, int64_t Mc, double sign) {
    double s, c;
    sincospi(2.0 * (double)e / (double)Mc, &s, &c);
    return make_double2(c, sign * s);
}
// e^{sign * pi * i * ((n * b) mod 2P) / P} for integers with n * b < 2^53: the product is exact in float64, the quotient by
// 2P is estimated with a reciprocal, and one fused multiply-add leaves the exact remainder (corrected by at most one period) --
// a dozen float64 instructions where a 64-bit integer modulo by a run-time divisor costs over a hundred.
FF_HD double2 chirp_prod(int64_t n, int64_t b, int64_t P, int sign) {
    const double two_p = (double)(2 * P), inv_p = 1.0 / (double)P;
    const double prod = (double)n * (double)b;
    const double q = floor(prod * (0.5 * inv_p));
    double r = fma(-q, two_p, prod);
    if (r < 0.0) r += two_p;
    if (r >= two_p) r -= two_p;
    double s, c;
    sincospi(r * inv_p, &s, &c);
    return make_double2(c, sign > 0 ? s : -s);
}
// v mod m for non-negative integers held exactly in float64 (v < 2^53), inv_m = 1 / m
FF_HD double mod_exact(double v, double m, double inv_m) {
    const double q = floor(v * inv_m);
    double r = fma(-q, m, v);
    if (r < 0.0) r += m;
    if (r >= m) r -= m;
    return r;
}
FF_HD double2 cis_pi(double x, int sign) {
    double s, c;
    sincospi(x, &s, &c);
    return make_double2(c, sign > 0 ? s : -s);
}
// the chirp e^{sign pi i n^2 / P} (0 <= n <= 2^26) and the chirp times the linear phase e^{sign 2 pi i n k / P} (0 <= n, k < P <= 2^26)
FF_HD double2 chirp_at(int64_t n, int64_t P, int sign) { return chirp_prod(n, n, P, sign); }
FF_HD double2 chirp_shift_at(int64_t n, int64_t k, int64_t P, int sign) {
    int64_t b = n + 2 * k;                                  // < 3P: one subtraction brings it below 2P, so n * b < 2 P^2 <= 2^53
    if (b >= 2 * P) b -= 2 * P;
    return chirp_prod(n, b, P, sign);
}

// 32-byte global accesses (two complex doubles of one thread): Blackwell's 256-bit ld / st.  The contiguous-tile side of a pass
// has every thread own 16 consecutive elements; 16-byte accesses there would touch half a sector per lane and instruction.
FF_HD void load2(const double2* p, double2* a, double2* b) {
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a->x), "=d"(a->y), "=d"(b->x), "=d"(b->y) : "l"(p) : "memory");
#else
    *a = p[0]; *b = p[1];
#endif
}
FF_HD void store2(double2* p, double2 a, double2 b) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
#else
    p[0] = a; p[1] = b;
#endif
}

enum { LD_PLAIN = 0, LD_PAIR = 1, LD_FILTER = 2 };

struct PassArgs {
    double2* a;             // [n_sig][sig_stride] complex work buffer, transformed in place
    int64_t sig_stride;
    int64_t M;              // transform length
    int64_t Mc;             // block size of this pass (= R * S)
    int contig;             // S == 1 (the last forward / first inverse pass): a tile is TC consecutive blocks
    int64_t ncols;          // columns in all: M / R (strided: (block, r) pairs; contiguous: blocks)
    // loads of the first forward pass
    int load_op;
    const double* x;        // LD_PAIR: real signals [n_x][n_in]; paired: transform p = blockIdx.y takes rows 2p (real part) and
                            // 2p + 1 (imaginary part, zero if absent); not paired: row p alone
    int paired;
    int64_t n_in;           // LD_PAIR: signal length; LD_FILTER: the filter covers lags (-n_in, n_out)
    int n_x;
    int64_t P, k0;          // chirp period and first bin
    int sign;
    int64_t n_out;
    // stores
    const double2* mul;     // forward: multiply by mul[i] (filter spectrum in the same order) before the store
    int64_t n_keep;         // inverse: elements i >= n_keep of a signal are not stored
};

// position of tile row `row`, column `c`: strided tiles are [row][c] (lanes run over c), contiguous tiles are [c][row + row / N2]
// (lanes run over q0 in stage 1 and over k1 in stage 2: the pad makes the stride-N2 accesses of stage 2 conflict-free)
template <int R, int N2, int TC>
FF_HD int tile_pos(int row, int c, int contig) {
    return contig ? c * (R + R / N2) + row + row / N2 : row * TC + c;
}
// + 2 TC entries behind the tile: b_c = w_Mc^{r0 + c} and its N1-th power (each from its own exactly reduced sincospi), the bases of
// column c's inter-pass twiddles w_Mc^{r (k1 + N1 k0)} = b^{k1} (b^{N1})^{k0}: written by TC threads (stage 1 of the forward,
// col_bases() + a barrier in front of the inverse), so the other threads need neither a sincospi nor a modulo for them
template <int R, int N2, int TC>
constexpr int tile_elems() { return TC * (R + R / N2) + 2 * TC; }
template <int R, int N2, int TC>
constexpr int colbase_offset() { return TC * (R + R / N2); }

template <int N1, int N2, int TC>
FF_HD double2 pass_load(const PassArgs& p, int sig, const double2* base, int64_t idx) {
    if (p.load_op == LD_PLAIN) return base[idx];
    if (p.load_op == LD_PAIR) {
        if (idx >= p.n_in) return make_double2(0.0, 0.0);
        const double re = p.x[(int64_t)(p.paired ? 2 * sig : sig) * p.n_in + idx];
        const double im = p.paired && 2 * sig + 1 < p.n_x ? p.x[(int64_t)(2 * sig + 1) * p.n_in + idx] : 0.0;
        return cmulf(make_double2(re, im), chirp_shift_at(idx, p.k0, p.P, p.sign));
    }
    // LD_FILTER: b[i mod M] = conj(chirp(i)) for lags i in (-n_in, n_out)
    if (idx < p.n_out) return chirp_at(idx, p.P, -p.sign);
    if (p.M - idx < p.n_in) return chirp_at(p.M - idx, p.P, -p.sign);
    return make_double2(0.0, 0.0);
}

// bases of column c's inter-pass twiddles (sign -1 forward, +1 inverse)
template <int N1, int N2, int TC>
FF_HD void write_col_bases(const PassArgs& p, int64_t r0, int c, int sign, double2* tile) {
    constexpr int R = N1 * N2;
    const double r = (double)(r0 + c), mc = (double)p.Mc, inv_mc = 1.0 / mc;
    tile[colbase_offset<R, N2, TC>() + c] = cis_pi(2.0 * r * inv_mc, sign);                                  // r < Mc / R: no reduction needed
    tile[colbase_offset<R, N2, TC>() + TC + c] = cis_pi(2.0 * mod_exact(r * (double)N1, mc, inv_mc) * inv_mc, sign);
}

// ---- forward pass -----------------------------------------------------------------------------------------------------
// One CTA = one tile of TC columns; `tile` holds tile_elems() values.  tile_x = blockIdx.x, sig = blockIdx.y; the two stage
// bodies are called once per thread with a block barrier in between.
template <int N1, int N2, int TC>
FF_HD void tile_origin(const PassArgs& p, int tile_x, int64_t* origin, int64_t* S, int64_t* r0, int* tc_eff) {
    constexpr int R = N1 * N2;
    *S = p.Mc / R;
    const int64_t left = p.ncols - (int64_t)tile_x * TC;
    *tc_eff = left < TC ? (int)left : TC;
    if (p.contig) { *origin = (int64_t)tile_x * TC * R; *r0 = 0; }
    else {
        const int64_t g0 = (int64_t)tile_x * TC, blk0 = g0 / *S;
        *r0 = g0 - blk0 * *S;
        *origin = blk0 * p.Mc + *r0;
    }
}

template <int N1, int N2, int TC, int NT>
FF_HD void fwd_stage1(const PassArgs& p, int tid, int tile_x, int sig, double2* tile) {
    constexpr int R = N1 * N2;
    int64_t origin, S, r0;
    int tc_eff;
    tile_origin<N1, N2, TC>(p, tile_x, &origin, &S, &r0, &tc_eff);
    const double2* base = p.a + (int64_t)sig * p.sig_stride;
    for (int it = tid; it < TC * N2; it += NT) {
        const int c = p.contig ? it / N2 : it % TC, q0 = p.contig ? it % N2 : it / TC;
        if (c >= tc_eff) continue;
        double2 v[N1];
        if (p.load_op == LD_PAIR && !p.contig && p.P <= ((int64_t)1 << 25)) {
            // chirp of the N1 elements of this thread (stride D = N2 S) from ONE exact evaluation and a second-order recurrence:
            //   phi(n + D) - phi(n) = pi D (2n + D + 2 k0) / P,   and that difference grows by 2 pi D^2 / P per step
            // (three sincospi per thread instead of one per element; every product below stays under 2^53)
            const int64_t n0 = origin + (int64_t)q0 * S + c, D = (int64_t)N2 * S;
#pragma unroll
            for (int q1 = 0; q1 < N1; ++q1) v[q1] = make_double2(0.0, 0.0);
            if (n0 < p.n_in) {
                const double two_p = (double)(2 * p.P), inv_p = 1.0 / (double)p.P, inv_2p = 0.5 * inv_p;
                const double dm = mod_exact((double)D, two_p, inv_2p);
                const double bm = mod_exact((double)(2 * n0 + D + 2 * p.k0), two_p, inv_2p);
                double2 ch = chirp_shift_at(n0, p.k0, p.P, p.sign);
                double2 ratio = cis_pi(mod_exact(bm * dm, two_p, inv_2p) * inv_p, p.sign);
                const double2 grow = cis_pi(mod_exact(2.0 * dm * dm, two_p, inv_2p) * inv_p, p.sign);
                const double* xr = p.x + (int64_t)(p.paired ? 2 * sig : sig) * p.n_in;
                const double* xi = p.paired && 2 * sig + 1 < p.n_x ? p.x + (int64_t)(2 * sig + 1) * p.n_in : nullptr;
#pragma unroll
                for (int q1 = 0; q1 < N1; ++q1) {
                    const int64_t idx = n0 + (int64_t)q1 * D;
                    if (idx < p.n_in) v[q1] = cmulf(make_double2(xr[idx], xi ? xi[idx] : 0.0), ch);
                    ch = cmulf(ch, ratio);
                    ratio = cmulf(ratio, grow);
                }
            }
        } else {
#pragma unroll
            for (int q1 = 0; q1 < N1; ++q1) {
                const int q = N2 * q1 + q0;
                const int64_t idx = p.contig ? origin + (int64_t)c * R + q : origin + (int64_t)q * S + c;
                v[q1] = pass_load<N1, N2, TC>(p, sig, base, idx);
            }
        }
        Dft<N1, -1>::run(v);
        if (!p.contig && q0 == 0) write_col_bases<N1, N2, TC>(p, r0, c, -1, tile);
        const double2 wq = tw_base<R>(q0);                        // w_R^{q0 k1} as a running product over k1
        double2 w = wq;
        tile[tile_pos<R, N2, TC>(q0, c, p.contig)] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < N1; ++k1) {
            tile[tile_pos<R, N2, TC>(N2 * k1 + q0, c, p.contig)] = cmulf(v[k1], w);
            w = cmulf(w, wq);
        }
    }
}

template <int N1, int N2, int TC, int NT>
FF_HD void fwd_stage2(const PassArgs& p, int tid, int tile_x, int sig, const double2* tile) {
    constexpr int R = N1 * N2;
    int64_t origin, S, r0;
    int tc_eff;
    tile_origin<N1, N2, TC>(p, tile_x, &origin, &S, &r0, &tc_eff);
    double2* base = p.a + (int64_t)sig * p.sig_stride;
    for (int it = tid; it < TC * N1; it += NT) {
        const int c = p.contig ? it / N1 : it % TC, k1 = p.contig ? it % N1 : it / TC;
        if (c >= tc_eff) continue;
        double2 v[N2];
#pragma unroll
        for (int q0 = 0; q0 < N2; ++q0) v[q0] = tile[tile_pos<R, N2, TC>(N2 * k1 + q0, c, p.contig)];
        Dft<N2, -1>::run(v);
        if (S > 1) {        // inter-pass twiddle w_Mc^{r (k1 + N1 k0)} = b^{k1} (b^{N1})^{k0}, b = w_Mc^r from stage 1 (no sincospi, no modulo)
            const double2 b = tile[colbase_offset<R, N2, TC>() + c], step = tile[colbase_offset<R, N2, TC>() + TC + c];
            double2 b8;
            double2 w = cpow16(b, k1, &b8);
#pragma unroll
            for (int k0 = 0; k0 < N2; ++k0) {
                v[k0] = cmulf(v[k0], w);
                w = cmulf(w, step);
            }
        }
        if (p.contig) {         // 16 consecutive elements per thread: 32-byte accesses
            const int64_t i0 = origin + (int64_t)c * R + N2 * k1;
#pragma unroll
            for (int k0 = 0; k0 < N2; k0 += 2) {
                double2 o0 = v[k0], o1 = v[k0 + 1];
                if (p.mul) {
                    double2 m0, m1;
                    load2(p.mul + i0 + k0, &m0, &m1);
                    o0 = cmulf(o0, m0);
                    o1 = cmulf(o1, m1);
                }
                store2(base + i0 + k0, o0, o1);
            }
        } else {
#pragma unroll
            for (int k0 = 0; k0 < N2; ++k0) {
                const int64_t idx = origin + (int64_t)(N2 * k1 + k0) * S + c;
                double2 o = v[k0];
                if (p.mul) o = cmulf(o, p.mul[idx]);
                base[idx] = o;
            }
        }
    }
}

// ---- inverse pass (exact reverse of the forward pass, unnormalised: a forward + inverse transform multiplies by M) ------
// inv_stage0 (TC threads) + barrier, inv_stage2, barrier, inv_stage1
template <int N1, int N2, int TC, int NT>
FF_HD void inv_stage0(const PassArgs& p, int tid, int tile_x, double2* tile) {
    int64_t origin, S, r0;
    int tc_eff;
    tile_origin<N1, N2, TC>(p, tile_x, &origin, &S, &r0, &tc_eff);
    if (!p.contig && tid < tc_eff) write_col_bases<N1, N2, TC>(p, r0, tid, +1, tile);
}

template <int N1, int N2, int TC, int NT>
FF_HD void inv_stage2(const PassArgs& p, int tid, int tile_x, int sig, double2* tile) {
    constexpr int R = N1 * N2;
    int64_t origin, S, r0;
    int tc_eff;
    tile_origin<N1, N2, TC>(p, tile_x, &origin, &S, &r0, &tc_eff);
    const double2* base = p.a + (int64_t)sig * p.sig_stride;
    for (int it = tid; it < TC * N1; it += NT) {
        const int c = p.contig ? it / N1 : it % TC, k1 = p.contig ? it % N1 : it / TC;
        if (c >= tc_eff) continue;
        double2 v[N2];
        if (p.contig) {
            const int64_t i0 = origin + (int64_t)c * R + N2 * k1;
#pragma unroll
            for (int k0 = 0; k0 < N2; k0 += 2) load2(base + i0 + k0, &v[k0], &v[k0 + 1]);
        } else {
#pragma unroll
            for (int k0 = 0; k0 < N2; ++k0) v[k0] = base[origin + (int64_t)(N2 * k1 + k0) * S + c];
        }
        if (S > 1) {        // conjugate inter-pass twiddles from the column bases (inv_stage0)
            const double2 b = tile[colbase_offset<R, N2, TC>() + c], step = tile[colbase_offset<R, N2, TC>() + TC + c];
            double2 b8;
            double2 w = cpow16(b, k1, &b8);
#pragma unroll
            for (int k0 = 0; k0 < N2; ++k0) {
                v[k0] = cmulf(v[k0], w);
                w = cmulf(w, step);
            }
        }
        Dft<N2, +1>::run(v);
        const double2 wk = tw_base<R>(k1);                        // conj(w_R^{q0 k1}) as a running product over q0
        double2 wi = wk;
        tile[tile_pos<R, N2, TC>(N2 * k1, c, p.contig)] = v[0];
#pragma unroll
        for (int q0 = 1; q0 < N2; ++q0) {
            tile[tile_pos<R, N2, TC>(N2 * k1 + q0, c, p.contig)] = cmulc(v[q0], wi);
            wi = cmulf(wi, wk);
        }
    }
}

template <int N1, int N2, int TC, int NT>
FF_HD void inv_stage1(const PassArgs& p, int tid, int tile_x, int sig, const double2* tile) {
    constexpr int R = N1 * N2;
    int64_t origin, S, r0;
    int tc_eff;
    tile_origin<N1, N2, TC>(p, tile_x, &origin, &S, &r0, &tc_eff);
    double2* base = p.a + (int64_t)sig * p.sig_stride;
    for (int it = tid; it < TC * N2; it += NT) {
        const int c = p.contig ? it / N2 : it % TC, q0 = p.contig ? it % N2 : it / TC;
        if (c >= tc_eff) continue;
        // pruned outputs: the smallest index this item stores is that of row q0 (q1 = 0)
        if ((p.contig ? origin + (int64_t)c * R + q0 : origin + (int64_t)q0 * S + c) >= p.n_keep) continue;
        double2 v[N1];
#pragma unroll
        for (int k1 = 0; k1 < N1; ++k1) v[k1] = tile[tile_pos<R, N2, TC>(N2 * k1 + q0, c, p.contig)];
        Dft<N1, +1>::run(v);
#pragma unroll
        for (int q1 = 0; q1 < N1; ++q1) {
            const int q = N2 * q1 + q0;
            const int64_t idx = p.contig ? origin + (int64_t)c * R + q : origin + (int64_t)q * S + c;
            if (idx < p.n_keep) base[idx] = v[q1];
        }
    }
}

// ---- plan ------------------------------------------------------------------------------------------------------------
// M = f * 2^b with f in {1, 3, 9}: an optional first pass of radix 192 = 12 x 16 or 144 = 9 x 16, then power-of-two passes of
// 6..8 bits (64 = 8 x 8, 128 = 16 x 8, 256 = 16 x 16).  ok == false: not such a length (the caller keeps the generic kernels).
struct FastPass { int n1, n2, tc; int64_t mc; };
struct FastPlan { int n; FastPass p[8]; bool ok; };

inline FastPlan make_fast_plan(int64_t M) {
    FastPlan pl;
    pl.n = 0; pl.ok = false;
    int64_t rest = M, mc = M;
    if (M >= 144 && M % 144 == 0 && ((M / 144) & (M / 144 - 1)) == 0) {
        pl.p[pl.n++] = {9, 16, 16, mc}; rest = M / 144; mc = rest;
    } else if (M >= 192 && M % 192 == 0 && ((M / 192) & (M / 192 - 1)) == 0) {
        pl.p[pl.n++] = {12, 16, 16, mc}; rest = M / 192; mc = rest;
    }
    if (rest < 1 || (rest & (rest - 1)) != 0) return pl;
    int bits = 0;
    while (((int64_t)1 << bits) < rest) ++bits;
    if (bits == 0) { pl.ok = pl.n > 0; return pl; }
    const int np = (bits + 7) / 8;
    if (bits < 6 * np || pl.n + np > 8) return pl;
    int left = bits;
    for (int i = 0; i < np; ++i) {
        const int b = (left + (np - i) - 1) / (np - i);       // spread the bits evenly, larger radices first
        left -= b;
        if (b == 8) pl.p[pl.n++] = {16, 16, 16, mc};
        else if (b == 7) pl.p[pl.n++] = {16, 8, 32, mc};
        else pl.p[pl.n++] = {8, 8, 64, mc};
        mc >>= b;
    }
    pl.ok = true;
    return pl;
}

// smallest length >= need among 2^a, 9 * 2^(a-3) and 3 * 2^(a-1) that make_fast_plan accepts; 0 if none below 2^30
inline int64_t fast_length_at_least(int64_t need) {
    int64_t best = 0;
    for (int a = 6; a <= 30; ++a) {
        const int64_t cand[3] = {(int64_t)1 << a, a >= 7 ? (int64_t)9 << (a - 3) : 0, a >= 7 ? (int64_t)3 << (a - 1) : 0};
        for (int j = 0; j < 3; ++j)
            if (cand[j] >= need && (best == 0 || cand[j] < best) && make_fast_plan(cand[j]).ok) best = cand[j];
        if (best && ((int64_t)1 << a) >= best) break;
    }
    return best;
}

}  // namespace ff
