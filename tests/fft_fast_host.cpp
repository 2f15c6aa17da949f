// CPU harness for multimodalsignal_b200/csrc/fft_fast.cuh (test infrastructure; built and run by tests/test_fft_fast_host.py).
// The pass bodies are the ones the CUDA kernels call; here a CTA is a pair of loops over `tid` with the block barrier between
// them.  Checks, per length M:
//   (1) DC: position 0 of the forward transform holds sum(x);  (2) inverse(forward(x)) == M x;
//   (3) the convolution theorem in the transform's own frequency order -- the property the chirp-z resampler relies on:
//       inverse(forward(a) .* forward(b)) == M * (a circularly convolved with b), b sparse so the direct sum is cheap;
//   (4) the fused hooks: chirp pair loads == plain loads of a pre-multiplied buffer, filter loads == a generated filter,
//       multiplied stores == a separate product, pruned stores leave the tail untouched.
// Prints one line per length and exits non-zero on a failure.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../multimodalsignal_b200/csrc/fft_fast.cuh"

using namespace ff;

template <int N1, int N2, int TC>
static void run_pass(const PassArgs& p, int n_sig, bool inverse) {
    constexpr int R = N1 * N2, NT = 256;
    const int64_t tiles = (p.ncols + TC - 1) / TC;
    std::vector<double2> tile(tile_elems<R, N2, TC>());
    for (int sig = 0; sig < n_sig; ++sig)
        for (int64_t tx = 0; tx < tiles; ++tx) {
            if (!inverse) {
                for (int tid = 0; tid < NT; ++tid) fwd_stage1<N1, N2, TC, NT>(p, tid, (int)tx, sig, tile.data());
                for (int tid = 0; tid < NT; ++tid) fwd_stage2<N1, N2, TC, NT>(p, tid, (int)tx, sig, tile.data());
            } else {
                for (int tid = 0; tid < NT; ++tid) inv_stage0<N1, N2, TC, NT>(p, tid, (int)tx, tile.data());
                for (int tid = 0; tid < NT; ++tid) inv_stage2<N1, N2, TC, NT>(p, tid, (int)tx, sig, tile.data());
                for (int tid = 0; tid < NT; ++tid) inv_stage1<N1, N2, TC, NT>(p, tid, (int)tx, sig, tile.data());
            }
        }
}

static void dispatch(const FastPass& fp, const PassArgs& p, int n_sig, bool inverse) {
    if (fp.n1 == 16 && fp.n2 == 16) run_pass<16, 16, 16>(p, n_sig, inverse);
    else if (fp.n1 == 16 && fp.n2 == 8) run_pass<16, 8, 32>(p, n_sig, inverse);
    else if (fp.n1 == 8 && fp.n2 == 8) run_pass<8, 8, 64>(p, n_sig, inverse);
    else if (fp.n1 == 9 && fp.n2 == 16) run_pass<9, 16, 16>(p, n_sig, inverse);
    else if (fp.n1 == 12 && fp.n2 == 16) run_pass<12, 16, 16>(p, n_sig, inverse);
    else { fprintf(stderr, "no instantiation for %d x %d\n", fp.n1, fp.n2); exit(2); }
}

// hooks: load (first forward pass), mul (last forward pass), n_keep (last inverse pass = pass 0)
static void transform(double2* a, int64_t stride, int64_t M, int n_sig, bool inverse, const PassArgs* hooks) {
    const FastPlan pl = make_fast_plan(M);
    if (!pl.ok) { fprintf(stderr, "no plan for %lld\n", (long long)M); exit(2); }
    for (int ii = 0; ii < pl.n; ++ii) {
        const int i = inverse ? pl.n - 1 - ii : ii;
        PassArgs p;
        memset(&p, 0, sizeof(p));
        p.a = a; p.sig_stride = stride; p.M = M; p.Mc = pl.p[i].mc; p.contig = i == pl.n - 1;
        p.ncols = M / (pl.p[i].n1 * pl.p[i].n2);
        p.n_keep = M;
        if (hooks) {
            if (!inverse && i == 0) {
                p.load_op = hooks->load_op; p.x = hooks->x; p.paired = hooks->paired; p.n_in = hooks->n_in; p.n_x = hooks->n_x; p.P = hooks->P; p.k0 = hooks->k0;
                p.sign = hooks->sign; p.n_out = hooks->n_out;
            }
            if (!inverse && i == pl.n - 1) p.mul = hooks->mul;
            if (inverse && i == 0 && hooks->n_keep > 0) p.n_keep = hooks->n_keep;
        }
        dispatch(pl.p[i], p, n_sig, inverse);
    }
}

static double frand() { return (double)rand() / RAND_MAX - 0.5; }
static double maxdiff(const double2* a, const double2* b, int64_t n) {
    double m = 0;
    for (int64_t i = 0; i < n; ++i) {
        m = fmax(m, fabs(a[i].x - b[i].x));
        m = fmax(m, fabs(a[i].y - b[i].y));
    }
    return m;
}

static int check_length(int64_t M) {
    const int n_sig = 2;
    std::vector<double2> x(n_sig * M), a, b(M), fb;
    srand((unsigned)M);
    for (auto& v : x) v = make_double2(frand(), frand());
    int bad = 0;
    const double tol = 1e-12 * (double)M;

    // (1) + (2)
    a = x;
    transform(a.data(), M, M, n_sig, false, nullptr);
    for (int s = 0; s < n_sig; ++s) {
        double sx = 0, sy = 0;
        for (int64_t i = 0; i < M; ++i) { sx += x[s * M + i].x; sy += x[s * M + i].y; }
        if (fabs(a[s * M].x - sx) > tol || fabs(a[s * M].y - sy) > tol) { printf("  DC mismatch (signal %d)\n", s); ++bad; }
    }
    transform(a.data(), M, M, n_sig, true, nullptr);
    for (auto& v : a) { v.x /= (double)M; v.y /= (double)M; }
    const double e_rt = maxdiff(a.data(), x.data(), n_sig * M);
    if (e_rt > 1e-13 * 64) { printf("  round trip error %.3e\n", e_rt); ++bad; }

    // (3) sparse b with 5 taps
    const int64_t taps[5] = {0, 1, M / 3, M / 2 + 1, M - 1};
    double2 tv[5];
    for (auto& v : b) v = make_double2(0, 0);
    for (int j = 0; j < 5; ++j) { tv[j] = make_double2(frand(), frand()); b[taps[j]] = tv[j]; }
    fb = b;
    transform(fb.data(), M, M, 1, false, nullptr);
    a = x;
    {
        PassArgs hooks;
        memset(&hooks, 0, sizeof(hooks));
        hooks.mul = fb.data();                              // the product rides on the last forward pass
        transform(a.data(), M, M, n_sig, false, &hooks);
    }
    transform(a.data(), M, M, n_sig, true, nullptr);
    double e_cv = 0;
    for (int s = 0; s < n_sig; ++s)
        for (int64_t i = 0; i < M; i += (M > 4096 ? 7 : 1)) {
            double2 acc = make_double2(0, 0);
            for (int j = 0; j < 5; ++j) acc = cadd(acc, cmulf(tv[j], x[s * M + ((i - taps[j]) % M + M) % M]));
            e_cv = fmax(e_cv, fabs(a[s * M + i].x / (double)M - acc.x));
            e_cv = fmax(e_cv, fabs(a[s * M + i].y / (double)M - acc.y));
        }
    if (e_cv > 1e-13 * 64) { printf("  convolution error %.3e\n", e_cv); ++bad; }

    // (4) hooks: paired chirp loads, filter loads, pruned stores
    const int64_t n_in = M / 2 + 3, P = n_in, k0 = n_in - 5, n_out = M / 8 + 1;
    std::vector<double> xr(3 * n_in);
    for (auto& v : xr) v = frand();
    std::vector<double2> ref(2 * M), got(2 * M);
    for (int p = 0; p < 2; ++p)
        for (int64_t n = 0; n < M; ++n) {
            double2 v = make_double2(0, 0);
            if (n < n_in) {
                const double re = xr[(2 * p) * n_in + n], im = 2 * p + 1 < 3 ? xr[(2 * p + 1) * n_in + n] : 0.0;
                v = cmulf(make_double2(re, im), chirp_shift_at(n, k0, P, -1));
            }
            ref[p * M + n] = v;
        }
    transform(ref.data(), M, M, 2, false, nullptr);
    for (auto& v : got) v = make_double2(123.0, 456.0);     // the first pass must not read the buffer
    {
        PassArgs hooks;
        memset(&hooks, 0, sizeof(hooks));
        hooks.load_op = LD_PAIR; hooks.paired = 1; hooks.x = xr.data(); hooks.n_in = n_in; hooks.n_x = 3; hooks.P = P; hooks.k0 = k0; hooks.sign = -1;
        transform(got.data(), M, M, 2, false, &hooks);
    }
    const double e_ld = maxdiff(ref.data(), got.data(), 2 * M);
    if (e_ld > 1e-13 * 64) { printf("  paired chirp loads differ by %.3e\n", e_ld); ++bad; }       // chirp by recurrence: not bit-equal

    for (int64_t i = 0; i < M; ++i) {
        double2 v = make_double2(0, 0);
        if (i < n_out) v = chirp_at(i, P, +1);
        else if (M - i < n_in) v = chirp_at(M - i, P, +1);
        ref[i] = v;
        got[i] = make_double2(9.0, 9.0);
    }
    if (n_out + n_in - 1 <= M) {
        transform(ref.data(), M, M, 1, false, nullptr);
        PassArgs hooks;
        memset(&hooks, 0, sizeof(hooks));
        hooks.load_op = LD_FILTER; hooks.n_in = n_in; hooks.n_out = n_out; hooks.P = P; hooks.sign = -1;
        transform(got.data(), M, M, 1, false, &hooks);
        const double e_f = maxdiff(ref.data(), got.data(), M);
        if (e_f > 0) { printf("  filter loads differ by %.3e\n", e_f); ++bad; }
    }
    // pruned inverse: the first n_keep outputs equal the full inverse, the rest keep what the previous pass left
    a = x;
    transform(a.data(), M, M, 1, false, nullptr);
    std::vector<double2> full(a.begin(), a.begin() + M), pruned(a.begin(), a.begin() + M);
    transform(full.data(), M, M, 1, true, nullptr);
    {
        PassArgs hooks;
        memset(&hooks, 0, sizeof(hooks));
        hooks.n_keep = M / 5 + 1;
        transform(pruned.data(), M, M, 1, true, &hooks);
        const double e_p = maxdiff(full.data(), pruned.data(), hooks.n_keep);
        if (e_p > 0) { printf("  pruned inverse differs by %.3e\n", e_p); ++bad; }
    }
    const FastPlan pl = make_fast_plan(M);
    printf("M = %lld (%d passes:", (long long)M, pl.n);
    for (int i = 0; i < pl.n; ++i) printf(" %dx%d", pl.p[i].n1, pl.p[i].n2);
    printf("): round trip %.2e, convolution %.2e -> %s\n", e_rt, e_cv, bad ? "FAIL" : "ok");
    return bad;
}

int main(int argc, char** argv) {
    int bad = 0;
    // register DFTs against the definition
    {
        double worst = 0;
        auto chk = [&](auto tag, int N) {
            (void)tag;
        };
        (void)chk;
#define CHECK_DFT(N)                                                                                          \
    for (int sign = -1; sign <= 1; sign += 2) {                                                               \
        double2 v[N], r[N];                                                                                   \
        for (int i = 0; i < N; ++i) v[i] = make_double2(frand(), frand());                                    \
        for (int k = 0; k < N; ++k) {                                                                         \
            long double sx = 0, sy = 0;                                                                       \
            for (int n = 0; n < N; ++n) {                                                                     \
                const long double ang = sign * 2.0L * 3.14159265358979323846264338327950288L * ((n * k) % N) / N; \
                sx += v[n].x * cosl(ang) - v[n].y * sinl(ang);                                                \
                sy += v[n].x * sinl(ang) + v[n].y * cosl(ang);                                                \
            }                                                                                                 \
            r[k] = make_double2((double)sx, (double)sy);                                                      \
        }                                                                                                     \
        if (sign < 0) Dft<N, -1>::run(v); else Dft<N, +1>::run(v);                                            \
        worst = fmax(worst, maxdiff(v, r, N));                                                                \
    }
        CHECK_DFT(2) CHECK_DFT(3) CHECK_DFT(4) CHECK_DFT(8) CHECK_DFT(9) CHECK_DFT(12) CHECK_DFT(16)
        printf("register DFTs 2/3/4/8/9/12/16: max error %.2e\n", worst);
        if (worst > 1e-14) ++bad;
    }
    std::vector<int64_t> lens;
    for (int i = 1; i < argc; ++i) lens.push_back(atoll(argv[i]));
    if (lens.empty()) lens = {64, 128, 256, 144, 192, 4096, 8192, 16384, 65536, 144 * 64, 192 * 128, 144 * 4096, 1 << 18, 1 << 19};
    for (int64_t M : lens) bad += check_length(M);
    if (fast_length_at_least(576000) != 589824 || fast_length_at_least(4584000) != 4718592 || fast_length_at_least(70) != 128 ||
        fast_length_at_least(3000) != 4096 || fast_length_at_least(200000) != 262144) {
        printf("fast_length_at_least: unexpected choice (%lld %lld %lld %lld %lld)\n", (long long)fast_length_at_least(576000),
               (long long)fast_length_at_least(4584000), (long long)fast_length_at_least(70), (long long)fast_length_at_least(3000),
               (long long)fast_length_at_least(200000));
        ++bad;
    }
    printf(bad ? "FAILED\n" : "all ok\n");
    return bad ? 1 : 0;
}
