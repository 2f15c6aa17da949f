// GRU recurrence (reference models.py:56-63 nn.GRU; PyTorch gate order r,z,n; h0 = 0), forward and
// reverse-time backward, as persistent kernels: one CTA owns R batch rows of one direction for the
// whole sequence, so there is no inter-CTA communication and one block barrier per time step.
//
// Layout of a CTA: 4*H threads, thread (j, q) = (tid >> 2, tid & 3).
//   forward : thread (j,q) keeps W_hh[g*H + j, q*H/4 .. (q+1)*H/4) for the three gates g in registers
//             (3*H/4 floats), multiplies by the matching quarter of h (shared memory, broadcast loads)
//             and the four partial sums of a hidden unit meet by two warp-shuffle butterflies; the gate
//             non-linearities and the state update then run in the same threads -> one barrier per step.
//   backward: thread (k,q) keeps column k of W_hh for gate rows [g*H + q*H/4, +H/4); the step's
//             (d r_pre, d z_pre, d q) vector goes through shared memory (double-buffered).
// Everything the step needs from global memory (input projections, stashed gates, upstream
// gradients) is prefetched PF steps ahead into registers, so only shared memory and the shuffles
// sit on the serial dependency chain.  All arithmetic is fp32 (tolerance, SURVEY §7 hard part 2).
#include "mms_common.cuh"

namespace mms {

constexpr int GRU_MAX_DIRS = 2;
constexpr int PF = 4;   // prefetch distance in time steps

struct GruFwdParams {
    mms_gru_dir_fwd dir[GRU_MAX_DIRS];
    int B;
    float p;
    uint64_t seed, offset;
    const int64_t* offset_dev;
};
struct GruBwdParams {
    mms_gru_dir_bwd dir[GRU_MAX_DIRS];
    int B;
    float p;
    uint64_t seed, offset;
    const int64_t* offset_dev;
};

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

template <int H, int R>
__global__ void __launch_bounds__(4 * H) gru_fwd_kernel(const GruFwdParams prm) {
    constexpr int KS = H / 4;
    const mms_gru_dir_fwd& d = prm.dir[blockIdx.y];
    const int tid = threadIdx.x, j = tid >> 2, q = tid & 3;
    const int b0 = blockIdx.x * R;
    const int B = prm.B;

    __shared__ __align__(16) float hsm[2][R][H];

    // recurrent weights of this thread: rows g*H + j, columns [q*KS, q*KS + KS)
    float w[3][KS];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int i = 0; i < KS; ++i) w[g][i] = __ldg(d.w_hh + (size_t)(g * H + j) * H + q * KS + i);
    float bh[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) bh[g] = (q == 0) ? __ldg(d.b_hh + g * H + j) : 0.f;

    for (int i = tid; i < 2 * R * H; i += 4 * H) (&hsm[0][0][0])[i] = 0.f;

    DropRng rng;
    const bool do_drop = d.hs_drop != nullptr;
    if (do_drop) rng.init(prm.seed, resolve_offset(prm.offset, prm.offset_dev), prm.p);

    int bb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) bb[r] = min(b0 + r, B - 1);     // clamp loads; stores are guarded

    // gi ring: lane q < 3 holds gate q's input projection of unit j
    float ring[PF][R];
    const int gq = q < 3 ? q : 2;
#pragma unroll
    for (int u = 0; u < PF; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ring[u][r] = 0.f;
            if (u < d.nsteps) {
                const int t = d.t0 + u * d.dt;
                ring[u][r] = __ldg(d.gi + (size_t)bb[r] * d.gi_bs + (size_t)t * d.gi_ts + gq * H + j);
            }
        }
    __syncthreads();

    int cur = 0;
    for (int s0 = 0; s0 < d.nsteps; s0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int s = s0 + u;
            if (s < d.nsteps) {
                const int t = d.t0 + s * d.dt;
                float gi_own[R];
#pragma unroll
                for (int r = 0; r < R; ++r) gi_own[r] = ring[u][r];
                if (s + PF < d.nsteps) {
                    const int tn = d.t0 + (s + PF) * d.dt;
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        ring[u][r] = __ldg(d.gi + (size_t)bb[r] * d.gi_bs + (size_t)tn * d.gi_ts + gq * H + j);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float a0 = bh[0], a1 = bh[1], a2 = bh[2];
                    const float4* hv = reinterpret_cast<const float4*>(&hsm[cur][r][q * KS]);
#pragma unroll
                    for (int i4 = 0; i4 < KS / 4; ++i4) {
                        const float4 h4 = hv[i4];
                        a0 = fmaf(w[0][4 * i4 + 0], h4.x, a0); a1 = fmaf(w[1][4 * i4 + 0], h4.x, a1); a2 = fmaf(w[2][4 * i4 + 0], h4.x, a2);
                        a0 = fmaf(w[0][4 * i4 + 1], h4.y, a0); a1 = fmaf(w[1][4 * i4 + 1], h4.y, a1); a2 = fmaf(w[2][4 * i4 + 1], h4.y, a2);
                        a0 = fmaf(w[0][4 * i4 + 2], h4.z, a0); a1 = fmaf(w[1][4 * i4 + 2], h4.z, a1); a2 = fmaf(w[2][4 * i4 + 2], h4.z, a2);
                        a0 = fmaf(w[0][4 * i4 + 3], h4.w, a0); a1 = fmaf(w[1][4 * i4 + 3], h4.w, a1); a2 = fmaf(w[2][4 * i4 + 3], h4.w, a2);
                    }
                    // lanes 0 and 1 fold their input projection into the r / z partial sums
                    if (q == 0) a0 += gi_own[r];
                    if (q == 1) a1 += gi_own[r];
                    a0 = quad_sum(a0);
                    a1 = quad_sum(a1);
                    a2 = quad_sum(a2);                                        // = W_hn h + b_hn
                    const float gin = __shfl_sync(0xffffffffu, gi_own[r], 2, 4);   // lane 2 of the quad
                    const float rg = sigmoid_f(a0);
                    const float zg = sigmoid_f(a1);
                    const float ng = tanhf(fmaf(rg, a2, gin));
                    const float hp = hsm[cur][r][j];
                    const float hn = fmaf(zg, hp - ng, ng);                   // (1-z)*n + z*h
                    const bool live = (b0 + r) < B;
                    if (q == 0) {
                        hsm[cur ^ 1][r][j] = hn;
                        if (live) d.hs[(size_t)(b0 + r) * d.hs_bs + (size_t)t * d.hs_ts + j] = hn;
                    }
                    if (q == 1 && do_drop && live) {
                        const size_t e = (size_t)(b0 + r) * d.hs_bs + (size_t)t * d.hs_ts + j;
                        d.hs_drop[e] = hn * rng.mult((uint64_t)d.drop_base + e);
                    }
                    if (d.stash && live) {
                        const float sv = q == 0 ? rg : (q == 1 ? zg : (q == 2 ? ng : a2));
                        d.stash[(size_t)(b0 + r) * d.st_bs + (size_t)t * d.st_ts + q * H + j] = sv;
                    }
                }
                __syncthreads();
                cur ^= 1;
            }
        }
    }
}

template <int H, int R>
__global__ void __launch_bounds__(4 * H) gru_bwd_kernel(const GruBwdParams prm) {
    constexpr int KS = H / 4;
    const mms_gru_dir_bwd& d = prm.dir[blockIdx.y];
    const int tid = threadIdx.x, k = tid >> 2, q = tid & 3;
    const int b0 = blockIdx.x * R;
    const int B = prm.B;

    __shared__ __align__(16) float dgh[2][R][3 * H];

    // column k of W_hh, gate rows [g*H + q*KS, +KS)
    float w[3][KS];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int i = 0; i < KS; ++i) w[g][i] = __ldg(d.w_hh + (size_t)(g * H + q * KS + i) * H + k);

    DropRng rng;
    const bool do_mask = d.drop_mask != 0 && d.dout != nullptr;
    if (do_mask) rng.init(prm.seed, resolve_offset(prm.offset, prm.offset_dev), prm.p);

    int bb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) bb[r] = min(b0 + r, B - 1);

    // initial recurrent gradient: optional projection of the head gradient (dlast = dhid @ W0)
    float dh[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float s = 0.f;
        if (d.dh_head) {
            for (int i = 0; i < HEAD_HID; ++i)
                s = fmaf(__ldg(d.dh_head + (size_t)bb[r] * HEAD_HID + i), __ldg(d.w0 + (size_t)i * d.w0_ld + d.w0_col + k), s);
        }
        dh[r] = s;
    }

    // ring of per-step inputs: stash (r,z,n,qq), h_prev, dout -- every lane of the quad loads all
    struct StepIn { float r, z, n, qq, hp, dout; };
    StepIn ring[PF][R];
    auto load_step = [&](int s, int r) {
        StepIn v;
        const int t = d.t0 + s * d.dt;
        const float* sp = d.stash + (size_t)bb[r] * d.st_bs + (size_t)t * d.st_ts + k;
        v.r = __ldg(sp);
        v.z = __ldg(sp + H);
        v.n = __ldg(sp + 2 * H);
        v.qq = __ldg(sp + 3 * H);
        v.hp = s > 0 ? __ldg(d.hs + (size_t)bb[r] * d.hs_bs + (size_t)(t - d.dt) * d.hs_ts + k) : 0.f;
        float g = 0.f;
        if (d.dout) {
            const size_t e = (size_t)bb[r] * d.do_bs + (size_t)t * d.do_ts + k;
            g = __ldg(d.dout + e);
            if (do_mask) g *= rng.mult((uint64_t)d.drop_base + e);
        }
        if (d.dout_last && s == d.nsteps - 1) g += __ldg(d.dout_last + (size_t)bb[r] * d.dl_ld + k);
        v.dout = g;
        return v;
    };
    // steps are visited in reverse forward order: s = nsteps-1 ... 0; ring slot u <-> visit index
#pragma unroll
    for (int u = 0; u < PF; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int s = d.nsteps - 1 - u;
            if (s >= 0) ring[u][r] = load_step(s, r);
            else ring[u][r] = StepIn{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        }

    int buf = 0;
    for (int v0 = 0; v0 < d.nsteps; v0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int v = v0 + u;
            if (v < d.nsteps) {
                const int s = d.nsteps - 1 - v;
                const int t = d.t0 + s * d.dt;
                StepIn in[R];
#pragma unroll
                for (int r = 0; r < R; ++r) in[r] = ring[u][r];
                if (s - PF >= 0) {
#pragma unroll
                    for (int r = 0; r < R; ++r) ring[u][r] = load_step(s - PF, r);
                }
                float dhz[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const StepIn& x = in[r];
                    const float dht = dh[r] + x.dout;
                    const float dn = dht * (1.f - x.z);
                    const float dz = dht * (x.hp - x.n);
                    const float dnp = dn * (1.f - x.n * x.n);
                    const float dq = dnp * x.r;
                    const float dr = dnp * x.qq;
                    const float dzp = dz * x.z * (1.f - x.z);
                    const float drp = dr * x.r * (1.f - x.r);
                    dhz[r] = dht * x.z;
                    if (q < 3) dgh[buf][r][q * H + k] = q == 0 ? drp : (q == 1 ? dzp : dq);
                    if ((b0 + r) < B) {
                        const float ov = q == 0 ? drp : (q == 1 ? dzp : (q == 2 ? dnp : dq));
                        d.D[(size_t)(b0 + r) * d.d_bs + (size_t)t * d.d_ts + q * H + k] = ov;
                    }
                }
                __syncthreads();
                if (s > 0) {     // the gradient flowing into h_{-1} = h0 is not needed
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float a = 0.f;
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            const float4* gv = reinterpret_cast<const float4*>(&dgh[buf][r][g * H + q * KS]);
#pragma unroll
                            for (int i4 = 0; i4 < KS / 4; ++i4) {
                                const float4 g4 = gv[i4];
                                a = fmaf(w[g][4 * i4 + 0], g4.x, a);
                                a = fmaf(w[g][4 * i4 + 1], g4.y, a);
                                a = fmaf(w[g][4 * i4 + 2], g4.z, a);
                                a = fmaf(w[g][4 * i4 + 3], g4.w, a);
                            }
                        }
                        dh[r] = dhz[r] + quad_sum(a);
                    }
                }
                buf ^= 1;
            }
        }
    }
}

template <int H>
static int gru_fwd_dispatch(const GruFwdParams& prm, int ndirs, cudaStream_t st) {
    const int B = prm.B;
    // rows per CTA: keep at most ~2 waves of CTAs on 148 SMs
    int R = 1;
    while (R < 4 && (int64_t)cdiv(B, R) * ndirs > 296) R *= 2;
    dim3 grid(cdiv(B, R), ndirs);
    if (R == 1) gru_fwd_kernel<H, 1><<<grid, 4 * H, 0, st>>>(prm);
    else if (R == 2) gru_fwd_kernel<H, 2><<<grid, 4 * H, 0, st>>>(prm);
    else gru_fwd_kernel<H, 4><<<grid, 4 * H, 0, st>>>(prm);
    MMS_LAUNCH_CHECK("gru_fwd_kernel");
    return MMS_OK;
}

template <int H>
static int gru_bwd_dispatch(const GruBwdParams& prm, int ndirs, cudaStream_t st) {
    const int B = prm.B;
    int R = 1;
    while (R < 4 && (int64_t)cdiv(B, R) * ndirs > 296) R *= 2;
    dim3 grid(cdiv(B, R), ndirs);
    if (R == 1) gru_bwd_kernel<H, 1><<<grid, 4 * H, 0, st>>>(prm);
    else if (R == 2) gru_bwd_kernel<H, 2><<<grid, 4 * H, 0, st>>>(prm);
    else gru_bwd_kernel<H, 4><<<grid, 4 * H, 0, st>>>(prm);
    MMS_LAUNCH_CHECK("gru_bwd_kernel");
    return MMS_OK;
}

int launch_gru_fwd(const mms_gru_dir_fwd* dirs, int ndirs, int B, int H, float p, uint64_t seed, uint64_t offset,
                   const int64_t* offset_dev, cudaStream_t st) {
    MMS_REQUIRE(ndirs >= 1 && ndirs <= GRU_MAX_DIRS && B > 0, "gru_recur_fwd: bad direction count / batch");
    MMS_REQUIRE(H == 64 || H == 32, "gru_recur_fwd: hidden size %d not supported (32 or 64)", H);
    GruFwdParams prm;
    for (int i = 0; i < ndirs; ++i) {
        prm.dir[i] = dirs[i];
        MMS_REQUIRE(dirs[i].gi && dirs[i].w_hh && dirs[i].b_hh && dirs[i].hs && dirs[i].nsteps >= 0, "gru_recur_fwd: null pointer");
    }
    for (int i = ndirs; i < GRU_MAX_DIRS; ++i) prm.dir[i] = dirs[0];
    prm.B = B; prm.p = p; prm.seed = seed; prm.offset = offset; prm.offset_dev = offset_dev;
    return H == 64 ? gru_fwd_dispatch<64>(prm, ndirs, st) : gru_fwd_dispatch<32>(prm, ndirs, st);
}

int launch_gru_bwd(const mms_gru_dir_bwd* dirs, int ndirs, int B, int H, float p, uint64_t seed, uint64_t offset,
                   const int64_t* offset_dev, cudaStream_t st) {
    MMS_REQUIRE(ndirs >= 1 && ndirs <= GRU_MAX_DIRS && B > 0, "gru_recur_bwd: bad direction count / batch");
    MMS_REQUIRE(H == 64 || H == 32, "gru_recur_bwd: hidden size %d not supported (32 or 64)", H);
    GruBwdParams prm;
    for (int i = 0; i < ndirs; ++i) {
        prm.dir[i] = dirs[i];
        MMS_REQUIRE(dirs[i].w_hh && dirs[i].stash && dirs[i].hs && dirs[i].D && dirs[i].nsteps >= 0, "gru_recur_bwd: null pointer");
        MMS_REQUIRE(!dirs[i].dh_head || dirs[i].w0, "gru_recur_bwd: dh_head needs w0");
    }
    for (int i = ndirs; i < GRU_MAX_DIRS; ++i) prm.dir[i] = dirs[0];
    prm.B = B; prm.p = p; prm.seed = seed; prm.offset = offset; prm.offset_dev = offset_dev;
    return H == 64 ? gru_bwd_dispatch<64>(prm, ndirs, st) : gru_bwd_dispatch<32>(prm, ndirs, st);
}

}  // namespace mms

using namespace mms;

extern "C" int mms_gru_recur_fwd(const mms_gru_dir_fwd* dirs_host, int32_t ndirs, int32_t B, int32_t H, float dropout_p,
                                 uint64_t rng_seed, uint64_t rng_offset, const int64_t* rng_offset_dev, mms_stream_t stream) {
    MMS_REQUIRE(dirs_host, "gru_recur_fwd: null dirs");
    return launch_gru_fwd(dirs_host, ndirs, B, H, dropout_p, rng_seed, rng_offset, rng_offset_dev, (cudaStream_t)stream);
}
extern "C" int mms_gru_recur_bwd(const mms_gru_dir_bwd* dirs_host, int32_t ndirs, int32_t B, int32_t H, float dropout_p,
                                 uint64_t rng_seed, uint64_t rng_offset, const int64_t* rng_offset_dev, mms_stream_t stream) {
    MMS_REQUIRE(dirs_host, "gru_recur_bwd: null dirs");
    return launch_gru_bwd(dirs_host, ndirs, B, H, dropout_p, rng_seed, rng_offset, rng_offset_dev, (cudaStream_t)stream);
}
