// K1 (proposal), K3 (accept / commit / statistics / adaptation) and the MALA kernels.  One thread
// owns a chain's scalar work (everything SoA, coalesced across chains); the reductions of the
// sweep's partial sums and the covariance update of mid-sized models are shared by the 8 "slices"
// (threads) a CTA assigns to every chain.  Compiled with -fmad=false: the reference never
// contracts a*b+c, and the replay parity tests compare eps, running moments and trajectories
// bit-for-bit against the CPU oracle.
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
#include "dev_state.cuh"
#include "philox.cuh"
#include "step_kernels.h"
#include "tma.cuh"
#include "sweep.h"

namespace extmcmc {

namespace {
constexpr double kLog2Pi = 1.8378770664093454835606594728112;

// logpdf(prior, theta_loc) on the update's own coordinates (src/updates.jl:104,
// src/priors.jl:18-39).
__device__ __forceinline__ double log_prior_family(int kind, const double *pp, const double *th, int n) {
    switch (kind) {
    case EXTMCMC_PRIOR_IMPROPER: return 0.0;  // priors.jl:19
    case EXTMCMC_PRIOR_IMPROPER_POS: {        // -sum(log.(th)), priors.jl:26
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += log(th[i]);
        return -s;
    }
    case EXTMCMC_PRIOR_NORMAL: {
        const double m = pp[0], sd = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            const double z = (th[i] - m) / sd;
            s += -(z * z + kLog2Pi) / 2.0 - log(sd);
        }
        return s;
    }
    case EXTMCMC_PRIOR_GAMMA: {
        const double k = pp[0], sc = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] > 0.0)) return -INFINITY;
            s += -lgamma(k) - k * log(sc) + (k - 1.0) * log(th[i]) - th[i] / sc;
        }
        return s;
    }
    case EXTMCMC_PRIOR_UNIFORM: {
        const double a = pp[0], b = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] >= a && th[i] <= b)) return -INFINITY;
            s += -log(b - a);
        }
        return s;
    }
    case EXTMCMC_PRIOR_EXPONENTIAL: { /* Exponential(scale): log(rate) - rate x, rate = 1/scale; -Inf for x < 0 */
        const double rate = 1.0 / pp[0];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] >= 0.0)) return -INFINITY;
            s += log(rate) - rate * th[i];
        }
        return s;
    }
    case EXTMCMC_PRIOR_INV_GAMMA: { /* InverseGamma(a, sc): a log sc - lgamma(a) - (a + 1) log x - sc/x */
        const double a = pp[0], sc = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] > 0.0)) return -INFINITY;
            s += a * log(sc) - lgamma(a) - (a + 1.0) * log(th[i]) - sc / th[i];
        }
        return s;
    }
    case EXTMCMC_PRIOR_BETA: { /* Beta(a, b): (a-1) log x + (b-1) log1p(-x) - logbeta(a, b) on (0, 1) */
        const double a = pp[0], b = pp[1];
        const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] > 0.0 && th[i] < 1.0)) return -INFINITY;
            s += (a - 1.0) * log(th[i]) + (b - 1.0) * log1p(-th[i]) - lbeta;
        }
        return s;
    }
    case EXTMCMC_PRIOR_LOGNORMAL: { /* LogNormal(m, sd): logpdf(Normal(m, sd), log x) - log x */
        const double m = pp[0], sd = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (!(th[i] > 0.0)) return -INFINITY;
            const double lx = log(th[i]), z = (lx - m) / sd;
            s += (-(z * z + kLog2Pi) / 2.0 - log(sd)) - lx;
        }
        return s;
    }
    case EXTMCMC_PRIOR_CAUCHY: { /* Cauchy(m, sc): -(log1p(z^2) + log(pi) + log(sc)) */
        const double m = pp[0], sc = pp[1];
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            const double z = (th[i] - m) / sc;
            s += -(log1p(z * z) + 1.1447298858494001741434273513531 + log(sc));
        }
        return s;
    }
    }
    return NAN;
}

__device__ __forceinline__ double log_prior(const DevUpdate &u, const double *th) {
    if (u.prior != EXTMCMC_PRIOR_PRODUCT) return log_prior_family(u.prior, u.prior_params, th, u.n_coords);
    // ProductPrior (priors.jl:82-88): lp = 0.0; lp += logpdf(dist_k, th[idx_k])
    const int K = (int)u.prior_params[0];
    double lp = 0.0;
    int off = 0;
    for (int k = 0; k < K; ++k) {
        const double *f = u.prior_params + 1 + 4 * k;
        const int dim = (int)f[1];
        lp += log_prior_family((int)f[0], f + 2, th + off, dim);
        off += dim;
    }
    return lp;
}

// logpdf(rw::UniformRandomWalk, from, to) (random_walk.jl:88-94): only positive-
// constrained coordinates contribute, -log(2 eps_i) - log(to_i).
__device__ __forceinline__ double log_q_unif(const DevUpdate &u, const double *eps, const double *to) {
    double s = 0.0;
    for (int i = 0; i < u.n_coords; ++i) {
        const double t = u.pos[i] ? (-log(2.0 * eps[i]) - log(to[i])) : 0.0;
        s = (i == 0) ? t : s + t;
    }
    return s;
}

// ---- Gaussian random walks (random_walk.jl:123-232), n <= kMaxGaussCoords -------------------

// Lower Cholesky factor of Symmetric(S) (upper triangle of the column-major n x n S).
__device__ __forceinline__ bool chol_lower_sym_upper(const double *S, int n, double *L) {
    for (int j = 0; j < n; ++j) {
        double s = S[j + j * n];
        for (int k = 0; k < j; ++k) s -= L[j + k * n] * L[j + k * n];
        if (!(s > 0.0) || isinf(s)) return false;
        const double ljj = sqrt(s);
        L[j + j * n] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double a = S[j + i * n];
            for (int k = 0; k < j; ++k) a -= L[i + k * n] * L[j + k * n];
            L[i + j * n] = a / ljj;
        }
    }
    return true;
}

__device__ __forceinline__ double mvn_logpdf_chol(const double *L, int n, const double *mu, const double *x) {
    double z[kMaxGaussCoords], sq = 0.0, logdet = 0.0;
    for (int r = 0; r < n; ++r) {
        double a = x[r] - mu[r];
        for (int k = 0; k < r; ++k) a -= L[r + k * n] * z[k];
        z[r] = a / L[r + r * n];
        sq += z[r] * z[r];
        logdet += log(L[r + r * n]);
    }
    return -((double)n * kLog2Pi + 2.0 * logdet) / 2.0 - sq / 2.0;
}

// Sigma of a Gaussian walk as a dense local matrix: shared sigA, or this chain's sigB
__device__ __forceinline__ void load_sigma(const DevUpdate &u, bool useB, int64_t C, int64_t c, double *S) {
    const int nn = u.n_coords * u.n_coords;
    if (useB) for (int k = 0; k < nn; ++k) S[k] = u.sigB[(int64_t)k * C + c];
    else      for (int k = 0; k < nn; ++k) S[k] = u.sigA[k];
}

// logpdf(rw::GaussianRandomWalk, from, to) random_walk.jl:163-171, on transformed copies
__device__ __forceinline__ double log_q_gauss(const DevUpdate &u, const double *L, const double *from,
                                              const double *to) {
    const int n = u.n_coords;
    double tf[kMaxGaussCoords], tt[kMaxGaussCoords], s = 0.0;
    for (int i = 0; i < n; ++i) if (u.pos[i]) s += log(to[i]);
    const double logJ = -s;
    for (int i = 0; i < n; ++i) {
        tf[i] = u.pos[i] ? log(from[i]) : from[i];
        tt[i] = u.pos[i] ? log(to[i]) : to[i];
    }
    return mvn_logpdf_chol(L, n, tf, tt) + logJ;
}

// q(from -> to) - both directions share the factorisations LA / LB computed once per step
__device__ __forceinline__ double log_q_any(const DevUpdate &u, const double *eps, const double *LA,
                                            const double *LB, const double *from, const double *to) {
    if (u.kernel == EXTMCMC_KERNEL_RW_UNIFORM) return log_q_unif(u, eps, to);
    if (u.kernel == EXTMCMC_KERNEL_RW_GAUSS) return log_q_gauss(u, LA, from, to);
    // GaussianRandomWalkMix random_walk.jl:229-232 (no log-sum-exp guard, as in the reference)
    const double lpA = log_q_gauss(u, LA, from, to), lpB = log_q_gauss(u, LB, from, to);
    return log((1.0 - u.lambda) * exp(lpA) + u.lambda * exp(lpB));
}

// Per-chain law constants of a parameter vector.
//   GSN_IID_1D: lawc = { mu, c0 = -(log 2pi + 2 log sqrt(var))/2, 1/(2 var) },
//               ll = N c0 - S/(2 var),  S = sum (x - mu)^2     (gsn_target.jl:15-29, d = 1)
//   GSN_MV(d):  lawc = { mu[d], W = inv(L) lower-tri row-major, c0 },  Sigma = L L' built from the
//               UPPER triangle of the d x d block of theta (Symmetric(triu(S)), gsn_target.jl:19);
//               ll = N c0 - S/2,  S = sum |W (x - mu)|^2,  c0 = -(d log 2pi + 2 sum log L_ii)/2
__device__ __forceinline__ void law_prepare(const DevState &d, int64_t c, const double *full,
                                            int64_t stride) {
    if (d.law == EXTMCMC_LAW_GSN_IID_1D) {
        const double mu = full[0], var = full[stride];
        double c0, inv2;
        if (!(var > 0.0) || isinf(var)) {
            c0 = NAN; inv2 = NAN;
            *d.err_flag = 1;
        } else {
            c0 = -(kLog2Pi + 2.0 * log(sqrt(var))) / 2.0;
            inv2 = 0.5 / var;
        }
        d.lawc[c] = mu;
        d.lawc[d.C + c] = c0;
        d.lawc[2 * d.C + c] = inv2;
    } else if (d.law == EXTMCMC_LAW_HIER_NORMAL) {
        // no constants: the sweep reads theta_1..G straight from the state array it is given
        const double tau = full[(int64_t)(d.G + 1) * stride];
        if (!(tau > 0.0) || isinf(tau)) *d.err_flag = 1;
    } else if (d.law == EXTMCMC_LAW_GSN_MV) {
        const int n = d.obs_dim;
        double L[kMaxObsDim * kMaxObsDim], W[kMaxObsDim * kMaxObsDim];
        bool bad = false;
        // A[i][j] (i >= j) = theta[n + j + i*n]: entry (row j, col i) of the column-major block
        for (int j = 0; j < n && !bad; ++j) {
            double s = full[(int64_t)(n + j + j * n) * stride];
            for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
            if (!(s > 0.0) || isinf(s)) { bad = true; break; }
            const double ljj = sqrt(s);
            L[j * n + j] = ljj;
            for (int i = j + 1; i < n; ++i) {
                double a = full[(int64_t)(n + j + i * n) * stride];
                for (int k = 0; k < j; ++k) a -= L[i * n + k] * L[j * n + k];
                L[i * n + j] = a / ljj;
            }
        }
        double logdet = 0.0;
        if (!bad) {
            // W = inv(L): forward substitution column by column
            for (int j = 0; j < n; ++j) {
                W[j * n + j] = 1.0 / L[j * n + j];
                for (int i = j + 1; i < n; ++i) {
                    double a = 0.0;
                    for (int k = j; k < i; ++k) a -= L[i * n + k] * W[k * n + j];
                    W[i * n + j] = a / L[i * n + i];
                }
                logdet += log(L[j * n + j]);
            }
        } else {
            *d.err_flag = 1;
        }
        for (int j = 0; j < n; ++j) d.lawc[(int64_t)j * d.C + c] = full[(int64_t)j * stride];
        int w = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j, ++w) d.lawc[(int64_t)(n + w) * d.C + c] = bad ? NAN : W[i * n + j];
        d.lawc[(int64_t)(d.lawc_k - 1) * d.C + c] = bad ? NAN : -((double)n * kLog2Pi + 2.0 * logdet) / 2.0;
    }
}

// th: this chain's parameter vector the sums were computed for (stride C)
__device__ __forceinline__ double law_finalize(const DevState &d, int64_t c, double S, const double *th) {
    if (d.law == EXTMCMC_LAW_LOGISTIC) return S;  // the logistic sweep finishes ll itself
    if (d.law == EXTMCMC_LAW_GSN_IID_1D)  // N*c0 - S/(2 var)
        return (double)d.n_obs_total * d.lawc[d.C + c] - S * d.lawc[2 * d.C + c];
    if (d.law == EXTMCMC_LAW_HIER_NORMAL) {
        // sum_gj logN(y_gj; th_g, 1) + sum_g logN(th_g; mu, tau^2)
        const int G = d.G;
        const double mu = th[(int64_t)G * d.C], tau = th[(int64_t)(G + 1) * d.C];
        if (!(tau > 0.0) || isinf(tau)) return NAN;
        double dev2 = 0.0;
        for (int g = 0; g < G; ++g) { const double dv = th[(int64_t)g * d.C] - mu; dev2 += dv * dv; }
        return -0.5 * (double)d.n_obs_total * kLog2Pi - S / 2.0 +
               (double)G * (-0.5 * kLog2Pi - log(tau)) - dev2 / (2.0 * tau * tau);
    }
    return (double)d.n_obs_total * d.lawc[(int64_t)(d.lawc_k - 1) * d.C + c] - S / 2.0;
}
}  // namespace

// The schedule element and its update entry are read by every thread dozens of times;
// stage them in shared memory once per CTA (a dependent chain of global loads otherwise).
struct StepCtx {
    StepDesc sd;
    DevUpdate u;
};
__device__ __forceinline__ void load_step_ctx(StepCtx *ctx, const DevState &d, const StepDesc *descs, int k) {
    static_assert(sizeof(StepDesc) % 4 == 0 && sizeof(DevUpdate) % 4 == 0, "word copies");
    const uint32_t *src = reinterpret_cast<const uint32_t *>(descs + k);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&ctx->sd);
    for (int i = threadIdx.x; i < (int)(sizeof(StepDesc) / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    const uint32_t *us = reinterpret_cast<const uint32_t *>(d.upd + ctx->sd.pidx);
    uint32_t *ud = reinterpret_cast<uint32_t *>(&ctx->u);
    for (int i = threadIdx.x; i < (int)(sizeof(DevUpdate) / 4); i += blockDim.x) ud[i] = us[i];
    __syncthreads();
}

// ---------------------------------------------------------------------------------
// K1: proposal!  (src/updates.jl:191-196, rand(::UniformRandomWalk) random_walk.jl:65-73)
//     + set_proposal! (src/run.jl:221-240): writes the full proposal and the law
//     constants the sweep consumes.
// ---------------------------------------------------------------------------------
// state_regs / eps_regs: optional register (or shared-memory, element stride sstride) copies of the
// current full state and of this update's eps that a caller already holds; they spare the dependent
// global loads on the accept kernel's critical path.
__device__ __forceinline__ void propose_chain(const DevState &d, const StepDesc &sd, const DevUpdate &u,
                                              int64_t c, const double *state_regs = nullptr,
                                              const double *eps_regs = nullptr, int sstride = 1) {
    const int n = u.n_coords;
    double th[kMaxCoords], prop[kMaxCoords];
    for (int i = 0; i < n; ++i)
        th[i] = state_regs ? state_regs[u.coords[i] * sstride] : d.theta[(int64_t)u.coords[i] * d.C + c];

    uint32_t used = 0;
    if (d.rng_mode == EXTMCMC_RNG_REPLAY) {
        for (int i = 0; i < n; ++i)
            prop[i] = d.rp_prop[((int64_t)sd.replay_row * d.p_u_max + i) * d.C + c];
    } else {
        ChainStepStream rng(d.seed, (uint64_t)(d.chain_offset + c), sd.mcmciter, sd.pidx);
        for (;;) {
            if (u.kernel == EXTMCMC_KERNEL_RW_UNIFORM) {
                for (int i = 0; i < n; ++i) {
                    const double r = rng.next();
                    const double e = eps_regs ? eps_regs[i] : u.eps[(int64_t)i * d.C + c];
                    const double a = -e, b = e;
                    const double U = a + (b - a) * r;  // rand(Uniform(-eps, eps))
                    prop[i] = u.pos[i] ? th[i] * exp(U) : th[i] + U;  // random_walk.jl:72
                }
            } else {
                // rand(rw::GaussianRandomWalk[Mix]) random_walk.jl:145-151,213-227
                bool useB = false;
                if (u.kernel == EXTMCMC_KERNEL_RW_GAUSS_MIX) useB = rng.next() <= u.lambda;  // Bernoulli(lambda)
                double Sg[kMaxGaussCoords * kMaxGaussCoords], L[kMaxGaussCoords * kMaxGaussCoords], z[kMaxGaussCoords];
                load_sigma(u, useB, d.C, c, Sg);
                for (int q = 0; q < n; q += 2) {  // randn via Box-Muller on the uniform stream
                    const double u1 = rng.next(), u2 = rng.next();
                    const double rad = sqrt(-2.0 * log(u1));
                    double sn, cs;
                    sincospi(2.0 * u2, &sn, &cs);
                    z[q] = rad * cs;
                    if (q + 1 < n) z[q + 1] = rad * sn;
                }
                if (!chol_lower_sym_upper(Sg, n, L)) {
                    *d.err_flag = 1;  // reference: PosDefException from MvNormal(theta, Sigma)
                    for (int i = 0; i < n; ++i) prop[i] = NAN;
                    break;
                }
                for (int i = 0; i < n; ++i) {
                    double t = u.pos[i] ? log(th[i]) : th[i];
                    double a = 0.0;
                    for (int k = 0; k <= i; ++k) a += L[i + k * n] * z[k];
                    t = a + t;
                    prop[i] = u.pos[i] ? exp(t) : t;
                }
            }
            // whole-vector redraw while the prior is exactly -Inf (updates.jl:193-195)
            if (!(log_prior(u, prop) == -INFINITY)) break;
            if (rng.j > 60000u) break;
        }
        used = rng.j;
    }
    d.n_used[c] = used;
    for (int i = 0; i < n; ++i) d.prop_loc[(int64_t)i * d.C + c] = prop[i];
    // full proposal = current state with the update's coordinates replaced (run.jl:237-239)
    for (int j = 0; j < d.p; ++j)
        d.prop_full[(int64_t)j * d.C + c] = state_regs ? state_regs[j * sstride] : d.theta[(int64_t)j * d.C + c];
    for (int i = 0; i < n; ++i) d.prop_full[(int64_t)u.coords[i] * d.C + c] = prop[i];
    law_prepare(d, c, d.prop_full + c, d.C);
}

__global__ void __launch_bounds__(256)
propose_kernel(DevState d, const StepDesc *__restrict__ descs, int k) {
    __shared__ StepCtx ctx;
    load_step_ctx(&ctx, d, descs, k);
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    propose_chain(d, ctx.sd, ctx.u, c);
}

// Law constants of the CURRENT state (extmcmc_eval_loglik).
__global__ void __launch_bounds__(256) prepare_current_kernel(DevState d) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    law_prepare(d, c, d.theta + c, d.C);
}

// Fixed-order reduction of partial[.][C] over the segments (and observation groups).  A 256-thread
// CTA owns 256 / SL chains; slice j of a chain adds rows j, j + SL, j + 2 SL, ... (loads issued in
// batches of 8 before the first add), then the SL slice sums are combined in slice order.  The
// order depends only on the row count and SL, never on timing, so results are reproducible.
// SL = 1 (thread per chain) for few rows, 8 for many rows (cfg 2: 148), 32 for a handful of chains
// with hundreds of rows (cfg 5).
constexpr int kRedThreads = 256;


template <int kRedSlices>
__device__ __forceinline__ double reduce_segments(const DevState &d, double *sh /*[SL][256/SL]*/) {
    constexpr int kRedChains = kRedThreads / kRedSlices;
    const int lane_c = threadIdx.x % kRedChains;
    const int slice = threadIdx.x / kRedChains;
    const int64_t c = (int64_t)blockIdx.x * kRedChains + lane_c;
    double s = 0.0;
    if (c < d.C) {
        const double *p = d.partial + c;
        const int nrows = d.S * d.G;   // all segments of all observation groups
        // all loads of a batch are issued before the first add (memory-level parallelism: after a
        // large sweep these come from DRAM); the adds stay in increasing row order
        int i = slice;
        for (; i + 7 * kRedSlices < nrows; i += 8 * kRedSlices) {
            double a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = p[(int64_t)(i + q * kRedSlices) * d.C];
#pragma unroll
            for (int q = 0; q < 8; ++q) s += a[q];
        }
        for (; i < nrows; i += kRedSlices) s += p[(int64_t)i * d.C];
    }
    if (kRedSlices == 1) return s;
    sh[slice * kRedChains + lane_c] = s;
    __syncthreads();
    double tot = 0.0;
    if (slice == 0) {
#pragma unroll
        for (int j = 0; j < kRedSlices; ++j) tot += sh[j * kRedChains + lane_c];
    }
    return tot;  // valid in threads with slice == 0
}

// Per-group sums of both orders, out[q][g][c] = sum_i partial[q][g*S + i][c] (segment order): what a
// rank contributes to the all-reduce of a gradient sweep under observation sharding.  The MALA
// kernels then read `out` as a partial buffer with one segment per group.
__global__ void __launch_bounds__(256) reduce_group_sums_kernel(DevState d, double *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t C = d.C;
    if (idx >= (int64_t)2 * d.G * C) return;
    const int64_t r = idx / C, c = idx % C;          // r = q * G + g
    const double *p = d.partial + (r * d.S) * C + c;  // rows (q G + g) S .. + S of partial[2][G S][C]
    double s = 0.0;
    int i = 0;
    for (; i + 3 < d.S; i += 4) {
        double a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = p[(int64_t)(i + q) * C];
#pragma unroll
        for (int q = 0; q < 4; ++q) s += a[q];
    }
    for (; i < d.S; ++i) s += p[(int64_t)i * C];
    out[idx] = s;
}

template <int SL>
__global__ void __launch_bounds__(256) reduce_partials_kernel(DevState d) {
    __shared__ double sh[kRedThreads];
    constexpr int kRedChains = kRedThreads / SL;
    const double tot = reduce_segments<SL>(d, sh);
    const int64_t c = (int64_t)blockIdx.x * kRedChains + (threadIdx.x % kRedChains);
    if ((threadIdx.x / kRedChains) == 0 && c < d.C) d.ssum[c] = tot;
}

__device__ __forceinline__ unsigned long long exchange_tag(const DevState &d, const StepDesc &sd) {
    return (d.epoch << 40) | (unsigned long long)(sd.seq + 1);
}

// Reduce this rank's partial sums and push them to every rank (itself included) through peer
// pointers; the last CTA to finish raises this rank's flag on every peer (threadfence pattern).
template <int SL>
__global__ void __launch_bounds__(256)
reduce_push_kernel(DevState d, const StepDesc *__restrict__ descs, int k) {
    __shared__ double sh[kRedThreads];
    __shared__ bool last;
    constexpr int kRedChains = kRedThreads / SL;
    // PDL chain sweep -> reduce_push -> accept: start early, let the accept kernel start early
    // too (its prologue then overlaps the sweep), and wait for the sweep before touching its sums
    griddep_launch_dependents();
    const StepDesc sd = descs[k];
    const int parity = (int)(sd.seq & 1);
    griddep_wait();
    const double tot = reduce_segments<SL>(d, sh);
    const int64_t c = (int64_t)blockIdx.x * kRedChains + (threadIdx.x % kRedChains);
    if ((threadIdx.x / kRedChains) == 0 && c < d.C) {
        const int64_t slot = ((int64_t)parity * d.world + d.rank) * d.C + c;
        for (int q = 0; q < d.world; ++q) d.peer_rx[q][slot] = tot;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(d.push_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *d.push_counter = 0;
        __threadfence_system();
        const unsigned long long tag = exchange_tag(d, sd);
        for (int q = 0; q < d.world; ++q) {
            unsigned long long *f = d.peer_flag[q] + (parity * d.world + d.rank);
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(tag) : "memory");
        }
    }
}

// Wait until every rank's sums of this step have landed, then add them in rank order.
__device__ __forceinline__ double wait_and_combine(const DevState &d, const StepDesc &sd, int64_t c) {
    const int parity = (int)(sd.seq & 1);
    if (threadIdx.x == 0) {
        const unsigned long long tag = exchange_tag(d, sd);
        const long long t0 = clock64();
        for (int r = 0; r < d.world; ++r) {
            const unsigned long long *f = d.my_flag + (parity * d.world + r);
            unsigned long long v;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
                if (v < tag && clock64() - t0 > 6000000000ll) { *d.err_flag = 2; v = tag; }  // ~3 s
            } while (v < tag);
        }
    }
    __syncthreads();
    double s = 0.0;
    if (c < d.C)
        for (int r = 0; r < d.world; ++r) s += __ldcg(d.my_rx + ((int64_t)parity * d.world + r) * d.C + c);
    return s;
}

__global__ void __launch_bounds__(256) finalize_loglik_kernel(DevState d, double *ll_out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    ll_out[c] = law_finalize(d, c, d.ssum[c], d.theta + c);
}

// ---------------------------------------------------------------------------------
// What follows an accept/reject decision, shared by the random-walk and MALA paths:
// register_accept_reject_results! (src/run.jl:299-335) + set_chain_param! (:312-320, the
// caller has already committed theta) + update_stats! (src/chain_statistics.jl:41-66) +
// update_adaptation! (src/run.jl:136-173, src/transition_kernels/adaptation.jl:273-329).
// n_eps = entries of the update's step-size vector (p_u for the uniform walk, 1 for MALA).
// ---------------------------------------------------------------------------------
// Everything the post-decision code reads, fetched into registers BEFORE the decision is known
// (small parameter vectors only): under PDL these loads are issued while the sweep is still
// running, so that after the sweep only the partial sums remain on the critical path.
constexpr int kPreP = 4;
struct Prefetch {
    double state[kPreP];        // current full state theta
    double prop[kPreP];         // full proposal
    double mean[kPreP];
    double cov[kPreP * kPreP];
    double ra_prev;
    int acc_out;
    int32_t adapt_prop, adapt_acc;
    int64_t tot_prop, tot_acc;
    double eps[kMaxCoords];     // this update's step sizes
    double new_state[kPreP];    // out: state after the decision
};

__device__ __forceinline__ void prefetch_chain(const DevState &d, const StepDesc &sd, const DevUpdate &u,
                                               int64_t c, Prefetch &pf) {
    const int64_t C = d.C;
    const int p = d.p;
    for (int j = 0; j < p; ++j) {
        pf.state[j] = d.theta[(int64_t)j * C + c];
        pf.prop[j] = d.prop_full[(int64_t)j * C + c];
    }
    if (d.stats_mode != 2) {
        for (int j = 0; j < p; ++j) pf.mean[j] = d.mean[(int64_t)j * C + c];
        const int nc = d.stats_mode == 0 ? p * p : p;
        for (int j = 0; j < nc; ++j) pf.cov[j] = d.cov[(int64_t)j * C + c];
    }
    pf.ra_prev = sd.ra_prev_valid ? u.ra_val[c] : 0.0;
    pf.acc_out = sd.acc_out_valid ? (int)u.acc_ring[(sd.mcmciter % d.W) * C + c] : 0;
    pf.adapt_prop = u.adapt_prop[c];
    pf.adapt_acc = u.adapt_acc[c];
    pf.tot_prop = u.tot_prop[c];
    pf.tot_acc = u.tot_acc[c];
}

// Cooperative path for models with more than a handful of parameters (cfg 4: p = 10, full p x p
// covariance).  A CTA owns NCH chains; the chain's own thread stages the committed state and the
// OLD running mean in shared memory (and writes the new mean), then, after a barrier, ALL threads
// of the CTA update the NCH x p x p covariance entries -- same arithmetic, one entry per thread per
// pass, coalesced along the chain axis -- instead of one thread walking p^2 dependent loads.
constexpr int kCoopP = 32;   // largest p served this way (2 x kCoopP x NCH doubles of shared memory)
struct CoopStage {
    double *t;   // [p][nch] committed state (also feeds the fused next proposal)
    double *m;   // [p][nch] running mean before this step; nullptr unless the full covariance is kept
    double *mn;  // [p][nch] running mean after this step (spares two divisions per covariance entry)
    int nch, ch;
};

template <int NCH>
__device__ __forceinline__ void update_cov_coop(const DevState &d, int64_t N, int64_t c0, const double *sh_t,
                                                const double *sh_m, const double *sh_n) {
    const int p = d.p;
    const int64_t C = d.C;
    const double f_old = (double)(N - 1) / (double)N;
    const double f_new = (double)(N + 1) / (double)N;
    const int total = NCH * p * p;
    const int nt = (int)blockDim.x;
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * nt) {
        double cv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * nt;
            const int ch = i % NCH, e = i / NCH;
            cv[q] = (i < total && c0 + ch < C) ? d.cov[(int64_t)e * C + c0 + ch] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * nt;
            const int ch = i % NCH, e = i / NCH;
            if (i < total && c0 + ch < C) {
                const int a = e % p, b = e / p;
                const double ta = sh_t[a * NCH + ch], tb = sh_t[b * NCH + ch];
                const double ma_old = sh_m[a * NCH + ch], mb_old = sh_m[b * NCH + ch];
                const double ma_new = sh_n[a * NCH + ch], mb_new = sh_n[b * NCH + ch];
                const double old_sum_sq = f_old * cv[q] + ma_old * mb_old;
                const double new_sum_sq = old_sum_sq + (ta * tb) / (double)N;
                d.cov[(int64_t)e * C + c0 + ch] = new_sum_sq - f_new * (ma_new * mb_new);
            }
        }
    }
}

__device__ __forceinline__ void post_decision(const DevState &d, const StepDesc &sd, const DevUpdate &u,
                                              int64_t c, bool accepted, double ll_new, double ll_prop,
                                              int n_eps, Prefetch *pf = nullptr, const CoopStage *cs = nullptr) {
    const int64_t C = d.C;
    d.ll[c] = ll_new;
    // history row (state_history / state_proposal_history / ll_history / acceptance_history)
    const int64_t slot = sd.seq % d.H;
    if (pf) {
        // register path: identical arithmetic, no loads
        const int p = d.p;
        for (int j = 0; j < p; ++j) {
            pf->new_state[j] = accepted ? pf->prop[j] : pf->state[j];
            d.h_theta[(slot * p + j) * C + c] = pf->new_state[j];
            d.h_prop[(slot * p + j) * C + c] = pf->prop[j];
        }
        d.h_ll[slot * C + c] = ll_new;
        d.h_llp[slot * C + c] = ll_prop;
        d.h_acc[slot * C + c] = accepted ? 1 : 0;
        const int64_t N = sd.stat_n;
        if (d.stats_mode != 2) {
            const double f_old = (double)(N - 1) / (double)N;
            const double f_mean = (double)N / (double)(N + 1);
            const double f_new = (double)(N + 1) / (double)N;
            double nm[kPreP];
            for (int a = 0; a < p; ++a) nm[a] = pf->mean[a] * f_mean + pf->new_state[a] / (double)(N + 1);
            if (d.stats_mode == 0) {
                for (int b = 0; b < p; ++b)
                    for (int a = 0; a < p; ++a) {
                        const double old_sum_sq = f_old * pf->cov[a + b * p] + pf->mean[a] * pf->mean[b];
                        const double new_sum_sq = old_sum_sq + (pf->new_state[a] * pf->new_state[b]) / (double)N;
                        d.cov[(int64_t)(a + b * p) * C + c] = new_sum_sq - f_new * (nm[a] * nm[b]);
                    }
            } else {
                for (int a = 0; a < p; ++a) {
                    const double old_sum_sq = f_old * pf->cov[a] + pf->mean[a] * pf->mean[a];
                    const double new_sum_sq = old_sum_sq + (pf->new_state[a] * pf->new_state[a]) / (double)N;
                    d.cov[(int64_t)a * C + c] = new_sum_sq - f_new * (nm[a] * nm[a]);
                }
            }
            for (int a = 0; a < p; ++a) d.mean[(int64_t)a * C + c] = nm[a];
        }
        {
            const int W = d.W;
            const int64_t mn = (int64_t)W < N ? (int64_t)W : N;
            u.ra_val[c] = (pf->ra_prev * (double)W + (double)((int)accepted - pf->acc_out)) / (double)mn;
            u.acc_ring[(sd.mcmciter % W) * C + c] = accepted ? 1 : 0;
        }
        u.tot_prop[c] = pf->tot_prop + 1;
        u.tot_acc[c] = pf->tot_acc + (accepted ? 1 : 0);
        if (u.adapt_kind == EXTMCMC_ADAPT_UNIF_RW || u.adapt_kind == EXTMCMC_ADAPT_MALA) {
            int32_t prop_n = pf->adapt_prop + 1;
            int32_t acc_n = pf->adapt_acc + (accepted ? 1 : 0);
            if (prop_n >= u.adapt_every_k) {
                const double r = (double)sd.mcmciter / (double)u.adapt_every_k - u.offset;
                const double delta = u.scale / sqrt(r > 1.0 ? r : 1.0);
                const double a_r = (double)acc_n / (double)prop_n;
                prop_n = 0; acc_n = 0;
                const double sgn = (a_r > u.target) ? 1.0 : -1.0;
                for (int i = 0; i < n_eps; ++i) {
                    double e = pf->eps[i] + sgn * delta;
                    e = e < u.vmax ? e : u.vmax;
                    e = e > u.vmin ? e : u.vmin;
                    pf->eps[i] = e;
                    u.eps[(int64_t)i * C + c] = e;
                }
            }
            u.adapt_prop[c] = prop_n;
            u.adapt_acc[c] = acc_n;
        }
        return;
    }
    const int64_t N = sd.stat_n;
    const double f_old = (double)(N - 1) / (double)N;
    const double f_mean = (double)N / (double)(N + 1);
    const double f_new = (double)(N + 1) / (double)N;
    const bool coop_full = cs && cs->m;        // full covariance left to update_cov_coop
    const bool diag = d.stats_mode == 1;
    // History row, staging and -- diagonal statistics / cooperative path -- update_stats!
    // (chain_statistics.jl:46-51, verbatim arithmetic); loads in batches of 4 ahead of the stores
    // (the stores may alias the loads as far as the compiler knows, so a plain loop serialises).
    for (int j0 = 0; j0 < d.p; j0 += 4) {
        double t[4], pr[4], m[4], cv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (j0 + q < d.p) {
                t[q] = d.theta[(int64_t)(j0 + q) * C + c];
                pr[q] = d.prop_full[(int64_t)(j0 + q) * C + c];
                if (coop_full || diag) m[q] = d.mean[(int64_t)(j0 + q) * C + c];
                if (diag) cv[q] = d.cov[(int64_t)(j0 + q) * C + c];
            }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (j0 + q < d.p) {
                const int j = j0 + q;
                d.h_theta[(slot * d.p + j) * C + c] = t[q];
                d.h_prop[(slot * d.p + j) * C + c] = pr[q];
                if (cs) cs->t[j * cs->nch + cs->ch] = t[q];
                if (coop_full || diag) {
                    const double m_new = m[q] * f_mean + t[q] / (double)(N + 1);
                    if (coop_full) { cs->m[j * cs->nch + cs->ch] = m[q]; cs->mn[j * cs->nch + cs->ch] = m_new; }
                    if (diag) {
                        const double old_sum_sq = f_old * cv[q] + m[q] * m[q];
                        const double new_sum_sq = old_sum_sq + (t[q] * t[q]) / (double)N;
                        d.cov[(int64_t)j * C + c] = new_sum_sq - f_new * (m_new * m_new);
                    }
                    d.mean[(int64_t)j * C + c] = m_new;
                }
            }
    }
    d.h_ll[slot * C + c] = ll_new;
    d.h_llp[slot * C + c] = ll_prop;
    d.h_acc[slot * C + c] = accepted ? 1 : 0;

    // full covariance by this thread alone (more than kCoopP parameters, or no spare threads)
    if (d.stats_mode == 0 && !coop_full) {
        const int p = d.p;
        // covariance first (it needs the old mean), column by column
        for (int b = 0; b < p; ++b) {
            const double tb = d.theta[(int64_t)b * C + c];
            const double mb_old = d.mean[(int64_t)b * C + c];
            const double mb_new = mb_old * f_mean + tb / (double)(N + 1);
            for (int a = 0; a < p; ++a) {
                const double ta = d.theta[(int64_t)a * C + c];
                const double ma_old = d.mean[(int64_t)a * C + c];
                const double ma_new = ma_old * f_mean + ta / (double)(N + 1);
                const int64_t idx = ((int64_t)(a + b * p)) * C + c;
                const double old_sum_sq = f_old * d.cov[idx] + ma_old * mb_old;
                const double new_sum_sq = old_sum_sq + (ta * tb) / (double)N;
                d.cov[idx] = new_sum_sq - f_new * (ma_new * mb_new);
            }
        }
        for (int a = 0; a < p; ++a) {
            const int64_t idx = (int64_t)a * C + c;
            d.mean[idx] = d.mean[idx] * f_mean + d.theta[idx] / (double)(N + 1);
        }
    }
    // rolling acceptance rate (chain_statistics.jl:53-64)
    {
        const int W = d.W;
        const double ra_prev = sd.ra_prev_valid ? u.ra_val[c] : 0.0;
        const int64_t rslot = sd.mcmciter % W;
        const int acc_out = sd.acc_out_valid ? (int)u.acc_ring[rslot * C + c] : 0;
        const int64_t mn = (int64_t)W < N ? (int64_t)W : N;
        u.ra_val[c] = (ra_prev * (double)W + (double)((int)accepted - acc_out)) / (double)mn;
        u.acc_ring[rslot * C + c] = accepted ? 1 : 0;
    }

    // update_adaptation! -- only the update whose turn it is registers (run.jl:176-177)
    u.tot_prop[c] += 1;
    u.tot_acc[c] += accepted ? 1 : 0;
    if (u.adapt_kind == EXTMCMC_ADAPT_UNIF_RW || u.adapt_kind == EXTMCMC_ADAPT_MALA) {
        int32_t prop_n = u.adapt_prop[c] + 1;                      // register! :292-295
        int32_t acc_n = u.adapt_acc[c] + (accepted ? 1 : 0);
        if (prop_n >= u.adapt_every_k) {                           // time_to_update :302-304
            const double r = (double)sd.mcmciter / (double)u.adapt_every_k - u.offset;
            const double delta = u.scale / sqrt(r > 1.0 ? r : 1.0);  // compute_delta :312-319
            const double a_r = (double)acc_n / (double)prop_n;       // acceptance_rate :242-244
            prop_n = 0; acc_n = 0;                                   // reset! :263-266
            const double sgn = (a_r > u.target) ? 1.0 : -1.0;
            for (int i = 0; i < n_eps; ++i) {                        // compute_eps :326-329
                double e = u.eps[(int64_t)i * C + c] + sgn * delta;
                e = e < u.vmax ? e : u.vmax;
                e = e > u.vmin ? e : u.vmin;
                u.eps[(int64_t)i * C + c] = e;
            }
        }
        u.adapt_prop[c] = prop_n;
        u.adapt_acc[c] = acc_n;
    }
    // HaarioTypeAdaptation registers on EVERY update step of ANY update (adaptation.jl:399-414),
    // on that update's view of the (already committed) global state, log-transformed copy.
    if (d.n_haario > 0) {
        for (int v = 0; v < d.NU; ++v) {
            const DevUpdate &w = (v == sd.pidx) ? u : d.upd[v];
            if (w.adapt_kind != EXTMCMC_ADAPT_HAARIO) continue;
            const int m = w.n_coords;
            double t[kMaxGaussCoords], om[kMaxGaussCoords], nm[kMaxGaussCoords];
            for (int i = 0; i < m; ++i) {
                const double x = d.theta[(int64_t)w.coords[i] * C + c];
                t[i] = w.pos[i] ? log(x) : x;
                om[i] = w.hmean[(int64_t)i * C + c];
            }
            const int64_t hn = sd.stat_n;  // adpt.N: starts at 1, +1 per registration = per executed step
            const double f_old = (double)(hn - 1) / (double)hn, f_mean = (double)hn / (double)(hn + 1);
            const double f_new = (double)(hn + 1) / (double)hn;
            for (int i = 0; i < m; ++i) {
                nm[i] = om[i] * f_mean + t[i] / (double)(hn + 1);
                w.hmean[(int64_t)i * C + c] = nm[i];
            }
            const bool ready = (v == sd.pidx) && sd.haario_ready;
            for (int b = 0; b < m; ++b)
                for (int a = 0; a < m; ++a) {
                    const int64_t idx = (int64_t)(a + b * m) * C + c;
                    const double old_sum_sq = f_old * w.hcov[idx] + om[a] * om[b];
                    const double new_sum_sq = old_sum_sq + (t[a] * t[b]) / (double)hn;
                    const double cv = new_sum_sq - f_new * (nm[a] * nm[b]);
                    w.hcov[idx] = cv;
                    // readjust!(rw::GaussianRandomWalkMix, ...) adaptation.jl:422-426
                    if (ready) w.sigB[idx] = (2.38 * 2.38) / (double)m * cv;
                }
        }
    }
}

__device__ __forceinline__ double draw_exp(const DevState &d, const StepDesc &sd, int64_t c) {
    if (d.rng_mode == EXTMCMC_RNG_REPLAY) return d.rp_exp[(int64_t)sd.replay_row * d.C + c];
    ChainStepStream rng(d.seed, (uint64_t)(d.chain_offset + c), sd.mcmciter, sd.pidx, d.n_used[c]);
    return -log(rng.next());  // rand(Exponential(1.0)), run.jl:278
}

// ---------------------------------------------------------------------------------
// K3: accept_reject! (src/run.jl:268-281) for the random-walk updates.
// ---------------------------------------------------------------------------------
template <int SL>
__global__ void __launch_bounds__(256)
accept_kernel(DevState d, const StepDesc *__restrict__ descs, int k, int fuse_next) {
    // 256 threads = (256/SL) chains x SL reduction slices; slice 0 carries on with the chain
    constexpr int kRedChains = kRedThreads / SL;
    __shared__ double sh[kRedThreads];
    __shared__ StepCtx ctx, ctx_next;
    // staging of the cooperative covariance update (only the sliced layouts have spare threads)
    constexpr int kStage = SL >= 8 ? kCoopP * kRedChains : 1;
    __shared__ double sh_t[kStage], sh_m[kStage], sh_n[kStage];
    // PDL: this kernel may have been scheduled while the likelihood sweep is still running.
    // Everything up to griddep_wait() only READS state that was final before the sweep started
    // (chain state, proposal, step sizes, law constants, RNG counters) -- the transition-density
    // and prior terms and the Exp(1) draw are computed here, hidden behind the sweep.
    griddep_launch_dependents();
    load_step_ctx(&ctx, d, descs, k);
    if (fuse_next == 1) load_step_ctx(&ctx_next, d, descs, k + 1);
    const int64_t c = (int64_t)blockIdx.x * kRedChains + (threadIdx.x % kRedChains);
    const bool worker = (threadIdx.x / kRedChains) == 0 && c < d.C;
    const StepDesc &sd = ctx.sd;
    const DevUpdate &u = ctx.u;
    const int n = u.n_coords;
    const int64_t C = d.C;
    double th[kMaxCoords], prop[kMaxCoords], eps[kMaxCoords];
    double q_back = 0.0, q_fwd = 0.0, lp_prop = 0.0, lp_cur = 0.0, E = 0.0, ll_cur = 0.0;
    if (worker) {
        // update_workspaces! (run.jl:101-112): ll of the previously executed update; on the
        // very first element it is still the initial -Inf (workspaces.jl:425)
        ll_cur = sd.first ? -INFINITY : d.ll[c];
        for (int i = 0; i < n; ++i) {
            th[i] = d.theta[(int64_t)u.coords[i] * C + c];
            prop[i] = d.prop_loc[(int64_t)i * C + c];
            eps[i] = u.kernel == EXTMCMC_KERNEL_RW_UNIFORM ? u.eps[(int64_t)i * C + c] : 0.0;
        }
        if (u.kernel == EXTMCMC_KERNEL_RW_UNIFORM) {
            q_back = log_q_unif(u, eps, th);    // theta° -> theta
            q_fwd = log_q_unif(u, eps, prop);   // theta -> theta°
        } else {
            double Sg[kMaxGaussCoords * kMaxGaussCoords], LA[kMaxGaussCoords * kMaxGaussCoords],
                LB[kMaxGaussCoords * kMaxGaussCoords];
            load_sigma(u, false, C, c, Sg);
            bool ok = chol_lower_sym_upper(Sg, n, LA);
            if (u.kernel == EXTMCMC_KERNEL_RW_GAUSS_MIX) {
                load_sigma(u, true, C, c, Sg);
                ok = chol_lower_sym_upper(Sg, n, LB) && ok;
            }
            if (!ok) {
                *d.err_flag = 1;
                q_back = NAN;
            } else {
                q_back = log_q_any(u, eps, LA, LB, prop, th);   // theta° -> theta
                q_fwd = log_q_any(u, eps, LA, LB, th, prop);    // theta -> theta°
            }
        }
        lp_prop = log_prior(u, prop);
        lp_cur = log_prior(u, th);
        E = draw_exp(d, sd, c);
    }
    // register path for small models without Haario adaptation (cfg 1, 2, 5)
    const bool use_pf = d.p <= kPreP && d.n_haario == 0;
    Prefetch pf;
    double law0 = 0.0, law1 = 0.0, eps_next[kMaxCoords];
    bool next_eps_ok = false;
    if (worker && use_pf) {
        prefetch_chain(d, sd, u, c, pf);
        for (int i = 0; i < n; ++i) pf.eps[i] = eps[i];
        if (d.law == EXTMCMC_LAW_GSN_IID_1D) { law0 = d.lawc[C + c]; law1 = d.lawc[2 * C + c]; }
        // step sizes of the NEXT element's update (fused proposal); not when it is this very update,
        // whose eps the adaptation below may still change
        if (fuse_next == 1 && ctx_next.sd.pidx != sd.pidx && ctx_next.u.kernel == EXTMCMC_KERNEL_RW_UNIFORM) {
            for (int i = 0; i < ctx_next.u.n_coords; ++i) eps_next[i] = ctx_next.u.eps[(int64_t)i * C + c];
            next_eps_ok = true;
        }
    }

    griddep_wait();   // the sweep (and, under sharding, the exchange) has finished
    double S;
    if (d.p2p) {
        S = wait_and_combine(d, ctx.sd, c);
    } else if (d.use_ssum) {
        S = c < d.C ? d.ssum[c] : 0.0;
    } else {
        S = reduce_segments<SL>(d, sh);
    }
    // full covariance of a model with more than kPreP parameters: all slices share the work
    const bool stage = SL >= 8 && !use_pf && d.p <= kCoopP;   // CTA-uniform
    const bool coop = stage && d.stats_mode == 0;
    if (worker) {
        const double ll_prop = (use_pf && d.law == EXTMCMC_LAW_GSN_IID_1D)
                                   ? (double)d.n_obs_total * law0 - S * law1   // = law_finalize, constants prefetched
                                   : law_finalize(d, c, S, d.prop_full + c);
        // llr, strictly left to right (run.jl:271-277)
        double llr = ll_prop - ll_cur;
        llr = llr + q_back;
        llr = llr - q_fwd;
        llr = llr + lp_prop;
        llr = llr - lp_cur;

        const bool accepted = E > -llr;  // NaN compares false -> reject
        const double ll_new = accepted ? ll_prop : ll_cur;
        if (accepted)
            for (int i = 0; i < n; ++i) d.theta[(int64_t)u.coords[i] * C + c] = prop[i];
        const int n_eps = u.kernel == EXTMCMC_KERNEL_RW_UNIFORM ? n : 0;
        const CoopStage cs{sh_t, coop ? sh_m : nullptr, sh_n, kRedChains, (int)(threadIdx.x % kRedChains)};
        post_decision(d, sd, u, c, accepted, ll_new, ll_prop, n_eps, use_pf ? &pf : nullptr, stage ? &cs : nullptr);
    }
    if (coop) {
        __syncthreads();
        update_cov_coop<kRedChains>(d, sd.stat_n, (int64_t)blockIdx.x * kRedChains, sh_t, sh_m, sh_n);
    }
    if (!worker) return;
    // proposal of the NEXT schedule element of this block, fused here: the chain's thread
    // already holds its freshly committed state, and one launch per update step is saved
    if (fuse_next == 2) {
        law_prepare(d, c, d.theta + c, C);   // what prepare_current_kernel would do
    } else if (fuse_next) {
        if (use_pf) {
            const double *en = next_eps_ok ? eps_next
                               : (ctx_next.sd.pidx == sd.pidx && ctx_next.u.kernel == EXTMCMC_KERNEL_RW_UNIFORM) ? pf.eps
                                                                                                            : nullptr;
            propose_chain(d, ctx_next.sd, ctx_next.u, c, pf.new_state, en);
        } else if (stage) {
            propose_chain(d, ctx_next.sd, ctx_next.u, c, sh_t + (threadIdx.x % kRedChains), nullptr, kRedChains);
        } else {
            propose_chain(d, ctx_next.sd, ctx_next.u, c);
        }
    }
}

// ---------------------------------------------------------------------------------
// Gradient path (MALAUpdate; the reference only has the hooks: MCMCGradientBasedUpdate
// src/types.jl:24, compute_gradients_and_momenta! src/updates.jl:129-133 called at
// src/run.jl:110,259, the `∇ll` buffer src/workspaces.jl:417).
//
// grad_finalize_kernel: fixed-order reduction of the sweep's partial sums per chain (and
// per observation group), then ll and d ll / d theta for ALL p parameters.
//   GSN_IID_1D : d/dmu = T/var, d/dvar = -N/(2 var) + S/(2 var^2)
//   HIER_NORMAL: theta = [th_1..th_G, mu, tau], y_gj ~ N(th_g, 1), th_g ~ N(mu, tau^2) (the
//                hierarchical term lives in the law because priors only see their own
//                coordinates, src/run.jl:374-385):
//                d/dth_g = T_g - (th_g - mu)/tau^2, d/dmu = sum_g (th_g - mu)/tau^2,
//                d/dtau = -G/tau + sum_g (th_g - mu)^2 / tau^3
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void grad_finalize_chain(const DevState &d, int64_t c, const double *__restrict__ src,
                                                    double *__restrict__ ll_out, double *__restrict__ grad_out) {
    const int64_t C = d.C;
    const int G = d.G, S = d.S;
    const int64_t rows = (int64_t)G * S;
    if (d.law == EXTMCMC_LAW_GSN_IID_1D) {
        double s2 = 0.0, s1 = 0.0;
        for (int i = 0; i < S; ++i) { s2 += d.partial[(int64_t)i * C + c]; s1 += d.partial[(rows + i) * C + c]; }
        const double var = src[C + c];
        ll_out[c] = law_finalize(d, c, s2, src + c);
        grad_out[c] = s1 / var;
        grad_out[C + c] = -(double)d.n_obs_total / (2.0 * var) + s2 / (2.0 * var * var);
    } else if (d.law == EXTMCMC_LAW_HIER_NORMAL) {
        const double mu = src[(int64_t)G * C + c], tau = src[(int64_t)(G + 1) * C + c];
        if (!(tau > 0.0) || isinf(tau)) *d.err_flag = 1;   // the current state never went through law_prepare
        const double it2 = 1.0 / (tau * tau);
        double s2_tot = 0.0, dmu = 0.0, dev2 = 0.0;
        for (int g = 0; g < G; ++g) {
            double s2 = 0.0, s1 = 0.0;
            for (int i = 0; i < S; ++i) {
                s2 += d.partial[((int64_t)g * S + i) * C + c];
                s1 += d.partial[(rows + (int64_t)g * S + i) * C + c];
            }
            s2_tot += s2;
            const double dv = src[(int64_t)g * C + c] - mu;
            grad_out[(int64_t)g * C + c] = s1 - dv * it2;
            dmu += dv * it2;
            dev2 += dv * dv;
        }
        grad_out[(int64_t)G * C + c] = dmu;
        grad_out[(int64_t)(G + 1) * C + c] = -(double)G / tau + dev2 * it2 / tau;
        ll_out[c] = law_finalize(d, c, s2_tot, src + c);
    }
}

// The MALA step kernels run as CTAs of kMalaChains chains x kMalaSlices slices (256 threads): the
// slices share the per-group segment sums of the hierarchical law and the covariance update, the
// chain's own thread (slice 0) does the scalar work.  Same sums, same order as grad_finalize_chain.
constexpr int kMalaChains = 32, kMalaSlices = 8;
constexpr int kCoopG = 16;   // most observation groups reduced this way
__device__ __forceinline__ void grad_finalize_coop(const DevState &d, int64_t c0, const double *__restrict__ src,
                                                   double *__restrict__ ll_out, double *__restrict__ grad_out,
                                                   double *sh2, double *sh1 /*[kCoopG][kMalaChains] each*/) {
    const int ch = threadIdx.x % kMalaChains, slice = threadIdx.x / kMalaChains;
    const int64_t c = c0 + ch, C = d.C;
    const int G = d.G, S = d.S;
    if (d.law != EXTMCMC_LAW_HIER_NORMAL || G > kCoopG) {   // CTA-uniform
        if (slice == 0 && c < C) grad_finalize_chain(d, c, src, ll_out, grad_out);
        return;
    }
    const int64_t rows = (int64_t)G * S;
    if (c < C)
        for (int g = slice; g < G; g += kMalaSlices) {
            const double *p2 = d.partial + ((int64_t)g * S) * C + c;
            const double *p1 = d.partial + (rows + (int64_t)g * S) * C + c;
            double s2 = 0.0, s1 = 0.0;
            int i = 0;
            for (; i + 3 < S; i += 4) {   // loads first, adds in segment order
                double a[4], b[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { a[q] = p2[(int64_t)(i + q) * C]; b[q] = p1[(int64_t)(i + q) * C]; }
#pragma unroll
                for (int q = 0; q < 4; ++q) { s2 += a[q]; s1 += b[q]; }
            }
            for (; i < S; ++i) { s2 += p2[(int64_t)i * C]; s1 += p1[(int64_t)i * C]; }
            sh2[g * kMalaChains + ch] = s2;
            sh1[g * kMalaChains + ch] = s1;
        }
    __syncthreads();
    if (slice != 0 || c >= C) return;
    const double mu = src[(int64_t)G * C + c], tau = src[(int64_t)(G + 1) * C + c];
    if (!(tau > 0.0) || isinf(tau)) *d.err_flag = 1;   // the current state never went through law_prepare
    const double it2 = 1.0 / (tau * tau);
    double s2_tot = 0.0, dmu = 0.0, dev2 = 0.0;
    for (int g = 0; g < G; ++g) {
        s2_tot += sh2[g * kMalaChains + ch];
        const double dv = src[(int64_t)g * C + c] - mu;
        grad_out[(int64_t)g * C + c] = sh1[g * kMalaChains + ch] - dv * it2;
        dmu += dv * it2;
        dev2 += dv * dv;
    }
    grad_out[(int64_t)G * C + c] = dmu;
    grad_out[(int64_t)(G + 1) * C + c] = -(double)G / tau + dev2 * it2 / tau;
    ll_out[c] = law_finalize(d, c, s2_tot, src + c);
}

__global__ void __launch_bounds__(128)
grad_finalize_kernel(DevState d, const double *__restrict__ src, double *__restrict__ ll_out,
                     double *__restrict__ grad_out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    grad_finalize_chain(d, c, src, ll_out, grad_out);
}

// d log prior / d theta_i for the priors that have one on the device
__device__ __forceinline__ double prior_grad(const DevUpdate &u, double th) {
    if (u.prior == EXTMCMC_PRIOR_NORMAL) return -(th - u.prior_params[0]) / (u.prior_params[1] * u.prior_params[1]);
    return 0.0;  // ImproperPrior
}
__device__ __forceinline__ double prior_logpdf1(const DevUpdate &u, double th) {
    if (u.prior == EXTMCMC_PRIOR_NORMAL) {
        const double z = (th - u.prior_params[0]) / u.prior_params[1];
        return -(z * z + kLog2Pi) / 2.0 - log(u.prior_params[1]);
    }
    return 0.0;
}

// K5a: MALA proposal  theta° = theta + (tau^2/2) g(theta) + tau z,  g = grad(ll + log prior)
__global__ void __launch_bounds__(kMalaChains * kMalaSlices)
mala_propose_kernel(DevState d, const StepDesc *__restrict__ descs, int k, int finalize_cur,
                    double *__restrict__ ll_scratch) {
    __shared__ StepCtx ctx;
    __shared__ double sh2[kCoopG * kMalaChains], sh1[kCoopG * kMalaChains];
    // PDL (as in accept_kernel): the schedule element and update entry are staged while the
    // preceding sweep is still running; nothing the predecessor writes is read before the wait
    griddep_launch_dependents();
    load_step_ctx(&ctx, d, descs, k);
    griddep_wait();
    const int64_t c0 = (int64_t)blockIdx.x * kMalaChains;
    const int64_t c = c0 + threadIdx.x % kMalaChains;
    // the sweep just before this kernel evaluated the CURRENT state: finish its sums here
    // (gradient of the current state) instead of in a kernel of its own
    if (finalize_cur) grad_finalize_coop(d, c0, d.theta, ll_scratch, d.grad_cur, sh2, sh1);
    if (threadIdx.x >= kMalaChains || c >= d.C) return;
    const StepDesc &sd = ctx.sd;
    const DevUpdate &u = ctx.u;
    const int64_t C = d.C;
    const int n = u.n_coords;
    const double tau = u.eps[c], h2 = tau * tau / 2.0;
    for (int j0 = 0; j0 < d.p; j0 += 4) {   // prop_full <- theta, loads ahead of the stores
        double t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) if (j0 + q < d.p) t[q] = d.theta[(int64_t)(j0 + q) * C + c];
#pragma unroll
        for (int q = 0; q < 4; ++q) if (j0 + q < d.p) d.prop_full[(int64_t)(j0 + q) * C + c] = t[q];
    }
    if (d.rng_mode == EXTMCMC_RNG_REPLAY) {
        for (int i = 0; i < n; ++i)
            d.prop_full[(int64_t)u.coords_dev[i] * C + c] = d.rp_prop[((int64_t)sd.replay_row * d.p_u_max + i) * C + c];
        d.n_used[c] = 0;
    } else {
        ChainStepStream rng(d.seed, (uint64_t)(d.chain_offset + c), sd.mcmciter, sd.pidx);
        for (int i = 0; i < n; i += 2) {
            const double u1 = rng.next(), u2 = rng.next();
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincospi(2.0 * u2, &sn, &cs);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (i + q >= n) break;
                const int64_t j = u.coords_dev[i + q];
                const double th = d.theta[j * C + c];
                const double g = d.grad_cur[j * C + c] + prior_grad(u, th);
                d.prop_full[j * C + c] = th + h2 * g + tau * (rad * (q ? sn : cs));
            }
        }
        d.n_used[c] = rng.j;
    }
    law_prepare(d, c, d.prop_full + c, C);
}

// K5b: MALA accept/reject.  log q(a -> b) = -|b - a - (tau^2/2) g(a)|^2 / (2 tau^2) (the
// normalising constant is the same in both directions and is left out).
// mala_decide: the chain's own thread -- decision, commit, history, counters; sh_t / sh_m: staging
// (see CoopStage; with sh_m the covariance update is left to update_cov_coop).
__device__ __forceinline__ void mala_decide(const DevState &d, const StepCtx &ctx, int64_t c, double *sh_t,
                                            double *sh_m, double *sh_n, int ch) {
    const StepDesc &sd = ctx.sd;
    const DevUpdate &u = ctx.u;
    const int64_t C = d.C;
    const int n = u.n_coords;
    const double tau = u.eps[c], h2 = tau * tau / 2.0;
    double qf = 0.0, qb = 0.0, lp_prop = 0.0, lp_cur = 0.0;
    for (int i0 = 0; i0 < n; i0 += 2) {   // loads of two coordinates in flight; sums in index order
        double a[2], b[2], ga[2], gb[2];
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (i0 + q < n) {
                const int64_t j = u.coords_dev[i0 + q];
                a[q] = d.theta[j * C + c];
                b[q] = d.prop_full[j * C + c];
                ga[q] = d.grad_cur[j * C + c];
                gb[q] = d.grad_prop[j * C + c];
            }
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (i0 + q < n) {
                const double gaq = ga[q] + prior_grad(u, a[q]);
                const double gbq = gb[q] + prior_grad(u, b[q]);
                const double rf = b[q] - a[q] - h2 * gaq, rb = a[q] - b[q] - h2 * gbq;
                qf += rf * rf;
                qb += rb * rb;
                lp_prop += prior_logpdf1(u, b[q]);
                lp_cur += prior_logpdf1(u, a[q]);
            }
    }
    const double inv = 1.0 / (2.0 * tau * tau);
    qf = -qf * inv;  // theta -> theta°
    qb = -qb * inv;  // theta° -> theta
    const double ll_prop = d.ll_prop[c];
    const double ll_cur = sd.first ? -INFINITY : d.ll[c];
    double llr = ll_prop - ll_cur;  // same association as run.jl:271-277
    llr = llr + qb;
    llr = llr - qf;
    llr = llr + lp_prop;
    llr = llr - lp_cur;
    const double E = draw_exp(d, sd, c);
    const bool accepted = E > -llr;
    const double ll_new = accepted ? ll_prop : ll_cur;
    if (accepted) {
        for (int i = 0; i < n; ++i) {
            const int64_t j = u.coords_dev[i];
            d.theta[j * C + c] = d.prop_full[j * C + c];
        }
        for (int j0 = 0; j0 < d.p; j0 += 4) {   // grad_cur <- grad_prop
            double g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) if (j0 + q < d.p) g[q] = d.grad_prop[(int64_t)(j0 + q) * C + c];
#pragma unroll
            for (int q = 0; q < 4; ++q) if (j0 + q < d.p) d.grad_cur[(int64_t)(j0 + q) * C + c] = g[q];
        }
    }
    const CoopStage cs{sh_t, sh_m, sh_n, kMalaChains, ch};
    post_decision(d, sd, u, c, accepted, ll_new, ll_prop, 1, nullptr, sh_t ? &cs : nullptr);
}

__global__ void __launch_bounds__(kMalaChains * kMalaSlices)
mala_accept_kernel(DevState d, const StepDesc *__restrict__ descs, int k, int finalize_prop, int fuse_next) {
    __shared__ StepCtx ctx, ctx_next;
    __shared__ double sh2[kCoopG * kMalaChains], sh1[kCoopG * kMalaChains];
    __shared__ double sh_t[kCoopP * kMalaChains], sh_m[kCoopP * kMalaChains], sh_n[kCoopP * kMalaChains];
    griddep_launch_dependents();
    load_step_ctx(&ctx, d, descs, k);
    if (fuse_next) load_step_ctx(&ctx_next, d, descs, k + 1);
    griddep_wait();   // the gradient sweep of the proposal has finished
    const int64_t c0 = (int64_t)blockIdx.x * kMalaChains;
    const int ch = threadIdx.x % kMalaChains;
    const int64_t c = c0 + ch;
    if (finalize_prop) grad_finalize_coop(d, c0, d.prop_full, d.ll_prop, d.grad_prop, sh2, sh1);
    const bool worker = threadIdx.x < kMalaChains && c < d.C;
    const bool stage = d.p <= kCoopP;   // CTA-uniform
    const bool coop = stage && d.stats_mode == 0;
    if (worker) mala_decide(d, ctx, c, stage ? sh_t : nullptr, coop ? sh_m : nullptr, sh_n, ch);
    if (coop) {
        __syncthreads();
        update_cov_coop<kMalaChains>(d, ctx.sd.stat_n, c0, sh_t, sh_m, sh_n);
    }
    // next element is a random-walk update: issue its proposal here (one launch saved)
    if (worker && fuse_next) {
        if (stage) propose_chain(d, ctx_next.sd, ctx_next.u, c, sh_t + ch, nullptr, kMalaChains);
        else propose_chain(d, ctx_next.sd, ctx_next.u, c);
    }
}

// ---------------------------------------------------------------------------------
// utilities
// ---------------------------------------------------------------------------------
// x_i ~ N(mean, sd^2), i = global observation index; Box-Muller on the Philox stream
// keyed by (seed; counter = i/2).  BASELINE cfg 5 generates 1e9 observations in place.
__global__ void generate_obs_normal_kernel(double *obs, int64_t first, int64_t n, double mean,
                                           double sd, uint64_t seed) {
    const int64_t pair0 = first >> 1;
    const int64_t n_pairs = ((first + n + 1) >> 1) - pair0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs;
         q += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t pair = (uint64_t)(pair0 + q);
        const Philox4 w = philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), 0x0B5E0B5Eu, 0u,
                                        (uint32_t)seed, (uint32_t)(seed >> 32));
        const double u1 = u52_to_unit(w.w[0], w.w[1]), u2 = u52_to_unit(w.w[2], w.w[3]);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        const int64_t i0 = (int64_t)(pair << 1), i1 = i0 + 1;
        if (i0 >= first && i0 < first + n) obs[i0 - first] = mean + sd * (rad * cs);
        if (i1 >= first && i1 < first + n) obs[i1 - first] = mean + sd * (rad * sn);
    }
}

__global__ void flush_l2_kernel(double *buf, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        buf[i] = v;
}

// Dependent-free FP64 FMA chains: 8 accumulators x iters per thread.
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// Independent FP64 tensor-core MMA chains (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4): 8 accumulator
// pairs per warp, 512 flop per instruction.
__global__ void dmma_peak_kernel(double *out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = (double)threadIdx.x; c[i][1] = (double)i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
static inline int blocks_for(int64_t C) { return (int)((C + 255) / 256); }

void launch_propose(const DevState &d, const StepDesc *descs, int k, cudaStream_t st) {
    propose_kernel<<<blocks_for(d.C), 256, 0, st>>>(d, descs, k);
}
static inline int red_blocks_for(int64_t C, int sl) {
    const int ch = kRedThreads / sl;
    return (int)((C + ch - 1) / ch);
}
// reduction slices per chain: 1 (thread per chain) for few segments; 8; 32 for a handful of chains
// with hundreds of segments (cfg 5), so that the cold loads of the partial sums overlap
static inline int slices_for(const DevState &d) {
    // a full covariance of more than kPreP parameters is updated by all slices (update_cov_coop)
    // (and its state is staged in shared memory for the fused next proposal)
    const bool stage = !(d.p <= kPreP && d.n_haario == 0) && d.p <= kCoopP;
    if (!stage && (d.use_ssum || d.S * d.G <= 16)) return 1;
    return d.C <= 8 ? 32 : 8;
}
void launch_accept(const DevState &d, const StepDesc *descs, int k, int fuse_next, cudaStream_t st) {
    // PDL attribute: the accept kernel's prologue overlaps the tail of the sweep
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (pdl_mask() >> 1) & 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (slices_for(d) == 32) {
        cfg.gridDim = dim3(red_blocks_for(d.C, 32));
        cudaLaunchKernelEx(&cfg, accept_kernel<32>, d, descs, k, fuse_next);
    } else if (slices_for(d) == 8) {
        cfg.gridDim = dim3(red_blocks_for(d.C, 8));
        cudaLaunchKernelEx(&cfg, accept_kernel<8>, d, descs, k, fuse_next);
    } else {
        cfg.gridDim = dim3(red_blocks_for(d.C, 1));
        cudaLaunchKernelEx(&cfg, accept_kernel<1>, d, descs, k, fuse_next);
    }
}
void launch_grad_finalize(const DevState &d, const double *src, double *ll_out, double *grad_out,
                          cudaStream_t st) {
    grad_finalize_kernel<<<(int)((d.C + 127) / 128), 128, 0, st>>>(d, src, ll_out, grad_out);
}
void launch_mala_propose(const DevState &d, const StepDesc *descs, int k, int finalize_cur, double *ll_scratch,
                         cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((d.C + kMalaChains - 1) / kMalaChains));
    cfg.blockDim = dim3(kMalaChains * kMalaSlices);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (pdl_mask() >> 1) & 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, mala_propose_kernel, d, descs, k, finalize_cur, ll_scratch);
}
void launch_mala_accept(const DevState &d, const StepDesc *descs, int k, int finalize_prop, int fuse_next,
                        cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((d.C + kMalaChains - 1) / kMalaChains));
    cfg.blockDim = dim3(kMalaChains * kMalaSlices);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (pdl_mask() >> 1) & 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, mala_accept_kernel, d, descs, k, finalize_prop, fuse_next);
}
void launch_prepare_current(const DevState &d, cudaStream_t st) {
    prepare_current_kernel<<<blocks_for(d.C), 256, 0, st>>>(d);
}
void launch_reduce_group_sums(const DevState &d, double *out, cudaStream_t st) {
    const int64_t n = (int64_t)2 * d.G * d.C;
    reduce_group_sums_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(d, out);
}
void launch_reduce_partials(const DevState &d, cudaStream_t st) {
    if (d.S * d.G > 16)
        reduce_partials_kernel<8><<<red_blocks_for(d.C, 8), 256, 0, st>>>(d);
    else
        reduce_partials_kernel<1><<<red_blocks_for(d.C, 1), 256, 0, st>>>(d);
}
void launch_reduce_push(const DevState &d, const StepDesc *descs, int k, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (pdl_mask() >> 1) & 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (d.S * d.G > 16) {
        cfg.gridDim = dim3(red_blocks_for(d.C, 8));
        cudaLaunchKernelEx(&cfg, reduce_push_kernel<8>, d, descs, k);
    } else {
        cfg.gridDim = dim3(red_blocks_for(d.C, 1));
        cudaLaunchKernelEx(&cfg, reduce_push_kernel<1>, d, descs, k);
    }
}
void launch_finalize_loglik(const DevState &d, double *ll_out, cudaStream_t st) {
    finalize_loglik_kernel<<<blocks_for(d.C), 256, 0, st>>>(d, ll_out);
}
void launch_generate_obs_normal(double *obs, int64_t first, int64_t n, double mean, double sd,
                                uint64_t seed, int num_sms, cudaStream_t st) {
    generate_obs_normal_kernel<<<num_sms * 8, 256, 0, st>>>(obs, first, n, mean, sd, seed);
}
void launch_flush_l2(double *buf, int64_t n, int num_sms, cudaStream_t st) {
    flush_l2_kernel<<<num_sms * 8, 256, 0, st>>>(buf, n, 1.0);
}
void launch_dmma_peak(double *out, int iters, int num_sms, cudaStream_t st) {
    dmma_peak_kernel<<<num_sms * 8, 256, 0, st>>>(out, iters, 1.0000001, 1e-9);
}
void launch_fp64_peak(double *out, int iters, int num_sms, cudaStream_t st) {
    fp64_peak_kernel<<<num_sms * 8, 256, 0, st>>>(out, iters, 1.0000001, 1e-9);
}

}  // namespace extmcmc
