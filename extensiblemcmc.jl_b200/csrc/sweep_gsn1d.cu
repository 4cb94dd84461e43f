// K2 -- the hot loop: full-data Gaussian log-likelihood sweep for all chains.
//
// Replaces loglikelihood(P::GsnTargetLaw, observs) (src/example/gsn_target.jl:23-29,
// reached through compute_ll! src/run.jl:251-260 and src/workspaces.jl:236-238),
// which in the reference is a sequential per-observation logpdf(MvNormal) loop for
// ONE chain.  Here every chain's sum  S_c = sum_i (x_i - mu_c)^2  is accumulated in
// one pass over the shared observation set; the O(1)-per-chain-step constants
// (-N/2 log(2 pi sigma^2), 1/(2 sigma^2)) are applied by the accept kernel.
// Observations are streamed, not summarised: the sweep is the stand-in for a
// general per-observation likelihood, so no sufficient-statistics shortcut is used.
//
// Data movement: each CTA streams its observation segment HBM/L2 -> shared memory
// with 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers, in a
// multi-stage ring; nobody touches an observation through the LSU global path.
//
// Two mappings, chosen by the host from the shapes:
//   "chains" (many chains, cfg 2): thread <-> R chains held in registers; every thread
//       of the CTA reads the same observation pair from shared memory (a broadcast
//       LDS.128 feeds 4R FP64 instructions), so the FP64 pipe is the only busy unit.
//       grid = (segments, chain groups); per-segment partial sums go to partial[S][C].
//   "obs" (few chains, huge N, cfg 5): thread <-> observations of the tile, all CB
//       chains' accumulators in registers; a fixed-order shuffle + shared-memory
//       block reduction ends each CTA.  HBM-bound: 8 bytes per observation per sweep.
// Both are deterministic: fixed tiling, fixed reduction order, no atomics.
//
// The same kernels serve the hierarchical-normal law (BASELINE cfg 4): observations come in G
// groups, group g has its own per-chain mean mu[g][C] (gridDim.z = G), and with GRAD = true the
// first-order sum  T = sum_i (x_i - mu)  is accumulated next to S for gradient-based updates.
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "dev_state.cuh"
#include "sweep.h"
#include "tma.cuh"
#include "exchange.cuh"

namespace extmcmc {

// Segment s of S over the n observations of one group, boundaries on even indices so that
// every bulk copy starts 16-byte aligned (group starts are even in the padded device layout).
__device__ __forceinline__ void segment_bounds(int64_t n_obs, int s, int S, int64_t &lo, int64_t &hi) {
    const int64_t n_pairs = (n_obs + 1) >> 1;
    lo = 2 * ((int64_t)s * n_pairs / S);
    hi = 2 * ((int64_t)(s + 1) * n_pairs / S);
    if (hi > n_obs) hi = n_obs;
}

// ---------------------------------------------------------------------------------
// "chains" mapping
// ---------------------------------------------------------------------------------
template <int R, int NT, int TILE, int STAGES, bool GRAD>
__global__ void __launch_bounds__(NT)
sweep_gsn1d_chains_kernel(Gsn1dArgs a) {
    __shared__ __align__(128) double tile[STAGES][TILE];
    __shared__ __align__(8) uint64_t bar[STAGES];
    const int tid = threadIdx.x;
    const int seg = blockIdx.x, g = blockIdx.z;
    const int S = a.S;
    const int64_t C = a.C;
    const int64_t cbase = (int64_t)blockIdx.y * (NT * R);
    const double *__restrict__ obs = a.obs + a.goff[g];
    const double *__restrict__ mu = a.mu + (int64_t)g * C;

    int64_t lo, hi;
    segment_bounds(a.glen[g], seg, S, lo, hi);
    const int64_t len = hi - lo;
    const int n_tiles = (int)((len + TILE - 1) / TILE);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t) {
        const int st = t % STAGES;
        const int64_t off = (int64_t)t * TILE;
        const int cnt = (int)((len - off) < (int64_t)TILE ? (len - off) : (int64_t)TILE);
        const uint32_t bytes = (uint32_t)((cnt + 1) >> 1) * 16u;  // padded device buffer
        mbar_expect_tx(&bar[st], bytes);
        bulk_g2s(&tile[st][0], obs + lo + off, bytes, &bar[st]);
    };
    if (tid == 0)
        for (int t = 0; t < STAGES && t < n_tiles; ++t) issue(t);

    // Everything above touches only the (constant) observations.  The per-chain means come from
    // the preceding kernel (proposal / accept): under PDL this kernel may have started before
    // that one finished, so wait for it here, then let our own successor start early.
    griddep_wait();
    griddep_launch_dependents();
    double m[R], acc[R], accT[GRAD ? R : 1];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t c = cbase + (int64_t)r * NT + tid;
        m[r] = c < C ? mu[c] : 0.0;
        acc[r] = 0.0;
        if (GRAD) accT[r] = 0.0;
    }

    for (int t = 0; t < n_tiles; ++t) {
        const int st = t % STAGES;
        mbar_wait(&bar[st], (uint32_t)(t / STAGES) & 1u);
        const int64_t off = (int64_t)t * TILE;
        const int cnt = (int)((len - off) < (int64_t)TILE ? (len - off) : (int64_t)TILE);
        const double2 *xs = reinterpret_cast<const double2 *>(&tile[st][0]);
        const int np = cnt >> 1;
#pragma unroll 4
        for (int i = 0; i < np; ++i) {
            const double2 x = xs[i];  // same address in every thread: broadcast LDS.128
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double d0 = x.x - m[r];
                acc[r] = fma(d0, d0, acc[r]);
                if (GRAD) accT[r] += d0;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double d1 = x.y - m[r];
                acc[r] = fma(d1, d1, acc[r]);
                if (GRAD) accT[r] += d1;
            }
        }
        if (cnt & 1) {
            const double x = tile[st][cnt - 1];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double d0 = x - m[r];
                acc[r] = fma(d0, d0, acc[r]);
                if (GRAD) accT[r] += d0;
            }
        }
        __syncthreads();  // everyone is done with this stage before it is refilled
        if (tid == 0 && t + STAGES < n_tiles) issue(t + STAGES);
    }
    const int64_t row = (int64_t)g * S + seg, rows = (int64_t)gridDim.z * S;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t c = cbase + (int64_t)r * NT + tid;
        if (c < C) {
            a.partial[row * C + c] = acc[r];
            if (GRAD) a.partial[(rows + row) * C + c] = accT[r];
        }
    }
}

// ---------------------------------------------------------------------------------
// "obs" mapping
// ---------------------------------------------------------------------------------
template <int CB, int NT, int TILE, int STAGES, bool GRAD>
__global__ void __launch_bounds__(NT)
sweep_gsn1d_obs_kernel(Gsn1dArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *tile = reinterpret_cast<double *>(smem_raw);                    // [STAGES][TILE]
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + STAGES * TILE);     // [STAGES]
    double *red = reinterpret_cast<double *>(bar + STAGES);                 // [2][NT/32][CB]
    const int tid = threadIdx.x;
    const int seg = blockIdx.x, g = blockIdx.z;
    const int S = a.S;
    const int64_t C = a.C;
    const double *__restrict__ obs = a.obs + a.goff[g];
    const double *__restrict__ mu = a.mu + (int64_t)g * C;

    int64_t lo, hi;
    segment_bounds(a.glen[g], seg, S, lo, hi);
    const int64_t len = hi - lo;
    const int n_tiles = (int)((len + TILE - 1) / TILE);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const uint64_t pol = a.stream_hint ? l2_policy_evict_first() : 0ull;
    auto issue = [&](int t) {
        const int st = t % STAGES;
        const int64_t off = (int64_t)t * TILE;
        const int cnt = (int)((len - off) < (int64_t)TILE ? (len - off) : (int64_t)TILE);
        const uint32_t bytes = (uint32_t)((cnt + 1) >> 1) * 16u;
        mbar_expect_tx(&bar[st], bytes);
        if (a.stream_hint) bulk_g2s_hint(tile + st * TILE, obs + lo + off, bytes, &bar[st], pol);
        else bulk_g2s(tile + st * TILE, obs + lo + off, bytes, &bar[st]);
    };
    if (tid == 0)
        for (int t = 0; t < STAGES && t < n_tiles; ++t) issue(t);

    griddep_wait();               // the means below are written by the preceding kernel (PDL)
    griddep_launch_dependents();
    double m[CB], acc[CB], accT[GRAD ? CB : 1];
#pragma unroll
    for (int c = 0; c < CB; ++c) {
        m[c] = c < C ? mu[c] : 0.0;
        acc[c] = 0.0;
        if (GRAD) accT[c] = 0.0;
    }

    auto eat = [&](const double2 x) {
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            const double d0 = x.x - m[c];
            acc[c] = fma(d0, d0, acc[c]);
            const double d1 = x.y - m[c];
            acc[c] = fma(d1, d1, acc[c]);
            if (GRAD) { accT[c] += d0; accT[c] += d1; }
        }
    };
    for (int t = 0; t < n_tiles; ++t) {
        const int st = t % STAGES;
        mbar_wait(&bar[st], (uint32_t)(t / STAGES) & 1u);
        const int64_t off = (int64_t)t * TILE;
        const int cnt = (int)((len - off) < (int64_t)TILE ? (len - off) : (int64_t)TILE);
        const double2 *xs = reinterpret_cast<const double2 *>(tile + st * TILE);
        if (cnt == TILE) {
#pragma unroll
            for (int k = 0; k < TILE / 2 / NT; ++k) eat(xs[k * NT + tid]);  // consecutive threads, consecutive 16 B
        } else {
            const int np = cnt >> 1;
            for (int i = tid; i < np; i += NT) eat(xs[i]);
            if ((cnt & 1) && tid == 0) {
                const double x = tile[st * TILE + cnt - 1];
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                    const double d0 = x - m[c];
                    acc[c] = fma(d0, d0, acc[c]);
                    if (GRAD) accT[c] += d0;
                }
            }
        }
        __syncthreads();
        if (tid == 0 && t + STAGES < n_tiles) issue(t + STAGES);
    }

    // fixed-order block reduction: xor-shuffle tree inside each warp, then warp 0..W-1
    constexpr int NW = NT / 32;
#pragma unroll
    for (int c = 0; c < CB; ++c) {
        double v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[(tid >> 5) * CB + c] = v;
        if (GRAD) {
            double w = accT[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            if ((tid & 31) == 0) red[(NW + (tid >> 5)) * CB + c] = w;
        }
    }
    __syncthreads();
    const int64_t row = (int64_t)g * S + seg, rows = (int64_t)gridDim.z * S;
    if (tid < CB && tid < C) {
        double v = 0.0;
        for (int w = 0; w < NW; ++w) v += red[w * CB + tid];
        a.partial[row * C + tid] = v;
        if (GRAD) {
            double t2 = 0.0;
            for (int w = 0; w < NW; ++w) t2 += red[(NW + w) * CB + tid];
            a.partial[(rows + row) * C + tid] = t2;
        }
    }
    if (GRAD || a.tail_mode == 0) return;

    // ---- fused tail: last CTA out reduces over the segments and delivers the totals ----------
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (tid == 0) last = atomicAdd(a.tail_counter, 1u) == gridDim.x * gridDim.z - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // CB chains x (NT / CB) slices; slice j adds rows j, j + NS, ... (independent loads, L2-hot),
    // then the slice sums are combined in slice order: a fixed order, like reduce_segments
    constexpr int NS = NT / CB;
    const int ch = tid % CB, sl = tid / CB;
    double acc2 = 0.0;
    if (ch < C)
        for (int64_t i = sl; i < rows; i += NS) acc2 += __ldcg(a.partial + i * C + ch);
    double *sh = reinterpret_cast<double *>(smem_raw);   // the tile ring is free now
    sh[sl * CB + ch] = acc2;
    __syncthreads();
    double tot = 0.0;
    if (tid < CB && tid < C)
        for (int j = 0; j < NS; ++j) tot += sh[j * CB + tid];
    if (a.tail_mode == 1) {
        if (tid < CB && tid < C) a.ssum[tid] = tot;
    } else {
        // one thread per (peer, chain): a tagged 16-byte cell straight into that rank's buffer
        __syncthreads();
        if (tid < CB) sh[tid] = tot;
        __syncthreads();
        const StepDesc *sd = reinterpret_cast<const StepDesc *>(a.descs) + a.k;
        const int parity = (int)(sd->xseq & 1);
        const uint32_t tag = (uint32_t)(sd->xseq + 1);
        for (int i = tid; i < a.world * CB; i += NT) {
            const int q = i / CB, ch = i % CB;
            if (ch < C) ll_store(a.peer_rx[q] + ll_cell(parity, a.world, a.rank, C, ch), sh[ch], tag);
        }
    }
    if (tid == 0) *a.tail_counter = 0u;
}

// ---------------------------------------------------------------------------------
// host side: shape -> plan -> launch
// ---------------------------------------------------------------------------------
namespace {
constexpr int kChainsNT = 128, kChainsTile = 1024, kChainsStages = 2;
constexpr int kObsNT = 256, kObsTile = 2048, kObsStages = 4;
constexpr size_t kObsSmem(int cb) {
    return (size_t)kObsStages * kObsTile * 8 + kObsStages * 8 + (size_t)2 * (kObsNT / 32) * cb * 8;
}

// launch with the PDL attribute: the kernel may be scheduled while its predecessor drains
template <typename Kern>
void launch_pdl(Kern kern, dim3 grid, int threads, size_t smem, cudaStream_t st, const Gsn1dArgs &a, bool early = false) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (pdl_mask() & 1) || (early && (pdl_mask() & 4));
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, a);
}

template <int R>
void launch_chains(const SweepPlan &pl, const Gsn1dArgs &a, bool grad, cudaStream_t st) {
    dim3 grid(pl.S, pl.groups, a.G);
    if (grad)
        launch_pdl(sweep_gsn1d_chains_kernel<R, kChainsNT, kChainsTile, kChainsStages, true>, grid, kChainsNT, 0, st, a);
    else
        launch_pdl(sweep_gsn1d_chains_kernel<R, kChainsNT, kChainsTile, kChainsStages, false>, grid, kChainsNT, 0, st, a);
}
template <int CB>
cudaError_t prep_obs() {
    cudaError_t e = cudaFuncSetAttribute(sweep_gsn1d_obs_kernel<CB, kObsNT, kObsTile, kObsStages, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kObsSmem(CB));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(sweep_gsn1d_obs_kernel<CB, kObsNT, kObsTile, kObsStages, true>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kObsSmem(CB));
}
template <int CB>
void launch_obs(const SweepPlan &pl, const Gsn1dArgs &a, bool grad, cudaStream_t st) {
    dim3 grid(pl.S, 1, a.G);
    if (grad)
        launch_pdl(sweep_gsn1d_obs_kernel<CB, kObsNT, kObsTile, kObsStages, true>, grid, kObsNT, kObsSmem(CB), st, a, true);
    else
        launch_pdl(sweep_gsn1d_obs_kernel<CB, kObsNT, kObsTile, kObsStages, false>, grid, kObsNT, kObsSmem(CB), st, a, true);
}
}  // namespace

// n_obs: observations of the LARGEST group (all of them when G = 1)
// EXTMCMC_PDL bits: 1 = every sweep may start early (the "chains" mapping loses its balanced wave:
// measured 45 % slower at cfg 2), 2 = the step kernels start early (prologue behind the sweep),
// 4 = the "obs" mapping (few chains, one CTA wave of long segments) may start early: its first tiles
// are in flight while the step kernel decides (cfg 5, 125 M observations per GPU: -2.4 %).
int pdl_mask() {
    static const int m = [] { const char *e = getenv("EXTMCMC_PDL"); return e ? atoi(e) : 6; }();
    return m;
}

SweepPlan plan_sweep_gsn1d(int64_t C, int64_t n_obs, int force_variant, int num_sms, int G) {
    SweepPlan pl{};
    pl.G = G;
    const int64_t n_pairs = (n_obs + 1) / 2;
    bool chains = C > 32;
    int force_R = 0;  // force_variant 11, 12, 14, 18: "chains" mapping with R = 1, 2, 4, 8 (tests)
    if (force_variant > 10) { force_R = force_variant - 10; force_variant = SWEEP_VARIANT_CHAINS; }
    if (force_variant == SWEEP_VARIANT_CHAINS) chains = true;
    if (force_variant == SWEEP_VARIANT_OBS && C <= 32) chains = false;
    if (chains) {
        // chains per thread: the largest R in {8,4,2,1} that still yields >= 2 CTAs per SM
        // (small N limits the number of segments, so small problems trade registers for CTAs)
        const int64_t max_S = (n_pairs + 127) / 128;  // >= 256 observations per segment
        int R = 8, S = 1, groups = 1;
        for (;; R >>= 1) {
            groups = (int)((C + (int64_t)kChainsNT * R - 1) / ((int64_t)kChainsNT * R));
            // 8 CTAs of 128 threads x 62 registers fill an SM's register file: 8 warps per scheduler hide
            // the DADD -> DFMA latency that ptxas's register-minimal schedule leaves exposed (measured at
            // cfg 2: 0.4724 ms per sweep with 4 CTAs per SM, 0.4643 with 8; EXTMCMC_CHAINS_CTAS overrides)
            static const int ctas_per_sm = [] { const char *e = getenv("EXTMCMC_CHAINS_CTAS"); return e && atoi(e) > 0 ? atoi(e) : 8; }();
            S = (num_sms * ctas_per_sm + groups * G - 1) / (groups * G);
            if (S > max_S) S = (int)max_S;
            if (S < 1) S = 1;
            // The FP64 pipe of an SM is shared by its resident CTAs, so the sweep lasts as long as
            // the busiest SM: ceil(CTAs / SMs) segments of ceil(n_pairs / S) pairs (+ ~32 pairs of
            // per-CTA prologue).  Pick the S near the target that minimises that product -- with
            // grouped data (cfg 4: 64 x S CTAs) the nearest multiple of the SM count is not S itself.
            {
                auto cost = [&](int s_) {
                    const int64_t ctas = (int64_t)groups * G * s_;
                    return ((ctas + num_sms - 1) / num_sms) * ((n_pairs + s_ - 1) / s_ + 32);
                };
                int best = S;
                for (int s_ = S > 3 ? S - 3 : 1; s_ <= S + 3 && s_ <= max_S; ++s_)
                    if (cost(s_) < cost(best)) best = s_;
                S = best;
            }
            const bool fits = C >= (int64_t)kChainsNT * R;      // no mostly-empty thread tiles
            if (force_R ? R == force_R : (R == 1 || (fits && (int64_t)groups * S * G >= 2 * num_sms))) break;
            if (R == 1) break;
        }
        pl.variant = SWEEP_VARIANT_CHAINS;
        pl.R = R;
        pl.groups = groups;
        pl.S = S;
        pl.launches = 1;
        static const char *names[] = {"", "gsn1d_chains_R1", "gsn1d_chains_R2", "", "gsn1d_chains_R4",
                                      "", "", "", "gsn1d_chains_R8"};
        pl.name = names[R];
    } else {
        int CB = 1;
        while (CB < C && CB < 32) CB <<= 1;
        pl.variant = SWEEP_VARIANT_OBS;
        pl.R = CB;
        pl.groups = 1;
        int S = (num_sms * 3 + G - 1) / G;  // 3 CTAs x 64 KB of staging per SM
        const int64_t max_S = (n_pairs + 1023) / 1024;  // >= one 2048-observation tile
        if (S > max_S) S = (int)max_S;
        if (S < 1) S = 1;
        pl.S = S;
        pl.launches = 1;
        pl.name = CB == 1 ? "gsn1d_obs_C1" : CB == 2 ? "gsn1d_obs_C2" : CB == 4 ? "gsn1d_obs_C4"
                : CB == 8 ? "gsn1d_obs_C8" : CB == 16 ? "gsn1d_obs_C16" : "gsn1d_obs_C32";
    }
    return pl;
}

cudaError_t sweep_gsn1d_init() {
    cudaError_t e;
    if ((e = prep_obs<1>()) != cudaSuccess) return e;
    if ((e = prep_obs<2>()) != cudaSuccess) return e;
    if ((e = prep_obs<4>()) != cudaSuccess) return e;
    if ((e = prep_obs<8>()) != cudaSuccess) return e;
    if ((e = prep_obs<16>()) != cudaSuccess) return e;
    if ((e = prep_obs<32>()) != cudaSuccess) return e;
    return cudaSuccess;
}

void launch_sweep_gsn1d(const SweepPlan &pl, const Gsn1dArgs &a, bool grad, cudaStream_t st) {
    if (pl.variant == SWEEP_VARIANT_CHAINS) {
        switch (pl.R) {
        case 1: launch_chains<1>(pl, a, grad, st); break;
        case 2: launch_chains<2>(pl, a, grad, st); break;
        case 4: launch_chains<4>(pl, a, grad, st); break;
        default: launch_chains<8>(pl, a, grad, st); break;
        }
    } else {
        switch (pl.R) {
        case 1: launch_obs<1>(pl, a, grad, st); break;
        case 2: launch_obs<2>(pl, a, grad, st); break;
        case 4: launch_obs<4>(pl, a, grad, st); break;
        case 8: launch_obs<8>(pl, a, grad, st); break;
        case 16: launch_obs<16>(pl, a, grad, st); break;
        default: launch_obs<32>(pl, a, grad, st); break;
        }
    }
}

}  // namespace extmcmc
