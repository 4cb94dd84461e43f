// K2 for the general-d GsnTargetLaw (src/example/gsn_target.jl:1-29): theta = [mu; vec(Sigma)],
// loglikelihood = sum_i logpdf(MvNormal(mu, Symmetric(triu(Sigma))), x_i).
//
// Per chain the proposal kernel factorises Sigma = L L' and stores W = inv(L) (lower
// triangular) next to mu in lawc; the sweep accumulates  S_c = sum_i |W_c (x_i - mu_c)|^2
// and the accept kernel finishes ll = N c0 - S/2 with c0 = -(d log 2pi + logdet Sigma)/2.
// Same data movement as the 1-D sweep: TMA bulk copies (UBLKCP) of observation tiles into a
// shared-memory ring, mbarrier completion, fixed tiling and reduction order, no atomics.
//
//   "chains" mapping (C > 32): one chain per thread (its mu and W in registers), the CTA's
//       threads read the same observation from shared memory (broadcast).
//   "obs" mapping (C <= 32): one chain per CTA column (blockIdx.y), threads stride over the
//       observations of the tile, fixed-order block reduction.
#include <cstdint>
#include <cuda_runtime.h>
#include "sweep.h"
#include "tma.cuh"

namespace extmcmc {

template <int D>
struct MvConst {
    double mu[D];
    double W[D * (D + 1) / 2];  // row-major lower triangle: W00, W10, W11, W20, ...
    __device__ __forceinline__ void load(const double *lawc, int64_t C, int64_t c) {
#pragma unroll
        for (int j = 0; j < D; ++j) mu[j] = lawc[(int64_t)j * C + c];
#pragma unroll
        for (int j = 0; j < D * (D + 1) / 2; ++j) W[j] = lawc[(int64_t)(D + j) * C + c];
    }
    __device__ __forceinline__ double quad(const double *x) const {
        double dd[D];
#pragma unroll
        for (int j = 0; j < D; ++j) dd[j] = x[j] - mu[j];
        double q = 0.0;
        int w = 0;
#pragma unroll
        for (int r = 0; r < D; ++r) {
            double z = W[w++] * dd[0];
#pragma unroll
            for (int k = 1; k <= r; ++k) z = fma(W[w++], dd[k], z);
            q = fma(z, z, q);
        }
        return q;
    }
};

// observation-index segment [lo, hi) of segment s; lo*D is even so bulk copies stay 16 B aligned
__device__ __forceinline__ void mv_segment(int64_t n_obs, int s, int S, int64_t &lo, int64_t &hi) {
    const int64_t n_pairs = (n_obs + 1) >> 1;
    lo = 2 * ((int64_t)s * n_pairs / S);
    hi = 2 * ((int64_t)(s + 1) * n_pairs / S);
    if (hi > n_obs) hi = n_obs;
}

template <int D, int NT, int TILE, int STAGES, bool OBS_MAPPED>
__global__ void __launch_bounds__(NT)
sweep_gsnmv_kernel(const double *__restrict__ obs, int64_t n_obs, const double *__restrict__ lawc,
                   int64_t C, double *__restrict__ partial, int S) {
    __shared__ __align__(128) double tile[STAGES][TILE * D];
    __shared__ __align__(8) uint64_t bar[STAGES];
    __shared__ double red[NT / 32];
    const int tid = threadIdx.x;
    const int seg = blockIdx.x;
    const int64_t c = OBS_MAPPED ? (int64_t)blockIdx.y : (int64_t)blockIdx.y * NT + tid;

    int64_t lo, hi;
    mv_segment(n_obs, seg, S, lo, hi);
    const int64_t len = hi - lo;
    const int n_tiles = (int)((len + TILE - 1) / TILE);

    MvConst<D> k;
    const bool live = c < C;
    k.load(lawc, C, live ? c : 0);
    double acc = 0.0;

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
        const uint32_t bytes = (uint32_t)(((int64_t)cnt * D + 1) >> 1) * 16u;  // padded buffer
        mbar_expect_tx(&bar[st], bytes);
        bulk_g2s(&tile[st][0], obs + (lo + off) * D, bytes, &bar[st]);
    };
    if (tid == 0)
        for (int t = 0; t < STAGES && t < n_tiles; ++t) issue(t);

    for (int t = 0; t < n_tiles; ++t) {
        const int st = t % STAGES;
        mbar_wait(&bar[st], (uint32_t)(t / STAGES) & 1u);
        const int64_t off = (int64_t)t * TILE;
        const int cnt = (int)((len - off) < (int64_t)TILE ? (len - off) : (int64_t)TILE);
        const double *xs = &tile[st][0];
        if (OBS_MAPPED) {
            for (int i = tid; i < cnt; i += NT) acc += k.quad(xs + i * D);
        } else {
#pragma unroll 2
            for (int i = 0; i < cnt; ++i) acc += k.quad(xs + i * D);  // broadcast reads
        }
        __syncthreads();
        if (tid == 0 && t + STAGES < n_tiles) issue(t + STAGES);
    }
    if (OBS_MAPPED) {
        double v = acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < NT / 32; ++w) s += red[w];
            partial[(int64_t)seg * C + c] = s;
        }
    } else if (live) {
        partial[(int64_t)seg * C + c] = acc;
    }
}

// d > 8: mu and W of a chain no longer fit the register file (152 doubles at d = 16); they live in
// shared memory, one column per thread ("chains" mapping, conflict-free: lane = column) or a single
// broadcast column ("obs" mapping); the difference vector stays in registers.
template <int D, int NT, int TILE, int STAGES, bool OBS_MAPPED>
__global__ void __launch_bounds__(NT)
sweep_gsnmv_smem_kernel(const double *__restrict__ obs, int64_t n_obs, const double *__restrict__ lawc,
                        int64_t C, double *__restrict__ partial, int S) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int K = D + D * (D + 1) / 2;
    constexpr int NC = OBS_MAPPED ? 1 : NT;
    double *tile = reinterpret_cast<double *>(smem_raw);                       // [STAGES][TILE * D]
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + STAGES * TILE * D);    // [STAGES]
    double *cs = reinterpret_cast<double *>(bar + STAGES + (STAGES & 1));      // [K][NC]
    double *red = cs + K * NC;                                                 // [NT / 32]
    const int tid = threadIdx.x;
    const int seg = blockIdx.x;
    const int64_t c = OBS_MAPPED ? (int64_t)blockIdx.y : (int64_t)blockIdx.y * NT + tid;
    const bool live = c < C;
    const int col = OBS_MAPPED ? 0 : tid;

    int64_t lo, hi;
    mv_segment(n_obs, seg, S, lo, hi);
    const int64_t len = hi - lo;
    const int n_tiles = (int)((len + TILE - 1) / TILE);
    if (OBS_MAPPED) {
        for (int j = tid; j < K; j += NT) cs[j] = lawc[(int64_t)j * C + c];
    } else {
        for (int j = 0; j < K; ++j) cs[j * NC + col] = lawc[(int64_t)j * C + (live ? c : 0)];
    }
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
        const uint32_t bytes = (uint32_t)(((int64_t)cnt * D + 1) >> 1) * 16u;  // padded buffer
        mbar_expect_tx(&bar[st], bytes);
        bulk_g2s(tile + st * TILE * D, obs + (lo + off) * D, bytes, &bar[st]);
    };
    if (tid == 0)
        for (int t = 0; t < STAGES && t < n_tiles; ++t) issue(t);
    auto quad = [&](const double *x) {
        double dd[D];
#pragma unroll
        for (int j = 0; j < D; ++j) dd[j] = x[j] - cs[j * NC + col];
        double q = 0.0;
        int w = D;
#pragma unroll
        for (int r = 0; r < D; ++r) {
            double z = cs[(w++) * NC + col] * dd[0];
#pragma unroll
            for (int k = 1; k <= r; ++k) z = fma(cs[(w++) * NC + col], dd[k], z);
            q = fma(z, z, q);
        }
        return q;
    };
    double acc = 0.0;
    for (int t = 0; t < n_tiles; ++t) {
        const int st = t % STAGES;
        mbar_wait(&bar[st], (uint32_t)(t / STAGES) & 1u);
        const int64_t off = (int64_t)t * TILE;
        const int cnt = (int)((len - off) < (int64_t)TILE ? (len - off) : (int64_t)TILE);
        const double *xs = tile + st * TILE * D;
        if (OBS_MAPPED) {
            for (int i = tid; i < cnt; i += NT) acc += quad(xs + i * D);
        } else {
            for (int i = 0; i < cnt; ++i) acc += quad(xs + i * D);  // broadcast reads of the observation
        }
        __syncthreads();
        if (tid == 0 && t + STAGES < n_tiles) issue(t + STAGES);
    }
    if (OBS_MAPPED) {
        double v = acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < NT / 32; ++w) s += red[w];
            partial[(int64_t)seg * C + c] = s;
        }
    } else if (live) {
        partial[(int64_t)seg * C + c] = acc;
    }
}

namespace {
constexpr int kNT = 128, kStages = 2;
constexpr int kBigNT = 64;     // chains per CTA of the shared-memory variant (d > 8)
template <int D> constexpr size_t kBigSmem(bool obs_mapped) {
    return (size_t)kStages * ((1024 / D) & ~1) * D * 8 + (kStages + (kStages & 1)) * 8 +
           (size_t)(D + D * (D + 1) / 2) * (obs_mapped ? 1 : kBigNT) * 8 + (kBigNT / 32) * 8;
}
template <int D>
void launch_big(const SweepPlan &pl, const double *obs, int64_t n_obs, const double *lawc, int64_t C,
                double *partial, cudaStream_t st) {
    dim3 grid(pl.S, pl.groups);
    constexpr int TILE = (1024 / D) & ~1;
    if (pl.variant == SWEEP_VARIANT_OBS) {
        auto k = sweep_gsnmv_smem_kernel<D, kBigNT, TILE, kStages, true>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigSmem<D>(true));
        k<<<grid, kBigNT, kBigSmem<D>(true), st>>>(obs, n_obs, lawc, C, partial, pl.S);
    } else {
        auto k = sweep_gsnmv_smem_kernel<D, kBigNT, TILE, kStages, false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigSmem<D>(false));
        k<<<grid, kBigNT, kBigSmem<D>(false), st>>>(obs, n_obs, lawc, C, partial, pl.S);
    }
}
// 8 KB of observations per stage, an even number of observations per tile
template <int D> constexpr int kTileObs() { return (1024 / D) & ~1; }

template <int D>
void launch_d(const SweepPlan &pl, const double *obs, int64_t n_obs, const double *lawc, int64_t C,
              double *partial, cudaStream_t st) {
    dim3 grid(pl.S, pl.groups);
    if (pl.variant == SWEEP_VARIANT_OBS)
        sweep_gsnmv_kernel<D, kNT, kTileObs<D>(), kStages, true><<<grid, kNT, 0, st>>>(obs, n_obs, lawc, C, partial, pl.S);
    else
        sweep_gsnmv_kernel<D, kNT, kTileObs<D>(), kStages, false><<<grid, kNT, 0, st>>>(obs, n_obs, lawc, C, partial, pl.S);
}
}  // namespace

SweepPlan plan_sweep_gsnmv(int d, int64_t C, int64_t n_obs, int force_variant, int num_sms) {
    SweepPlan pl{};
    pl.D = d;
    bool chains = C > 32;
    if (force_variant == SWEEP_VARIANT_CHAINS) chains = true;
    if (force_variant == SWEEP_VARIANT_OBS && C <= 65535) chains = false;
    pl.variant = chains ? SWEEP_VARIANT_CHAINS : SWEEP_VARIANT_OBS;
    pl.R = 1;
    const int nt = d > 8 ? kBigNT : kNT;
    pl.groups = chains ? (int)((C + nt - 1) / nt) : (int)C;
    int S = (num_sms * 4 + pl.groups - 1) / pl.groups;
    const int64_t max_S = ((n_obs + 1) / 2 + 127) / 128;  // >= 256 observations per segment
    if (S > max_S) S = (int)max_S;
    if (S < 1) S = 1;
    pl.S = S;
    pl.launches = 1;
    pl.name = chains ? "gsnmv_chains" : "gsnmv_obs";
    return pl;
}

void launch_sweep_gsnmv(const SweepPlan &pl, const double *obs, int64_t n_obs, const double *lawc,
                        int64_t C, double *partial, cudaStream_t st) {
    switch (pl.D) {
    case 2: launch_d<2>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 3: launch_d<3>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 4: launch_d<4>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 5: launch_d<5>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 6: launch_d<6>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 7: launch_d<7>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 8: launch_d<8>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 9: launch_big<9>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 10: launch_big<10>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 11: launch_big<11>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 12: launch_big<12>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 13: launch_big<13>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 14: launch_big<14>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 15: launch_big<15>(pl, obs, n_obs, lawc, C, partial, st); break;
    case 16: launch_big<16>(pl, obs, n_obs, lawc, C, partial, st); break;
    default: break;
    }
}

}  // namespace extmcmc
